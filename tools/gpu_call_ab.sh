set -x
mkdir -p gpurun_out
for rep in 1 2; do
for pair in 0 1; do
FA2_BWD_PAIR=$pair timeout -s KILL 300 python bench.py --steps 50 --no-cpu-baseline --no-verify --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pair=$pair steps=50', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done; done
