mkdir -p gpurun_out
for rep in 1 2 3; do
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('e2e', round(e['ms_per_step'],2), round(e['value'],1), round(e['pcie']['duplex_gbs_per_direction_per_gpu'],1), round(e['pcie']['e2e_frac_of_floor'],3), 'step', round(d['ms_per_step'],3))"
done
timeout -s KILL 300 python -m pytest tests -x -q -m gpu -k "host_api or chunk or cli" 2>&1 | tail -2
