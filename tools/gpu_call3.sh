set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest3.log 2>&1; tail -8 gpurun_out/r2_pytest3.log
timeout -s KILL 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2_bench_C3.json 2> gpurun_out/r2_bench_C3.err; tail -3 gpurun_out/r2_bench_C3.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_C3.json'))
print({k:d[k] for k in ('value','ms_per_step','kernel_ms','gpu_launches','clocks')}); print(d['verify']['max_abs_err'], d['verify']['ok']); print(d['roofline']['frac'], d['roofline']['frac_of_nominal_2250'])
print(d['e2e'])
PY
for w in D B A; do timeout -s KILL 200 python bench.py --workload $w --steps 20 --no-cpu-baseline > gpurun_out/r2_bench_${w}3.json 2> gpurun_out/r2_bench_${w}3.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${w}3.json'))
print({k:d[k] for k in ('value','ms_per_step','kernel_ms')}, d['verify']['ok'])
PY
done
