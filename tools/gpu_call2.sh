set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest2.log 2>&1; tail -25 gpurun_out/r2_pytest2.log
timeout -s KILL 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2_bench_C2.json 2> gpurun_out/r2_bench_C2.err; tail -3 gpurun_out/r2_bench_C2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_C2.json'))
print({k:d[k] for k in ('value','ms_per_step','kernel_ms','gpu_launches','verify')})
print(d['e2e'])
PY
for w in A D; do timeout -s KILL 200 python bench.py --workload $w --steps 20 --no-cpu-baseline > gpurun_out/r2_bench_${w}2.json 2> gpurun_out/r2_bench_${w}2.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${w}2.json'))
print({k:d[k] for k in ('value','ms_per_step','kernel_ms')}); print(d['e2e'])
PY
done
