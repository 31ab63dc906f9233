set -x
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_C.json 2> gpurun_out/r2g_bench_C.err; tail -2 gpurun_out/r2g_bench_C.err; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_C.json')); print('C', d['ms_per_step'], d['value'], d['gpu_launches'], d['kernel_ms'], d['verify']['ok'], d['e2e']['value'])"
timeout -s KILL 200 python bench.py --workload A --steps 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('A', d['ms_per_step'], d['gpu_launches'])"
timeout -s KILL 300 python -m pytest tests -x -q -m gpu -k "cabi or symbol or exports or forward_vs_golden" 2>&1 | tail -2
