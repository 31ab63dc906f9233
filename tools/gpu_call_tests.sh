set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
for w in A B D; do timeout -s KILL 200 python bench.py --workload $w --steps 20 --no-cpu-baseline > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_$w.json')); print('$w', d['ms_per_step'], d['value'], d['kernel_ms'], d['verify']['ok'], d['e2e']['value'])"; done
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_C_20.json 2> gpurun_out/r2f_bench_C_20.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_C_20.json')); print('C', d['ms_per_step'], d['value'], d['kernel_ms'], d['verify']['ok'], d['e2e']['value'], d['e2e']['pcie'], d['clocks'])"
