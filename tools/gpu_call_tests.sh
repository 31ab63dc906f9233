set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; tail -6 gpurun_out/r2c_pytest.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_C.json 2> gpurun_out/r2c_bench_C.err; tail -3 gpurun_out/r2c_bench_C.err; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_C.json')); print(d['ms_per_step'], d['value'], d['kernel_ms'], d['clocks'], d['e2e']['value'], d['verify']['ok'])"
