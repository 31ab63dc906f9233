set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1; tail -3 gpurun_out/r2_final_pytest.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_final_bench.json')); print('C', d['ms_per_step'], d['value'], d['gpu_launches'], d['verify']['ok'], d['e2e']['value'], d['roofline']['frac'])"
timeout -s KILL 200 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
