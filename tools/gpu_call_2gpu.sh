set -x
mkdir -p gpurun_out
nvidia-smi -L; nvidia-smi topo -m | head -8
timeout -s KILL 900 python -m pytest tests -x -q -m gpu -k "two_gpus or sequence_split or host_api" > gpurun_out/r2_pytest_2gpu.log 2>&1; tail -25 gpurun_out/r2_pytest_2gpu.log
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -5 gpurun_out/r2_bench_n2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','verify')})
print(json.dumps(d.get('strong_scaling'), indent=1)); print(d['e2e'])
PY
timeout -s KILL 300 python tools/check_multi_gpu_cli.py 2 > gpurun_out/r2_cli_multi_gpu.log 2>&1; tail -5 gpurun_out/r2_cli_multi_gpu.log
