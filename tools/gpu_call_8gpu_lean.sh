set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 300 python -m pytest tests/test_gpu_backward.py -x -q -m gpu -k "two_gpus or sequence_split" > gpurun_out/r2_pytest_8gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_8gpu.log
timeout -s KILL 400 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; tail -3 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['verify']['ok'])
ss=d['strong_scaling']; print(ss['device']['T1_ms'], ss['device']['TN_ms'], ss['device']['efficiency']); print(ss['host_api']['kernel_ms'], ss['host_api']['wall_ms'], ss['host_api']['equal_to_1gpu']); print(d['e2e'])
PY
