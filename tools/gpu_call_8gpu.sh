set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8; ls /sys/devices/system/node/ | grep node; nproc; free -g | head -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 1. multi-GPU parity tests
timeout -s KILL 300 python -m pytest tests/test_gpu_backward.py -x -q -m gpu -k "two_gpus or sequence_split" > gpurun_out/r2_pytest_8gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_8gpu.log
# 2. headline bench at 8 ranks (weak scaling + strong scaling legs)
timeout -s KILL 400 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; tail -3 gpurun_out/r2_bench_n8.err
timeout -s KILL 300 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
# 3. sequence-length sweep at 1 / 2 / 4 / 8 GPUs (BASELINE.json configs[4])
timeout -s KILL 200 python tools/sweep.py > gpurun_out/r2_sweep_n1.csv 2> gpurun_out/r2_sweep.err
for n in 2 4 8; do timeout -s KILL 200 $TR --nproc-per-node $n --master-port 2953$n tools/sweep.py > gpurun_out/r2_sweep_n$n.csv 2>> gpurun_out/r2_sweep.err; done
tail -3 gpurun_out/r2_sweep_n8.csv
# 4. sequence split vs slab split through the host API on 8 GPUs
timeout -s KILL 300 python tools/seq_split_bench.py 8 > gpurun_out/r2_seq_split_n8.csv 2> gpurun_out/r2_seq_split_n8.err; cat gpurun_out/r2_seq_split_n8.csv
# 5. CLI wall time at config C: serial (reference order) vs streamed, 1 and 8 GPUs
python - <<'PY'
import numpy as np, os
d = "/tmp/data/B8_H32_S4096_D128"; os.makedirs(d, exist_ok=True)
rng = np.random.default_rng(42)
for n in "QKV":
    rng.standard_normal((8, 32, 4096, 128), dtype=np.float32).tofile(f"{d}/{n}.bin")
PY
CLI=cuda-flash-attention_b200/FlashAttention
for i in 1 2; do
( /usr/bin/time -f "serial   gpus=1 wall %e s" env FA2_CLI_STREAM=0 $CLI fa2 forward_backward fp32 /tmp/data/B8_H32_S4096_D128 | grep -E "Kernel|Total" ) 2>&1
( /usr/bin/time -f "streamed gpus=1 wall %e s" $CLI fa2 forward_backward fp32 /tmp/data/B8_H32_S4096_D128 | grep -E "Kernel|Total" ) 2>&1
( /usr/bin/time -f "streamed gpus=8 wall %e s" $CLI fa2 forward_backward fp32 /tmp/data/B8_H32_S4096_D128 --gpus 8 | grep -E "Kernel|Total" ) 2>&1
done > gpurun_out/r2_cli_wall_configC.log 2>&1; cat gpurun_out/r2_cli_wall_configC.log
