# Round-2 evidence run (one B200): bench lines, ncu launch list, ncu full capture of the two tcgen05 kernels, CLI wall time.
set -x
R=gpurun_out
mkdir -p $R
timeout -s KILL 400 python bench.py > $R/r2_bench_final_n1.json 2> $R/r2_bench_final_n1.err
timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 1 > $R/r2_bench_final_reference.json 2>> $R/r2_bench_final_n1.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $R/r2_ncu_plain.log 2>&1 && \
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $R/r2_ncu_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $R/r2_ncu_launch_run.log 2>&1
python tools/profile_one.py 8 32 4096 128 2 > $R/r2_ncu_plain2.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k "regex:fa2_(fwd|bwd2)_kernel" -s 2 -c 2 -o $R/r2_full python tools/profile_one.py 8 32 4096 128 2 > $R/r2_ncu_full_run.log 2>&1
ncu -i $R/r2_full.ncu-rep --page raw --csv > $R/r2_ncu_full_fwd_bwd_raw.csv 2>/dev/null
ncu -i $R/r2_full.ncu-rep --page details > $R/r2_ncu_full_fwd_bwd_details.txt 2>/dev/null
ls -la $R/r2_full.ncu-rep
# CLI wall time at config C: the reference's serial order vs the streamed pipeline
python - <<'PY'
import numpy as np, os
d = "/tmp/data/B8_H32_S4096_D128"; os.makedirs(d, exist_ok=True)
rng = np.random.default_rng(42)
for n in "QKV":
    rng.standard_normal((8, 32, 4096, 128), dtype=np.float32).tofile(f"{d}/{n}.bin")
PY
CLI=cuda-flash-attention_b200/FlashAttention
for i in 1 2; do
  echo "== serial (FA2_CLI_STREAM=0)"; { time FA2_CLI_STREAM=0 $CLI fa2 forward_backward fp32 /tmp/data/B8_H32_S4096_D128 | grep -E "Kernel|Total" ; } 2>&1
  echo "== streamed"; { time $CLI fa2 forward_backward fp32 /tmp/data/B8_H32_S4096_D128 | grep -E "Kernel|Total" ; } 2>&1
done > $R/r2_cli_wall_configC.log 2>&1
cat $R/r2_cli_wall_configC.log
