set -x
mkdir -p gpurun_out
timeout -s KILL 120 python tools/timeline_bwd.py 2 32 4096 128 > gpurun_out/r2b_timeline_bwd2_ptdp.log 2>&1; tail -75 gpurun_out/r2b_timeline_bwd2_ptdp.log
python tools/profile_one.py 8 32 4096 128 2 > gpurun_out/r2b_ncu_plain.log 2>&1 && \
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k "regex:fa2_bwd2_kernel" -s 1 -c 1 -o gpurun_out/r2b_bwd2_ptdp python tools/profile_one.py 8 32 4096 128 2 > gpurun_out/r2b_ncu_run.log 2>&1
ls -la gpurun_out/r2b_bwd2_ptdp.ncu-rep
