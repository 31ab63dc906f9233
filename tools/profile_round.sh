set -x
timeout -s KILL 400 python bench.py > gpurun_out/bench_v3_n1.json 2> gpurun_out/bench_v3_n1.err
timeout -s KILL 300 python bench.py --impl reference > gpurun_out/bench_v3_reference.json 2>> gpurun_out/bench_v3_n1.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_bench_steps2_v3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_run.log 2>&1
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:fa2_ -c 3 -o gpurun_out/full_v3 python tools/profile_one.py 8 32 4096 128 1 > gpurun_out/ncu_full_run.log 2>&1
ncu -i gpurun_out/full_v3.ncu-rep --page raw --csv > gpurun_out/ncu_full_fwd_bwd_raw_v3.csv 2>/dev/null
ncu -i gpurun_out/full_v3.ncu-rep --page details > gpurun_out/ncu_full_fwd_bwd_details_v3.txt 2>/dev/null
timeout -s KILL 100 python tools/timeline_fwd.py 8 32 4096 128 > gpurun_out/timeline_fwd_configC_v3.log 2>&1
timeout -s KILL 100 python tools/timeline_bwd.py 8 32 4096 128 > gpurun_out/timeline_bwd_configC_v3.log 2>&1
ls -la gpurun_out/
