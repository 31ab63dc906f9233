set -x
mkdir -p gpurun_out
nvidia-smi -L; ls /sys/devices/system/node/ | head; nproc; nvidia-smi topo -m 2>/dev/null | head -20
timeout -s KILL 400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest1.log 2>&1; tail -15 gpurun_out/r2_pytest1.log
timeout -s KILL 300 python bench.py --steps 20 > gpurun_out/r2_bench_C.json 2> gpurun_out/r2_bench_C.err; tail -3 gpurun_out/r2_bench_C.err; cat gpurun_out/r2_bench_C.json
for w in A B D; do timeout -s KILL 200 python bench.py --workload $w --steps 20 --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; tail -2 gpurun_out/r2_bench_$w.err; cat gpurun_out/r2_bench_$w.json; done
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -3 gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_ref.json
