set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_backward.py -x -q -m gpu -k "two_gpus or sequence_split" > gpurun_out/r2_pytest_2gpu_b.log 2>&1; tail -25 gpurun_out/r2_pytest_2gpu_b.log
timeout -s KILL 600 python tools/seq_split_bench.py 2 > gpurun_out/r2_seq_split_n2.csv 2> gpurun_out/r2_seq_split_n2.err; cat gpurun_out/r2_seq_split_n2.csv; tail -5 gpurun_out/r2_seq_split_n2.err
