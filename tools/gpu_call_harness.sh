set -x
mkdir -p gpurun_out/harness
cd cuda-flash-attention_b200
timeout -s KILL 400 python test_flash_attention2.py --mode both --experiment --save-results --output-dir ../gpurun_out/harness --no-stop-on-failure > ../gpurun_out/harness/experiment.log 2>&1; tail -15 ../gpurun_out/harness/experiment.log
timeout -s KILL 300 python test_flash_attention2.py --mode both --seqlen-experiment --save-results --output-dir ../gpurun_out/harness --no-stop-on-failure > ../gpurun_out/harness/seqlen.log 2>&1; tail -5 ../gpurun_out/harness/seqlen.log
cd ..; ls -la gpurun_out/harness
timeout -s KILL 200 python tools/sweep.py > gpurun_out/r2f_sweep_n1.csv 2> gpurun_out/r2f_sweep.err; tail -3 gpurun_out/r2f_sweep_n1.csv
