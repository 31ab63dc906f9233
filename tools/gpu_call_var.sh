set -x
mkdir -p gpurun_out
V=cuda-flash-attention_b200/build/variants
timeout -s KILL 400 python tools/lib_variants.py $V/base.so $V/split.so $V/all3.so $V/ptdp.so $V/ptfix.so $V/ptnosplit.so $V/splitonly.so $V/base_b.so $V/split_b.so $V/all3_b.so $V/ptdp_b.so 2>&1 | tee gpurun_out/r2b_variants2.log
nvidia-smi --query-gpu=power.limit,power.default_limit,power.max_limit,clocks.max.sm --format=csv
