set -x
V=cuda-flash-attention_b200/build/variants
timeout -s KILL 400 python tools/lib_variants.py $V/cur.so $V/fakeP.so $V/fakeDS.so $V/fakeBoth.so 2>&1 | tail -5
