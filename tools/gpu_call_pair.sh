set -x
mkdir -p gpurun_out
FA2_BWD_PAIR=1 timeout -s KILL 120 python tools/pair_check.py > gpurun_out/r2_pair_check.log 2>&1; echo "rc $?"; cat gpurun_out/r2_pair_check.log | tail -9
for pl in 0 1 2 0 1; do FA2_BWD_PAIR=1 FA2_BWD2_POLY=$pl timeout -s KILL 120 python tools/pair_check.py time 2>&1 | tail -2 | sed "s/^/poly $pl: /"; done
FA2_BWD_PAIR=1 FA2_BWD2_POLY=2 timeout -s KILL 120 python tools/pair_check.py 2>&1 | head -8
FA2_BWD_PAIR=1 timeout -s KILL 120 python tools/timeline_bwd.py 2 32 4096 128 > gpurun_out/r2_timeline_bwd2_v6.log 2>&1; tail -30 gpurun_out/r2_timeline_bwd2_v6.log
