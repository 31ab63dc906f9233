set -x
mkdir -p gpurun_out
timeout -s KILL 100 python tools/timeline_fwd.py 2 8 512 64 2>&1 | tail -4
timeout -s KILL 100 python tools/timeline_bwd.py 2 8 512 64 2>&1 | tail -3
timeout -s KILL 100 python tools/timeline_fwd.py 4 16 1024 64 2>&1 | tail -3
timeout -s KILL 100 python tools/timeline_bwd.py 4 16 1024 64 2>&1 | tail -3
python tools/profile_one.py 2 8 512 64 5 > /dev/null 2>&1 && timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -c 40 --csv --log-file gpurun_out/r2c_ncu_small_A.csv python tools/profile_one.py 2 8 512 64 5 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c_ncu_small_A.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); 
for r in rows[1:]:
    print(r[ik][:60], r[im], r[iv])
PY
