#!/usr/bin/env bash
# round 2, GPU call 32: ablation of the tc5p pipeline (timing only, wrong results): no epilogue stores / one tap's MMAs / both
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 0 1 2 3; do
IMVS_TUNE_TC5P_DBG=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c32_launches_dbg$v.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c32_ncu_$v.log 2>&1
done
python - <<'PY'
import csv
res={}
for v in range(4):
    rows=[r for r in csv.reader(open(f'gpurun_out/r2c32_launches_dbg{v}.csv')) if len(r)>10]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    res[v]=[(r[ki][:70], float(r[vi])/1000) for r in rows[1:21]]
for i in range(20):
    print(res[0][i][0], ' | '.join(f"{res[v][i][1]:.1f}" for v in range(4)))
PY
