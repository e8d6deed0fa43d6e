#!/bin/bash
# round 2, GPU call 17 (1 GPU): SAC loop timing + per-kernel launch list; train kernel launched twice back to back (instruction-cache experiment)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/sac_probe.py 16384 20 > gpurun_out/c17_sac.txt 2>&1
timeout 300 python tools/sac_probe.py 4096 20 >> gpurun_out/c17_sac.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/c17_sac_launches.csv python tools/sac_probe.py 16384 6 > gpurun_out/c17_sac_ncu.log 2>&1
UAVRL_TC_TRACE=1 UAVRL_TRAIN_TWICE=1 timeout 200 python tools/tc_trace.py 2>&1 | grep "train_trace" | tail -3 > gpurun_out/c17_trace.txt
grep sac_probe gpurun_out/c17_sac.txt; cat gpurun_out/c17_trace.txt
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/c17_sac_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
agg=collections.OrderedDict()
for r in rows[hdr+2:]:
    if len(r)>5: agg.setdefault(r[4].split('(')[0][:48],[]).append(float(r[-1]))
for k,v in agg.items(): print("%-50s n=%3d mean %.1f us"%(k,len(v),sum(v)/len(v)/1e3))
PY
