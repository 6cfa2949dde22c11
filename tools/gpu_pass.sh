#!/bin/bash
# pass-kernel timing at a given batch for MMREC_CF_DEBUG variants:  bash tools/gpu_pass.sh "0 4" 20000
for dbg in $1; do
  MMREC_CF_DEBUG=$dbg timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:cf_ -c 14 --csv --log-file gpurun_out/pass_dbg$dbg.csv python tools/bench_score.py --paths auto --reps 1 --batch ${2:-20000} > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/pass_dbg$dbg.csv")) if len(r)>10 and r[0].isdigit()]
print("dbg=$dbg", [(r[4].split("(")[0][-24:], r[-1]) for r in rows[-8:]])
PY
done
