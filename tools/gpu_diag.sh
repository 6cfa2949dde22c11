#!/bin/bash
# diagnostics of the scoring kernels: variants + ncu --set full captures
tag=${1:-dg}
mkdir -p gpurun_out
for dbg in 0 1 2 3; do
  MMREC_CF_DEBUG=$dbg timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cf_pass -c 12 --csv --log-file gpurun_out/${tag}_dbg$dbg.csv python tools/bench_score.py --paths auto --reps 2 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}_dbg$dbg.csv")) if len(r)>10 and r[0].isdigit()]
print("dbg=$dbg", [ (r[4].split('<')[1][:4], r[-1], r[-2]) for r in rows[-4:]])
PY
done
for k in cf_pass_kernel:2 cf_final_kernel:1 cf_thr_kernel:1; do
    name=${k%%:*}; skip=${k##*:}
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -o gpurun_out/${tag}_$name -f python tools/bench_score.py --paths auto --reps 2 > /dev/null 2>&1
done
ls -la gpurun_out | grep "${tag}_"
