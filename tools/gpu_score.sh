#!/bin/bash
# GPU-side check of the scoring path only:  gpurun --timeout 1500 -- 'bash tools/gpu_score.sh [tag]'
tag=${1:-sc}
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "score or fused or topk or full_size" > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log; tail -12 gpurun_out/${tag}_pytest.log | cut -c1-250
timeout 500 python -m pytest tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/${tag}_configs.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_configs.log; tail -12 gpurun_out/${tag}_configs.log | cut -c1-250
MMREC_DEBUG=1 timeout 200 python tools/bench_score.py --paths auto,tc > gpurun_out/${tag}_score.log 2>&1; tail -4 gpurun_out/${tag}_score.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv python tools/bench_score.py --paths auto --reps 2 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/${tag}_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][:60]; v=float(r[-1].replace(',','')); v*= {'ns':1.0,'us':1e3,'ms':1e6}.get(r[-2],1.0); agg.setdefault(name,[]).append(v)
for k,v in agg.items(): print(f"{k:60s} n={len(v):3d} med={sorted(v)[len(v)//2]/1e3:8.2f} us")
PY
