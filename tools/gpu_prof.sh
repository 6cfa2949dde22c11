#!/bin/bash
# ncu --set full captures of named kernels while tools/bench_score.py runs:  bash tools/gpu_prof.sh <tag> kernel:skip [kernel:skip ...]
tag=$1; shift
mkdir -p gpurun_out
for k in "$@"; do
    name=${k%%:*}; skip=${k##*:}
    timeout 200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$name -s $skip -c 1 -o gpurun_out/${tag}_$name -f python tools/bench_score.py --paths auto --reps 2 > /dev/null 2>&1
done
ls -la gpurun_out | grep "${tag}_"
