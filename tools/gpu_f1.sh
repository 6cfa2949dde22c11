#!/bin/bash
# GPU-side check of the f1 / f4 / a5b work:  gpurun --timeout 900 -- 'bash tools/gpu_f1.sh [tag]'
tag=${1:-f1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|error|rc=" gpurun_out/${tag}_pytest.log | tail -5; grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_pytest.log | cut -c1-220 | head -30
timeout 200 python tools/bench_train.py > gpurun_out/${tag}_train.log 2>&1; tail -2 gpurun_out/${tag}_train.log | cut -c1-1000
MMREC_DGRAD_CLUSTER=1 timeout 200 python tools/bench_train.py --F 4096 > gpurun_out/${tag}_train_cluster.log 2>&1; tail -1 gpurun_out/${tag}_train_cluster.log | cut -c1-1000
timeout 500 python bench.py > gpurun_out/${tag}_bench.log 2>&1; tail -1 gpurun_out/${tag}_bench.log | cut -c1-200
# ncu --set full: launches of bench_train --reps 1 are 4 per timed op (3 warm-up + 1)
for k in "linear_dgrad_kernel:5:dgrad_adam" "adam_multi_kernel:1:adam_multi" "linear_wgrad_kernel:1:wgrad" "index_sum_rows_kernel:1:index_sum_rows"; do
    name=$(echo $k | cut -d: -f1); skip=$(echo $k | cut -d: -f2); out=$(echo $k | cut -d: -f3)
    timeout 150 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -o gpurun_out/${tag}_$out -f \
        python tools/bench_train.py --F 4096 --reps 1 > /dev/null 2>&1
done
ls gpurun_out | grep "^${tag}_"
