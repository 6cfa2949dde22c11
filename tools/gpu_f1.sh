#!/bin/bash
# GPU-side check of the f1 / f4 / a5b work:  gpurun --timeout 900 -- 'bash tools/gpu_f1.sh [tag]'
tag=${1:-f1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|error|rc=" gpurun_out/${tag}_pytest.log | tail -5; grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_pytest.log | cut -c1-220 | head -30
timeout 200 python tools/bench_train.py > gpurun_out/${tag}_train.log 2>&1; tail -3 gpurun_out/${tag}_train.log | cut -c1-900
timeout 500 python bench.py > gpurun_out/${tag}_bench.log 2>&1; tail -1 gpurun_out/${tag}_bench.log | cut -c1-200
for k in "dgrad_kernel.64..512..1:dgrad_adam" "linear_wgrad_kernel:wgrad" "mgcn_fuse_kernel:mgcn_fuse"; do
    name=${k%%:*}; out=${k##*:}
    if [ "$out" = "mgcn_fuse" ]; then cmd="python -m pytest tests/test_gpu_fuse.py -q -m gpu -k 63000"; else cmd="python tools/bench_train.py --F 4096 --reps 1"; fi
    timeout 150 ncu --set full --clock-control none --import-source on -k regex:$name -s 1 -c 1 -o gpurun_out/${tag}_$out -f $cmd > /dev/null 2>&1
done
ls gpurun_out | grep "^${tag}_"
