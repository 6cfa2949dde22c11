#!/bin/bash
# Round-end style check on one GPU:  gpurun --timeout 400 -- 'bash tools/gpu_final.sh [tag]'
# parity suite, smoke, and the N = 1 baselines of the sharded drivers (FULL=1 adds the default bench line)
tag=${1:-fin}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|error|rc=" gpurun_out/${tag}_pytest.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_pytest.log | cut -c1-220 | head -20
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log | cut -c1-300
if [ -n "$FULL" ]; then timeout 300 python bench.py > gpurun_out/${tag}_bench.log 2>&1; tail -1 gpurun_out/${tag}_bench.log | cut -c1-200; fi
timeout 90 python bench.py --gpus 1 --sharded --no-cpu-baseline > gpurun_out/${tag}_bench_n1_sharded.log 2>&1; tail -1 gpurun_out/${tag}_bench_n1_sharded.log | cut -c1-250
timeout 100 python bench.py --gpus 1 --workload xls --no-cpu-baseline > gpurun_out/${tag}_xls_n1.log 2>&1; tail -1 gpurun_out/${tag}_xls_n1.log | cut -c1-250
ls gpurun_out | grep "^${tag}_"
