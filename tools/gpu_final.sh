#!/bin/bash
# Round-end style check on one GPU:  gpurun --timeout 600 -- 'bash tools/gpu_final.sh [tag]'
tag=${1:-fin}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|error|rc=" gpurun_out/${tag}_pytest.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_pytest.log | cut -c1-220 | head -20
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/${tag}_bench.log 2>&1; tail -1 gpurun_out/${tag}_bench.log | cut -c1-200
timeout 120 python bench.py --gpus 1 --sharded --no-cpu-baseline > gpurun_out/${tag}_bench_n1_sharded.log 2>&1; tail -1 gpurun_out/${tag}_bench_n1_sharded.log | cut -c1-250
timeout 150 python bench.py --gpus 1 --workload xls --no-cpu-baseline > gpurun_out/${tag}_xls_n1.log 2>&1; tail -1 gpurun_out/${tag}_xls_n1.log | cut -c1-250
ls gpurun_out | grep "^${tag}_"
