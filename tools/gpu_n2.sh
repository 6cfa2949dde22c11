#!/bin/bash
# N = 2 check of the item-sharded path:  gpurun --gpus 2 --timeout 500 -- 'bash tools/gpu_n2.sh [tag]'
# (the N = 1 baselines of the same drivers -- `bench.py --gpus 1 --sharded`, `--workload xls` -- run in tools/gpu_final.sh on one GPU)
tag=${1:-n2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 python -m pytest tests/test_gpu_sharded.py -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log | cut -c1-300
timeout 150 $TR --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n2.log 2>&1; grep '^{' gpurun_out/${tag}_bench_n2.log | tail -1 | cut -c1-250
timeout 200 $TR --master-port 29542 bench.py --gpus 2 --workload xls --steps 20 --warmup 5 > gpurun_out/${tag}_xls_n2.log 2>&1; grep '^{' gpurun_out/${tag}_xls_n2.log | tail -1 | cut -c1-250
ls gpurun_out | grep "^${tag}_"
