set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r18_pytest.log; tail -5 gpurun_out/r18_pytest.log
MMREC_DEBUG=1 timeout 300 python tools/bench_score.py --paths fused > gpurun_out/r18_score.log 2>&1; tail -3 gpurun_out/r18_score.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r18_bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/r18_bench_n2.log; tail -2 gpurun_out/r18_bench_n2.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r18_bench_n1.log 2>&1; tail -1 gpurun_out/r18_bench_n1.log
