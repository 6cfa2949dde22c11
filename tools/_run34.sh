set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/r34_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r34_pytest.log; tail -3 gpurun_out/r34_pytest.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r34_bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/r34_bench_n2.log; tail -2 gpurun_out/r34_bench_n2.log | cut -c1-300
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fz_prep -c 12 --csv --log-file gpurun_out/r34_prep.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/r34_bench.log 2>&1; tail -1 gpurun_out/r34_bench.log | cut -c1-200
