set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/r30_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r30_pytest.log; tail -3 gpurun_out/r30_pytest.log
MMREC_DEBUG=1 timeout 90 python tools/bench_score.py --paths fused,tc > gpurun_out/r30_score.log 2>&1; tail -3 gpurun_out/r30_score.log
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/r30_bench.log 2>&1; tail -1 gpurun_out/r30_bench.log | cut -c1-300
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r30_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r30_ncu_bench.log 2>&1
