set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r21_pytest.log; tail -3 gpurun_out/r21_pytest.log
timeout 200 python bench.py > gpurun_out/r21_bench.log 2>&1; tail -1 gpurun_out/r21_bench.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r21_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r21_ncu_bench.log 2>&1
timeout 500 python tools/bench_spmm.py --users 1000000 --items 300000 --edges 16000000 --reps 5 > gpurun_out/r21_spmm_large.log 2>&1; tail -6 gpurun_out/r21_spmm_large.log
