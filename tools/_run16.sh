set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r16_pytest.log; tail -5 gpurun_out/r16_pytest.log
MMREC_DEBUG=1 timeout 300 python tools/bench_score.py --paths fused,tc > gpurun_out/r16_score.log 2>&1; tail -5 gpurun_out/r16_score.log
timeout 300 python tools/bench_project.py > gpurun_out/r16_project.log 2>&1; tail -5 gpurun_out/r16_project.log
timeout 300 python bench.py > gpurun_out/r16_bench.log 2>&1; tail -2 gpurun_out/r16_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r16_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r16_ncu_bench.log 2>&1
