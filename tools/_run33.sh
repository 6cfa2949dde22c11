set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/r33_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r33_pytest.log; tail -3 gpurun_out/r33_pytest.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r33_smoke.log 2>&1; tail -2 gpurun_out/r33_smoke.log
timeout 200 python bench.py > gpurun_out/r33_bench.log 2>&1; tail -1 gpurun_out/r33_bench.log | cut -c1-300
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fz_prep -c 12 --csv --log-file gpurun_out/r33_prep.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
