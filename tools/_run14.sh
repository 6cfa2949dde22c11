set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r14_pytest.log; tail -5 gpurun_out/r14_pytest.log
timeout 300 python tools/bench_score.py --paths fused,tc > gpurun_out/r14_score.log 2>&1; tail -4 gpurun_out/r14_score.log
timeout 300 python bench.py > gpurun_out/r14_bench.log 2>&1; tail -2 gpurun_out/r14_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r14_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r14_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:project_tc_kernel -c 1 -o gpurun_out/r14_project_tc -f python tools/bench_score.py --paths fused --reps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:project_tc_kernel -c 1 -o gpurun_out/r14_project_tc -f python bench.py --steps 1 --warmup 1 > gpurun_out/r14_ncu_pj.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_fused_kernel -s 2 -c 1 -o gpurun_out/r14_score_fused -f python bench.py --steps 1 --warmup 1 > gpurun_out/r14_ncu_fz.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_vec_kernel -s 4 -c 1 -o gpurun_out/r14_spmm -f python bench.py --steps 1 --warmup 1 > gpurun_out/r14_ncu_sp.log 2>&1
ls -la gpurun_out/
