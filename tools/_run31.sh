set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/r31_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r31_pytest.log; tail -3 gpurun_out/r31_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r31_smoke.log 2>&1; tail -2 gpurun_out/r31_smoke.log
timeout 200 python bench.py > gpurun_out/r31_bench.log 2>&1; tail -1 gpurun_out/r31_bench.log | cut -c1-300
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r31_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r31_ncu_bench.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:spmm_vec_kernel -s 4 -c 1 -o gpurun_out/r31_spmm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:project_tc_kernel -c 1 -o gpurun_out/r31_project_tc -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:score_fused_kernel -s 2 -c 1 -o gpurun_out/r31_score_fused -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:fused_select_kernel -s 2 -c 1 -o gpurun_out/r31_select -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:fz_prep_kernel -s 2 -c 1 -o gpurun_out/r31_prep -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls gpurun_out | grep r31
