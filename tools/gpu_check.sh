#!/bin/bash
# The GPU-side check of a change, as one `gpurun` call:   gpurun --timeout 1300 -- 'bash tools/gpu_check.sh [tag]'
# parity tests, smoke, bench line, launch list and one `ncu --set full` capture per hot kernel -> gpurun_out/<tag>_*.
# Every step has its own timeout (the first python start on a fresh box can take a minute or two by itself).
tag=${1:-chk}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.log 2>&1; tail -1 gpurun_out/${tag}_bench.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_ncu_bench.log 2>&1
for k in ${KERNELS:-}; do
    name=${k%%:*}; skip=${k##*:}
    timeout 120 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -o gpurun_out/${tag}_$name -f \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
done
ls gpurun_out | grep "^${tag}_"
# then, here:  python tools/summarize_ncu.py launches gpurun_out/<tag>_launches.csv profiles/rNN_launches_bench.md "<title>"
#              python tools/summarize_ncu.py full gpurun_out/<tag>_<kernel>.ncu-rep profiles/rNN_<kernel>.md <kernel>
