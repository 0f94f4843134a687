#!/bin/bash
# ncu captures for one round (run on the GPU box through gpurun; see B200_PROFILING.md).
#   bash profiles/run_ncu.sh r1 [steps]   -> gpurun_out/launches_r1.csv, gpurun_out/prof_*_r1.ncu-rep
# steps: any of "l" (launch list), "c" (Chamfer dense kernel, full set), "s" (stage kernels, full set); default all
set -u
R=${1:-r1}
STEPS=${2:-lcs}
mkdir -p gpurun_out
# 1. every launch of the default bench command with its device time (cold-cache, serialised: compare shares)
if [[ $STEPS == *l* ]]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/bench_under_ncu_${R}.log 2>&1
fi
# 2. the dominant kernel (nn_pair_kernel since round 2b; nn_kernel with DUSTY_CHAMFER_PRUNE=0), full set, at the bench's own size
if [[ $STEPS == *c* ]]; then
ncu --set full --clock-control none --import-source on -k regex:"nn_pair|nn_kernel" -s 1 -c 1 -f -o gpurun_out/prof_chamfer_${R} \
    python bench.py --steps 1 --warmup 1 --skip-extras > gpurun_out/prof_chamfer_${R}.log 2>&1
fi
# 3. head + projection, real-scan preprocess, FPS and the merged-origin Chamfer kernel (last launch of each)
if [[ $STEPS == *s* ]]; then
ncu --set full --clock-control none --import-source on -k regex:"head_project|fps_|scan_preprocess|nn_kernel|nn_walk|nn_pair|prep_sort" -s 6 -c 10 -f \
    -o gpurun_out/prof_stages_${R} python profiles/stage_driver.py > gpurun_out/prof_stages_${R}.log 2>&1
fi
