#!/bin/bash
# Mid-round check on one B200 (under gpurun): the resident-tile GPU tests, stage timelines of the tile kernels and a
# --set full capture of one step's lane-kernel launches at 2^17 rows (quick_bench: 18 lane launches per step, 5 steps per size).
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tile_step" > $OUT/r02_gputests_tile.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_gputests_tile.log
tail -3 $OUT/r02_gputests_tile.log
{
TIMELINE=1 python tools/tile_prof.py ode 32 1 1048576
TIMELINE=1 python tools/tile_prof.py heat 32 1 262144
K=50 TIMELINE=1 python tools/tile_prof.py fredholm 32 1 4096
K=1024 python tools/tile_prof.py fredholm 32 1 16384
} 2>&1 | grep -v "^No batch\|ReLU sel" > $OUT/r02_tile_timeline.txt
tail -5 $OUT/r02_tile_timeline.txt
ncu --set full --clock-control none -k regex:lane_gemm -s 108 -c 18 -o $OUT/r02_lane_gemm_final python tools/quick_bench.py > $OUT/ncu_lane.log 2>&1
tail -3 $OUT/ncu_lane.log
