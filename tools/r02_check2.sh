#!/bin/bash
# Resident-tile step after a change: kernel-level GPU tests, then throughput + stage timelines of the small-hidden-size configs.
set -u
OUT=gpurun_out
TAG=${1:-x}
mkdir -p $OUT
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $OUT/r02_gputests_kernels_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_gputests_kernels_$TAG.log
tail -3 $OUT/r02_gputests_kernels_$TAG.log
{
TIMELINE=1 STAGES_PER_LINE=16 python tools/tile_prof.py ode 32 1 1048576
TIMELINE=1 STAGES_PER_LINE=16 python tools/tile_prof.py heat 32 1 262144
python tools/tile_prof.py heat 32 1 1048576
python tools/tile_prof.py fhn 32 2 262144
K=50 python tools/tile_prof.py fredholm 32 1 4096
python tools/tile_prof.py heat 64 3 4096
} 2>&1 | grep -v "^No batch\|ReLU sel" > $OUT/r02_tile_timeline_$TAG.txt
grep "rows/s\|totals" $OUT/r02_tile_timeline_$TAG.txt | cut -c1-250
