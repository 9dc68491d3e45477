#!/bin/bash
# throughput (+ optional stage timelines) of the small-hidden-size configs on one B200
set -u
OUT=gpurun_out
TAG=${1:-x}
mkdir -p $OUT
{
TIMELINE=${TL:-} STAGES_PER_LINE=16 python tools/tile_prof.py ode 32 1 1048576
TIMELINE=${TL:-} STAGES_PER_LINE=16 python tools/tile_prof.py heat 32 1 262144
python tools/tile_prof.py heat 32 1 1048576
python tools/tile_prof.py fhn 32 2 262144
K=50 python tools/tile_prof.py fredholm 32 1 4096
python tools/tile_prof.py heat 64 3 4096
} 2>&1 | grep -v "^No batch\|ReLU sel" > $OUT/r02_tile_timeline_$TAG.txt
grep "rows/s\|totals" $OUT/r02_tile_timeline_$TAG.txt | cut -c1-250
