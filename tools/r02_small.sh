#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $OUT/r02_gputests_kernels_small.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_gputests_kernels_small.log
tail -3 $OUT/r02_gputests_kernels_small.log
python -m pytest tests/test_gpu_modules.py -m gpu -q -x -k "philox or cuda_graph_driver or fredholm" > $OUT/r02_gputests_mod_small.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_gputests_mod_small.log
tail -3 $OUT/r02_gputests_mod_small.log
python tools/small_batch.py 2>&1 | grep -v "^Iteration\|^Total\|^No batch\|ReLU sel" > $OUT/r02_small_batch_b.txt
cat $OUT/r02_small_batch_b.txt
for c in ode fredholm; do
python bench.py --config $c --no-cpu-baseline --no-cuda-eager --steps 3 --warmup 3 > $OUT/r02_bench_d_$c.json 2> $OUT/r02_bench_d_$c.err
python - <<PY
import json
try:
    j = json.load(open("$OUT/r02_bench_d_$c.json"))
    print("$c", "%.4g rows/s" % j["value"], {k: round(v, 1) for k, v in (j.get("driver_latency") or {}).items() if k.endswith("iteration")}, (j.get("driver_latency") or {}).get("error"))
except Exception as e:
    print("$c", "unreadable", e)
PY
done
