#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py -m gpu -q -x -k "philox or cuda_graph_driver" > $OUT/r02_gputests_philox.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_gputests_philox.log
tail -15 $OUT/r02_gputests_philox.log
for c in ode fredholm heat; do
python bench.py --config $c --no-cpu-baseline --no-cuda-eager --steps 3 --warmup 3 > $OUT/r02_bench_c_$c.json 2> $OUT/r02_bench_c_$c.err
python - <<PY
import json
try:
    j = json.load(open("$OUT/r02_bench_c_$c.json"))
    print("$c", "%.4g rows/s" % j["value"], {k: round(v, 1) for k, v in (j.get("driver_latency") or {}).items() if k.endswith("iteration")}, (j.get("driver_latency") or {}).get("error"))
except Exception as e:
    print("$c", "unreadable", e)
PY
done
