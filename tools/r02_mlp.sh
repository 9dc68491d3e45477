#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $OUT/r02_gputests_kernels_mlp.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_gputests_kernels_mlp.log
tail -3 $OUT/r02_gputests_kernels_mlp.log
python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "fhn or mlp" > $OUT/r02_gputests_large_mlp.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_gputests_large_mlp.log
tail -3 $OUT/r02_gputests_large_mlp.log
FL="--no-cpu-baseline --no-cuda-eager --no-driver-latency"
python bench.py --config fhn $FL > $OUT/r02_bench_b_fhn.json 2> $OUT/r02_bench_b_fhn.err
python bench.py --config heat --net mlp $FL > $OUT/r02_bench_b_heat_mlp.json 2> $OUT/r02_bench_b_heat_mlp.err
for f in fhn heat_mlp; do python - <<PY
import json
try:
    j = json.load(open("$OUT/r02_bench_b_$f.json")); r = j["roofline"]
    print("$f", "%.4g rows/s %.3f ms launches %d" % (j["value"], j["ms_per_step"], j["gpu_launches"]), {k: (round(v["ms_per_step"], 2), v["launches_per_step"]) for k, v in r["kernel_classes"].items()})
except Exception as e:
    print("$f", "unreadable", e)
PY
done
