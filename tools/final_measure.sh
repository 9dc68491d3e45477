#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun): GPU tests, one bench line per BASELINE config, the
# reference arm, the ncu launch list of the default bench command and --set full captures of the dominant kernels.
# Everything lands in gpurun_out/; tools/summarize_ncu.py turns the ncu outputs into profiles/r02_*.txt.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/r02_gputests_final.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_gputests_final.log
tail -2 $OUT/r02_gputests_final.log
python bench.py > $OUT/r02_bench_heat.json 2> $OUT/r02_bench_heat.err
for c in ode fhn fredholm; do python bench.py --config $c > $OUT/r02_bench_$c.json 2> $OUT/r02_bench_$c.err; done
python bench.py --config fhn --net dgm > $OUT/r02_bench_fhn_dgm.json 2> $OUT/r02_bench_fhn_dgm.err
python bench.py --config heat --hidden 32 --layers 1 > $OUT/r02_bench_heat_h32l1.json 2> $OUT/r02_bench_heat_h32l1.err
python bench.py --config heat --net mlp > $OUT/r02_bench_heat_mlp.json 2> $OUT/r02_bench_heat_mlp.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_bench_reference.json 2> $OUT/r02_bench_reference.err
python tools/small_batch.py 2>&1 | grep -v "^Iteration\|^Total\|^No batch\|ReLU sel" > $OUT/r02_small_batch.txt
# ncu: launch list of the same bench command (short run), then full captures (each program has already exited 0 above)
FL="--no-cpu-baseline --no-cuda-eager --no-driver-latency --steps 2 --warmup 3 --rows-per-gpu 131072"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/r02_launches_heat.csv python bench.py $FL > $OUT/ncu_heat.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches_ode.csv python bench.py --config ode $FL > $OUT/ncu_ode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_step -c 1 -o $OUT/r02_tile_ode_final python tools/tile_prof.py ode 32 1 1048576 1 > $OUT/ncu_tile_ode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_step -c 1 -o $OUT/r02_tile_heat32_final python tools/tile_prof.py heat 32 1 262144 1 > $OUT/ncu_tile_heat.log 2>&1
ncu --set full --clock-control none -k regex:lane_gemm -s 372 -c 6 -o $OUT/r02_lane_gemm_final python tools/quick_bench.py > $OUT/ncu_lane.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgrad_res -s 66 -c 2 -o $OUT/r02_dgrad_res_final python tools/quick_bench.py > $OUT/ncu_dgrad.log 2>&1
for f in heat ode fhn fredholm fhn_dgm heat_h32l1 heat_mlp; do python - <<PY
import json
try:
    j = json.load(open("$OUT/r02_bench_$f.json")); r = j["roofline"]
    print("$f", "%.4g rows/s %.3f ms e2e %.4g launches %d | %s frac %.3f | cpu %s | eager-cuda %s | drv %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["gpu_launches"], r.get("bound"), r.get("frac", 0), (j.get("cpu_baseline") or {}).get("value"), (j.get("cuda_eager_baseline") or {}).get("value"), {k: round(v) for k, v in (j.get("driver_latency") or {}).items() if k.endswith("iteration")}))
except Exception as e:
    print("$f", "unreadable", e)
PY
done
cat $OUT/r02_small_batch.txt
