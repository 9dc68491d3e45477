#!/bin/bash
# BASELINE configs[4] on hardware: heat + dgm_net.DGM(2,1,128,3) data-parallel over N GPUs of one box --
# weak scaling at 2^21 rows per GPU (2^24 global at N = 8) and strong scaling at a fixed global 2^24 rows -- and the two
# independent-trial sweeps, one trial per GPU.   usage: tools/run_config5.sh N [studies]     (outputs: gpurun_out/r02_c5_*)
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
if [ "$N" -gt 1 ]; then
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
else
  RUN="python"
fi
FLAGS="--no-cpu-baseline --no-cuda-eager --no-driver-latency"
if [ "${2:-}" != "ray" ]; then
$RUN bench.py --gpus $N --rows-per-gpu 2097152 $FLAGS > $OUT/r02_c5_weak_n$N.json 2> $OUT/r02_c5_weak_n$N.err
$RUN bench.py --gpus $N --scaling strong --global-rows 16777216 $FLAGS > $OUT/r02_c5_strong_n$N.json 2> $OUT/r02_c5_strong_n$N.err
fi
if [ "${2:-}" = "ray" ]; then
  $RUN -m differential_equations_dnn_b200.optimize_heat_ray --num-samples 10 --out $OUT/r02_c5_ray_n$N.json > /dev/null 2> $OUT/r02_c5_ray_n$N.err
  python -c "import json; j=json.load(open('$OUT/r02_c5_ray_n$N.json')); print('ray wall_s %.1f' % j['wall_s'], [round(t['loss'],8) for t in j['trials']])"
  exit 0
fi
if [ "${2:-}" = "studies" ]; then
  if [ "$N" -gt 1 ]; then M="-m"; else M="-m"; fi
  $RUN -m differential_equations_dnn_b200.optimize_heat_ray --num-samples 10 --out $OUT/r02_c5_ray_n$N.json > /dev/null 2> $OUT/r02_c5_ray_n$N.err
  $RUN -m differential_equations_dnn_b200.batchsize_effect_heat --n-iters 15000 --n-runs 5 --fix-batch-size --fresh-net --out $OUT/r02_c5_bs_n$N.json > /dev/null 2> $OUT/r02_c5_bs_n$N.err
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/r02_c5_*_n$N.json")):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    if "value" in j:
        print(f, "rows/s %.4g ms/step %.2f n_gpus %d global_rows %s" % (j["value"], j["ms_per_step"], j["n_gpus"], j["config"]["global_rows"]))
    else:
        print(f, "wall_s %.1f world %d best/first" % (j["wall_s"], j["world_size"]), j.get("best") or [(t["config"], round(t["loss"], 6)) for t in j["trials"][:3]])
PY
