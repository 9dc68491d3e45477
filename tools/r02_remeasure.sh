set -u
OUT=gpurun_out
python bench.py > $OUT/r02_bench_heat.json 2> $OUT/r02_bench_heat.err
python bench.py --config fhn > $OUT/r02_bench_fhn.json 2> $OUT/r02_bench_fhn.err
python bench.py --config fhn --net dgm > $OUT/r02_bench_fhn_dgm.json 2> $OUT/r02_bench_fhn_dgm.err
python bench.py --config heat --net mlp > $OUT/r02_bench_heat_mlp.json 2> $OUT/r02_bench_heat_mlp.err
FL="--no-cpu-baseline --no-cuda-eager --no-driver-latency --steps 2 --warmup 3 --rows-per-gpu 131072"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/r02_launches_heat.csv python bench.py $FL > $OUT/ncu_heat.log 2>&1
for f in heat fhn fhn_dgm heat_mlp; do python -c "
import json
j=json.load(open('$OUT/r02_bench_$f.json')); print('$f', j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'])"; done
