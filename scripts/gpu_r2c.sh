#!/bin/bash
# Round 2, call C: fp16 forward (mlp_fwd5) parity + timing, regime test, GPU-eager denominator.
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fwd_fp16.py -x -q -s > $OUT/pytest_fwd5_$TAG.log 2>&1; echo "pytest fwd5 rc=$?"; tail -25 $OUT/pytest_fwd5_$TAG.log
timeout 300 python scripts/time_fwd.py > $OUT/time_fwd_$TAG.txt 2>&1; echo "time_fwd rc=$?"; cat $OUT/time_fwd_$TAG.txt | tail -12
timeout 600 python -m pytest tests/test_gpu_dropin_train.py -x -q -k default_cuda > $OUT/pytest_regime_$TAG.log 2>&1; tail -3 $OUT/pytest_regime_$TAG.log
timeout 120 python oracle/twin.py eager --mode train --steps 5 --warmup 2 > $OUT/eager_train_$TAG.json 2> $OUT/eager_train_$TAG.err; echo "eager train rc=$?"; tail -1 $OUT/eager_train_$TAG.json
timeout 120 python oracle/twin.py eager --mode render --steps 5 --warmup 2 > $OUT/eager_render_$TAG.json 2> $OUT/eager_render_$TAG.err; echo "eager render rc=$?"; tail -1 $OUT/eager_render_$TAG.json
for fp in split fp16; do for gp in split fp16; do
  if [ $fp = fp16 ] && [ $gp = split ]; then continue; fi
  CNERF_FWD_PRECISION=$fp timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --grad-precision $gp > $OUT/bench_train_${TAG}_${fp}_$gp.json 2> $OUT/bench_train_${TAG}_${fp}_$gp.err; echo "bench fwd=$fp grad=$gp rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_train_${TAG}_${fp}_$gp.json"))
    print("fwd=$fp grad=$gp", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["ms_per_step"], 3), d["clocks"]["sm_mhz"], {k.replace("cnerf_mlp_", ""): round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
except Exception as e:
    print("no result", e)
PY
done; done
for fp in split fp16; do
  CNERF_FWD_PRECISION=$fp timeout 300 python bench.py --mode render --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_render_${TAG}_$fp.json 2> $OUT/bench_render_${TAG}_$fp.err
  python -c "
import json; d = json.load(open('$OUT/bench_render_${TAG}_$fp.json')); print('render fwd=$fp', round(d['ms_per_step'], 3), 'ms', round(d['value']), 'rays/s e2e', round(d['e2e']['value']))"
done
