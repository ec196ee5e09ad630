#!/bin/bash
# Round 2, call E: tight-issue fwd5 + replicated weight streams + wide dW stages: tests, phase profile, timings, bench; multi-seed PSNR twin.
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fwd_fp16.py tests/test_gpu_mlp_bwd.py tests/test_gpu_kernels.py tests/test_gpu_render.py tests/test_gpu_configs.py -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_$TAG.log
timeout 300 python scripts/prof_phases.py > $OUT/prof_phases_$TAG.txt 2>&1; echo "phases rc=$?"; cat $OUT/prof_phases_$TAG.txt | head -60
timeout 300 python scripts/time_fwd.py > $OUT/time_fwd_$TAG.txt 2>&1; cat $OUT/time_fwd_$TAG.txt
for combo in split:split split:fp16 fp16:fp16; do
  fp=${combo%%:*}; gp=${combo##*:}
  timeout 300 python bench.py --steps 20 --warmup 3 --quick --fwd-precision $fp --grad-precision $gp > $OUT/bench_${TAG}_${fp}_$gp.json 2> $OUT/bench_${TAG}_${fp}_$gp.err; echo "bench $combo rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${TAG}_${fp}_$gp.json"))
    print("$combo", round(d["ms_per_step"], 3), "ms (host launch", round(d["config"]["ms_per_step_host_launch"], 3), ") e2e", round(d["e2e"]["ms_per_step"], 3), d["clocks"]["sm_mhz"], {k.replace("cnerf_mlp_", ""): round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
except Exception as e:
    print("no result", e)
PY
done
for fp in split fp16; do
  timeout 300 python bench.py --mode render --steps 20 --warmup 3 --quick --fwd-precision $fp > $OUT/bench_render_${TAG}_$fp.json 2> $OUT/bench_render_${TAG}_$fp.err
  python -c "
import json; d = json.load(open('$OUT/bench_render_${TAG}_$fp.json')); print('render fwd=$fp', round(d['ms_per_step'], 3), 'ms', round(d['value']), 'rays/s e2e', round(d['e2e']['value']))"
done
timeout 1500 python scripts/psnr_twin.py --iters 600 --seeds 0 1 2 --repo-modes split:split split:fp16 fp16:fp16 --out $OUT/psnr_twin_$TAG.json > $OUT/psnr_twin_$TAG.log 2>&1; echo "psnr twin rc=$?"; tail -3 $OUT/psnr_twin_$TAG.log | cut -c1-1500
