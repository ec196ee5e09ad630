#!/bin/bash
# Round 2, call F: chain5 + single weight-stream copy: tests, timings, bench; sanitizer; PSNR twin with 8 training views.
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fwd_fp16.py tests/test_gpu_mlp_bwd.py tests/test_gpu_render.py -x -q -s > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|cos|Error|error" $OUT/pytest_$TAG.log | tail -12
for combo in split:fp16 fp16:fp16; do
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
CNERF_CHAIN1=single timeout 300 python bench.py --steps 20 --warmup 3 --quick --fwd-precision fp16 --grad-precision fp16 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('single-tile chain:', round(d['ms_per_step'],3), {k.replace('cnerf_mlp_',''): round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize.py > $OUT/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/sanitizer_memcheck_$TAG.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize.py > $OUT/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -4 $OUT/sanitizer_racecheck_$TAG.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/sanitize.py > $OUT/sanitizer_synccheck_$TAG.log 2>&1; echo "synccheck rc=$?"; tail -4 $OUT/sanitizer_synccheck_$TAG.log
