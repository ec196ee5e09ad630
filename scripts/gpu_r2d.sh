#!/bin/bash
# Round 2, call D: full GPU suite, phase profile + ncu capture of the fp16 forward, new bench line, twin with the fp16 forward.
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_$TAG.log
timeout 300 python scripts/prof_phases.py > $OUT/prof_phases_$TAG.txt 2>&1; echo "phases rc=$?"; grep -A10 "^fwd5" $OUT/prof_phases_$TAG.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'mlp_fused5_kernel' --launch-skip 5 -c 1 -f -o $OUT/prof_fwd5_$TAG python scripts/time_fwd.py > $OUT/ncu_fwd5_$TAG.log 2>&1; echo "ncu fwd5 rc=$?"
python scripts/ncu_summary.py full $OUT/prof_fwd5_$TAG.ncu-rep $OUT/ncu_fwd5_summary_$TAG.txt > /dev/null 2>&1; head -30 $OUT/ncu_fwd5_summary_$TAG.txt
for fp in split fp16; do
  CNERF_FWD_PRECISION=$fp timeout 900 python bench.py --steps 20 --warmup 3 --grad-precision fp16 --no-quality > $OUT/bench_full_${TAG}_$fp.json 2> $OUT/bench_full_${TAG}_$fp.err; echo "bench full $fp rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_full_${TAG}_$fp.json"))
    print("fwd=$fp", round(d["ms_per_step"], 3), "ms (host launch", round(d["config"]["ms_per_step_host_launch"], 3), ") e2e", round(d["e2e"]["ms_per_step"], 3), "graph", d["config"]["cuda_graph"], d["config"]["cuda_graph_error"], d["clocks"]["sm_mhz"])
    print("   kernels", {k.replace("cnerf_mlp_", ""): round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
    r = d["render"]; print("   render", round(r["ms_per_step"], 3), "ms", round(r["value"]), "rays/s; image", round(r["image"]["ms_per_image"], 1), "ms")
    print("   cpu", d["cpu_baseline"]); print("   eager", d.get("gpu_eager"))
except Exception as e:
    print("no result", e)
PY
done
CNERF_FWD_PRECISION=fp16 CNERF_GRAD_PRECISION=fp16 timeout 600 python oracle/twin.py twin --kind blender --root /tmp/twin_b --iters 600 --res 400 --eval-views 2 --arms repo > $OUT/twin_blender_${TAG}_fwd16.json 2> $OUT/twin_blender_${TAG}_fwd16.err; echo "twin fwd16 rc=$?"; tail -1 $OUT/twin_blender_${TAG}_fwd16.json | cut -c1-500
