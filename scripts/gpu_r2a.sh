#!/bin/bash
# Round 2, call A: GPU parity tests with the precision-mode backward, reduced-MMA table, bench A/B of the three backward modes.
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -15 $OUT/pytest_$TAG.log
for mode in split dw16 fp16 split dw16 fp16; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --grad-precision $mode > $OUT/bench_train_${TAG}_$mode.json 2> $OUT/bench_train_${TAG}_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_train_${TAG}_$mode.json"))
    print("$mode", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["ms_per_step"], 3), d["clocks"]["sm_mhz"], {k.replace("cnerf_mlp_", ""): round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
except Exception as e:
    print("$mode: no result", e)
PY
done
timeout 900 python scripts/mma_terms.py > $OUT/mma_terms_$TAG.txt 2>&1; echo "mma_terms rc=$?"; tail -12 $OUT/mma_terms_$TAG.txt
