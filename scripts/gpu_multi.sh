#!/bin/bash
# Multi-GPU check (run under `gpurun --gpus N`): 2-rank NCCL gradient-equality test, then bench.py at the given rank counts.
#   bash scripts/gpu_multi.sh TAG "2" | "2:weak,strong 4:weak 8:weak,strong"
TAG=${1:-r2m}; NS=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -s > $OUT/pytest_multi_$TAG.log 2>&1; echo "pytest multi rc=$?"; tail -4 $OUT/pytest_multi_$TAG.log
PORT=29511
for ITEM in $NS; do
  N=${ITEM%%:*}; MODES=${ITEM#*:}; if [ "$MODES" = "$ITEM" ]; then MODES="weak,strong"; fi
  for SC in ${MODES//,/ }; do
    PORT=$((PORT+1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 3 --scaling $SC > $OUT/bench_${TAG}_${N}gpu_$SC.json 2> $OUT/bench_${TAG}_${N}gpu_$SC.err; echo "bench N=$N $SC rc=$?"
    python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${TAG}_${N}gpu_$SC.json").read().strip().splitlines()[-1])
    r = d.get("render") or {}
    print("N=$N $SC", round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s  (host launch", round(d["config"]["ms_per_step_host_launch"], 3), "ms) e2e", round(d["e2e"]["value"]), "graph", d["config"]["cuda_graph"], d["config"]["cuda_graph_error"],
          "| render", round(r.get("value", 0)), "rays/s image", round((r.get("image") or {}).get("ms_per_image", 0), 1), "ms")
except Exception as e:
    print("N=$N $SC: no result", e); print(open("$OUT/bench_${TAG}_${N}gpu_$SC.err").read()[-1500:])
PY
  done
done
