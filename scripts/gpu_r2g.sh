#!/bin/bash
# Round 2, call G: PSNR twin with 8 training views (well-conditioned), 3 seeds; ncu launch list of the fp16 step.
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python scripts/psnr_twin.py --iters 1000 --seeds 0 1 2 --train-views 8 --root /tmp/cnerf_psnr_twin8 --repo-modes split:split fp16:fp16 --out $OUT/psnr_twin8_$TAG.json > $OUT/psnr_twin8_$TAG.log 2>&1; echo "psnr twin rc=$?"; tail -2 $OUT/psnr_twin8_$TAG.log | cut -c1-1500
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_${TAG}_train.csv python bench.py --steps 2 --warmup 3 --quick --no-graph --fwd-precision fp16 --grad-precision fp16 > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
python scripts/ncu_summary.py launches $OUT/launches_${TAG}_train.csv $OUT/launches_${TAG}_train.txt | head -30
