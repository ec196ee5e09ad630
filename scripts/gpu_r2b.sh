#!/bin/bash
# Round 2, call B: the unmodified scripts through the drop-in (tests), PSNR twin on the Blender scene, GPU-eager / CPU denominators.
TAG=${1:-r2b}; ITERS=${2:-600}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
nproc > $OUT/nproc_$TAG.txt
timeout 1500 python -m pytest tests/test_gpu_dropin_train.py -x -q > $OUT/pytest_dropin_$TAG.log 2>&1; echo "pytest dropin rc=$?"; tail -15 $OUT/pytest_dropin_$TAG.log
timeout 120 python oracle/twin.py eager --mode train --steps 5 --warmup 2 > $OUT/eager_train_$TAG.json 2> $OUT/eager_train_$TAG.err; echo "eager train rc=$?"; tail -1 $OUT/eager_train_$TAG.json
timeout 120 python oracle/twin.py eager --mode render --steps 5 --warmup 2 > $OUT/eager_render_$TAG.json 2> $OUT/eager_render_$TAG.err; echo "eager render rc=$?"; tail -1 $OUT/eager_render_$TAG.json
timeout 300 python oracle/twin.py eager --device cpu --mode train --steps 1 --warmup 1 > $OUT/cpu_train_$TAG.json 2> $OUT/cpu_train_$TAG.err; echo "cpu train rc=$?"; tail -1 $OUT/cpu_train_$TAG.json
timeout 1500 python oracle/twin.py twin --kind blender --root /tmp/twin_b --iters $ITERS --res 400 --eval-views 2 > $OUT/twin_blender_$TAG.json 2> $OUT/twin_blender_$TAG.err; echo "twin rc=$?"; tail -1 $OUT/twin_blender_$TAG.json | cut -c1-600
cp /tmp/twin_b/log_blender_ref.txt $OUT/twin_blender_ref_$TAG.log 2>/dev/null; cp /tmp/twin_b/log_blender_repo.txt $OUT/twin_blender_repo_$TAG.log 2>/dev/null
for mode in dw16 fp16; do
  CNERF_GRAD_PRECISION=$mode timeout 600 python oracle/twin.py twin --kind blender --root /tmp/twin_b --iters $ITERS --res 400 --eval-views 2 --arms repo > $OUT/twin_blender_${TAG}_$mode.json 2>> $OUT/twin_blender_$TAG.err; echo "twin $mode rc=$?"; tail -1 $OUT/twin_blender_${TAG}_$mode.json | cut -c1-400
done
