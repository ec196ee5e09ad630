#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (train + render), ncu launch list and full captures of the MLP kernels.
# Usage: gpurun --timeout 1700 -- 'bash scripts/gpu_check.sh TAG'
TAG=${1:-r1c}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_$TAG.log
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/bench_train_$TAG.json 2> $OUT/bench_train_$TAG.err; echo "bench train rc=$?"; cat $OUT/bench_train_$TAG.json
timeout 300 python bench.py --mode render --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_render_$TAG.json 2> $OUT/bench_render_$TAG.err; echo "bench render rc=$?"; cat $OUT/bench_render_$TAG.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"; cat $OUT/bench_ref_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_${TAG}_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'mlp_fused3_kernel|mlp_bwd_data3_kernel' --launch-skip 8 -c 4 -f -o $OUT/prof_fwd_chain_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full1_$TAG.log 2>&1; echo "ncu full fwd/chain rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'mlp_bwd_weight_kernel' --launch-skip 40 -c 3 -f -o $OUT/prof_dw_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full2_$TAG.log 2>&1; echo "ncu full dw rc=$?"
ls -la $OUT
CNERF_MLP_IMPL=3 timeout 200 python scripts/prof_phases.py > $OUT/prof_phases_$TAG.txt 2>&1; echo "phase profile rc=$?"
timeout 200 python scripts/umma_rate.py > $OUT/umma_rate_$TAG.txt 2>&1; echo "umma rate rc=$?"
timeout 200 python scripts/host_time.py 2>&1 | head -4 > $OUT/host_time_$TAG.txt; echo "host time rc=$?"
