#!/bin/bash
# Round 2, call H: fused dW launch + fp16 defaults: full GPU suite, smoke, default bench (all legs), ncu launch list + full captures.
TAG=${1:-r2h}
OUT=gpurun_out
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; echo "bench default rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_default_$TAG.json"))
    print(round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s  e2e", round(d["e2e"]["ms_per_step"], 3), "ms graph", d["config"]["cuda_graph"], d["config"]["cuda_graph_error"], d["clocks"])
    print("   kernels", {k.replace("cnerf_mlp_", ""): round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, "launches", d["gpu_launches"])
    print("   roofline", {k: d["roofline"][k] for k in ("bound", "achieved", "peak", "frac", "tensor_frac", "mlp_step_tensor_frac")})
    r = d["render"]; print("   render", round(r["ms_per_step"], 3), "ms", round(r["value"]), "rays/s frac", round(r["roofline"]["frac"], 3), "; image", round(r["image"]["ms_per_image"], 1), "ms")
    print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"]); print("   eager", d["gpu_eager"]["train"], d["gpu_eager"]["render"])
    q = d["quality"]; print("   quality", {k: q.get(k) for k in ("psnr_repo", "psnr_ref", "delta_db", "seconds", "error")})
except Exception as e:
    print("no result", e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"; cut -c1-600 $OUT/bench_ref_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_${TAG}_train.csv python bench.py --steps 2 --warmup 3 --quick --no-graph > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
python scripts/ncu_summary.py launches $OUT/launches_${TAG}_train.csv $OUT/launches_${TAG}_train.txt | head -14
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'mlp_fused5_kernel|mlp_bwd_data5_kernel|mlp_bwd_weight_kernel|mlp_heads_grad_kernel' --launch-skip 16 -c 8 -f -o $OUT/prof_step_$TAG python bench.py --steps 2 --warmup 3 --quick --no-graph > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py full $OUT/prof_step_$TAG.ncu-rep $OUT/ncu_full_summary_$TAG.txt > /dev/null 2>&1; grep -E "Kernel Name|gpu__time_duration|dram__bytes|tensor_cycles_active_realtime" $OUT/ncu_full_summary_$TAG.txt | cut -c1-400
timeout 300 python scripts/prof_phases.py > $OUT/prof_phases_$TAG.txt 2>&1
timeout 600 python oracle/twin.py twin --kind blender --root /tmp/twin_sw --iters 400 --res 200 --eval-views 2 --train-views 8 --arms repo > $OUT/twin_sameweights_$TAG.json 2> $OUT/twin_sameweights_$TAG.err; echo "twin same-weights rc=$?"; python -c "
import json; d=json.loads(open('$OUT/twin_sameweights_$TAG.json').read().strip().splitlines()[-1]); print(d['repo'].get('same_weights_render_parity'), d['repo'].get('psnr'), d['repo'].get('error'))"
