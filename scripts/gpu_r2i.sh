#!/bin/bash
# Round 2, call I: CTA-pair fp16 forward (mlp_fwd6): parity + timing; trained-weights parity quantiles for both forward precisions.
TAG=${1:-r2i}
OUT=gpurun_out
mkdir -p $OUT
CNERF_FWD_PAIR=1 timeout 300 python scripts/test_pair.py > $OUT/pair_$TAG.txt 2>&1; echo "pair rc=$?"; tail -14 $OUT/pair_$TAG.txt
timeout 120 python scripts/time_fwd.py 2>&1 | grep "infer" > $OUT/time_fwd_$TAG.txt; cat $OUT/time_fwd_$TAG.txt
CNERF_FWD_PAIR=1 timeout 300 python bench.py --mode render --steps 20 --warmup 3 --quick > $OUT/bench_render_pair_$TAG.json 2> $OUT/bench_render_pair_$TAG.err; python -c "
import json; d = json.load(open('$OUT/bench_render_pair_$TAG.json')); print('render pair', round(d['ms_per_step'], 3), 'ms', round(d['value']), 'rays/s e2e', round(d['e2e']['value']))"
for fp in fp16 split; do
  CNERF_FWD_PRECISION=$fp timeout 600 python oracle/twin.py twin --kind blender --root /tmp/twin_sw_$fp --iters 600 --res 200 --eval-views 2 --train-views 8 --arms repo > $OUT/twin_sameweights_${TAG}_$fp.json 2> $OUT/twin_sameweights_${TAG}_$fp.err; echo "twin $fp rc=$?"
  python -c "
import json; d=json.loads(open('$OUT/twin_sameweights_${TAG}_$fp.json').read().strip().splitlines()[-1]); print('$fp', d['repo'].get('same_weights_render_parity'), d['repo'].get('psnr'), d['repo'].get('error'))"
done
