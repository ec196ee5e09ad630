#!/bin/bash
# A/B of two builds of libcnerf.so on the same box: bash scripts/ab_lib.sh consistentnerf_b200/libcnerf_head.so [mode]
OTHER=$1; MODE=${2:-train}
for rep in 1 2 3; do
  for lib in "$OTHER" ""; do
    CNERF_LIB=$lib CNERF_SKIP_BUILD=1 timeout 300 python bench.py --mode $MODE --steps 20 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/ab.json
    python -c "
import json; d=json.load(open('/tmp/ab.json')); print('${lib:-current}', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], [round(v['ms_per_step'],3) for v in d['kernels'].values()])"
  done
done
