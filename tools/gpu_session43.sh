#!/bin/bash
# fLDA: shared-memory tile capacity (TMVB_TILE_CAP_MAX) against occupancy
O=gpurun_out; mkdir -p $O
for cap in 0 16 32 48 64 96; do
  if [ $cap = 0 ]; then unset TMVB_TILE_CAP_MAX; else export TMVB_TILE_CAP_MAX=$cap; fi
  timeout 200 python bench.py --config nsf_flda_k50 --also none --no-cpu-baseline --steps 20 --warmup 5 > $O/s43_flda_cap$cap.json 2> $O/s43_flda_cap$cap.err
  python - <<PY
import json
for l in open('$O/s43_flda_cap$cap.json'):
    if l.startswith('{'):
        d=json.loads(l); print('cap=$cap', 'ms/step %.4f'%d['ms_per_step'], 'estep', d['roofline'].get('kernel_ms'), 'e2e', d['e2e'].get('ms_per_step'), 'parity', d['parity']['max_rel_vs_oracle'])
PY
done
