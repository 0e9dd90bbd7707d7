#!/bin/bash
# fLDA register-state kernel: 200 vs 168 registers per thread
O=gpurun_out; mkdir -p $O
for r in 168 200; do
  TMVB_FLDA_MAXREG=$r timeout 200 python bench.py --config nsf_flda_k50 --also none --no-cpu-baseline --steps 20 --warmup 5 > $O/s48_flda_r$r.json 2> $O/s48_flda_r$r.err
  python - <<PY
import json
for l in open('$O/s48_flda_r$r.json'):
    if l.startswith('{'):
        d=json.loads(l); print('maxreg=$r', 'ms/step %.4f'%d['ms_per_step'], 'estep', d['roofline'].get('kernel_ms'), 'e2e', d['e2e'].get('ms_per_step'), 'parity', d['parity']['max_rel_vs_oracle'])
PY
done
