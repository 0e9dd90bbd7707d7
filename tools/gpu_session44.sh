#!/bin/bash
# fLDA register-state E-step variant: parity tests, then the bench line with TMVB_FLDA_REG = 0 (tile kernel, 64-token tile), 1 (200 registers), 2 (uncapped)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_flda_gpu.py -q -m gpu -x > $O/s44_pytest.log 2>&1; tail -8 $O/s44_pytest.log
for v in 0 1 2; do
  TMVB_FLDA_REG=$v timeout 200 python bench.py --config nsf_flda_k50 --also none --no-cpu-baseline --steps 20 --warmup 5 > $O/s44_flda_v$v.json 2> $O/s44_flda_v$v.err
  python - <<PY
import json
for l in open('$O/s44_flda_v$v.json'):
    if l.startswith('{'):
        d=json.loads(l); print('reg=$v', 'ms/step %.4f'%d['ms_per_step'], 'estep', d['roofline'].get('kernel_ms'), 'e2e', d['e2e'].get('ms_per_step'), 'parity', d['parity']['max_rel_vs_oracle'])
PY
done
