#!/bin/bash
# fLDA register-state kernel with one / two warps per document: parity tests, racecheck, bench A/B
O=gpurun_out; mkdir -p $O
TMVB_FLDA_REG=2 timeout 600 python -m pytest tests/test_flda_gpu.py -q -m gpu -k "flda or filtered" > $O/s46_pytest_reg2.log 2>&1; tail -6 $O/s46_pytest_reg2.log
TMVB_FLDA_REG=1 timeout 600 python -m pytest tests/test_flda_gpu.py -q -m gpu -k "flda_elbo or ragged" > $O/s46_pytest_reg1.log 2>&1; tail -3 $O/s46_pytest_reg1.log
TMVB_FLDA_REG=2 timeout 240 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_flda_gpu.py -q -m gpu -k "ragged" > $O/s46_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed|Race reported|error" $O/s46_racecheck.log | head -8
for v in 0 1 2; do
  TMVB_FLDA_REG=$v timeout 200 python bench.py --config nsf_flda_k50 --also none --no-cpu-baseline --steps 20 --warmup 5 > $O/s46_flda_v$v.json 2> $O/s46_flda_v$v.err
  python - <<PY
import json
for l in open('$O/s46_flda_v$v.json'):
    if l.startswith('{'):
        d=json.loads(l); print('reg=$v', 'ms/step %.4f'%d['ms_per_step'], 'estep', d['roofline'].get('kernel_ms'), 'e2e', d['e2e'].get('ms_per_step'), 'parity', d['parity']['max_rel_vs_oracle'])
PY
done
