#!/bin/bash
# host mirror (update_host! of gamma / Elogtheta inside the last E-step): tests, then the default bench line with and without it
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_lda_gpu.py -q -m gpu -k "host_mirror or golden or fused_iteration" > $O/s41_pytest.log 2>&1; tail -15 $O/s41_pytest.log
for m in 1 0; do
  TMVB_HOST_MIRROR=$m timeout 300 python bench.py --also none --no-cpu-baseline > $O/s41_bench_m$m.json 2> $O/s41_bench_m$m.err
  python - <<PY
import json
for l in open('$O/s41_bench_m$m.json'):
    if l.startswith('{'):
        d=json.loads(l); print('mirror=$m', 'ms/step %.4f'%d['ms_per_step'], 'estep', d['roofline']['kernel_ms'], 'e2e', d['e2e'], 'iter10', d['e2e_iter10']['ms_per_iteration'], 'parity', d['parity']['max_rel_vs_oracle'])
PY
done
