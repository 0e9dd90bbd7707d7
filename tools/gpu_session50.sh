#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_lda_gpu.py tests/test_ctm_gpu.py tests/test_ctpf_gpu.py -q -m gpu -x -k "error or golden or host_mirror or second_call or checkelbo" > $O/s50_pytest.log 2>&1; tail -3 $O/s50_pytest.log
timeout 200 python bench.py --also none --no-cpu-baseline > $O/s50_bench.json 2> $O/s50_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/s50_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.4f'%d['ms_per_step'], 'estep', d['roofline']['kernel_ms'], 'e2e', d['e2e']['ms_per_step'], 'iter10', d['e2e_iter10']['ms_per_iteration'], 'parity', d['parity']['max_rel_vs_oracle'])
PY
