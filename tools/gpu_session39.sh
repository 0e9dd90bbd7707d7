#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_ctpf_gpu.py tests/test_lda_gpu.py -q -m gpu > $O/s39_pytest.log 2>&1; tail -3 $O/s39_pytest.log
python bench.py --steps 20 --warmup 5 --config citeu_ctpf_k30 --also none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['name'], 'ms/step %.4f'%d['ms_per_step'], 'estep %.3f'%d['roofline']['kernel_ms'], 'e2e %.3f'%d['e2e']['ms_per_step'], 'parity', d['parity'].get('max_rel_vs_oracle'))"
