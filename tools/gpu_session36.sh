#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_ctm_gpu.py -q -m gpu > $O/s36_pytest.log 2>&1; tail -3 $O/s36_pytest.log
timeout 600 python -m pytest tests/test_flda_gpu.py tests/test_next_rows_gpu.py -q -m gpu -k "fctm or predict_ctm" > $O/s36_pytest2.log 2>&1; tail -3 $O/s36_pytest2.log
for c in citeu_ctm_k30 citeu_fctm_k30; do
python bench.py --steps 10 --warmup 3 --config $c --also none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['name'], 'ms/step %.4f'%d['ms_per_step'], 'estep %.3f'%d['roofline']['kernel_ms'], 'parity', d['parity'].get('max_rel_vs_oracle'))"
done > $O/s36_ctm.log 2>&1
cat $O/s36_ctm.log
