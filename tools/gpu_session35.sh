#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_ctm_gpu.py -q -m gpu > $O/s35_pytest.log 2>&1; tail -3 $O/s35_pytest.log
timeout 600 python -m pytest tests/test_flda_gpu.py -q -m gpu -k fctm > $O/s35_pytest2.log 2>&1; tail -3 $O/s35_pytest2.log
for so in "" topicmodelsvb.jl_b200/libtmvb_ctm_12.so topicmodelsvb.jl_b200/libtmvb_ctm_8.so topicmodelsvb.jl_b200/libtmvb_ctm_smem.so; do
for c in citeu_ctm_k30 citeu_fctm_k30; do
TMVB_SO=$so python bench.py --steps 10 --warmup 3 --config $c --also none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('so=$so', d['config']['name'], 'ms/step %.4f'%d['ms_per_step'], 'estep %.3f'%d['roofline']['kernel_ms'], 'parity', d['parity'].get('max_rel_vs_oracle'))"
done
done > $O/s35_ctm_variants.log 2>&1
cat $O/s35_ctm_variants.log
