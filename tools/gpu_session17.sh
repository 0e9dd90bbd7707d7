#!/bin/bash
# recs tests first (new tcgen05 kernel: own timeout), then the whole GPU suite, then a default bench
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_ctpf_gpu.py -x -q -m gpu -k recs -s > $O/s17_recs.log 2>&1
tail -15 $O/s17_recs.log
timeout 1500 python -m pytest tests -x -q -m gpu > $O/s17_pytest.log 2>&1
tail -5 $O/s17_pytest.log
python bench.py --steps 20 --warmup 5 > $O/s17_bench.json 2> $O/s17_bench.err
tail -c 300 $O/s17_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/s17_bench.json'):
    if not l.startswith('{'): continue
    d=json.loads(l)
    print(d['config'].get('name'), 'ms/step %.4f'%d['ms_per_step'], 'roof', {k:d['roofline'].get(k) for k in ('kernel','kernel_ms','frac','achieved')}, 'e2e', d['e2e'].get('ms_per_step'), 'parity', d.get('parity',{}).get('max_rel_vs_oracle'), 'cpu', d.get('cpu_baseline',{}).get('value'))
    for k,v in d.items():
        if k.startswith('also'): print(k, json.dumps(v)[:1500])
PY
