#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_ctm_gpu.py tests/test_ctpf_gpu.py -q -m gpu > $O/s27_pytest.log 2>&1
tail -3 $O/s27_pytest.log
timeout 600 python -m pytest tests/test_lda_gpu.py -q -m gpu -k "threshold or nsf_shaped" > $O/s27_pytest2.log 2>&1
tail -3 $O/s27_pytest2.log
for c in citeu_ctm_k30 citeu_ctpf_k30 nsf_flda_k50 citeu_fctm_k30; do
python bench.py --steps 20 --warmup 5 --config $c --also none --no-cpu-baseline > $O/s27_bench_$c.json 2> $O/s27_bench_$c.err
tail -c 200 $O/s27_bench_$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s27_bench_*.json')):
    for l in open(f):
        if not l.startswith('{'): continue
        d=json.loads(l)
        it=[v for k,v in d.items() if k.startswith('e2e_iter')]
        print(d['config']['name'], 'ms/step %.4f'%d['ms_per_step'], 'estep %.3f'%d['roofline']['kernel_ms'], 'e2e %.3f'%d['e2e']['ms_per_step'], 'iterN %.3f'%it[0]['ms_per_iteration'], 'parity', d['parity'].get('max_rel_vs_oracle'), 'h2d', d['e2e']['h2d_bytes_per_step'], 'd2h', d['e2e']['d2h_bytes_per_step'])
PY
