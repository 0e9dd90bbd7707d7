#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -k "argument_errors or filtered_predict or fresh" > $O/s23_pytest.log 2>&1
tail -4 $O/s23_pytest.log
python bench.py --steps 20 --warmup 5 > $O/s23_bench.json 2> $O/s23_bench.err
tail -c 300 $O/s23_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/s23_bench.json'):
    if not l.startswith('{'): continue
    d=json.loads(l)
    def show(n,d):
        if 'error' in d: print(n, 'ERROR', d['error']); return
        it=[v for k,v in d.items() if k.startswith('e2e_iter')]
        print(n, 'ms/step %.4f'%d['ms_per_step'], 'estep %.3f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3f'%d['e2e']['ms_per_step'], 'iterN %.3f'%(it[0]['ms_per_iteration'] if it else -1), 'parity', d.get('parity',{}).get('max_rel_vs_oracle'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'sec', d.get('bench_seconds'))
    show(d['config']['name'], d)
    for n,c in d.get('configs',{}).items(): show(n,c)
PY
export TMVB_GRAPH=0
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:flda_estep -c 6 -o /tmp/s23_flda python tools/prof_run.py --config nsf_flda_k50 > $O/s23_full_flda.log 2>&1
ncu -i /tmp/s23_flda.ncu-rep --page raw --csv > $O/r2_full_nsf_flda_k50_raw.csv 2>/dev/null
ncu -i /tmp/s23_flda.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r2_full_nsf_flda_k50_source.csv.gz
tail -3 $O/s23_full_flda.log
