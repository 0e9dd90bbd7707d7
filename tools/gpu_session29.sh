#!/bin/bash
# final N=1 evidence of the round: the default bench line (all six configurations), the reference arm, launch list of the default config
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/r2_clocks.csv &
SMI=$!
( time python bench.py ) > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
tail -c 400 $O/r2_bench_n1.err
kill $SMI
( time python bench.py --impl reference --steps 10 --warmup 3 ) > $O/r2_bench_reference_n1.json 2> $O/r2_bench_reference_n1.err
tail -c 200 $O/r2_bench_reference_n1.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n1.json','gpurun_out/r2_bench_reference_n1.json'):
    for l in open(f):
        if not l.startswith('{'): continue
        d=json.loads(l)
        def show(n,d):
            if 'error' in d: print(n,'ERROR',d['error']); return
            it=[v for k,v in d.items() if k.startswith('e2e_iter')]
            print(n, 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d.get('ms_per_step_min_med_max'), 'estep', d.get('roofline',{}).get('kernel_ms'), 'frac', d.get('roofline',{}).get('frac'), 'traffic', d.get('roofline',{}).get('traffic'), 'e2e', d['e2e'].get('ms_per_step'), 'iterN', (it[0]['ms_per_iteration'] if it else None), 'parity', d.get('parity',{}).get('max_rel_vs_oracle'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'launches', d.get('gpu_launches'))
        show(d['config']['name'], d)
        for n,c in d.get('configs',{}).items(): show(n,c)
PY
