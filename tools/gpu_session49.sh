#!/bin/bash
# final N=1 evidence of the round (final build): default bench line (six configurations), reference arm, fLDA tests, fLDA ncu capture
O=gpurun_out; mkdir -p $O
md5sum topicmodelsvb.jl_b200/libtmvb.so > $O/r2_lib_md5.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/r2_clocks.csv &
SMI=$!
( time python bench.py ) > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
tail -c 300 $O/r2_bench_n1.err
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
            print(n, 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d.get('ms_per_step_min_med_max'), 'estep', d.get('roofline',{}).get('kernel_ms'), 'frac', d.get('roofline',{}).get('frac'), 'e2e', d['e2e'].get('ms_per_step'), 'iterN', (it[0]['ms_per_iteration'] if it else None), 'parity', d.get('parity',{}).get('max_rel_vs_oracle'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'launches', d.get('gpu_launches'))
        show(d['config']['name'], d)
        for n,c in d.get('configs',{}).items(): show(n,c)
PY
timeout 300 python -m pytest tests/test_flda_gpu.py -q -m gpu > $O/r2_pytest_flda_final.log 2>&1; tail -3 $O/r2_pytest_flda_final.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:flda_estep -c 8 -o /tmp/s49_flda python tools/prof_run.py --config nsf_flda_k50 > $O/s49_full_flda.log 2>&1
ncu -i /tmp/s49_flda.ncu-rep --page raw --csv > $O/r2_full_nsf_flda_k50_raw.csv 2>/dev/null
ncu -i /tmp/s49_flda.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r2_full_nsf_flda_k50_source.csv.gz
ls -la $O/r2_full_nsf_flda_k50_*
