#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/s9_pytest.log 2>&1
tail -8 $O/s9_pytest.log
python bench.py --steps 20 --warmup 5 --also none > $O/s9_bench.json 2> $O/s9_bench.err
tail -c 300 $O/s9_bench.err; python -c "
import json; d=json.load(open('$O/s9_bench.json')); print('ms/step', d['ms_per_step'], 'estep', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'], 'iter10', d['e2e_iter10']['ms_per_iteration'], 'parity', d['parity']['max_rel_vs_oracle'])"
TMVB_GRAPH=0 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hyb -o /tmp/s9_full python tools/prof_run.py --config nsf_lda_k50 > $O/s9_full.log 2>&1
ncu -i /tmp/s9_full.ncu-rep --page raw --csv > $O/s9_full_raw.csv 2>/dev/null
ncu -i /tmp/s9_full.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/s9_full_source.csv.gz
