#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -k "fresh or recs or fused" -s > $O/s19_pytest.log 2>&1
tail -4 $O/s19_pytest.log
python bench.py --steps 20 --warmup 5 --also none --no-cpu-baseline > $O/s19_bench.json 2> $O/s19_bench.err
tail -c 300 $O/s19_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/s19_bench.json'):
    if not l.startswith('{'): continue
    d=json.loads(l)
    print(d['config'].get('name'), 'ms/step %.4f'%d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), 'iter10', d['e2e_iter10']['ms_per_iteration'], 'parity', d.get('parity',{}).get('max_rel_vs_oracle'))
PY
export TMVB_GRAPH=0
for c in citeu_ctm_k30 citeu_ctpf_k30; do
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s19_launches_e2e_$c.csv python tools/prof_run.py --config $c --e2e > $O/s19_prof_e2e_$c.log 2>&1
done
python tools/dev_e2e_profile.py > $O/s19_e2e_profile.log 2>&1; head -14 $O/s19_e2e_profile.log
