#!/bin/bash
O=gpurun_out; mkdir -p $O
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 30 --warmup 10 --also none > $O/s14_bench_n$n.json 2> $O/s14_bench_n$n.err
tail -c 400 $O/s14_bench_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 5 --config synth_lda_k200 --also none > $O/s14_bench_k200_n8.json 2> $O/s14_bench_k200_n8.err
tail -c 400 $O/s14_bench_k200_n8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s14_bench_*.json')):
    l=[x for x in open(f).read().splitlines() if x.startswith('{')]
    if not l: print(f,'EMPTY'); continue
    d=json.loads(l[-1])
    print(f, 'N%d ms/step %.4f estep %.4f e2e ms %.3f iterN %.3f parity %s %s'%(d['n_gpus'],d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], [v for k,v in d.items() if k.startswith('e2e_iter')][0]['ms_per_iteration'], d['parity'].get('max_rel_vs_oracle'), [round(x,3) for x in d['ms_per_step_min_med_max']]))
PY
