#!/bin/bash
# 8 GPUs: NSF N=8 / N=4 bench lines, cfg4 N=8, the reference arm is run on one GPU box separately
O=gpurun_out; mkdir -p $O
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 30 --warmup 10 --also none > $O/r2_bench_n${n}_nsf_lda_k50.json 2> $O/s25_bench_n$n.err
tail -c 300 $O/s25_bench_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 10 --warmup 5 --config synth_lda_k200 --also none > $O/r2_bench_n8_synth_lda_k200.json 2> $O/s25_bench_k200_n8.err
tail -c 300 $O/s25_bench_k200_n8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n[48]_*.json')):
    l=[x for x in open(f).read().splitlines() if x.startswith('{')]
    if not l: print(f,'EMPTY'); continue
    d=json.loads(l[-1])
    print(f, 'N%d ms/step %.4f estep %.4f e2e ms %.3f iterN %.3f parity %s %s'%(d['n_gpus'],d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], [v for k,v in d.items() if k.startswith('e2e_iter')][0]['ms_per_iteration'], d['parity'].get('max_rel_vs_oracle'), [round(x,3) for x in d['ms_per_step_min_med_max']]))
    print('   ', d.get('ms_per_step_rank0'))
PY
