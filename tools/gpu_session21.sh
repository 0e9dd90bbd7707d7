#!/bin/bash
# 2 GPUs: the multi-GPU parity tests (LDA exchange paths, sharded CTM / CTPF) and the N=2 bench line
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_next_rows_gpu.py -q -m gpu -k "two_gpu" > $O/s21_pytest_2gpu.log 2>&1
tail -5 $O/s21_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 30 --warmup 10 --also none > $O/r2_bench_n2_nsf_lda_k50.json 2> $O/s21_bench_n2.err
tail -c 300 $O/s21_bench_n2.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench_n2_nsf_lda_k50.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print('N2 ms/step %.4f'%d['ms_per_step'], d['ms_per_step_min_med_max'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity'].get('max_rel_vs_oracle'))
PY
