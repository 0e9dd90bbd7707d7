#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_lda_gpu.py tests/test_next_rows_gpu.py -x -q -m gpu > $O/s10_pytest.log 2>&1
tail -12 $O/s10_pytest.log
python bench.py --steps 20 --warmup 5 --also none > $O/s10_bench_n1.json 2> $O/s10_bench_n1.err
tail -c 300 $O/s10_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s10_bench_n1.json')); print('N1 ms/step', d['ms_per_step'], 'estep', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'], 'iter10', d['e2e_iter10']['ms_per_iteration'], 'parity', d['parity']['max_rel_vs_oracle'], 'launches', d['gpu_launches'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 5 --also none > $O/s10_bench_n2.json 2> $O/s10_bench_n2.err
tail -c 600 $O/s10_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s10_bench_n2.json')); print('N2 ms/step', d['ms_per_step'], 'estep', d['roofline']['kernel_ms'], 'e2e ms', d['e2e']['ms_per_step'], 'iter10', d['e2e_iter10']['ms_per_iteration'], 'parity', d['parity']['max_rel_vs_oracle'], d['config']['exchange'][:40])
PY
python tools/dev_e2e_profile.py > $O/s10_e2e_profile.log 2>&1
head -40 $O/s10_e2e_profile.log
