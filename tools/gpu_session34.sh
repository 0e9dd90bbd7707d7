#!/bin/bash
O=gpurun_out; mkdir -p $O
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2990$n tools/dev_exchange_time.py 2>&1 | grep "world\|rank"
done > $O/r2_exchange_kernel_time.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 tools/dev_exchange_time.py synth_lda_k200 2>&1 | grep "world\|rank" >> $O/r2_exchange_kernel_time.txt
cat $O/r2_exchange_kernel_time.txt
