#!/bin/bash
O=gpurun_out; mkdir -p $O
for rep in 1 2; do
for so in "" topicmodelsvb.jl_b200/variants/libtmvb_dot4.so; do
echo "== TMVB_SO=$so"
TMVB_SO=$so ITERS=6 python tools/dev_hyb.py nsf default
TMVB_SO=$so ITERS=3 M=100000 python tools/dev_hyb.py k200 default
done
done > $O/s33_dot4.log 2>&1
cut -c1-150 $O/s33_dot4.log
