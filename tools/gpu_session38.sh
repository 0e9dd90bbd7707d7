#!/bin/bash
O=gpurun_out; mkdir -p $O
A="1:2:0,1:3:0,1:4:0,1:5:0,1:6:0,2:4:0,2:5:0,2:6:0,4:5:0,4:6:0,4:6:1"
B="1:2:0,1:4:0,1:6:0,2:4:0,2:6:0,4:6:0,4:6:1"
C="1:3:0,1:5:0,1:6:0,2:4:0,2:6:0,4:6:0,4:6:1"
D="1:2:0,1:4:0,1:5:0,1:6:0,2:4:0,2:6:0,4:6:0,4:6:1"
for w in 8 4; do
for spec in "$A" "$B" "$C" "$D"; do
for st in 4 8; do
WORLD=$w TMVB_STREAMS=$st TMVB_HYB_CLASSES="$spec" ITERS=8 python tools/dev_shard_estep.py
done
done
done > $O/s38_classes.log 2>&1
cut -c1-200 $O/s38_classes.log
