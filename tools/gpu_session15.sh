#!/bin/bash
O=gpurun_out; mkdir -p $O
for st in 4 6 8 12; do TMVB_STREAMS=$st python tools/dev_shard_estep.py; done > $O/s15_shard.log 2>&1
TMVB_STREAMS=12 TMVB_HYB_CLASSES="1:2:0,1:4:0,1:6:0,2:4:0,2:6:0,4:6:0,4:6:1" python tools/dev_shard_estep.py >> $O/s15_shard.log 2>&1
TMVB_STREAMS=12 TMVB_HYB_CLASSES="1:3:0,1:6:0,2:6:0,4:6:0,4:6:1" python tools/dev_shard_estep.py >> $O/s15_shard.log 2>&1
for st in 4 12; do WORLD=1 TMVB_STREAMS=$st python tools/dev_shard_estep.py; done >> $O/s15_shard.log 2>&1
for st in 4 12; do WORLD=4 TMVB_STREAMS=$st python tools/dev_shard_estep.py; done >> $O/s15_shard.log 2>&1
cat $O/s15_shard.log
