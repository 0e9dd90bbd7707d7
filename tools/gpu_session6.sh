#!/bin/bash
O=gpurun_out; mkdir -p $O
SPEC="1:2:0,1:3:0,1:4:0,2:3:0,2:4:0,4:3:0,4:4:0,4:4:1"
for v in base opq nds both sts; do
  echo "== $v" >> $O/s6_variants.log
  TMVB_SO=topicmodelsvb.jl_b200/variants/libtmvb_$v.so python tools/dev_hyb.py nsf "$SPEC" >> $O/s6_variants.log 2>&1
  TMVB_SO=topicmodelsvb.jl_b200/variants/libtmvb_$v.so M=100000 python tools/dev_hyb.py k200 "4:3:0,4:3:3,4:3:2,4:3:1" >> $O/s6_variants.log 2>&1
done
cat $O/s6_variants.log
