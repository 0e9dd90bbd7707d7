#!/bin/bash
set -x
O=gpurun_out; mkdir -p $O
python tools/dev_hyb.py nsf reg default "1:2:0,1:3:0,1:4:0,2:3:0,2:4:0,4:3:0,4:4:0,4:4:1" "1:2:0,2:2:0,2:3:0,2:4:0,4:3:0,4:4:0,4:4:1" "1:2:0,1:4:0,2:3:0,2:4:0,4:3:0,4:3:2,4:3:1" "1:2:0,1:4:0,2:3:0,2:3:6,4:3:0,4:4:0,4:4:1" "1:2:0,1:4:0,2:3:0,2:2:6,4:3:2,4:3:1" > $O/s3_hyb_nsf.log 2>&1
cat $O/s3_hyb_nsf.log
M=100000 python tools/dev_hyb.py k200 reg default "4:4:2,4:4:1" "4:3:0,4:3:3,4:3:2,4:3:1" > $O/s3_hyb_k200.log 2>&1
cat $O/s3_hyb_k200.log
TMVB_GRAPH=0 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hyb -o /tmp/s3_full python tools/prof_run.py --config nsf_lda_k50 > $O/s3_full.log 2>&1
ncu -i /tmp/s3_full.ncu-rep --page raw --csv > $O/s3_full_raw.csv 2>/dev/null
ncu -i /tmp/s3_full.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/s3_full_source.csv.gz
ls -la $O | tail -8
