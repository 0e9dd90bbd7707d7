#!/bin/bash
set -x
O=gpurun_out; mkdir -p $O
python tools/dev_hyb.py nsf "1:2:0,1:3:0,1:4:0,2:3:0,2:4:0,4:3:0,4:4:0,4:4:1" "1:2:0,1:3:0,2:2:0,2:3:0,2:4:0,4:3:0,4:4:0,4:4:1" "1:2:0,1:3:0,1:4:0,2:3:0,2:4:0,4:3:0,4:3:2,4:3:1" > $O/s5_hyb_nsf.log 2>&1
cat $O/s5_hyb_nsf.log
M=100000 python tools/dev_hyb.py k200 default "4:3:0,4:3:3,4:3:2,4:3:1" > $O/s5_hyb_k200.log 2>&1
cat $O/s5_hyb_k200.log
timeout 900 python -m pytest tests/test_lda_gpu.py -x -q -m gpu > $O/s5_pytest_lda.log 2>&1
tail -5 $O/s5_pytest_lda.log
TMVB_GRAPH=0 TMVB_HYB_CLASSES="1:2:0,1:3:0,1:4:0,2:3:0,2:4:0,4:3:0,4:4:0,4:4:1" ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hyb -o /tmp/s5_full python tools/prof_run.py --config nsf_lda_k50 > $O/s5_full.log 2>&1
ncu -i /tmp/s5_full.ncu-rep --page raw --csv > $O/s5_full_raw.csv 2>/dev/null
ncu -i /tmp/s5_full.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/s5_full_source.csv.gz
