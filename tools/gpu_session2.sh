#!/bin/bash
set -x
O=gpurun_out; mkdir -p $O
python tools/dev_hyb.py nsf reg default "1:2:0,1:2:12,2:3:5,4:3:2,4:3:1" "1:2:0,1:4:9,2:3:5,4:3:2,4:3:1" "1:2:0,1:6:8,2:3:5,4:3:2,4:3:1" "1:2:0,1:6:8,2:6:4,4:3:2,4:3:1" "1:2:0,1:3:10,2:4:4,4:4:2,4:3:1" "1:2:0,1:3:10,2:3:5,2:6:4,4:3:1" > $O/s2_hyb_nsf.log 2>&1
cat $O/s2_hyb_nsf.log
M=100000 python tools/dev_hyb.py k200 reg default "4:4:2,4:4:1" "4:3:2,4:3:1" > $O/s2_hyb_k200.log 2>&1
cat $O/s2_hyb_k200.log
timeout 900 python -m pytest tests/test_lda_gpu.py -x -q -m gpu > $O/s2_pytest_lda.log 2>&1
tail -15 $O/s2_pytest_lda.log
