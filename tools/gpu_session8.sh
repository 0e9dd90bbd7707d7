#!/bin/bash
O=gpurun_out; mkdir -p $O
python tools/dev_hyb.py nsf default "1:2:0,1:3:0,1:4:0,1:5:0,1:6:0,2:4:0,2:5:0,2:6:0,4:5:0,4:6:0,4:6:1" > $O/s8_nsf.log 2>&1
cat $O/s8_nsf.log
M=100000 python tools/dev_hyb.py k200 default "4:3:0,4:4:0,4:5:0,4:6:0,4:6:2,4:6:1" "4:4:0,4:6:0,4:6:2,4:6:1" "4:3:0,4:4:0,4:4:2,4:4:1" > $O/s8_k200.log 2>&1
cat $O/s8_k200.log
timeout 900 python -m pytest tests/test_lda_gpu.py -x -q -m gpu > $O/s8_pytest_lda.log 2>&1
tail -5 $O/s8_pytest_lda.log
