#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_flda_gpu.py -q -m gpu -s > $O/s22_filtered.log 2>&1
tail -40 $O/s22_filtered.log
timeout 900 python -m pytest tests/test_ctm_gpu.py tests/test_lda_gpu.py -q -m gpu > $O/s22_other.log 2>&1
tail -3 $O/s22_other.log
