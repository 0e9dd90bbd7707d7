#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_next_rows_gpu.py -q -m gpu -k "two_gpu" -x > $O/s30_pytest_2gpu.log 2>&1
tail -30 $O/s30_pytest_2gpu.log
