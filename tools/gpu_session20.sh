#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_flda_gpu.py -q -m gpu -s > $O/s20_flda.log 2>&1
tail -40 $O/s20_flda.log
timeout 600 python -m pytest tests -q -m gpu -k "fresh or elbo_trajectory_small" > $O/s20_other.log 2>&1
tail -3 $O/s20_other.log
