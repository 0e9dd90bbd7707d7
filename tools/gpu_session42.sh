#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_lda_gpu.py -q -m gpu -k "host_mirror" > $O/s42_pytest.log 2>&1; tail -8 $O/s42_pytest.log
timeout 300 python tools/dev_e2e_profile2.py nsf_lda_k50 > $O/s42_e2e.log 2>&1; head -24 $O/s42_e2e.log
