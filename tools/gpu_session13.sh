#!/bin/bash
O=gpurun_out; mkdir -p $O
CAPS=0 python tools/dev_ctm.py ctm ctpf > $O/s13_ctm.log 2>&1; cat $O/s13_ctm.log
timeout 1500 python -m pytest tests/test_ctm_gpu.py tests/test_ctpf_gpu.py tests/test_next_rows_gpu.py -x -q -m gpu > $O/s13_pytest.log 2>&1
tail -5 $O/s13_pytest.log
