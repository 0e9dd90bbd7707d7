#!/bin/bash
# the full GPU suite on the final state of the round
O=gpurun_out; mkdir -p $O
md5sum topicmodelsvb.jl_b200/libtmvb.so > $O/r2_lib_md5_last.txt
timeout 120 python -m pytest tests -q -m gpu > $O/r2_pytest_last.log 2>&1
tail -4 $O/r2_pytest_last.log
