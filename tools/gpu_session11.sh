#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/s11_pytest.log 2>&1
tail -15 $O/s11_pytest.log
