#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -q -m gpu > $O/r2_pytest_final.log 2>&1
tail -5 $O/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; tail -2 $O/r2_smoke.log
