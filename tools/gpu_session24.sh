#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu -k "argument_errors or filtered_predict" > $O/s24_pytest.log 2>&1
tail -3 $O/s24_pytest.log
for w in 8 4 2 1; do
for k in 1 2 3 4 6 8; do
WORLD=$w TMVB_DOCS_PER_CTA=$k python tools/dev_shard_estep.py | sed "s/^/dpc=$k /"
done
done > $O/s24_dpc.log 2>&1
cat $O/s24_dpc.log
