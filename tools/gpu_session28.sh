#!/bin/bash
O=gpurun_out; mkdir -p $O
for c in citeu_ctpf_k30 citeu_ctm_k30 nsf_flda_k50; do python tools/dev_e2e_profile2.py $c 2>&1 | grep -v "^$" | head -26; done > $O/s28_e2e.log 2>&1
cat $O/s28_e2e.log
