#!/bin/bash
O=gpurun_out; mkdir -p $O
export TMVB_GRAPH=0
md5sum topicmodelsvb.jl_b200/libtmvb.so > $O/r2b_lib_md5.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_citeu_ctm_k30.csv python tools/prof_run.py --config citeu_ctm_k30 > $O/s37_prof.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:estep -c 4 -o /tmp/s37_ctm python tools/prof_run.py --config citeu_ctm_k30 > $O/s37_full.log 2>&1
ncu -i /tmp/s37_ctm.ncu-rep --page raw --csv > $O/r2_full_citeu_ctm_k30_raw.csv 2>/dev/null
ncu -i /tmp/s37_ctm.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r2_full_citeu_ctm_k30_source.csv.gz
tail -2 $O/s37_full.log
