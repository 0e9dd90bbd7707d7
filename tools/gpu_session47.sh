#!/bin/bash
# final GPU suite + smoke of the round, and the ncu --set full capture of one fLDA E-step of the final build
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/r2_pytest_final.log 2>&1
tail -5 $O/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; tail -2 $O/r2_smoke.log
md5sum topicmodelsvb.jl_b200/libtmvb.so > $O/r2_lib_md5.txt
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:flda_estep -c 8 -o /tmp/s47_flda python tools/prof_run.py --config nsf_flda_k50 > $O/s47_full_flda.log 2>&1
tail -2 $O/s47_full_flda.log
ncu -i /tmp/s47_flda.ncu-rep --page raw --csv > $O/r2_full_nsf_flda_k50_raw.csv 2>/dev/null
ncu -i /tmp/s47_flda.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r2_full_nsf_flda_k50_source.csv.gz
ls -la $O/r2_full_nsf_flda_k50_*
