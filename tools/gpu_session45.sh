#!/bin/bash
# ncu --set full of one fLDA E-step (64-token tile) for the per-line stall picture
O=gpurun_out; mkdir -p $O
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:flda_estep -c 4 -o /tmp/s45_flda python tools/prof_run.py --config nsf_flda_k50 > $O/s45_full_flda.log 2>&1
tail -3 $O/s45_full_flda.log
ncu -i /tmp/s45_flda.ncu-rep --page raw --csv > $O/r2_full_nsf_flda_k50_raw.csv 2>/dev/null
ncu -i /tmp/s45_flda.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r2_full_nsf_flda_k50_source.csv.gz
ls -la $O/r2_full_nsf_flda_k50_*
