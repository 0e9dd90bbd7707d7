#!/bin/bash
# r2 evidence from one binary: GPU suite, launch lists and full ncu captures of every E-step kernel + the recs kernels
O=gpurun_out; mkdir -p $O
T=r2
md5sum topicmodelsvb.jl_b200/libtmvb.so > $O/${T}_lib_md5.txt
timeout 1500 python -m pytest tests -q -m gpu -s > $O/${T}_pytest.log 2>&1
tail -5 $O/${T}_pytest.log
export TMVB_GRAPH=0
for c in nsf_lda_k50 citeu_ctm_k30 citeu_ctpf_k30; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_$c.csv python tools/prof_run.py --config $c > $O/${T}_prof_$c.log 2>&1
done
TMVB_CFG4_M=200000 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_synth_lda_k200.csv python tools/prof_run.py --config synth_lda_k200 > $O/${T}_prof_synth_lda_k200.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_recs.csv python tools/prof_recs.py > $O/${T}_prof_recs.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_e2e_nsf.csv python tools/prof_run.py --config nsf_lda_k50 --e2e > $O/${T}_prof_e2e.log 2>&1
full() {  # name, kernel regex, extra ncu args, script, env...
  local name=$1; shift
  local rx=$1; shift
  local extra=$1; shift
  local script=$1; shift
  env "$@" ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$rx $extra -o /tmp/${T}_full_$name python $script > $O/${T}_full_$name.log 2>&1
  ncu -i /tmp/${T}_full_$name.ncu-rep --page raw --csv > $O/${T}_full_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${T}_full_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/${T}_full_${name}_source.csv.gz
  ls -la /tmp/${T}_full_$name.ncu-rep
}
full nsf_lda_k50 estep "" "tools/prof_run.py --config nsf_lda_k50" A=1
full citeu_ctm_k30 estep "-c 4" "tools/prof_run.py --config citeu_ctm_k30" A=1
full citeu_ctpf_k30 estep "-c 4" "tools/prof_run.py --config citeu_ctpf_k30" A=1
full synth_lda_k200 estep "-c 4" "tools/prof_run.py --config synth_lda_k200" TMVB_CFG4_M=100000
full recs "recs_scores|DeviceSegmentedSort" "-c 6" tools/prof_recs.py A=1
full mstep "normalize|colsum|alpha|elbo" "-c 6" "tools/prof_run.py --config nsf_lda_k50" A=1
du -sh $O; tail -3 $O/${T}_prof_*.log
