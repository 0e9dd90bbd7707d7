#!/bin/bash
# GPU call: bench of all four configurations + launch lists + full ncu captures of the four E-step kernels.
# The .ncu-rep files are reduced to CSV pages on the box (gpurun_out/ may carry at most 64 MiB back).
set -x
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-s1}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/${T}_clocks.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
tail -c 600 $O/${T}_bench.err
kill $SMI
export TMVB_GRAPH=0
for c in nsf_lda_k50 citeu_ctm_k30 citeu_ctpf_k30; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_$c.csv python tools/prof_run.py --config $c > $O/${T}_prof_$c.log 2>&1
done
TMVB_CFG4_M=200000 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_synth_lda_k200.csv python tools/prof_run.py --config synth_lda_k200 > $O/${T}_prof_synth_lda_k200.log 2>&1
full() {  # name, extra ncu args, env...
  local name=$1; shift
  local extra=$1; shift
  env "$@" ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:estep $extra -o /tmp/${T}_full_$name python tools/prof_run.py --config $name > $O/${T}_full_$name.log 2>&1
  ncu -i /tmp/${T}_full_$name.ncu-rep --page raw --csv > $O/${T}_full_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${T}_full_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/${T}_full_${name}_source.csv.gz
  ls -la /tmp/${T}_full_$name.ncu-rep
}
full nsf_lda_k50 "" A=1
full citeu_ctm_k30 "-c 4" A=1
full citeu_ctpf_k30 "-c 4" A=1
full synth_lda_k200 "-c 3" TMVB_CFG4_M=100000
du -sh $O; ls -la $O | tail -30
