#!/bin/bash
# compute-sanitizer over the kernels added this round (small cases): memcheck, then racecheck of the shared-memory exchanges
O=gpurun_out; mkdir -p $O
export TMVB_GRAPH=0
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_flda_gpu.py tests/test_ctpf_gpu.py -q -m gpu -x -k "ragged or trajectory_small or recs_small or recs_ties or filtered_predict" > $O/s32_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" $O/s32_memcheck.log; tail -4 $O/s32_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_flda_gpu.py -q -m gpu -x -k "flda_ragged or test_flda_elbo_trajectory_small or test_fctm_elbo_trajectory_small" > $O/s32_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -c "hazard" $O/s32_racecheck.log; tail -4 $O/s32_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_lda_gpu.py tests/test_ctm_gpu.py -q -m gpu -x -k "fresh or ragged or argument_errors" > $O/s32_memcheck2.log 2>&1
echo "memcheck2 rc=$?"; tail -3 $O/s32_memcheck2.log
