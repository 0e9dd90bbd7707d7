#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29801 tools/dev_e2e_profile_dist.py 2>&1 | grep -v "^$\|OMP_NUM\|\*\*\*" | head -40 > $O/s31_e2e_dist.log
cat $O/s31_e2e_dist.log
