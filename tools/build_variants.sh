#!/bin/bash
# Developer tool: A/B builds of the hybrid LDA kernel.  usage: tools/build_variants.sh tag1:"-DX=1 -DY=0" tag2:"..."
# -> topicmodelsvb.jl_b200/variants/libtmvb_<tag>.so (select with TMVB_SO=...); only the tmvb_lda_hyb_*.cu units are recompiled.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /dev/null
C=topicmodelsvb.jl_b200/csrc
mkdir -p topicmodelsvb.jl_b200/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for spec in "$@"; do
  tag=${spec%%:*}; defs=${spec#*:}
  (
    mkdir -p /tmp/var_$tag
    for f in $C/tmvb_lda_hyb_*.cu; do
      /usr/local/cuda/bin/nvcc $FLAGS $defs -c $f -o /tmp/var_$tag/$(basename ${f%.cu}).o &
    done
    wait
    others=$(ls $C/*.o | grep -v tmvb_lda_hyb_)
    /usr/local/cuda/bin/nvcc $FLAGS -shared $others /tmp/var_$tag/*.o -o topicmodelsvb.jl_b200/variants/libtmvb_$tag.so
    echo built $tag
  ) &
done
wait
ls -la topicmodelsvb.jl_b200/variants/
