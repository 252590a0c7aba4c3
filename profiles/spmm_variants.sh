#!/bin/bash
# A/B builds of libb200rec.so that differ only in spmm.cu's compile-time knobs (selected at run time with B200REC_LIB):
#   profiles/spmm_variants.sh   ->  profiles/bin/libb200rec_<name>.so
set -e
cd "$(dirname "$0")/../inductive-recommendation_b200/csrc"
make -j8 > /dev/null
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
OTHERS="abi.o plan.o spmm_stream.o bpr.o score.o score_tc.o p2p.o infonce.o"
build() {  # name, defines...
  name=$1; shift
  nvcc $FLAGS "$@" -c spmm.cu -o ../../profiles/bin/spmm_$name.o 2> ../../profiles/bin/spmm_$name.ptxas.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../profiles/bin/libb200rec_$name.so ../../profiles/bin/spmm_$name.o $OTHERS -lcudart
  echo "$name: $(grep -A2 'spmm_items_kernelILi16ELi1ELb1ELb0ELb0ELb0ELb0ELb0' ../../profiles/bin/spmm_$name.ptxas.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
}
build base
build minb3 -DB200REC_SPMM_MINB=3
build eb2 -DB200REC_SPMM_EBMUL=2
build eb2_minb3 -DB200REC_SPMM_EBMUL=2 -DB200REC_SPMM_MINB=3
build pf -DB200REC_SPMM_PREFETCH=1
build pf_minb3 -DB200REC_SPMM_PREFETCH=1 -DB200REC_SPMM_MINB=3
build pf_eb2_minb3 -DB200REC_SPMM_PREFETCH=1 -DB200REC_SPMM_EBMUL=2 -DB200REC_SPMM_MINB=3
