#!/bin/bash
# A/B builds of libb200rec.so that differ in one translation unit's compile-time knobs (selected at run time with
# B200REC_LIB=profiles/bin/libb200rec_<name>.so):   profiles/ab_build.sh score_tc name1 "-DX=1" name2 "-DX=2 -DY=3" ...
set -e
unit=$1; shift
cd "$(dirname "$0")/../inductive-recommendation_b200/csrc"
make -j8 > /dev/null
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
OTHERS=$(ls *.o | grep -v "^$unit.o$" | tr '\n' ' ')
mkdir -p ../../profiles/bin
while [ $# -gt 0 ]; do
  name=$1; defs=$2; shift 2
  nvcc $FLAGS $defs -c $unit.cu -o ../../profiles/bin/${unit}_$name.o 2> ../../profiles/bin/${unit}_$name.ptxas.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../profiles/bin/libb200rec_$name.so ../../profiles/bin/${unit}_$name.o $OTHERS -lcudart
  echo "built $name ($defs)"
done
