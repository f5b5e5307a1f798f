#!/bin/bash
# Diagnostic builds of the library with parts of sa_pipe.cu switched off (-DSAP_X=n; results are WRONG, timing only):
# 2 one MMA k step of eight, 3 gather without global loads, 4 epilogues without smem stores, 5 max without global
# stores, 6 epilogues without the TMEM load.  Usage: tools/sap_variants.sh 2 3 4 ; then on the GPU box
#   DEMF_B200_LIB=demf_b200/_variants/libdemf_x3.so python tools/sa_pipe_bench.py
set -e
cd "$(dirname "$0")/.."
python -m demf_b200.build >/dev/null
mkdir -p demf_b200/_variants
objs=$(ls demf_b200/csrc/build/*.o | grep -v sa_pipe.o)
for x in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DSAP_X=$x \
       -c demf_b200/csrc/sa_pipe.cu -o demf_b200/_variants/sa_pipe_x$x.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o demf_b200/_variants/libdemf_x$x.so $objs demf_b200/_variants/sa_pipe_x$x.o
done
ls -la demf_b200/_variants/*.so
