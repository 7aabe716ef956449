#!/bin/bash
# Builds gtn_applications_b200/lib/libwfst_b200_prof.so: the product library with ONE translation
# unit recompiled with -DWFST_PROFILE (per-role clock accounting printed by block 0).
# usage: tools/build_prof_lib.sh ctc_tick.cu ; then run with WFST_B200_LIB=<that .so>
set -e
cd "$(dirname "$0")/../gtn_applications_b200/csrc"
src=${1:-ctc_tick.cu}; obj=${src%.cu}.o
make -j8 >/dev/null
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -DWFST_PROFILE $EXTRA -c $src -o /tmp/prof_$obj
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../lib/libwfst_b200_prof.so \
  $(ls *.o | grep -v "^$obj$") /tmp/prof_$obj -L/usr/lib/gcc/x86_64-linux-gnu/13 -Xlinker --no-as-needed -lstdc++ -lpthread
echo built ../lib/libwfst_b200_prof.so
