#!/bin/bash
# Builds gtn_applications_b200/lib/libwfst_b200_prof.so: the library with -DWFST_PROFILE
# (per-role clock64 accounting printed by block 0).  Not used by the product or the tests.
set -e
cd "$(dirname "$0")/../gtn_applications_b200/csrc"
mkdir -p /tmp/wfst_prof
for f in capi lattice lattice_pair ctc_fast ctc_chain ctc_solo ctc_exact ctc_pair_cs0 ctc_pair_cs32 ctc_pair_cs64 ctc_pair_cs128 lsm asg_dense viterbi; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -DWFST_PROFILE -c $f.cu -o /tmp/wfst_prof/$f.o &
done
wait
g++ -O2 -std=c++17 -fPIC -fvisibility=hidden -pthread -c graph.cpp -o /tmp/wfst_prof/graph.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libwfst_b200_prof.so /tmp/wfst_prof/*.o -L/usr/lib/gcc/x86_64-linux-gnu/13 -Xlinker --no-as-needed -lstdc++ -lpthread
