// Paired fast CTC kernel, translation unit for p-tile row stride 64 (0 = host dispatch).
#define WFST_PAIR_CS 64
#include "ctc_pair.cuh"
