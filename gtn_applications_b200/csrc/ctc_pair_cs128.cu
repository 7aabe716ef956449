// Paired fast CTC kernel, translation unit for p-tile row stride 128 (0 = host dispatch).
#define WFST_PAIR_CS 128
#include "ctc_pair.cuh"
