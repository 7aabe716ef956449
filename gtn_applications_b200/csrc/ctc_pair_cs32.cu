// Paired fast CTC kernel, translation unit for p-tile row stride 32 (0 = host dispatch).
#define WFST_PAIR_CS 32
#include "ctc_pair.cuh"
