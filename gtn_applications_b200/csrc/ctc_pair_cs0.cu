// Paired fast CTC kernel, translation unit for p-tile row stride 0 (0 = host dispatch).
#define WFST_PAIR_CS 0
#include "ctc_pair.cuh"
