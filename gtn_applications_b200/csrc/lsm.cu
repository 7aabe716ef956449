// log_softmax over the class axis and its backward, restricted to the utterances an
// `active` mask selects.  Used by the fused logits -> CTC entry point for the (rare)
// utterances the scaled-probability kernel hands to the log-semiring kernel: those need
// materialised log-probabilities (criterions/ctc.py:107) and the chain rule back to logits.
#include "common.cuh"
#include "launchers.h"

namespace wfst {

// one block per utterance (the unselected ones leave at once), one warp per frame
__global__ void lsm_rows_kernel(const float* x, const int* active, int T, int C, float* out) {
  const int b = blockIdx.x;
  if (active && active[b] == 0) return;
  const int lane = threadIdx.x & 31;
  for (int t = threadIdx.x >> 5; t < T; t += blockDim.x >> 5) {
    const float* row = x + ((size_t)b * T + t) * C;
    float* o = out + ((size_t)b * T + t) * C;
    float m = kNegInf;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    const float base = (m == kNegInf) ? 0.f : m;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __expf(row[c] - base);
    s = warp_sum(s);
    const float lse = base + logf(s);
    for (int c = lane; c < C; c += 32) o[c] = row[c] - lse;
  }
}

// out = g - exp(lsm) * sum_c g   (d/d logits of a function of log_softmax(logits))
__global__ void lsm_backward_kernel(const float* lsm, const float* g, const int* active, int T, int C,
                                    float* out) {
  const int b = blockIdx.x;
  if (active && active[b] == 0) return;
  const int lane = threadIdx.x & 31;
  for (int t = threadIdx.x >> 5; t < T; t += blockDim.x >> 5) {
    const size_t base = ((size_t)b * T + t) * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += g[base + c];
    s = warp_sum(s);
    for (int c = lane; c < C; c += 32) out[base + c] = g[base + c] - __expf(lsm[base + c]) * s;
  }
}

int launch_lsm_rows(const float* x, const int* active, int B, int T, int C, float* out, cudaStream_t st) {
  if (B <= 0 || T <= 0) return WFST_OK;
  lsm_rows_kernel<<<B, 512, 0, st>>>(x, active, T, C, out);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

int launch_lsm_backward(const float* lsm, const float* g, const int* active, int B, int T, int C,
                        float* out, cudaStream_t st) {
  if (B <= 0 || T <= 0) return WFST_OK;
  lsm_backward_kernel<<<B, 512, 0, st>>>(lsm, g, active, T, C, out);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

}  // namespace wfst
