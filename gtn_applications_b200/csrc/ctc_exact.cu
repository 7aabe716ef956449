// Float64 log-semiring CTC forward/backward: the kernel that recomputes the utterances the
// scaled-probability kernels (ctc_chain.cu, ctc_pair.cuh, ctc_fast.cu) flag.  Same contract as
// the reference's create_ctc_graph -> intersect -> forward_score -> backward
// (criterions/ctc.py:15-29,40-51,78-81): scores[b] = log Z_b, gradE[b] = -grad_scale[b] * posterior.
//
// Why float64 storage: flagged utterances are the steep ones (emissions far from the target),
// where log alpha / log beta of the states that carry the posterior sit thousands of nats below
// the frame maximum; a float32 log value of magnitude 3000 has an ulp of 2.4e-4, more than the
// 1e-4 parity gate, no matter how the sums are organised.  alpha, beta and their history are
// therefore doubles; the transcendental work stays in float32, applied to DIFFERENCES that are
// formed exactly in double and lie in [-88, 0] (relative error 1e-7).
//
// One block per flagged utterance (others return at once), thread per state (strided), alpha
// ping-pong in shared memory, alpha history [T, S] in the workspace; per-frame posteriors are
// summed by label in a fixed-point shared tile (2^-30 units, integer atomics: order
// independent, so results are reproducible).
#include "common.cuh"
#include "launchers.h"

namespace wfst {
namespace exactk {

constexpr int kNT = 384;    // one state per thread up to S = 384 (cfg2: 353), two up to 768
constexpr float kFix = 1073741824.f;   // 2^30

struct Args {
  const float* E;
  const int* targets;
  const int* offsets;
  int B, T, C, blank;
  const float* grad_scale;
  float* scores;
  float* gradE;
  double* hist;      // [B][T][stride]
  int stride;
  const int* active;
};

__device__ __forceinline__ double lse3(double a0, double a1, double a2) {
  const double m = fmax(a0, fmax(a1, a2));
  if (m == -INFINITY) return -INFINITY;
  const float s = __expf((float)(a0 - m)) + __expf((float)(a1 - m)) + __expf((float)(a2 - m));
  return m + (double)__logf(s);
}

__global__ void __launch_bounds__(kNT) ctc_exact_kernel(Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  if (a.active && a.active[b] == 0) return;
  const int tid = threadIdx.x, T = a.T, C = a.C;
  const int* y = a.targets + a.offsets[b];
  const int L = a.offsets[b + 1] - a.offsets[b];
  const int S = 2 * L + 1;
  double* buf0 = reinterpret_cast<double*>(smem_raw);
  double* buf1 = buf0 + S;
  int* lab = reinterpret_cast<int*>(buf1 + S);
  unsigned* gt0 = reinterpret_cast<unsigned*>(lab + S);      // two tiles: one is cleared while the other is summed
  unsigned char* skip = reinterpret_cast<unsigned char*>(gt0 + 2 * C);
  __shared__ double zsh;
  for (int s = tid; s < S; s += kNT) {
    int l = a.blank, sk = 0;
    if (s & 1) {
      const int n = s >> 1;
      l = min(max(y[n], 0), C - 1);
      sk = (n >= 1 && y[n - 1] != y[n]) ? 1 : 0;
    }
    lab[s] = l;
    skip[s] = (unsigned char)sk;
  }
  __syncthreads();
  const float* Eb = a.E + (size_t)b * T * C;
  double* H = a.hist + (size_t)b * T * a.stride;
  float* G = a.gradE ? a.gradE + (size_t)b * T * C : nullptr;
  if (T == 0) {
    if (tid == 0) a.scores[b] = S == 1 ? 0.f : kNegInf;
    return;
  }
  // ---- alpha
  double* prev = buf0;
  double* cur = buf1;
  for (int s = tid; s < S; s += kNT) {
    const double v = s < 2 ? (double)Eb[lab[s]] : -INFINITY;
    prev[s] = v;
    H[s] = v;
  }
  __syncthreads();
  // one state per thread per round (S <= 2 kNT keeps the emissions of a thread's states in two
  // registers, fetched one frame ahead so that the L2 latency hides behind the frame before)
  const int s0 = tid, s1 = tid + kNT;
  const bool two = S <= 2 * kNT;
  float en0 = 0.f, en1 = 0.f;
  if (two && T > 1) {
    if (s0 < S) en0 = Eb[(size_t)C + lab[s0]];
    if (s1 < S) en1 = Eb[(size_t)C + lab[s1]];
  }
  for (int t = 1; t < T; ++t) {
    const float* Et = Eb + (size_t)t * C;
    double* Ht = H + (size_t)t * a.stride;
    if (two) {
      const float e0 = en0, e1 = en1;
      if (t + 1 < T) {
        if (s0 < S) en0 = Et[C + lab[s0]];
        if (s1 < S) en1 = Et[C + lab[s1]];
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int s = r ? s1 : s0;
        if (s < S) {
          const double a0 = prev[s];
          const double a1 = s >= 1 ? prev[s - 1] : -INFINITY;
          const double a2 = (s >= 2 && skip[s]) ? prev[s - 2] : -INFINITY;
          const double v = lse3(a0, a1, a2) + (double)(r ? e1 : e0);
          cur[s] = v;
          Ht[s] = v;
        }
      }
    } else {
      for (int s = tid; s < S; s += kNT) {
        const double a0 = prev[s];
        const double a1 = s >= 1 ? prev[s - 1] : -INFINITY;
        const double a2 = (s >= 2 && skip[s]) ? prev[s - 2] : -INFINITY;
        const double v = lse3(a0, a1, a2) + (double)Et[lab[s]];
        cur[s] = v;
        Ht[s] = v;
      }
    }
    __syncthreads();
    double* tmp = prev; prev = cur; cur = tmp;
  }
  if (tid == 0) {
    const double z = lse3(prev[S - 1], S > 1 ? prev[S - 2] : -INFINITY, -INFINITY);
    zsh = z;
    a.scores[b] = (float)z;
  }
  __syncthreads();
  const double Z = zsh;
  if (!G) return;
  if (!(Z > -INFINITY) || !(Z < INFINITY)) {   // infeasible (or +inf / NaN emissions): zero gradient
    for (size_t k = tid; k < (size_t)T * C; k += kNT) G[k] = 0.f;
    return;
  }
  // ---- beta (without the emission of its own frame), posteriors, gradient rows
  const float gsc = -(a.grad_scale ? a.grad_scale[b] : 1.f) / kFix;
  double* beta = prev;   // both buffers are free now
  double* be = cur;
  for (int s = tid; s < S; s += kNT) beta[s] = (s >= S - 2) ? 0.0 : -INFINITY;
  __syncthreads();
  float eb0 = 0.f, eb1 = 0.f;
  double hb0 = -INFINITY, hb1 = -INFINITY;
  if (two) {
    const float* Et = Eb + (size_t)(T - 1) * C;
    const double* Ht = H + (size_t)(T - 1) * a.stride;
    if (s0 < S) { eb0 = Et[lab[s0]]; hb0 = Ht[s0]; }
    if (s1 < S) { eb1 = Et[lab[s1]]; hb1 = Ht[s1]; }
  }
  for (int c = tid; c < 2 * C; c += kNT) gt0[c] = 0u;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    const float* Et = Eb + (size_t)t * C;
    const double* Ht = H + (size_t)t * a.stride;
    unsigned* gt = gt0 + ((T - 1 - t) & 1) * C;
    if (two) {
      const float e0 = eb0, e1 = eb1;
      const double h0 = hb0, h1 = hb1;
      if (t > 0) {      // next frame's emissions and alpha row
        if (s0 < S) { eb0 = Et[lab[s0] - C]; hb0 = Ht[s0 - a.stride]; }
        if (s1 < S) { eb1 = Et[lab[s1] - C]; hb1 = Ht[s1 - a.stride]; }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int s = r ? s1 : s0;
        if (s < S) {
          const double bs = beta[s];
          const double g = (r ? h1 : h0) + bs - Z;
          if (g > -80.0) {
            const float p = __expf((float)fmin(g, 0.0));
            atomicAdd(&gt[lab[s]], (unsigned)(p * kFix + 0.5f));
          }
          be[s] = bs + (double)(r ? e1 : e0);
        }
      }
    } else {
      for (int s = tid; s < S; s += kNT) {
        const double bs = beta[s];
        const double g = Ht[s] + bs - Z;
        if (g > -80.0) {
          const float p = __expf((float)fmin(g, 0.0));
          atomicAdd(&gt[lab[s]], (unsigned)(p * kFix + 0.5f));
        }
        be[s] = bs + (double)Et[lab[s]];
      }
    }
    __syncthreads();
    for (int c = tid; c < C; c += kNT) {
      G[(size_t)t * C + c] = gsc * (float)gt[c];
      gt[c] = 0u;     // summed into again two frames from now, after two more barriers
    }
    for (int s = tid; s < S; s += kNT) {
      const double b1 = s + 1 < S ? be[s + 1] : -INFINITY;
      const double b2 = (s + 2 < S && skip[s + 2]) ? be[s + 2] : -INFINITY;
      beta[s] = lse3(be[s], b1, b2);
    }
    __syncthreads();
  }
}

static size_t smem_bytes(int S, int C) { return (size_t)S * (2 * sizeof(double) + sizeof(int) + 1) + (size_t)2 * C * sizeof(unsigned) + 16; }

}  // namespace exactk

bool ctc_exact_eligible(int T, int C, int max_target_len) {
  (void)T;
  return exactk::smem_bytes(2 * max_target_len + 1, C) <= (size_t)200 * 1024;
}

size_t ctc_exact_hist_bytes(int B, int T, int max_target_len) {
  const int stride = (2 * max_target_len + 1 + 1) & ~1;
  return align_up((size_t)B * (T > 0 ? T : 1) * stride * sizeof(double), 256);
}

int launch_ctc_exact(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                     int blank, int max_target_len, const float* grad_scale, float* scores,
                     float* gradE, void* hist, const int* active, cudaStream_t st) {
  using namespace exactk;
  Args a{};
  a.E = E; a.targets = targets; a.offsets = offsets; a.B = B; a.T = T; a.C = C; a.blank = blank;
  a.grad_scale = grad_scale; a.scores = scores; a.gradE = gradE;
  a.hist = (double*)hist;
  a.stride = (2 * max_target_len + 1 + 1) & ~1;
  a.active = active;
  const size_t smem = smem_bytes(2 * max_target_len + 1, C);
  if (smem > 48 * 1024)
    WFST_CUDA_CHECK(cudaFuncSetAttribute(ctc_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_exact_kernel<<<B, kNT, smem, st>>>(a);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

}  // namespace wfst
