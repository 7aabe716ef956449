// Dense ASG full-connect lattice for sm_100a: emissions x the bigram transition graph of
// ASGLossFunction.create_transitions_graph (criterions/asg.py:53-69), i.e. the denominator
// term forward_score(intersect(g_em, g_tr)) of asg.py:114 and its gtn.backward (asg.py:158).
// The composed lattice has T*C*C arcs (0.9 M per utterance at C=30, T=1000) but it is a
// dense layered graph: one frame is a matrix-vector product with the C x C transition
// matrix.  One warp per utterance, lane = label, the matrix row (and column) of the lane in
// registers, the previous frame's vector broadcast through 128 bytes of shared memory:
//   forward   a_t[i] = p_t[i] * sum_j W[i,j] a^_{t-1}[j],   a^_t = a_t / sum_i a_t[i]
//   backward  b_{t-1}[j] = sum_i W[i,j] p_t[i] b^_t[i],     b^ normalised the same way
// in the probability domain (W = exp(tr - max), p_t = exp(E_t - max_c E_t)); the removed
// scales are accumulated in float64, so Z is exact to float32 rounding of the per-frame sums.
// Posteriors are normalised per frame (every path crosses a frame / a frame boundary exactly
// once), which needs no global constant:
//   dZ/dE[t,i]   = a^_t[i] b^_t[i] / sum_i(.)
//   dZ/dtr[i,j]  = W[i,j] * sum_t a^_{t-1}[j] p_t[i] b^_t[i] / (c_t sum_i a^_t[i] b^_t[i])
// The forward vectors go to the caller's workspace ([T, C] per utterance, written and read
// once, coalesced).  Handles C <= 32 and T >= 1; other shapes use the generic lattice kernel.
#include <atomic>
#include "common.cuh"
#include "launchers.h"

namespace wfst {

extern int g_asg_dense_single;

namespace {
constexpr int kCP = 36;      // padded row length: a multiple of 4 whose quarter is odd
constexpr int kWarps = 4;    // utterances per block
constexpr int kPF = 4;       // frames of emissions / history prefetched ahead
constexpr unsigned kAll = 0xffffffffu;

__device__ __forceinline__ float dot_row(const float (&row)[kCP], const float* bc) {
  const float4* v = reinterpret_cast<const float4*>(bc);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int q = 0; q < kCP / 4; ++q) {
    const float4 a = v[q];   // same address in every lane: broadcast
    s0 = fmaf(row[4 * q + 0], a.x, s0);
    s1 = fmaf(row[4 * q + 1], a.y, s1);
    s2 = fmaf(row[4 * q + 2], a.z, s2);
    s3 = fmaf(row[4 * q + 3], a.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(32 * kWarps) asg_fcc_dense_kernel(
    const float* __restrict__ E, const float* __restrict__ tr, int B, int T, int C,
    const float* __restrict__ grad_scale, float sign, float* __restrict__ scores,
    float* __restrict__ gradE, int accumulate, float* __restrict__ gradTr, float* __restrict__ hist) {
  __shared__ __align__(16) float bcast[kWarps][2][kCP];
  __shared__ float red[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kWarps + warp;
  const bool valid = lane < C;

  // scale of the transition weights: max over the bigram part / over the start part
  if (threadIdx.x < 32) {
    float m = kNegInf, m0 = kNegInf;
    for (int k = lane; k < C * C; k += 32) m = fmaxf(m, tr[C + k]);
    for (int k = lane; k < C; k += 32) m0 = fmaxf(m0, tr[k]);
    m = warp_max(m);
    m0 = warp_max(m0);
    if (lane == 0) {
      red[0] = (m == kNegInf) ? 0.f : m;
      red[1] = (m0 == kNegInf) ? 0.f : m0;
    }
  }
  for (int k = threadIdx.x; k < kWarps * 2 * kCP; k += blockDim.x) (&bcast[0][0][0])[k] = 0.f;
  __syncthreads();
  const float wmax = red[0], w0max = red[1];
  if (b >= B) return;

  // row i of W (into label i from j) and column i of W (from label i into j), in registers
  float Wr[kCP], Wc[kCP];
#pragma unroll
  for (int j = 0; j < kCP; ++j) {
    Wr[j] = (valid && j < C) ? __expf(tr[C + lane * C + j] - wmax) : 0.f;
    Wc[j] = (valid && j < C) ? __expf(tr[C + j * C + lane] - wmax) : 0.f;
  }
  const float w0 = valid ? __expf(tr[lane] - w0max) : 0.f;

  const float* Eb = E + (size_t)b * T * C;
  float* hA = hist + (size_t)b * T * (C + 1);   // [T][C] normalised forward vectors
  float* hC = hA + (size_t)T * C;               // [T] their normalisers
  float* bc0 = bcast[warp][0];
  float* bc1 = bcast[warp][1];

  // ------------------------------------------------------------------ forward
  double logz = (double)w0max + (double)(T - 1) * (double)wmax;
  float ahat = 0.f;
  float xn[kPF];
#pragma unroll
  for (int k = 0; k < kPF; ++k) xn[k] = (valid && k < T) ? __ldg(Eb + (size_t)k * C + lane) : kNegInf;
  for (int t0 = 0; t0 < T; t0 += kPF) {
    float xc[kPF];
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      xc[k] = xn[k];
      const int tn = t0 + kPF + k;
      xn[k] = (valid && tn < T) ? __ldg(Eb + (size_t)tn * C + lane) : kNegInf;
    }
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      const int t = t0 + k;
      if (t >= T) break;
      const float m = warp_max(xc[k]);
      const float base = (m == kNegInf) ? 0.f : m;
      const float p = valid ? __expf(xc[k] - base) : 0.f;
      float a;
      if (t == 0) {
        a = p * w0;
      } else {
        bc0[lane] = ahat;
        __syncwarp();
        a = p * dot_row(Wr, bc0);
        __syncwarp();
      }
      // the normaliser only has to keep the vector in range: the maximum costs one CREDUX on the
      // dependency chain where a sum costs five shuffle steps (every formula below holds for
      // any positive c_t as long as a^_t = a_t / c_t)
      const float c = warp_max(a);
      ahat = c > 0.f ? __fdividef(a, c) : 0.f;
      logz += (double)base;            // the log c_t terms are summed after the sweep, T/32 per lane
      if (valid) hA[(size_t)t * C + lane] = ahat;
      if (lane == 0) hC[t] = c;
    }
  }
  {
    // Z = sum_t (base_t + log c_t) + log sum_i a^_{T-1}[i]
    __syncwarp();
    double lc = 0.0;
    for (int t = lane; t < T; t += 32) lc += (double)logf(hC[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lc += __shfl_xor_sync(kAll, lc, o);
    const float tail = warp_sum(ahat);
    logz += lc + (double)logf(tail);
  }
  if (lane == 0) scores[b] = (float)logz;

  if (gradE == nullptr && gradTr == nullptr) return;
  const float gs = sign * (grad_scale ? grad_scale[b] : 1.f);

  // ------------------------------------------------------------------ backward
  float acc[kCP];
#pragma unroll
  for (int j = 0; j < kCP; ++j) acc[j] = 0.f;
  float bhat = valid ? 1.f : 0.f;
  float acur = ahat;   // a^_{T-1}
  float xq[kPF], aq[kPF], cq[kPF];
  // queue entry k of a chunk starting at t0 (descending): x_t, a^_{t-1}, c_t for t = t0 - k
#pragma unroll
  for (int k = 0; k < kPF; ++k) {
    const int t = T - 1 - k;
    xq[k] = (valid && t >= 0) ? __ldg(Eb + (size_t)t * C + lane) : kNegInf;
    aq[k] = (valid && t >= 1) ? hA[(size_t)(t - 1) * C + lane] : 0.f;
    cq[k] = (t >= 0) ? hC[t] : 1.f;
  }
  float* gEb = gradE ? gradE + (size_t)b * T * C : nullptr;
  for (int t0 = T - 1; t0 >= 0; t0 -= kPF) {
    float xc[kPF], ac[kPF], cc[kPF];
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      xc[k] = xq[k]; ac[k] = aq[k]; cc[k] = cq[k];
      const int t = t0 - kPF - k;
      xq[k] = (valid && t >= 0) ? __ldg(Eb + (size_t)t * C + lane) : kNegInf;
      aq[k] = (valid && t >= 1) ? hA[(size_t)(t - 1) * C + lane] : 0.f;
      cq[k] = (t >= 0) ? hC[t] : 1.f;
    }
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      const int t = t0 - k;
      if (t < 0) break;
      const float m = warp_max(xc[k]);
      const float base = (m == kNegInf) ? 0.f : m;
      const float p = valid ? __expf(xc[k] - base) : 0.f;
      const float g = acur * bhat;
      const float G = warp_sum(g);
      const float gamma = G > 0.f ? __fdividef(g, G) : 0.f;
      if (gEb && valid) {
        float* dst = gEb + (size_t)t * C + lane;
        *dst = accumulate ? *dst + gs * gamma : gs * gamma;
      }
      if (t == 0) {
        // arcs 0 -> i+1 (asg.py:60-62): posterior of starting in label i
        if (gradTr && valid && gamma != 0.f) atomicAdd(&gradTr[lane], gs * gamma);
        break;
      }
      const float r = p * bhat;
      const float n = cc[k] * G;
      const float rn = n > 0.f ? __fdividef(r, n) : 0.f;
      bc0[lane] = ac[k];     // a^_{t-1}
      bc1[lane] = r;
      __syncwarp();
      {
        const float4* v = reinterpret_cast<const float4*>(bc0);
#pragma unroll
        for (int q = 0; q < kCP / 4; ++q) {
          const float4 a4 = v[q];
          acc[4 * q + 0] = fmaf(a4.x, rn, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(a4.y, rn, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(a4.z, rn, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(a4.w, rn, acc[4 * q + 3]);
        }
      }
      const float bn = dot_row(Wc, bc1);
      __syncwarp();
      const float nb = warp_max(bn);
      bhat = nb > 0.f ? __fdividef(bn, nb) : 0.f;
      acur = ac[k];
    }
  }
  if (gradTr && valid) {
    // arcs j+1 -> i+1 with weight index C + i*C + j (asg.py:63-67)
#pragma unroll
    for (int j = 0; j < kCP; ++j) {
      if (j < C) {
        const float v = gs * Wr[j] * acc[j];
        if (v != 0.f) atomicAdd(&gradTr[C + lane * C + j], v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Gradient version, one block of SIX warps per utterance.
//
// Two recurrence warps meet in the middle: A runs alpha up over [0, Th) while B runs beta down
// over [Th, T); they exchange a^_{Th-1} and b^_{Th-1} through shared memory (one named barrier);
// then A runs beta down over its half against its stored alpha vectors and B runs alpha up over
// its half against its stored beta vectors.  The kernel is a pure latency chain (a few warps
// per SM), so in that second phase a recurrence warp does NOTHING but the recurrence: per
// frame it leaves the vectors the gradient needs in a shared-memory ring (A: r_t = p_t b^_t and
// b^_t; B: a^_{t-1}, a^_t and c_t), eight frames per chunk, and two helper warps per direction
// take alternate frames of the previous chunk: posterior sum (five shuffle steps), gradient
// row, the rank-one update of the transition-gradient accumulators (32 FMAs against a
// broadcast vector).  Measured at cfg3: with that bookkeeping on the recurrence warp (the
// previous two-warp kernel) 0.45 ms alone and 0.67 ms next to the force-align kernel; a timing
// run with the bookkeeping deleted: 0.30 ms, ASG call 0.69 -> 0.45 ms.  One named barrier per
// chunk and direction (recurrence warp + its two helpers), double-buffered chunks.
// ---------------------------------------------------------------------------------------
constexpr int kChunk = 8;                 // frames per ring chunk
constexpr int kEnt = 2 * kCP;             // floats per ring entry: V0[kCP] (16-byte aligned, broadcast), V1[kCP]
constexpr int kGradWarps = 6;             // A, B, helpers of A (2), helpers of B (2)

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// named barriers 1, 2: a recurrence warp and its two helpers (immediate ids: the block reserves 5 barriers, not 16)
__device__ __forceinline__ void tick_bar(int role) {
  if (role == 0) asm volatile("bar.sync 1, 96;" ::: "memory");
  else asm volatile("bar.sync 2, 96;" ::: "memory");
}

__global__ void __launch_bounds__(32 * kGradWarps, 3) asg_fcc_dense_split_kernel(
    const float* __restrict__ E, const float* __restrict__ tr, int B, int T, int C,
    const float* __restrict__ grad_scale, float sign, float* __restrict__ scores,
    float* __restrict__ gradE, int accumulate, float* __restrict__ gradTr, float* __restrict__ hist) {
  __shared__ __align__(16) float bcast[2][2][kCP];            // recurrence warps, phase 1
  __shared__ __align__(16) float hb[4][kCP];                  // helpers' broadcast vectors
  __shared__ __align__(16) float ring[2][2][kChunk][kEnt];    // [direction][chunk parity][frame][entry]
  __shared__ __align__(16) float xch[2][kCP];
  __shared__ double zpart[2];
  __shared__ float red[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int role = warp < 2 ? warp : (warp - 2) >> 1;     // direction: 0 = A's half, 1 = B's half
  const bool helper = warp >= 2;
  const int hidx = (warp - 2) & 1;                          // which of the direction's two helpers
  const int b = blockIdx.x;
  const bool valid = lane < C;
  if (threadIdx.x < 32) {
    float m = kNegInf, m0 = kNegInf;
    for (int k = lane; k < C * C; k += 32) m = fmaxf(m, tr[C + k]);
    for (int k = lane; k < C; k += 32) m0 = fmaxf(m0, tr[k]);
    m = warp_max(m);
    m0 = warp_max(m0);
    if (lane == 0) {
      red[0] = (m == kNegInf) ? 0.f : m;
      red[1] = (m0 == kNegInf) ? 0.f : m0;
    }
  }
  for (int k = threadIdx.x; k < 2 * 2 * kCP; k += blockDim.x) (&bcast[0][0][0])[k] = 0.f;
  for (int k = threadIdx.x; k < 4 * kCP; k += blockDim.x) (&hb[0][0])[k] = 0.f;
  for (int k = threadIdx.x; k < 2 * 2 * kChunk * kEnt; k += blockDim.x) (&ring[0][0][0][0])[k] = 0.f;
  for (int k = threadIdx.x; k < 2 * kCP; k += blockDim.x) (&xch[0][0])[k] = 0.f;
  __syncthreads();
  const float wmax = red[0], w0max = red[1];

  const float* Eb = E + (size_t)b * T * C;
  float* hV = hist + (size_t)b * T * (C + 1);   // [T][C]: a^_t for t < Th (A), b^_t for t >= Th (B)
  float* hC = hV + (size_t)T * C;               // [T] normalisers c_t of the alpha vectors
  const int Th = T / 2;                          // >= 1 (launcher: T >= 2)
  const float gs = sign * (grad_scale ? grad_scale[b] : 1.f);
  float* gEb = gradE ? gradE + (size_t)b * T * C : nullptr;
  // phase 2 of direction d walks nfr frames in chunks of kChunk: A descends from Th-1, B ascends from Th
  const int nfr = role == 0 ? Th : T - Th;
  const int nch = (nfr + kChunk - 1) / kChunk;

  if (helper) {
    // ================================================================ helper warps
    float acc[kCP];
#pragma unroll
    for (int j = 0; j < kCP; ++j) acc[j] = 0.f;
    float* mybc = hb[warp - 2];
    constexpr int kH = kChunk / 2;              // frames of a chunk per helper
    // what a frame needs from global memory, fetched one chunk ahead:
    //   A's half: a^_t, a^_{t-1}, c_t;   B's half: b^_t, E_t (for p_t)
    float g0[kH], g1[kH], g2[kH];
    auto frame_of = [&](int n, int i) {         // i-th frame of this helper in chunk n (-1: none)
      const int k = n * kChunk + 2 * i + hidx;
      if (k >= nfr) return -1;
      return role == 0 ? Th - 1 - k : Th + k;
    };
    auto prefetch = [&](int n) {
#pragma unroll
      for (int i = 0; i < kH; ++i) {
        const int t = frame_of(n, i);
        if (role == 0) {
          g0[i] = (valid && t >= 0) ? hV[(size_t)t * C + lane] : 0.f;
          g1[i] = (valid && t >= 1) ? hV[(size_t)(t - 1) * C + lane] : 0.f;
          g2[i] = (t >= 0) ? hC[t] : 1.f;
        } else {
          g0[i] = (valid && t >= 0) ? hV[(size_t)t * C + lane] : 0.f;
          g1[i] = (valid && t >= 0) ? __ldg(Eb + (size_t)t * C + lane) : kNegInf;
          g2[i] = 0.f;
        }
      }
    };
    tick_bar(role);                     // the meeting is over: the stored vectors are visible
    prefetch(0);
    for (int n = 0; n < nch; ++n) {
      tick_bar(role);                   // chunk n is complete
      float c0[kH], c1[kH], c2[kH];
#pragma unroll
      for (int i = 0; i < kH; ++i) { c0[i] = g0[i]; c1[i] = g1[i]; c2[i] = g2[i]; }
      if (n + 1 < nch) prefetch(n + 1);
#pragma unroll
      for (int i = 0; i < kH; ++i) {
        const int t = frame_of(n, i);
        if (t < 0) break;
        const float* ent = ring[role][n & 1][2 * i + hidx];
        float gq, rnum, cden;
        const float* avec;                      // a^_{t-1}, broadcastable
        if (role == 0) {
          // V0 = r_t = p_t b^_t, V1 = b^_t
          gq = c0[i] * ent[kCP + lane];
          rnum = ent[lane];
          cden = c2[i];
          mybc[lane] = c1[i];
          avec = mybc;
        } else {
          // V0 = a^_{t-1}, V1 = a^_t, V1[32] = c_t
          const float m = warp_max(c1[i]);
          const float base = (m == kNegInf) ? 0.f : m;
          const float p = valid ? __expf(c1[i] - base) : 0.f;
          gq = ent[kCP + lane] * c0[i];
          rnum = p * c0[i];
          cden = ent[kCP + 32];
          avec = ent;
        }
        const float G = warp_sum(gq);
        const float gamma = G > 0.f ? __fdividef(gq, G) : 0.f;
        if (gEb && valid) {
          float* dst = gEb + (size_t)t * C + lane;
          *dst = accumulate ? *dst + gs * gamma : gs * gamma;
        }
        if (t == 0) {                           // only in A's half: the start arcs
          if (gradTr && valid && gamma != 0.f) atomicAdd(&gradTr[lane], gs * gamma);
          break;
        }
        const float nn = cden * G;
        const float rn = nn > 0.f ? __fdividef(rnum, nn) : 0.f;
        __syncwarp();
        {
          const float4* v = reinterpret_cast<const float4*>(avec);
#pragma unroll
          for (int q = 0; q < kCP / 4; ++q) {
            const float4 a4 = v[q];
            acc[4 * q + 0] = fmaf(a4.x, rn, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(a4.y, rn, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(a4.z, rn, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(a4.w, rn, acc[4 * q + 3]);
          }
        }
        __syncwarp();
      }
    }
    if (gradTr && valid) {
#pragma unroll
      for (int j = 0; j < kCP; ++j) {
        if (j < C) {
          const float v = gs * __expf(tr[C + lane * C + j] - wmax) * acc[j];
          if (v != 0.f) atomicAdd(&gradTr[C + lane * C + j], v);
        }
      }
    }
    return;
  }

  // ================================================================== recurrence warps
  // the lane's row of W for the alpha recursions, its column for the beta recursions: a warp
  // needs one of them per phase, so ONE register array is refilled at the meeting (the block
  // must leave registers for the force-align blocks that share the SM: 3 blocks per SM)
  float Wx[kCP];
  auto load_row = [&]() {
#pragma unroll
    for (int j = 0; j < kCP; ++j) Wx[j] = (valid && j < C) ? __expf(tr[C + lane * C + j] - wmax) : 0.f;
  };
  auto load_col = [&]() {
#pragma unroll
    for (int j = 0; j < kCP; ++j) Wx[j] = (valid && j < C) ? __expf(tr[C + j * C + lane] - wmax) : 0.f;
  };
  const float w0 = valid ? __expf(tr[lane] - w0max) : 0.f;
  float* bc0 = bcast[warp][0];
  float* bc1 = bcast[warp][1];
  double logz = 0.0;

  if (role == 0) {
    // ------------------------------------------------------------ A: alpha up over [0, Th)
    logz = (double)w0max + (double)(Th - 1) * (double)wmax;
    load_row();
    float ahat = 0.f;
    float xn[kPF];
#pragma unroll
    for (int k = 0; k < kPF; ++k) xn[k] = (valid && k < Th) ? __ldg(Eb + (size_t)k * C + lane) : kNegInf;
    for (int t0 = 0; t0 < Th; t0 += kPF) {
      float xc[kPF];
#pragma unroll
      for (int k = 0; k < kPF; ++k) {
        xc[k] = xn[k];
        const int tn = t0 + kPF + k;
        xn[k] = (valid && tn < Th) ? __ldg(Eb + (size_t)tn * C + lane) : kNegInf;
      }
#pragma unroll
      for (int k = 0; k < kPF; ++k) {
        const int t = t0 + k;
        if (t >= Th) break;
        const float m = warp_max(xc[k]);
        const float base = (m == kNegInf) ? 0.f : m;
        const float p = valid ? __expf(xc[k] - base) : 0.f;
        float av;
        if (t == 0) {
          av = p * w0;
        } else {
          bc0[lane] = ahat;
          __syncwarp();
          av = p * dot_row(Wx, bc0);
          __syncwarp();
        }
        const float c = warp_max(av);
        ahat = c > 0.f ? __fdividef(av, c) : 0.f;
        logz += (double)base;
        if (valid) hV[(size_t)t * C + lane] = ahat;
        if (lane == 0) hC[t] = c;
      }
    }
    xch[0][lane] = ahat;                         // a^_{Th-1} for B
    asm volatile("bar.sync 3, 64;" ::: "memory");
    float bhat = xch[1][lane];                   // b^_{Th-1} from B
    {
      __syncwarp();
      double lc = 0.0;
      for (int t = lane; t < Th; t += 32) lc += (double)logf(hC[t]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lc += __shfl_xor_sync(kAll, lc, o);
      logz += lc;
    }
    // ------------------------------------------------------------ A: beta down over [0, Th)
    load_col();
    tick_bar(role);                      // my stored vectors are visible to my helpers
    float xq[kPF];
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      const int t = Th - 1 - k;
      xq[k] = (valid && t >= 0) ? __ldg(Eb + (size_t)t * C + lane) : kNegInf;
    }
    for (int n = 0; n < nch; ++n) {
#pragma unroll
      for (int h = 0; h < kChunk / kPF; ++h) {
        float xc[kPF];
#pragma unroll
        for (int k = 0; k < kPF; ++k) {
          xc[k] = xq[k];
          const int t = Th - 1 - (n * kChunk + h * kPF + kPF + k);
          xq[k] = (valid && t >= 0) ? __ldg(Eb + (size_t)t * C + lane) : kNegInf;
        }
#pragma unroll
        for (int k = 0; k < kPF; ++k) {
          const int t = Th - 1 - (n * kChunk + h * kPF + k);
          if (t < 0) break;
          float* ent = ring[0][n & 1][h * kPF + k];
          const float m = warp_max(xc[k]);
          const float base = (m == kNegInf) ? 0.f : m;
          const float p = valid ? __expf(xc[k] - base) : 0.f;
          const float r = p * bhat;
          ent[lane] = r;
          ent[kCP + lane] = bhat;
          if (t == 0) break;
          __syncwarp();
          const float bn = dot_row(Wx, ent);
          const float nb = warp_max(bn);
          bhat = nb > 0.f ? __fdividef(bn, nb) : 0.f;
        }
      }
      tick_bar(role);                    // chunk n is complete
    }
  } else {
    // ------------------------------------------------------------ B: beta down over [Th, T)
    load_col();
    float bhat = valid ? 1.f : 0.f;
    float xq[kPF];
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      const int t = T - 1 - k;
      xq[k] = (valid && t >= Th) ? __ldg(Eb + (size_t)t * C + lane) : kNegInf;
    }
    for (int t0 = T - 1; t0 >= Th; t0 -= kPF) {
      float xc[kPF];
#pragma unroll
      for (int k = 0; k < kPF; ++k) {
        xc[k] = xq[k];
        const int t = t0 - kPF - k;
        xq[k] = (valid && t >= Th) ? __ldg(Eb + (size_t)t * C + lane) : kNegInf;
      }
#pragma unroll
      for (int k = 0; k < kPF; ++k) {
        const int t = t0 - k;
        if (t < Th) break;
        const float m = warp_max(xc[k]);
        const float base = (m == kNegInf) ? 0.f : m;
        const float p = valid ? __expf(xc[k] - base) : 0.f;
        if (valid) hV[(size_t)t * C + lane] = bhat;          // b^_t
        bc1[lane] = p * bhat;
        __syncwarp();
        const float bn = dot_row(Wx, bc1);
        __syncwarp();
        const float nb = warp_max(bn);
        bhat = nb > 0.f ? __fdividef(bn, nb) : 0.f;           // b^_{t-1}
      }
    }
    xch[1][lane] = bhat;                         // b^_{Th-1} for A
    asm volatile("bar.sync 3, 64;" ::: "memory");
    float ahat = xch[0][lane];                   // a^_{Th-1} from A
    // ------------------------------------------------------------ B: alpha up over [Th, T)
    load_row();
    tick_bar(role);                      // my stored vectors are visible to my helpers
    logz = (double)(T - Th) * (double)wmax;
    float xn[kPF];
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      const int t = Th + k;
      xn[k] = (valid && t < T) ? __ldg(Eb + (size_t)t * C + lane) : kNegInf;
    }
    for (int n = 0; n < nch; ++n) {
#pragma unroll
      for (int h = 0; h < kChunk / kPF; ++h) {
        float xc[kPF];
#pragma unroll
        for (int k = 0; k < kPF; ++k) {
          xc[k] = xn[k];
          const int tn = Th + n * kChunk + h * kPF + kPF + k;
          xn[k] = (valid && tn < T) ? __ldg(Eb + (size_t)tn * C + lane) : kNegInf;
        }
#pragma unroll
        for (int k = 0; k < kPF; ++k) {
          const int t = Th + n * kChunk + h * kPF + k;
          if (t >= T) break;
          float* ent = ring[1][n & 1][h * kPF + k];
          const float m = warp_max(xc[k]);
          const float base = (m == kNegInf) ? 0.f : m;
          const float p = valid ? __expf(xc[k] - base) : 0.f;
          ent[lane] = ahat;                                  // a^_{t-1}
          __syncwarp();
          const float av = p * dot_row(Wx, ent);
          const float c = warp_max(av);
          const float anew = c > 0.f ? __fdividef(av, c) : 0.f;
          ent[kCP + lane] = anew;
          if (lane == 0) {
            ent[kCP + 32] = c;
            hC[t] = c;
          }
          ahat = anew;
          logz += (double)base;
        }
      }
      tick_bar(role);                    // chunk n is complete
    }
    {
      __syncwarp();
      double lc = 0.0;
      for (int t = Th + lane; t < T; t += 32) lc += (double)logf(hC[t]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lc += __shfl_xor_sync(kAll, lc, o);
      const float tail = warp_sum(ahat);
      logz += lc + (double)logf(tail);
    }
  }
  if (lane == 0) zpart[role] = logz;
  asm volatile("bar.sync 4, 64;" ::: "memory");
  if (role == 0 && lane == 0) scores[b] = (float)(zpart[0] + zpart[1]);
}

// ---------------------------------------------------------------------------------------
// Best path through emissions x the same bigram graph (ASG.viterbi, criterions/asg.py:217-226:
// gtn.viterbi_path(gtn.intersect(g_em, g_tr))).  One warp per utterance, lane = label; the
// lane's transition row in registers, the previous frame's scores broadcast through shared
// memory, one byte of back-pointer per (frame, label) in shared memory, backtrace by lane 0.
// Ties go to the lowest previous label / lowest final label: the order of the in-lists of
// create_transitions_graph (asg.py:60-67) under GTN's strict '>' (see viterbi.cu).
__global__ void __launch_bounds__(32) asg_viterbi_dense_kernel(
    const float* __restrict__ E, const float* __restrict__ tr, int T, int C,
    float* __restrict__ scores, int32_t* __restrict__ labels) {
  extern __shared__ __align__(16) unsigned char vsm[];
  float* bc = reinterpret_cast<float*>(vsm);            // [kCP] scores of the previous frame
  unsigned char* bp = vsm + kCP * sizeof(float);         // [T][32]
  const int lane = threadIdx.x, b = blockIdx.x;
  const bool valid = lane < C;
  float Wr[kCP];
#pragma unroll
  for (int j = 0; j < kCP; ++j) Wr[j] = (valid && j < C) ? tr[C + lane * C + j] : kNegInf;
  for (int k = lane; k < kCP; k += 32) bc[k] = kNegInf;
  const float* Eb = E + (size_t)b * T * C;
  float xn[kPF];
#pragma unroll
  for (int k = 0; k < kPF; ++k) xn[k] = (valid && k < T) ? __ldg(Eb + (size_t)k * C + lane) : kNegInf;
  float delta = kNegInf;
  for (int t0 = 0; t0 < T; t0 += kPF) {
    float xc[kPF];
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      xc[k] = xn[k];
      const int tn = t0 + kPF + k;
      xn[k] = (valid && tn < T) ? __ldg(Eb + (size_t)tn * C + lane) : kNegInf;
    }
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      const int t = t0 + k;
      if (t >= T) break;
      float best = kNegInf;
      int arg = 255;
      if (t == 0) {
        best = valid ? tr[lane] : kNegInf;               // arcs 0 -> i+1
      } else {
        __syncwarp();
        bc[lane] = delta;
        __syncwarp();
        const float4* v = reinterpret_cast<const float4*>(bc);
#pragma unroll
        for (int q = 0; q < kCP / 4; ++q) {
          const float4 d = v[q];
          float c;
          c = d.x + Wr[4 * q + 0]; if (c > best) { best = c; arg = 4 * q + 0; }
          c = d.y + Wr[4 * q + 1]; if (c > best) { best = c; arg = 4 * q + 1; }
          c = d.z + Wr[4 * q + 2]; if (c > best) { best = c; arg = 4 * q + 2; }
          c = d.w + Wr[4 * q + 3]; if (c > best) { best = c; arg = 4 * q + 3; }
        }
      }
      delta = valid ? best + xc[k] : kNegInf;
      bp[(size_t)t * 32 + lane] = (unsigned char)arg;
    }
  }
  // best final label, lowest index on ties
  float m = warp_max(delta);
  const unsigned hit = __ballot_sync(kAll, valid && delta == m && m != kNegInf);
  int cur = hit ? __ffs(hit) - 1 : -1;
  if (lane == 0) scores[b] = hit ? m : kNegInf;
  __syncwarp();
  if (lane == 0) {
    for (int t = T - 1; t >= 0; --t) {
      const int prev = (cur >= 0) ? bp[(size_t)t * 32 + cur] : 255;
      bp[(size_t)t * 32] = (unsigned char)(cur >= 0 ? cur : 255);   // row t is done: reuse its first byte
      cur = (t > 0 && prev != 255) ? prev : ((t > 0) ? -1 : cur);
    }
  }
  __syncwarp();
  int32_t* lb = labels + (size_t)b * T;
  for (int t = lane; t < T; t += 32) {
    const int v = bp[(size_t)t * 32];
    lb[t] = (v == 255) ? -1 : v;
  }
}
}  // namespace

bool asg_viterbi_dense_eligible(int T, int C) {
  return T >= 1 && C >= 1 && C <= 32 && (size_t)T * 32 + kCP * sizeof(float) <= 200 * 1024;
}

int launch_asg_viterbi_dense(const float* E, const float* tr, int B, int T, int C, float* scores,
                             int32_t* labels, cudaStream_t st) {
  const size_t smem = (size_t)T * 32 + kCP * sizeof(float);
  WFST_CUDA_CHECK(cudaFuncSetAttribute(asg_viterbi_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  asg_viterbi_dense_kernel<<<B, 32, smem, st>>>(E, tr, T, C, scores, labels);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

int g_asg_dense_single = 0;   // test hook (wfst_debug_force_generic_ctc(3)): one warp per utterance only
bool asg_fcc_dense_eligible(int T, int C) { return T >= 1 && C >= 1 && C <= 32; }

int launch_asg_fcc_dense(const float* E, const float* tr, int B, int T, int C, const float* grad_scale,
                         float sign, float* scores, float* gradE, int accumulate, float* gradTr,
                         float* hist, cudaStream_t st) {
  // both kernels ask for the largest shared-memory carveout although they need little: an SM
  // configured for them can then also host the lattice blocks that run next to them
  // (function attributes are per device: remember which devices have been configured)
  static std::atomic<unsigned long long> carveout_set{0};
  int devi = 0;
  cudaGetDevice(&devi);
  const unsigned long long bit = 1ull << (devi & 63);
  if (!(carveout_set.load() & bit)) {
    cudaFuncSetAttribute(asg_fcc_dense_split_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(asg_fcc_dense_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    carveout_set.fetch_or(bit);
  }
  // six warps per utterance (two that meet in the middle + their helpers) when gradients are wanted and T allows
  if (T >= 8 && (gradE || gradTr) && !g_asg_dense_single)
    asg_fcc_dense_split_kernel<<<B, 32 * kGradWarps, 0, st>>>(
        E, tr, B, T, C, grad_scale, sign, scores, gradE, accumulate, gradTr, hist);
  else
    asg_fcc_dense_kernel<<<(B + kWarps - 1) / kWarps, 32 * kWarps, 0, st>>>(
        E, tr, B, T, C, grad_scale, sign, scores, gradE, accumulate, gradTr, hist);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

}  // namespace wfst
