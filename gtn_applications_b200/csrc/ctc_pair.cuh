// Paired fast CTC forward+backward for sm_100a.  Replaces, for CTC, the reference's
// per-utterance create_ctc_graph -> intersect -> forward_score -> backward
// (criterions/ctc.py:15-29,40-51,78-81).
//
// One thread block handles TWO utterances at once.  Every value of the recursion is a
// packed pair of floats (x = first utterance, y = second) and all arithmetic on it is a
// packed FP32 instruction (add/mul/fma.f32x2 -> SASS FADD2 / FMUL2 / FFMA2), which halves
// the issue slots the dependent per-frame chain needs.  Twelve warps per block:
//   L0 "live alpha"  time ascending,  index j = s
//   L1 "live beta"   time descending, index j = Sp-2-s   (mirror that keeps the parity of s)
//   RC<d>.0, RC<d>.1 recompute the OPPOSITE recursion of L<d> over alternate segments from a
//                    checkpoint and multiply it, frame by frame, with what L<d> stored
//   X<d>.<u>         sum a segment's per-state posteriors of utterance u over states with
//                    equal label and send the [8, C] gradient tile to HBM (bulk async store)
//   P<d>             producer of direction d: TMA-loads [8, C] emission tiles of both
//                    utterances and turns them into p[t,c] = exp(E[t,c] - max_c E[t,c]) tiles
//                    (one plane per utterance, rows in the order L<d> steps through them)
// All recursion warps run the SAME code: label-type states sit at odd slots in both
// orientations, the direction is a runtime value.  The recursion is
//   v'[j] = (v[j] + v[j-1] + skip[j] * v[j-2]) * p_t[lab j]
// (the beta recursion written for beta~_t(s) = p_t(lab s) * beta_t(s) is the alpha recursion
// on the reversed target and reversed time).  Lane l owns K consecutive slots; the left
// neighbour's last value arrives by one shuffle per utterance and frame.  Values are float32
// mantissas with one power-of-two exponent per lane and utterance, renormalised every 16
// frames ("event").
//
// Schedule (meet in the middle + recompute; nothing of size T x S leaves the SM):
//   phase 1: L0 sweeps segments [0, nA), L1 sweeps [nA, nseg) downwards; each writes one
//            checkpoint (K pairs + 2 exponents per lane) per 8-frame segment.
//   meeting: Z = sum_s alpha(s) * (successor sum of beta~)(s) at the boundary.
//   phase 2: L0 continues upwards through [nA, nseg) and stores the pre-emission sums
//            ("abar") of every frame of a segment into a shared-memory ring.  An RC0 warp
//            then re-runs the beta recursion over that segment from L1's checkpoint, rescaled
//            once per segment so that  w * abar = posterior * Zm  with no further factor,
//            writes the label-state products back IN PLACE and the blank partial sums
//            per lane next to them.  X0.u gathers the products of utterance u in
//            label-sorted order (4 positions per lane, segmented suffix sums by shuffles)
//            and stores the gradient tile.  L1 / RC1 / X1 mirror this downwards through [0, nA).
//   fused mode (Args::fused): E holds raw logits; the producers also form the softmax
//            denominators, store the softmax term of the gradient themselves, and X adds
//            the posterior term with a bulk reduce (see produce_range / role_reduce).
//
// Robustness: a frame's posteriors sum to one.  Every segment's total is checked against
// rows * Z (2e-5, finite); a violation (float32 range exceeded inside a window, which can
// only lose mass or produce inf/NaN), Z out of range, a blank label inside the target or a
// label layout that does not fit flag the utterance in `hazard`, and the log-semiring
// kernel (lattice.cuh, CtcTopo) recomputes it on the GPU.  No CPU fallback.
#pragma once
#include "common.cuh"
#include "launchers.h"

#ifndef WFST_PAIR_WAIT
#define WFST_PAIR_WAIT 0
#endif

namespace wfst {
namespace pairk {

#ifdef WFST_PROFILE
__device__ long long g_tl[8][64];
#define PROF_TL(slot, d, k2) do { if (blockIdx.x == 0 && (d) == 0 && (k2) >= 0 && (k2) < 64 && (threadIdx.x & 31) == 0) \
  g_tl[slot][k2] = clock64(); } while (0)
#define PROF_DECL long long pf_t0 = clock64(), pf_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF_MARK(i) do { long long pf_t1 = clock64(); pf_acc[i] += pf_t1 - pf_t0; pf_t0 = pf_t1; } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_TL(tag, d, k2)
#endif

typedef unsigned long long p2;   // two packed floats: low = utterance 0, high = utterance 1

constexpr int kSeg = 8;               // frames per segment / tile
constexpr int kEventEvery = 2;        // lanes are renormalised every kEventEvery segments
constexpr int kUndef = -(1 << 19);    // "no exponent": lane holds only zeros
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNB = 4;                // p-tile ring depth per direction
constexpr int kNR = 3;                // raw (TMA) staging slots per producer warp
constexpr int kMaxPass = 4;           // label reduction: up to 4 x 32 chunk slots
constexpr int kU = 4;                 // frames unrolled in the hot loops (code size vs address updates)

struct Args {
  const float* E;
  const int* targets;
  const int* offsets;
  int B, T, C, blank;
  const float* grad_scale;
  float* z_out;     // [B] log Z
  float* gradE;     // [B, T, C] or null
  float* ckpt;      // [blocks][nseg][32][2K+4]
  int* hazard;      // [B]
  int nseg, nA, NAB;
  int fused;        // 1: E holds raw logits; the kernel applies log_softmax and returns d/d logits
};

// ---- packed pairs -----------------------------------------------------------------
__device__ __forceinline__ p2 pk(float x, float y) {
  p2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float lo(p2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a));
  return x;
}
__device__ __forceinline__ float hi(p2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a));
  return y;
}
__device__ __forceinline__ p2 add2(p2 a, p2 b) {
  p2 r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ p2 mul2(p2 a, p2 b) {
  p2 r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) {
  p2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ p2 shfl_up2(p2 a) {
  return pk(__shfl_up_sync(kFull, lo(a), 1), __shfl_up_sync(kFull, hi(a), 1));
}

// ---- shared-state-space accesses on 32-bit addresses ------------------------------
__device__ __forceinline__ float lds(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int ldsi(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ p2 lds64(uint32_t a) {
  p2 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void lds128(uint32_t a, p2& x, p2& y) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a));
}
__device__ __forceinline__ void sts(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void stsi(uint32_t a, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, p2 v) {
  asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, p2 x, p2 y) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(a), "l"(x), "l"(y) : "memory");
}

// 2^x for x <= 0 (one MUFU; results below the normal range flush to zero, which is what a
// probability that small is worth here)
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ bool defined_exp(int e) { return e > kUndef / 2; }
__device__ __forceinline__ float pow2i(int d) {  // 2^d for d in [-126, 127]
  return __uint_as_float((uint32_t)(d + 127) << 23);
}
// 2^d clamped: 0 below the normal range, 2^126 above it (callers bound d from above)
__device__ __forceinline__ float pow2c(int d) { return (d < -126) ? 0.f : pow2i(min(d, 126)); }

// ---- mbarriers ---------------------------------------------------------------------
constexpr int kMaxAB = 4;
constexpr int kBarPFull = 0;                       // [2][kNB]    p tile ready (P -> L, RC)
constexpr int kBarPEmpty = kBarPFull + 2 * kNB;    // [2][kNB]    p tile released (count 2)
constexpr int kBarTma = kBarPEmpty + 2 * kNB;      // [2][kNR]    raw tiles landed
constexpr int kBarAFull = kBarTma + 2 * kNR;       // [2][kMaxAB] abar segment ready (L -> RC)
constexpr int kBarCFull = kBarAFull + 2 * kMaxAB;  // [2][kMaxAB] products ready (RC -> X)
constexpr int kBarAEmpty = kBarCFull + 2 * kMaxAB; // [2][kMaxAB] segment buffer free (X -> L)
constexpr int kBarZ = kBarAEmpty + 2 * kMaxAB;     // Z published (L1 -> everyone)
constexpr int kSD = 12;                            // > the furthest P can run ahead of X (kNB + kMaxAB tiles)
constexpr int kBarSDone = kBarZ + 1;               // [2][kSD]    fused mode: softmax tile of phase-2 tile i is in HBM (P -> X)
constexpr int kNumBars = kBarSDone + 2 * kSD;

__device__ __forceinline__ void bar_init(uint32_t bars, int idx, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8u * idx), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint32_t bars, int idx, uint32_t count = 1) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8u * idx), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bars, int idx, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8u * idx), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bars, int idx, uint32_t parity) {
#if WFST_PAIR_WAIT == 1
  // plain polling
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WFSTP_BW_%=:\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WFSTP_BD_%=;\n"
      "bra WFSTP_BW_%=;\n"
      "WFSTP_BD_%=:\n"
      "}\n" ::"r"(bars + 8u * idx), "r"(parity) : "memory");
#elif WFST_PAIR_WAIT == 2
  // try_wait with the default (implementation-defined) suspend time
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WFSTP_BW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WFSTP_BD_%=;\n"
      "bra WFSTP_BW_%=;\n"
      "WFSTP_BD_%=:\n"
      "}\n" ::"r"(bars + 8u * idx), "r"(parity) : "memory");
#else
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WFSTP_BW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WFSTP_BD_%=;\n"
      "bra WFSTP_BW_%=;\n"
      "WFSTP_BD_%=:\n"
      "}\n" ::"r"(bars + 8u * idx), "r"(parity), "r"(0x989680u) : "memory");
#endif
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- shared memory layout (in floats) ---------------------------------------------
struct Layout {
  size_t raw, out, abuf, bnd, lexp, ptile, bars, zx, xtab, colpos, slotlab, hist, total;
};
__host__ __device__ inline Layout make_layout(int K, int C, int CS, int NAB) {
  Layout L;
  const size_t rawsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
  const size_t RS = 32 * ((size_t)K / 2 + 1) + 4, BS = 32 * ((size_t)K + 1) + 4, Cp = (size_t)CS + 1;
  size_t p = 0;
  L.raw = p;    p += 2 * (size_t)kNR * 2 * rawsz + 32;   // [d][slot][u][8*C] (+ slack: the producers over-read)
  L.out = p;    p += 2 * 2 * 2 * rawsz;                  // [d][ob][u][8*C]
  L.abuf = p;   p += 2 * (size_t)NAB * kSeg * 2 * RS;    // [d][buf][row][4 pad + 32 x (K/2 + 1) label slots][u]
  L.bnd = p;    p += 2 * (size_t)NAB * 2 * BS;           // [d][buf][4 pad + 32 x (K + 1) slots][u]: L's state at the segment boundary
  L.lexp = p;   p += 2 * (size_t)NAB * 2 * 32;           // [d][buf][u][lane] (int)
  L.ptile = p;  p += 2 * (size_t)kNB * kSeg * 2 * Cp;    // [d][buf][row][u][Cp]
  p = (p + 3) & ~(size_t)3;
  L.bars = p;   p += 2 * kNumBars;
  p = (p + 3) & ~(size_t)3;
  L.zx = p;     p += 12;                                 // [u]{Zm, eZ, ok, -}, [d] partial sum of row maxima (double)
  L.xtab = p;   p += 2 * 2 * (size_t)kMaxPass * 32 * 4;  // [d][u][pass][lane]{4 x u16 offsets, flags | label << 8, -}
  L.colpos = p; p += 2 * 4 * 32 * kMaxPass;              // [u][column] -> target position
  L.slotlab = p; p += 2 * (32 * kMaxPass + 8);           // [u][slot] -> label
  L.hist = p;   p += 2 * ((size_t)C + 4);                // [u] counting-sort scratch
  L.total = p + 4;
  return L;
}

struct Smem {
  uint32_t raw, out, abuf, bnd, lexp, ptile, bars, zx, xtab;
  int* colpos;
  int* slotlab;
  int* hist;
  float* out_gen;   // generic pointer of the out tiles (non-TMA store path)
};

// per-block constants (arrays are only ever indexed by unrolled constants)
struct Ctx {
  int lane, T, C, nseg, nA, NAB;
  int b[2], L[2];
  const int* y[2];
  bool live[2];       // utterance exists and passed the setup checks
  bool want_grad;
  uint32_t rawsz;
};

__device__ __forceinline__ int seg_of(int dir, int k, int nseg) { return dir == 0 ? k : nseg - 1 - k; }

// ---- per-lane topology --------------------------------------------------------------
template <int K>
struct Topo {
  uint32_t labofs[2][K / 2];  // byte offset in a p-tile row of the label of odd slot 2q+1
                              // (column C, always zero, for padding slots), per utterance
  p2 skipm[K / 2];            // 1 if the skip arc into odd slot 2q+1 exists
};

// orientation o: slot j holds state s = j (o = 0) or s = Sp-2-j (o = 1)
template <int K>
__device__ __forceinline__ void build_topo(Topo<K>& tp, const Ctx& cx, int o, int plane) {
  constexpr int Sp = 32 * K;
#pragma unroll
  for (int q = 0; q < K / 2; ++q) {
    const int j = cx.lane * K + 2 * q + 1;
    const int s = o == 0 ? j : Sp - 2 - j;
    float sk[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int L = cx.L[u];
      int col = cx.C;
      sk[u] = 0.f;
      if (s >= 1 && s < 2 * L + 1) {
        const int n = (s - 1) >> 1;
        col = min(max(cx.y[u][n], 0), cx.C - 1);
        const int n2 = o == 0 ? n - 1 : n + 1;   // two positions earlier IN THIS ORIENTATION
        if (n2 >= 0 && n2 < L && cx.y[u][n2] != cx.y[u][n]) sk[u] = 1.f;
      }
      tp.labofs[u][q] = 4u * (uint32_t)(u * plane + col);
    }
    tp.skipm[q] = pk(sk[0], sk[1]);
  }
}

// The p values one frame needs.  A p tile is [8 rows][2 utterances][CS + 1 columns] (one plane
// per utterance: the labels of a gather then fall into distinct banks); its rows are in the
// order the live warp of the direction consumes them.
template <int K>
struct PRow {
  p2 pl[K / 2];
  p2 pb;
};
template <int K>
struct TileAddr {
  uint32_t la[2][K / 2];   // row-0 address of each odd slot's label
  uint32_t pba;            // row-0 address of the blank pair
};
template <int K>
__device__ __forceinline__ TileAddr<K> tile_addr(const Topo<K>& tp, uint32_t pt, uint32_t blank_ofs) {
  TileAddr<K> t;
#pragma unroll
  for (int q = 0; q < K / 2; ++q) {
    t.la[0][q] = pt + tp.labofs[0][q];
    t.la[1][q] = pt + tp.labofs[1][q];
  }
  t.pba = pt + blank_ofs;
  return t;
}
template <int K, int CS>
__device__ __forceinline__ void advance_rows(TileAddr<K>& t, int rows) {
  const uint32_t o = (uint32_t)(rows * (8 * (CS + 1)));
#pragma unroll
  for (int q = 0; q < K / 2; ++q) {
    t.la[0][q] += o;
    t.la[1][q] += o;
  }
  t.pba += o;
}
template <int K, int CS>
__device__ __forceinline__ PRow<K> load_prow(const TileAddr<K>& t, int it) {
  PRow<K> p;
  const uint32_t o = (uint32_t)it * (8u * (CS + 1));
#pragma unroll
  for (int q = 0; q < K / 2; ++q) p.pl[q] = pk(lds(t.la[0][q] + o), lds(t.la[1][q] + o));
  p.pb = pk(lds(t.pba + o), lds(t.pba + o + 4u * (CS + 1)));
  return p;
}

// One frame.  v: with-emission values of the previous frame (own scale); on return this
// frame's with-emission values and, if WANT_ABAR, abar = the pre-emission sums.  f converts
// the left neighbour's scale to ours (0 in lane 0).
template <int K, bool WANT_ABAR>
__device__ __forceinline__ void step(p2 (&v)[K], p2 (&abar)[K], const Topo<K>& tp, const PRow<K>& p, p2 f) {
  const p2 in1 = mul2(shfl_up2(v[K - 1]), f);
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    const p2 a1 = (i >= 1) ? v[i - 1] : in1;
    p2 s = add2(v[i], a1);
    if (i & 1) {
      const p2 a2 = (i >= 2) ? v[i - 2] : in1;
      s = fma2(tp.skipm[i >> 1], a2, s);
      if (WANT_ABAR) abar[i] = s;
      v[i] = mul2(s, p.pl[i >> 1]);
    } else {
      if (WANT_ABAR) abar[i] = s;
      v[i] = mul2(s, p.pb);
    }
  }
}

// Event: renormalise the lane (max mantissa in [1,2)) and make the lane exponents
// consistent from left to right (the direction mass flows):
//   * a lane that holds only zeros takes the exponent of its left neighbour, so mass
//     arriving during the next 16 frames arrives unscaled;
//   * a lane with own mass never sits more than D below its left neighbour, where D is
//     small enough that a wave crossing several lanes inside one 16-frame window cannot
//     overflow: D * (lanes crossed) + log2(3^16) < 127.
// This is the prefix composition of the maps x -> max(c, x - d) with (c, d) = (own
// exponent, D) or (undefined, 0), which is associative: a 5-step warp scan.  "Undefined" is
// any value below kUndef / 2, so max / minus need no special cases; (c, d) travel in one
// shuffle as c * 2048 + d (d < 2048).  Both utterances are scanned side by side.
template <int K>
__device__ __forceinline__ void event2(p2 (&v)[K], int (&e)[2], p2& f, int lane) {
  constexpr int kChain = (32 + K - 1) / K + 1;   // lanes a wave can cross in 16 frames
  constexpr int D = 96 / kChain;
  float m[2] = {lo(v[0]), hi(v[0])};
#pragma unroll
  for (int i = 1; i < K; ++i) {
    m[0] = fmaxf(m[0], lo(v[i]));
    m[1] = fmaxf(m[1], hi(v[i]));
  }
  int ex[2], eown[2], c[2], d[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    ex[u] = min(max((int)((__float_as_uint(m[u]) >> 23) & 0xffu) - 127, -126), 126);
    const bool has = m[u] > 0.f;
    if (!has) ex[u] = 0;
    eown[u] = has ? (defined_exp(e[u]) ? e[u] : 0) + ex[u] : kUndef;
    c[u] = eown[u];
    d[u] = has ? D : 0;
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int pc[2], pd[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int packed = __shfl_up_sync(kFull, c[u] * 2048 + d[u], o);
      pd[u] = packed & 2047;
      pc[u] = packed >> 11;            // arithmetic shift: floor((c * 2048 + d) / 2048) = c
    }
    if (lane >= o) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        c[u] = max(c[u], max(pc[u], kUndef) - d[u]);
        d[u] += pd[u];
      }
    }
  }
  int t[2];           // total power-of-two shift applied to the lane
  float fv[2];
  bool deep = false;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int E = defined_exp(c[u]) ? c[u] : kUndef;
    t[u] = -ex[u] + ((defined_exp(E) && defined_exp(eown[u])) ? eown[u] - E : 0);   // second term <= 0
    e[u] = E;
    const int el = __shfl_up_sync(kFull, E, 1);
    fv[u] = (lane == 0 || !defined_exp(el) || !defined_exp(E)) ? 0.f : pow2c(el - E);   // el - E <= D
    deep = deep || t[u] < -126;
  }
  f = pk(fv[0], fv[1]);
  {
    const p2 s1 = pk(pow2i(max(t[0], -126)), pow2i(max(t[1], -126)));
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = mul2(v[i], s1);
  }
  if (__any_sync(kFull, deep)) {   // a lane pushed far below its own maximum: second factor
    const p2 s2 = pk(pow2c(t[0] - max(t[0], -126)), pow2c(t[1] - max(t[1], -126)));
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = mul2(v[i], s2);
  }
}

// checkpoint: per lane 2K+4 floats (K pairs, then the two exponents), 128-bit accesses
template <int K>
__host__ __device__ constexpr int ck_floats() { return 2 * K + 4; }
template <int K>
__device__ __forceinline__ void ckpt_store(float* base, const p2 (&v)[K], const int (&e)[2], int lane) {
  ulonglong2* p = reinterpret_cast<ulonglong2*>(base + (size_t)lane * ck_floats<K>());
#pragma unroll
  for (int i = 0; i < K; i += 2) p[i >> 1] = make_ulonglong2(v[i], v[i + 1]);
  p[K >> 1] = make_ulonglong2(pk(__int_as_float(e[0]), __int_as_float(e[1])), 0ull);
}
template <int K>
__device__ __forceinline__ void ckpt_load(const float* base, p2 (&v)[K], int (&e)[2], int lane) {
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(base + (size_t)lane * ck_floats<K>());
#pragma unroll
  for (int i = 0; i < K; i += 2) {
    const ulonglong2 q = p[i >> 1];
    v[i] = q.x;
    v[i + 1] = q.y;
  }
  const ulonglong2 q = p[K >> 1];
  e[0] = __float_as_int(lo(q.x));
  e[1] = __float_as_int(hi(q.x));
}

// p-tile ring of one direction, as seen by a consumer warp
struct PTileRing {
  uint32_t bars, base, tile_bytes;
  int dir;
  uint32_t phase;   // bit per buffer
  __device__ __forceinline__ uint32_t wait(int k) {
    const int buf = k % kNB;
    bar_wait(bars, kBarPFull + dir * kNB + buf, (phase >> buf) & 1u);
    phase ^= 1u << buf;
    return base + (uint32_t)buf * tile_bytes;
  }
  // for a consumer that does not see every tile: the parity of use k of its buffer
  __device__ __forceinline__ uint32_t wait_use(int k) {
    const int buf = k % kNB;
    bar_wait(bars, kBarPFull + dir * kNB + buf, (uint32_t)(k / kNB) & 1u);
    return base + (uint32_t)buf * tile_bytes;
  }
  __device__ __forceinline__ void release(int k, int lane, uint32_t count) {
    __syncwarp();
    if (lane == 0) bar_arrive(bars, kBarPEmpty + dir * kNB + (k % kNB), count);
  }
};

// ---------------------------------------------------------------------------
// P<d>: producer of direction d.  Direction 0 consumes tiles 0,1,...; direction 1 consumes
// nseg-1, nseg-2, ...  Phase-1 tiles first; the phase-2 tiles only once Z is known to be
// usable for at least one utterance.
// ---------------------------------------------------------------------------
struct ProducerState {
  int fetched, converted;
  uint32_t tma_phase, tma_used, empty_phase;
  double msum[2];
};

template <int CS>
__device__ __forceinline__ void produce_range(const Args& a, const Smem& sm, const Ctx& cx, ProducerState& ps,
                                              const int d, const int kbeg, const int kcnt, const bool phase1) {
  constexpr int NPL = CS / 4;      // labels per lane: a lane converts a quarter of one row
  const int lane = cx.lane, T = cx.T, C = cx.C, nseg = cx.nseg;
  const uint32_t rawsz = cx.rawsz;
  const int fr = lane & 7, part = lane >> 3;                       // 8 frames x 4 label quarters
  const int c0 = (C * part) / 4, cnt = (C * (part + 1)) / 4 - c0;
  const uint32_t raw0 = sm.raw + 4u * (uint32_t)(d * kNR * 2) * rawsz;
  const int tbar = kBarTma + d * kNR;
  auto issue_raw = [&](int k) {
    const int tile = seg_of(d, k, nseg);
    const int rows = min(kSeg, T - tile * kSeg);
    const int slot = ps.fetched % kNR;
    const uint32_t bytes = (uint32_t)rows * C * 4u;
    const float* src0 = a.E + ((size_t)cx.b[0] * T + (size_t)tile * kSeg) * C;
    const float* src1 = a.E + ((size_t)cx.b[1] * T + (size_t)tile * kSeg) * C;
    const uint32_t dst0 = raw0 + 4u * (uint32_t)(slot * 2) * rawsz, dst1 = dst0 + 4u * rawsz;
    const bool tma = (bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(src0) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(src1) & 15) == 0;
    if (tma) {
      if (lane == 0) {
        bar_expect_tx(sm.bars, tbar + slot, 2u * bytes);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst0),
            "l"(src0), "r"(bytes), "r"(sm.bars + 8u * (tbar + slot))
            : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst1),
            "l"(src1), "r"(bytes), "r"(sm.bars + 8u * (tbar + slot))
            : "memory");
      }
      ps.tma_used |= 1u << slot;
    } else {
      for (int q = lane; q < rows * C; q += 32) {
        sts(dst0 + 4u * q, __ldg(src0 + q));
        sts(dst1 + 4u * q, __ldg(src1 + q));
      }
      ps.tma_used &= ~(1u << slot);
      __syncwarp();
    }
    ++ps.fetched;
  };
  int kf = 0;   // next entry to fetch (relative)
  PROF_DECL;
  for (int i = 0; i < kcnt; ++i) {
    PROF_MARK(2);
    while (kf < kcnt && ps.fetched < ps.converted + kNR) issue_raw(kbeg + kf++);
    const int k = kbeg + i;
    const int slot = ps.converted % kNR;
    const int tile = seg_of(d, k, nseg);
    const int rows = min(kSeg, T - tile * kSeg);
    const int buf = k % kNB;
    if (k >= kNB) {  // wait until the consumers have released this p-tile buffer
      bar_wait(sm.bars, kBarPEmpty + d * kNB + buf, (ps.empty_phase >> buf) & 1u);
      ps.empty_phase ^= 1u << buf;
    }
    PROF_MARK(0);
    if ((ps.tma_used >> slot) & 1u) {
      bar_wait(sm.bars, tbar + slot, (ps.tma_phase >> slot) & 1u);
      ps.tma_phase ^= 1u << slot;
    }
    PROF_MARK(1);
    const bool live = fr < rows;
    const uint32_t er0 = raw0 + 4u * ((uint32_t)(slot * 2) * rawsz + (uint32_t)(fr * C + c0));
    const uint32_t er1 = er0 + 4u * rawsz;
    // tile row = the step at which the live warp of direction d consumes frame fr
    const int trow = d == 0 ? fr : rows - 1 - fr;
    const uint32_t pt = sm.ptile + (uint32_t)((d * kNB + buf) * kSeg + (live ? trow : 0)) * (8u * (CS + 1)) + 4u * (uint32_t)c0;
    // loads run past the lane's quarter (into the next row / the slack behind the staging
    // area); those elements are masked
    float ev0[NPL], ev1[NPL];
#pragma unroll
    for (int i2 = 0; i2 < NPL; ++i2) {
      ev0[i2] = lds(er0 + 4u * i2);
      ev1[i2] = lds(er1 + 4u * i2);
    }
    float mx0 = kNegInf, mx1 = kNegInf;
#pragma unroll
    for (int i2 = 0; i2 < NPL; ++i2) {
      if (i2 >= cnt) { ev0[i2] = kNegInf; ev1[i2] = kNegInf; }
      mx0 = fmaxf(mx0, ev0[i2]);
      mx1 = fmaxf(mx1, ev1[i2]);
    }
    if (!live) { mx0 = kNegInf; mx1 = kNegInf; }
    mx0 = fmaxf(mx0, __shfl_xor_sync(kFull, mx0, 8));
    mx1 = fmaxf(mx1, __shfl_xor_sync(kFull, mx1, 8));
    mx0 = fmaxf(mx0, __shfl_xor_sync(kFull, mx0, 16));
    mx1 = fmaxf(mx1, __shfl_xor_sync(kFull, mx1, 16));
    // a row that is entirely -inf keeps p = 0 (dead frame); +inf / NaN rows surface through
    // the certificate
    const float base0 = (mx0 == kNegInf) ? 0.f : mx0, base1 = (mx1 == kNegInf) ? 0.f : mx1;
    const float nb0 = -base0 * 1.4426950408889634f, nb1 = -base1 * 1.4426950408889634f;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i2 = 0; i2 < NPL; ++i2) {
      ev0[i2] = ex2_fast(fmaf(ev0[i2], 1.4426950408889634f, nb0));   // masked elements: 2^-inf = 0
      ev1[i2] = ex2_fast(fmaf(ev1[i2], 1.4426950408889634f, nb1));
      s0 += ev0[i2];
      s1 += ev1[i2];
      if (live && i2 < cnt) {
        sts(pt + 4u * i2, ev0[i2]);
        sts(pt + 4u * (CS + 1) + 4u * i2, ev1[i2]);
      }
    }
    if (a.fused) {
      // E holds logits: log_softmax(x)_c = x_c - max - log(sum_c 2^...).  The recursion runs on the
      // unnormalised p, so log Z = log Z~ - sum_t log s_t; d loss / d logits = gs * (p / s - posterior):
      // the first term leaves from here (the staged raw tile is overwritten in place and stored),
      // X adds the second with a bulk reduce once this store is known to have landed.
      s0 += __shfl_xor_sync(kFull, s0, 8);
      s1 += __shfl_xor_sync(kFull, s1, 8);
      s0 += __shfl_xor_sync(kFull, s0, 16);
      s1 += __shfl_xor_sync(kFull, s1, 16);
      if (live && part == 0 && phase1) {
        ps.msum[0] -= (double)logf(s0);
        ps.msum[1] -= (double)logf(s1);
      }
      if (!phase1) {
        if (i > 0) {   // the previous tile's store has landed: tell X
          if (lane == 0) {
            bulk_wait_all<0>();
            bar_arrive(sm.bars, kBarSDone + d * kSD + ((i - 1) % kSD));
          }
          __syncwarp();
        }
        const float g0 = (a.grad_scale ? a.grad_scale[cx.b[0]] : 1.f) / s0;
        const float g1 = (a.grad_scale ? a.grad_scale[cx.b[1]] : 1.f) / s1;
#pragma unroll
        for (int i2 = 0; i2 < NPL; ++i2) {
          if (live && i2 < cnt) {
            sts(er0 + 4u * i2, ev0[i2] * g0);
            sts(er1 + 4u * i2, ev1[i2] * g1);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const uint32_t bytes = (uint32_t)rows * C * 4u;
          const uint32_t src0 = raw0 + 4u * (uint32_t)(slot * 2) * rawsz;
          float* dst0 = a.gradE + ((size_t)cx.b[0] * T + (size_t)tile * kSeg) * C;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst0), "r"(src0), "r"(bytes) : "memory");
          if (cx.b[1] != cx.b[0]) {
            float* dst1 = a.gradE + ((size_t)cx.b[1] * T + (size_t)tile * kSeg) * C;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst1), "r"(src0 + 4u * rawsz), "r"(bytes) : "memory");
          }
          bulk_commit();
          bulk_wait_read<0>();   // the staging slot may be refilled
        }
        __syncwarp();
      }
    } else if (live && part == 0 && phase1) {
      ps.msum[0] += (double)base0;
      ps.msum[1] += (double)base1;
    }
    __syncwarp();
    if (lane == 0) bar_arrive(sm.bars, kBarPFull + d * kNB + buf);
    if (!phase1) PROF_TL(7, d, i);
    ++ps.converted;
  }
  if (a.fused && !phase1 && kcnt > 0) {
    if (lane == 0) {
      bulk_wait_all<0>();
      bar_arrive(sm.bars, kBarSDone + d * kSD + ((kcnt - 1) % kSD));
    }
    __syncwarp();
  }
  PROF_MARK(2);
#ifdef WFST_PROFILE
  if (blockIdx.x == 0 && lane == 0)
    printf("P%d (phase1 %d) cycles: wait_pempty %lld wait_tma %lld convert+issue %lld\n", d, (int)phase1, pf_acc[0], pf_acc[1], pf_acc[2]);
#endif
}

template <int CS>
__device__ __forceinline__ void role_producer(const Args& a, const Smem& sm, const Ctx& cx, const int d) {
  const int lane = cx.lane, nseg = cx.nseg, nA = cx.nA;
  const int n1 = d == 0 ? nA : nseg - nA;
  ProducerState ps;
  ps.fetched = 0; ps.converted = 0; ps.tma_phase = 0u; ps.tma_used = 0u; ps.empty_phase = 0u;
  ps.msum[0] = 0.0; ps.msum[1] = 0.0;
  produce_range<CS>(a, sm, cx, ps, d, 0, n1, true);
  // loss: log Z = log(Zm) + eZ ln2 + sum_t max_t; the two producers each hold the row maxima
  // of their phase-1 half
#pragma unroll
  for (int u = 0; u < 2; ++u) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ps.msum[u] += __shfl_xor_sync(kFull, ps.msum[u], o);
  }
  double* msh = reinterpret_cast<double*>(__cvta_shared_to_generic(sm.zx + 32u));   // [u] of P0
  if (d == 0 && lane == 0) { msh[0] = ps.msum[0]; msh[1] = ps.msum[1]; }
  named_sync(2, 64);
  bar_wait(sm.bars, kBarZ, 0u);
  bool any = false;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const float Zm = lds(sm.zx + 16u * u);
    const int eZ = ldsi(sm.zx + 16u * u + 4u);
    const bool ok = lds(sm.zx + 16u * u + 8u) != 0.f;
    any = any || ok;
    if (d == 1 && lane == 0 && (u == 0 || cx.b[1] != cx.b[0]))
      a.z_out[cx.b[u]] = ok ? (float)(log((double)Zm) + (double)eZ * 0.6931471805599453 + (ps.msum[u] + msh[u])) : kNegInf;
  }
  if (!cx.want_grad || !any) return;
  produce_range<CS>(a, sm, cx, ps, d, n1, nseg - n1, false);
}

// ---------------------------------------------------------------------------
// L: the live sweep of direction d (whole warp; d is a runtime value so that L0 and L1
// share one copy of the code).
// ---------------------------------------------------------------------------
template <int K, int CS>
__device__ __forceinline__ void role_live(const Args& a, const Smem& sm, const Ctx& cx, const int d) {
  constexpr int Sp = 32 * K;
  constexpr uint32_t ROWB = 8u * (32 * (K / 2 + 1) + 4);   // abar row: 4 pad pairs + 32 lane blocks of K/2 label slots + 1 pad
  constexpr uint32_t BNDB = 8u * (32 * (K + 1) + 4);       // boundary row: 4 pad pairs + 32 lane blocks of K slots + 1 pad
  constexpr uint32_t TILEB = 8u * (CS + 1) * kSeg;
  const int lane = cx.lane, nseg = cx.nseg, nA = cx.nA, T = cx.T, NAB = cx.NAB;
  float* ck = a.ckpt + (size_t)blockIdx.x * nseg * 32 * ck_floats<K>();
  Topo<K> tp;
  build_topo<K>(tp, cx, d, CS + 1);

  p2 v[K], abar[K];
  int e[2] = {kUndef, kUndef};
  p2 f = pk(0.f, 0.f);
  {
    // virtual pre-frame state: all mass on the start slot of this orientation
    float x[2][K];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int S = 2 * cx.L[u] + 1;
      const int jstart = d == 0 ? 0 : Sp - 1 - S;
      const bool mine = (jstart / K == lane);
      const int jm = jstart % K;
#pragma unroll
      for (int i = 0; i < K; ++i) x[u][i] = (mine && jm == i) ? 1.f : 0.f;
      if (mine) e[u] = 0;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = pk(x[0][i], x[1][i]);
  }
  PTileRing ring{sm.bars, sm.ptile + (uint32_t)d * kNB * TILEB, TILEB, d, 0u};
  const uint32_t blank_ofs = 4u * (uint32_t)a.blank;

  // ------------------------------------------------------------------ phase 1
  const int n1 = d == 0 ? nA : nseg - nA;
  PROF_DECL;
  for (int k = 0; k < n1; ++k) {
    const int seg = seg_of(d, k, nseg);
    const int rows = min(kSeg, T - seg * kSeg);
    PROF_MARK(2);
    if (k % kEventEvery == 0) event2<K>(v, e, f, lane);
    ckpt_store<K>(ck + (size_t)seg * 32 * ck_floats<K>(), v, e, lane);
    PROF_MARK(0);
    const TileAddr<K> ta = tile_addr<K>(tp, ring.wait(k), blank_ofs);
    PROF_MARK(1);
    if (rows == kSeg) {
      // kU frames per trip; the row after the last one is prefetched too (it is the first row of
      // the next trip, or a harmless read past the tile)
      PRow<K> nx = load_prow<K, CS>(ta, 0);
#pragma unroll
      for (int it = 0; it < kSeg; ++it) {   // fully unrolled: in phase 1 only L and P compete for the instruction cache
        const PRow<K> cur = nx;
        if (it + 1 < kSeg) nx = load_prow<K, CS>(ta, it + 1);
        step<K, false>(v, abar, tp, cur, f);
      }
    } else {
#pragma unroll 1
      for (int it = 0; it < rows; ++it) {
        const PRow<K> cur = load_prow<K, CS>(ta, it);
        step<K, false>(v, abar, tp, cur, f);
      }
    }
    ring.release(k, lane, 2);   // no recompute warp reads phase-1 tiles
  }

  // ------------------------------------------------------------------ meeting: Z
  // L0 publishes its state in the layout of an abar row (buffer 0, row 0 of direction 0);
  // L1 combines it with the successor sums of its own state.
  const uint32_t myblock = 8u * (uint32_t)(4 + lane * (K / 2 + 1));   // my label slots in an abar row
  const uint32_t mybnd = 8u * (uint32_t)(4 + lane * (K + 1));         // my slots in a boundary row
  if (d == 0) {
#pragma unroll
    for (int i = 0; i < K; ++i) sts64(sm.bnd + mybnd + 8u * i, v[i]);
    stsi(sm.lexp + 4u * (uint32_t)lane, e[0]);
    stsi(sm.lexp + 4u * (uint32_t)(32 + lane), e[1]);
  }
  named_sync(1, 64);
  if (d == 1) {
    event2<K>(v, e, f, lane);     // consistent exponents / f for the shuffle below
    p2 bb[K];
    {
      const p2 in1 = mul2(shfl_up2(v[K - 1]), f);
#pragma unroll
      for (int i = K - 1; i >= 0; --i) {
        const p2 a1 = (i >= 1) ? v[i - 1] : in1;
        p2 s = add2(v[i], a1);
        if (i & 1) {
          const p2 a2 = (i >= 2) ? v[i - 2] : in1;
          s = fma2(tp.skipm[i >> 1], a2, s);
        }
        bb[i] = s;
      }
    }
    // partner of my slot i is slot K-2-i of L0's lane 31-lane; of my slot K-1, the last
    // slot of L0's lane 30-lane (= the element just before that block)
    const uint32_t pblock = sm.bnd + 8u * (uint32_t)(4 + (31 - lane) * (K + 1));
    p2 av[K];
#pragma unroll
    for (int i = 0; i < K; ++i) av[i] = lds64(pblock + 8u * i);
    const p2 ext = lds64(pblock - 16u);
    p2 Pm = pk(0.f, 0.f);
#pragma unroll
    for (int i = 0; i <= K - 2; ++i) Pm = fma2(bb[i], av[K - 2 - i], Pm);
    const p2 Px = mul2(bb[K - 1], ext);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float pm = u ? hi(Pm) : lo(Pm), px = u ? hi(Px) : lo(Px);
      const int ea = ldsi(sm.lexp + 4u * (uint32_t)(u * 32 + 31 - lane));
      const int eb = lane < 31 ? ldsi(sm.lexp + 4u * (uint32_t)(u * 32 + 30 - lane)) : kUndef;
      int Em = kUndef, Ex = kUndef;
      if (pm > 0.f && defined_exp(e[u]) && defined_exp(ea)) Em = e[u] + ea;
      if (px > 0.f && defined_exp(e[u]) && defined_exp(eb)) Ex = e[u] + eb;
      int Emax = max(Em, Ex);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) Emax = max(Emax, __shfl_xor_sync(kFull, Emax, o));
      float contrib = 0.f;
      if (defined_exp(Em)) contrib += pm * pow2c(Em - Emax);
      if (defined_exp(Ex)) contrib += px * pow2c(Ex - Emax);
      contrib = warp_sum(contrib);
      if (lane == 0) {
        const bool ok = defined_exp(Emax) && contrib > 0.f && contrib < 3.0e38f;
        int ex = 0;
        float Zm = 1.f;
        if (ok) {
          ex = (int)((__float_as_uint(contrib) >> 23) & 0xffu) - 127;
          ex = min(max(ex, -126), 126);
          Zm = contrib * pow2i(-ex);
        }
        sts(sm.zx + 16u * u, Zm);
        stsi(sm.zx + 16u * u + 4u, ok ? Emax + ex : 0);
        sts(sm.zx + 16u * u + 8u, ok ? 1.f : 0.f);
        // reason 2: infeasible or out of range -- the log-semiring kernel decides
        if (!ok && (u == 0 || cx.b[1] != cx.b[0])) atomicOr(&a.hazard[cx.b[u]], 2);
      }
    }
    __syncwarp();
  }
  named_sync(1, 64);
  if (d == 1 && lane == 0) bar_arrive(sm.bars, kBarZ);
  const bool any_ok = (lds(sm.zx + 8u) != 0.f) || (lds(sm.zx + 24u) != 0.f);
  if (!cx.want_grad || !any_ok) return;

  // ------------------------------------------------------------------ phase 2
  const int n2 = nseg - n1;
  uint32_t aempty_phase = 0u;
  for (int k2 = 0, buf = 0; k2 < n2; ++k2, buf = (buf + 1 == NAB) ? 0 : buf + 1) {
    const int k = n1 + k2;
    const int seg = seg_of(d, k, nseg);
    const int rows = min(kSeg, T - seg * kSeg);
    PROF_MARK(6);
    if (k % kEventEvery == 0) event2<K>(v, e, f, lane);
    PROF_MARK(3);
    if (k2 >= NAB) {   // X has consumed the segment that used this buffer
      bar_wait(sm.bars, kBarAEmpty + d * kMaxAB + buf, (aempty_phase >> buf) & 1u);
      aempty_phase ^= 1u << buf;
    }
    PROF_MARK(4);
    PROF_TL(0, d, k2);
    {
      const uint32_t le = sm.lexp + 4u * (uint32_t)((d * NAB + buf) * 64 + lane);
      stsi(le, e[0]);
      stsi(le + 128u, e[1]);
      // state at the segment boundary: the recompute warp checks Z against it (certificate)
      const uint32_t bb = sm.bnd + (uint32_t)(d * NAB + buf) * BNDB + mybnd;
#pragma unroll
      for (int i = 0; i < K; ++i) sts64(bb + 8u * i, v[i]);
    }
    TileAddr<K> ta = tile_addr<K>(tp, ring.wait(k), blank_ofs);
    PROF_MARK(5);
    PROF_TL(1, d, k2);
    // rows of the segment buffer are in step order, like the p tile
    uint32_t ar = sm.abuf + (uint32_t)((d * NAB + buf) * kSeg) * ROWB + myblock;
    if (rows == kSeg) {
      PRow<K> nx = load_prow<K, CS>(ta, 0);
#pragma unroll 1
      for (int h = 0; h < kSeg / kU; ++h) {
#pragma unroll
        for (int it = 0; it < kU; ++it) {
          const PRow<K> cur = nx;
          nx = load_prow<K, CS>(ta, it + 1);
          step<K, true>(v, abar, tp, cur, f);
#pragma unroll
          for (int q = 0; q < K / 2; ++q) sts64(ar + (uint32_t)it * ROWB + 8u * q, abar[2 * q + 1]);
        }
        advance_rows<K, CS>(ta, kU);
        ar += kU * ROWB;
      }
      ar -= kSeg * ROWB;
    } else {
#pragma unroll 1
      for (int it = 0; it < rows; ++it) {
        const PRow<K> cur = load_prow<K, CS>(ta, it);
        step<K, true>(v, abar, tp, cur, f);
#pragma unroll
        for (int q = 0; q < K / 2; ++q) sts64(ar + (uint32_t)it * ROWB + 8u * q, abar[2 * q + 1]);
      }
    }
    __syncwarp();
    if (lane == 0) bar_arrive(sm.bars, kBarAFull + d * kMaxAB + buf);
    ring.release(k, lane, 1);
    PROF_TL(2, d, k2);
  }
  // certificate, last leg: the sweep must arrive with total mass Z on the two slots that end the
  // chain in this orientation (the recompute warps check every earlier segment boundary)
  if (n2 > 0) {
    int bad = 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int S = 2 * cx.L[u] + 1;
      const int jend = d == 0 ? S - 1 : Sp - 2;          // last state of the chain in this orientation
      const float Zm = lds(sm.zx + 16u * u);
      const int eZ = ldsi(sm.zx + 16u * u + 4u);
      const bool okz = lds(sm.zx + 16u * u + 8u) != 0.f;
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const int j = lane * K + i;
        const float x = u ? hi(v[i]) : lo(v[i]);
        if (j == jend || (j == jend - 1 && j >= 0)) part += x;
      }
      if (part != 0.f) part *= defined_exp(e[u]) ? pow2c(e[u] - eZ) : 0.f;
      const float tot = warp_sum(part);
      if (okz && cx.live[u] && !(fabsf(tot - Zm) <= 2e-5f * Zm)) bad |= 8 << u;
    }
    if (bad && lane == 0) {
      if (bad & 8) atomicOr(&a.hazard[cx.b[0]], 8);
      if ((bad & 16) && cx.b[1] != cx.b[0]) atomicOr(&a.hazard[cx.b[1]], 8);
    }
  }
  PROF_MARK(6);
#ifdef WFST_PROFILE
  if (blockIdx.x == 0 && lane == 0)
    printf("L%d cycles: ph1 event+ckpt %lld wait_ptile %lld steps %lld | ph2 event %lld wait_aempty %lld wait_ptile %lld steps %lld\n",
           d, pf_acc[0], pf_acc[1], pf_acc[2], pf_acc[3], pf_acc[4], pf_acc[5], pf_acc[6]);
#endif
}

// ---------------------------------------------------------------------------
// RC: serves L<d>.  Runs the opposite orientation over the segments of L<d>'s phase 2 from
// the checkpoints the other live warp wrote in phase 1, against the step order of L<d>,
// and multiplies with L<d>'s stored abar rows.  Before each segment the lane is rescaled so
// that its effective exponent is eZ - eL(partner lane): products need no further factor.
// ---------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void rc_frame(p2 (&w)[K], const Topo<K>& tp, const PRow<K>& cur, p2 f2, p2 h2, uint32_t arow) {
  // partner of my odd slot i (<= K-3) is label slot (K-3-i)/2 of the partner block; of my slot
  // K-1, the last label slot of the block before it (one pad pair in between)
  p2 av[K / 2], dummy[K];
#pragma unroll
  for (int q = 0; q < K / 2 - 1; ++q) av[q] = lds64(arow + 8u * q);
  const p2 ext = lds64(arow - 16u);
  step<K, false>(w, dummy, tp, cur, f2);
#pragma unroll
  for (int q = 0; q < K / 2 - 1; ++q) sts64(arow + 8u * q, mul2(w[K - 3 - 2 * q], av[q]));
  sts64(arow - 16u, mul2(mul2(w[K - 1], ext), h2));
}

template <int K, int CS>
__device__ __forceinline__ void role_rc(const Args& a, const Smem& sm, const Ctx& cx, const int d, const int x) {
  constexpr int Sp = 32 * K;
  constexpr uint32_t ROWB = 8u * (32 * (K / 2 + 1) + 4);   // abar row: 4 pad pairs + 32 lane blocks of K/2 label slots + 1 pad
  constexpr uint32_t BNDB = 8u * (32 * (K + 1) + 4);       // boundary row: 4 pad pairs + 32 lane blocks of K slots + 1 pad
  constexpr uint32_t TILEB = 8u * (CS + 1) * kSeg;
  const int lane = cx.lane, nseg = cx.nseg, nA = cx.nA, T = cx.T, NAB = cx.NAB;
  bar_wait(sm.bars, kBarZ, 0u);      // phase 1 (and every checkpoint) is complete
  const bool ok0 = lds(sm.zx + 8u) != 0.f, ok1 = lds(sm.zx + 24u) != 0.f;
  if (!cx.want_grad || !(ok0 || ok1)) return;
  const int eZ[2] = {ldsi(sm.zx + 4u), ldsi(sm.zx + 20u)};
  const int n1 = d == 0 ? nA : nseg - nA;
  const int n2 = nseg - n1;
  Topo<K> tp;
  build_topo<K>(tp, cx, 1 - d, CS + 1);
  const float* ck = a.ckpt + (size_t)blockIdx.x * nseg * 32 * ck_floats<K>();
  PTileRing ring{sm.bars, sm.ptile + (uint32_t)d * kNB * TILEB, TILEB, d, 0u};
  const uint32_t blank_ofs = 4u * (uint32_t)a.blank;
  const uint32_t pblock = 8u * (uint32_t)(4 + (31 - lane) * (K / 2 + 1));   // partner block in an abar row
  const uint32_t pbnd = 8u * (uint32_t)(4 + (31 - lane) * (K + 1));        // partner block in a boundary row
  int bad = 0;   // reason 4: scale overflow when pairing live and recomputed values
  p2 w[K];
  int ew[2];
  // the two RC warps of a direction take alternate segments
  if (x < n2) ckpt_load<K>(ck + (size_t)seg_of(d, n1 + x, nseg) * 32 * ck_floats<K>(), w, ew, lane);
  PROF_DECL;
  for (int k2 = x; k2 < n2; k2 += 2) {
    const int k = n1 + k2;
    const int seg = seg_of(d, k, nseg);
    const int rows = min(kSeg, T - seg * kSeg);
    const int buf = k2 % NAB;
    PROF_MARK(2);
    bar_wait(sm.bars, kBarAFull + d * kMaxAB + buf, (uint32_t)(k2 / NAB) & 1u);
    PROF_MARK(0);
    PROF_TL(3, d, k2);
    // scales
    float g[2], h[2], fr[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t le = sm.lexp + 4u * (uint32_t)((d * NAB + buf) * 64 + u * 32);
      const int ep = ldsi(le + 4u * (uint32_t)(31 - lane));                       // partner lane
      const int ex = lane < 31 ? ldsi(le + 4u * (uint32_t)(30 - lane)) : kUndef;  // partner of slot K-1
      const int el = lane > 0 ? ldsi(le + 4u * (uint32_t)(32 - lane)) : kUndef;   // partner of my left neighbour
      g[u] = 0.f; h[u] = 0.f; fr[u] = 0.f;
      if (defined_exp(ep)) {
        if (defined_exp(ew[u])) {
          const int dd = ew[u] + ep - eZ[u];
          if (dd > 126) bad |= 4 << u;
          else g[u] = pow2c(dd);
        }
        if (defined_exp(ex)) h[u] = pow2c(ex - ep);     // <= 2^D by the event invariant
        if (defined_exp(el)) fr[u] = pow2c(ep - el);    // <= 2^D likewise
      }
    }
    {
      const p2 g2 = pk(g[0], g[1]);
#pragma unroll
      for (int i = 0; i < K; ++i) w[i] = mul2(w[i], g2);
    }
    const p2 f2 = pk(fr[0], fr[1]), h2 = pk(h[0], h[1]);
    PROF_MARK(2);
    TileAddr<K> ta = tile_addr<K>(tp, ring.wait_use(k), blank_ofs);
    PROF_MARK(1);
    const uint32_t ar = sm.abuf + (uint32_t)((d * NAB + buf) * kSeg) * ROWB + pblock;
    if (rows == kSeg) {
      // against L's step order, kU frames per trip; th / ah / bh point at the trip's lowest row
      TileAddr<K> th = ta;
      advance_rows<K, CS>(th, kSeg - kU);
      uint32_t ah = ar + (kSeg - kU) * ROWB;
      PRow<K> nx = load_prow<K, CS>(th, kU - 1);
#pragma unroll 1
      for (int h = 0; h < kSeg / kU; ++h) {
#pragma unroll
        for (int it = kU - 1; it >= 0; --it) {
          const PRow<K> cur = nx;
          nx = load_prow<K, CS>(th, it - 1);   // it = 0: the last row of the next trip (or a harmless read)
          rc_frame<K>(w, tp, cur, f2, h2, ah + (uint32_t)it * ROWB);
        }
        advance_rows<K, CS>(th, -kU);
        ah -= kU * ROWB;
      }
    } else {
#pragma unroll 1
      for (int it = rows - 1; it >= 0; --it) {
        const PRow<K> cur = load_prow<K, CS>(ta, it);
        rc_frame<K>(w, tp, cur, f2, h2, ar + (uint32_t)it * ROWB);
      }
    }
    {
      // certificate: sum_s v_L(s) * (successor sum of w)(s) at the segment boundary must be Z
      // (float32 range can only be exceeded by losing mass or producing inf / NaN)
      const p2 in1 = mul2(shfl_up2(w[K - 1]), f2);
      const uint32_t bb = sm.bnd + (uint32_t)(d * NAB + buf) * BNDB + pbnd;
      p2 acc = pk(0.f, 0.f);
#pragma unroll
      for (int i = K - 1; i >= 0; --i) {
        const p2 a1 = (i >= 1) ? w[i - 1] : in1;
        p2 sx = add2(w[i], a1);
        if (i & 1) {
          const p2 a2 = (i >= 2) ? w[i - 2] : in1;
          sx = fma2(tp.skipm[i >> 1], a2, sx);
        }
        if (i == K - 1) acc = fma2(mul2(sx, h2), lds64(bb - 16u), acc);
        else acc = fma2(sx, lds64(bb + 8u * (K - 2 - i)), acc);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float tot = warp_sum(u ? hi(acc) : lo(acc));
        const float Zm = lds(sm.zx + 16u * u);
        const bool chk = (u ? ok1 : ok0) && cx.live[u];
        if (chk && !(fabsf(tot - Zm) <= 2e-5f * Zm)) bad |= 16 << u;
      }
    }
    // next checkpoint (consumed at the top of the next iteration)
    if (k2 + 2 < n2) ckpt_load<K>(ck + (size_t)seg_of(d, k + 2, nseg) * 32 * ck_floats<K>(), w, ew, lane);
    __syncwarp();
    if (lane == 0) bar_arrive(sm.bars, kBarCFull + d * kMaxAB + buf);
    ring.release(k, lane, 1);
    PROF_TL(4, d, k2);
  }
  PROF_MARK(2);
#ifdef WFST_PROFILE
  if (blockIdx.x == 0 && lane == 0)
    printf("RC%d.%d cycles: wait_afull %lld wait_ptile %lld compute %lld\n", d, x, pf_acc[0], pf_acc[1], pf_acc[2]);
#endif
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0) {
    if (bad & 4) atomicOr(&a.hazard[cx.b[0]], 4);
    if ((bad & 8) && cx.b[1] != cx.b[0]) atomicOr(&a.hazard[cx.b[1]], 4);
    if (bad & 16) atomicOr(&a.hazard[cx.b[0]], 8);
    if ((bad & 32) && cx.b[1] != cx.b[0]) atomicOr(&a.hazard[cx.b[1]], 8);
  }
}

// ---------------------------------------------------------------------------
// X<d, u>: per-label reduction of a segment's posteriors of utterance u in direction d +
// gradient tile store.  (d, u are runtime values: the four X warps share one copy of the code.)
// ---------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void role_reduce(const Args& a, const Smem& sm, const Ctx& cx, const int d, const int ux) {
  constexpr int Sp = 32 * K;
  constexpr uint32_t ROWB = 8u * (32 * (K / 2 + 1) + 4);   // abar row: 4 pad pairs + 32 lane blocks of K/2 label slots + 1 pad
  constexpr uint32_t BNDB = 8u * (32 * (K + 1) + 4);       // boundary row: 4 pad pairs + 32 lane blocks of K slots + 1 pad
  const int lane = cx.lane, nseg = cx.nseg, nA = cx.nA, T = cx.T, C = cx.C, NAB = cx.NAB;
  bar_wait(sm.bars, kBarZ, 0u);
  const bool ok0 = lds(sm.zx + 8u) != 0.f, ok1 = lds(sm.zx + 24u) != 0.f;
  if (!cx.want_grad || !(ok0 || ok1)) return;
  const int n1 = d == 0 ? nA : nseg - nA;
  const int n2 = nseg - n1;
  const uint32_t rawsz = cx.rawsz;
  const uint32_t blank_ofs = 4u * (uint32_t)a.blank;
  const int bu = ux ? cx.b[1] : cx.b[0];
  const bool dup = ux == 1 && cx.b[1] == cx.b[0];
  const float Zm = lds(sm.zx + 16u * ux);
  const float kappa = -(a.grad_scale ? a.grad_scale[bu] : 1.f) / Zm;
  const bool act = (ux ? cx.live[1] && ok1 : cx.live[0] && ok0);
  const int nslots = sm.hist[ux * (C + 4) + C], maxch = sm.hist[ux * (C + 4) + C + 1];
  const int npass = (nslots + 31) >> 5;
  // per-pass constants of this lane -> shared table (one 128-bit load per pass later):
  // row-relative byte offsets of the 4 sorted positions of my chunk slot, link flags, label
  const uint32_t xt = sm.xtab + 16u * (uint32_t)(((d * 2 + ux) * kMaxPass) * 32 + lane);
  {
    const int* slotlab = sm.slotlab + ux * (32 * kMaxPass + 8);
    const int* colpos = sm.colpos + ux * (4 * 32 * kMaxPass);
    for (int p = 0; p < kMaxPass; ++p) {
      const int gsl = 32 * p + lane;
      const int me = slotlab[gsl];
      // links never leave the pass: a label's slots do not straddle a multiple of 32
      uint32_t fl = 0u;
      if (me >= 0 && lane + 1 < 32 && slotlab[gsl + 1] == me) fl |= 1u;
      if (me >= 0 && lane + 2 < 32 && slotlab[gsl + 2] == me) fl |= 2u;
      if (me >= 0 && lane + 4 < 32 && slotlab[gsl + 4] == me) fl |= 4u;
      if (me >= 0 && (lane == 0 || slotlab[gsl - 1] != me)) fl |= 8u;
      uint32_t off[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int n = colpos[4 * gsl + q];
        const int jl = d == 0 ? 2 * n + 1 : Sp - 3 - 2 * n;
        off[q] = n >= 0 ? 8u * (uint32_t)(4 + (jl / K) * (K / 2 + 1) + (jl % K) / 2) + 4u * ux : 4u * ux;   // row pad 0 is always zero
      }
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(xt + 512u * p), "r"(off[0] | (off[1] << 16)),
                   "r"(off[2] | (off[3] << 16)), "r"(fl | ((uint32_t)max(me, 0) << 8)), "r"(0u)
                   : "memory");
    }
  }
  __syncwarp();
  int bad = 0;
  uint32_t cfull_phase = 0u;
  float* gE = a.gradE + (size_t)bu * T * C;
  PROF_DECL;
  for (int k2 = 0, buf = 0; k2 < n2; ++k2, buf = (buf + 1 == NAB) ? 0 : buf + 1) {
    const int seg = seg_of(d, n1 + k2, nseg);
    const int rows = min(kSeg, T - seg * kSeg);
    const int ob = k2 & 1;
    PROF_MARK(1);
    bar_wait(sm.bars, kBarCFull + d * kMaxAB + buf, (cfull_phase >> buf) & 1u);
    cfull_phase ^= 1u << buf;
    PROF_MARK(0);
    if (ux == 0) PROF_TL(5, d, k2);
    if (act) {
      if (lane == 0) bulk_wait_read<1>();   // the store that last read this out buffer is done
      __syncwarp();
      const uint32_t ab = sm.abuf + (uint32_t)((d * NAB + buf) * kSeg) * ROWB;
      const uint32_t ot = sm.out + 4u * (uint32_t)(((d * 2 + ob) * 2 + ux) * rawsz);
      // buffer row j holds the frame of step j: frame row r = j (d = 0) or rows-1-j (d = 1)
      const int rsign = d == 0 ? 1 : -1, rbase = d == 0 ? 0 : rows - 1;
      float rs[kSeg];     // per-row sum of the label posteriors this lane is head of
#pragma unroll
      for (int j = 0; j < kSeg; ++j) rs[j] = 0.f;
#pragma unroll 1
      for (int p = 0; p < npass; ++p) {
        uint32_t w0, w1, w2, w3;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(xt + 512u * p));
        const uint32_t o0 = ab + (w0 & 0xffffu), o1 = ab + (w0 >> 16), o2 = ab + (w1 & 0xffffu), o3 = ab + (w1 >> 16);
        const float lk1 = (w2 & 1u) ? 1.f : 0.f, lk2 = (w2 & 2u) ? 1.f : 0.f, lk4 = (w2 & 4u) ? 1.f : 0.f;
        const bool head = (w2 & 8u) != 0u;
        const int me = (int)(w2 >> 8);
        float c[kSeg];
#pragma unroll
        for (int j = 0; j < kSeg; ++j) {   // rows >= `rows` hold finite stale data; never stored
          const uint32_t ro = (uint32_t)j * ROWB;
          c[j] = (lds(o0 + ro) + lds(o1 + ro)) + (lds(o2 + ro) + lds(o3 + ro));
        }
#pragma unroll
        for (int j = 0; j < kSeg; ++j) c[j] = fmaf(__shfl_down_sync(kFull, c[j], 1), lk1, c[j]);
        if (maxch > 2) {
#pragma unroll
          for (int j = 0; j < kSeg; ++j) c[j] = fmaf(__shfl_down_sync(kFull, c[j], 2), lk2, c[j]);
        }
        if (maxch > 4) {
#pragma unroll
          for (int j = 0; j < kSeg; ++j) c[j] = fmaf(__shfl_down_sync(kFull, c[j], 4), lk4, c[j]);
        }
        if (head) {
          uint32_t dsto = ot + 4u * (uint32_t)(me + rbase * C);
          const int32_t dstep = 4 * rsign * C;
#pragma unroll
          for (int j = 0; j < kSeg; ++j) {
            if (j < rows) {
              sts(dsto, c[j] * kappa);
              rs[j] += c[j];
            }
            dsto += dstep;
          }
        }
      }
      // blank posterior of a frame = Zm - (sum of its label posteriors): the posteriors of a frame
      // sum to Zm, which the recompute warp certifies at every segment boundary.  Transposed
      // reduction of the 8 row sums: 3 halving steps, then 2 full ones; lanes with lane % 4 == 0
      // end up with the total of row (lane >> 2).
      {
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0, up4 = (lane & 4) != 0;
        float h4[4], h2v[2], h1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = up16 ? rs[i] : rs[i + 4];
          h4[i] = (up16 ? rs[i + 4] : rs[i]) + __shfl_xor_sync(kFull, send, 16);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = up8 ? h4[i] : h4[i + 2];
          h2v[i] = (up8 ? h4[i + 2] : h4[i]) + __shfl_xor_sync(kFull, send, 8);
        }
        {
          const float send = up4 ? h2v[0] : h2v[1];
          h1 = (up4 ? h2v[1] : h2v[0]) + __shfl_xor_sync(kFull, send, 4);
        }
        h1 += __shfl_xor_sync(kFull, h1, 2);
        h1 += __shfl_xor_sync(kFull, h1, 1);
        const int row = (up16 ? 4 : 0) + (up8 ? 2 : 0) + (up4 ? 1 : 0);
        if ((lane & 3) == 0 && row < rows)
          sts(ot + 4u * (uint32_t)((rbase + rsign * row) * C) + blank_ofs, fmaxf(Zm - h1, 0.f) * kappa);
      }
      __syncwarp();
      if (lane == 0) bar_arrive(sm.bars, kBarAEmpty + d * kMaxAB + buf);   // products / partials consumed
      if (!dup) {
        const int n = rows * C;
        float* dst = gE + (size_t)seg * kSeg * C;
        const bool tma = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n & 3) == 0);
        if (a.fused) {
          // the softmax term of this tile is already in HBM (P<d> stored it; SDone says it landed)
          bar_wait(sm.bars, kBarSDone + d * kSD + (k2 % kSD), (uint32_t)(k2 / kSD) & 1u);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0)
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(ot),
                         "r"((uint32_t)n * 4u)
                         : "memory");
        } else if (tma) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(ot),
                         "r"((uint32_t)n * 4u)
                         : "memory");
        } else {
          const float* src = sm.out_gen + (size_t)((d * 2 + ob) * 2 + ux) * rawsz;
          for (int q = lane; q < n; q += 32) dst[q] = src[q];
        }
      }
      if (lane == 0) bulk_commit();   // one group per segment (possibly empty)
      __syncwarp();
      if (ux == 0) PROF_TL(6, d, k2);
    } else {
      if (lane == 0) bar_arrive(sm.bars, kBarAEmpty + d * kMaxAB + buf);
    }
  }
  PROF_MARK(1);
#ifdef WFST_PROFILE
  if (blockIdx.x == 0 && lane == 0) printf("X%d.%d cycles: wait_cfull %lld compute %lld\n", d, ux, pf_acc[0], pf_acc[1]);
  if (blockIdx.x == 0 && lane == 0 && d == 0 && ux == 0) {
    const long long t0 = g_tl[1][16];
    for (int k2 = 16; k2 < 24; ++k2)
      printf("TL k2=%d: L abuf %lld ptile %lld done %lld | RC start %lld done %lld | X start %lld done %lld | P produced %lld\n", k2,
             g_tl[0][k2] - t0, g_tl[1][k2] - t0, g_tl[2][k2] - t0, g_tl[3][k2] - t0, g_tl[4][k2] - t0, g_tl[5][k2] - t0,
             g_tl[6][k2] - t0, g_tl[7][k2] - t0);
  }
#endif
  if (lane == 0) bulk_wait_all<0>();
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0 && !dup) atomicOr(&a.hazard[bu], 8);
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <int K, int CS>
__global__ void __launch_bounds__(384, 1) ctc_pair_kernel(Args a) {
  extern __shared__ __align__(16) float smem_raw[];
  constexpr int Sp = 32 * K;
  const int warp = threadIdx.x >> 5;
  const int NT = 384;
  const int C = a.C;
  Ctx cx;
  cx.lane = threadIdx.x & 31;
  cx.T = a.T; cx.C = C; cx.nseg = a.nseg; cx.nA = a.nA; cx.NAB = a.NAB;
  cx.b[0] = 2 * blockIdx.x;
  cx.b[1] = min(2 * blockIdx.x + 1, a.B - 1);   // odd B: the last block runs its utterance twice
  cx.want_grad = a.gradE != nullptr;
  cx.rawsz = (uint32_t)((kSeg * C + 3) & ~3);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    cx.y[u] = a.targets + a.offsets[cx.b[u]];
    cx.L[u] = a.offsets[cx.b[u] + 1] - a.offsets[cx.b[u]];
  }
  const Layout lay = make_layout(K, C, CS, a.NAB);
  Smem sm;
  {
    const uint32_t base = smem_u32(smem_raw);
    sm.raw = base + 4u * (uint32_t)lay.raw;
    sm.out = base + 4u * (uint32_t)lay.out;
    sm.abuf = base + 4u * (uint32_t)lay.abuf;
    sm.bnd = base + 4u * (uint32_t)lay.bnd;
    sm.lexp = base + 4u * (uint32_t)lay.lexp;
    sm.ptile = base + 4u * (uint32_t)lay.ptile;
    sm.bars = base + 4u * (uint32_t)lay.bars;
    sm.zx = base + 4u * (uint32_t)lay.zx;
    sm.xtab = base + 4u * (uint32_t)lay.xtab;
    sm.colpos = reinterpret_cast<int*>(smem_raw + lay.colpos);
    sm.slotlab = reinterpret_cast<int*>(smem_raw + lay.slotlab);
    sm.hist = reinterpret_cast<int*>(smem_raw + lay.hist);
    sm.out_gen = smem_raw + lay.out;
  }

  // ------------------------------------------------------------------ setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumBars; ++i) {
      const bool two = (i >= kBarPEmpty && i < kBarPEmpty + 2 * kNB) || (i >= kBarAEmpty && i < kBarAEmpty + 2 * kMaxAB);
      bar_init(sm.bars, i, two ? 2u : 1u);
    }
    fence_barrier_init();
  }
  // zero everything up to the barriers: p-tile padding columns, gradient tiles (labels that do
  // not occur in the target keep a zero gradient), abar row pads, stale rows stay finite
  {
    float4* z = reinterpret_cast<float4*>(smem_raw);
    for (size_t k = threadIdx.x; k < lay.bars / 4; k += NT) z[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int k = threadIdx.x; k < 2 * (C + 4); k += NT) sm.hist[k] = 0;
  for (int k = threadIdx.x; k < 2 * (32 * kMaxPass + 8); k += NT) sm.slotlab[k] = -1;
  for (int k = threadIdx.x; k < 2 * 4 * 32 * kMaxPass; k += NT) sm.colpos[k] = -1;
  __syncthreads();
  // Counting sort of the target positions by label -> columns of the label-sorted order.
  // The order is organised in chunk slots of 4 columns; a label with n occurrences owns
  // ceil(n/4) consecutive slots that never straddle a multiple of 32 slots (one pass of the
  // reduction = 32 slots, one per lane).
  int flag[2] = {0, 0};
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    int* hist = sm.hist + u * (C + 4);
    int has_blank = 0, oob = 0;
    for (int n = threadIdx.x; n < cx.L[u]; n += NT) {
      const int yy = cx.y[u][n];
      if (yy < 0 || yy >= C) { oob = 1; continue; }
      atomicAdd(&hist[yy], 1);
      has_blank |= (yy == a.blank);
    }
    // a target that contains the blank label shares a gradient column between a label state
    // and the blank states: leave it to the log-semiring kernel (reason 1); so are labels
    // outside [0, C)
    if (__syncthreads_or(has_blank | oob)) flag[u] = 1;
    if (2 * cx.L[u] + 1 > Sp - 1) flag[u] = 1;
  }
  if (threadIdx.x < 2) {
    const int u = threadIdx.x;
    int* hist = sm.hist + u * (C + 4);
    int* slotlab = sm.slotlab + u * (32 * kMaxPass + 8);
    int slot = 0, maxch = 0;
    for (int c = 0; c < C; ++c) {
      const int cnt = hist[c];
      hist[c] = 0;
      if (cnt == 0) continue;
      const int nch = (cnt + 3) >> 2;
      if ((slot & 31) + nch > 32) slot = (slot + 31) & ~31;
      hist[c] = slot * 4;                // becomes the column cursor of label c
      for (int j = 0; j < nch && slot + j < 32 * kMaxPass; ++j) slotlab[slot + j] = c;
      slot += nch;
      maxch = max(maxch, nch);
    }
    hist[C] = slot;
    hist[C + 1] = maxch;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    int* hist = sm.hist + u * (C + 4);
    // layouts the reduction cannot hold go to the log-semiring kernel (reason 16)
    if (hist[C] > 32 * kMaxPass || hist[C + 1] > 8) flag[u] |= 16;
    if (!flag[u]) {
      int* colpos = sm.colpos + u * (4 * 32 * kMaxPass);
      for (int n = threadIdx.x; n < cx.L[u]; n += NT) {
        const int col = atomicAdd(&hist[cx.y[u][n]], 1);
        colpos[col] = n;
      }
    }
    cx.live[u] = flag[u] == 0;
    if (flag[u] && threadIdx.x == 0 && (u == 0 || cx.b[1] != cx.b[0])) atomicOr(&a.hazard[cx.b[u]], flag[u]);
  }
  __syncthreads();
  if (!cx.live[0] && !cx.live[1]) return;
  // a flagged utterance rides along with the target of the live one; its outputs are
  // suppressed (its emissions are still read, nothing is written)
  if (!cx.live[0]) { cx.y[0] = cx.y[1]; cx.L[0] = cx.L[1]; }
  if (!cx.live[1]) { cx.y[1] = cx.y[0]; cx.L[1] = cx.L[0]; }

  // warp -> scheduler partition is warp % 4 (measured: a live warp runs ~35 % slower while a
  // recompute warp is active on its partition, but every other placement tried -- both
  // reduction warps, or producer + reduction warp, next to the live warp -- was slower overall)
  //   partition 0: L0  RC0.1 X0.1      partition 2: RC0.0 X0.0 P0
  //   partition 1: L1  RC1.1 X1.1      partition 3: RC1.0 X1.0 P1
  if (warp < 2) role_live<K, CS>(a, sm, cx, warp);
  else if (warp < 6) role_rc<K, CS>(a, sm, cx, warp & 1, (warp - 2) >> 1);     // RC0.0 RC1.0 RC0.1 RC1.1
  else if (warp < 10) role_reduce<K>(a, sm, cx, warp & 1, (warp - 6) >> 1);    // X0.0 X1.0 X0.1 X1.1
  else role_producer<CS>(a, sm, cx, warp - 10);
}

static int pick_k(int max_target_len) {
  const int S = 2 * max_target_len + 1;
  const int ks[] = {4, 8, 12, 16};
  for (int k : ks)
    if (32 * k - 1 >= S) return k;
  return 0;
}

// p-tile row stride in pairs: the smallest instantiated value that holds C + 1 columns
static int pick_cs(int C) {
  if (C + 1 <= 32) return 32;
  if (C + 1 <= 64) return 64;
  if (C + 1 <= 128) return 128;
  return 0;
}

static int pick_nab(int K, int C, int CS) {
  for (int nab = kMaxAB; nab >= 2; --nab)
    if (make_layout(K, C, CS, nab).total * sizeof(float) <= 227 * 1024) return nab;
  return 0;
}

template <int K, int CS>
static int launch_kc(const Args& a, cudaStream_t st) {
  const size_t smem = make_layout(K, a.C, CS, a.NAB).total * sizeof(float);
  auto kern = ctc_pair_kernel<K, CS>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(a.B + 1) / 2, 384, smem, st>>>(a);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

template <int CS>
static int launch_c(const Args& a, int K, cudaStream_t st) {
  switch (K) {
    case 4: return launch_kc<4, CS>(a, st);
    case 8: return launch_kc<8, CS>(a, st);
    case 12: return launch_kc<12, CS>(a, st);
    case 16: return launch_kc<16, CS>(a, st);
  }
  set_error("no paired CTC instantiation for K=%d", K);
  return WFST_ERR_UNSUPPORTED;
}

}  // namespace pairk

#ifndef WFST_PAIR_CS
#error "compile with -DWFST_PAIR_CS=32|64|128|0 (0: host-side dispatch only)"
#endif

#if WFST_PAIR_CS == 32
int launch_ctc_pair_cs32(const pairk::Args& a, int K, cudaStream_t st) { return pairk::launch_c<32>(a, K, st); }
#elif WFST_PAIR_CS == 64
int launch_ctc_pair_cs64(const pairk::Args& a, int K, cudaStream_t st) { return pairk::launch_c<64>(a, K, st); }
#elif WFST_PAIR_CS == 128
int launch_ctc_pair_cs128(const pairk::Args& a, int K, cudaStream_t st) { return pairk::launch_c<128>(a, K, st); }
#else
int launch_ctc_pair_cs32(const pairk::Args& a, int K, cudaStream_t st);
int launch_ctc_pair_cs64(const pairk::Args& a, int K, cudaStream_t st);
int launch_ctc_pair_cs128(const pairk::Args& a, int K, cudaStream_t st);

// fused log-softmax mode moves gradient tiles with bulk copies only: every tile must be
// 16-byte aligned and a multiple of 16 bytes long
bool ctc_pair_fused_eligible(int T, int C, int max_target_len) {
  if (!ctc_pair_eligible(T, C, max_target_len)) return false;
  return ((size_t)T * C) % 4 == 0 && ((size_t)(T % pairk::kSeg) * C) % 4 == 0;
}

bool ctc_pair_eligible(int T, int C, int max_target_len) {
  if (T < 1) return false;
  const int K = pairk::pick_k(max_target_len), CS = pairk::pick_cs(C);
  if (K == 0 || CS == 0) return false;
  return pairk::pick_nab(K, C, CS) != 0;
}

size_t ctc_pair_workspace_bytes(int B, int T, int max_target_len) {
  const int K = pairk::pick_k(max_target_len);
  const int nseg = (T + pairk::kSeg - 1) / pairk::kSeg;
  return align_up((size_t)((B + 1) / 2) * nseg * 32 * (2 * K + 4) * sizeof(float), 256) +
         align_up((size_t)B * sizeof(int), 256);
}

int launch_ctc_pair(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                    int blank, int max_target_len, const float* grad_scale, float* z_out,
                    float* gradE, void* workspace, int** hazard_out, int fused, cudaStream_t st) {
  using namespace pairk;
  const int K = pick_k(max_target_len), CS = pick_cs(C);
  Args a{};
  a.E = E; a.targets = targets; a.offsets = offsets; a.B = B; a.T = T; a.C = C; a.blank = blank;
  a.grad_scale = grad_scale; a.z_out = z_out; a.gradE = gradE;
  a.nseg = (T + kSeg - 1) / kSeg;
  a.nA = a.nseg / 2;
  a.NAB = pick_nab(K, C, CS);
  a.fused = fused;
  a.ckpt = (float*)workspace;
  a.hazard = (int*)((char*)workspace +
                    align_up((size_t)((B + 1) / 2) * a.nseg * 32 * (2 * K + 4) * sizeof(float), 256));
  *hazard_out = a.hazard;
  WFST_CUDA_CHECK(cudaMemsetAsync(a.hazard, 0, (size_t)B * sizeof(int), st));
  switch (CS) {
    case 32: return launch_ctc_pair_cs32(a, K, st);
    case 64: return launch_ctc_pair_cs64(a, K, st);
    case 128: return launch_ctc_pair_cs128(a, K, st);
  }
  set_error("no paired CTC instantiation for C=%d", C);
  return WFST_ERR_UNSUPPORTED;
}
#endif

}  // namespace wfst
