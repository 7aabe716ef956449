// Chain-split scaled CTC forward+backward for sm_100a.  Replaces, for CTC, the reference's
// per-utterance create_ctc_graph -> intersect -> forward_score -> backward
// (criterions/ctc.py:15-29,40-51,78-81).
//
// One thread block per utterance.  The CTC chain (S = 2L+1 states, padded to Sp = 32*K*W
// slots) is split over W warps: global lane gl = 32*w + lane owns the K consecutive slots
// [K*gl, K*gl + K).  BOTH TIME DIRECTIONS of the utterance travel in one register: every
// value is a packed pair of floats (x = alpha, swept upwards from frame 0; y = beta~, swept
// downwards from frame T-1, stored at the mirrored slot j = Sp-2-s so that label states sit at
// odd slots in both orientations) and all arithmetic on it is add/mul/fma.f32x2 (SASS FADD2 /
// FMUL2 / FFMA2).  The recursion, in either orientation, is
//   v'[j] = (v[j] + v[j-1] + skip[j] * v[j-2]) * p_t[lab j],   p_t[c] = exp(E[t,c] - max_c E[t,c])
// (beta~_t(s) = p_t(lab s) * beta_t(s) obeys the alpha recursion on the reversed target and
// reversed time).  Within a warp the left neighbour's last slot arrives by shuffle; across
// warps the chain is SKEWED BY ONE STEP (= 8 frames): warp w-1 runs one step ahead of warp w
// and leaves the values of its last slot, frame by frame, in a small shared-memory ring
// (mbarrier full/empty per step).  Values are float32 mantissas with one power-of-two
// exponent per lane and direction, renormalised every 16 frames by a warp scan that takes
// the exponent of the neighbouring warp's last lane as its carry-in.
//
// Time is cut symmetrically: direction 0 owns the frames [0, Th), direction 1 the frames
// [Th, T), both in `nsd` steps of 8 frames counted from their own end of the utterance plus one
// partial step next to the meeting point, so the two directions always have the same number
// of steps.
//   phase 1: the W "live" warps sweep both halves at once, writing one checkpoint (their
//            registers) per step to the workspace.
//   meeting: Z = sum_s alpha_{Th-1}(s) * (successor sum of beta~_Th)(s).
//   phase 2: the live warps continue into the other half (alpha upwards through [Th, T), beta
//            downwards through [0, Th)) and store the pre-emission sums of the label states of
//            every frame into a ring of step buffers.  W "recompute" warps re-run the OPPOSITE
//            recursion over each step from the checkpoint phase 1 left there, rescaled once per
//            step so that  w * abar = posterior * Zm  with no further factor, and write the
//            products back in place.  One reduction warp per direction sums a frame's products
//            by label (lane = class, per-class lists of row offsets built at setup), takes the
//            blank posterior as Zm - sum(labels) and sends the [8, C] gradient tile to HBM with
//            a bulk async store.  Two producer warps fetch [8, C] emission tiles by bulk async
//            copy (TMA) and turn them into transposed p tiles.
// Nothing of size T x S leaves the SM except the checkpoints (one chain state per 8 frames).
//
// Robustness (as in ctc_pair.cuh): every step boundary is certified,
// sum_s v_live(s) * (successor sum of w)(s) = Z within 2e-5; a violation (float32 range
// exceeded, which can only lose mass or produce inf/NaN), Z out of range, a scale that does
// not fit, a blank / out-of-range label inside the target or a label histogram the reduction
// table cannot hold flag the utterance in `hazard`, and the log-semiring kernel (lattice.cuh)
// recomputes it on the GPU.  No CPU fallback.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "launchers.h"

namespace wfst {
namespace chaink {

typedef unsigned long long p2;   // two packed floats: low = direction 0, high = direction 1

#ifdef WFST_PROFILE
#define PROF_DECL long long pf_t0 = clock64(), pf_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF_MARK(i) do { long long pf_t1 = clock64(); pf_acc[i] += pf_t1 - pf_t0; pf_t0 = pf_t1; } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#endif

constexpr int kSeg = 8;               // frames per step / tile
constexpr int kEventEvery = 2;        // lanes are renormalised every kEventEvery steps
constexpr int kUndef = -(1 << 19);    // "no exponent": lane holds only zeros
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNR = 3;                // raw (TMA) staging slots per producer warp
constexpr int kRD = 4;                // depth of the warp-to-warp chain rings
constexpr int kRS = 1;                // recompute warp sets: set r takes the steps k2 = r (mod kRS)
constexpr int kMaxAB = 8;             // step buffers (abar / products), at most
constexpr int kMaxNB = 12;            // p tiles, at most
constexpr int kMaxList = 48;          // reduction table: sum over class rounds of the longest list
constexpr int kMaxW = 4;
constexpr int kRingPairs = 10;        // ring entry: 9 boundary pairs + {e0, e1}

struct Args {
  const float* E;
  const int* targets;
  const int* offsets;
  int B, T, C, blank;
  const float* grad_scale;
  float* z_out;     // [B] log Z
  float* gradE;     // [B, T, C] or null
  float* ckpt;      // [B][nsd][32 W][2K+4]
  int* hazard;      // [B]
  int nsd;          // steps per direction and phase
  int nfull;        // full (8-frame) steps per direction
  int r0, r1;       // frames of the partial step of direction 0 / 1 (next to the meeting point)
  int Th;           // first frame of direction 1's half
  int NAB, NB;
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
__device__ __forceinline__ p2 swap2(p2 a) { return pk(hi(a), lo(a)); }
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
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ p2 lds64(uint32_t a) {
  p2 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
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

// 2^x for x <= 0 (one MUFU; results below the normal range flush to zero)
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
constexpr int kBarPFull = 0;                              // [kMaxNB]       p tile ready (P0 + P1 -> live, RC)
constexpr int kBarPEmpty = kBarPFull + kMaxNB;            // [kMaxNB]       p tile released (count 2W)
constexpr int kBarTma = kBarPEmpty + kMaxNB;              // [2][kNR]       raw tiles landed
constexpr int kBarLFull = kBarTma + 2 * kNR;              // [kMaxW][kRD]   live chain ring entry written (w-1 -> w)
constexpr int kBarLEmpty = kBarLFull + kMaxW * kRD;       // [kMaxW][kRD]   ... consumed
constexpr int kBarRFull = kBarLEmpty + kMaxW * kRD;       // [kRS][kMaxW][kRD]   recompute chain rings
constexpr int kBarREmpty = kBarRFull + kRS * kMaxW * kRD;
constexpr int kBarAFull = kBarREmpty + kRS * kMaxW * kRD;       // [kMaxAB]       abar rows of a step stored (count W; live -> RC)
constexpr int kBarXFull = kBarAFull + kMaxAB;             // [kMaxAB]       products ready (count W; RC -> X)
constexpr int kBarAEmpty = kBarXFull + kMaxAB;            // [kMaxAB]       step buffer free (count 2; X -> live)
constexpr int kBarZ = kBarAEmpty + kMaxAB;                // Z published
constexpr int kNumBars = kBarZ + 1;

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
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WFSTC_BW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WFSTC_BD_%=;\n"
      "bra WFSTC_BW_%=;\n"
      "WFSTC_BD_%=:\n"
      "}\n" ::"r"(bars + 8u * idx), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- geometry ------------------------------------------------------------------------
template <int K, int W>
struct Geo {
  static constexpr int NL = 32 * W;            // lanes of the chain
  static constexpr int Sp = K * NL;            // slots
  static constexpr int HL = K / 2;             // label slots per lane
  // an abar row is two planes (direction 0, direction 1) of PADA zero words + SA words per lane
  static constexpr int SA = (HL & 1) ? HL : HL + 1;     // odd: conflict-free 32-bit accesses
  static constexpr int PADA = 4;
  static constexpr int ROWP = PADA + SA * NL + 4;       // words per plane
  static constexpr uint32_t PLANEB = 4u * ROWP;
  static constexpr uint32_t ROWB = 8u * ROWP;
  static constexpr uint32_t EXTB = 4u * (SA - HL + 1);  // distance from a lane block back to the previous block's last label
  static constexpr int SB = K;                 // boundary row: pairs per lane (touched once per step: bank conflicts do not matter)
  static constexpr int BNDP = 4 + SB * NL + 2;
  static constexpr uint32_t BNDB = 8u * BNDP;
  static constexpr int CKF = 2 * K + 4;        // checkpoint floats per lane
  static constexpr int NWARPS = W + kRS * W + 4;   // live, recompute sets, X0 X1, P0 P1
  static constexpr int NT = 32 * NWARPS;
};

// shared memory layout (in floats)
struct Layout {
  size_t raw, out, abuf, bnd, lexp, cert, ptile, ringL, ringR, bars, zx, ytab, xtab, hist, total;
  size_t zero_end;
};
template <int K, int W>
__host__ __device__ inline Layout make_layout(int C, int NAB, int NB) {
  using G = Geo<K, W>;
  Layout L;
  const size_t rawsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
  const size_t CP = (size_t)C + 1;
  size_t p = 0;
  L.raw = p;    p += 2 * (size_t)kNR * rawsz + 32;                  // [d][slot][8*C] (+ slack)
  L.out = p;    p += 2 * 2 * rawsz;                                 // [d][ob][8*C]
  L.abuf = p;   p += (size_t)NAB * kSeg * G::ROWP * 2;              // [buf][row][plane][word]
  L.bnd = p;    p += (size_t)NAB * G::BNDP * 2;                     // [buf][pair]: live state at the step boundary
  L.lexp = p;   p += (size_t)NAB * 2 * G::NL;                       // [buf][d][gl] (int)
  L.cert = p;   p += (size_t)NAB * W * 2;                           // [buf][w] (pair)
  L.ptile = p;  p += (size_t)NB * 2 * CP * 9 + 8;                   // [buf][d][col][9]
  p = (p + 3) & ~(size_t)3;
  L.ringL = p;  p += (size_t)(W + 1) * kRD * kRingPairs * 2;        // [w][slot]{9 boundary pairs, e0, e1}
  L.ringR = p;  p += (size_t)kRS * (W + 1) * kRD * kRingPairs * 2;
  p = (p + 3) & ~(size_t)3;
  L.zero_end = p;
  L.bars = p;   p += 2 * kNumBars;
  p = (p + 3) & ~(size_t)3;
  L.zx = p;     p += 32;   // Zm, eZ, ok, -, msum(double), zpart[kMaxW]{contrib, Emax}, endacc[2]
  L.ytab = p;   p += (size_t)G::Sp / 2 + 4;                         // targets of the utterance
  L.xtab = p;   p += (size_t)2 * kMaxList * 32 / 2;                 // [d][entry][lane] (u16 row offsets)
  L.hist = p;   p += (size_t)C + 16;                                // per-class counts; then per-round {nmax, base}
  L.total = p + 4;
  return L;
}

struct Smem {
  uint32_t raw, out, abuf, bnd, lexp, cert, ptile, ringL, ringR, bars, zx, xtab;
  int* ytab;
  int* hist;
  unsigned short* xtab_gen;
  float* out_gen;
};

struct Ctx {
  int lane, T, C, CP, L, b;
  int nsd, nfull, r0, r1, Th, NAB, NB;
  bool want_grad;
  uint32_t rawsz;
};

// phase-1 step k of direction d covers `rows` frames starting at `lo`
__device__ __forceinline__ int seg_rows(const Ctx& cx, int d, int k) { return k < cx.nfull ? kSeg : (d == 0 ? cx.r0 : cx.r1); }
__device__ __forceinline__ int seg_lo(const Ctx& cx, int d, int k) {
  if (d == 0) return kSeg * k;
  return k < cx.nfull ? cx.T - kSeg * (k + 1) : cx.Th;
}
// what component c works on at global tile kt (phase 1: kt < nsd, own half; phase 2: the
// other direction's steps, last one first)
__device__ __forceinline__ void comp_seg(const Ctx& cx, int c, int kt, int& lo_, int& rows) {
  const int d = kt < cx.nsd ? c : 1 - c;
  const int k = kt < cx.nsd ? kt : 2 * cx.nsd - 1 - kt;
  lo_ = seg_lo(cx, d, k);
  rows = seg_rows(cx, d, k);
}

// ---- per-lane topology --------------------------------------------------------------
template <int K>
struct Topo {
  uint32_t labofs[2][K / 2];  // byte offset in a p tile of row 0 of the label of odd slot 2q+1, per component
  uint32_t pbofs[2];          // ... of the blank
  p2 skipm[K / 2];            // 1 if the skip arc into odd slot 2q+1 exists
};

// component c runs orientation c (live) or 1-c (recompute: swapped = true); orientation o
// holds state s = j (o = 0) or s = Sp-2-j (o = 1) at slot j.  Component c reads plane c of a p tile.
template <int K, int W>
__device__ __forceinline__ void build_topo(Topo<K>& tp, const Ctx& cx, const int* ytab, int gl, bool swapped, int blank) {
  constexpr int Sp = Geo<K, W>::Sp;
#pragma unroll
  for (int q = 0; q < K / 2; ++q) {
    const int j = gl * K + 2 * q + 1;
    float sk[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int o = swapped ? 1 - c : c;
      const int s = o == 0 ? j : Sp - 2 - j;
      int col = cx.C;   // padding slots read the zero column
      sk[c] = 0.f;
      if (s >= 1 && s < 2 * cx.L + 1) {
        const int n = (s - 1) >> 1;
        col = min(max(ytab[n], 0), cx.C - 1);
        const int n2 = o == 0 ? n - 1 : n + 1;   // two positions earlier IN THIS ORIENTATION
        if (n2 >= 0 && n2 < cx.L && ytab[n2] != ytab[n]) sk[c] = 1.f;
      }
      tp.labofs[c][q] = 36u * (uint32_t)(c * cx.CP + col);
    }
    tp.skipm[q] = pk(sk[0], sk[1]);
  }
  tp.pbofs[0] = 36u * (uint32_t)blank;
  tp.pbofs[1] = 36u * (uint32_t)(cx.CP + blank);
}

template <int K>
struct PRow {
  p2 pl[K / 2];
  p2 pb;
};
template <int K>
struct TileAddr {
  uint32_t la[2][K / 2];
  uint32_t pba[2];
};
template <int K>
__device__ __forceinline__ TileAddr<K> tile_addr(const Topo<K>& tp, uint32_t pt) {
  TileAddr<K> t;
#pragma unroll
  for (int q = 0; q < K / 2; ++q) {
    t.la[0][q] = pt + tp.labofs[0][q];
    t.la[1][q] = pt + tp.labofs[1][q];
  }
  t.pba[0] = pt + tp.pbofs[0];
  t.pba[1] = pt + tp.pbofs[1];
  return t;
}
// row `it` of the tile (4 * it is an immediate when `it` is)
template <int K>
__device__ __forceinline__ PRow<K> load_prow(const TileAddr<K>& t, int it) {
  PRow<K> p;
  const uint32_t o = 4u * (uint32_t)it;
#pragma unroll
  for (int q = 0; q < K / 2; ++q) p.pl[q] = pk(lds(t.la[0][q] + o), lds(t.la[1][q] + o));
  p.pb = pk(lds(t.pba[0] + o), lds(t.pba[1] + o));
  return p;
}

// One frame.  v: with-emission values of the previous frame (own scale); on return this
// frame's with-emission values and, if WANT_ABAR, abar = the pre-emission sums of the label
// slots.  in1: the left neighbour's last slot, already converted to this lane's scale.
template <int K, bool WANT_ABAR>
__device__ __forceinline__ void step(p2 (&v)[K], p2 (&abar)[K / 2], const Topo<K>& tp, const PRow<K>& p, p2 in1) {
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    const p2 a1 = (i >= 1) ? v[i - 1] : in1;
    p2 s = add2(v[i], a1);
    if (i & 1) {
      const p2 a2 = (i >= 2) ? v[i - 2] : in1;
      s = fma2(tp.skipm[i >> 1], a2, s);
      if (WANT_ABAR) abar[i >> 1] = s;
      v[i] = mul2(s, p.pl[i >> 1]);
    } else {
      v[i] = mul2(s, p.pb);
    }
  }
}

// the left neighbour's last slot: by shuffle, lane 0 takes the value the previous warp of the
// chain left in the ring (zero for the first warp: its ring is never written)
__device__ __forceinline__ p2 left_in(p2 last, uint32_t ring_addr, int lane, p2 f) {
  p2 left = shfl_up2(last);
  const p2 bv = lds64(ring_addr);
  if (lane == 0) left = bv;
  return mul2(left, f);
}
// same with the ring value already in a register
__device__ __forceinline__ p2 left_in_reg(p2 last, p2 bv, int lane, p2 f) {
  p2 left = shfl_up2(last);
  if (lane == 0) left = bv;
  return mul2(left, f);
}

// Event: renormalise the lane (max mantissa in [1,2)) and make the lane exponents
// consistent from left to right (the direction mass flows):
//   * a lane that holds only zeros takes the exponent of its left neighbour, so mass
//     arriving during the next 16 frames arrives unscaled;
//   * a lane with own mass never sits more than D below its left neighbour, where D is
//     small enough that a wave crossing several lanes inside one 16-frame window cannot
//     overflow: D * (lanes crossed) + log2(3^16) < 127.
// This is the prefix composition of the maps x -> max(c, x - d) with (c, d) = (own
// exponent, D) or (undefined, 0), applied to the exponent Ein of the previous warp's last lane
// (undefined for the first warp).  "Undefined" is any value below kUndef / 2.
template <int K>
__device__ __forceinline__ void event2(p2 (&v)[K], int (&e)[2], p2& f, int lane, const int (&Ein)[2]) {
  constexpr int kChain = (2 * kSeg * kEventEvery + K - 1) / K + 1;   // lanes a wave can cross in a window
  constexpr int D = 96 / kChain;
  float m[2] = {lo(v[0]), hi(v[0])};
#pragma unroll
  for (int i = 1; i < K; ++i) {
    m[0] = fmaxf(m[0], lo(v[i]));
    m[1] = fmaxf(m[1], hi(v[i]));
  }
  // With m_l = number of lanes 0..l that hold mass, the composition collapses to a prefix
  // maximum:  E_l = max( max_{l' <= l, mass} (eown_l' + D m_l'),  Ein ) - D m_l .
  int ex[2], eown[2], val[2], dm[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    ex[u] = min(max((int)((__float_as_uint(m[u]) >> 23) & 0xffu) - 127, -126), 126);
    const bool has = m[u] > 0.f;
    if (!has) ex[u] = 0;
    eown[u] = has ? (defined_exp(e[u]) ? e[u] : 0) + ex[u] : kUndef;
    dm[u] = D * __popc(__ballot_sync(kFull, has) & (0xffffffffu >> (31 - lane)));
    val[u] = has ? eown[u] + dm[u] : 2 * kUndef;
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int u = 0; u < 2; ++u) val[u] = max(val[u], __shfl_up_sync(kFull, val[u], o));   // lanes < o get their own value back
  }
  int t[2];           // total power-of-two shift applied to the lane
  float fv[2];
  bool deep = false;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int cin = max(val[u], defined_exp(Ein[u]) ? Ein[u] : 2 * kUndef) - dm[u];
    const int E = defined_exp(cin) ? cin : kUndef;
    t[u] = -ex[u] + ((defined_exp(E) && defined_exp(eown[u])) ? eown[u] - E : 0);   // second term <= 0
    e[u] = E;
    int el = __shfl_up_sync(kFull, E, 1);
    if (lane == 0) el = Ein[u];
    fv[u] = (!defined_exp(el) || !defined_exp(E)) ? 0.f : pow2c(el - E);   // el - E <= D
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
__device__ __forceinline__ void ckpt_store(float* base, const p2 (&v)[K], const int (&e)[2]) {
  ulonglong2* p = reinterpret_cast<ulonglong2*>(base);
#pragma unroll
  for (int i = 0; i < K; i += 2) p[i >> 1] = make_ulonglong2(v[i], v[i + 1]);
  p[K >> 1] = make_ulonglong2(pk(__int_as_float(e[0]), __int_as_float(e[1])), 0ull);
}
// loads with the components exchanged (what direction 1 stored is what the recompute of
// direction 0's half needs, and vice versa)
template <int K>
__device__ __forceinline__ void ckpt_load_swapped(const float* base, p2 (&v)[K], int (&e)[2]) {
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(base);
#pragma unroll
  for (int i = 0; i < K; i += 2) {
    const ulonglong2 q = p[i >> 1];
    v[i] = swap2(q.x);
    v[i + 1] = swap2(q.y);
  }
  const ulonglong2 q = p[K >> 1];
  e[0] = __float_as_int(hi(q.x));
  e[1] = __float_as_int(lo(q.x));
}

// p-tile ring, as seen by a consumer warp that takes every tile from `first` on
struct PRing {
  int buf;
  uint32_t par;
  __device__ __forceinline__ void init(int first, int NB) { buf = first % NB; par = (uint32_t)(first / NB) & 1u; }
  __device__ __forceinline__ void next(int NB) {
    if (++buf == NB) { buf = 0; par ^= 1u; }
  }
};
__device__ __forceinline__ uint32_t ptile_wait(const Smem& sm, const Ctx& cx, const PRing& r) {
  bar_wait(sm.bars, kBarPFull + r.buf, r.par);
  return sm.ptile + 4u * (uint32_t)(r.buf * 2 * cx.CP * 9);
}
__device__ __forceinline__ void ptile_release(const Smem& sm, const Ctx& cx, PRing& r, uint32_t count) {
  __syncwarp();
  if (cx.lane == 0) bar_arrive(sm.bars, kBarPEmpty + r.buf, count);
  r.next(cx.NB);
}

// ---------------------------------------------------------------------------
// P<c>: producer of component c's p tiles (plane c of every tile): phase-1 tiles first; the
// phase-2 tiles only once Z is known to be usable.
// ---------------------------------------------------------------------------
struct ProducerState {
  int fetched, converted;
  uint32_t tma_phase, tma_used;
  int pbuf;            // p-tile buffer of the next tile
  uint32_t ppar;       // parity of its "empty" barrier
  double msum;
};

template <int W>
__device__ __forceinline__ void produce_range(const Args& a, const Smem& sm, const Ctx& cx, ProducerState& ps,
                                              const int c, const int kbeg, const int kcnt, const bool phase1) {
  const int lane = cx.lane, T = cx.T, C = cx.C;
  const uint32_t rawsz = cx.rawsz;
  const int fr = lane & 7, part = lane >> 3;                       // 8 frames x 4 label quarters
  const int c0 = (C * part) / 4, c1 = (C * (part + 1)) / 4;
  const uint32_t raw0 = sm.raw + 4u * (uint32_t)(c * kNR) * rawsz;
  const int tbar = kBarTma + c * kNR;
  const float* Eb = a.E + (size_t)cx.b * T * C;
  auto issue_raw = [&](int kt) {
    int lo_, rows;
    comp_seg(cx, c, kt, lo_, rows);
    const int slot = ps.fetched % kNR;
    const uint32_t bytes = (uint32_t)rows * C * 4u;
    const float* src = Eb + (size_t)lo_ * C;
    const uint32_t dst = raw0 + 4u * (uint32_t)slot * rawsz;
    const bool tma = rows > 0 && (bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    if (tma) {
      if (lane == 0) {
        bar_expect_tx(sm.bars, tbar + slot, bytes);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
            "l"(src), "r"(bytes), "r"(sm.bars + 8u * (tbar + slot))
            : "memory");
      }
      ps.tma_used |= 1u << slot;
    } else {
      for (int q = lane; q < rows * C; q += 32) sts(dst + 4u * q, __ldg(src + q));
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
    PROF_MARK(3);
    const int kt = kbeg + i;
    const int slot = ps.converted % kNR;
    int lo_, rows;
    comp_seg(cx, c, kt, lo_, rows);
    const int buf = ps.pbuf;
    if (kt >= cx.NB) {  // wait until the consumers have released this p-tile buffer
      bar_wait(sm.bars, kBarPEmpty + buf, ps.ppar);
    }
    if (++ps.pbuf == cx.NB) { ps.pbuf = 0; if (kt >= cx.NB) ps.ppar ^= 1u; }
    PROF_MARK(0);
    if ((ps.tma_used >> slot) & 1u) {
      bar_wait(sm.bars, tbar + slot, (ps.tma_phase >> slot) & 1u);
      ps.tma_phase ^= 1u << slot;
    }
    PROF_MARK(1);
    const bool live = fr < rows;
    const uint32_t er = raw0 + 4u * ((uint32_t)slot * rawsz + (uint32_t)(fr * C));
    // tile row = the step at which component c consumes frame fr (c = 0 ascends, c = 1 descends)
    const int trow = c == 0 ? fr : rows - 1 - fr;
    const uint32_t pt = sm.ptile + 4u * (uint32_t)((buf * 2 + c) * cx.CP * 9) + 4u * (uint32_t)(live ? trow : 0);
    // a row that is entirely -inf keeps p = 0 (dead frame); +inf / NaN rows surface through
    // the certificate
    float base;
    if (c1 - c0 <= 8) {
      // the lane's quarter of the row stays in registers between the two passes
      float ev[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) ev[q] = (live && c0 + q < c1) ? lds(er + 4u * (uint32_t)(c0 + q)) : kNegInf;
      float mx = fmaxf(fmaxf(fmaxf(ev[0], ev[1]), fmaxf(ev[2], ev[3])), fmaxf(fmaxf(ev[4], ev[5]), fmaxf(ev[6], ev[7])));
      mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, 16));
      base = (mx == kNegInf) ? 0.f : mx;
      const float nb = -base * 1.4426950408889634f;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (live && c0 + q < c1) sts(pt + 36u * (uint32_t)(c0 + q), ex2_fast(fmaf(ev[q], 1.4426950408889634f, nb)));
    } else {
      float mx = kNegInf;
      if (live)
        for (int cc = c0; cc < c1; ++cc) mx = fmaxf(mx, lds(er + 4u * cc));
      mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, 16));
      base = (mx == kNegInf) ? 0.f : mx;
      const float nb = -base * 1.4426950408889634f;
      if (live) {
#pragma unroll 4
        for (int cc = c0; cc < c1; ++cc)
          sts(pt + 36u * (uint32_t)cc, ex2_fast(fmaf(lds(er + 4u * cc), 1.4426950408889634f, nb)));
      }
    }
    if (live && part == 0 && phase1) ps.msum += (double)base;
    __syncwarp();
    if (lane == 0) bar_arrive(sm.bars, kBarPFull + buf);
    ++ps.converted;
  }
  PROF_MARK(2);
#ifdef WFST_PROFILE
  if (blockIdx.x == 0 && lane == 0)
    printf("P%d (phase1 %d) cycles: wait_pempty %lld wait_tma %lld convert %lld issue %lld\n", c, (int)phase1, pf_acc[0], pf_acc[1], pf_acc[2], pf_acc[3]);
#endif
}

template <int W>
__device__ __forceinline__ void role_producer(const Args& a, const Smem& sm, const Ctx& cx, const int c) {
  const int lane = cx.lane, nsd = cx.nsd;
  ProducerState ps;
  ps.fetched = 0; ps.converted = 0; ps.tma_phase = 0u; ps.tma_used = 0u;
  ps.pbuf = 0; ps.ppar = 0u;
  ps.msum = 0.0;
  produce_range<W>(a, sm, cx, ps, c, 0, nsd, true);
  // loss: log Z = log(Zm) + eZ ln2 + sum_t max_t; the two producers each hold the row maxima
  // of their phase-1 half
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ps.msum += __shfl_xor_sync(kFull, ps.msum, o);
  double* msh = reinterpret_cast<double*>(__cvta_shared_to_generic(sm.zx + 16u));
  if (c == 0 && lane == 0) msh[0] = ps.msum;
  named_sync(2, 64);
  bar_wait(sm.bars, kBarZ, 0u);
  const float Zm = lds(sm.zx);
  const int eZ = ldsi(sm.zx + 4u);
  const bool ok = lds(sm.zx + 8u) != 0.f;
  if (c == 1 && lane == 0)
    a.z_out[cx.b] = ok ? (float)(log((double)Zm) + (double)eZ * 0.6931471805599453 + (ps.msum + msh[0])) : kNegInf;
  if (!cx.want_grad || !ok) return;
  produce_range<W>(a, sm, cx, ps, c, nsd, nsd, false);
}

// ---------------------------------------------------------------------------
// live warp w of the chain
// ---------------------------------------------------------------------------
template <int K, int W>
__device__ __forceinline__ void role_live(const Args& a, const Smem& sm, const Ctx& cx, const int w) {
  using G = Geo<K, W>;
  constexpr int Sp = G::Sp, NL = G::NL, HL = G::HL;
  const int lane = cx.lane, gl = 32 * w + lane, nsd = cx.nsd, NAB = cx.NAB;
  const int S = 2 * cx.L + 1;
  float* ck = a.ckpt + ((size_t)cx.b * nsd * NL + gl) * G::CKF;
  Topo<K> tp;
  build_topo<K, W>(tp, cx, sm.ytab, gl, false, a.blank);

  p2 v[K], abar[HL];
  int e[2] = {kUndef, kUndef};
  p2 f = pk(0.f, 0.f);
  {
    // virtual pre-frame state: all mass on the start slot of each orientation
    float x[2][K];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int jstart = c == 0 ? 0 : Sp - 1 - S;
      const bool mine = (jstart / K == gl);
      const int jm = jstart % K;
#pragma unroll
      for (int i = 0; i < K; ++i) x[c][i] = (mine && jm == i) ? 1.f : 0.f;
      if (mine) e[c] = 0;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = pk(x[0][i], x[1][i]);
  }
  const uint32_t ring_in0 = sm.ringL + 8u * (uint32_t)(w * kRD * kRingPairs);
  const uint32_t ring_out0 = sm.ringL + 8u * (uint32_t)((w + 1) * kRD * kRingPairs);
  uint32_t rin = ring_in0, rout = ring_out0;
  PRing pr;
  pr.init(0, cx.NB);
  const bool has_partial = nsd > cx.nfull;

  // start of global step g: take the left warp's ring entry, renormalise if due, open my own entry
  PROF_DECL;
  auto step_begin = [&](int g, bool ev) {
    const int slot = g % kRD;
    rin = ring_in0 + 8u * (uint32_t)(slot * kRingPairs);
    rout = ring_out0 + 8u * (uint32_t)(slot * kRingPairs);
    PROF_MARK(0);
    if (w > 0) bar_wait(sm.bars, kBarLFull + w * kRD + slot, (uint32_t)(g / kRD) & 1u);
    PROF_MARK(1);
    if (ev) {
      int Ein[2] = {kUndef, kUndef};
      if (w > 0) { Ein[0] = ldsi(rin + 72u); Ein[1] = ldsi(rin + 76u); }
      event2<K>(v, e, f, lane, Ein);
    }
    PROF_MARK(2);
    if (w < W - 1) {
      if (g >= kRD) bar_wait(sm.bars, kBarLEmpty + (w + 1) * kRD + slot, (uint32_t)(g / kRD - 1) & 1u);
      PROF_MARK(3);
      if (lane == 31) {
        sts64(rout, v[K - 1]);
        stsi(rout + 72u, e[0]);
        stsi(rout + 76u, e[1]);
      }
    }
  };
  auto step_end = [&](int g) {
    const int slot = g % kRD;
    __syncwarp();
    if (lane == 0) {
      if (w < W - 1) bar_arrive(sm.bars, kBarLFull + (w + 1) * kRD + slot);
      if (w > 0) bar_arrive(sm.bars, kBarLEmpty + w * kRD + slot);
    }
  };
  // frames of a partial step: component c is frozen from frame rows_c on
  auto slow_frames = [&](const TileAddr<K>& ta, int rx, int ry, uint32_t ar, bool want_abar) {
    const int nfr = max(rx, ry);
#pragma unroll 1
    for (int it = 0; it < nfr; ++it) {
      const PRow<K> cur = load_prow<K>(ta, it);
      p2 old[K];
#pragma unroll
      for (int i = 0; i < K; ++i) old[i] = v[i];
      const p2 in1 = left_in(v[K - 1], rin + 8u * (uint32_t)it, lane, f);
      step<K, true>(v, abar, tp, cur, in1);
      const bool kx = it < rx, ky = it < ry;
#pragma unroll
      for (int i = 0; i < K; ++i) v[i] = pk(kx ? lo(v[i]) : lo(old[i]), ky ? hi(v[i]) : hi(old[i]));
      if (want_abar) {
#pragma unroll
        for (int q = 0; q < HL; ++q) {
          sts(ar + (uint32_t)it * G::ROWB + 4u * q, lo(abar[q]));
          sts(ar + G::PLANEB + (uint32_t)it * G::ROWB + 4u * q, hi(abar[q]));
        }
      }
      if (lane == 31) sts64(rout + 8u * (uint32_t)(it + 1), v[K - 1]);
    }
  };

  // ------------------------------------------------------------------ phase 1
  for (int g = 0; g < nsd; ++g) {
    step_begin(g, g % kEventEvery == 0);
    ckpt_store<K>(ck + (size_t)g * NL * G::CKF, v, e);
    const bool partial = has_partial && g == cx.nfull;
    PROF_MARK(0);
    const TileAddr<K> ta = tile_addr<K>(tp, ptile_wait(sm, cx, pr));
    PROF_MARK(4);
    if (!partial) {
      PRow<K> nx = load_prow<K>(ta, 0);
      p2 bvn = lds64(rin);
#pragma unroll
      for (int it = 0; it < kSeg; ++it) {
        const PRow<K> cur = nx;
        const p2 bv = bvn;
        if (it + 1 < kSeg) { nx = load_prow<K>(ta, it + 1); bvn = lds64(rin + 8u * (it + 1)); }
        const p2 in1 = left_in_reg(v[K - 1], bv, lane, f);
        step<K, false>(v, abar, tp, cur, in1);
        if (lane == 31) sts64(rout + 8u * (it + 1), v[K - 1]);
      }
    } else {
      slow_frames(ta, cx.r0, cx.r1, 0u, false);
    }
    ptile_release(sm, cx, pr, 2);   // no recompute warp reads phase-1 tiles
    step_end(g);
  }

  PROF_MARK(0);
#ifdef WFST_PROFILE
  if (blockIdx.x == 0 && lane == 0)
    printf("L%d phase 1 cycles: work %lld wait_ring %lld event %lld wait_ring_empty %lld wait_ptile %lld\n", w, pf_acc[0], pf_acc[1], pf_acc[2], pf_acc[3], pf_acc[4]);
  for (int i = 0; i < 8; ++i) pf_acc[i] = 0;
#endif
  // ------------------------------------------------------------------ meeting: Z
  // every live warp renormalises (consistent exponents for the successor sums below) and
  // publishes its state in the layout of a boundary row (buffer 0); then
  //   Z = sum over y-slots of (successor sum of beta~)(slot) * alpha(partner slot).
  step_begin(nsd, true);
  const uint32_t mybnd = 8u * (uint32_t)(4 + gl * G::SB);       // my slots in a boundary row
  const uint32_t myabar = 4u * (uint32_t)(G::PADA + gl * G::SA);
  {
#pragma unroll
    for (int i = 0; i < K; ++i) sts64(sm.bnd + mybnd + 8u * i, v[i]);
    stsi(sm.lexp + 4u * (uint32_t)gl, e[0]);
    stsi(sm.lexp + 4u * (uint32_t)(NL + gl), e[1]);
  }
  step_end(nsd);
  named_sync(1, 32 * W);
  {
    p2 bb[K];
    {
      const p2 in1 = left_in(v[K - 1], rin, lane, f);
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
    // partner of my slot i is slot K-2-i of lane NL-1-gl; of my slot K-1, the last slot of lane NL-2-gl
    const int pl = NL - 1 - gl;
    const uint32_t pblock = sm.bnd + 8u * (uint32_t)(4 + pl * G::SB);
    float pm = 0.f;
#pragma unroll
    for (int i = 0; i <= K - 2; ++i) pm = fmaf(hi(bb[i]), lo(lds64(pblock + 8u * (K - 2 - i))), pm);
    const float px = hi(bb[K - 1]) * lo(lds64(pblock - 8u));
    const int ea = ldsi(sm.lexp + 4u * (uint32_t)pl);
    const int eb = pl > 0 ? ldsi(sm.lexp + 4u * (uint32_t)(pl - 1)) : kUndef;
    int Em = kUndef, Ex = kUndef;
    if (pm > 0.f && defined_exp(e[1]) && defined_exp(ea)) Em = e[1] + ea;
    if (px > 0.f && defined_exp(e[1]) && defined_exp(eb)) Ex = e[1] + eb;
    int Emax = max(Em, Ex);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Emax = max(Emax, __shfl_xor_sync(kFull, Emax, o));
    float contrib = 0.f;
    if (defined_exp(Em)) contrib += pm * pow2c(Em - Emax);
    if (defined_exp(Ex)) contrib += px * pow2c(Ex - Emax);
    contrib = warp_sum(contrib);
    if (lane == 0) {
      sts(sm.zx + 32u + 8u * (uint32_t)w, contrib);
      stsi(sm.zx + 36u + 8u * (uint32_t)w, Emax);
    }
  }
  named_sync(1, 32 * W);
  if (w == 0 && lane == 0) {
    int Emax = kUndef;
    for (int i = 0; i < W; ++i) Emax = max(Emax, ldsi(sm.zx + 36u + 8u * (uint32_t)i));
    float tot = 0.f;
    for (int i = 0; i < W; ++i) {
      const int Ei = ldsi(sm.zx + 36u + 8u * (uint32_t)i);
      if (defined_exp(Ei)) tot += lds(sm.zx + 32u + 8u * (uint32_t)i) * pow2c(Ei - Emax);
    }
    const bool ok = defined_exp(Emax) && tot > 0.f && tot < 3.0e38f;
    int ex = 0;
    float Zm = 1.f;
    if (ok) {
      ex = (int)((__float_as_uint(tot) >> 23) & 0xffu) - 127;
      ex = min(max(ex, -126), 126);
      Zm = tot * pow2i(-ex);
    }
    sts(sm.zx, Zm);
    stsi(sm.zx + 4u, ok ? Emax + ex : 0);
    sts(sm.zx + 8u, ok ? 1.f : 0.f);
    // reason 2: infeasible or out of range -- the log-semiring kernel decides
    if (!ok) atomicOr(&a.hazard[cx.b], 2);
    bar_arrive(sm.bars, kBarZ);
  }
  bar_wait(sm.bars, kBarZ, 0u);
  const bool okz = lds(sm.zx + 8u) != 0.f;
  if (!cx.want_grad || !okz) return;

  // ------------------------------------------------------------------ phase 2
  uint32_t rpar = 0u;   // (k2 / NAB) & 1
  for (int k2 = 0, buf = 0; k2 < nsd; ++k2, rpar ^= (buf + 1 == NAB), buf = (buf + 1 == NAB) ? 0 : buf + 1) {
    const int g = nsd + 1 + k2;
    step_begin(g, g % kEventEvery == 0);
    PROF_MARK(0);
    if (k2 >= NAB) bar_wait(sm.bars, kBarAEmpty + buf, rpar ^ 1u);
    PROF_MARK(5);
    {
      const uint32_t le = sm.lexp + 4u * (uint32_t)(buf * 2 * NL + gl);
      stsi(le, e[0]);
      stsi(le + 4u * NL, e[1]);
      // state at the step boundary: the recompute warps check Z against it (certificate)
      const uint32_t bb = sm.bnd + (uint32_t)buf * G::BNDB + mybnd;
#pragma unroll
      for (int i = 0; i < K; ++i) sts64(bb + 8u * i, v[i]);
    }
    // component c continues through the other direction's steps, last one (the partial one) first
    const bool partial = has_partial && k2 == 0;
    PROF_MARK(0);
    const TileAddr<K> ta = tile_addr<K>(tp, ptile_wait(sm, cx, pr));
    PROF_MARK(4);
    const uint32_t ar = sm.abuf + (uint32_t)(buf * kSeg) * G::ROWB + myabar;
    if (!partial) {
      PRow<K> nx = load_prow<K>(ta, 0);
      p2 bvn = lds64(rin);
#pragma unroll
      for (int it = 0; it < kSeg; ++it) {
        const PRow<K> cur = nx;
        const p2 bv = bvn;
        if (it + 1 < kSeg) { nx = load_prow<K>(ta, it + 1); bvn = lds64(rin + 8u * (it + 1)); }
        const p2 in1 = left_in_reg(v[K - 1], bv, lane, f);
        step<K, true>(v, abar, tp, cur, in1);
#pragma unroll
        for (int q = 0; q < HL; ++q) {
          sts(ar + (uint32_t)it * G::ROWB + 4u * q, lo(abar[q]));
          sts(ar + G::PLANEB + (uint32_t)it * G::ROWB + 4u * q, hi(abar[q]));
        }
        if (lane == 31) sts64(rout + 8u * (it + 1), v[K - 1]);
      }
    } else {
      slow_frames(ta, cx.r1, cx.r0, ar, true);
    }
    __syncwarp();
    if (lane == 0) bar_arrive(sm.bars, kBarAFull + buf);
    ptile_release(sm, cx, pr, 1);
    step_end(g);
  }
  PROF_MARK(0);
#ifdef WFST_PROFILE
  if (blockIdx.x == 0 && lane == 0)
    printf("L%d phase 2 cycles: work %lld wait_ring %lld event %lld wait_ring_empty %lld wait_ptile %lld wait_aempty %lld\n", w, pf_acc[0], pf_acc[1], pf_acc[2], pf_acc[3], pf_acc[4], pf_acc[5]);
#endif
  // certificate, last leg: the sweep must arrive with total mass Z on the two slots that end the
  // chain in each orientation (the recompute warps check every earlier step boundary)
  {
    const float Zm = lds(sm.zx);
    const int eZ = ldsi(sm.zx + 4u);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int jend = c == 0 ? S - 1 : Sp - 2;          // last state of the chain in this orientation
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const int j = gl * K + i;
        const float x = c ? hi(v[i]) : lo(v[i]);
        if (j == jend || (j == jend - 1 && j >= 0)) part += x;
      }
      if (part != 0.f) part *= defined_exp(e[c]) ? pow2c(e[c] - eZ) : 0.f;
      const float tot = warp_sum(part);
      // the two end slots sit in at most two warps: a sum of two terms is order independent
      if (lane == 0 && tot != 0.f) atomicAdd(reinterpret_cast<float*>(__cvta_shared_to_generic(sm.zx + 64u + 4u * c)), tot);
    }
    named_sync(1, 32 * W);
    if (w == 0 && lane == 0) {
      const float t0 = lds(sm.zx + 64u), t1 = lds(sm.zx + 68u);
      if (!(fabsf(t0 - Zm) <= 2e-5f * Zm) || !(fabsf(t1 - Zm) <= 2e-5f * Zm)) atomicOr(&a.hazard[cx.b], 8);
    }
  }
}

// ---------------------------------------------------------------------------
// RC warp w: runs the opposite orientation of each component over the steps of the live
// warps' phase 2, from the checkpoints phase 1 wrote, against the live step order, and
// multiplies with the stored abar rows.  Before each step the lane is rescaled so that its
// effective exponent is eZ - e_live(partner lane): products need no further factor.
// ---------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void rc_frame(p2 (&w)[K], const Topo<K>& tp, const PRow<K>& cur, p2 in1, p2 h2, uint32_t arow,
                                         uint32_t extb, uint32_t planeb) {
  // partner of my odd slot i (<= K-3) is label slot (K-3-i)/2 of the partner block; of my slot
  // K-1, the last label slot of the block before it
  p2 av[K / 2], dummy[K / 2];
#pragma unroll
  for (int q = 0; q < K / 2 - 1; ++q) av[q] = pk(lds(arow + 4u * q), lds(arow + planeb + 4u * q));
  const p2 ext = pk(lds(arow - extb), lds(arow + planeb - extb));
  step<K, false>(w, dummy, tp, cur, in1);
#pragma unroll
  for (int q = 0; q < K / 2 - 1; ++q) {
    const p2 pr = mul2(w[K - 3 - 2 * q], av[q]);
    sts(arow + 4u * q, lo(pr));
    sts(arow + planeb + 4u * q, hi(pr));
  }
  const p2 pe = mul2(mul2(w[K - 1], ext), h2);
  sts(arow - extb, lo(pe));
  sts(arow + planeb - extb, hi(pe));
}

template <int K, int W>
__device__ __forceinline__ void role_rc(const Args& a, const Smem& sm, const Ctx& cx, const int w, const int set) {
  using G = Geo<K, W>;
  constexpr int NL = G::NL;
  const int lane = cx.lane, gl = 32 * w + lane, nsd = cx.nsd, NAB = cx.NAB;
  bar_wait(sm.bars, kBarZ, 0u);      // phase 1 (and every checkpoint) is complete
  const bool ok = lds(sm.zx + 8u) != 0.f;
  if (!cx.want_grad || !ok) return;
  const int eZ = ldsi(sm.zx + 4u);
  Topo<K> tp;
  build_topo<K, W>(tp, cx, sm.ytab, gl, true, a.blank);
  const float* ck = a.ckpt + ((size_t)cx.b * nsd * NL + gl) * G::CKF;
  const int pl = NL - 1 - gl;
  const uint32_t pabar = 4u * (uint32_t)(G::PADA + pl * G::SA);   // partner block in a plane of an abar row
  const uint32_t pbnd = 8u * (uint32_t)(4 + pl * G::SB);          // partner block in a boundary row
  const uint32_t ring_in0 = sm.ringR + 8u * (uint32_t)((set * (W + 1) + w) * kRD * kRingPairs);
  const uint32_t ring_out0 = sm.ringR + 8u * (uint32_t)((set * (W + 1) + w + 1) * kRD * kRingPairs);
  const int rbar = set * kMaxW * kRD;   // this set's chain barriers
  int bad = 0;   // reason 4: scale overflow when pairing live and recomputed values
  p2 wv[K];
  int ew[2];
  if (set < nsd) ckpt_load_swapped<K>(ck + (size_t)(nsd - 1 - set) * NL * G::CKF, wv, ew);
  PRing pr;
  pr.init(nsd + set, cx.NB);
  const bool has_partial = nsd > cx.nfull;
  PROF_DECL;
  // this set's j-th step is k2 = set + kRS j
  int buf = set % NAB;
  uint32_t rpar = (uint32_t)(set / NAB) & 1u;   // (k2 / NAB) & 1
  for (int k2 = set, j = 0; k2 < nsd; k2 += kRS, ++j) {
    const int slot = j % kRD;
    const uint32_t rin = ring_in0 + 8u * (uint32_t)(slot * kRingPairs);
    const uint32_t rout = ring_out0 + 8u * (uint32_t)(slot * kRingPairs);
    const bool partial = has_partial && k2 == 0;
    const int rx = partial ? cx.r1 : kSeg, ry = partial ? cx.r0 : kSeg;
    PROF_MARK(0);
    bar_wait(sm.bars, kBarAFull + buf, rpar);
    PROF_MARK(1);
    // scales
    float gsc[2], hsc[2], frs[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t le = sm.lexp + 4u * (uint32_t)((buf * 2 + c) * NL);
      const int ep = ldsi(le + 4u * (uint32_t)pl);                          // partner lane
      const int ex = pl > 0 ? ldsi(le + 4u * (uint32_t)(pl - 1)) : kUndef;  // partner of slot K-1
      const int el = gl > 0 ? ldsi(le + 4u * (uint32_t)(pl + 1)) : kUndef;  // partner of my left neighbour
      gsc[c] = 0.f; hsc[c] = 0.f; frs[c] = 0.f;
      if (defined_exp(ep)) {
        if (defined_exp(ew[c])) {
          const int dd = ew[c] + ep - eZ;
          if (dd > 126) bad |= 4;
          else gsc[c] = pow2c(dd);
        }
        if (defined_exp(ex)) hsc[c] = pow2c(ex - ep);     // <= 2^D by the event invariant
        if (defined_exp(el)) frs[c] = pow2c(ep - el);     // <= 2^D likewise
      }
    }
    {
      const p2 g2 = pk(gsc[0], gsc[1]);
#pragma unroll
      for (int i = 0; i < K; ++i) wv[i] = mul2(wv[i], g2);
    }
    const p2 f2 = pk(frs[0], frs[1]), h2 = pk(hsc[0], hsc[1]);
    // chain ring: my left neighbour's entry of this step; open my own
    PROF_MARK(0);
    if (w > 0) bar_wait(sm.bars, kBarRFull + rbar + w * kRD + slot, (uint32_t)(j / kRD) & 1u);
    if (w < W - 1) {
      if (j >= kRD) bar_wait(sm.bars, kBarREmpty + rbar + (w + 1) * kRD + slot, (uint32_t)(j / kRD - 1) & 1u);
      if (lane == 31) sts64(rout, wv[K - 1]);
    }
    PROF_MARK(2);
    const TileAddr<K> ta = tile_addr<K>(tp, ptile_wait(sm, cx, pr));
    PROF_MARK(3);
    const uint32_t ar = sm.abuf + (uint32_t)(buf * kSeg) * G::ROWB + pabar;
    const int nfr = max(rx, ry);
    if (!partial) {
      // against the live step order
      PRow<K> nx = load_prow<K>(ta, kSeg - 1);
      p2 bvn = lds64(rin);
#pragma unroll
      for (int it = kSeg - 1; it >= 0; --it) {
        const PRow<K> cur = nx;
        const p2 bv = bvn;
        if (it > 0) { nx = load_prow<K>(ta, it - 1); bvn = lds64(rin + 8u * (kSeg - it)); }
        const p2 in1 = left_in_reg(wv[K - 1], bv, lane, f2);
        rc_frame<K>(wv, tp, cur, in1, h2, ar + (uint32_t)it * G::ROWB, G::EXTB, G::PLANEB);
        if (lane == 31) sts64(rout + 8u * (kSeg - it), wv[K - 1]);
      }
    } else {
#pragma unroll 1
      for (int it = nfr - 1; it >= 0; --it) {
        const PRow<K> cur = load_prow<K>(ta, it);
        p2 old[K];
#pragma unroll
        for (int i = 0; i < K; ++i) old[i] = wv[i];
        const p2 in1 = left_in(wv[K - 1], rin + 8u * (uint32_t)(nfr - 1 - it), lane, f2);
        rc_frame<K>(wv, tp, cur, in1, h2, ar + (uint32_t)it * G::ROWB, G::EXTB, G::PLANEB);
        const bool kx = it < rx, ky = it < ry;
#pragma unroll
        for (int i = 0; i < K; ++i) wv[i] = pk(kx ? lo(wv[i]) : lo(old[i]), ky ? hi(wv[i]) : hi(old[i]));
        if (lane == 31) sts64(rout + 8u * (uint32_t)(nfr - it), wv[K - 1]);
      }
    }
    {
      // certificate: sum_s v_live(s) * (successor sum of w)(s) at the step boundary must be Z
      // (float32 range can only be exceeded by losing mass or producing inf / NaN)
      const p2 in1 = left_in(wv[K - 1], rin + 8u * (uint32_t)nfr, lane, f2);
      const uint32_t bb = sm.bnd + (uint32_t)buf * G::BNDB + pbnd;
      p2 acc = pk(0.f, 0.f);
#pragma unroll
      for (int i = K - 1; i >= 0; --i) {
        const p2 a1 = (i >= 1) ? wv[i - 1] : in1;
        p2 sx = add2(wv[i], a1);
        if (i & 1) {
          const p2 a2 = (i >= 2) ? wv[i - 2] : in1;
          sx = fma2(tp.skipm[i >> 1], a2, sx);
        }
        if (i == K - 1) acc = fma2(mul2(sx, h2), lds64(bb - 8u), acc);
        else acc = fma2(sx, lds64(bb + 8u * (K - 2 - i)), acc);
      }
      const float t0 = warp_sum(lo(acc)), t1 = warp_sum(hi(acc));
      if (lane == 0) sts64(sm.cert + 8u * (uint32_t)(buf * W + w), pk(t0, t1));
    }
    // next checkpoint (consumed at the top of the next iteration)
    if (k2 + kRS < nsd) ckpt_load_swapped<K>(ck + (size_t)(nsd - 1 - kRS - k2) * NL * G::CKF, wv, ew);
    __syncwarp();
    if (lane == 0) {
      bar_arrive(sm.bars, kBarXFull + buf);
      if (w < W - 1) bar_arrive(sm.bars, kBarRFull + rbar + (w + 1) * kRD + slot);
      if (w > 0) bar_arrive(sm.bars, kBarREmpty + rbar + w * kRD + slot);
    }
    ptile_release(sm, cx, pr, 1);
#pragma unroll
    for (int r = 1; r < kRS; ++r) pr.next(cx.NB);   // the other sets' tiles
    buf += kRS;
    if (buf >= NAB) { buf -= NAB; rpar ^= 1u; }
  }
  PROF_MARK(0);
#ifdef WFST_PROFILE
  if (blockIdx.x == 0 && lane == 0)
    printf("RC%d cycles: work %lld wait_afull %lld wait_ring %lld wait_ptile %lld\n", w, pf_acc[0], pf_acc[1], pf_acc[2], pf_acc[3]);
#endif
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0) atomicOr(&a.hazard[cx.b], bad);
}

// ---------------------------------------------------------------------------
// X<c>: per-label reduction of a step's posteriors of component c + gradient tile store.
// Lane = class (classes beyond 32 in further rounds); entry i of a round holds, per lane, the
// row offset of the i-th occurrence of the lane's class (or of a zero pad).
// ---------------------------------------------------------------------------
constexpr int kRegList = 16;   // entries of a class list the reduction keeps in registers

template <int K, int W>
__device__ __forceinline__ void role_reduce(const Args& a, const Smem& sm, const Ctx& cx, const int c) {
  using G = Geo<K, W>;
  const int lane = cx.lane, nsd = cx.nsd, T = cx.T, C = cx.C, NAB = cx.NAB;
  const int rounds = (C + 31) >> 5;
  const int* rinfo = sm.hist + C;   // per round {nmax, base}
  // One round of classes: the lane's list of row offsets lives in registers, ordered (while
  // phase 1 runs) so that, slot by slot, the lanes of the warp read distinct banks: every slot
  // each lane proposes one of its next three entries, the lowest lane wins a contested bank.
  const int nmax0 = rinfo[0];
  bool reg_lists = rounds == 1;
  uint32_t offs[kRegList];
  int nslots = 0;
  if (reg_lists) {
    const uint32_t tb = sm.xtab + 2u * (uint32_t)(c * kMaxList * 32 + lane);   // entry k at tb + 64 k
    const int n = lane < C ? sm.hist[lane] : 0;
    int done = 0;
#pragma unroll
    for (int sl = 0; sl < kRegList; ++sl) {
      uint32_t taken = 0u, mine = 0xffffu;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int k = done + t;
        const bool cand = mine == 0xffffu && k < n;
        const uint32_t off = cand ? lds_u16(tb + 64u * (uint32_t)k) : 0u;
        const uint32_t bank = (off >> 2) & 31u;
        const bool okb = cand && !((taken >> bank) & 1u);
        const unsigned peers = __match_any_sync(kFull, okb ? bank : 32u + (uint32_t)lane);
        const bool win = okb && (__ffs(peers) - 1) == lane;
        if (win) {
          mine = off;
          if (t > 0) {   // swap it with the first entry still to be placed: those stay contiguous, the table a permutation
            const uint32_t first = lds_u16(tb + 64u * (uint32_t)done);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(tb + 64u * (uint32_t)k), "h"((unsigned short)first) : "memory");
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(tb + 64u * (uint32_t)done), "h"((unsigned short)off) : "memory");
          }
        }
        taken |= __reduce_or_sync(kFull, win ? (1u << bank) : 0u);
      }
      if (mine != 0xffffu) ++done;
      offs[sl] = mine;
      if (__any_sync(kFull, mine != 0xffffu)) nslots = sl + 1;
    }
    reg_lists = __all_sync(kFull, done == n);   // else: the table (a permutation of itself) is walked from shared memory
  } else {
#pragma unroll
    for (int i = 0; i < kRegList; ++i) offs[i] = 0xffffu;
  }
  (void)nmax0;
  bar_wait(sm.bars, kBarZ, 0u);
  const bool ok = lds(sm.zx + 8u) != 0.f;
  if (!cx.want_grad || !ok) return;
  const uint32_t rawsz = cx.rawsz;
  const float Zm = lds(sm.zx);
  const float kappa = -(a.grad_scale ? a.grad_scale[cx.b] : 1.f) / Zm;
  const bool has_partial = nsd > cx.nfull;
  int bad = 0;
  float* gE = a.gradE + (size_t)cx.b * T * C;
  PROF_DECL;
  uint32_t rpar = 0u;
  for (int k2 = 0, buf = 0; k2 < nsd; ++k2, rpar ^= (buf + 1 == NAB), buf = (buf + 1 == NAB) ? 0 : buf + 1) {
    // component c works through the other direction's steps, the partial one first
    const int kk = nsd - 1 - k2;
    const int rows = (has_partial && k2 == 0) ? (c == 0 ? cx.r1 : cx.r0) : kSeg;
    const int lo_ = seg_lo(cx, 1 - c, kk);
    const int ob = k2 & 1;
    PROF_MARK(0);
    bar_wait(sm.bars, kBarXFull + buf, rpar);
    PROF_MARK(1);
    {
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const p2 t = lds64(sm.cert + 8u * (uint32_t)(buf * W + i));
        tot += c ? hi(t) : lo(t);
      }
      if (!(fabsf(tot - Zm) <= 2e-5f * Zm)) bad = 8;
    }
    if (lane == 0) bulk_wait_read<1>();   // the store that last read this out buffer is done
    __syncwarp();
    const uint32_t ab = sm.abuf + (uint32_t)(buf * kSeg) * G::ROWB + (uint32_t)c * G::PLANEB;
    const uint32_t ot = sm.out + 4u * (uint32_t)((c * 2 + ob) * rawsz);
    // buffer row j holds the frame of step j: frame row r = j (c = 0) or rows-1-j (c = 1)
    const int rsign = c == 0 ? 1 : -1, rbase = c == 0 ? 0 : rows - 1;
    float rs[kSeg];     // per-row sum of the label posteriors of this lane's classes
    if (reg_lists) {
#pragma unroll
      for (int j = 0; j < kSeg; ++j) rs[j] = 0.f;
#pragma unroll
      for (int i0 = 0; i0 < kRegList; i0 += 4) {
        if (i0 < nslots) {   // warp-uniform; lanes without an entry in a slot read the zero word in front of the plane
          float t[4][kSeg];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t o = ab + (offs[i0 + i] != 0xffffu ? offs[i0 + i] : 0u);
#pragma unroll
            for (int j = 0; j < kSeg; ++j) t[i][j] = lds(o + (uint32_t)j * G::ROWB);   // rows >= `rows` hold finite stale data; never stored
          }
#pragma unroll
          for (int j = 0; j < kSeg; ++j) rs[j] += (t[0][j] + t[1][j]) + (t[2][j] + t[3][j]);
        }
      }
      if (lane < C && lane != a.blank) {
        uint32_t dsto = ot + 4u * (uint32_t)(lane + rbase * C);
        const int32_t dstep = 4 * rsign * C;
#pragma unroll
        for (int j = 0; j < kSeg; ++j) {
          if (j < rows) sts(dsto, rs[j] * kappa);
          dsto += dstep;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < kSeg; ++j) rs[j] = 0.f;
      for (int r = 0; r < rounds; ++r) {
        const int nmax = rinfo[2 * r], base = rinfo[2 * r + 1];
        const int cls = 32 * r + lane;
        float acc[kSeg];
#pragma unroll
        for (int j = 0; j < kSeg; ++j) acc[j] = 0.f;
        uint32_t xt = sm.xtab + 2u * (uint32_t)((c * kMaxList + base) * 32 + lane);
#pragma unroll 2
        for (int i = 0; i < nmax; ++i, xt += 64u) {
          const uint32_t o = ab + lds_u16(xt);
#pragma unroll
          for (int j = 0; j < kSeg; ++j) acc[j] += lds(o + (uint32_t)j * G::ROWB);
        }
        if (cls < C && cls != a.blank) {
          uint32_t dsto = ot + 4u * (uint32_t)(cls + rbase * C);
          const int32_t dstep = 4 * rsign * C;
#pragma unroll
          for (int j = 0; j < kSeg; ++j) {
            if (j < rows) sts(dsto, acc[j] * kappa);
            dsto += dstep;
          }
        }
#pragma unroll
        for (int j = 0; j < kSeg; ++j) rs[j] += acc[j];
      }
    }
    // blank posterior of a frame = Zm - (sum of its label posteriors): the posteriors of a frame
    // sum to Zm, which the recompute warps certify at every step boundary.  Transposed
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
        sts(ot + 4u * (uint32_t)((rbase + rsign * row) * C + a.blank), fmaxf(Zm - h1, 0.f) * kappa);
    }
    __syncwarp();
    if (lane == 0) bar_arrive(sm.bars, kBarAEmpty + buf);   // products consumed
    if (rows > 0) {
      const int n = rows * C;
      float* dst = gE + (size_t)lo_ * C;
      const bool tma = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n & 3) == 0);
      if (tma) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(ot),
                       "r"((uint32_t)n * 4u)
                       : "memory");
      } else {
        const float* src = sm.out_gen + (size_t)(c * 2 + ob) * rawsz;
        for (int q = lane; q < n; q += 32) dst[q] = src[q];
      }
    }
    if (lane == 0) bulk_commit();   // one group per step (possibly empty)
    __syncwarp();
  }
  PROF_MARK(0);
#ifdef WFST_PROFILE
  if (blockIdx.x == 0 && lane == 0) printf("X%d cycles: work %lld wait_xfull %lld\n", c, pf_acc[0], pf_acc[1]);
#endif
  if (lane == 0) bulk_wait_all<0>();
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0) atomicOr(&a.hazard[cx.b], 8);
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <int K, int W>
__global__ void __launch_bounds__(Geo<K, W>::NT, (Geo<K, W>::NT <= 384 ? 2 : 1)) ctc_chain_kernel(Args a) {
  extern __shared__ __align__(16) float smem_raw[];
#ifdef WFST_PROFILE
  const long long pf_setup0 = clock64();
#endif
  using G = Geo<K, W>;
  constexpr int NT = G::NT;
  const int warp = threadIdx.x >> 5;
  const int C = a.C;
  Ctx cx;
  cx.lane = threadIdx.x & 31;
  cx.T = a.T; cx.C = C; cx.CP = C + 1;
  cx.nsd = a.nsd; cx.nfull = a.nfull; cx.r0 = a.r0; cx.r1 = a.r1; cx.Th = a.Th;
  cx.NAB = a.NAB; cx.NB = a.NB;
  cx.b = blockIdx.x;
  cx.want_grad = a.gradE != nullptr;
  cx.rawsz = (uint32_t)((kSeg * C + 3) & ~3);
  const int* y = a.targets + a.offsets[cx.b];
  cx.L = a.offsets[cx.b + 1] - a.offsets[cx.b];
  const Layout lay = make_layout<K, W>(C, a.NAB, a.NB);
  Smem sm;
  {
    const uint32_t base = smem_u32(smem_raw);
    sm.raw = base + 4u * (uint32_t)lay.raw;
    sm.out = base + 4u * (uint32_t)lay.out;
    sm.abuf = base + 4u * (uint32_t)lay.abuf;
    sm.bnd = base + 4u * (uint32_t)lay.bnd;
    sm.lexp = base + 4u * (uint32_t)lay.lexp;
    sm.cert = base + 4u * (uint32_t)lay.cert;
    sm.ptile = base + 4u * (uint32_t)lay.ptile;
    sm.ringL = base + 4u * (uint32_t)lay.ringL;
    sm.ringR = base + 4u * (uint32_t)lay.ringR;
    sm.bars = base + 4u * (uint32_t)lay.bars;
    sm.zx = base + 4u * (uint32_t)lay.zx;
    sm.xtab = base + 4u * (uint32_t)lay.xtab;
    sm.ytab = reinterpret_cast<int*>(smem_raw + lay.ytab);
    sm.hist = reinterpret_cast<int*>(smem_raw + lay.hist);
    sm.xtab_gen = reinterpret_cast<unsigned short*>(smem_raw + lay.xtab);
    sm.out_gen = smem_raw + lay.out;
  }

  // ------------------------------------------------------------------ setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumBars; ++i) {
      uint32_t cnt = 1u;
      if (i >= kBarPFull && i < kBarPFull + kMaxNB) cnt = 2u;
      else if (i >= kBarPEmpty && i < kBarPEmpty + kMaxNB) cnt = 2u * W;
      else if (i >= kBarAFull && i < kBarAFull + kMaxAB) cnt = W;
      else if (i >= kBarXFull && i < kBarXFull + kMaxAB) cnt = W;
      else if (i >= kBarAEmpty && i < kBarAEmpty + kMaxAB) cnt = 2u;
      bar_init(sm.bars, i, cnt);
    }
    fence_barrier_init();
  }
  // zero everything up to the barriers: p-tile padding columns, row pads, rings of the first
  // warps (never written), stale rows stay finite
  {
    float4* z = reinterpret_cast<float4*>(smem_raw);
    for (size_t k = threadIdx.x; k < lay.zero_end / 4; k += NT) z[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = threadIdx.x; k < 32; k += NT) smem_raw[lay.zx + k] = 0.f;
  }
  int flag = 0;
  {
    int has_bad = 0;
    for (int n = threadIdx.x; n < cx.L && n < G::Sp / 2; n += NT) {
      const int yy = y[n];
      sm.ytab[n] = yy;
      if (yy < 0 || yy >= C || yy == a.blank) has_bad = 1;
    }
    // a target that contains the blank label shares a gradient column between a label state
    // and the blank states: leave it to the log-semiring kernel (reason 1); so are labels
    // outside [0, C)
    if (__syncthreads_or(has_bad)) flag = 1;
    if (2 * cx.L + 1 > G::Sp - 1) flag = 1;
  }
  if (!flag) {
    // per-class counts (integer atomics: order independent)
    for (int cc = threadIdx.x; cc < C + 16; cc += NT) sm.hist[cc] = 0;
    __syncthreads();
    for (int n = threadIdx.x; n < cx.L; n += NT) atomicAdd(&sm.hist[sm.ytab[n]], 1);
    __syncthreads();
    const int rounds = (C + 31) >> 5;
    if (warp == 0) {
      int base = 0;
      for (int r = 0; r < rounds; ++r) {
        const int cc = 32 * r + cx.lane;
        const int nm = __reduce_max_sync(kFull, cc < C ? sm.hist[cc] : 0);
        if (cx.lane == 0) { sm.hist[C + 2 * r] = nm; sm.hist[C + 2 * r + 1] = base; }
        base += nm;
      }
      if (cx.lane == 0) sm.hist[C + 2 * rounds] = base;
    }
    __syncthreads();
    // tables the reduction cannot hold go to the log-semiring kernel (reason 16)
    if (sm.hist[C + 2 * rounds] > kMaxList) flag = 16;
    if (!flag) {
      // position n becomes entry (number of earlier positions with the same label) of its class:
      // a deterministic order, so the sums of the reduction do not depend on scheduling
      for (int n = threadIdx.x; n < cx.L; n += NT) {
        const int cc = sm.ytab[n];
        int rank = 0;
#pragma unroll 4
        for (int m = 0; m < n; ++m) rank += (sm.ytab[m] == cc);
        const int base = sm.hist[C + 2 * (cc >> 5) + 1], ln = cc & 31;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const int j = d == 0 ? 2 * n + 1 : G::Sp - 3 - 2 * n;
          const int off = 4 * (G::PADA + (j / K) * G::SA + (j % K) / 2);
          sm.xtab_gen[(d * kMaxList + base + rank) * 32 + ln] = (unsigned short)off;
        }
      }
      for (int cc = threadIdx.x; cc < 32 * rounds; cc += NT) {
        const int r = cc >> 5, ln = cc & 31;
        const int nm = sm.hist[C + 2 * r], base = sm.hist[C + 2 * r + 1];
        for (int i = cc < C ? sm.hist[cc] : 0; i < nm; ++i) {
          sm.xtab_gen[(base + i) * 32 + ln] = 0;                 // word 0 of a plane is always zero
          sm.xtab_gen[(kMaxList + base + i) * 32 + ln] = 0;
        }
      }
    }
  }
  if (flag && threadIdx.x == 0) atomicOr(&a.hazard[cx.b], flag);
  __syncthreads();
  if (flag) return;
#ifdef WFST_PROFILE
  if (blockIdx.x == 0 && threadIdx.x == 0) printf("setup cycles: %lld\n", clock64() - pf_setup0);
#endif

  // warp -> scheduler partition is warp % 4
  constexpr int WR = W + kRS * W;
  if (warp < W) role_live<K, W>(a, sm, cx, warp);
  else if (warp < WR) role_rc<K, W>(a, sm, cx, (warp - W) % W, (warp - W) / W);
  else if (warp < WR + 2) role_reduce<K, W>(a, sm, cx, warp - WR);
  else role_producer<W>(a, sm, cx, warp - WR - 2);
}

// ---- host side ----------------------------------------------------------------------
// (slots per lane, warps per chain) by target length: short per-frame work per warp first
struct Cfg {
  int K, W;
  bool automatic;   // false: only when forced (WFST_CHAIN_CFG)
};
// the chain warps are skewed by a whole step, so few warps with more slots win once the chain
// needs more than two warps
static const Cfg kCfgs[] = {{4, 1, true}, {4, 2, true}, {4, 3, false}, {4, 4, false}, {6, 3, true}, {6, 4, true}, {6, 2, true}};
constexpr int kNumCfgs = (int)(sizeof(kCfgs) / sizeof(kCfgs[0]));

// test / tuning hook: WFST_CHAIN_CFG="K,W" forces a configuration for targets it can hold
static int g_force_k = -1, g_force_w = 0;
static void read_forced() {
  if (g_force_k >= 0) return;
  g_force_k = 0;
  const char* s = getenv("WFST_CHAIN_CFG");
  if (s) {
    int k = 0, w = 0;
    if (sscanf(s, "%d,%d", &k, &w) == 2) { g_force_k = k; g_force_w = w; }
  }
}

static int pick_cfg(int max_target_len) {
  const int S = 2 * max_target_len + 1;
  read_forced();
  int first = -1;
  for (int i = 0; i < kNumCfgs; ++i) {
    if (32 * kCfgs[i].K * kCfgs[i].W - 1 < S) continue;
    if (kCfgs[i].K == g_force_k && kCfgs[i].W == g_force_w) return i;
    if (!kCfgs[i].automatic) continue;
    if (first < 0 || 32 * kCfgs[i].K * kCfgs[i].W < 32 * kCfgs[first].K * kCfgs[first].W) first = i;   // smallest chain that holds the target
  }
  return first;
}

template <int K, int W>
static bool pick_bufs(int C, int& NAB, int& NB, size_t& bytes) {
  // step buffers for the live -> recompute -> reduce pipeline and p tiles for live + recompute;
  // two blocks per SM when that fits, else what fits in one
  const int nab_want = min(2 * W + 1, kMaxAB), nb_want = min(2 * W + 3, kMaxNB);
  const size_t two = (size_t)(113 * 1024), one = (size_t)(227 * 1024);
  for (int pass = 0; pass < 2; ++pass) {
    const size_t lim = pass == 0 ? two : one;
    const int nab_min = pass == 0 ? max(nab_want - 1, 3) : 3, nb_min = pass == 0 ? max(nb_want - 2, 4) : 4;
    for (int nab = nab_want; nab >= nab_min; --nab) {
      for (int nb = nb_want; nb >= nb_min; --nb) {
        const size_t b = make_layout<K, W>(C, nab, nb).total * sizeof(float);
        if (b <= lim) { NAB = nab; NB = nb; bytes = b; return true; }
      }
    }
  }
  return false;
}

template <int K, int W>
static int launch_kw(const Args& a, size_t smem, cudaStream_t st) {
  auto kern = ctc_chain_kernel<K, W>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  kern<<<a.B, Geo<K, W>::NT, smem, st>>>(a);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

#define WFST_CHAIN_DISPATCH(idx, EXPR)            \
  switch (idx) {                                  \
    case 0: { constexpr int K = 4, W = 1; EXPR; } break; \
    case 1: { constexpr int K = 4, W = 2; EXPR; } break; \
    case 2: { constexpr int K = 4, W = 3; EXPR; } break; \
    case 3: { constexpr int K = 4, W = 4; EXPR; } break; \
    case 4: { constexpr int K = 6, W = 3; EXPR; } break; \
    case 5: { constexpr int K = 6, W = 4; EXPR; } break; \
    case 6: { constexpr int K = 6, W = 2; EXPR; } break; \
    default: break;                               \
  }

static bool pick_bufs_cfg(int idx, int C, int& NAB, int& NB, size_t& bytes) {
  bool ok = false;
  WFST_CHAIN_DISPATCH(idx, ok = (pick_bufs<K, W>(C, NAB, NB, bytes)));
  return ok;
}

}  // namespace chaink

// test / tuning hook: prefer configuration (K, W) for targets it can hold; (0, 0) = automatic
int ctc_chain_force_config(int K, int W) {
  chaink::read_forced();
  chaink::g_force_k = K;
  chaink::g_force_w = W;
  return 0;
}

bool ctc_chain_eligible(int T, int C, int max_target_len) {
  if (T < 1 || C + 1 > 128) return false;
  const int idx = chaink::pick_cfg(max_target_len);
  if (idx < 0) return false;
  int nab, nb;
  size_t bytes;
  return chaink::pick_bufs_cfg(idx, C, nab, nb, bytes);
}

static int chain_nsd(int T) {
  const int a = T / 16, R = T - 16 * a;
  return a + (R > 0 ? 1 : 0);
}

static size_t chain_ckpt_bytes(int B, int T, int idx) {
  const chaink::Cfg c = chaink::kCfgs[idx];
  return align_up((size_t)B * chain_nsd(T) * 32 * c.W * (2 * c.K + 4) * sizeof(float), 256);
}

size_t ctc_chain_workspace_bytes(int B, int T, int max_target_len) {
  const int idx = chaink::pick_cfg(max_target_len);
  if (idx < 0) return 0;
  return chain_ckpt_bytes(B, T, idx) + align_up((size_t)B * sizeof(int), 256);
}

int launch_ctc_chain(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                     int blank, int max_target_len, const float* grad_scale, float* z_out,
                     float* gradE, void* workspace, int** hazard_out, cudaStream_t st) {
  using namespace chaink;
  const int idx = pick_cfg(max_target_len);
  Args a{};
  a.E = E; a.targets = targets; a.offsets = offsets; a.B = B; a.T = T; a.C = C; a.blank = blank;
  a.grad_scale = grad_scale; a.z_out = z_out; a.gradE = gradE;
  a.nfull = T / 16;
  const int R = T - 16 * a.nfull;
  a.r0 = (R + 1) / 2;
  a.r1 = R / 2;
  a.nsd = a.nfull + (R > 0 ? 1 : 0);
  a.Th = kSeg * a.nfull + a.r0;
  size_t smem = 0;
  if (idx < 0 || !pick_bufs_cfg(idx, C, a.NAB, a.NB, smem)) {
    set_error("no chain CTC configuration for C=%d L=%d", C, max_target_len);
    return WFST_ERR_UNSUPPORTED;
  }
  a.ckpt = (float*)workspace;
  a.hazard = (int*)((char*)workspace + chain_ckpt_bytes(B, T, idx));
  *hazard_out = a.hazard;
  WFST_CUDA_CHECK(cudaMemsetAsync(a.hazard, 0, (size_t)B * sizeof(int), st));
  int rc = WFST_ERR_UNSUPPORTED;
  WFST_CHAIN_DISPATCH(idx, rc = (launch_kw<K, W>(a, smem, st)));
  if (rc == WFST_ERR_UNSUPPORTED) set_error("no chain CTC instantiation for configuration %d", idx);
  return rc;
}

}  // namespace wfst
