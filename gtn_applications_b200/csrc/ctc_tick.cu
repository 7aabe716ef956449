// Chain-split scaled CTC forward+backward on a TICK schedule.
// Replaces, for CTC, the reference's per-utterance create_ctc_graph -> intersect ->
// forward_score -> backward (criterions/ctc.py:15-29,40-51,78-81).
//
// Same algorithm, numerics, certificate and warp roles as ctc_solo.cu (read its header and
// ctc_chain.cu's first): one block per utterance, the chain of S = 2L+1 states padded to
// Sp = 32*K*W slots and split over W warps per time direction, meet in the middle, phase 2 =
// live sweep + recompute from checkpoints + per-label reduction.  What differs is how the
// roles hand their buffers to each other.  ctc_solo.cu guards every resource (p tile, chain
// ring entry, step buffer, products) with its own pair of mbarriers; measured there
// (DESIGN.md §7.1): of 1270 cycles per phase-1 step only 380 are the 8 frames, and a run with
// every body skipped still takes 0.119 of 0.228 ms -- the step protocol is the kernel.  Here
// ALL roles of a component advance in lock step, one 8-frame step per TICK, and a single
// named barrier per tick and component replaces every full/empty pair: role r works at tick m
// on step m - delay(r), and every buffer is a ring indexed by the step number whose depth covers
// the delays between its writer and its last reader:
//
//   phase 1 (tick n)   P[c]: tile n+1      live[c][w]: step n - w
//   phase 2 (tick m)   P[c]: tile m+1      live[c][w]: step m - w      rc[c][w]: step m - W - w
//                      X[c]: step m - 2W
//
//   p tiles      ring of NB >= 2W+1 (written one tick ahead, last read by rc[W-1])
//   step buffers ring of NAB = 2W (abar rows, boundary state, lane exponents: live -> rc)
//   tiles        ring of W+1 (integer gradient tile, certificate terms: rc -> X)
//   chain rings  2 entries per warp boundary (written at tick t, read at tick t+1)
//
// The only mbarriers left are the producer's own TMA completion barriers.  Lanes are
// renormalised on EVEN TICKS (not even steps): the chain warps of a direction are skewed by one
// step, so events on even steps would put a 500-cycle event into every tick; lane 0 of a warp
// therefore refreshes its inbound scale factor from the ring entry at every step.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "launchers.h"

namespace wfst {
namespace tickk {

constexpr int kSeg = 8;               // frames per step / tile
constexpr int kEventEvery = 2;        // lanes are renormalised every kEventEvery ticks
constexpr int kUndef = -(1 << 19);    // "no exponent": lane holds only zeros
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNR = 3;                // raw (TMA) staging slots per producer warp, at most (Args::NR of them used)
constexpr int kRD = 2;                // entries per chain ring (step parity)
constexpr int kMaxAB = 8;             // step buffers per component, at most
constexpr int kMaxNB = 12;            // p tiles per component, at most
constexpr int kRingF = 12;            // floats per ring entry: 9 boundary values, the exponent, pad

#define WFST_HAZ(ptr, bits) atomicOr(ptr, bits)

struct Args {
  const float* E;
  const int* targets;
  const int* offsets;
  int B, T, C, blank;
  const float* grad_scale;
  float* z_out;     // [B] log Z
  float* gradE;     // [B, T, C] or null
  float* ckpt;      // [B][2][nsd][32 W][CKF]
  int* hazard;      // [B]
  int nsd;          // steps per direction and phase
  int nfull;        // full (8-frame) steps per direction
  int r0, r1;       // frames of the partial step of direction 0 / 1 (next to the meeting point)
  int Th;           // first frame of direction 1's half
  int NAB, NB, NR, NO;
};

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
__device__ __forceinline__ void sts(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void stsi(uint32_t a, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
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


// ---- mbarriers: the producers' TMA completion barriers only ------------------------
constexpr int kBarTma = 0;            // [2][kNR] raw tiles landed
constexpr int kNumBars = 2 * kNR;

__device__ __forceinline__ void bar_init(uint32_t bars, int idx, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8u * idx), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bars, int idx, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8u * idx), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bars, int idx, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WFSTT_BW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WFSTT_BD_%=;\n"
      "bra WFSTT_BW_%=;\n"
      "WFSTT_BD_%=:\n"
      "}\n" ::"r"(bars + 8u * idx), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// ---- ticks: one named barrier per component and phase ------------------------------
// phase 1: live[c][0..W) + P[c]; phase 2: live[c], rc[c], X[c], P[c]
// WFST_PROFILE: per-role cycles spent working / idle per tick (block 0 prints).  Clocks are
// read BEFORE the barrier only (a clock read after BAR.SYNC.DEFER_BLOCKING does not wait for
// the barrier): every warp posts its arrival time, the release time of a tick is the latest
// arrival among the component's warps.
struct Prof {
#ifdef WFST_PROFILE
  long long work, idle, wmax, prev_rel, mine;
  volatile long long* tab;   // [2][16] in shared memory
  unsigned mask;
  int warp, par;
  __device__ __forceinline__ Prof() : work(0), idle(0), wmax(0), tab(nullptr), mask(0), warp(0), par(0) { prev_rel = clock64(); mine = prev_rel; }
  __device__ __forceinline__ void setup(float* smem_prof, unsigned m) { tab = reinterpret_cast<volatile long long*>(smem_prof); mask = m; warp = threadIdx.x >> 5; prev_rel = clock64(); }
  long long acc[6] = {0, 0, 0, 0, 0, 0}, tm = 0;
  __device__ __forceinline__ void mark0() { tm = clock64(); }
  __device__ __forceinline__ void mark(int i) { const long long t = clock64(); acc[i] += t - tm; tm = t; }
  __device__ __forceinline__ void before() { mine = clock64(); tab[par * 16 + warp] = mine; }
  __device__ __forceinline__ void after() {
    long long rel = 0;
    for (int i = 0; i < 16; ++i) if ((mask >> i) & 1u) { const long long t = tab[par * 16 + i]; rel = t > rel ? t : rel; }
    const long long d = mine - prev_rel;
    work += d; if (d > wmax) wmax = d;
    idle += rel - mine;
    prev_rel = rel; par ^= 1;
  }
  __device__ __forceinline__ void report(const char* role, int c, int w, int phase, int lane) {
    if (blockIdx.x == 0 && lane == 0) printf("%s[%d][%d] phase %d: work %lld idle %lld longest %lld marks %lld %lld %lld %lld %lld %lld\n", role, c, w, phase, work, idle, wmax, acc[0], acc[1], acc[2], acc[3], acc[4], acc[5]);
    work = 0; idle = 0; wmax = 0; prev_rel = clock64();
    for (int i = 0; i < 6; ++i) acc[i] = 0;
  }
#else
  __device__ __forceinline__ void setup(float*, unsigned) {}
  __device__ __forceinline__ void mark0() {}
  __device__ __forceinline__ void mark(int) {}
  __device__ __forceinline__ void before() {}
  __device__ __forceinline__ void after() {}
  __device__ __forceinline__ void report(const char*, int, int, int, int) {}
#endif
};
template <int W>
__device__ __forceinline__ unsigned tick_mask(int c, int phase) {
  unsigned m = (((1u << W) - 1u) << (c * W)) | (1u << (4 * W + 2 + c));
  if (phase == 2) m |= (((1u << W) - 1u) << (2 * W + c * W)) | (1u << (4 * W + c));
  return m;
}
template <int W>
__device__ __forceinline__ void tick1(int c, Prof& pf) { pf.before(); named_sync(8 + c, 32 * (W + 1)); pf.after(); }
template <int W>
__device__ __forceinline__ void tick2(int c, Prof& pf) { pf.before(); named_sync(10 + c, 32 * (2 * W + 2)); pf.after(); }

// ---- geometry ------------------------------------------------------------------------
template <int K, int W>
struct Geo {
  static constexpr int NL = 32 * W;            // lanes of the chain
  static constexpr int Sp = K * NL;            // slots
  static constexpr int HL = K / 2;             // label slots per lane
  static constexpr int SA = (HL & 1) ? HL : HL + 1;     // abar row: words per lane (odd: conflict-free)
  static constexpr int PADA = 4;               // zero words in front of an abar row
  static constexpr int ROWW = PADA + SA * NL + 4;       // words per abar row
  static constexpr uint32_t ROWB = 4u * ROWW;
  static constexpr uint32_t EXTB = 4u * (SA - HL + 1);  // distance from a lane block back to the previous block's last label
  static constexpr int BNDW = 4 + K * NL + 4;  // words of a boundary row (touched once per step)
  static constexpr uint32_t BNDB = 4u * BNDW;
  static constexpr int CKF = (K + 2 + 3) & ~3; // checkpoint floats per lane: K values, the exponent, pad
  static constexpr int NWARPS = 4 * W + 4;     // live[2][W], rc[2][W], X[2], P[2]
  static constexpr int NT = 32 * NWARPS;
};

// shared memory layout (in floats); every per-component region is [2][...]
struct Layout {
  size_t raw, out, abuf, bnd, lexp, cert, gacc, ptile, ringL, ringR, bars, zx, ytab, prof, total;
  size_t abuf_c, bnd_c, lexp_c, cert_c, gacc_c, ptile_c, ring_c;   // size of one component's part
  size_t zero_end;
};
template <int K, int W>
__host__ __device__ inline Layout make_layout(int C, int NAB, int NB, int NR, int NO) {
  const int NG = W + 1;   // tiles / certificate terms: written by rc, read by X
  using G = Geo<K, W>;
  Layout L;
  const size_t rawsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
  const size_t CP = (size_t)C + 1;
  size_t p = 0;
  L.raw = p;    p += 2 * (size_t)NR * rawsz + 32;                  // [c][slot][8*C] (+ slack)
  L.out = p;    p += 2 * (size_t)NO * rawsz;                        // [c][ob][8*C]
  L.abuf_c = (size_t)NAB * kSeg * G::ROWW;
  L.abuf = p;   p += 2 * L.abuf_c;                                  // [c][buf][row][word]
  L.bnd_c = (size_t)NAB * G::BNDW;
  L.bnd = p;    p += 2 * L.bnd_c;                                   // [c][buf][word]: live state at the step boundary
  L.lexp_c = (size_t)NAB * G::NL;
  L.lexp = p;   p += 2 * L.lexp_c;                                  // [c][buf][gl] (int)
  L.cert_c = (size_t)NG * G::NL;
  L.cert = p;   p += 2 * L.cert_c;                                  // [c][buf][w]
  L.gacc_c = (size_t)NG * kSeg * CP;
  L.gacc = p;   p += 2 * L.gacc_c;                                  // [c][buf][row][class] (int, 2^-29 units)
  L.ptile_c = (size_t)NB * CP * 9 + 8;
  L.ptile = p;  p += 2 * L.ptile_c;                                 // [c][buf][col][9]
  p = (p + 3) & ~(size_t)3;
  L.ring_c = (size_t)(W + 1) * kRD * kRingF;
  L.ringL = p;  p += 2 * L.ring_c;                                  // [c][w][slot]{9 boundary values, e, pad}
  L.ringR = p;  p += 2 * L.ring_c;
  p = (p + 3) & ~(size_t)3;
  L.zero_end = p;
  L.bars = p;   p += 2 * kNumBars;
  p = (p + 3) & ~(size_t)3;
  L.zx = p;     p += 32;   // Zm, eZ, ok, -, msum(double), zpart[kMaxW]{contrib, Emax}, endacc[2]
  L.ytab = p;   p += (size_t)G::Sp / 2 + 4;                         // targets of the utterance
  p = (p + 3) & ~(size_t)3;
  L.prof = p;
#ifdef WFST_PROFILE
  p += 2 * 2 * 16 * 2;   // [phase-1 / phase-2 ticks][parity][warp] clocks
#endif
  L.total = p + 4;
  return L;
}

struct Smem {
  uint32_t raw, out, abuf, bnd, lexp, cert, gacc, ptile, ringL, ringR, bars, zx;
  uint32_t abuf_c, bnd_c, lexp_c, cert_c, gacc_c, ptile_c, ring_c;   // bytes per component
  int* ytab;
  float* out_gen;
  float* prof_gen;
};

struct Ctx {
  int lane, T, C, CP, L, b;
  int nsd, nfull, r0, r1, Th, NAB, NB, NR, NO;
  bool want_grad;
  uint32_t rawsz;
};

// phase-1 step k of direction d covers `rows` frames starting at `lo`
__device__ __forceinline__ int seg_rows(const Ctx& cx, int d, int k) { return k < cx.nfull ? kSeg : (d == 0 ? cx.r0 : cx.r1); }
__device__ __forceinline__ int seg_lo(const Ctx& cx, int d, int k) {
  if (d == 0) return kSeg * k;
  return k < cx.nfull ? cx.T - kSeg * (k + 1) : cx.Th;
}
// what component c works on at its tile kt (phase 1: kt < nsd, own half; phase 2: the other
// direction's steps, last one first)
__device__ __forceinline__ void comp_seg(const Ctx& cx, int c, int kt, int& lo_, int& rows) {
  const int d = kt < cx.nsd ? c : 1 - c;
  const int k = kt < cx.nsd ? kt : 2 * cx.nsd - 1 - kt;
  lo_ = seg_lo(cx, d, k);
  rows = seg_rows(cx, d, k);
}

// ---- per-lane topology --------------------------------------------------------------
template <int K>
struct Topo {
  uint32_t labofs[K / 2];  // byte offset in a p tile of row 0 of the label of odd slot 2q+1
  uint32_t pbofs;          // ... of the blank
  float skipm[K / 2];      // 1 if the skip arc into odd slot 2q+1 exists
};

// orientation o holds state s = j (o = 0) or s = Sp-2-j (o = 1) at slot j
template <int K, int W>
__device__ __forceinline__ void build_topo(Topo<K>& tp, const Ctx& cx, const int* ytab, int gl, int o, int blank) {
  constexpr int Sp = Geo<K, W>::Sp;
#pragma unroll
  for (int q = 0; q < K / 2; ++q) {
    const int j = gl * K + 2 * q + 1;
    const int s = o == 0 ? j : Sp - 2 - j;
    int col = cx.C;   // padding slots read the zero column
    float sk = 0.f;
    if (s >= 1 && s < 2 * cx.L + 1) {
      const int n = (s - 1) >> 1;
      col = min(max(ytab[n], 0), cx.C - 1);
      const int n2 = o == 0 ? n - 1 : n + 1;   // two positions earlier IN THIS ORIENTATION
      if (n2 >= 0 && n2 < cx.L && ytab[n2] != ytab[n]) sk = 1.f;
    }
    tp.labofs[q] = 36u * (uint32_t)col;
    tp.skipm[q] = sk;
  }
  tp.pbofs = 36u * (uint32_t)blank;
}

template <int K>
struct PRow {
  float pl[K / 2];
  float pb;
};
template <int K>
struct TileAddr {
  uint32_t la[K / 2];
  uint32_t pba;
};
template <int K>
__device__ __forceinline__ TileAddr<K> tile_addr(const Topo<K>& tp, uint32_t pt) {
  TileAddr<K> t;
#pragma unroll
  for (int q = 0; q < K / 2; ++q) t.la[q] = pt + tp.labofs[q];
  t.pba = pt + tp.pbofs;
  return t;
}
// row `it` of the tile (4 * it is an immediate when `it` is)
template <int K>
__device__ __forceinline__ PRow<K> load_prow(const TileAddr<K>& t, int it) {
  PRow<K> p;
  const uint32_t o = 4u * (uint32_t)it;
#pragma unroll
  for (int q = 0; q < K / 2; ++q) p.pl[q] = lds(t.la[q] + o);
  p.pb = lds(t.pba + o);
  return p;
}

// One frame.  v: with-emission values of the previous frame (own scale); on return this
// frame's with-emission values and, if WANT_ABAR, abar = the pre-emission sums of the label
// slots.  in1: the left neighbour's last slot, already converted to this lane's scale.
template <int K, bool WANT_ABAR>
__device__ __forceinline__ void step(float (&v)[K], float (&abar)[K / 2], const Topo<K>& tp, const PRow<K>& p, float in1) {
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    const float a1 = (i >= 1) ? v[i - 1] : in1;
    float s = v[i] + a1;
    if (i & 1) {
      const float a2 = (i >= 2) ? v[i - 2] : in1;
      s = fmaf(tp.skipm[i >> 1], a2, s);
      if (WANT_ABAR) abar[i >> 1] = s;
      v[i] = s * p.pl[i >> 1];
    } else {
      v[i] = s * p.pb;
    }
  }
}

__device__ __forceinline__ float ring_ld(uint32_t a) { return lds(a); }
__device__ __forceinline__ void ring_st(uint32_t a, float v, int lane) {
  if (lane == 31) sts(a, v);
}
// the left neighbour's last slot: by shuffle; lane 0 takes the value the previous warp of the
// chain left in the ring (zero for the first warp: its ring is never written)
__device__ __forceinline__ float left_in(float last, float bv, int lane, float f) {
  float left = __shfl_up_sync(kFull, last, 1);
  if (lane == 0) left = bv;
  return left * f;
}

// Event: renormalise the lane (max mantissa in [1,2)) and make the lane exponents consistent
// from left to right (see ctc_chain.cu): with m_l = number of lanes 0..l that hold mass,
//   E_l = max( max_{l' <= l, mass} (eown_l' + D m_l'),  Ein ) - D m_l ,
// Ein = the exponent of the previous warp's last lane (undefined for the first warp).
template <int K>
__device__ __forceinline__ void event1(float (&v)[K], int& e, float& f, int lane, int Ein) {
  constexpr int kChain = (2 * kSeg * kEventEvery + K - 1) / K + 1;   // lanes a wave can cross in a window
  constexpr int D = 96 / kChain;
  float m = v[0];
#pragma unroll
  for (int i = 1; i < K; ++i) m = fmaxf(m, v[i]);
  int ex = min(max((int)((__float_as_uint(m) >> 23) & 0xffu) - 127, -126), 126);
  const bool has = m > 0.f;
  if (!has) ex = 0;
  const int eown = has ? (defined_exp(e) ? e : 0) + ex : kUndef;
  const int dm = D * __popc(__ballot_sync(kFull, has) & (0xffffffffu >> (31 - lane)));
  int val = has ? eown + dm : 2 * kUndef;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) val = max(val, __shfl_up_sync(kFull, val, o));   // lanes < o get their own value back
  const int cin = max(val, defined_exp(Ein) ? Ein : 2 * kUndef) - dm;
  const int E = defined_exp(cin) ? cin : kUndef;
  const int t = -ex + ((defined_exp(E) && defined_exp(eown)) ? eown - E : 0);   // second term <= 0
  e = E;
  int el = __shfl_up_sync(kFull, E, 1);
  if (lane == 0) el = Ein;
  f = (!defined_exp(el) || !defined_exp(E)) ? 0.f : pow2c(el - E);   // el - E <= D
  {
    const float s1 = pow2i(max(t, -126));
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] *= s1;
  }
  if (__any_sync(kFull, t < -126)) {   // a lane pushed far below its own maximum: second factor
    const float s2 = pow2c(t - max(t, -126));
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] *= s2;
  }
}

// checkpoint: per lane CKF floats (K values, then the exponent), 128-bit accesses
template <int K, int CKF>
__device__ __forceinline__ void ckpt_store(float* base, const float (&v)[K], int e) {
  float t[CKF];
#pragma unroll
  for (int i = 0; i < CKF; ++i) t[i] = i < K ? v[i] : (i == K ? __int_as_float(e) : 0.f);
  float4* p = reinterpret_cast<float4*>(base);
#pragma unroll
  for (int i = 0; i < CKF / 4; ++i) p[i] = make_float4(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]);
}
template <int K, int CKF>
__device__ __forceinline__ void ckpt_load(const float* base, float (&v)[K], int& e) {
  float t[CKF];
  const float4* p = reinterpret_cast<const float4*>(base);
#pragma unroll
  for (int i = 0; i < CKF / 4; ++i) {
    const float4 q = p[i];
    t[4 * i] = q.x; t[4 * i + 1] = q.y; t[4 * i + 2] = q.z; t[4 * i + 3] = q.w;
  }
#pragma unroll
  for (int i = 0; i < K; ++i) v[i] = t[i];
  e = __float_as_int(t[K]);
}

// p tile buffer `buf` of component c
__device__ __forceinline__ uint32_t ptile_addr(const Smem& sm, const Ctx& cx, int c, int buf) {
  return sm.ptile + (uint32_t)c * sm.ptile_c + 4u * (uint32_t)(buf * cx.CP * 9);
}

// ---------------------------------------------------------------------------
// P[c]: producer of component c's p tiles, one tick ahead of live[c][0]: phase-1 tiles first;
// the phase-2 tiles only once Z is known to be usable.  Lane = class (classes beyond 32 in
// further groups): a row maximum is one warp reduction, the transposed tile is written with
// conflict-free stores.  Raw tiles arrive by TMA, up to kNR in flight (the only mbarriers).
// ---------------------------------------------------------------------------
struct ProducerState {
  int fetched, converted;
  uint32_t tma_phase, tma_used;
  int pbuf;            // p-tile buffer of the next tile
  double msum;
};

// raw tiles of [.., kt_end) in flight, at most kNR beyond the last converted one
__device__ __forceinline__ void producer_fetch(const Args& a, const Smem& sm, const Ctx& cx, ProducerState& ps,
                                               const int c, const int kt_end) {
  const int lane = cx.lane, T = cx.T, C = cx.C;
  const uint32_t rawsz = cx.rawsz;
  const uint32_t raw0 = sm.raw + 4u * (uint32_t)(c * cx.NR) * rawsz;
  const int tbar = kBarTma + c * cx.NR;
  const float* Eb = a.E + (size_t)cx.b * T * C;
  while (ps.fetched < kt_end && ps.fetched < ps.converted + cx.NR) {
    int lo_, rows;
    comp_seg(cx, c, ps.fetched, lo_, rows);
    const int slot = ps.fetched % cx.NR;
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
  }
}

// tile kt (== ps.converted) from its raw slot into its p-tile buffer
__device__ __forceinline__ void producer_convert(const Smem& sm, const Ctx& cx, ProducerState& ps, const int c,
                                                 const int kt, const bool phase1, Prof& pf) {
  pf.mark0();
  const int lane = cx.lane, C = cx.C;
  const uint32_t rawsz = cx.rawsz;
  const uint32_t raw0 = sm.raw + 4u * (uint32_t)(c * cx.NR) * rawsz;
  const int tbar = kBarTma + c * cx.NR;
  const int groups = (C + 31) >> 5;
  const int slot = ps.converted % cx.NR;
  int lo_, rows;
  comp_seg(cx, c, kt, lo_, rows);
  const int buf = ps.pbuf;
  if (++ps.pbuf == cx.NB) ps.pbuf = 0;
  if ((ps.tma_used >> slot) & 1u) {
    bar_wait(sm.bars, tbar + slot, (ps.tma_phase >> slot) & 1u);
    ps.tma_phase ^= 1u << slot;
  }
  pf.mark(0);
  const uint32_t er = raw0 + 4u * (uint32_t)slot * rawsz + 4u * (uint32_t)lane;
  const uint32_t pt = ptile_addr(sm, cx, c, buf) + 36u * (uint32_t)lane;
  // tile row = the step at which component c consumes the frame (c = 0 ascends, c = 1 descends).
  // A row that is entirely -inf keeps p = 0 (dead frame); +inf / NaN rows surface through the
  // certificate.
  float base[kSeg];
  if (groups == 1 && rows == kSeg) {
    // the common case (full tile, C <= 32) without row predicates
    const bool valid = lane < C;
    const uint32_t e0 = valid ? er : raw0 + 4u * (uint32_t)slot * rawsz;   // idle lanes re-read class 0: the maximum is unchanged
    float x[kSeg];
#pragma unroll
    for (int r = 0; r < kSeg; ++r) x[r] = lds(e0 + 4u * (uint32_t)(r * C));
#pragma unroll
    for (int r = 0; r < kSeg; ++r) {
      const float mx = warp_max(x[r]);
      base[r] = (mx == kNegInf) ? 0.f : mx;
    }
    if (valid) {
      const uint32_t p0 = pt + (c == 0 ? 0u : 4u * (kSeg - 1));
      const int32_t dp = c == 0 ? 4 : -4;
#pragma unroll
      for (int r = 0; r < kSeg; ++r)
        sts(p0 + (uint32_t)(r * dp), ex2_fast((x[r] - base[r]) * 1.4426950408889634f));
    }
  } else if (groups == 1) {
    const bool valid = lane < C;
    float x[kSeg];
#pragma unroll
    for (int r = 0; r < kSeg; ++r) x[r] = (valid && r < rows) ? lds(er + 4u * (uint32_t)(r * C)) : kNegInf;
#pragma unroll
    for (int r = 0; r < kSeg; ++r) {
      const float mx = warp_max(x[r]);
      base[r] = (mx == kNegInf) ? 0.f : mx;
    }
#pragma unroll
    for (int r = 0; r < kSeg; ++r) {
      if (valid && r < rows) {
        const int trow = c == 0 ? r : rows - 1 - r;
        sts(pt + 4u * (uint32_t)trow, ex2_fast(fmaf(x[r], 1.4426950408889634f, -base[r] * 1.4426950408889634f)));
      }
    }
  } else {
    float mx[kSeg];
#pragma unroll
    for (int r = 0; r < kSeg; ++r) mx[r] = kNegInf;
    for (int g = 0; g < groups; ++g) {
      const bool valid = 32 * g + lane < C;
#pragma unroll
      for (int r = 0; r < kSeg; ++r)
        if (valid && r < rows) mx[r] = fmaxf(mx[r], lds(er + 4u * (uint32_t)(r * C + 32 * g)));
    }
#pragma unroll
    for (int r = 0; r < kSeg; ++r) {
      const float m = warp_max(mx[r]);
      base[r] = (m == kNegInf) ? 0.f : m;
    }
    for (int g = 0; g < groups; ++g) {
      const bool valid = 32 * g + lane < C;
#pragma unroll
      for (int r = 0; r < kSeg; ++r) {
        if (valid && r < rows) {
          const int trow = c == 0 ? r : rows - 1 - r;
          const float x = lds(er + 4u * (uint32_t)(r * C + 32 * g));
          sts(pt + 36u * (uint32_t)(32 * g) + 4u * (uint32_t)trow, ex2_fast(fmaf(x, 1.4426950408889634f, -base[r] * 1.4426950408889634f)));
        }
      }
    }
  }
  if (phase1) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kSeg; ++r) s += r < rows ? base[r] : 0.f;
    ps.msum += (double)s;
  }
  pf.mark(1);
  ++ps.converted;
}

template <int W>
__device__ __forceinline__ void role_producer(const Args& a, const Smem& sm, const Ctx& cx, const int c) {
  const int lane = cx.lane, nsd = cx.nsd;
  Prof pf;
  pf.setup(sm.prof_gen, tick_mask<W>(c, 1));
  ProducerState ps;
  ps.fetched = 0; ps.converted = 0; ps.tma_phase = 0u; ps.tma_used = 0u;
  ps.pbuf = 0;
  ps.msum = 0.0;
  producer_fetch(a, sm, cx, ps, c, nsd);
  producer_convert(sm, cx, ps, c, 0, true, pf);
  producer_fetch(a, sm, cx, ps, c, nsd);
  tick1<W>(c, pf);
  for (int n = 0; n < nsd + W; ++n) {
    if (n + 1 < nsd) {
      producer_convert(sm, cx, ps, c, n + 1, true, pf);
      producer_fetch(a, sm, cx, ps, c, nsd);
      pf.mark(2);
    }
    tick1<W>(c, pf);
  }
  pf.report("P", c, 0, 1, lane);
  // loss: log Z = log(Zm) + eZ ln2 + sum_t max_t; the two producers each hold the row maxima
  // of their phase-1 half (every lane holds the same sum)
  double* msh = reinterpret_cast<double*>(__cvta_shared_to_generic(sm.zx + 16u));
  if (c == 0 && lane == 0) msh[0] = ps.msum;
  __syncthreads();   // meeting A
  __syncthreads();   // meeting B
  __syncthreads();   // meeting C: Z published
  const float Zm = lds(sm.zx);
  const int eZ = ldsi(sm.zx + 4u);
  const bool ok = lds(sm.zx + 8u) != 0.f;
  if (c == 1 && lane == 0)
    a.z_out[cx.b] = ok ? (float)(log((double)Zm) + (double)eZ * 0.6931471805599453 + (ps.msum + msh[0])) : kNegInf;
  if (!cx.want_grad || !ok) return;
  pf.setup(sm.prof_gen + 64, tick_mask<W>(c, 2));
  producer_fetch(a, sm, cx, ps, c, 2 * nsd);
  producer_convert(sm, cx, ps, c, nsd, false, pf);
  producer_fetch(a, sm, cx, ps, c, 2 * nsd);
  tick2<W>(c, pf);
  for (int m = 0; m < nsd + 2 * W; ++m) {
    if (m + 1 < nsd) {
      producer_convert(sm, cx, ps, c, nsd + m + 1, false, pf);
      producer_fetch(a, sm, cx, ps, c, 2 * nsd);
      pf.mark(2);
    }
    tick2<W>(c, pf);
  }
  pf.report("P", c, 0, 2, lane);
}

// ---------------------------------------------------------------------------
// live[c][w]: warp w of component c's chain
// ---------------------------------------------------------------------------
template <int K, int W>
__device__ __forceinline__ void role_live(const Args& a, const Smem& sm, const Ctx& cx, const int c, const int w) {
  using G = Geo<K, W>;
  constexpr int Sp = G::Sp, NL = G::NL, HL = G::HL;
  const int lane = cx.lane, gl = 32 * w + lane, nsd = cx.nsd, NAB = cx.NAB;
  const int S = 2 * cx.L + 1;
  float* ck = a.ckpt + (((size_t)cx.b * 2 + c) * nsd * NL + gl) * G::CKF;
  Topo<K> tp;
  build_topo<K, W>(tp, cx, sm.ytab, gl, c, a.blank);

  float v[K], abar[HL];
  int e = kUndef;
  float f = 0.f;
  {
    // virtual pre-frame state: all mass on the start slot of this orientation
    const int jstart = c == 0 ? 0 : Sp - 1 - S;
    const bool mine = (jstart / K == gl);
    const int jm = jstart % K;
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = (mine && jm == i) ? 1.f : 0.f;
    if (mine) e = 0;
  }
  const uint32_t ring_in0 = sm.ringL + (uint32_t)c * sm.ring_c + 4u * (uint32_t)(w * kRD * kRingF);
  const uint32_t ring_out0 = ring_in0 + 4u * (uint32_t)(kRD * kRingF);
  uint32_t rin = ring_in0, rout = ring_out0;
  const bool has_partial = nsd > cx.nfull;

  // start of global step g: the left warp's ring entry of this step was written one tick ago;
  // renormalise if due (else lane 0 refreshes its inbound factor: the left warp may have
  // renormalised since), open my own entry
  auto step_begin = [&](int g, bool ev) {
    const int slot = g & 1;
    rin = ring_in0 + 4u * (uint32_t)(slot * kRingF);
    rout = ring_out0 + 4u * (uint32_t)(slot * kRingF);
    const int Ein = w > 0 ? ldsi(rin + 36u) : kUndef;
    if (ev) {
      event1<K>(v, e, f, lane, Ein);
    } else if (W > 1 && lane == 0) {
      f = (defined_exp(Ein) && defined_exp(e)) ? pow2c(Ein - e) : 0.f;
    }
    if (w < W - 1 && lane == 31) {
      sts(rout, v[K - 1]);
      stsi(rout + 36u, e);
    }
  };
  // frames of a partial step
  auto slow_frames = [&](const TileAddr<K>& ta, int rows, uint32_t ar, bool want_abar) {
#pragma unroll 1
    for (int it = 0; it < rows; ++it) {
      const PRow<K> cur = load_prow<K>(ta, it);
      const float in1 = left_in(v[K - 1], lds(rin + 4u * (uint32_t)it), lane, f);
      step<K, true>(v, abar, tp, cur, in1);
      if (want_abar) {
#pragma unroll
        for (int q = 0; q < HL; ++q) sts(ar + (uint32_t)it * G::ROWB + 4u * q, abar[q]);
      }
      if (lane == 31) sts(rout + 4u * (uint32_t)(it + 1), v[K - 1]);
    }
  };

  const uint32_t mybnd = 4u * (uint32_t)(4 + gl * K);       // my slots in a boundary row
  const uint32_t myabar = 4u * (uint32_t)(G::PADA + gl * G::SA);

  // ------------------------------------------------------------------ phase 1
  Prof pf;
  pf.setup(sm.prof_gen, tick_mask<W>(c, 1));
  int pbuf = 0;
  tick1<W>(c, pf);   // tile 0 is there
  for (int n = 0; n < nsd + W; ++n) {
    const int g = n - w;
    if (g >= 0 && g < nsd) {
      step_begin(g, (n % kEventEvery) == 0 || g == 0);
      ckpt_store<K, G::CKF>(ck + (size_t)g * NL * G::CKF, v, e);
      const bool partial = has_partial && g == cx.nfull;
      const TileAddr<K> ta = tile_addr<K>(tp, ptile_addr(sm, cx, c, pbuf));
      if (++pbuf == cx.NB) pbuf = 0;
      if (!partial) {
        PRow<K> nx = load_prow<K>(ta, 0);
        float bvn = ring_ld(rin);
#pragma unroll
        for (int it = 0; it < kSeg; ++it) {
          const PRow<K> cur = nx;
          const float bv = bvn;
          if (it + 1 < kSeg) { nx = load_prow<K>(ta, it + 1); bvn = ring_ld(rin + 4u * (it + 1)); }
          const float in1 = left_in(v[K - 1], bv, lane, f);
          step<K, false>(v, abar, tp, cur, in1);
          ring_st(rout + 4u * (it + 1), v[K - 1], lane);
        }
      } else {
        slow_frames(ta, c == 0 ? cx.r0 : cx.r1, 0u, false);
      }
    } else if (g == nsd) {
      // meeting: every live warp renormalises (consistent exponents for the successor sums
      // below); the alpha warps publish their state in the layout of a boundary row (buffer 0)
      step_begin(nsd, true);
      if (c == 0) {
#pragma unroll
        for (int i = 0; i < K; ++i) sts(sm.bnd + mybnd + 4u * i, v[i]);
        stsi(sm.lexp + 4u * (uint32_t)gl, e);
      }
    }
    tick1<W>(c, pf);
  }

  pf.report("live", c, w, 1, lane);
  // ------------------------------------------------------------------ meeting: Z
  // the beta warps form   Z = sum over their slots of (successor sum of beta~)(slot) * alpha(partner slot).
  __syncthreads();   // A: both chains have arrived, the alpha state is published
  if (c == 1) {
    float bb[K];
    {
      const float in1 = left_in(v[K - 1], lds(rin), lane, f);
#pragma unroll
      for (int i = K - 1; i >= 0; --i) {
        const float a1 = (i >= 1) ? v[i - 1] : in1;
        float s = v[i] + a1;
        if (i & 1) {
          const float a2 = (i >= 2) ? v[i - 2] : in1;
          s = fmaf(tp.skipm[i >> 1], a2, s);
        }
        bb[i] = s;
      }
    }
    // partner of my slot i is slot K-2-i of lane NL-1-gl; of my slot K-1, the last slot of lane NL-2-gl
    const int pl = NL - 1 - gl;
    const uint32_t pblock = sm.bnd + 4u * (uint32_t)(4 + pl * K);
    float pm = 0.f;
#pragma unroll
    for (int i = 0; i <= K - 2; ++i) pm = fmaf(bb[i], lds(pblock + 4u * (K - 2 - i)), pm);
    const float px = bb[K - 1] * lds(pblock - 4u);
    const int ea = ldsi(sm.lexp + 4u * (uint32_t)pl);
    const int eb = pl > 0 ? ldsi(sm.lexp + 4u * (uint32_t)(pl - 1)) : kUndef;
    int Em = kUndef, Ex = kUndef;
    if (pm > 0.f && defined_exp(e) && defined_exp(ea)) Em = e + ea;
    if (px > 0.f && defined_exp(e) && defined_exp(eb)) Ex = e + eb;
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
  __syncthreads();   // B
  if (c == 1 && w == 0 && lane == 0) {
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
      ex = min(max(ex, -126), 125);
      Zm = tot * pow2i(-ex) * 0.5f;   // in [0.5, 1): the fixed-point posteriors stay below 1
      ex += 1;
    }
    sts(sm.zx, Zm);
    stsi(sm.zx + 4u, ok ? Emax + ex : 0);
    sts(sm.zx + 8u, ok ? 1.f : 0.f);
    // reason 2: infeasible or out of range -- the fallback kernel decides
    if (!ok) WFST_HAZ(&a.hazard[cx.b], 2);
  }
  __syncthreads();   // C: Z published
  const bool okz = lds(sm.zx + 8u) != 0.f;
  if (!cx.want_grad || !okz) return;

  // ------------------------------------------------------------------ phase 2
  const uint32_t abuf = sm.abuf + (uint32_t)c * sm.abuf_c, bndb = sm.bnd + (uint32_t)c * sm.bnd_c,
                 lexpb = sm.lexp + (uint32_t)c * sm.lexp_c;
  pbuf = nsd % cx.NB;
  int buf = 0;
  pf.setup(sm.prof_gen + 64, tick_mask<W>(c, 2));
  tick2<W>(c, pf);   // the first phase-2 tile is there
  for (int m = 0; m < nsd + 2 * W; ++m) {
    const int k2 = m - w;
    if (k2 >= 0 && k2 < nsd) {
      step_begin(nsd + 1 + k2, (m % kEventEvery) == 0);
      {
        stsi(lexpb + 4u * (uint32_t)(buf * NL + gl), e);
        // state at the step boundary: the recompute warps check Z against it (certificate)
        const uint32_t bb = bndb + (uint32_t)buf * G::BNDB + mybnd;
#pragma unroll
        for (int i = 0; i < K; ++i) sts(bb + 4u * i, v[i]);
      }
      // component c continues through the other direction's steps, last one (the partial one) first
      const bool partial = has_partial && k2 == 0;
      const TileAddr<K> ta = tile_addr<K>(tp, ptile_addr(sm, cx, c, pbuf));
      if (++pbuf == cx.NB) pbuf = 0;
      const uint32_t ar = abuf + (uint32_t)(buf * kSeg) * G::ROWB + myabar;
      if (!partial) {
        PRow<K> nx = load_prow<K>(ta, 0);
        float bvn = ring_ld(rin);
#pragma unroll
        for (int it = 0; it < kSeg; ++it) {
          const PRow<K> cur = nx;
          const float bv = bvn;
          if (it + 1 < kSeg) { nx = load_prow<K>(ta, it + 1); bvn = ring_ld(rin + 4u * (it + 1)); }
          const float in1 = left_in(v[K - 1], bv, lane, f);
          step<K, true>(v, abar, tp, cur, in1);
#pragma unroll
          for (int q = 0; q < HL; ++q) sts(ar + (uint32_t)it * G::ROWB + 4u * q, abar[q]);
          ring_st(rout + 4u * (it + 1), v[K - 1], lane);
        }
      } else {
        slow_frames(ta, c == 0 ? cx.r1 : cx.r0, ar, true);
      }
      if (++buf == NAB) buf = 0;
    }
    tick2<W>(c, pf);
  }
  pf.report("live", c, w, 2, lane);
  // certificate, last leg: the sweep must arrive with total mass Z on the two slots that end the
  // chain in this orientation (the recompute warps check every earlier step boundary)
  {
    const float Zm = lds(sm.zx);
    const int eZ = ldsi(sm.zx + 4u);
    const int jend = c == 0 ? S - 1 : Sp - 2;          // last state of the chain in this orientation
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int j = gl * K + i;
      if (j == jend || (j == jend - 1 && j >= 0)) part += v[i];
    }
    if (part != 0.f) part *= defined_exp(e) ? pow2c(e - eZ) : 0.f;
    const float tot = warp_sum(part);
    // the two end slots sit in at most two warps: a sum of two terms is order independent
    if (lane == 0 && tot != 0.f) atomicAdd(reinterpret_cast<float*>(__cvta_shared_to_generic(sm.zx + 64u + 4u * c)), tot);
    if (c == 0) named_sync(3, 32 * W); else named_sync(4, 32 * W);
    if (w == 0 && lane == 0) {
      const float t0 = lds(sm.zx + 64u + 4u * c);
      if (!(fabsf(t0 - Zm) <= 2e-5f * Zm)) WFST_HAZ(&a.hazard[cx.b], 8);
    }
  }
}

// ---------------------------------------------------------------------------
// rc[c][w]: runs direction 1-c over the steps of live[c]'s phase 2, from the checkpoints
// live[1-c] wrote in phase 1, against live[c]'s step order, and multiplies with the stored abar
// rows.  Before each step the lane is rescaled so that its effective exponent is
// eZ - e_live(partner lane): products need no further factor.
// ---------------------------------------------------------------------------
// Fixed point in 2^-23 units without a conversion: for 0 <= x < 1 the float 1 + x has the
// exponent of 1.0 and its mantissa field is round(x * 2^23); the integer difference to the bits
// of 1.0 stays continuous at x = 1.  Zm is kept in [0.5, 1), so every posterior is below 1.
constexpr float kFix = 8388608.f;
__device__ __forceinline__ void red_add(uint32_t a, float w, float av) {
  const int v = __float_as_int(fmaf(w, av, 1.f)) - 0x3f800000;
  asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// gt: this frame's row of the step's gradient tile; gofs[q]: byte offset of the class of my odd
// slot 2q+1 in a tile row (the dump column for padding slots)
template <int K>
__device__ __forceinline__ void rc_frame(float (&w)[K], const Topo<K>& tp, const PRow<K>& cur, float in1, float h, uint32_t arow,
                                         uint32_t extb, uint32_t gt, const uint32_t (&gofs)[K / 2]) {
  // partner of my odd slot i (<= K-3) is label slot (K-3-i)/2 of the partner block; of my slot
  // K-1, the last label slot of the block before it
  float av[K / 2], dummy[K / 2];
#pragma unroll
  for (int q = 0; q < K / 2 - 1; ++q) av[q] = lds(arow + 4u * q);
  const float ext = lds(arow - extb);
  step<K, false>(w, dummy, tp, cur, in1);
  // posteriors (times Zm) of my label states, summed by class with integer atomics: the sum does
  // not depend on the order
#pragma unroll
  for (int q = 0; q < K / 2 - 1; ++q) red_add(gt + gofs[(K - 3 - 2 * q) >> 1], w[K - 3 - 2 * q], av[q]);
  red_add(gt + gofs[(K - 1) >> 1], w[K - 1] * h, ext);
}

template <int K, int W>
__device__ __forceinline__ void role_rc(const Args& a, const Smem& sm, const Ctx& cx, const int c, const int w) {
  using G = Geo<K, W>;
  constexpr int NL = G::NL;
  const int lane = cx.lane, gl = 32 * w + lane, nsd = cx.nsd, NAB = cx.NAB;
  __syncthreads();   // meeting A
  __syncthreads();   // meeting B
  __syncthreads();   // meeting C: phase 1 (and every checkpoint) is complete, Z published
  const bool ok = lds(sm.zx + 8u) != 0.f;
  if (!cx.want_grad || !ok) return;
  const int eZ = ldsi(sm.zx + 4u);
  Topo<K> tp;
  build_topo<K, W>(tp, cx, sm.ytab, gl, 1 - c, a.blank);
  const float* ck = a.ckpt + (((size_t)cx.b * 2 + (1 - c)) * nsd * NL + gl) * G::CKF;
  const int pl = NL - 1 - gl;
  const uint32_t pabar = 4u * (uint32_t)(G::PADA + pl * G::SA);   // partner block in an abar row
  const uint32_t pbnd = 4u * (uint32_t)(4 + pl * K);              // partner block in a boundary row
  const uint32_t ring_in0 = sm.ringR + (uint32_t)c * sm.ring_c + 4u * (uint32_t)(w * kRD * kRingF);
  const uint32_t ring_out0 = ring_in0 + 4u * (uint32_t)(kRD * kRingF);
  const uint32_t abuf = sm.abuf + (uint32_t)c * sm.abuf_c, bndb = sm.bnd + (uint32_t)c * sm.bnd_c,
                 lexpb = sm.lexp + (uint32_t)c * sm.lexp_c, certb = sm.cert + (uint32_t)c * sm.cert_c;
  uint32_t gofs[K / 2];
#pragma unroll
  for (int q = 0; q < K / 2; ++q) gofs[q] = tp.labofs[q] / 9u;   // 36 col -> 4 col
  const uint32_t gaccb = sm.gacc + (uint32_t)c * sm.gacc_c;
  const uint32_t growb = 4u * (uint32_t)cx.CP;
  int bad = 0;   // reason 4: scale overflow when pairing live and recomputed values
  float wv[K], nv[K];
  int ew, ne;
  // the checkpoint of a step is fetched while the step before it runs
  ckpt_load<K, G::CKF>(ck + (size_t)(nsd - 1) * NL * G::CKF, nv, ne);
  const bool has_partial = nsd > cx.nfull;
  int pbuf = nsd % cx.NB, buf = 0, gbuf = 0;
  Prof pf;
  pf.setup(sm.prof_gen + 64, tick_mask<W>(c, 2));
  tick2<W>(c, pf);
  for (int m = 0; m < nsd + 2 * W; ++m) {
    const int k2 = m - W - w;
    if (k2 >= 0 && k2 < nsd) {
#pragma unroll
      for (int i = 0; i < K; ++i) wv[i] = nv[i];
      ew = ne;
      if (k2 + 1 < nsd) ckpt_load<K, G::CKF>(ck + (size_t)(nsd - 2 - k2) * NL * G::CKF, nv, ne);
      const int slot = k2 & 1;
      const uint32_t rin = ring_in0 + 4u * (uint32_t)(slot * kRingF);
      const uint32_t rout = ring_out0 + 4u * (uint32_t)(slot * kRingF);
      const bool partial = has_partial && k2 == 0;
      const int rows = partial ? (c == 0 ? cx.r1 : cx.r0) : kSeg;
      // scales
      float gsc = 0.f, hsc = 0.f, frs = 0.f;
      {
        const uint32_t le = lexpb + 4u * (uint32_t)(buf * NL);
        const int ep = ldsi(le + 4u * (uint32_t)pl);                          // partner lane
        const int ex = pl > 0 ? ldsi(le + 4u * (uint32_t)(pl - 1)) : kUndef;  // partner of slot K-1
        const int el = gl > 0 ? ldsi(le + 4u * (uint32_t)(pl + 1)) : kUndef;  // partner of my left neighbour
        if (defined_exp(ep)) {
          if (defined_exp(ew)) {
            const int dd = ew + ep - eZ;
            if (dd > 126) bad |= 4;
            else gsc = pow2c(dd);
          }
          if (defined_exp(ex)) hsc = pow2c(ex - ep);     // <= 2^D by the event invariant
          if (defined_exp(el)) frs = pow2c(ep - el);     // <= 2^D likewise
        }
      }
#pragma unroll
      for (int i = 0; i < K; ++i) wv[i] *= gsc;
      // chain ring: my left neighbour wrote its entry of this step one tick ago; open my own
      if (w < W - 1 && lane == 31) sts(rout, wv[K - 1]);
      const TileAddr<K> ta = tile_addr<K>(tp, ptile_addr(sm, cx, c, pbuf));
      if (++pbuf == cx.NB) pbuf = 0;
      const uint32_t ar = abuf + (uint32_t)(buf * kSeg) * G::ROWB + pabar;
      const uint32_t gt0 = gaccb + (uint32_t)(gbuf * kSeg) * growb;
      if (!partial) {
        // against the live step order
        PRow<K> nx = load_prow<K>(ta, kSeg - 1);
        float bvn = ring_ld(rin);
#pragma unroll
        for (int it = kSeg - 1; it >= 0; --it) {
          const PRow<K> cur = nx;
          const float bv = bvn;
          if (it > 0) { nx = load_prow<K>(ta, it - 1); bvn = ring_ld(rin + 4u * (kSeg - it)); }
          const float in1 = left_in(wv[K - 1], bv, lane, frs);
          rc_frame<K>(wv, tp, cur, in1, hsc, ar + (uint32_t)it * G::ROWB, G::EXTB, gt0 + (uint32_t)it * growb, gofs);
          ring_st(rout + 4u * (kSeg - it), wv[K - 1], lane);
        }
      } else {
#pragma unroll 1
        for (int it = rows - 1; it >= 0; --it) {
          const PRow<K> cur = load_prow<K>(ta, it);
          const float in1 = left_in(wv[K - 1], lds(rin + 4u * (uint32_t)(rows - 1 - it)), lane, frs);
          rc_frame<K>(wv, tp, cur, in1, hsc, ar + (uint32_t)it * G::ROWB, G::EXTB, gt0 + (uint32_t)it * growb, gofs);
          if (lane == 31) sts(rout + 4u * (uint32_t)(rows - it), wv[K - 1]);
        }
      }
      {
        // certificate: sum_s v_live(s) * (successor sum of w)(s) at the step boundary must be Z
        // (float32 range can only be exceeded by losing mass or producing inf / NaN)
        const float in1 = left_in(wv[K - 1], lds(rin + 4u * (uint32_t)rows), lane, frs);
        const uint32_t bb = bndb + (uint32_t)buf * G::BNDB + pbnd;
        float acc = 0.f;
#pragma unroll
        for (int i = K - 1; i >= 0; --i) {
          const float a1 = (i >= 1) ? wv[i - 1] : in1;
          float sx = wv[i] + a1;
          if (i & 1) {
            const float a2 = (i >= 2) ? wv[i - 2] : in1;
            sx = fmaf(tp.skipm[i >> 1], a2, sx);
          }
          if (i == K - 1) acc = fmaf(sx * hsc, lds(bb - 4u), acc);
          else acc = fmaf(sx, lds(bb + 4u * (K - 2 - i)), acc);
        }
        sts(certb + 4u * (uint32_t)(gbuf * NL + gl), acc);   // X sums the lanes' terms
      }
      if (++buf == NAB) buf = 0;
      if (++gbuf == W + 1) gbuf = 0;
    }
    tick2<W>(c, pf);
  }
  pf.report("rc", c, w, 2, lane);
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0) WFST_HAZ(&a.hazard[cx.b], bad);
}

// ---------------------------------------------------------------------------
// X[c]: turns a step's integer tile (label posteriors summed by class by the recompute warps)
// into the [8, C] gradient tile and stores it.  Lane = class (classes beyond 32 in further
// rounds).
// ---------------------------------------------------------------------------
template <int K, int W>
__device__ __forceinline__ void role_reduce(const Args& a, const Smem& sm, const Ctx& cx, const int c) {
  const int lane = cx.lane, nsd = cx.nsd, T = cx.T, C = cx.C, NAB = cx.NAB;
  const int rounds = (C + 31) >> 5;
  __syncthreads();   // meeting A
  __syncthreads();   // meeting B
  __syncthreads();   // meeting C: Z published
  const bool ok = lds(sm.zx + 8u) != 0.f;
  if (!cx.want_grad || !ok) return;
  const uint32_t rawsz = cx.rawsz;
  const float Zm = lds(sm.zx);
  const float kappa = -(a.grad_scale ? a.grad_scale[cx.b] : 1.f) / Zm;
  const bool has_partial = nsd > cx.nfull;
  const uint32_t gaccb = sm.gacc + (uint32_t)c * sm.gacc_c, certb = sm.cert + (uint32_t)c * sm.cert_c;
  const uint32_t growb = 4u * (uint32_t)cx.CP;
  int bad = 0;
  float* gE = a.gradE + (size_t)cx.b * T * C;
  int buf = 0;
  Prof pf;
  pf.setup(sm.prof_gen + 64, tick_mask<W>(c, 2));
  tick2<W>(c, pf);
  for (int m = 0; m < nsd + 2 * W; ++m) {
    const int k2 = m - 2 * W;
    if (k2 >= 0) {
      // component c works through the other direction's steps, the partial one first
      const int kk = nsd - 1 - k2;
      const int rows = (has_partial && k2 == 0) ? (c == 0 ? cx.r1 : cx.r0) : kSeg;
      const int lo_ = seg_lo(cx, 1 - c, kk);
      const int ob = cx.NO == 2 ? (k2 & 1) : 0;
      pf.mark0();
      {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < W; ++i) tot += lds(certb + 4u * (uint32_t)(buf * 32 * W + 32 * i + lane));
        tot = warp_sum(tot);
        if (!(fabsf(tot - Zm) <= 2e-5f * Zm)) bad = 8;
      }
      if (lane == 0) {   // the store that last read this out buffer is done
        if (cx.NO == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
      }
      __syncwarp();
      pf.mark(0);
      const uint32_t gt = gaccb + (uint32_t)(buf * kSeg) * growb;
      const uint32_t ot = sm.out + 4u * (uint32_t)((c * cx.NO + ob) * rawsz);
      // tile row j holds the frame of step j: frame row r = j (c = 0) or rows-1-j (c = 1)
      const int rsign = c == 0 ? 1 : -1, rbase = c == 0 ? 0 : rows - 1;
      float rs[kSeg];     // per-row sum of the label posteriors of this lane's classes
#pragma unroll
      for (int j = 0; j < kSeg; ++j) rs[j] = 0.f;
      for (int r = 0; r < rounds; ++r) {
        const int cls = 32 * r + lane;
        const bool mine = cls < C && cls != a.blank;
        // lane cls == C takes the padding column; the lanes beyond it have nothing to do
        const uint32_t ga = gt + 4u * (uint32_t)(cls < C ? cls : C);
        float acc[kSeg];
#pragma unroll
        for (int j = 0; j < kSeg; ++j) acc[j] = cls <= C ? (float)ldsi(ga + (uint32_t)j * growb) * (1.f / kFix) : 0.f;
        if (cls <= C) {
#pragma unroll
          for (int j = 0; j < kSeg; ++j) stsi(ga + (uint32_t)j * growb, 0);   // the tile is clean for its next step
        }
        if (mine) {
          uint32_t dsto = ot + 4u * (uint32_t)(cls + rbase * C);
          const int32_t dstep = 4 * rsign * C;
#pragma unroll
          for (int j = 0; j < kSeg; ++j) {
            if (j < rows) sts(dsto, acc[j] * kappa);
            dsto += dstep;
          }
#pragma unroll
          for (int j = 0; j < kSeg; ++j) rs[j] += acc[j];
        }
      }
      pf.mark(1);
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
      pf.mark(2);
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
          __syncwarp();
          const float* src = sm.out_gen + (size_t)(c * cx.NO + ob) * rawsz;
          for (int q = lane; q < n; q += 32) dst[q] = src[q];
        }
      }
      if (lane == 0) bulk_commit();   // one group per step (possibly empty)
      __syncwarp();
      pf.mark(3);
      if (++buf == W + 1) buf = 0;
    }
    tick2<W>(c, pf);
  }
  pf.report("X", c, 0, 2, lane);
  if (lane == 0) bulk_wait_all<0>();
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0) WFST_HAZ(&a.hazard[cx.b], 8);
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
// W = 3 needs more than half an SM of shared memory: one block per SM, twice the registers
template <int K, int W>
__global__ void __launch_bounds__(Geo<K, W>::NT, (W <= 2 ? 2 : 1)) ctc_tick_kernel(Args a) {
  extern __shared__ __align__(16) float smem_raw[];
  using G = Geo<K, W>;
  constexpr int NT = G::NT;
  const int warp = threadIdx.x >> 5;
  const int C = a.C;
  Ctx cx;
  cx.lane = threadIdx.x & 31;
  cx.T = a.T; cx.C = C; cx.CP = C + 1;
  cx.nsd = a.nsd; cx.nfull = a.nfull; cx.r0 = a.r0; cx.r1 = a.r1; cx.Th = a.Th;
  cx.NAB = a.NAB; cx.NB = a.NB; cx.NR = a.NR; cx.NO = a.NO;
  cx.b = blockIdx.x;
  cx.want_grad = a.gradE != nullptr;
  cx.rawsz = (uint32_t)((kSeg * C + 3) & ~3);
  const int* y = a.targets + a.offsets[cx.b];
  cx.L = a.offsets[cx.b + 1] - a.offsets[cx.b];
  const Layout lay = make_layout<K, W>(C, a.NAB, a.NB, a.NR, a.NO);
  Smem sm;
  {
    const uint32_t base = smem_u32(smem_raw);
    sm.raw = base + 4u * (uint32_t)lay.raw;
    sm.out = base + 4u * (uint32_t)lay.out;
    sm.abuf = base + 4u * (uint32_t)lay.abuf;
    sm.bnd = base + 4u * (uint32_t)lay.bnd;
    sm.lexp = base + 4u * (uint32_t)lay.lexp;
    sm.cert = base + 4u * (uint32_t)lay.cert;
    sm.gacc = base + 4u * (uint32_t)lay.gacc;
    sm.ptile = base + 4u * (uint32_t)lay.ptile;
    sm.ringL = base + 4u * (uint32_t)lay.ringL;
    sm.ringR = base + 4u * (uint32_t)lay.ringR;
    sm.bars = base + 4u * (uint32_t)lay.bars;
    sm.zx = base + 4u * (uint32_t)lay.zx;
    sm.abuf_c = 4u * (uint32_t)lay.abuf_c;
    sm.bnd_c = 4u * (uint32_t)lay.bnd_c;
    sm.lexp_c = 4u * (uint32_t)lay.lexp_c;
    sm.cert_c = 4u * (uint32_t)lay.cert_c;
    sm.gacc_c = 4u * (uint32_t)lay.gacc_c;
    sm.ptile_c = 4u * (uint32_t)lay.ptile_c;
    sm.ring_c = 4u * (uint32_t)lay.ring_c;
    sm.ytab = reinterpret_cast<int*>(smem_raw + lay.ytab);
    sm.out_gen = smem_raw + lay.out;
    sm.prof_gen = smem_raw + lay.prof;
  }

  // flags: cleared by the block that owns them (no memset node in front of the kernel)
  if (threadIdx.x == 0) a.hazard[cx.b] = 0;

  // ------------------------------------------------------------------ setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumBars; ++i) bar_init(sm.bars, i, 1u);
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
    // and the blank states: leave it to the fallback kernel (reason 1); so are labels
    // outside [0, C)
    if (__syncthreads_or(has_bad)) flag = 1;
    if (2 * cx.L + 1 > G::Sp - 1) flag = 1;
  }
  if (flag && threadIdx.x == 0) WFST_HAZ(&a.hazard[cx.b], flag);
  __syncthreads();
  if (flag) return;

  // warps: live[0][0..W), live[1][0..W), rc[0][..], rc[1][..], X[0], X[1], P[0], P[1]
  if (warp < 2 * W) role_live<K, W>(a, sm, cx, warp / W, warp % W);
  else if (warp < 4 * W) role_rc<K, W>(a, sm, cx, (warp - 2 * W) / W, (warp - 2 * W) % W);
  else if (warp < 4 * W + 2) role_reduce<K, W>(a, sm, cx, warp - 4 * W);
  else role_producer<W>(a, sm, cx, warp - 4 * W - 2);
}

// ---- host side ----------------------------------------------------------------------
struct Cfg {
  int K, W;
  bool automatic;   // false: only when forced
};
static const Cfg kCfgs[] = {{4, 1, true}, {4, 2, true}, {6, 2, true}, {6, 3, true}, {4, 3, false}};
constexpr int kNumCfgs = (int)(sizeof(kCfgs) / sizeof(kCfgs[0]));
static int g_force_k = 0, g_force_w = 0;

static int pick_cfg(int max_target_len) {
  const int S = 2 * max_target_len + 1;
  int first = -1;
  for (int i = 0; i < kNumCfgs; ++i) {
    if (32 * kCfgs[i].K * kCfgs[i].W - 1 < S) continue;
    if (kCfgs[i].K == g_force_k && kCfgs[i].W == g_force_w) return i;
    if (!kCfgs[i].automatic) continue;
    if (first < 0 || 32 * kCfgs[i].K * kCfgs[i].W < 32 * kCfgs[first].K * kCfgs[first].W) first = i;
  }
  return first;
}

// ring depths the tick schedule needs: a step buffer lives from live[0]'s tick to X's (2W ticks
// later), a p tile from the tick before live[0]'s to rc[W-1]'s
template <int K, int W>
static bool pick_bufs(int C, int& NAB, int& NB, int& NR, int& NO, size_t& bytes) {
  NAB = 2 * W;       // abar rows, boundary state, lane exponents: live -> rc
  NB = 2 * W + 1;
  if (NAB > kMaxAB || NB > kMaxNB) return false;
  // raw tiles in flight / gradient staging tiles: 3 / 2 where they fit, then 2 / 2, then 2 / 1
  static const int tries[3][2] = {{kNR, 2}, {2, 2}, {2, 1}};
  for (int i = 0; i < 3; ++i) {
    NR = tries[i][0]; NO = tries[i][1];
    bytes = make_layout<K, W>(C, NAB, NB, NR, NO).total * sizeof(float);
    if (bytes <= (size_t)(227 * 1024)) return true;
  }
  return false;
}

template <int K, int W>
static int launch_kw(const Args& a, size_t smem, cudaStream_t st) {
  auto kern = ctc_tick_kernel<K, W>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  kern<<<a.B, Geo<K, W>::NT, smem, st>>>(a);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

#define WFST_TICK_DISPATCH(idx, EXPR)             \
  switch (idx) {                                  \
    case 0: { constexpr int K = 4, W = 1; EXPR; } break; \
    case 1: { constexpr int K = 4, W = 2; EXPR; } break; \
    case 2: { constexpr int K = 6, W = 2; EXPR; } break; \
    case 3: { constexpr int K = 6, W = 3; EXPR; } break; \
    case 4: { constexpr int K = 4, W = 3; EXPR; } break; \
    default: break;                               \
  }

static bool pick_bufs_cfg(int idx, int C, int& NAB, int& NB, int& NR, int& NO, size_t& bytes) {
  bool ok = false;
  WFST_TICK_DISPATCH(idx, ok = (pick_bufs<K, W>(C, NAB, NB, NR, NO, bytes)));
  return ok;
}
static int ckf_of(int idx) { return (kCfgs[idx].K + 2 + 3) & ~3; }

}  // namespace tickk

int ctc_tick_force_config(int K, int W) {
  tickk::g_force_k = K;
  tickk::g_force_w = W;
  return 0;
}

bool ctc_tick_eligible(int T, int C, int max_target_len) {
  if (T < 1 || C + 1 > 128) return false;
  const int idx = tickk::pick_cfg(max_target_len);
  if (idx < 0) return false;
  int nab, nb, nr, no;
  size_t bytes;
  return tickk::pick_bufs_cfg(idx, C, nab, nb, nr, no, bytes);
}

// two blocks of the selected configuration fit on one SM (what makes this kernel the faster one
// for batches beyond one block per SM)
bool ctc_tick_two_per_sm(int T, int C, int max_target_len) {
  if (T < 1 || C + 1 > 128) return false;
  const int idx = tickk::pick_cfg(max_target_len);
  if (idx < 0) return false;
  int nab, nb, nr, no;
  size_t bytes;
  if (!tickk::pick_bufs_cfg(idx, C, nab, nb, nr, no, bytes)) return false;
  return bytes + 1024 <= (size_t)(114 * 1024) && 32 * (4 * tickk::kCfgs[idx].W + 4) <= 512;
}

static int tick_nsd(int T) {
  const int a = T / 16, R = T - 16 * a;
  return a + (R > 0 ? 1 : 0);
}
static size_t tick_ckpt_bytes(int B, int T, int idx) {
  return align_up((size_t)B * 2 * tick_nsd(T) * 32 * tickk::kCfgs[idx].W * tickk::ckf_of(idx) * sizeof(float), 256);
}

size_t ctc_tick_workspace_bytes(int B, int T, int max_target_len) {
  // the largest over the configurations a forced one could select
  size_t n = 0;
  const int S = 2 * max_target_len + 1;
  for (int i = 0; i < tickk::kNumCfgs; ++i) {
    if (32 * tickk::kCfgs[i].K * tickk::kCfgs[i].W - 1 < S) continue;
    const size_t m = tick_ckpt_bytes(B, T, i);
    if (m > n && (tickk::kCfgs[i].automatic || (tickk::kCfgs[i].K == tickk::g_force_k && tickk::kCfgs[i].W == tickk::g_force_w))) n = m;
  }
  const int idx = tickk::pick_cfg(max_target_len);
  if (idx >= 0) n = tick_ckpt_bytes(B, T, idx);
  return n + align_up((size_t)B * sizeof(int), 256);
}

int launch_ctc_tick(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                    int blank, int max_target_len, const float* grad_scale, float* z_out,
                    float* gradE, void* workspace, int** hazard_out, cudaStream_t st) {
  using namespace tickk;
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
  if (idx < 0 || !pick_bufs_cfg(idx, C, a.NAB, a.NB, a.NR, a.NO, smem)) {
    set_error("no tick-chain CTC configuration for C=%d L=%d", C, max_target_len);
    return WFST_ERR_UNSUPPORTED;
  }
  a.ckpt = (float*)workspace;
  a.hazard = (int*)((char*)workspace + tick_ckpt_bytes(B, T, idx));
  *hazard_out = a.hazard;
  int rc = WFST_ERR_UNSUPPORTED;
  WFST_TICK_DISPATCH(idx, rc = (launch_kw<K, W>(a, smem, st)));
  if (rc == WFST_ERR_UNSUPPORTED) set_error("no tick-chain CTC instantiation for configuration %d", idx);
  return rc;
}

}  // namespace wfst
