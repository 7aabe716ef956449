// Scaled-probability ASG force-align lattice on the TICK schedule of ctc_tick.cu (read its header
// for the schedule: one named barrier per 8-frame step and component instead of per-resource
// mbarriers, rings sized by writer-reader distance, node posteriors summed by class with
// fixed-point shared-memory atomics in the recompute warps): forward_score + backward of
//   intersect(intersect(g_fal, g_transitions), g_emissions)      (criterions/asg.py:53-81,103-115,158)
// as a chain recursion, on the machinery of ctc_solo.cu (read that header first): one block per
// utterance, one warp set per time direction, meet in the middle, recompute from checkpoints,
// per-label reduction by lane = class.
//
// The force-align graph is a chain of L+1 nodes (asg.py:71-81): node n >= 1 carries label
// y[n-1] on its self loop and on the arc from node n-1.  Composed with the bigram graph the arc
// weights are tr[1+y[n-1]][y[n-1]] (self), tr[1+y[n-1]][y[n-2]] (advance) and tr[0][y[0]] (the arc
// out of the start node, used at frame 0).  In the probability domain, per frame,
//   v'[n] = (v[n] * self[n] + v[n-1] * adv[n]) * p_t[y[n-1]],   self/adv = exp(tr - max tr)
// with node 0 (source) holding mass 1 before the first frame (its own self / emission are 0) and,
// for the beta~ recursion, a virtual sink node L+1 feeding node L with coefficient 1.
// Orientation 0 keeps node n at slot n, orientation 1 at slot Sp-1-n: partners are mirror images
// lane by lane, slot by slot.
//
// Gradients: the live sweep stores BOTH pre-emission terms of every node and frame; the
// recompute warp multiplies them with its value of the partner node -> posteriors of the self
// and the advance arc (accumulated over time in registers, scattered into grad_transitions with
// one atomic per node at the end) and, summed, the node posterior that the reduction warp sums
// by label into the emission gradient.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "launchers.h"

namespace wfst {
namespace falk {

constexpr int kSeg = 8;               // frames per step / tile
constexpr int kEventEvery = 1;        // lanes are renormalised every step: label-only chains with exp(tr - max tr) on every arc decay ~7 bits per frame
constexpr int kUndef = -(1 << 19);    // "no exponent": lane holds only zeros
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNR = 3;                // raw (TMA) staging slots per producer warp
constexpr int kRD = 2;                // entries per chain ring (step parity)
constexpr int kMaxAB = 8;             // step buffers per component, at most
constexpr int kMaxNB = 12;            // p tiles per component, at most
constexpr int kRingF = 12;            // floats per ring slot: 9 boundary values, the exponent, pad

#ifdef WFST_FAL_DEBUG
#define WFST_HAZ(ptr, bits) do { if (blockIdx.x < 3) printf("hazard b=%d bits=%d line %d warp %d\n", (int)blockIdx.x, (int)(bits), __LINE__, (int)(threadIdx.x >> 5)); atomicOr(ptr, bits); } while (0)
#else
#define WFST_HAZ(ptr, bits) atomicOr(ptr, bits)
#endif

struct Args {
  const float* E;
  const float* tr;  // [C+1, C] transitions: row 0 = scores out of <s>, row 1+i = scores into label i
  const int* targets;
  const int* offsets;
  int B, T, C;
  const float* grad_scale;
  float sign;       // multiplies grad_scale (-1 in ASG: loss = Z_fcc - Z_fal)
  float* z_out;     // [B] log Z
  float* gradE;     // [B, T, C] or null (overwritten)
  float* gradTr;    // [C+1, C] or null (accumulated with atomics)
  float* ckpt;      // [B][2][nsd][32 W][CKF]
  int* hazard;      // [B]
  int nsd;          // steps per direction and phase
  int nfull;        // full (8-frame) steps per direction
  int r0, r1;       // frames of the partial step of direction 0 / 1 (next to the meeting point)
  int Th;           // first frame of direction 1's half
  int NAB, NB;
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
  static constexpr int Sp = K * NL;            // slots (nodes 0 .. L+1 of the chain + padding)
  static constexpr int SA = 2 * K + 1;         // term row: {self term, advance term} per slot, words per lane (odd: conflict-free)
  static constexpr int PADA = 4;               // zero words in front of a term row
  static constexpr int ROWW = PADA + SA * NL + 4;       // words per term row
  static constexpr uint32_t ROWB = 4u * ROWW;
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
__host__ __device__ inline Layout make_layout(int C, int NAB, int NB) {
  using G = Geo<K, W>;
  const int NG = W + 1;   // tiles / certificate terms: written by rc, read by X
  Layout L;
  const size_t rawsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
  const size_t CP = (size_t)C + 1;
  size_t p = 0;
  L.raw = p;    p += 2 * (size_t)kNR * rawsz + 32;                  // [c][slot][8*C] (+ slack)
  L.out = p;    p += 2 * 2 * rawsz;                                 // [c][ob][8*C]
  L.abuf_c = (size_t)NAB * kSeg * G::ROWW;
  L.abuf = p;   p += 2 * L.abuf_c;                                  // [c][buf][row][word]
  L.bnd_c = (size_t)NAB * G::BNDW;
  L.bnd = p;    p += 2 * L.bnd_c;                                   // [c][buf][word]: live state at the step boundary
  L.lexp_c = (size_t)NAB * G::NL;
  L.lexp = p;   p += 2 * L.lexp_c;                                  // [c][buf][gl] (int)
  L.cert_c = (size_t)NG * G::NL;
  L.cert = p;   p += 2 * L.cert_c;                                  // [c][buf][gl]: the lanes' certificate terms
  L.gacc_c = (size_t)NG * kSeg * CP;
  L.gacc = p;   p += 2 * L.gacc_c;                                  // [c][buf][row][class] (int, 2^-23 units)
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
  L.zx = p;     p += 32;   // Zm, eZ, ok, trmax, msum(double), zpart[kMaxW]{contrib, Emax}, endacc[2], bad flag
  L.ytab = p;   p += (size_t)G::Sp + 4;                             // targets of the utterance
  p = (p + 3) & ~(size_t)3;
  L.prof = p;
#ifdef WFST_PROFILE
  p += 2 * 2 * 16 * 2;
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
  float trmax;
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
  uint32_t labofs[K];   // byte offset in a p tile of row 0 of the label of slot i (the zero column for nodes without one)
  float selfc[K];       // exp(self-loop weight - max tr) of the node at slot i
  float advc[K];        // exp(weight of the arc from the node at slot i-1 - max tr)
};

// orientation o keeps node n at slot n (o = 0) or Sp-1-n (o = 1); nodes: 0 = source, 1..L = the
// target positions (node n carries label y[n-1]), L+1 = sink of the beta~ recursion
template <int K, int W>
__device__ __forceinline__ void build_topo(Topo<K>& tp, const Ctx& cx, const int* ytab, const float* tr, int gl, int o) {
  constexpr int Sp = Geo<K, W>::Sp;
  const int C = cx.C, L = cx.L;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int j = gl * K + i;
    const int n = o == 0 ? j : Sp - 1 - j;
    int col = C;
    float sf = 0.f, ad = 0.f;
    if (n >= 1 && n <= L) {
      const int y = ytab[n - 1];
      col = y;
      sf = __expf(tr[(1 + y) * C + y] - cx.trmax);
      if (o == 0) ad = n == 1 ? __expf(tr[y] - cx.trmax) : __expf(tr[(1 + y) * C + ytab[n - 2]] - cx.trmax);
      else ad = n == L ? 1.f : __expf(tr[(1 + ytab[n]) * C + y] - cx.trmax);   // the arc n -> n+1, or the sink
    }
    tp.labofs[i] = 36u * (uint32_t)col;
    tp.selfc[i] = sf;
    tp.advc[i] = ad;
  }
}

template <int K>
struct PRow {
  float pl[K];
};
template <int K>
struct TileAddr {
  uint32_t la[K];
};
template <int K>
__device__ __forceinline__ TileAddr<K> tile_addr(const Topo<K>& tp, uint32_t pt) {
  TileAddr<K> t;
#pragma unroll
  for (int i = 0; i < K; ++i) t.la[i] = pt + tp.labofs[i];
  return t;
}
// row `it` of the tile (4 * it is an immediate when `it` is)
template <int K>
__device__ __forceinline__ PRow<K> load_prow(const TileAddr<K>& t, int it) {
  PRow<K> p;
  const uint32_t o = 4u * (uint32_t)it;
#pragma unroll
  for (int i = 0; i < K; ++i) p.pl[i] = lds(t.la[i] + o);
  return p;
}

// One frame.  v: with-emission values of the previous frame (own scale); on return this
// frame's with-emission values and, if WANT_TERMS, the two pre-emission terms of every slot.
// in1: the left neighbour's last slot, already converted to this lane's scale.
template <int K, bool WANT_TERMS>
__device__ __forceinline__ void step(float (&v)[K], float (&ts)[K], float (&ta)[K], const Topo<K>& tp, const PRow<K>& p, float in1) {
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    const float a1 = (i >= 1) ? v[i - 1] : in1;
    if (WANT_TERMS) {
      ts[i] = v[i] * tp.selfc[i];
      ta[i] = a1 * tp.advc[i];
      v[i] = (ts[i] + ta[i]) * p.pl[i];
    } else {
      v[i] = fmaf(v[i], tp.selfc[i], a1 * tp.advc[i]) * p.pl[i];
    }
  }
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
  constexpr int kChain = (kSeg * kEventEvery + K - 1) / K + 1;   // lanes a wave can cross in a window (one slot per frame)
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
  const uint32_t raw0 = sm.raw + 4u * (uint32_t)(c * kNR) * rawsz;
  const int tbar = kBarTma + c * kNR;
  const float* Eb = a.E + (size_t)cx.b * T * C;
  while (ps.fetched < kt_end && ps.fetched < ps.converted + kNR) {
    int lo_, rows;
    comp_seg(cx, c, ps.fetched, lo_, rows);
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
  }
}

// tile kt (== ps.converted) from its raw slot into its p-tile buffer
__device__ __forceinline__ void producer_convert(const Smem& sm, const Ctx& cx, ProducerState& ps, const int c,
                                                 const int kt, const bool phase1, Prof& pf) {
  pf.mark0();
  const int lane = cx.lane, C = cx.C;
  const uint32_t rawsz = cx.rawsz;
  const uint32_t raw0 = sm.raw + 4u * (uint32_t)(c * kNR) * rawsz;
  const int tbar = kBarTma + c * kNR;
  const int groups = (C + 31) >> 5;
  const int slot = ps.converted % kNR;
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
    a.z_out[cx.b] = ok ? (float)(log((double)Zm) + (double)eZ * 0.6931471805599453 + (ps.msum + msh[0]) +
                                 (double)cx.T * (double)cx.trmax)   // every path takes T arcs scaled by exp(-max tr)
                           : kNegInf;
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
  constexpr int Sp = G::Sp, NL = G::NL;
  const int lane = cx.lane, gl = 32 * w + lane, nsd = cx.nsd, NAB = cx.NAB;
  float* ck = a.ckpt + (((size_t)cx.b * 2 + c) * nsd * NL + gl) * G::CKF;
  Topo<K> tp;
  build_topo<K, W>(tp, cx, sm.ytab, a.tr, gl, c);

  float v[K], ts[K], ta[K];
  int e = kUndef;
  float f = 0.f;
  {
    // before the first frame all mass sits on the source (orientation 0: slot 0) / the sink
    // (orientation 1: slot Sp-2-L), whose own coefficients and emission are zero
    const int jstart = c == 0 ? 0 : Sp - 2 - cx.L;
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
  // renormalise (every step here), open my own entry
  auto step_begin = [&](int g) {
    const int slot = g & 1;
    rin = ring_in0 + 4u * (uint32_t)(slot * kRingF);
    rout = ring_out0 + 4u * (uint32_t)(slot * kRingF);
    const int Ein = w > 0 ? ldsi(rin + 36u) : kUndef;
    event1<K>(v, e, f, lane, Ein);
    if (w < W - 1 && lane == 31) {
      sts(rout, v[K - 1]);
      stsi(rout + 36u, e);
    }
  };
  // frames of a partial step
  auto slow_frames = [&](const TileAddr<K>& tad, int rows, uint32_t ar, bool want_terms) {
#pragma unroll 1
    for (int it = 0; it < rows; ++it) {
      const PRow<K> cur = load_prow<K>(tad, it);
      const float in1 = left_in(v[K - 1], lds(rin + 4u * (uint32_t)it), lane, f);
      step<K, true>(v, ts, ta, tp, cur, in1);
      if (want_terms) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
          sts(ar + (uint32_t)it * G::ROWB + 8u * i, ts[i]);
          sts(ar + (uint32_t)it * G::ROWB + 8u * i + 4u, ta[i]);
        }
      }
      if (lane == 31) sts(rout + 4u * (uint32_t)(it + 1), v[K - 1]);
    }
  };

  const uint32_t mybnd = 4u * (uint32_t)(4 + gl * K);       // my slots in a boundary row
  const uint32_t myrow = 4u * (uint32_t)(G::PADA + gl * G::SA);

  // ------------------------------------------------------------------ phase 1
  Prof pf;
  pf.setup(sm.prof_gen, tick_mask<W>(c, 1));
  int pbuf = 0;
  tick1<W>(c, pf);   // tile 0 is there
  for (int n = 0; n < nsd + W; ++n) {
    const int g = n - w;
    if (g >= 0 && g < nsd) {
      step_begin(g);
      ckpt_store<K, G::CKF>(ck + (size_t)g * NL * G::CKF, v, e);
      const bool partial = has_partial && g == cx.nfull;
      const TileAddr<K> tad = tile_addr<K>(tp, ptile_addr(sm, cx, c, pbuf));
      if (++pbuf == cx.NB) pbuf = 0;
      if (!partial) {
        PRow<K> nx = load_prow<K>(tad, 0);
        float bvn = lds(rin);
#pragma unroll
        for (int it = 0; it < kSeg; ++it) {
          const PRow<K> cur = nx;
          const float bv = bvn;
          if (it + 1 < kSeg) { nx = load_prow<K>(tad, it + 1); bvn = lds(rin + 4u * (it + 1)); }
          const float in1 = left_in(v[K - 1], bv, lane, f);
          step<K, false>(v, ts, ta, tp, cur, in1);
          if (lane == 31) sts(rout + 4u * (it + 1), v[K - 1]);
        }
      } else {
        slow_frames(tad, c == 0 ? cx.r0 : cx.r1, 0u, false);
      }
    } else if (g == nsd) {
      // meeting: every live warp renormalises (consistent exponents for the successor sums
      // below); the alpha warps publish their state in the layout of a boundary row (buffer 0)
      step_begin(nsd);
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
    // partner of my slot i is slot K-1-i of lane NL-1-gl
    const int pl = NL - 1 - gl;
    const uint32_t pblock = sm.bnd + 4u * (uint32_t)(4 + pl * K);
    const float in1 = left_in(v[K - 1], lds(rin), lane, f);
    float pm = 0.f;
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
      const float a1 = (i >= 1) ? v[i - 1] : in1;
      const float bb = fmaf(v[i], tp.selfc[i], a1 * tp.advc[i]);
      pm = fmaf(bb, lds(pblock + 4u * (K - 1 - i)), pm);
    }
    const int ea = ldsi(sm.lexp + 4u * (uint32_t)pl);
    int Em = kUndef;
    if (pm > 0.f && defined_exp(e) && defined_exp(ea)) Em = e + ea;
    int Emax = Em;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Emax = max(Emax, __shfl_xor_sync(kFull, Emax, o));
    float contrib = defined_exp(Em) ? pm * pow2c(Em - Emax) : 0.f;
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
#ifdef WFST_FAL_DEBUG
    if (blockIdx.x < 3) printf("b=%d Z: tot %g Emax %d ok %d trmax %g\n", (int)blockIdx.x, tot, Emax, (int)ok, cx.trmax);
#endif
    // reason 2: infeasible or out of range -- the log-semiring kernel decides
    if (!ok) { WFST_HAZ(&a.hazard[cx.b], 2); stsi(sm.zx + 72u, 1); }
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
      step_begin(nsd + 1 + k2);
      {
        stsi(lexpb + 4u * (uint32_t)(buf * NL + gl), e);
        // state at the step boundary: the recompute warps check Z against it (certificate)
        const uint32_t bb = bndb + (uint32_t)buf * G::BNDB + mybnd;
#pragma unroll
        for (int i = 0; i < K; ++i) sts(bb + 4u * i, v[i]);
      }
      // component c continues through the other direction's steps, last one (the partial one) first
      const bool partial = has_partial && k2 == 0;
      const TileAddr<K> tad = tile_addr<K>(tp, ptile_addr(sm, cx, c, pbuf));
      if (++pbuf == cx.NB) pbuf = 0;
      const uint32_t ar = abuf + (uint32_t)(buf * kSeg) * G::ROWB + myrow;
      if (!partial) {
        PRow<K> nx = load_prow<K>(tad, 0);
        float bvn = lds(rin);
#pragma unroll
        for (int it = 0; it < kSeg; ++it) {
          const PRow<K> cur = nx;
          const float bv = bvn;
          if (it + 1 < kSeg) { nx = load_prow<K>(tad, it + 1); bvn = lds(rin + 4u * (it + 1)); }
          const float in1 = left_in(v[K - 1], bv, lane, f);
          step<K, true>(v, ts, ta, tp, cur, in1);
#pragma unroll
          for (int i = 0; i < K; ++i) {
            sts(ar + (uint32_t)it * G::ROWB + 8u * i, ts[i]);
            sts(ar + (uint32_t)it * G::ROWB + 8u * i + 4u, ta[i]);
          }
          if (lane == 31) sts(rout + 4u * (it + 1), v[K - 1]);
        }
      } else {
        slow_frames(tad, c == 0 ? cx.r1 : cx.r0, ar, true);
      }
      if (++buf == NAB) buf = 0;
    }
    tick2<W>(c, pf);
  }
  pf.report("live", c, w, 2, lane);
}

// ---------------------------------------------------------------------------
// rc[c][w]: runs direction 1-c over the steps of live[c]'s phase 2, from the checkpoints
// live[1-c] wrote in phase 1, against live[c]'s step order; the products of its values with the
// stored terms are the arc posteriors (kept in registers, summed over time) and, added, the node
// posteriors (written back in place of the self term for the reduction warp).
// ---------------------------------------------------------------------------
template <int K, int W>
__device__ __forceinline__ void role_rc(const Args& a, const Smem& sm, const Ctx& cx, const int c, const int w,
                                        float (&gself)[K], float (&gadv)[K], bool& ran) {
  using G = Geo<K, W>;
  constexpr int NL = G::NL;
  const int lane = cx.lane, gl = 32 * w + lane, nsd = cx.nsd, NAB = cx.NAB;
#pragma unroll
  for (int i = 0; i < K; ++i) { gself[i] = 0.f; gadv[i] = 0.f; }
  ran = false;
  __syncthreads();   // meeting A
  __syncthreads();   // meeting B
  __syncthreads();   // meeting C: phase 1 (and every checkpoint) is complete, Z published
  const bool ok = lds(sm.zx + 8u) != 0.f;
  if (!cx.want_grad || !ok) return;
  ran = true;
  const int eZ = ldsi(sm.zx + 4u);
  Topo<K> tp;
  build_topo<K, W>(tp, cx, sm.ytab, a.tr, gl, 1 - c);
  const float* ck = a.ckpt + (((size_t)cx.b * 2 + (1 - c)) * nsd * NL + gl) * G::CKF;
  const int pl = NL - 1 - gl;
  const uint32_t prow = 4u * (uint32_t)(G::PADA + pl * G::SA);    // partner block in a term row
  const uint32_t pbnd = 4u * (uint32_t)(4 + pl * K);              // partner block in a boundary row
  const uint32_t ring_in0 = sm.ringR + (uint32_t)c * sm.ring_c + 4u * (uint32_t)(w * kRD * kRingF);
  const uint32_t ring_out0 = ring_in0 + 4u * (uint32_t)(kRD * kRingF);
  const uint32_t abuf = sm.abuf + (uint32_t)c * sm.abuf_c, bndb = sm.bnd + (uint32_t)c * sm.bnd_c,
                 lexpb = sm.lexp + (uint32_t)c * sm.lexp_c, certb = sm.cert + (uint32_t)c * sm.cert_c;
  int bad = 0;   // reason 4: scale overflow when pairing live and recomputed values
  uint32_t gofs[K];
#pragma unroll
  for (int i = 0; i < K; ++i) gofs[i] = tp.labofs[i] / 9u;   // 36 col -> 4 col (the padding column for nodes without a label)
  const uint32_t gaccb = sm.gacc + (uint32_t)c * sm.gacc_c;
  const uint32_t growb = 4u * (uint32_t)cx.CP;
  float wv[K], nv[K], dts[K], dta[K];
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
    float gsc = 0.f, frs = 0.f;
    {
      const uint32_t le = lexpb + 4u * (uint32_t)(buf * NL);
      const int ep = ldsi(le + 4u * (uint32_t)pl);                          // partner lane
      const int el = gl > 0 ? ldsi(le + 4u * (uint32_t)(pl + 1)) : kUndef;  // partner of my left neighbour
      if (defined_exp(ep)) {
        if (defined_exp(ew)) {
          const int dd = ew + ep - eZ;
          // a lane without mass carries an inherited exponent: its scale is irrelevant
          float wmax = wv[0];
#pragma unroll
          for (int i = 1; i < K; ++i) wmax = fmaxf(wmax, wv[i]);
          if (dd > 126) { if (wmax > 0.f) bad |= 4; }
          else gsc = pow2c(dd);
        }
        if (defined_exp(el)) frs = pow2c(ep - el);     // <= 2^D by the event invariant
      }
    }
#pragma unroll
    for (int i = 0; i < K; ++i) wv[i] *= gsc;
    // chain ring: my left neighbour wrote its entry of this step one tick ago; open my own
    if (w < W - 1 && lane == 31) sts(rout, wv[K - 1]);
    const TileAddr<K> tad = tile_addr<K>(tp, ptile_addr(sm, cx, c, pbuf));
    if (++pbuf == cx.NB) pbuf = 0;
    const uint32_t ar = abuf + (uint32_t)(buf * kSeg) * G::ROWB + prow;
    const uint32_t gt0 = gaccb + (uint32_t)(gbuf * kSeg) * growb;
    // In the lower half (c = 1) a stored term at frame t belongs to an arc taken at frame t+1; the
    // arcs into frame Th are already counted by the upper half: the first row of the first step
    // contributes node posteriors only.
    auto frame = [&](int it, float bv, const PRow<K>& cur, bool count_arcs) {
      const uint32_t arow = ar + (uint32_t)it * G::ROWB;
      float tsv[K], tav[K];
#pragma unroll
      for (int i = 0; i < K; ++i) {     // partner of my slot i is slot K-1-i of the partner block
        tsv[i] = lds(arow + 8u * (K - 1 - i));
        tav[i] = lds(arow + 8u * (K - 1 - i) + 4u);
      }
      const float in1 = left_in(wv[K - 1], bv, lane, frs);
      step<K, false>(wv, dts, dta, tp, cur, in1);
      // node posteriors (times Zm < 1) summed by class with integer atomics in 2^-23 units: the
      // float 1 + x has the exponent of 1.0 and round(x 2^23) as its mantissa field
      const uint32_t gt = gt0 + (uint32_t)it * growb;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const float ps = tsv[i] * wv[i], pa = tav[i] * wv[i];
        if (count_arcs) { gself[i] += ps; gadv[i] += pa; }
        const int q = __float_as_int((ps + pa) + 1.f) - 0x3f800000;
        asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(gt + gofs[i]), "r"(q) : "memory");
      }
    };
    if (!partial) {
      // against the live step order
      PRow<K> nx = load_prow<K>(tad, kSeg - 1);
      float bvn = lds(rin);
#pragma unroll
      for (int it = kSeg - 1; it >= 0; --it) {
        const PRow<K> cur = nx;
        const float bv = bvn;
        if (it > 0) { nx = load_prow<K>(tad, it - 1); bvn = lds(rin + 4u * (kSeg - it)); }
        frame(it, bv, cur, !(c == 1 && k2 == 0 && it == 0));
        if (lane == 31) sts(rout + 4u * (kSeg - it), wv[K - 1]);
      }
    } else {
#pragma unroll 1
      for (int it = rows - 1; it >= 0; --it) {
        const PRow<K> cur = load_prow<K>(tad, it);
        frame(it, lds(rin + 4u * (uint32_t)(rows - 1 - it)), cur, !(c == 1 && it == 0));
        if (lane == 31) sts(rout + 4u * (uint32_t)(rows - it), wv[K - 1]);
      }
    }
    {
      // certificate: sum_n v_live(n) * (successor sum of w)(n) at the step boundary must be Z
      // (float32 range can only be exceeded by losing mass or producing inf / NaN)
      const float in1 = left_in(wv[K - 1], lds(rin + 4u * (uint32_t)rows), lane, frs);
      const uint32_t bb = bndb + (uint32_t)buf * G::BNDB + pbnd;
      float acc = 0.f;
#pragma unroll
      for (int i = K - 1; i >= 0; --i) {
        const float a1 = (i >= 1) ? wv[i - 1] : in1;
        const float sx = fmaf(wv[i], tp.selfc[i], a1 * tp.advc[i]);
        acc = fmaf(sx, lds(bb + 4u * (K - 1 - i)), acc);
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
  if (bad && lane == 0) { WFST_HAZ(&a.hazard[cx.b], bad); stsi(sm.zx + 72u, 1); }
}
// ---------------------------------------------------------------------------
// X[c]: turns a step's integer tile (node posteriors summed by class by the recompute warps) into
// the [8, C] gradient tile and stores it.  Lane = class (classes beyond 32 in further rounds).
// ---------------------------------------------------------------------------
template <int K, int W>
__device__ __forceinline__ void role_reduce(const Args& a, const Smem& sm, const Ctx& cx, const int c) {
  const int lane = cx.lane, nsd = cx.nsd, T = cx.T, C = cx.C;
  const int rounds = (C + 31) >> 5;
  __syncthreads();   // meeting A
  __syncthreads();   // meeting B
  __syncthreads();   // meeting C: Z published
  const bool ok = lds(sm.zx + 8u) != 0.f;
  if (!cx.want_grad || !ok) return;
  const uint32_t rawsz = cx.rawsz;
  const float Zm = lds(sm.zx);
  const float kappa = a.sign * (a.grad_scale ? a.grad_scale[cx.b] : 1.f) / Zm * (1.f / 8388608.f);
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
      const int ob = k2 & 1;
      {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < W; ++i) tot += lds(certb + 4u * (uint32_t)(buf * 32 * W + 32 * i + lane));
        tot = warp_sum(tot);
        if (!(fabsf(tot - Zm) <= 2e-5f * Zm)) bad = 8;
      }
      if (lane == 0) bulk_wait_read<1>();   // the store that last read this out buffer is done
      __syncwarp();
      const uint32_t gt = gaccb + (uint32_t)(buf * kSeg) * growb;
      const uint32_t ot = sm.out + 4u * (uint32_t)((c * 2 + ob) * rawsz);
      // tile row j holds the frame of step j: frame row r = j (c = 0) or rows-1-j (c = 1)
      const int rsign = c == 0 ? 1 : -1, rbase = c == 0 ? 0 : rows - 1;
      for (int r = 0; r < rounds; ++r) {
        const int cls = 32 * r + lane;
        // lane cls == C takes the padding column; the lanes beyond it have nothing to do
        if (cls <= C) {
          const uint32_t ga = gt + 4u * (uint32_t)cls;
          int acc[kSeg];
#pragma unroll
          for (int j = 0; j < kSeg; ++j) acc[j] = ldsi(ga + (uint32_t)j * growb);
#pragma unroll
          for (int j = 0; j < kSeg; ++j) stsi(ga + (uint32_t)j * growb, 0);   // the tile is clean for its next step
          if (cls < C) {
            uint32_t dsto = ot + 4u * (uint32_t)(cls + rbase * C);
            const int32_t dstep = 4 * rsign * C;
#pragma unroll
            for (int j = 0; j < kSeg; ++j) {
              if (j < rows) sts(dsto, (float)acc[j] * kappa);
              dsto += dstep;
            }
          }
        }
      }
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
          const float* src = sm.out_gen + (size_t)(c * 2 + ob) * rawsz;
          for (int q = lane; q < n; q += 32) dst[q] = src[q];
        }
      }
      if (lane == 0) bulk_commit();   // one group per step (possibly empty)
      __syncwarp();
      if (++buf == W + 1) buf = 0;
    }
    tick2<W>(c, pf);
  }
  pf.report("X", c, 0, 2, lane);
  if (lane == 0) bulk_wait_all<0>();
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0) { WFST_HAZ(&a.hazard[cx.b], 8); stsi(sm.zx + 72u, 1); }
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <int K, int W>
__global__ void __launch_bounds__(Geo<K, W>::NT, (Geo<K, W>::NT <= 512 ? 2 : 1)) asg_fal_chain_kernel(Args a) {
  extern __shared__ __align__(16) float smem_raw[];
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
  // max of the transition scores (every arc weight enters as exp(tr - max))
  {
    float m = kNegInf;
    for (int k = threadIdx.x; k < (C + 1) * C; k += NT) m = fmaxf(m, a.tr[k]);
    m = warp_max(m);
    __shared__ float wm[32];
    if (cx.lane == 0) wm[warp] = m;
    __syncthreads();
    float mm = wm[0];
    for (int k = 1; k < G::NWARPS; ++k) mm = fmaxf(mm, wm[k]);
    cx.trmax = (mm > kNegInf && mm < -kNegInf) ? mm : 0.f;
  }
  int flag = 0;
  {
    int has_bad = 0;
    for (int n = threadIdx.x; n < cx.L && n < G::Sp; n += NT) {
      const int yy = y[n];
      sm.ytab[n] = yy;
      if (yy < 0 || yy >= C) has_bad = 1;
    }
    // labels outside [0, C), empty targets, targets the chain cannot hold: the log-semiring kernel (reason 1)
    if (__syncthreads_or(has_bad)) flag = 1;
    if (cx.L < 1 || cx.L + 2 > G::Sp) flag = 1;
  }
  if (flag && threadIdx.x == 0) WFST_HAZ(&a.hazard[cx.b], flag);
  __syncthreads();
  if (flag) return;

  // warps: live[0][0..W), live[1][0..W), rc[0][..], rc[1][..], X[0], X[1], P[0], P[1]
  float gself[K], gadv[K];
  bool rc_ran = false;
  const bool is_rc = warp >= 2 * W && warp < 4 * W;
  const int rc_c = (warp - 2 * W) / W, rc_w = (warp - 2 * W) % W;
  if (warp < 2 * W) role_live<K, W>(a, sm, cx, warp / W, warp % W);
  else if (is_rc) role_rc<K, W>(a, sm, cx, rc_c, rc_w, gself, gadv, rc_ran);
  else if (warp < 4 * W + 2) role_reduce<K, W>(a, sm, cx, warp - 4 * W);
  else role_producer<W>(a, sm, cx, warp - 4 * W - 2);
  // ---- transition gradient: arc posteriors summed over time, one atomic per node and arc kind;
  // nothing from an utterance that was flagged (the log-semiring kernel redoes it)
  __syncthreads();
  if (!a.gradTr || ldsi(sm.zx + 72u) != 0 || lds(sm.zx + 8u) == 0.f || !cx.want_grad) return;
  const float Zm = lds(sm.zx);
  const float kap = a.sign * (a.grad_scale ? a.grad_scale[cx.b] : 1.f) / Zm;
  if (is_rc && rc_ran) {
    const int gl = 32 * rc_w + cx.lane;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int j = gl * K + i;
      const int n = rc_c == 0 ? G::Sp - 1 - j : j;      // rc[c] runs orientation 1-c
      if (n < 1 || n > cx.L) continue;
      const int yn = sm.ytab[n - 1];
      if (gself[i] != 0.f) atomicAdd(&a.gradTr[(1 + yn) * C + yn], kap * gself[i]);
      if (gadv[i] != 0.f) {
        if (rc_c == 0) {          // live alpha's advance term: the arc n-1 -> n (the start arc for n = 1)
          const int idx = n == 1 ? yn : (1 + yn) * C + sm.ytab[n - 2];
          atomicAdd(&a.gradTr[idx], kap * gadv[i]);
        } else if (n < cx.L) {    // live beta's advance term: the arc n -> n+1 (n = L: the sink, no parameter)
          atomicAdd(&a.gradTr[(1 + sm.ytab[n]) * C + yn], kap * gadv[i]);
        }
      }
    }
  }
  // the arc out of the start node is taken exactly once, at frame 0
  if (threadIdx.x == 0) atomicAdd(&a.gradTr[sm.ytab[0]], kap * Zm);
}

// ---- host side ----------------------------------------------------------------------
constexpr int kK = 6;

static int pick_w(int max_target_len) {
  for (int w = 1; w <= 2; ++w)
    if (32 * kK * w >= max_target_len + 2) return w;
  return 0;
}

// ring depths the tick schedule needs (ctc_tick.cu): step buffers live -> rc, p tiles P -> rc[W-1]
template <int K, int W>
static bool pick_bufs(int C, int& NAB, int& NB, size_t& bytes) {
  NAB = 2 * W;
  NB = 2 * W + 1;
  if (NAB > kMaxAB || NB > kMaxNB) return false;
  bytes = make_layout<K, W>(C, NAB, NB).total * sizeof(float);
  return bytes <= (size_t)(227 * 1024);
}
static bool pick_bufs_w(int W, int C, int& NAB, int& NB, size_t& bytes) {
  return W == 1 ? pick_bufs<kK, 1>(C, NAB, NB, bytes) : pick_bufs<kK, 2>(C, NAB, NB, bytes);
}

template <int K, int W>
static int launch_kw(const Args& a, size_t smem, cudaStream_t st) {
  auto kern = asg_fal_chain_kernel<K, W>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  kern<<<a.B, Geo<K, W>::NT, smem, st>>>(a);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

}  // namespace falk

bool asg_fal_chain_eligible(int T, int C, int max_target_len) {
  if (T < 2 || C + 1 > 128 || max_target_len < 1) return false;
  const int W = falk::pick_w(max_target_len);
  if (W == 0) return false;
  int nab, nb;
  size_t bytes;
  return falk::pick_bufs_w(W, C, nab, nb, bytes);
}

static int fal_nsd(int T) {
  const int a = T / 16, R = T - 16 * a;
  return a + (R > 0 ? 1 : 0);
}
static size_t fal_ckpt_bytes(int B, int T, int W) {
  return align_up((size_t)B * 2 * fal_nsd(T) * 32 * W * ((falk::kK + 2 + 3) & ~3) * sizeof(float), 256);
}

size_t asg_fal_chain_workspace_bytes(int B, int T, int max_target_len) {
  const int W = falk::pick_w(max_target_len);
  if (W == 0) return 0;
  return fal_ckpt_bytes(B, T, W) + align_up((size_t)B * sizeof(int), 256);
}

// scores[b] = log Z of the force-align lattice; gradE (overwritten, every element) =
// sign * grad_scale[b] * node posterior by label; gradTr += sign * grad_scale[b] * arc posterior.
// hazard_out: per utterance != 0 where the caller must run the log-semiring kernel instead.
int launch_asg_fal_chain(const float* E, const float* tr, const int* targets, const int* offsets, int B,
                         int T, int C, int max_target_len, const float* grad_scale, float sign,
                         float* scores, float* gradE, float* gradTr, void* workspace, int** hazard_out,
                         cudaStream_t st) {
  using namespace falk;
  const int W = pick_w(max_target_len);
  Args a{};
  a.E = E; a.tr = tr; a.targets = targets; a.offsets = offsets; a.B = B; a.T = T; a.C = C;
  a.grad_scale = grad_scale; a.sign = sign; a.z_out = scores; a.gradE = gradE; a.gradTr = gradTr;
  a.nfull = T / 16;
  const int R = T - 16 * a.nfull;
  a.r0 = (R + 1) / 2;
  a.r1 = R / 2;
  a.nsd = a.nfull + (R > 0 ? 1 : 0);
  a.Th = kSeg * a.nfull + a.r0;
  size_t smem = 0;
  if (W == 0 || !pick_bufs_w(W, C, a.NAB, a.NB, smem)) {
    set_error("no force-align chain configuration for C=%d L=%d", C, max_target_len);
    return WFST_ERR_UNSUPPORTED;
  }
  a.ckpt = (float*)workspace;
  a.hazard = (int*)((char*)workspace + fal_ckpt_bytes(B, T, W));
  *hazard_out = a.hazard;
  return W == 1 ? launch_kw<kK, 1>(a, smem, st) : launch_kw<kK, 2>(a, smem, st);
}

}  // namespace wfst
