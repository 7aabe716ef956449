// Fast CTC forward+backward for sm_100a: scaled-probability recursion with the
// whole CTC chain of an utterance resident in the registers of ONE warp per
// direction.  Replaces, for CTC, the reference's per-utterance
//   create_ctc_graph -> intersect -> forward_score -> backward
// (criterions/ctc.py:15-29,40-51,78-81).
//
// Layout of one thread block (= one utterance), 96 threads:
//   warp 0 "A": alpha direction (time ascending),  states j = s
//   warp 1 "B": beta  direction (time descending), states j = Sp-1-s  (mirrored)
//   warp 2 "P": producer — TMA-loads [16, C] emission tiles into shared memory and
//               turns them into p[t,c] = exp(E[t,c] - max_c E[t,c]) tiles for A and B
// Both directions run the SAME recursion (the beta recursion written for
// beta~_t(s) = p_t(lab s) * beta_t(s) is the alpha recursion on the reversed target
// and reversed time):  v'[j] = (v[j] + v[j-1] + skip[j] * v[j-2]) * p_t[lab j].
// Lane l owns K consecutive states j in [l*K, (l+1)*K); neighbours come from one or
// two warp shuffles per frame.  Values are float32 mantissas with one power-of-two
// exponent per lane, re-normalised every 16 frames ("event").
//
// Schedule (meet in the middle + recompute; nothing of size T x S ever leaves the SM):
//   phase 1: A sweeps segments [0, nA), B sweeps segments [nA, nseg) downwards; each
//            writes a checkpoint (K values + exponent per lane) per 16-frame segment.
//   meeting: Z = sum_s alpha(s) * beta(s) at the boundary.
//   phase 2: A continues upwards through [nA, nseg): per segment it re-runs the beta
//            recursion from B's checkpoint (stored in shared memory, scaled so that
//            stored * live = posterior * Zm), then advances alpha and multiplies;
//            B does the mirror image downwards through [0, nA).
//   The per-state posteriors of a segment are reduced over states with equal label by a
//   "transposed" pass (lane = frame) and leave as a [16, C] tile via a bulk async store.
//
// Robustness: a frame's posteriors must sum to one.  Every row sum is checked against Z
// (|sum - Z| <= 1e-3 Z, finite); any violation (possible only if float32 range was
// exceeded inside a 16-frame window) flags the utterance in `hazard`, and the
// log-semiring kernel (lattice.cuh, CtcTopo) recomputes it.  No CPU fallback.
#include "common.cuh"
#include "launchers.h"

namespace wfst {

#ifdef WFST_PROFILE
#define PROF_DECL long long pf_t0 = clock64(), pf_acc[8] = {0,0,0,0,0,0,0,0}
#define PROF_MARK(i) do { long long pf_t1 = clock64(); pf_acc[i] += pf_t1 - pf_t0; pf_t0 = pf_t1; } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#endif

constexpr int kSeg = 16;              // frames per segment / tile
constexpr int kUndef = -(1 << 19);    // "no exponent": lane holds only zeros
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNB = 2;                // p-tile ring depth per direction
constexpr int kNR = 4;                // raw (TMA) staging slots of the producer
constexpr int kMaxPass = 4;           // transposed pass: up to 4 x 32 chunk slots per gamma row

struct CtcFastArgs {
  const float* E;
  const int* targets;
  const int* offsets;
  int B, T, C, blank;
  const float* grad_scale;
  float* z_out;     // [B] log Z
  float* gradE;     // [B, T, C] or null
  float* ckpt;      // [B][2][nseg + 1][K + 1][32]
  int* hazard;      // [B]
  int nseg, nA;
  int Cp;           // p-tile row stride (odd, > C); column C is always 0
  int RS;           // row stride of the stored/gamma buffer (= 4 mod 32, >= Sp and >= gamma columns)
};

// Explicit shared-state-space accesses on 32-bit addresses: generic pointers made the
// compiler emit LD.E with 64-bit address arithmetic (5 instructions per load).
__device__ __forceinline__ float lds(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ int2 lds64i(uint32_t a) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}

__device__ __forceinline__ bool defined_exp(int e) { return e > kUndef / 2; }
__device__ __forceinline__ float pow2i(int d) {  // 2^d for d in [-126, 127]
  return __uint_as_float((uint32_t)(d + 127) << 23);
}

// ---------------------------------------------------------------------------
// Per-lane static description of its K states in one orientation.
// ---------------------------------------------------------------------------
template <int K>
struct LaneTopo {
  uint32_t labofs[K / 2];  // byte offset in a p-tile row of each label-type slot's label
                           // (column C, always zero, for padding slots)
  float skipm[K / 2];      // 1 if the skip arc into the label-type slot exists
  uint32_t gofs[K / 2];    // byte offset of the slot's posterior in the sorted gamma row
};

// ODD: label-type states sit at odd slots (orientation 0) or even slots (orientation 1)
template <int K, bool ODD>
__device__ __forceinline__ void build_topo(LaneTopo<K>& tp, int lane, const int* y, int L, int C,
                                           const int* gcol_of_pos, int dump_col) {
  constexpr int Sp = 32 * K;
  const int S = 2 * L + 1;
#pragma unroll
  for (int q = 0; q < K / 2; ++q) {
    const int i = ODD ? 2 * q + 1 : 2 * q;
    const int j = lane * K + i;
    const int s = ODD ? j : Sp - 1 - j;   // true state; label states have odd s
    int col = C;
    float sk = 0.f;
    int gc = dump_col;
    if (s >= 1 && s < S) {
      const int n = (s - 1) >> 1;
      col = y[n];
      gc = gcol_of_pos[n];
      // skip arc from two positions earlier IN THIS ORIENTATION
      const int n2 = ODD ? n - 1 : n + 1;
      if (n2 >= 0 && n2 < L && y[n2] != y[n]) sk = 1.f;
    }
    tp.labofs[q] = 4u * (uint32_t)col;
    tp.skipm[q] = sk;
    tp.gofs[q] = 4u * (uint32_t)gc;
  }
}

// The p values one frame needs: the label-type slots' labels and the blank.
template <int K>
struct PRow {
  float pl[K / 2];
  float pb;
};
template <int K>
__device__ __forceinline__ PRow<K> load_prow(const LaneTopo<K>& tp, uint32_t prow, uint32_t blank_ofs) {
  PRow<K> p;
#pragma unroll
  for (int q = 0; q < K / 2; ++q) p.pl[q] = lds(prow + tp.labofs[q]);
  p.pb = lds(prow + blank_ofs);
  return p;
}

// One frame of the recursion. v: with-emission values of the previous frame (own scale).
// On return v holds this frame's with-emission values and, if WANT_ABAR, abar the
// pre-emission sums.  f converts the left neighbour's scale to ours (0 in lane 0).
// The p values are loaded by the caller one frame ahead (the shared-memory accessors are
// volatile asm, so the software pipelining has to be explicit).
template <int K, bool ODD, bool WANT_ABAR>
__device__ __forceinline__ void step(float (&v)[K], float (&abar)[K], const LaneTopo<K>& tp,
                                     const PRow<K>& p, float f) {
  const float in1 = __shfl_up_sync(kFull, v[K - 1], 1) * f;
  float in2 = 0.f;
  if (!ODD) in2 = __shfl_up_sync(kFull, v[K - 2], 1) * f;
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    const bool lab = ((i & 1) == 1) == ODD;
    const float a1 = (i >= 1) ? v[i - 1] : in1;
    float s = v[i] + a1;
    if (lab) {
      const int q = i >> 1;
      const float a2 = (i >= 2) ? v[i - 2] : (i == 1 ? in1 : in2);
      s = fmaf(tp.skipm[q], a2, s);
      if (WANT_ABAR) abar[i] = s;
      v[i] = s * p.pl[q];
    } else {
      if (WANT_ABAR) abar[i] = s;
      v[i] = s * p.pb;
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
// exponent, D) or (-inf, 0), which is associative: a 5-step warp scan.
template <int K>
__device__ __forceinline__ void event(float (&v)[K], int& e, float& f, int lane) {
  constexpr int kChain = (32 + K - 1) / K + 1;   // lanes a wave can cross in 16 frames
  constexpr int D = 96 / kChain;
  float m = v[0];
#pragma unroll
  for (int i = 1; i < K; ++i) m = fmaxf(m, v[i]);
  int eown = kUndef;
  if (m > 0.f) {
    int ex = (int)((__float_as_uint(m) >> 23) & 0xffu) - 127;
    ex = min(max(ex, -126), 126);
    const float sc = pow2i(-ex);
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] *= sc;
    eown = (defined_exp(e) ? e : 0) + ex;
  }
  // (c, d) packed in one int: c * 2048 + d, d < 2048
  int c = eown, d = defined_exp(eown) ? D : 0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int packed = __shfl_up_sync(kFull, c * 2048 + d, o);
    const int pd = packed & 2047;
    const int pc = (packed - pd) / 2048;
    if (lane >= o) {
      if (defined_exp(pc)) c = defined_exp(c) ? max(c, pc - d) : pc - d;
      d += pd;
    }
  }
  const int E = c;
  if (defined_exp(E) && defined_exp(eown) && E != eown) {
    const int sh = eown - E;  // < 0
    const float sc = (sh < -126) ? 0.f : pow2i(sh);
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] *= sc;
  }
  e = defined_exp(E) ? E : kUndef;
  const int el = __shfl_up_sync(kFull, e, 1);
  if (lane == 0 || !defined_exp(el) || !defined_exp(e)) {
    f = 0.f;
  } else {
    const int dd = el - e;  // <= D by construction
    f = (dd < -126) ? 0.f : pow2i(min(dd, 126));
  }
}

// checkpoint I/O: ckpt[slot][lane], slot K holds the exponent
template <int K>
__device__ __forceinline__ void ckpt_store(float* base, const float (&v)[K], int e, int lane) {
#pragma unroll
  for (int i = 0; i < K; ++i) base[i * 32 + lane] = v[i];
  base[K * 32 + lane] = __int_as_float(e);
}
template <int K>
__device__ __forceinline__ void ckpt_load(const float* base, float (&v)[K], int& e, int lane) {
#pragma unroll
  for (int i = 0; i < K; ++i) v[i] = base[i * 32 + lane];
  e = __float_as_int(base[K * 32 + lane]);
}

// ---------------------------------------------------------------------------
// shared memory carve-up
// ---------------------------------------------------------------------------
template <int K>
struct FastSmem {
  static constexpr int Sp = 32 * K;
  // per direction
  // (two explicit members instead of arrays: indexing an array of pointers with the
  // runtime direction would push the whole struct into local memory)
  float *stored0, *stored1;   // [kSeg][RS] recomputed opposite-direction values; row r is re-used
                              // for the sorted per-state posteriors ("gamma") once it has been read
  float *pbk0, *pbk1;         // [kSeg][33] blank partials
  float *out0, *out1;         // [2][kSeg*C] output tiles (double buffered)
  float *ptile0, *ptile1;     // [kNB][kSeg][Cp]
  __device__ __forceinline__ float* stored(int d) const { return d ? stored1 : stored0; }
  __device__ __forceinline__ float* pbk(int d) const { return d ? pbk1 : pbk0; }
  __device__ __forceinline__ float* out(int d) const { return d ? out1 : out0; }
  __device__ __forceinline__ float* ptile(int d) const { return d ? ptile1 : ptile0; }
  // shared
  float* raw;         // [kNR][kSeg*C] TMA staging (16B aligned)
  int* gcolpos;       // [Sp/2] gamma column of target position n
  int* slotlab;       // [32*kMaxPass + 8] label of each 4-column chunk slot of the gamma row (-1: unused)
  int* hist;          // [C + 4] scratch for the counting sort; [C]: #slots, [C+1]: max chunks per label
  uint64_t* bars;     // full[2][kNB], empty[2][kNB], tma[2], zready
  float* zx;          // Zm, eZ (as int bits), valid flag
  double* msum;       // sum of per-frame maxima (phase-1 frames)
};

__host__ __device__ inline size_t fast_smem_floats(int K, int C, int Cp, int RS) {
  const int Sp = 32 * K;
  size_t per_dir = (size_t)kSeg * RS + kSeg * 33 + 2 * (((size_t)kSeg * C + 3) & ~3) +
                   (size_t)kNB * kSeg * Cp;
  per_dir = (per_dir + 3) & ~(size_t)3;
  size_t shared = kNR * (((size_t)kSeg * C + 3) & ~3) + Sp / 2 + (32 * kMaxPass + 8) + (C + 4) + 2 * 16 + 8 + 4;
  return 2 * per_dir + shared + 16;
}

template <int K>
__device__ __forceinline__ FastSmem<K> carve_fast(float* base, int C, int Cp, int RS) {
  constexpr int Sp = 32 * K;
  FastSmem<K> s;
  float* p = base;
  const size_t outsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
  s.raw = p; p += kNR * outsz;                     // 16B aligned (base is)
  s.out0 = p; p += 2 * outsz;
  s.out1 = p; p += 2 * outsz;
  s.stored0 = p; p += (size_t)kSeg * RS;
  s.stored1 = p; p += (size_t)kSeg * RS;
  s.bars = reinterpret_cast<uint64_t*>(p); p += 2 * 16;      // up to 16 barriers
  s.msum = reinterpret_cast<double*>(p); p += 4;
  s.zx = p; p += 4;
  s.pbk0 = p; p += kSeg * 33;
  s.pbk1 = p; p += kSeg * 33;
  s.ptile0 = p; p += (size_t)kNB * kSeg * Cp;
  s.ptile1 = p; p += (size_t)kNB * kSeg * Cp;
  s.gcolpos = reinterpret_cast<int*>(p); p += Sp / 2;
  s.slotlab = reinterpret_cast<int*>(p); p += 32 * kMaxPass + 8;
  s.hist = reinterpret_cast<int*>(p); p += C + 4;
  return s;
}

// barrier indices
__device__ __forceinline__ int bar_full(int d, int i) { return d * kNB + i; }
__device__ __forceinline__ int bar_empty(int d, int i) { return 2 * kNB + d * kNB + i; }
constexpr int kBarTma = 4 * kNB;      // + raw slot
constexpr int kBarZ = 4 * kNB + kNR;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------
// One direction of one utterance (a whole warp).  DIR 0: alpha, time ascending, states
// j = s.  DIR 1: beta, time descending, mirrored states j = Sp-1-s.
// ---------------------------------------------------------------------------
template <int K, int DIR>
__device__ __forceinline__ void run_direction(const CtcFastArgs& a, const FastSmem<K>& sm, int b, int lane) {
  constexpr int Sp = 32 * K;
  constexpr int dir = DIR;
  const int T = a.T, C = a.C, Cp = a.Cp, RS = a.RS;
  const int* y = a.targets + a.offsets[b];
  const int L = a.offsets[b + 1] - a.offsets[b];
  const int nseg = a.nseg, nA = a.nA;
  const bool want_grad = a.gradE != nullptr;
  const int dump_col = RS - 1;
  float* ck_own = a.ckpt + ((size_t)b * 2 + dir) * (size_t)(nseg + 1) * (K + 1) * 32;
  const float* ck_other = a.ckpt + ((size_t)b * 2 + (1 - dir)) * (size_t)(nseg + 1) * (K + 1) * 32;
  const int blank = a.blank;

  LaneTopo<K> tp_live, tp_rc;   // live orientation = dir, recompute orientation = 1 - dir
  if (dir == 0) {
    build_topo<K, true>(tp_live, lane, y, L, C, sm.gcolpos, dump_col);
    build_topo<K, false>(tp_rc, lane, y, L, C, sm.gcolpos, dump_col);
  } else {
    build_topo<K, false>(tp_live, lane, y, L, C, sm.gcolpos, dump_col);
    build_topo<K, true>(tp_rc, lane, y, L, C, sm.gcolpos, dump_col);
  }

  float v[K], abar[K];
  int e = kUndef;
  float f = 0.f;
#pragma unroll
  for (int i = 0; i < K; ++i) v[i] = 0.f;
  {
    // virtual pre-frame state: all mass on the start state of this orientation
    const int S = 2 * L + 1;
    const int jstart = (dir == 0) ? 0 : Sp - S;
    const bool mine = (jstart / K == lane);
    const int jm = jstart % K;
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = (mine && jm == i) ? 1.f : 0.f;   // selects, no dynamic index
    if (mine) e = 0;
  }

  PROF_DECL;
  uint32_t full_phase = 0u;   // bit per p-tile buffer
  auto seg_of = [&](int k) { return dir == 0 ? k : nseg - 1 - k; };
  const uint32_t ptile_u32 = smem_u32(sm.ptile(dir));
  const uint32_t row_bytes = 4u * (uint32_t)Cp;
  const uint32_t blank_ofs = 4u * (uint32_t)a.blank;
  // returns the shared address of the tile's first row
  auto wait_ptile = [&](int k) -> uint32_t {
    const int buf = k % kNB;
    mbar_wait(&sm.bars[bar_full(dir, buf)], (full_phase >> buf) & 1u);
    full_phase ^= 1u << buf;
    return ptile_u32 + (uint32_t)buf * kSeg * row_bytes;
  };
  auto release_ptile = [&](int k) {
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.bars[bar_empty(dir, k % kNB)]);
  };

  // ------------------------------------------------------------------ phase 1
  const int n1 = (dir == 0) ? nA : nseg - nA;
  for (int k = 0; k < n1; ++k) {
    const int seg = seg_of(k);
    const int rows = min(kSeg, T - seg * kSeg);
    event<K>(v, e, f, lane);
    ckpt_store<K>(ck_own + (size_t)seg * (K + 1) * 32, v, e, lane);
    PROF_MARK(0);
    const uint32_t pt = wait_ptile(k);
    PROF_MARK(1);
    if (dir == 0) {
      uint32_t pr = pt;
      PRow<K> nx = load_prow<K>(tp_live, pr, blank_ofs);
#pragma unroll 2
      for (int r = 0; r < rows; ++r) {
        const PRow<K> cur = nx;
        pr += (r + 1 < rows) ? row_bytes : 0u;
        nx = load_prow<K>(tp_live, pr, blank_ofs);
        step<K, true, false>(v, abar, tp_live, cur, f);
      }
    } else {
      uint32_t pr = pt + (uint32_t)(rows - 1) * row_bytes;
      PRow<K> nx = load_prow<K>(tp_live, pr, blank_ofs);
#pragma unroll 2
      for (int r = rows - 1; r >= 0; --r) {
        const PRow<K> cur = nx;
        pr -= (r > 0) ? row_bytes : 0u;
        nx = load_prow<K>(tp_live, pr, blank_ofs);
        step<K, false, false>(v, abar, tp_live, cur, f);
      }
    }
    release_ptile(k);
    PROF_MARK(2);
  }

  // ------------------------------------------------------------------ meeting: Z
  // A publishes its state (extra checkpoint slot nseg of its own area); B combines.
  if (dir == 0) ckpt_store<K>(ck_own + (size_t)nseg * (K + 1) * 32, v, e, lane);
  named_sync(1, 64);
  if (dir == 1) {
    event<K>(v, e, f, lane);     // consistent exponents / f for the shuffle below
    // pre-emission sums of B's next frame: bb = v[j] + v[j-1] + skip * v[j-2]
    float bb[K];
    {
      const float in1 = __shfl_up_sync(kFull, v[K - 1], 1) * f;
      const float in2 = __shfl_up_sync(kFull, v[K - 2], 1) * f;
#pragma unroll
      for (int i = K - 1; i >= 0; --i) {
        const bool lab = (i & 1) == 0;   // orientation 1: label-type states at even slots
        const float a1 = (i >= 1) ? v[i - 1] : in1;
        float s = v[i] + a1;
        if (lab) {
          const float a2 = (i >= 2) ? v[i - 2] : in2;
          s = fmaf(tp_live.skipm[i >> 1], a2, s);
        }
        bb[i] = s;
      }
    }
    float av[K];
    int ea;
    ckpt_load<K>(ck_other + (size_t)nseg * (K + 1) * 32, av, ea, 31 - lane);
    float P = 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) P = fmaf(bb[i], av[K - 1 - i], P);
    int Eabs = kUndef;
    if (P > 0.f && defined_exp(e) && defined_exp(ea)) Eabs = e + ea;
#ifdef WFST_DEBUG_Z
    if (b == 0) printf("Z lane %d: P=%g e=%d ea=%d f=%g v0=%g vK=%g bb0=%g av0=%g avK=%g\n", lane, P, e, ea, f, v[0], v[K-1], bb[0], av[0], av[K-1]);
#endif
    int Emax = Eabs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Emax = max(Emax, __shfl_xor_sync(kFull, Emax, o));
    float contrib = 0.f;
    if (defined_exp(Eabs)) {
      const int dd = Eabs - Emax;
      contrib = (dd < -126) ? 0.f : P * pow2i(dd);
    }
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
      sm.zx[0] = Zm;
      sm.zx[1] = __int_as_float(ok ? Emax + ex : 0);
      sm.zx[2] = ok ? 1.f : 0.f;
      if (!ok) a.hazard[b] = 2;   // reason 2: infeasible or out of range — the log-semiring kernel decides
    }
  }
  named_sync(1, 64);
  if (dir == 1 && lane == 0) mbar_arrive(&sm.bars[kBarZ]);
  const float Zm = sm.zx[0];
  const int eZ = __float_as_int(sm.zx[1]);
  const bool zok = sm.zx[2] != 0.f;
  if (!want_grad) return;

  // ------------------------------------------------------------------ phase 2
  const int n2 = (dir == 0) ? nseg - nA : nA;
  if (!zok) {
    // keep the producer's ring moving so that it can terminate
    for (int k2 = 0; k2 < n2; ++k2) { wait_ptile(n1 + k2); release_ptile(n1 + k2); }
    return;
  }
  const float gs = a.grad_scale ? a.grad_scale[b] : 1.f;
  const float kappa = -gs / Zm;
  float* gEb = a.gradE + (size_t)b * T * C;
  const uint32_t stored_u32 = smem_u32(sm.stored(dir));
  const uint32_t pbk_u32 = smem_u32(sm.pbk(dir));
  const uint32_t srow_bytes = 4u * (uint32_t)RS;
  const uint32_t my_block = 4u * (uint32_t)(lane * K);          // where my recomputed values go
  const uint32_t pair_block = 4u * (uint32_t)((31 - lane) * K); // the block of the lane paired with me
  const size_t outsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
  int bad = 0;   // reason bits: 4 = scale overflow in the recompute, 8 = row-sum certificate
  int obuf = 0;
  // transposed pass, labels: lane = chunk slot of the gamma row (32 slots per pass); the
  // slots of one label are adjacent lanes and are combined by a segmented suffix sum
  const int nslots = sm.hist[C];
  const int npass = (nslots + 31) >> 5;
  const int maxch = sm.hist[C + 1];
  int slab[kMaxPass];
  float lk1[kMaxPass], lk2[kMaxPass], lk4[kMaxPass];
  bool head[kMaxPass];
#pragma unroll
  for (int p = 0; p < kMaxPass; ++p) {
    const int g = 32 * p + lane;
    const int me = sm.slotlab[g];
    slab[p] = me;
    // links never leave the pass: a label's slots do not straddle a multiple of 32
    lk1[p] = (me >= 0 && lane + 1 < 32 && sm.slotlab[g + 1] == me) ? 1.f : 0.f;
    lk2[p] = (me >= 0 && lane + 2 < 32 && sm.slotlab[g + 2] == me) ? 1.f : 0.f;
    lk4[p] = (me >= 0 && lane + 4 < 32 && sm.slotlab[g + 4] == me) ? 1.f : 0.f;
    head[p] = me >= 0 && (lane == 0 || sm.slotlab[g - 1] != me);
  }
  // transposed pass, blank: lane = (frame, half) sums 16 of the 32 lane partials
  const int frm = lane & 15, hf = lane >> 4;
  const uint32_t pbrow_u32 = pbk_u32 + 4u * (uint32_t)(frm * 33 + hf * 16);
  const uint32_t slot_ofs = 16u * (uint32_t)lane;

  // checkpoint of the first phase-2 segment (the next one is prefetched inside the loop)
  float w[K];
  int ew;
  if (n2 > 0) ckpt_load<K>(ck_other + (size_t)seg_of(n1) * (K + 1) * 32, w, ew, lane);

  for (int k2 = 0; k2 < n2; ++k2) {
    const int k = n1 + k2;
    const int seg = seg_of(k);
    const int rows = min(kSeg, T - seg * kSeg);
    PROF_MARK(3);
    event<K>(v, e, f, lane);
    PROF_MARK(0);
    const uint32_t pt = wait_ptile(k);
    PROF_MARK(1);

    // ---- recompute the opposite direction over this segment in the complementary scale:
    // stored * live = posterior * Zm, i.e. exponent(stored lane) = eZ - exponent(live lane)
    {
      const int ex_live = __shfl_sync(kFull, e, 31 - lane);   // the live lane paired with me
      int erc = kUndef;
      float sc = 0.f;
      if (defined_exp(ex_live)) {
        erc = eZ - ex_live;
        if (defined_exp(ew)) {
          const int dd = ew - erc;
          if (dd > 126) bad |= 4;
          else sc = (dd < -126) ? 0.f : pow2i(dd);
        }
      }
#pragma unroll
      for (int i = 0; i < K; ++i) w[i] *= sc;
      const int el = __shfl_up_sync(kFull, erc, 1);
      float fr = 0.f;
      if (lane > 0 && defined_exp(el) && defined_exp(erc)) {
        const int dd = el - erc;
        fr = (dd < -126) ? 0.f : pow2i(min(dd, 126));
      }
      // the recompute walks the frames in the opposite order to the live sweep
      if (dir == 0) {
        uint32_t pr = pt + (uint32_t)(rows - 1) * row_bytes;
        uint32_t dst = stored_u32 + (uint32_t)(rows - 1) * srow_bytes + my_block;
        PRow<K> nx = load_prow<K>(tp_rc, pr, blank_ofs);
#pragma unroll 2
        for (int r = rows - 1; r >= 0; --r, dst -= srow_bytes) {
          const PRow<K> cur = nx;
          pr -= (r > 0) ? row_bytes : 0u;
          nx = load_prow<K>(tp_rc, pr, blank_ofs);
          step<K, false, false>(w, abar, tp_rc, cur, fr);
#pragma unroll
          for (int i = 0; i < K; i += 4) sts128(dst + 4u * i, w[i], w[i + 1], w[i + 2], w[i + 3]);
        }
      } else {
        uint32_t pr = pt;
        uint32_t dst = stored_u32 + my_block;
        PRow<K> nx = load_prow<K>(tp_rc, pr, blank_ofs);
#pragma unroll 2
        for (int r = 0; r < rows; ++r, dst += srow_bytes) {
          const PRow<K> cur = nx;
          pr += (r + 1 < rows) ? row_bytes : 0u;
          nx = load_prow<K>(tp_rc, pr, blank_ofs);
          step<K, true, false>(w, abar, tp_rc, cur, fr);
#pragma unroll
          for (int i = 0; i < K; i += 4) sts128(dst + 4u * i, w[i], w[i + 1], w[i + 2], w[i + 3]);
        }
      }
      // prefetch the checkpoint of the next segment (consumed at the top of the next iteration)
      if (k2 + 1 < n2) ckpt_load<K>(ck_other + (size_t)seg_of(k + 1) * (K + 1) * 32, w, ew, lane);
      __syncwarp();
    }
    PROF_MARK(4);

    // ---- live sweep over the segment: posterior(state) * Zm = abar * stored.
    // Row r+1's stored block and p values are fetched while row r is computed.  A block is
    // read by exactly one lane, which clears it right away so that the row can be re-used
    // for the sorted posteriors with every padding column reading as zero.
    {
      const int rstep = (dir == 0) ? 1 : -1;
      int r = (dir == 0) ? 0 : rows - 1;
      uint32_t pr = pt + (uint32_t)r * row_bytes;
      uint32_t row = stored_u32 + (uint32_t)r * srow_bytes;
      float stn[K];
      auto fetch_block = [&](uint32_t rw) {
#pragma unroll
        for (int i = 0; i < K; i += 4) {
          const float4 q = lds128(rw + pair_block + 4u * i);
          stn[i] = q.x; stn[i + 1] = q.y; stn[i + 2] = q.z; stn[i + 3] = q.w;
          sts128(rw + pair_block + 4u * i, 0.f, 0.f, 0.f, 0.f);
        }
      };
      fetch_block(row);
      PRow<K> nx = load_prow<K>(tp_live, pr, blank_ofs);
#pragma unroll 2
      for (int it = 0; it < rows; ++it, r += rstep) {
        float st[K];
#pragma unroll
        for (int i = 0; i < K; ++i) st[i] = stn[i];
        const PRow<K> cur = nx;
        const uint32_t row_cur = row;
        const bool more = it + 1 < rows;
        if (more) {
          pr = (dir == 0) ? pr + row_bytes : pr - row_bytes;
          row = (dir == 0) ? row + srow_bytes : row - srow_bytes;
          fetch_block(row);
        }
        nx = load_prow<K>(tp_live, pr, blank_ofs);
        __syncwarp();   // every lane has fetched (and cleared) its block of row_cur
        if (dir == 0) step<K, true, true>(v, abar, tp_live, cur, f);
        else step<K, false, true>(v, abar, tp_live, cur, f);
        float pbsum = 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const bool lab = ((i & 1) == 1) == (dir == 0);
          const float g = abar[i] * st[K - 1 - i];
          if (lab) sts(row_cur + tp_live.gofs[i >> 1], g);
          else pbsum += g;
        }
        sts(pbk_u32 + 4u * (uint32_t)(r * 33 + lane), pbsum);
      }
    }
    release_ptile(k);
    __syncwarp();
    PROF_MARK(5);

    // ---- transposed pass: sum the posteriors per label and write the [rows, C] tile
    {
      float* ot = sm.out(dir) + (size_t)obuf * outsz;
      const uint32_t ot_u32 = smem_u32(ot);
      if (lane == 0) bulk_wait_read<1>();   // the store that last read this buffer is done
      __syncwarp();
      PROF_MARK(6);
      float tot = 0.f;
#pragma unroll
      for (int p = 0; p < kMaxPass; ++p) {
        if (p < npass) {
          const uint32_t src = stored_u32 + 512u * (uint32_t)p + slot_ofs;
          const uint32_t dsto = ot_u32 + 4u * (uint32_t)max(slab[p], 0);
#pragma unroll
          for (int r0 = 0; r0 < kSeg; r0 += 8) {
            if (r0 < rows) {
              float c[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {   // rows >= `rows` hold finite stale data; never stored
                const float4 q = lds128(src + (uint32_t)(r0 + j) * srow_bytes);
                c[j] = (q.x + q.y) + (q.z + q.w);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) c[j] = fmaf(__shfl_down_sync(kFull, c[j], 1), lk1[p], c[j]);
              if (maxch > 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) c[j] = fmaf(__shfl_down_sync(kFull, c[j], 2), lk2[p], c[j]);
              }
              if (maxch > 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j) c[j] = fmaf(__shfl_down_sync(kFull, c[j], 4), lk4[p], c[j]);
              }
              if (head[p]) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (r0 + j < rows) {
                    sts(dsto + 4u * (uint32_t)((r0 + j) * C), c[j] * kappa);
                    tot += c[j];
                  }
                }
              }
            }
          }
        }
      }
      float bs = 0.f;
      if (frm < rows) {
#pragma unroll
        for (int q = 0; q < 16; ++q) bs += lds(pbrow_u32 + 4u * q);
      }
      tot += bs;
      const float bsum = bs + __shfl_xor_sync(kFull, bs, 16);
      if (frm < rows && hf == 0) sts(ot_u32 + 4u * (uint32_t)(frm * C) + blank_ofs, bsum * kappa);
      // certificate: the posteriors of every frame sum to one, i.e. the segment sums to
      // rows * Zm.  Range loss can only remove mass, so deficits cannot cancel.
      tot = warp_sum(tot);
      if (!(fabsf(tot - (float)rows * Zm) <= 2e-5f * (float)rows * Zm)) bad |= 8;
      PROF_MARK(7);
      float* dst = gEb + (size_t)seg * kSeg * C;
      const int n = rows * C;
      const bool tma = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n & 3) == 0);
      if (tma) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          bulk_s2g(dst, ot, (uint32_t)n * 4u);
          bulk_commit();
        }
      } else {
        __syncwarp();
        for (int q = lane; q < n; q += 32) dst[q] = ot[q];
      }
      obuf ^= 1;
      __syncwarp();
    }
    PROF_MARK(3);
  }
#ifdef WFST_PROFILE
  if (b == 0 && lane == 0)
    printf("dir %d cycles: event+ckpt %lld  wait_ptile %lld  phase1-steps %lld  fence+store %lld  recompute %lld  combine %lld  bulkwait %lld transposed %lld\n",
           dir, pf_acc[0], pf_acc[1], pf_acc[2], pf_acc[3], pf_acc[4], pf_acc[5], pf_acc[6], pf_acc[7]);
#endif
  if (lane == 0) bulk_wait_all<0>();
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0) atomicOr(&a.hazard[b], bad);
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(96, 2) ctc_fast_kernel(CtcFastArgs a) {
  constexpr int Sp = 32 * K;
  extern __shared__ __align__(16) float smem_raw[];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T, C = a.C, Cp = a.Cp, RS = a.RS;
  FastSmem<K> sm = carve_fast<K>(smem_raw, C, Cp, RS);
  const int* y = a.targets + a.offsets[b];
  const int L = a.offsets[b + 1] - a.offsets[b];
  const int nseg = a.nseg, nA = a.nA;
  const bool want_grad = a.gradE != nullptr;
  const float* Eb = a.E + (size_t)b * T * C;
  const int dump_col = RS - 1;

  // ------------------------------------------------------------------ setup
  if (threadIdx.x == 0) {
    for (int d = 0; d < 2; ++d)
      for (int i = 0; i < kNB; ++i) {
        mbar_init(&sm.bars[bar_full(d, i)], 1);
        mbar_init(&sm.bars[bar_empty(d, i)], 1);
      }
    for (int i = 0; i < kNR; ++i) mbar_init(&sm.bars[kBarTma + i], 1);
    mbar_init(&sm.bars[kBarZ], 1);
    fence_barrier_init();
  }
  // zero the regions that rely on it: p-tile padding columns, output tiles (labels that
  // do not occur in the target keep a zero gradient), blank partials
  for (int d = 0; d < 2; ++d) {
    for (int k = threadIdx.x; k < kNB * kSeg * Cp; k += 96) sm.ptile(d)[k] = 0.f;
    for (int k = threadIdx.x; k < 2 * (int)(((size_t)kSeg * C + 3) & ~(size_t)3); k += 96) sm.out(d)[k] = 0.f;
    for (int k = threadIdx.x; k < kSeg * 33; k += 96) sm.pbk(d)[k] = 0.f;
    for (int k = threadIdx.x; k < kSeg * RS; k += 96) sm.stored(d)[k] = 0.f;
  }
  // Counting sort of the target positions by label -> columns of the "gamma" row.  The row
  // is organised in chunk slots of 4 columns; a label with n occurrences owns ceil(n/4)
  // consecutive slots that never straddle a multiple of 32 slots (one pass of the
  // transposed reduction = 32 slots, one per lane).
  for (int k = threadIdx.x; k < C + 4; k += 96) sm.hist[k] = 0;
  for (int k = threadIdx.x; k < 32 * kMaxPass + 8; k += 96) sm.slotlab[k] = -1;
  __syncthreads();
  int has_blank = 0;
  for (int n = threadIdx.x; n < L; n += 96) {
    atomicAdd(&sm.hist[y[n]], 1);
    has_blank |= (y[n] == a.blank);
  }
  if (__syncthreads_or(has_blank)) {
    // a target that contains the blank label shares a gradient column between a label
    // state and the blank states: leave it to the log-semiring kernel
    if (threadIdx.x == 0) a.hazard[b] = 1;   // reason 1: blank label inside the target
    return;
  }
  if (threadIdx.x == 0) {
    int slot = 0, maxch = 0;
    for (int c = 0; c < C; ++c) {
      const int cnt = sm.hist[c];
      sm.hist[c] = 0;
      if (cnt == 0) continue;
      const int nch = (cnt + 3) >> 2;
      if ((slot & 31) + nch > 32) slot = (slot + 31) & ~31;
      sm.hist[c] = slot * 4;                // becomes the column cursor of label c
      for (int j = 0; j < nch && slot + j < 32 * kMaxPass; ++j) sm.slotlab[slot + j] = c;
      slot += nch;
      maxch = max(maxch, nch);
    }
    sm.hist[C] = slot;
    sm.hist[C + 1] = maxch;
  }
  __syncthreads();
  {
    const int nslots = sm.hist[C], maxch = sm.hist[C + 1];
    // layouts the transposed pass cannot hold (very many distinct labels for this K, or one
    // label more than 32 times) go to the log-semiring kernel
    if (nslots > 32 * kMaxPass || nslots * 4 > RS - 4 || maxch > 8) {
      if (threadIdx.x == 0) a.hazard[b] = 16;  // reason 16: gamma row layout does not fit
      return;
    }
  }
  for (int n = threadIdx.x; n < L; n += 96) sm.gcolpos[n] = atomicAdd(&sm.hist[y[n]], 1);
  __syncthreads();

  // ===================================================================== producer
  if (warp == 2) {
    // Direction 0 consumes tiles 0,1,...; direction 1 consumes nseg-1, nseg-2, ...
    // Without a gradient only the phase-1 tiles are needed.  The schedule is the sequence
    // (k, d), k = 0.., d = 0, 1, restricted to k < ntile[d].  Raw [16, C] tiles are
    // fetched kNR - 1 entries ahead (TMA latency >> conversion time).
    const int ntile0 = want_grad ? nseg : nA, ntile1 = want_grad ? nseg : (nseg - nA);
    const int total = max(ntile0, ntile1);
    const int fr = lane & 15, hh = lane >> 4;
    const int c0 = hh ? (C + 1) / 2 : 0, c1 = hh ? C : (C + 1) / 2;
    const int rawsz = (kSeg * C + 3) & ~3;
    const uint32_t raw_u32 = smem_u32(sm.raw);
    double msum = 0.0;
    uint32_t tma_phase = 0u;     // bit per raw slot
    uint32_t tma_used = 0u;      // bit per raw slot: the entry in it came through TMA
    uint32_t empty_phase = 0u;   // bit (d * kNB + buf)
    auto tile_of = [&](int d, int k) { return d == 0 ? k : nseg - 1 - k; };
    auto next_entry = [&](int& k, int& d) {
      do {
        if (d == 0) d = 1; else { d = 0; ++k; }
      } while (k < total && k >= (d == 0 ? ntile0 : ntile1));
    };
    auto issue_raw = [&](int tile, int slot) {
      const int rows = min(kSeg, T - tile * kSeg);
      const float* src = Eb + (size_t)tile * kSeg * C;
      const uint32_t bytes = (uint32_t)rows * C * 4u;
      float* dst = sm.raw + (size_t)slot * rawsz;
      const bool tma = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15u) == 0);
      if (tma) {
        if (lane == 0) {
          mbar_expect_tx(&sm.bars[kBarTma + slot], bytes);
          bulk_g2s(dst, src, bytes, &sm.bars[kBarTma + slot]);
        }
        tma_used |= 1u << slot;
      } else {
        for (int q = lane; q < rows * C; q += 32) dst[q] = __ldg(src + q);
        tma_used &= ~(1u << slot);
        __syncwarp();
      }
    };
    // cursor of the entry being converted (k, d) and of the next entry to fetch (fk, fd)
    int k = -1, d = 1;
    next_entry(k, d);
    int fk = k, fd = d, fetched = 0, converted = 0;
    while (k < total) {
      // keep the raw ring full; slot = entry index mod kNR; a slot is free once its previous
      // entry has been converted (entries are converted in order)
      while (fk < total && fetched < converted + kNR) {
        issue_raw(tile_of(fd, fk), fetched % kNR);
        ++fetched;
        next_entry(fk, fd);
      }
      const int slot = converted % kNR;
      const int tile = tile_of(d, k);
      const int rows = min(kSeg, T - tile * kSeg);
      const int buf = k % kNB;
      if (k >= kNB) {  // wait until the consumer has released this p-tile buffer
        const int bit = d * kNB + buf;
        mbar_wait(&sm.bars[bar_empty(d, buf)], (empty_phase >> bit) & 1u);
        empty_phase ^= 1u << bit;
      }
      if ((tma_used >> slot) & 1u) {
        mbar_wait(&sm.bars[kBarTma + slot], (tma_phase >> slot) & 1u);
        tma_phase ^= 1u << slot;
      }
      const uint32_t er = raw_u32 + 4u * (uint32_t)(slot * rawsz + fr * C);
      const uint32_t pt = smem_u32(sm.ptile(d)) + 4u * (uint32_t)(buf * kSeg * Cp + fr * Cp);
      float mx = kNegInf;
      if (fr < rows) {
#pragma unroll 8
        for (int c = c0; c < c1; ++c) mx = fmaxf(mx, lds(er + 4u * c));
      }
      mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, 16));
      if (fr < rows) {
        // a row that is entirely -inf keeps p = 0 (dead frame); +inf / NaN rows surface
        // through the certificate
        const float base = (mx == kNegInf) ? 0.f : mx;
        const float nb = -base * 1.4426950408889634f;
#pragma unroll 8
        for (int c = c0; c < c1; ++c)
          sts(pt + 4u * c, exp2f(fmaf(lds(er + 4u * c), 1.4426950408889634f, nb)));
        const bool phase1 = (d == 0) ? (tile < nA) : (tile >= nA);
        if (hh == 0 && phase1) msum += (double)base;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.bars[bar_full(d, buf)]);
      ++converted;
      next_entry(k, d);
    }
    // loss: log Z = log(Zm) + eZ ln2 + sum_t max_t
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(kFull, msum, o);
    mbar_wait(&sm.bars[kBarZ], 0u);
    if (lane == 0) {
      const float Zm = sm.zx[0];
      const int eZ = __float_as_int(sm.zx[1]);
      const bool ok = sm.zx[2] != 0.f;
      a.z_out[b] = ok ? (float)(log((double)Zm) + (double)eZ * 0.6931471805599453 + msum) : kNegInf;
    }
    return;
  }

  // ===================================================================== A / B
  if (warp == 0) run_direction<K, 0>(a, sm, b, lane);
  else run_direction<K, 1>(a, sm, b, lane);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int fast_pick_k(int max_target_len) {
  const int S = 2 * max_target_len + 1;
  const int ks[] = {4, 8, 12, 16, 20, 24};
  for (int k : ks)
    if (32 * k >= S) return k;
  return 0;
}

static void fast_dims(int K, int C, int& Cp, int& RS) {
  Cp = (C + 1) | 1;
  // the row holds the Sp recomputed values, later re-used for the chunk slots of the
  // sorted posteriors (<= Sp / 4 slots) and the dump column RS - 1
  RS = 32 * K + 4;
}

bool ctc_fast_eligible(int T, int C, int max_target_len) {
  if (T < 1) return false;
  const int K = fast_pick_k(max_target_len);
  if (K == 0) return false;
  int Cp, RS;
  fast_dims(K, C, Cp, RS);
  return fast_smem_floats(K, C, Cp, RS) * sizeof(float) <= 227 * 1024;
}

size_t ctc_fast_workspace_bytes(int B, int T, int max_target_len) {
  const int K = fast_pick_k(max_target_len);
  const int nseg = (T + kSeg - 1) / kSeg;
  return align_up((size_t)B * 2 * (nseg + 1) * (K + 1) * 32 * sizeof(float), 256) +
         align_up((size_t)B * sizeof(int), 256);
}

template <int K>
static int launch_fast_k(const CtcFastArgs& a, cudaStream_t st) {
  size_t smem = fast_smem_floats(K, a.C, a.Cp, a.RS) * sizeof(float);
  auto kern = ctc_fast_kernel<K>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<a.B, 96, smem, st>>>(a);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

int launch_ctc_fast(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                    int blank, int max_target_len, const float* grad_scale, float* z_out,
                    float* gradE, void* workspace, int** hazard_out, cudaStream_t st) {
  const int K = fast_pick_k(max_target_len);
  CtcFastArgs a{};
  a.E = E; a.targets = targets; a.offsets = offsets; a.B = B; a.T = T; a.C = C; a.blank = blank;
  a.grad_scale = grad_scale; a.z_out = z_out; a.gradE = gradE;
  a.nseg = (T + kSeg - 1) / kSeg;
  a.nA = a.nseg / 2;
  fast_dims(K, C, a.Cp, a.RS);
  a.ckpt = (float*)workspace;
  a.hazard = (int*)((char*)workspace +
                    align_up((size_t)B * 2 * (a.nseg + 1) * (K + 1) * 32 * sizeof(float), 256));
  *hazard_out = a.hazard;
  WFST_CUDA_CHECK(cudaMemsetAsync(a.hazard, 0, (size_t)B * sizeof(int), st));
  switch (K) {
    case 4: return launch_fast_k<4>(a, st);
    case 8: return launch_fast_k<8>(a, st);
    case 12: return launch_fast_k<12>(a, st);
    case 16: return launch_fast_k<16>(a, st);
    case 20: return launch_fast_k<20>(a, st);
    case 24: return launch_fast_k<24>(a, st);
  }
  set_error("no fast CTC instantiation for target length %d", max_target_len);
  return WFST_ERR_UNSUPPORTED;
}

}  // namespace wfst
