// Fast CTC forward+backward for sm_100a: scaled-probability recursion with the whole CTC
// chain of an utterance resident in the registers of one warp per direction, and the
// remaining stages of the computation pipelined over helper warps.  Replaces, for CTC,
// the reference's per-utterance
//   create_ctc_graph -> intersect -> forward_score -> backward
// (criterions/ctc.py:15-29,40-51,78-81).
//
// One thread block (= one utterance) has 7 warps:
//   L0 "live alpha"  time ascending,  states j = s
//   L1 "live beta"   time descending, states j = Sp-1-s (mirrored)
//   P  producer      TMA-loads [8, C] emission tiles and turns them into
//                    p[t,c] = exp(E[t,c] - max_c E[t,c]) tiles for both directions
//   R0, R1 recompute the OPPOSITE recursion of L0 / L1 over a segment from a checkpoint
//   X0, X1 reduce    the per-state posteriors of a segment over states with equal label
//                    and send the [8, C] gradient tile to HBM (bulk async store)
// Both directions run the SAME recursion (the beta recursion written for
// beta~_t(s) = p_t(lab s) * beta_t(s) is the alpha recursion on the reversed target and
// reversed time):  v'[j] = (v[j] + v[j-1] + skip[j] * v[j-2]) * p_t[lab j].
// Lane l owns K consecutive states; neighbours come from one or two warp shuffles per
// frame.  Values are float32 mantissas with one power-of-two exponent per lane,
// re-normalised every 16 frames ("event").
//
// Schedule (meet in the middle + recompute; nothing of size T x S ever leaves the SM):
//   phase 1: L0 sweeps segments [0, nA), L1 sweeps [nA, nseg) downwards; each writes a
//            checkpoint (K values + exponent per lane) per 8-frame segment.
//   meeting: Z = sum_s alpha(s) * beta(s) at the boundary.
//   phase 2: L0 continues upwards through [nA, nseg).  For each segment R0 has re-run the
//            beta recursion from L1's checkpoint into a shared-memory ring (it runs up to
//            two segments ahead); L0 advances alpha frame by frame and multiplies,
//            posterior(s) * Zm = abar(s) * stored(s) * 2^(eL + eR - eZ); X0 then sums the
//            posteriors per label (lane = 4-column chunk of a label-sorted row, segmented
//            suffix sum by shuffles) and stores the gradient tile.  L1 / R1 / X1 mirror
//            this downwards through [0, nA).
//
// Robustness: a frame's posteriors sum to one.  Every segment's total is checked against
// rows * Z (2e-5, finite); a violation (possible only if float32 range was exceeded inside
// a window, which can only lose mass), Z out of range, a blank label inside the target or a
// label layout that does not fit flag the utterance in `hazard`, and the log-semiring
// kernel (lattice.cuh, CtcTopo) recomputes it on the GPU.  No CPU fallback.
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

constexpr int kSeg = 8;               // frames per segment / tile
constexpr int kEventEvery = 2;        // lanes are renormalised every kEventEvery segments (16 frames)
constexpr int kUndef = -(1 << 19);    // "no exponent": lane holds only zeros
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNB = 4;                // p-tile ring depth per direction (the recompute warp runs ahead)
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
  float* ckpt;      // [B][2][nseg + 1][32][ck_stride(K)]
  int* hazard;      // [B]
  int nseg, nA;
  int Cp;           // p-tile row stride (odd, > C); column C is always 0
  int RS;           // row stride of the recomputed ("stored") rows: Sp + 4
  int GW;           // row stride of the sorted posterior ("gamma") rows: 4 * slots + 4
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

// checkpoint I/O: per lane a 16-byte aligned record of kCk(K) floats (K values, then the
// exponent), moved with 128-bit accesses
template <int K>
__host__ __device__ constexpr int ck_stride() { return (K + 1 + 3) & ~3; }
template <int K>
__device__ __forceinline__ void ckpt_store(float* base, const float (&v)[K], int e, int lane) {
  float4* p = reinterpret_cast<float4*>(base + (size_t)lane * ck_stride<K>());
#pragma unroll
  for (int i = 0; i < K; i += 4) p[i >> 2] = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  p[K >> 2] = make_float4(__int_as_float(e), 0.f, 0.f, 0.f);
}
template <int K>
__device__ __forceinline__ void ckpt_load(const float* base, float (&v)[K], int& e, int lane) {
  const float4* p = reinterpret_cast<const float4*>(base + (size_t)lane * ck_stride<K>());
#pragma unroll
  for (int i = 0; i < K; i += 4) {
    const float4 q = p[i >> 2];
    v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
  }
  e = __float_as_int(p[K >> 2].x);
}

// factor that converts the left neighbour's scale to this lane's (0 when either is undefined)
__device__ __forceinline__ float neighbor_factor(int e, int lane) {
  const int el = __shfl_up_sync(kFull, e, 1);
  if (lane == 0 || !defined_exp(el) || !defined_exp(e)) return 0.f;
  const int dd = el - e;
  return (dd < -126) ? 0.f : pow2i(min(dd, 126));
}

// ---------------------------------------------------------------------------
// shared memory
// ---------------------------------------------------------------------------
// mbarrier indices
constexpr int kBarPFull = 0;                   // [2][kNB]  p tile ready (producer -> L, R)
constexpr int kBarPEmpty = kBarPFull + 2 * kNB;  // [2][kNB]  p tile released (count 2: L and R)
constexpr int kBarTma = kBarPEmpty + 2 * kNB;  // [kNR]     raw tile landed
constexpr int kBarSFull = kBarTma + kNR;       // [2][2]    stored segment ready (R -> L)
constexpr int kBarSEmpty = kBarSFull + 4;      // [2][2]    stored segment consumed (L -> R)
constexpr int kBarGFull = kBarSEmpty + 4;      // [2][2]    gamma segment ready (L -> X)
constexpr int kBarGEmpty = kBarGFull + 4;      // [2][2]    gamma segment consumed (X -> L)
constexpr int kBarZ = kBarGEmpty + 4;          // Z published (L1 -> P, R, X); count 1
constexpr int kNumBars = kBarZ + 1;

struct FastSmem {
  // per direction d (two explicit members: indexing an array of pointers with a runtime
  // direction would push the struct into local memory)
  uint32_t stored[2];   // [2 buf][kSeg][RS]  recomputed opposite-direction values
  uint32_t sexp[2];     // [2 buf][32]        their lane exponents (int)
  uint32_t gam[2];      // [2 buf][kSeg][GW]  label-sorted per-state posteriors (pads stay 0)
  uint32_t pbk[2];      // [2 buf][kSeg][33]  blank partial sums per lane
  uint32_t out[2];      // [2 buf][kSeg*C]    gradient tiles
  uint32_t ptile[2];    // [kNB][kSeg][Cp]    p tiles
  uint32_t raw;         // [kNR][kSeg*C]      TMA staging
  uint32_t bars;        // [kNumBars] mbarriers
  int* gcolpos;         // [Sp/2] gamma column of target position n
  int* slotlab;         // [32*kMaxPass + 8] label of each 4-column chunk slot (-1: unused)
  int* hist;            // [C + 4] counting-sort scratch; [C]: #slots, [C+1]: max chunks per label
  float* zx;            // Zm, eZ (int bits), valid flag
  float* out_ptr[2];    // generic pointers of the out tiles (bulk store source)
};

__host__ __device__ inline size_t fast_smem_floats(int K, int C, int Cp, int RS, int GW) {
  const size_t outsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
  size_t per_dir = 2 * (size_t)kSeg * RS + 64 + 2 * (size_t)kSeg * GW + 2 * kSeg * 33 + 2 * outsz +
                   (size_t)kNB * kSeg * Cp;
  per_dir = (per_dir + 3) & ~(size_t)3;
  size_t shared = kNR * outsz + 2 * kNumBars + 16 * K + (32 * kMaxPass + 8) + (C + 4) + 8;
  return 2 * per_dir + shared + 16;
}

__device__ __forceinline__ FastSmem carve_fast(float* base, int K, int C, int Cp, int RS, int GW) {
  FastSmem s;
  float* p = base;
  const size_t outsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
  s.raw = smem_u32(p); p += kNR * outsz;                    // 16B aligned (base is)
  for (int d = 0; d < 2; ++d) { s.out[d] = smem_u32(p); s.out_ptr[d] = p; p += 2 * outsz; }
  for (int d = 0; d < 2; ++d) { s.stored[d] = smem_u32(p); p += 2 * (size_t)kSeg * RS; }
  for (int d = 0; d < 2; ++d) { s.gam[d] = smem_u32(p); p += 2 * (size_t)kSeg * GW; }
  s.bars = smem_u32(p); p += 2 * kNumBars;
  s.zx = p; p += 4;
  for (int d = 0; d < 2; ++d) { s.sexp[d] = smem_u32(p); p += 64; }
  for (int d = 0; d < 2; ++d) { s.pbk[d] = smem_u32(p); p += 2 * kSeg * 33; }
  for (int d = 0; d < 2; ++d) { s.ptile[d] = smem_u32(p); p += (size_t)kNB * kSeg * Cp; }
  s.gcolpos = reinterpret_cast<int*>(p); p += 16 * K;
  s.slotlab = reinterpret_cast<int*>(p); p += 32 * kMaxPass + 8;
  s.hist = reinterpret_cast<int*>(p); p += C + 4;
  return s;
}

// mbarrier helpers on 32-bit shared addresses
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
      "WFST_BW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WFST_BD_%=;\n"
      "bra WFST_BW_%=;\n"
      "WFST_BD_%=:\n"
      "}\n" ::"r"(bars + 8u * idx), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: the
  // warp sleeps in hardware until the phase completes instead of spinning on issue slots
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// common per-block constants
struct Ctx {
  int b, lane, T, C, Cp, RS, GW, L, nseg, nA;
  const int* y;
  bool want_grad;
};

__device__ __forceinline__ int seg_of(int dir, int k, int nseg) { return dir == 0 ? k : nseg - 1 - k; }

// ---------------------------------------------------------------------------
// P: producer.  Direction 0 consumes tiles 0,1,...; direction 1 consumes nseg-1, nseg-2, ...
// Without a gradient only the phase-1 tiles are needed.  The schedule is the sequence
// (k, d), k = 0.., d = 0, 1, restricted to k < ntile[d].  Raw [8, C] tiles are fetched
// kNR - 1 entries ahead (TMA latency >> conversion time).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void role_producer(const CtcFastArgs& a, const FastSmem& sm, const Ctx& cx) {
  const int lane = cx.lane, T = cx.T, C = cx.C, Cp = cx.Cp, nseg = cx.nseg, nA = cx.nA;
  const float* Eb = a.E + (size_t)cx.b * T * C;
  const int ntile0 = cx.want_grad ? nseg : nA, ntile1 = cx.want_grad ? nseg : (nseg - nA);
  const int total = max(ntile0, ntile1);
  const int fr = lane & 7, part = lane >> 3;                       // 8 frames x 4 label quarters
  const int c0 = (C * part) / 4, c1 = (C * (part + 1)) / 4;
  const int rawsz = (kSeg * C + 3) & ~3;
  double msum = 0.0;
  uint32_t tma_phase = 0u, tma_used = 0u, empty_phase = 0u;
  auto next_entry = [&](int& k, int& d) {
    do {
      if (d == 0) d = 1; else { d = 0; ++k; }
    } while (k < total && k >= (d == 0 ? ntile0 : ntile1));
  };
  auto issue_raw = [&](int tile, int slot) {
    const int rows = min(kSeg, T - tile * kSeg);
    const float* src = Eb + (size_t)tile * kSeg * C;
    const uint32_t bytes = (uint32_t)rows * C * 4u;
    const uint32_t dst = sm.raw + 4u * (uint32_t)(slot * rawsz);
    const bool tma = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15u) == 0);
    if (tma) {
      if (lane == 0) {
        bar_expect_tx(sm.bars, kBarTma + slot, bytes);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
            "l"(src), "r"(bytes), "r"(sm.bars + 8u * (kBarTma + slot))
            : "memory");
      }
      tma_used |= 1u << slot;
    } else {
      for (int q = lane; q < rows * C; q += 32) sts(dst + 4u * q, __ldg(src + q));
      tma_used &= ~(1u << slot);
      __syncwarp();
    }
  };
  int k = -1, d = 1;
  next_entry(k, d);
  int fk = k, fd = d, fetched = 0, converted = 0;
  PROF_DECL;
  while (k < total) {
    PROF_MARK(3);
    while (fk < total && fetched < converted + kNR) {
      issue_raw(seg_of(fd, fk, nseg), fetched % kNR);
      ++fetched;
      next_entry(fk, fd);
    }
    const int slot = converted % kNR;
    const int tile = seg_of(d, k, nseg);
    const int rows = min(kSeg, T - tile * kSeg);
    const int buf = k % kNB;
    PROF_MARK(0);
    if (k >= kNB) {  // wait until the consumers have released this p-tile buffer
      const int bit = d * kNB + buf;
      bar_wait(sm.bars, kBarPEmpty + bit, (empty_phase >> bit) & 1u);
      empty_phase ^= 1u << bit;
    }
    PROF_MARK(1);
    if ((tma_used >> slot) & 1u) {
      bar_wait(sm.bars, kBarTma + slot, (tma_phase >> slot) & 1u);
      tma_phase ^= 1u << slot;
    }
    PROF_MARK(2);
    const uint32_t er = sm.raw + 4u * (uint32_t)(slot * rawsz + fr * C);
    const uint32_t pt = sm.ptile[d] + 4u * (uint32_t)(buf * kSeg * Cp + fr * Cp);
    // register-blocked in chunks of 8 labels per lane: all loads of a chunk are issued
    // before any dependent instruction (the shared-memory accessors are volatile asm)
    const bool live = fr < rows;
    float mx = kNegInf;
    for (int cb = c0; cb < c1; cb += 8) {
      float ev[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ev[i] = lds(er + 4u * (uint32_t)min(cb + i, c1 - 1));
#pragma unroll
      for (int i = 0; i < 8; ++i) mx = fmaxf(mx, ev[i]);
    }
    if (!live) mx = kNegInf;
    mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, 16));
    // a row that is entirely -inf keeps p = 0 (dead frame); +inf / NaN rows surface through the
    // certificate
    const float base = (mx == kNegInf) ? 0.f : mx;
    const float nb = -base * 1.4426950408889634f;
    for (int cb = c0; cb < c1; cb += 8) {
      float ev[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ev[i] = lds(er + 4u * (uint32_t)min(cb + i, c1 - 1));
#pragma unroll
      for (int i = 0; i < 8; ++i) ev[i] = exp2f(fmaf(ev[i], 1.4426950408889634f, nb));
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (live && cb + i < c1) sts(pt + 4u * (uint32_t)(cb + i), ev[i]);
    }
    {
      const bool phase1 = (d == 0) ? (tile < nA) : (tile >= nA);
      if (live && part == 0 && phase1) msum += (double)base;
    }
    __syncwarp();
    if (lane == 0) bar_arrive(sm.bars, kBarPFull + d * kNB + buf);
    ++converted;
    next_entry(k, d);
  }
  PROF_MARK(3);
#ifdef WFST_PROFILE
  if (cx.b == 0 && lane == 0) printf("P cycles: issue %lld  wait_pempty %lld  wait_tma %lld  convert %lld\n", pf_acc[0], pf_acc[1], pf_acc[2], pf_acc[3]);
#endif
  // loss: log Z = log(Zm) + eZ ln2 + sum_t max_t
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(kFull, msum, o);
  bar_wait(sm.bars, kBarZ, 0u);
  if (lane == 0) {
    const float Zm = sm.zx[0];
    const int eZ = __float_as_int(sm.zx[1]);
    const bool ok = sm.zx[2] != 0.f;
    a.z_out[cx.b] = ok ? (float)(log((double)Zm) + (double)eZ * 0.6931471805599453 + msum) : kNegInf;
  }
}

// p-tile ring of one direction, as seen by a consumer warp
struct PTileRing {
  uint32_t bars, base, row_bytes;
  int dir;
  uint32_t phase;   // bit per buffer
  __device__ __forceinline__ uint32_t wait(int k) {
    const int buf = k % kNB;
    bar_wait(bars, kBarPFull + dir * kNB + buf, (phase >> buf) & 1u);
    phase ^= 1u << buf;
    return base + (uint32_t)buf * kSeg * row_bytes;
  }
  // consumers that skipped tiles [from, to) must still flip their phase bits for them
  __device__ __forceinline__ void skip(int from, int to) {
    for (int k = from; k < to; ++k) phase ^= 1u << (k % kNB);
  }
  __device__ __forceinline__ void release(int k, int lane, uint32_t count) {
    __syncwarp();
    if (lane == 0) bar_arrive(bars, kBarPEmpty + dir * kNB + (k % kNB), count);
  }
};

// ---------------------------------------------------------------------------
// R: recompute.  R<DIR> serves L<DIR>: it runs the OPPOSITE orientation over the segments
// of L<DIR>'s phase 2, from the checkpoints the other live warp wrote in phase 1, in its
// natural scale, into a two-segment ring; the lane exponents go along.
// ---------------------------------------------------------------------------
template <int K, int DIR>
__device__ __forceinline__ void role_recompute(const CtcFastArgs& a, const FastSmem& sm, const Ctx& cx) {
  constexpr bool ODD = (DIR == 1);   // orientation of the recompute = 1 - DIR; ODD <=> orientation 0
  const int lane = cx.lane, nseg = cx.nseg, nA = cx.nA, T = cx.T;
  bar_wait(sm.bars, kBarZ, 0u);      // phase 1 (and every checkpoint) is complete
  if (!cx.want_grad || sm.zx[2] == 0.f) return;
  const int n1 = (DIR == 0) ? nA : nseg - nA;
  const int n2 = nseg - n1;
  LaneTopo<K> tp;
  build_topo<K, ODD>(tp, lane, cx.y, cx.L, cx.C, sm.gcolpos, 0);
  const float* ck = a.ckpt + ((size_t)cx.b * 2 + (1 - DIR)) * (size_t)(nseg + 1) * ck_stride<K>() * 32;
  PTileRing ring{sm.bars, sm.ptile[DIR], 4u * (uint32_t)cx.Cp, DIR, 0u};
  ring.skip(0, n1);
  const uint32_t blank_ofs = 4u * (uint32_t)a.blank;
  const uint32_t srow = 4u * (uint32_t)cx.RS;
  const uint32_t my_block = 4u * (uint32_t)(lane * K);
  uint32_t sempty_phase = 0u;
  float w[K], dummy[K];
  int ew;
  if (n2 > 0) ckpt_load<K>(ck + (size_t)seg_of(DIR, n1, nseg) * ck_stride<K>() * 32, w, ew, lane);
  PROF_DECL;
  for (int k2 = 0; k2 < n2; ++k2) {
    const int k = n1 + k2;
    const int seg = seg_of(DIR, k, nseg);
    const int rows = min(kSeg, T - seg * kSeg);
    const int buf = k2 & 1;
    PROF_MARK(2);
    if (k2 >= 2) {
      bar_wait(sm.bars, kBarSEmpty + DIR * 2 + buf, (sempty_phase >> buf) & 1u);
      sempty_phase ^= 1u << buf;
    }
    PROF_MARK(0);
    const uint32_t pt = ring.wait(k);
    PROF_MARK(1);
    const float fr = neighbor_factor(ew, lane);
    const uint32_t sbase = sm.stored[DIR] + (uint32_t)buf * kSeg * srow + my_block;
    sts(sm.sexp[DIR] + 4u * (uint32_t)(buf * 32 + lane), __int_as_float(ew));
    // the recompute walks the frames in the opposite order to the live sweep
    if (DIR == 0) {
      uint32_t pr = pt + (uint32_t)(rows - 1) * ring.row_bytes;
      uint32_t dst = sbase + (uint32_t)(rows - 1) * srow;
      PRow<K> nx = load_prow<K>(tp, pr, blank_ofs);
#pragma unroll 1
      for (int r = rows - 1; r >= 0; --r, dst -= srow) {
        const PRow<K> cur = nx;
        pr -= (r > 0) ? ring.row_bytes : 0u;
        nx = load_prow<K>(tp, pr, blank_ofs);
        step<K, ODD, false>(w, dummy, tp, cur, fr);
#pragma unroll
        for (int i = 0; i < K; i += 4) sts128(dst + 4u * i, w[i], w[i + 1], w[i + 2], w[i + 3]);
      }
    } else {
      uint32_t pr = pt;
      uint32_t dst = sbase;
      PRow<K> nx = load_prow<K>(tp, pr, blank_ofs);
#pragma unroll 1
      for (int r = 0; r < rows; ++r, dst += srow) {
        const PRow<K> cur = nx;
        pr += (r + 1 < rows) ? ring.row_bytes : 0u;
        nx = load_prow<K>(tp, pr, blank_ofs);
        step<K, ODD, false>(w, dummy, tp, cur, fr);
#pragma unroll
        for (int i = 0; i < K; i += 4) sts128(dst + 4u * i, w[i], w[i + 1], w[i + 2], w[i + 3]);
      }
    }
    // next checkpoint (consumed at the top of the next iteration)
    if (k2 + 1 < n2) ckpt_load<K>(ck + (size_t)seg_of(DIR, k + 1, nseg) * ck_stride<K>() * 32, w, ew, lane);
    __syncwarp();
    if (lane == 0) bar_arrive(sm.bars, kBarSFull + DIR * 2 + buf);
    ring.release(k, lane, 1);
  }
  PROF_MARK(2);
#ifdef WFST_PROFILE
  if (cx.b == 0 && lane == 0)
    printf("R%d cycles: wait_sempty %lld  wait_ptile %lld  compute %lld\n", DIR, pf_acc[0], pf_acc[1], pf_acc[2]);
#endif
}

// ---------------------------------------------------------------------------
// X: per-label reduction of a segment's posteriors + gradient tile store.  The direction
// is a runtime argument and the pass loop is not unrolled: X0 and X1 share one small
// body (the warps of a block run seven different loops at once; keeping each of them
// small matters for the instruction cache).
// ---------------------------------------------------------------------------
__device__ __noinline__ void role_reduce(const CtcFastArgs& a, const FastSmem& sm, const Ctx& cx, int dir) {
  const int lane = cx.lane, nseg = cx.nseg, nA = cx.nA, T = cx.T, C = cx.C;
  bar_wait(sm.bars, kBarZ, 0u);
  if (!cx.want_grad || sm.zx[2] == 0.f) return;
  const float Zm = sm.zx[0];
  const float gs = a.grad_scale ? a.grad_scale[cx.b] : 1.f;
  const float kappa = -gs / Zm;
  const int n1 = (dir == 0) ? nA : nseg - nA;
  const int n2 = nseg - n1;
  float* gEb = a.gradE + (size_t)cx.b * T * C;
  const int outsz = (kSeg * C + 3) & ~3;
  const uint32_t grow = 4u * (uint32_t)cx.GW;
  const uint32_t blank_ofs = 4u * (uint32_t)a.blank;
  const uint32_t gam0 = dir ? sm.gam[1] : sm.gam[0];
  const uint32_t pbk0 = dir ? sm.pbk[1] : sm.pbk[0];
  const uint32_t out0 = dir ? sm.out[1] : sm.out[0];
  const float* outp = dir ? sm.out_ptr[1] : sm.out_ptr[0];
  // labels: lane = chunk slot (32 slots per pass); slots of one label are adjacent lanes
  const int nslots = sm.hist[C];
  const int npass = (nslots + 31) >> 5;
  const int maxch = sm.hist[C + 1];
  // blank: lane = (frame, quarter) sums 8 of the 32 lane partials
  const int frm = lane & 7, qtr = lane >> 3;
  int bad = 0;
  uint32_t gfull_phase = 0u;
  PROF_DECL;
  for (int k2 = 0; k2 < n2; ++k2) {
    const int seg = seg_of(dir, n1 + k2, nseg);
    const int rows = min(kSeg, T - seg * kSeg);
    const int buf = k2 & 1;
    PROF_MARK(1);
    bar_wait(sm.bars, kBarGFull + dir * 2 + buf, (gfull_phase >> buf) & 1u);
    gfull_phase ^= 1u << buf;
    PROF_MARK(0);
    const uint32_t ot = out0 + 4u * (uint32_t)(buf * outsz);
    if (lane == 0) bulk_wait_read<1>();   // the store that last read this out buffer is done
    __syncwarp();
    const uint32_t gbase = gam0 + (uint32_t)buf * kSeg * grow + 16u * (uint32_t)lane;
    float tot = 0.f;
#pragma unroll 1
    for (int p = 0; p < npass; ++p) {
      const int g = 32 * p + lane;
      const int me = sm.slotlab[g];
      // links never leave the pass: a label's slots do not straddle a multiple of 32
      const float lk1 = (me >= 0 && lane + 1 < 32 && sm.slotlab[g + 1] == me) ? 1.f : 0.f;
      const float lk2 = (me >= 0 && lane + 2 < 32 && sm.slotlab[g + 2] == me) ? 1.f : 0.f;
      const float lk4 = (me >= 0 && lane + 4 < 32 && sm.slotlab[g + 4] == me) ? 1.f : 0.f;
      const bool head = me >= 0 && (lane == 0 || sm.slotlab[g - 1] != me);
      const uint32_t src = gbase + 512u * (uint32_t)p;
      const uint32_t dsto = ot + 4u * (uint32_t)max(me, 0);
      float c[kSeg];
#pragma unroll
      for (int j = 0; j < kSeg; ++j) {   // rows >= `rows` hold finite stale data; never stored
        const float4 q = lds128(src + (uint32_t)j * grow);
        c[j] = (q.x + q.y) + (q.z + q.w);
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
#pragma unroll
        for (int j = 0; j < kSeg; ++j) {
          if (j < rows) {
            sts(dsto + 4u * (uint32_t)(j * C), c[j] * kappa);
            tot += c[j];
          }
        }
      }
    }
    float bs = 0.f;
    if (frm < rows) {
      const uint32_t pb = pbk0 + 4u * (uint32_t)(buf * kSeg * 33 + frm * 33 + qtr * 8);
#pragma unroll
      for (int q = 0; q < 8; ++q) bs += lds(pb + 4u * q);
    }
    tot += bs;
    float bsum = bs + __shfl_xor_sync(kFull, bs, 8);
    bsum += __shfl_xor_sync(kFull, bsum, 16);
    if (frm < rows && qtr == 0) sts(ot + 4u * (uint32_t)(frm * C) + blank_ofs, bsum * kappa);
    // certificate: the posteriors of every frame sum to one, i.e. the segment sums to rows * Zm.
    // Range loss can only remove mass, so deficits cannot cancel.
    tot = warp_sum(tot);
    if (!(fabsf(tot - (float)rows * Zm) <= 2e-5f * (float)rows * Zm)) bad |= 8;
    __syncwarp();
    if (lane == 0) bar_arrive(sm.bars, kBarGEmpty + dir * 2 + buf);   // gamma / pbk are consumed
    float* dst = gEb + (size_t)seg * kSeg * C;
    const int n = rows * C;
    const bool tma = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n & 3) == 0);
    if (tma) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(ot),
                     "r"((uint32_t)n * 4u)
                     : "memory");
        bulk_commit();
      }
    } else {
      const float* src = outp + (size_t)buf * outsz;
      for (int q = lane; q < n; q += 32) dst[q] = src[q];
    }
    __syncwarp();
  }
  PROF_MARK(1);
#ifdef WFST_PROFILE
  if (cx.b == 0 && lane == 0) printf("X%d cycles: wait_gfull %lld  compute %lld\n", dir, pf_acc[0], pf_acc[1]);
#endif
  if (lane == 0) bulk_wait_all<0>();
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0) atomicOr(&a.hazard[cx.b], bad);
}

// ---------------------------------------------------------------------------
// L: the live sweep of one direction (whole warp).  DIR 0: alpha, time ascending, states
// j = s.  DIR 1: beta, time descending, mirrored states j = Sp-1-s.
// ---------------------------------------------------------------------------
template <int K, int DIR>
__device__ __forceinline__ void role_live(const CtcFastArgs& a, const FastSmem& sm, const Ctx& cx) {
  constexpr int Sp = 32 * K;
  constexpr bool ODD = (DIR == 0);
  const int lane = cx.lane, nseg = cx.nseg, nA = cx.nA, T = cx.T, b = cx.b;
  float* ck_own = a.ckpt + ((size_t)b * 2 + DIR) * (size_t)(nseg + 1) * ck_stride<K>() * 32;
  const float* ck_other = a.ckpt + ((size_t)b * 2 + (1 - DIR)) * (size_t)(nseg + 1) * ck_stride<K>() * 32;
  LaneTopo<K> tp;
  build_topo<K, ODD>(tp, lane, cx.y, cx.L, cx.C, sm.gcolpos, cx.GW - 1);

  float v[K], abar[K];
  int e = kUndef;
  float f = 0.f;
  {
    // virtual pre-frame state: all mass on the start state of this orientation
    const int S = 2 * cx.L + 1;
    const int jstart = (DIR == 0) ? 0 : Sp - S;
    const bool mine = (jstart / K == lane);
    const int jm = jstart % K;
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = (mine && jm == i) ? 1.f : 0.f;   // selects, no dynamic index
    if (mine) e = 0;
  }
  PTileRing ring{sm.bars, sm.ptile[DIR], 4u * (uint32_t)cx.Cp, DIR, 0u};
  const uint32_t blank_ofs = 4u * (uint32_t)a.blank;

  // ------------------------------------------------------------------ phase 1
  PROF_DECL;
  const int n1 = (DIR == 0) ? nA : nseg - nA;
  for (int k = 0; k < n1; ++k) {
    const int seg = seg_of(DIR, k, nseg);
    const int rows = min(kSeg, T - seg * kSeg);
    if (k % kEventEvery == 0) event<K>(v, e, f, lane);
    ckpt_store<K>(ck_own + (size_t)seg * ck_stride<K>() * 32, v, e, lane);
    PROF_MARK(0);
    const uint32_t pt = ring.wait(k);
    PROF_MARK(1);
    uint32_t pr = (DIR == 0) ? pt : pt + (uint32_t)(rows - 1) * ring.row_bytes;
    PRow<K> nx = load_prow<K>(tp, pr, blank_ofs);
#pragma unroll 2
    for (int it = 0; it < rows; ++it) {
      const PRow<K> cur = nx;
      if (it + 1 < rows) pr = (DIR == 0) ? pr + ring.row_bytes : pr - ring.row_bytes;
      nx = load_prow<K>(tp, pr, blank_ofs);
      step<K, ODD, false>(v, abar, tp, cur, f);
    }
    ring.release(k, lane, 2);   // no recompute warp reads phase-1 tiles
    PROF_MARK(2);
  }

  // ------------------------------------------------------------------ meeting: Z
  // L0 publishes its state (extra checkpoint slot nseg of its own area); L1 combines.
  if (DIR == 0) ckpt_store<K>(ck_own + (size_t)nseg * ck_stride<K>() * 32, v, e, lane);
  named_sync(1, 64);
  if (DIR == 1) {
    event<K>(v, e, f, lane);     // consistent exponents / f for the shuffle below
    // pre-emission sums of L1's next frame: bb = v[j] + v[j-1] + skip * v[j-2]
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
          s = fmaf(tp.skipm[i >> 1], a2, s);
        }
        bb[i] = s;
      }
    }
    float av[K];
    int ea;
    ckpt_load<K>(ck_other + (size_t)nseg * ck_stride<K>() * 32, av, ea, 31 - lane);
    float P = 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) P = fmaf(bb[i], av[K - 1 - i], P);
    int Eabs = kUndef;
    if (P > 0.f && defined_exp(e) && defined_exp(ea)) Eabs = e + ea;
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
  if (DIR == 1 && lane == 0) bar_arrive(sm.bars, kBarZ);
  const int eZ = __float_as_int(sm.zx[1]);
  const bool zok = sm.zx[2] != 0.f;
  if (!cx.want_grad) return;

  // ------------------------------------------------------------------ phase 2
  const int n2 = nseg - n1;
  if (!zok) {
    // keep the producer's ring moving so that it can terminate (the R warp has left too)
    for (int k2 = 0; k2 < n2; ++k2) { ring.wait(n1 + k2); ring.release(n1 + k2, lane, 2); }
    return;
  }
  const uint32_t srow = 4u * (uint32_t)cx.RS;
  const uint32_t grow = 4u * (uint32_t)cx.GW;
  const uint32_t pair_block = 4u * (uint32_t)((31 - lane) * K);   // the block of the lane paired with me
  int bad = 0;   // reason bit 4: scale overflow when pairing live and recomputed values
  uint32_t sfull_phase = 0u, gempty_phase = 0u;

  for (int k2 = 0; k2 < n2; ++k2) {
    const int k = n1 + k2;
    const int seg = seg_of(DIR, k, nseg);
    const int rows = min(kSeg, T - seg * kSeg);
    const int buf = k2 & 1;
    PROF_MARK(3);
    if (k % kEventEvery == 0) event<K>(v, e, f, lane);
    PROF_MARK(0);
    const uint32_t pt = ring.wait(k);
    PROF_MARK(1);
    bar_wait(sm.bars, kBarSFull + DIR * 2 + buf, (sfull_phase >> buf) & 1u);
    sfull_phase ^= 1u << buf;
    PROF_MARK(4);
    // posterior * Zm = abar * stored * 2^(eL + eR - eZ): one factor per lane pair and segment
    float g = 0.f;
    {
      const int er = __float_as_int(lds(sm.sexp[DIR] + 4u * (uint32_t)(buf * 32 + 31 - lane)));
      if (defined_exp(e) && defined_exp(er)) {
        const int dd = e + er - eZ;
        if (dd > 126) bad |= 4;
        else g = (dd < -126) ? 0.f : pow2i(dd);
      }
    }
    if (k2 >= 2) {   // X has consumed the gamma / pbk buffer used two segments ago
      bar_wait(sm.bars, kBarGEmpty + DIR * 2 + buf, (gempty_phase >> buf) & 1u);
      gempty_phase ^= 1u << buf;
    }
    PROF_MARK(5);
    const uint32_t sbase = sm.stored[DIR] + (uint32_t)buf * kSeg * srow + pair_block;
    const uint32_t gbase = sm.gam[DIR] + (uint32_t)buf * kSeg * grow;
    const uint32_t pbase = sm.pbk[DIR] + 4u * (uint32_t)(buf * kSeg * 33 + lane);
    {
      int r = (DIR == 0) ? 0 : rows - 1;
      uint32_t pr = pt + (uint32_t)r * ring.row_bytes;
      float stn[K];
      auto fetch_block = [&](int rr) {
#pragma unroll
        for (int i = 0; i < K; i += 4) {
          const float4 q = lds128(sbase + (uint32_t)rr * srow + 4u * i);
          stn[i] = q.x; stn[i + 1] = q.y; stn[i + 2] = q.z; stn[i + 3] = q.w;
        }
      };
      fetch_block(r);
      PRow<K> nx = load_prow<K>(tp, pr, blank_ofs);
#pragma unroll 2
      for (int it = 0; it < rows; ++it) {
        float st[K];
#pragma unroll
        for (int i = 0; i < K; ++i) st[i] = stn[i] * g;
        const PRow<K> cur = nx;
        const int rc = r;
        if (it + 1 < rows) {
          r = (DIR == 0) ? r + 1 : r - 1;
          pr = (DIR == 0) ? pr + ring.row_bytes : pr - ring.row_bytes;
          fetch_block(r);
        }
        nx = load_prow<K>(tp, pr, blank_ofs);
        step<K, ODD, true>(v, abar, tp, cur, f);
        float pbsum = 0.f;
        const uint32_t gr = gbase + (uint32_t)rc * grow;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const bool lab = ((i & 1) == 1) == ODD;
          const float gm = abar[i] * st[K - 1 - i];
          if (lab) sts(gr + tp.gofs[i >> 1], gm);
          else pbsum += gm;
        }
        sts(pbase + 4u * (uint32_t)(rc * 33), pbsum);
      }
    }
    __syncwarp();
    if (lane == 0) {
      bar_arrive(sm.bars, kBarSEmpty + DIR * 2 + buf);   // R may refill this stored buffer
      bar_arrive(sm.bars, kBarGFull + DIR * 2 + buf);    // X may reduce this gamma buffer
    }
    ring.release(k, lane, 1);
    PROF_MARK(6);
  }
#ifdef WFST_PROFILE
  if (b == 0 && lane == 0)
    printf("L%d cycles: event+ckpt %lld  wait_ptile %lld  phase1-steps %lld  misc %lld  wait_sfull %lld  wait_gempty %lld  combine %lld\n",
           DIR, pf_acc[0], pf_acc[1], pf_acc[2], pf_acc[3], pf_acc[4], pf_acc[5], pf_acc[6]);
#endif
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0) atomicOr(&a.hazard[b], bad);
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(224, (K <= 12) ? 2 : 1) ctc_fast_kernel(CtcFastArgs a) {
  extern __shared__ __align__(16) float smem_raw[];
  const int warp = threadIdx.x >> 5;
  Ctx cx;
  cx.b = blockIdx.x;
  cx.lane = threadIdx.x & 31;
  cx.T = a.T; cx.C = a.C; cx.Cp = a.Cp; cx.RS = a.RS; cx.GW = a.GW;
  cx.nseg = a.nseg; cx.nA = a.nA;
  cx.y = a.targets + a.offsets[cx.b];
  cx.L = a.offsets[cx.b + 1] - a.offsets[cx.b];
  cx.want_grad = a.gradE != nullptr;
  const int C = a.C, L = cx.L;
  const int* y = cx.y;
  FastSmem sm = carve_fast(smem_raw, K, C, a.Cp, a.RS, a.GW);
  const int NT = blockDim.x;

  // ------------------------------------------------------------------ setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumBars; ++i) {
      const bool two = i >= kBarPEmpty && i < kBarPEmpty + 2 * kNB;
      bar_init(sm.bars, i, two ? 2u : 1u);
    }
    fence_barrier_init();
  }
  // zero what relies on it: p-tile padding columns, gradient tiles (labels that do not occur in
  // the target keep a zero gradient), gamma rows (padding columns of the chunk slots), blank partials
  {
    const int outsz = (kSeg * C + 3) & ~3;
    for (int d = 0; d < 2; ++d) {
      for (int k = threadIdx.x; k < kNB * kSeg * a.Cp; k += NT) sts(sm.ptile[d] + 4u * k, 0.f);
      for (int k = threadIdx.x; k < 2 * outsz; k += NT) sts(sm.out[d] + 4u * k, 0.f);
      for (int k = threadIdx.x; k < 2 * kSeg * a.GW; k += NT) sts(sm.gam[d] + 4u * k, 0.f);
      for (int k = threadIdx.x; k < 2 * kSeg * 33; k += NT) sts(sm.pbk[d] + 4u * k, 0.f);
    }
  }
  // Counting sort of the target positions by label -> columns of the "gamma" row.  The row
  // is organised in chunk slots of 4 columns; a label with n occurrences owns ceil(n/4)
  // consecutive slots that never straddle a multiple of 32 slots (one pass of the
  // transposed reduction = 32 slots, one per lane).
  for (int k = threadIdx.x; k < C + 4; k += NT) sm.hist[k] = 0;
  for (int k = threadIdx.x; k < 32 * kMaxPass + 8; k += NT) sm.slotlab[k] = -1;
  __syncthreads();
  int has_blank = 0;
  for (int n = threadIdx.x; n < L; n += NT) {
    const int yy = y[n];
    // a label outside [0, C) (device-resident targets are not validated on the host) must not
    // index the histogram or a p tile: flag the utterance, the log-semiring kernel clamps it
    if (yy < 0 || yy >= C) { has_blank = 1; continue; }
    atomicAdd(&sm.hist[yy], 1);
    has_blank |= (yy == a.blank);
  }
  if (__syncthreads_or(has_blank)) {
    // a target that contains the blank label shares a gradient column between a label state
    // and the blank states: leave it to the log-semiring kernel (so are out-of-range labels)
    if (threadIdx.x == 0) a.hazard[cx.b] = 1;   // reason 1
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
    // layouts the transposed pass cannot hold go to the log-semiring kernel
    if (nslots > 32 * kMaxPass || nslots * 4 > a.GW - 4 || maxch > 8) {
      if (threadIdx.x == 0) a.hazard[cx.b] = 16;  // reason 16
      return;
    }
  }
  for (int n = threadIdx.x; n < L; n += NT) sm.gcolpos[n] = atomicAdd(&sm.hist[y[n]], 1);
  __syncthreads();

  switch (warp) {
    case 0: role_live<K, 0>(a, sm, cx); break;
    case 1: role_live<K, 1>(a, sm, cx); break;
    case 2: role_producer(a, sm, cx); break;
    case 3: role_recompute<K, 0>(a, sm, cx); break;
    case 4: role_recompute<K, 1>(a, sm, cx); break;
    case 5: role_reduce(a, sm, cx, 0); break;
    default: role_reduce(a, sm, cx, 1); break;
  }
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

static void fast_dims(int K, int C, int max_target_len, int& Cp, int& RS, int& GW) {
  Cp = (C + 1) | 1;
  RS = 32 * K + 4;
  // chunk slots: ceil(n_c / 4) per label, plus the slots skipped so that no label straddles
  // a multiple of 32 (fewer than 8 per pass), plus the dump column
  const int labels = (C - 1 < max_target_len) ? C - 1 : max_target_len;
  int slots = (max_target_len + 3 * labels + 3) / 4 + 8;
  if (slots > 32 * kMaxPass) slots = 32 * kMaxPass;
  GW = 4 * slots + 4;
}

bool ctc_fast_eligible(int T, int C, int max_target_len) {
  if (T < 1) return false;
  const int K = fast_pick_k(max_target_len);
  if (K == 0) return false;
  int Cp, RS, GW;
  fast_dims(K, C, max_target_len, Cp, RS, GW);
  return fast_smem_floats(K, C, Cp, RS, GW) * sizeof(float) <= 227 * 1024;
}

size_t ctc_fast_workspace_bytes(int B, int T, int max_target_len) {
  const int K = fast_pick_k(max_target_len);
  const int nseg = (T + kSeg - 1) / kSeg;
  return align_up((size_t)B * 2 * (nseg + 1) * ((K + 4) & ~3) * 32 * sizeof(float), 256) +
         align_up((size_t)B * sizeof(int), 256);
}

template <int K>
static int launch_fast_k(const CtcFastArgs& a, cudaStream_t st) {
  size_t smem = fast_smem_floats(K, a.C, a.Cp, a.RS, a.GW) * sizeof(float);
  auto kern = ctc_fast_kernel<K>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<a.B, 224, smem, st>>>(a);
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
  fast_dims(K, C, max_target_len, a.Cp, a.RS, a.GW);
  a.ckpt = (float*)workspace;
  a.hazard = (int*)((char*)workspace +
                    align_up((size_t)B * 2 * (a.nseg + 1) * ((K + 4) & ~3) * 32 * sizeof(float), 256));
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
