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

constexpr int kSeg = 16;              // frames per segment / tile
constexpr int kUndef = -(1 << 20);    // "no exponent": lane holds only zeros
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNB = 2;                // p-tile ring depth per direction

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

__device__ __forceinline__ bool defined_exp(int e) { return e > kUndef / 2; }
__device__ __forceinline__ float pow2i(int d) {  // 2^d for d in [-126, 127]
  return __uint_as_float((uint32_t)(d + 127) << 23);
}

// ---------------------------------------------------------------------------
// Per-lane static description of its K states in one orientation.
// ---------------------------------------------------------------------------
template <int K>
struct LaneTopo {
  int labcol[K / 2];   // p-tile column of each label-type slot (C = zero column for padding)
  float skipm[K / 2];  // 1 if the skip arc into the label-type slot exists
  int gcol[K / 2];     // column of the slot's posterior in the sorted gamma row
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
    tp.labcol[q] = col;
    tp.skipm[q] = sk;
    tp.gcol[q] = gc;
  }
}

// One frame of the recursion. v: with-emission values of the previous frame (own scale).
// On return v holds this frame's with-emission values and, if WANT_ABAR, abar the
// pre-emission sums.  f converts the left neighbour's scale to ours (0 in lane 0).
template <int K, bool ODD, bool WANT_ABAR>
__device__ __forceinline__ void step(float (&v)[K], float (&abar)[K], const LaneTopo<K>& tp,
                                     const float* __restrict__ prow, int blank, float f) {
  float pl[K / 2];
#pragma unroll
  for (int q = 0; q < K / 2; ++q) pl[q] = prow[tp.labcol[q]];
  const float pb = prow[blank];
  const float in1 = __shfl_up_sync(kFull, v[K - 1], 1) * f;
  float in2 = 0.f;
  if (!ODD) in2 = __shfl_up_sync(kFull, v[K - 2], 1) * f;
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    constexpr bool dummy = false;
    (void)dummy;
    const bool lab = ((i & 1) == 1) == ODD;
    const float a1 = (i >= 1) ? v[i - 1] : in1;
    float s = v[i] + a1;
    if (lab) {
      const int q = i >> 1;
      const float a2 = (i >= 2) ? v[i - 2] : (i == 1 ? in1 : in2);
      s = fmaf(tp.skipm[q], a2, s);
      if (WANT_ABAR) abar[i] = s;
      v[i] = s * pl[q];
    } else {
      if (WANT_ABAR) abar[i] = s;
      v[i] = s * pb;
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
  int c = eown, d = defined_exp(eown) ? D : 0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int pc = __shfl_up_sync(kFull, c, o);
    const int pd = __shfl_up_sync(kFull, d, o);
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
  float* stored[2];   // [kSeg][RS] recomputed opposite-direction values; row r is re-used
                      // for the sorted per-state posteriors ("gamma") once it has been read
  float* pbk[2];      // [kSeg][33] blank partials
  float* out[2];      // [2][kSeg*C] output tiles (double buffered)
  float* ptile[2];    // [kNB][kSeg][Cp]
  // shared
  float* raw;         // [2][kSeg*C] TMA staging (16B aligned)
  int* gcolpos;       // [Sp/2] gamma column of target position n
  int* runs;          // [2][C+1][2] per half: (label, end column); terminated by label -1
  int* hist;          // [C + 2] scratch for the counting sort
  uint64_t* bars;     // full[2][kNB], empty[2][kNB], tma[2], zready
  float* zx;          // Zm, eZ (as int bits), valid flag
  double* msum;       // sum of per-frame maxima (phase-1 frames)
};

__host__ __device__ inline size_t fast_smem_floats(int K, int C, int Cp, int RS) {
  const int Sp = 32 * K;
  size_t per_dir = (size_t)kSeg * RS + kSeg * 33 + 2 * (((size_t)kSeg * C + 3) & ~3) +
                   (size_t)kNB * kSeg * Cp;
  per_dir = (per_dir + 3) & ~(size_t)3;
  size_t shared = 2 * (((size_t)kSeg * C + 3) & ~3) + Sp / 2 + 4 * (C + 1) + (C + 2) + 2 * 16 + 8 + 4;
  return 2 * per_dir + shared + 16;
}

template <int K>
__device__ __forceinline__ FastSmem<K> carve_fast(float* base, int C, int Cp, int RS) {
  constexpr int Sp = 32 * K;
  FastSmem<K> s;
  float* p = base;
  const size_t outsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
  s.raw = p; p += 2 * outsz;                       // 16B aligned (base is)
  for (int d = 0; d < 2; ++d) { s.out[d] = p; p += 2 * outsz; }
  for (int d = 0; d < 2; ++d) { s.stored[d] = p; p += (size_t)kSeg * RS; }
  s.bars = reinterpret_cast<uint64_t*>(p); p += 2 * 16;      // up to 16 barriers
  s.msum = reinterpret_cast<double*>(p); p += 4;
  s.zx = p; p += 4;
  for (int d = 0; d < 2; ++d) { s.pbk[d] = p; p += kSeg * 33; }
  for (int d = 0; d < 2; ++d) { s.ptile[d] = p; p += (size_t)kNB * kSeg * Cp; }
  s.gcolpos = reinterpret_cast<int*>(p); p += Sp / 2;
  s.runs = reinterpret_cast<int*>(p); p += 4 * (C + 1);
  s.hist = reinterpret_cast<int*>(p); p += C + 2;
  return s;
}

// barrier indices
__device__ __forceinline__ int bar_full(int d, int i) { return d * kNB + i; }
__device__ __forceinline__ int bar_empty(int d, int i) { return 2 * kNB + d * kNB + i; }
constexpr int kBarTma = 4 * kNB;      // +0, +1
constexpr int kBarZ = 4 * kNB + 2;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
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
    mbar_init(&sm.bars[kBarTma], 1);
    mbar_init(&sm.bars[kBarTma + 1], 1);
    mbar_init(&sm.bars[kBarZ], 1);
    fence_barrier_init();
  }
  // zero the regions that rely on it: p-tile padding columns, output tiles (labels that
  // do not occur in the target keep a zero gradient), blank partials
  for (int d = 0; d < 2; ++d) {
    for (int k = threadIdx.x; k < kNB * kSeg * Cp; k += 96) sm.ptile[d][k] = 0.f;
    for (int k = threadIdx.x; k < 2 * (int)(((size_t)kSeg * C + 3) & ~(size_t)3); k += 96) sm.out[d][k] = 0.f;
    for (int k = threadIdx.x; k < kSeg * 33; k += 96) sm.pbk[d][k] = 0.f;
    for (int k = threadIdx.x; k < kSeg * RS; k += 96) sm.stored[d][k] = 0.f;
  }
  // counting sort of the target positions by label -> gamma columns; each label's run is
  // padded to a multiple of 4 columns; labels are split into two halves (one per
  // half-warp of the transposed pass), the second half starting at a column = 16 mod 32
  for (int k = threadIdx.x; k < C + 2; k += 96) sm.hist[k] = 0;
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
    // serial over C labels (C is small); runs[h][r] = (label, end column)
    int half_target = (L + 1) / 2, seen = 0, col = 0, h = 0, r = 0;
    int* runs = sm.runs;
    for (int c = 0; c < C; ++c) {
      const int cnt = sm.hist[c];
      sm.hist[c] = col;                 // becomes the write cursor of label c
      if (cnt == 0) continue;
      const int width = (cnt + 3) & ~3;
      runs[(h * (C + 1) + r) * 2 + 0] = c;
      runs[(h * (C + 1) + r) * 2 + 1] = col + width;
      ++r;
      col += width;
      seen += cnt;
      if (h == 0 && seen >= half_target) {
        runs[(0 * (C + 1) + r) * 2 + 0] = -1;
        h = 1; r = 0;
        col = ((col + 15) & ~31) + 16;  // next column = 16 mod 32, >= col
        sm.hist[C] = col;               // first column of the second half
      }
    }
    if (h == 0) { runs[(0 * (C + 1) + r) * 2 + 0] = -1; h = 1; r = 0; sm.hist[C] = col; }
    runs[(1 * (C + 1) + r) * 2 + 0] = -1;
    sm.hist[C + 1] = col;               // first unused column (must be < RS - 1)
  }
  __syncthreads();
  // second-half start column is recomputed by each reader from runs; assign columns
  for (int n = threadIdx.x; n < L; n += 96) sm.gcolpos[n] = atomicAdd(&sm.hist[y[n]], 1);
  __syncthreads();

  // ===================================================================== producer
  if (warp == 2) {
    // Direction 0 consumes tiles 0,1,...; direction 1 consumes nseg-1, nseg-2, ...
    // Without a gradient only the phase-1 tiles are needed.
    const int ntile[2] = {want_grad ? nseg : nA, want_grad ? nseg : (nseg - nA)};
    const int total = max(ntile[0], ntile[1]);
    const int fr = lane & 15, hh = lane >> 4;
    const int c0 = hh ? (C + 1) / 2 : 0, c1 = hh ? C : (C + 1) / 2;
    const size_t rawsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
    double msum = 0.0;
    uint32_t tma_phase[2] = {0u, 0u};
    uint32_t empty_phase[2][kNB] = {{0u, 0u}, {0u, 0u}};
    auto tile_of = [&](int d, int k) { return d == 0 ? k : nseg - 1 - k; };
    // the schedule is the sequence (k, d), k = 0.., d = 0, 1, restricted to k < ntile[d]
    auto next_entry = [&](int& k, int& d) {
      do {
        if (d == 0) d = 1; else { d = 0; ++k; }
      } while (k < total && k >= ntile[d]);
    };
    auto issue_raw = [&](int tile, int slot) -> bool {
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
      } else {
        for (int q = lane; q < rows * C; q += 32) dst[q] = __ldg(src + q);
        __syncwarp();
      }
      return tma;
    };
    int k = 0, d = -1;
    {  // first entry
      int kk = 0, dd = 1; --kk;  // so that next_entry lands on (0, 0) if valid
      kk = -1; dd = 1;
      next_entry(kk, dd);
      k = kk; d = dd;
    }
    int slot = 0;
    bool cur_tma = false;
    if (k < total) cur_tma = issue_raw(tile_of(d, k), slot);
    while (k < total) {
      int nk = k, nd = d;
      next_entry(nk, nd);
      bool next_tma = false;
      if (nk < total) next_tma = issue_raw(tile_of(nd, nk), slot ^ 1);   // prefetch
      const int tile = tile_of(d, k);
      const int rows = min(kSeg, T - tile * kSeg);
      const int buf = k % kNB;
      if (k >= kNB) {  // wait until the consumer has released this p-tile buffer
        mbar_wait(&sm.bars[bar_empty(d, buf)], empty_phase[d][buf]);
        empty_phase[d][buf] ^= 1u;
      }
      if (cur_tma) {
        mbar_wait(&sm.bars[kBarTma + slot], tma_phase[slot]);
        tma_phase[slot] ^= 1u;
      }
      const float* er = sm.raw + (size_t)slot * rawsz + fr * C;
      float* pt = sm.ptile[d] + (size_t)buf * kSeg * Cp + fr * Cp;
      float mx = kNegInf;
      if (fr < rows)
        for (int c = c0; c < c1; ++c) mx = fmaxf(mx, er[c]);
      mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, 16));
      if (fr < rows) {
        // a row that is entirely -inf keeps p = 0 (dead frame); +inf / NaN rows surface
        // through the row-sum certificate
        const float base = (mx == kNegInf) ? 0.f : mx;
        for (int c = c0; c < c1; ++c) pt[c] = __expf(er[c] - base);
        const bool phase1 = (d == 0) ? (tile < nA) : (tile >= nA);
        if (hh == 0 && phase1) msum += (double)base;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.bars[bar_full(d, buf)]);
      k = nk; d = nd; slot ^= 1; cur_tma = next_tma;
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
  const int dir = warp;  // 0: alpha (ascending), 1: beta (descending, mirrored)
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
    if (jstart / K == lane) {
#pragma unroll
      for (int i = 0; i < K; ++i)
        if (i == jstart % K) v[i] = 1.f;
      e = 0;
    }
  }

  uint32_t full_phase[kNB] = {0u, 0u};
  auto seg_of = [&](int k) { return dir == 0 ? k : nseg - 1 - k; };
  auto wait_ptile = [&](int k) -> const float* {
    const int buf = k % kNB;
    mbar_wait(&sm.bars[bar_full(dir, buf)], full_phase[buf]);
    full_phase[buf] ^= 1u;
    return sm.ptile[dir] + (size_t)buf * kSeg * Cp;
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
    const float* pt = wait_ptile(k);
    if (dir == 0) {
      for (int r = 0; r < rows; ++r) step<K, true, false>(v, abar, tp_live, pt + r * Cp, blank, f);
    } else {
      for (int r = rows - 1; r >= 0; --r) step<K, false, false>(v, abar, tp_live, pt + r * Cp, blank, f);
    }
    release_ptile(k);
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
  float* stored = sm.stored[dir];
  float* pbk = sm.pbk[dir];
  const size_t outsz = ((size_t)kSeg * C + 3) & ~(size_t)3;
  int bad = 0;   // reason bits: 4 = scale overflow in the recompute, 8 = row-sum certificate
  int obuf = 0;
  const int* runs = sm.runs + (lane >> 4) * (C + 1) * 2;   // label runs of my half (transposed pass)
  const int half_col0 = (lane >> 4) ? sm.hist[C] : 0;

  for (int k2 = 0; k2 < n2; ++k2) {
    const int k = n1 + k2;
    const int seg = seg_of(k);
    const int rows = min(kSeg, T - seg * kSeg);
    event<K>(v, e, f, lane);
    const float* pt = wait_ptile(k);

    // ---- recompute the opposite direction over this segment in the complementary scale:
    // stored * live = posterior * Zm, i.e. exponent(stored lane) = eZ - exponent(live lane)
    {
      float w[K];
      int ew;
      ckpt_load<K>(ck_other + (size_t)seg * (K + 1) * 32, w, ew, lane);
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
        for (int r = rows - 1; r >= 0; --r) {
          step<K, false, false>(w, abar, tp_rc, pt + r * Cp, blank, fr);
          float4* dst = reinterpret_cast<float4*>(stored + (size_t)r * RS + lane * K);
#pragma unroll
          for (int i = 0; i < K; i += 4) dst[i >> 2] = make_float4(w[i], w[i + 1], w[i + 2], w[i + 3]);
        }
      } else {
        for (int r = 0; r < rows; ++r) {
          step<K, true, false>(w, abar, tp_rc, pt + r * Cp, blank, fr);
          float4* dst = reinterpret_cast<float4*>(stored + (size_t)r * RS + lane * K);
#pragma unroll
          for (int i = 0; i < K; i += 4) dst[i >> 2] = make_float4(w[i], w[i + 1], w[i + 2], w[i + 3]);
        }
      }
      __syncwarp();
    }

    // ---- live sweep over the segment: posterior(state) * Zm = abar * stored
    auto combine_row = [&](int r) {
      float st[K];
      float* row = stored + (size_t)r * RS;
      float4* src = reinterpret_cast<float4*>(row + (31 - lane) * K);
#pragma unroll
      for (int i = 0; i < K; i += 4) {
        const float4 q = src[i >> 2];
        st[i] = q.x; st[i + 1] = q.y; st[i + 2] = q.z; st[i + 3] = q.w;
        // the block is read by this lane only: clear it, so that the row can be re-used
        // for the sorted posteriors with every padding column reading as zero
        src[i >> 2] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
      if (dir == 0) step<K, true, true>(v, abar, tp_live, pt + r * Cp, blank, f);
      else step<K, false, true>(v, abar, tp_live, pt + r * Cp, blank, f);
      float pbsum = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const bool lab = ((i & 1) == 1) == (dir == 0);
        const float g = abar[i] * st[K - 1 - i];
        if (lab) row[tp_live.gcol[i >> 1]] = g;
        else pbsum += g;
      }
      pbk[r * 33 + lane] = pbsum;
    };
    if (dir == 0) { for (int r = 0; r < rows; ++r) combine_row(r); }
    else { for (int r = rows - 1; r >= 0; --r) combine_row(r); }
    release_ptile(k);
    __syncwarp();

    // ---- transposed pass: lane = (frame, half); sum the posteriors per label
    {
      const int frm = lane & 15, hf = lane >> 4;
      float* ot = sm.out[dir] + (size_t)obuf * outsz;
      if (lane == 0) bulk_wait_read<1>();   // the store that last read this buffer is done
      __syncwarp();
      float rowsum = 0.f, bsum = 0.f;
      if (frm < rows) {
        const float* grow = stored + (size_t)frm * RS;
        int col = half_col0;
        for (int r = 0; runs[2 * r] >= 0; ++r) {
          const int labc = runs[2 * r], endc = runs[2 * r + 1];
          float acc = 0.f;
          for (; col < endc; col += 4) {
            const float4 q = *reinterpret_cast<const float4*>(grow + col);
            acc += (q.x + q.y) + (q.z + q.w);
          }
          rowsum += acc;
          ot[frm * C + labc] = acc * kappa;
        }
        const float* pr = pbk + frm * 33 + hf * 16;
#pragma unroll
        for (int q = 0; q < 16; ++q) bsum += pr[q];
      }
      bsum += __shfl_xor_sync(kFull, bsum, 16);
      rowsum += __shfl_xor_sync(kFull, rowsum, 16);
      rowsum += bsum;
      if (frm < rows) {
        if (hf == 0) ot[frm * C + blank] = bsum * kappa;
        if (!(fabsf(rowsum - Zm) <= 1e-3f * Zm)) bad |= 8;
      }
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
  }
  if (lane == 0) bulk_wait_all<0>();
  bad = __reduce_or_sync(kFull, (unsigned)bad);
  if (bad && lane == 0) atomicOr(&a.hazard[b], bad);
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
  // gamma columns: <= Sp/2 posteriors + 3 padding per label + alignment of the second
  // half (< 48) + the dump column; the row also holds the Sp recomputed values
  int cols = 16 * K + 3 * C + 48 + 4;
  if (cols < 32 * K) cols = 32 * K;
  RS = ((cols + 31) / 32) * 32 + 4;
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
