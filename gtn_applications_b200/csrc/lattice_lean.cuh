// Lean variant of the log-semiring lattice kernel (lattice.cuh): same recursion, same
// numerics contract (float64 offsets, per-frame posterior normalisation, fixed-point
// posterior tile), same workspace layout — but the acceptor of the utterance is
// materialised ONCE into shared memory as packed arc records, and the per-frame loops
// run on 32-bit shared-memory addresses only.
//
//   in-arc  record (8 B): { 4 src | 4 label << 16, weight }   grouped by destination
//   out-arc record (8 B): { 4 dst | 4 label << 16, weight }   grouped by source
//   (both fields are byte offsets into the alpha row / the emission row: no shift per use)
//   node record   (4 B) : { first slot | end slot << 16 }     one for in-, one for out-arcs
//
// What replaces what: forward_score(intersect(emissions, A_b)) + gtn.backward of
// criterions/ctc.py:49-51,78; asg.py:111-113,158; stc.py:85-86,113; transducer.py:283-290,321.
//
// Why: the generic kernel walked CSR arrays in global memory through 64-bit generic
// pointers, twice per arc and frame (ncu, cfg4 transducer: 880 instructions per warp and
// frame step, 61 % issue-bound).  Here one arc evaluation is LDS.64 + 2 LDS + 2 FADD, the
// four first arcs of a node stay in registers between the max and the sum pass, alpha rows
// of the backward sweep are prefetched one frame ahead, and the label that most arcs carry
// (blank, for CTC-like graphs) is summed in registers + one warp reduction instead of
// hundreds of same-address shared-memory atomics per frame.
//
// Limits (the launcher falls back to the generic kernel beyond them): nodes, arc slots and
// < 65536, nodes and classes < 16 * 1024, everything must fit 227 KB of shared memory.
#pragma once

#include "lattice.cuh"

namespace wfst {

namespace lean {

__device__ __forceinline__ float lds_f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_u(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u2(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void red_add_u(uint32_t a, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offsets of the regions inside dynamic shared memory
struct Layout {
  uint32_t bars, red, rowsum, tile0, tile1, gtile, alpha0, alpha1, node_in, node_out, nflags, fw, perm,
      in_pack, out_pack, out_gidx, in_gidx, gw, xch, total;
};

__host__ __device__ inline Layout make_layout(int Kt, int C, int npad, int aslots, int want_gw) {
  Layout l;
  uint32_t tile = (((uint32_t)Kt * (uint32_t)C + 3u) & ~3u) * 4u;
  uint32_t o = 0;
  l.bars = o; o += 16;
  l.red = o; o += 40 * 4;
  l.rowsum = o; o += (((uint32_t)Kt + 3u) & ~3u) * 4u;
  o = (o + 15u) & ~15u;
  l.tile0 = o; o += tile;
  l.tile1 = o; o += tile;
  l.gtile = o; o += tile;
  l.alpha0 = o; o += (uint32_t)npad * 4u;
  l.alpha1 = o; o += (uint32_t)npad * 4u;
  l.node_in = o; o += (uint32_t)npad * 4u;
  l.node_out = o; o += (uint32_t)npad * 4u;
  l.fw = o; o += (uint32_t)npad * 4u;
  l.perm = o; o += (uint32_t)npad * 4u;
  l.nflags = o; o += (uint32_t)npad;       // npad is a multiple of 4
  o = (o + 7u) & ~7u;
  // 4 slots of padding: the DEG register slots of the last node are read unconditionally
  l.in_pack = o; o += (uint32_t)(aslots + 4) * 8u;
  l.out_pack = o; o += (uint32_t)(aslots + 4) * 8u;
  l.out_gidx = o; o += want_gw ? (uint32_t)aslots * 4u : 0u;
  l.in_gidx = o; o += want_gw ? (uint32_t)aslots * 4u : 0u;
  l.gw = o; o += want_gw ? (uint32_t)aslots * 4u : 0u;
  o = (o + 7u) & ~7u;
  l.xch = o; o += 8u + (uint32_t)npad * 4u;   // pair kernel: peer's boundary vector (+ its float64 offset)
  l.total = (o + 15u) & ~15u;
  return l;
}

struct Args {
  LatticeArgs a;     // E, T, C, grad_scale, sign, scores, gradE, accumulate, hist, offs, Kt, npad, active
  int aslots;        // arc slots reserved per direction
  int want_gw;       // arc-weight gradients wanted
  Layout lay;        // make_layout(Kt, C, npad, aslots, want_gw), computed by the launcher: the kernels
                     // read the offsets from the constant bank instead of recomputing them in the loops
};

// what a builder sees while it materialises the acceptor
struct Build {
  uint32_t node_in, node_out, nflags, fw, in_pack, out_pack, out_gidx, in_gidx;
  int want_gw;
  __device__ __forceinline__ void node(int v, uint32_t in_beg, uint32_t in_end, uint32_t out_beg,
                                       uint32_t out_end, int start, int accept, float final_w) const {
    sts_u(node_in + 4u * v, in_beg | (in_end << 16));
    sts_u(node_out + 4u * v, out_beg | (out_end << 16));
    sts_f(fw + 4u * v, final_w);
    uint8_t f = (uint8_t)((start ? 1 : 0) | (accept ? 2 : 0));
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(nflags + (uint32_t)v), "r"((uint32_t)f) : "memory");
  }
  __device__ __forceinline__ void in_arc(uint32_t slot, int src, int label, float w, int gidx) const {
    sts_u2(in_pack + 8u * slot, ((uint32_t)src << 2) | ((uint32_t)label << 18), __float_as_uint(w));
    if (want_gw) sts_u(in_gidx + 4u * slot, (uint32_t)gidx);
  }
  __device__ __forceinline__ void out_arc(uint32_t slot, int dst, int label, float w, int gidx) const {
    sts_u2(out_pack + 8u * slot, ((uint32_t)dst << 2) | ((uint32_t)label << 18), __float_as_uint(w));
    if (want_gw) sts_u(out_gidx + 4u * slot, (uint32_t)gidx);
  }
};

// NPT: nodes per thread; DEG: arcs of a node kept in registers between the max and the sum
// pass; TAIL: nodes may have more than DEG arcs (extra arcs are evaluated twice, from
// shared memory).  Builders with a known maximum degree set DEG to it and TAIL = false.
template <class Builder, int NPT>
__global__ void __launch_bounds__(1024, 1) lattice_lean_kernel(Args g, typename Builder::Params bp) {
  constexpr int DEG = Builder::kDeg;
  constexpr bool TAIL = Builder::kTail;
  extern __shared__ __align__(16) unsigned char smem_lean[];
  const LatticeArgs& a = g.a;
  const int b = blockIdx.x;
  if (a.active && a.active[b] == 0) return;
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
  const Layout& L = g.lay;
  const uint32_t sb = smem_u32(smem_lean);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_lean + L.bars);
  float* red = reinterpret_cast<float*>(smem_lean + L.red);
  float* gt = reinterpret_cast<float*>(smem_lean + L.gtile);
  const uint32_t tile_bytes = L.tile1 - L.tile0;
  auto tilep = [&](int buf) { return reinterpret_cast<float*>(smem_lean + L.tile0 + (uint32_t)buf * tile_bytes); };
  auto s_tile = [&](int buf) { return sb + L.tile0 + (uint32_t)buf * tile_bytes; };
  const uint32_t s_gt = sb + L.gtile, s_rowsum = sb + L.rowsum;
  const uint32_t s_node_in = sb + L.node_in, s_node_out = sb + L.node_out, s_flags = sb + L.nflags,
                 s_fw = sb + L.fw, s_in = sb + L.in_pack, s_out = sb + L.out_pack,
                 s_gidx = sb + L.out_gidx, s_gw = sb + L.gw;
  const uint32_t s_in_gidx = sb + L.in_gidx;

  Builder bld;
  bld.init(bp, b);
  const int N = bld.num_nodes();
  const int A = bld.num_slots();
  {
    Build bd{s_node_in, s_node_out, s_flags, s_fw, s_in, s_out, s_gidx, s_in_gidx, g.want_gw};
    // slots no arc owns (gaps of fixed-stride builders, the padding) are read by the register
    // slots of a node and discarded: they must address valid memory (node 0, label 0)
    for (int k = tid; k < g.aslots + 4; k += NT) {
      sts_u2(s_in + 8u * k, 0u, 0u);
      sts_u2(s_out + 8u * k, 0u, 0u);
    }
    __syncthreads();
    bld.build(bd);
    if (g.want_gw)
      for (int k = tid; k < A; k += NT) sts_f(s_gw + 4u * k, 0.f);
  }
  const int T = a.T, C = a.C, Kt = a.Kt;
  const int eb = a.e_mod > 0 ? b % a.e_mod : b;      // emission item of this block
  const float* Eb = a.E + (size_t)eb * T * C;
  float* hist = a.hist + (size_t)b * (T + 1) * a.hist_stride;
  double* offA = a.offs + (size_t)b * a.offs_stride;
  const int ntiles = (T + Kt - 1) / Kt;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  uint32_t phase = 0u;   // bit `buf` = parity the next wait on barrier `buf` expects
  __syncthreads();

  // Which nodes a thread owns.  A warp steps through the arcs of its 32 nodes in lock step, so
  // its cost is the LARGEST degree among them: builders with irregular graphs (kSort) hand out
  // the nodes in order of decreasing degree (counting sort on max(in, out) degree), snake-wise
  // over the rounds so that every warp gets heavy and light rounds.  History rows are indexed
  // by slot (thread + round), which keeps them coalesced in both sweeps.
  int vnode[NPT];
  if (Builder::kSort) {
    const uint32_t s_perm = sb + L.perm;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(red);     // 33 counters
    if (tid < 34) cnt[tid] = 0u;
    __syncthreads();
    auto key_of = [&](int v) {
      const uint32_t bi = lds_u(s_node_in + 4u * v), bo = lds_u(s_node_out + 4u * v);
      const uint32_t d = max((bi >> 16) - (bi & 0xffffu), (bo >> 16) - (bo & 0xffffu));
      return 31u - min(d, 31u);                             // bucket 0 = heaviest
    };
    for (int v = tid; v < N; v += NT) atomicAdd(&cnt[key_of(v) + 1], 1u);
    __syncthreads();
    if (tid == 0)
      for (int k = 1; k < 33; ++k) cnt[k] += cnt[k - 1];    // cnt[k] = first slot of bucket k
    __syncthreads();
    for (int v = tid; v < N; v += NT) sts_u(s_perm + 4u * atomicAdd(&cnt[key_of(v)], 1u), (uint32_t)v);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int q = j * NT + ((j & 1) ? NT - 1 - tid : tid);
      vnode[j] = (q < N) ? (int)lds_u(s_perm + 4u * q) : -1;
    }
    __syncthreads();                                        // `red` is reused below
  } else {
#pragma unroll
    for (int j = 0; j < NPT; ++j) vnode[j] = (tid + j * NT < N) ? tid + j * NT : -1;
  }
  auto slot_of = [&](int j) { return (Builder::kSort && (j & 1)) ? j * NT + NT - 1 - tid : j * NT + tid; };

  auto is_start = [&](int v) {
    uint32_t f;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(f) : "r"(s_flags + (uint32_t)v));
    return (f & 1u) != 0u;
  };
  auto is_accept = [&](int v) {
    uint32_t f;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(f) : "r"(s_flags + (uint32_t)v));
    return (f & 2u) != 0u;
  };
  auto tile_rows = [&](int i) { return min(Kt, T - i * Kt); };
  auto tile_tma_ok = [&](int i) {
    const float* src = Eb + (size_t)i * Kt * C;
    uint32_t bytes = (uint32_t)tile_rows(i) * C * 4u;
    return ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15u) == 0);
  };
  auto issue_tile = [&](int i, int buf) {
    const float* src = Eb + (size_t)i * Kt * C;
    int n = tile_rows(i) * C;
    if (tile_tma_ok(i)) {
      if (tid == 0) {
        mbar_expect_tx(&bars[buf], (uint32_t)n * 4u);
        bulk_g2s(tilep(buf), src, (uint32_t)n * 4u, &bars[buf]);
      }
    } else {
      float* dstp = tilep(buf);
      for (int k = tid; k < n; k += NT) dstp[k] = __ldg(src + k);
    }
  };
  auto wait_tile = [&](int i, int buf) {
    if (tile_tma_ok(i)) {
      mbar_wait(&bars[buf], (phase >> buf) & 1u);
      phase ^= 1u << buf;
    }
  };

  // label carried by most arcs: its posterior mass is summed in registers
  uint32_t cstar = 0;
  if (a.gradE != nullptr) {
    for (int c = tid; c < C; c += NT) sts_u(s_gt + 4u * c, 0u);
    __syncthreads();
    for (int v = tid; v < N; v += NT) {
      const uint32_t be = lds_u(s_node_out + 4u * v);
      for (uint32_t k = be & 0xffffu; k < (be >> 16); ++k) red_add_u(s_gt + (lds_u(s_out + 8u * k) >> 16), 1u);
    }
    __syncthreads();
    uint32_t best = 0;
    for (int c = tid; c < C; c += NT) {
      uint32_t n = min(lds_u(s_gt + 4u * c), 0xffffu);
      best = max(best, (n << 16) | (uint32_t)(0xffff - c));
    }
    best = __reduce_max_sync(0xffffffffu, best);
    uint32_t* redu = reinterpret_cast<uint32_t*>(red);
    if (lane == 0) redu[tid >> 5] = best;
    __syncthreads();
    best = (lane < ((NT + 31) >> 5)) ? redu[lane] : 0u;
    best = __reduce_max_sync(0xffffffffu, best);
    cstar = 0xffffu - (best & 0xffffu);
    if (cstar >= (uint32_t)C) cstar = 0;
    cstar <<= 2;     // byte offset, like the label field of the arc records
    __syncthreads();
  }

  // ------------------------------------------------------------- forward
  uint32_t cur = sb + L.alpha0, nxt = sb + L.alpha1;
  double cumA = 0.0;
  for (int v = tid; v < N; v += NT) sts_f(cur + 4u * v, is_start(v) ? 0.f : kNegInf);
#pragma unroll
  for (int j = 0; j < NPT; ++j)
    if (vnode[j] >= 0) hist[slot_of(j)] = is_start(vnode[j]) ? 0.f : kNegInf;
  if (ntiles > 0) issue_tile(0, 0);
  uint32_t be_in[NPT];
#pragma unroll
  for (int j = 0; j < NPT; ++j) be_in[j] = (vnode[j] >= 0) ? lds_u(s_node_in + 4u * vnode[j]) : 0u;
  __syncthreads();
  for (int i = 0; i < ntiles; ++i) {
    const int buf = i & 1;
    wait_tile(i, buf);
    if (i + 1 < ntiles) issue_tile(i + 1, buf ^ 1);
    const int rows = tile_rows(i);
    if (i % a.renorm_every == 0) {
      float pm = kNegInf;
      for (int v = tid; v < N; v += NT) pm = fmaxf(pm, lds_f(cur + 4u * v));
      const float mx = block_max(pm, red);
      if (mx != kNegInf && mx != -kNegInf && mx == mx) {
        for (int v = tid; v < N; v += NT) sts_f(cur + 4u * v, lds_f(cur + 4u * v) - mx);
        cumA += (double)mx;
      }
      __syncthreads();
    }
    if (tid == 0) { offA[2 * i] = cumA; offA[2 * i + 1] = 0.0; }
    for (int tt = 0; tt < rows; ++tt) {
      const uint32_t Et = s_tile(buf) + 4u * (uint32_t)(tt * C);
      float* hrow = hist + (size_t)(i * Kt + tt + 1) * a.hist_stride;
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        const int v = vnode[j];
        if (v >= 0) {
          const uint32_t k0 = be_in[j] & 0xffffu, ke = be_in[j] >> 16;
          // records first, then the gathers they address: independent loads stay in flight together
          uint2 rec[DEG];
#pragma unroll
          for (int d = 0; d < DEG; ++d) rec[d] = lds_u2(s_in + 8u * (k0 + d));
          float x[DEG];
#pragma unroll
          for (int d = 0; d < DEG; ++d) {
            const float av = lds_f(cur + (rec[d].x & 0xffffu)), ev = lds_f(Et + (rec[d].x >> 16));
            x[d] = (k0 + d < ke) ? av + ev + __uint_as_float(rec[d].y) : kNegInf;
          }
          float m = x[0];
#pragma unroll
          for (int d = 1; d < DEG; ++d) m = fmaxf(m, x[d]);
          auto eval = [&](uint32_t k) {
            const uint2 r = lds_u2(s_in + 8u * k);
            return lds_f(cur + (r.x & 0xffffu)) + lds_f(Et + (r.x >> 16)) + __uint_as_float(r.y);
          };
          if (TAIL)
            for (uint32_t k = k0 + DEG; k < ke; ++k) m = fmaxf(m, eval(k));
          float r = kNegInf;
          if (m != kNegInf) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < DEG; ++d) s += __expf(x[d] - m);
            if (TAIL)
              for (uint32_t k = k0 + DEG; k < ke; ++k) s += __expf(eval(k) - m);
            r = m + __logf(s);
          }
          sts_f(nxt + 4u * v, r);
          hrow[slot_of(j)] = r;
        }
      }
      __syncthreads();
      const uint32_t tmp = cur; cur = nxt; nxt = tmp;
    }
  }

  // ------------------------------------------------------------- Z
  float part = kNegInf;
  for (int v = tid; v < N; v += NT)
    if (is_accept(v)) part = log_add(part, lds_f(cur + 4u * v) + lds_f(s_fw + 4u * v));
  const float Zn = block_lse(part, red);
  const double Zd = (double)Zn + cumA;
  const float Z = (float)Zd;
  if (tid == 0) a.scores[b] = Z;
  const bool want_gE = a.gradE != nullptr;
  const bool want_gW = g.want_gw != 0;
  if (!want_gE && !want_gW) return;
  const float gs = a.sign * (a.grad_scale ? a.grad_scale[b] : 1.f);
  float* gEb = want_gE ? a.gradE + (size_t)eb * T * C : nullptr;
  const bool feasible = (Z != kNegInf) && (Z == Z) && (Z != -kNegInf);
  if (!feasible) {
    if (want_gE && !a.accumulate)
      for (size_t k = tid; k < (size_t)T * C; k += NT) gEb[k] = 0.f;
    bld.finish(s_gw, s_gidx, s_node_out, 0.f, g.want_gw, false);
    return;
  }

  // ------------------------------------------------------------- backward
  __syncthreads();
  for (int v = tid; v < N; v += NT) {
    const bool acc = is_accept(v);
    const float fwv = lds_f(s_fw + 4u * v), av = lds_f(cur + 4u * v);
    sts_f(nxt + 4u * v, acc ? fwv : kNegInf);
    if (acc && av != kNegInf) bld.add_final_grad(v, __expf(av + fwv - Zn) * gs);
  }
  double cumB = 0.0;
  if (ntiles > 0) issue_tile(ntiles - 1, (ntiles - 1) & 1);
  // alpha rows are prefetched one frame ahead (they come from HBM / L2)
  float au_next[NPT];
  uint32_t be_out[NPT];
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    const int u = vnode[j];
    au_next[j] = (u >= 0 && T > 0) ? hist[(size_t)(T - 1) * a.hist_stride + slot_of(j)] : kNegInf;
    be_out[j] = (u >= 0) ? lds_u(s_node_out + 4u * u) : 0u;
  }
  __syncthreads();
  for (int i = ntiles - 1; i >= 0; --i) {
    const int buf = i & 1;
    wait_tile(i, buf);
    if (i > 0) issue_tile(i - 1, buf ^ 1);
    const int rows = tile_rows(i);
    if (want_gE) {
      for (int k = tid; k < rows * C; k += NT) sts_u(s_gt + 4u * k, 0u);
      if (tid < rows) sts_u(s_rowsum + 4u * tid, 0u);
    }
    if ((ntiles - 1 - i) % a.renorm_every == 0) {
      float pm = kNegInf;
      for (int v = tid; v < N; v += NT) pm = fmaxf(pm, lds_f(nxt + 4u * v));
      const float mx = block_max(pm, red);
      if (mx != kNegInf && mx != -kNegInf && mx == mx) {
        for (int v = tid; v < N; v += NT) sts_f(nxt + 4u * v, lds_f(nxt + 4u * v) - mx);
        cumB += (double)mx;
      }
    }
    __syncthreads();
    const double oa_first = (i > 0) ? offA[2 * (i - 1)] : 0.0;
    const double oa_base = offA[2 * i];
    for (int tt = rows - 1; tt >= 0; --tt) {
      const uint32_t Et = s_tile(buf) + 4u * (uint32_t)(tt * C);
      const int t = i * Kt + tt;
      const uint32_t grow = s_gt + 4u * (uint32_t)(tt * C);
      const float dlt = (float)(((tt == 0) ? oa_first : oa_base) + cumB - Zd);
      float au[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        au[j] = au_next[j];
        au_next[j] = (vnode[j] >= 0 && t > 0) ? hist[(size_t)(t - 1) * a.hist_stride + slot_of(j)] : kNegInf;
      }
      uint32_t qstar = 0, qtot = 0;
      auto post = [&](float xv, uint32_t rx, uint32_t k, float off) {
        const float p = __expf(xv + off);
        if (want_gE) {
          const uint32_t q = __float2uint_rn(p * kFixOne);
          const uint32_t lab = rx >> 16;
          qtot += q;
          if (lab == cstar) qstar += q;
          else if (q != 0u) red_add_u(grow + lab, q);
        }
        if (want_gW && p != 0.f) sts_f(s_gw + 4u * k, lds_f(s_gw + 4u * k) + p);
      };
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        const int u = vnode[j];
        if (u >= 0) {
          const uint32_t k0 = be_out[j] & 0xffffu, ke = be_out[j] >> 16;
          uint2 rec[DEG];
#pragma unroll
          for (int d = 0; d < DEG; ++d) rec[d] = lds_u2(s_out + 8u * (k0 + d));
          float x[DEG];
#pragma unroll
          for (int d = 0; d < DEG; ++d) {
            const float bv = lds_f(nxt + (rec[d].x & 0xffffu)), ev = lds_f(Et + (rec[d].x >> 16));
            x[d] = (k0 + d < ke) ? ev + __uint_as_float(rec[d].y) + bv : kNegInf;
          }
          float m = x[0];
#pragma unroll
          for (int d = 1; d < DEG; ++d) m = fmaxf(m, x[d]);
          uint32_t rr = 0;
          auto eval = [&](uint32_t k, uint32_t& rx) {
            const uint2 r = lds_u2(s_out + 8u * k);
            rx = r.x;
            return lds_f(Et + (r.x >> 16)) + __uint_as_float(r.y) + lds_f(nxt + (r.x & 0xffffu));
          };
          if (TAIL)
            for (uint32_t k = k0 + DEG; k < ke; ++k) m = fmaxf(m, eval(k, rr));
          float r = kNegInf;
          if (m != kNegInf) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < DEG; ++d) s += __expf(x[d] - m);
            if (TAIL)
              for (uint32_t k = k0 + DEG; k < ke; ++k) s += __expf(eval(k, rr) - m);
            r = m + __logf(s);
            if (au[j] != kNegInf) {
              const float off = au[j] + dlt;
#pragma unroll
              for (int d = 0; d < DEG; ++d)
                if (x[d] != kNegInf) post(x[d], rec[d].x, k0 + d, off);
              if (TAIL)
                for (uint32_t k = k0 + DEG; k < ke; ++k) {
                  const float xv = eval(k, rr);
                  if (xv != kNegInf) post(xv, rr, k, off);
                }
            }
          }
          sts_f(cur + 4u * u, r);
        }
      }
      if (want_gE) {
        __syncwarp();
        qstar = __reduce_add_sync(0xffffffffu, qstar);
        qtot = __reduce_add_sync(0xffffffffu, qtot);
        if (lane == 0) {
          if (qstar) red_add_u(grow + cstar, qstar);
          if (qtot) red_add_u(s_rowsum + 4u * (uint32_t)tt, qtot);
        }
      }
      __syncthreads();
      const uint32_t tmp = cur; cur = nxt; nxt = tmp;
    }
    if (want_gE) {
      // the posteriors of a frame sum to one: normalise each row by its own sum
      // (removes the common-mode rounding error of alpha + beta - Z); the fixed-point unit cancels
      const uint32_t* gtu = reinterpret_cast<const uint32_t*>(gt);
      for (int r = 0; r < rows; ++r) {
        const uint32_t rs = lds_u(s_rowsum + 4u * (uint32_t)r);
        const float f = rs ? gs / (float)rs : 0.f;
        for (int c = tid; c < C; c += NT) gt[r * C + c] = (float)gtu[r * C + c] * f;
      }
      float* dst = gEb + (size_t)i * Kt * C;
      const int n = rows * C;
      const bool tma = !a.accumulate && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n & 3) == 0);
      if (tma) {
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
          bulk_s2g(dst, gt, (uint32_t)n * 4u);
          bulk_commit();
          bulk_wait_read<0>();
        }
      } else {
        __syncthreads();
        if (a.e_mod > 0) {
          // cross launch: other blocks (other graphs against the same emission item) add here too
          for (int k = tid; k < n; k += NT)
            if (gt[k] != 0.f) atomicAdd(&dst[k], gt[k]);
        } else if (a.accumulate) {
          for (int k = tid; k < n; k += NT) dst[k] += gt[k];
        } else {
          for (int k = tid; k < n; k += NT) dst[k] = gt[k];
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) bulk_wait_all<0>();
  __syncthreads();
  bld.finish(s_gw, s_gidx, s_node_out, gs, g.want_gw, false);
}


// =======================================================================================
// Pair kernel: the same lattice on a CLUSTER OF TWO thread blocks that meet in the middle.
// A block of the kernel above walks 2T dependent frame steps (alpha up, then beta down) and
// that chain, not memory, is what bounds it.  Here rank 0 owns the frames [0, Th) and rank 1
// the frames [Th, T) of the same utterance:
//   primary sweep    rank 0: alpha_0 -> alpha_Th (in-arcs, time up)     | rank 1: beta_T -> beta_Th (out-arcs, time down)
//   exchange         each block stores its boundary vector (and its float64 offset) into the
//                    PEER's shared memory (st.shared::cluster) + one cluster barrier;
//                    Z = LSE_v alpha_Th[v] + beta_Th[v] in both blocks
//   secondary sweep  rank 0: beta_Th -> beta_0 over its frames, posteriors with its alpha rows
//                    rank 1: alpha_Th -> alpha_T over its frames, posteriors with its beta rows
// so each block walks T frame steps and twice as many SMs work (B = 64 utterances: 128 SMs).
// Both ranks run ONE code path: a sweep gathers over the arc records of its direction,
//   R_{s+1}[n] = LSE_k  val[other node of k] + E[f(s), label k] + w_k,     f(s) = s | T-1-s
// the primary sweep stores R_1..R_{S-1} (history rows s, slot indexed) and the secondary sweep
// Q_s = step(Q_{s+1}) takes the posterior of arc k into node n as exp(x_k + R_s[n] - Z).
// =======================================================================================
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_peer(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_peer_f(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_peer_u2(uint32_t addr, uint32_t x, uint32_t y) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}

// TPN: threads per node.  1: a thread walks its nodes' arc lists alone (NPT nodes per thread).
// 32: a WARP per node (NPT nodes per warp) -- lane l takes arcs l, l + 32, ... of the list, the
// maximum and the sum are combined with shuffles, lane 0 stores the node's value.  For small
// dense acceptors (the n-gram transition graph of the transducer: 84 nodes x 81 arcs), where a
// thread per node leaves 84 threads walking 81 arcs each, three times per frame step.
template <class Builder, int NPT, bool GW, int TPN = 1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(1024, 1)
lattice_lean_pair_kernel(Args g, typename Builder::Params bp) {
  constexpr int DEG = Builder::kDeg;
  constexpr bool TAIL = Builder::kTail;
  extern __shared__ __align__(16) unsigned char smem_lean[];
  const LatticeArgs& a = g.a;
  const int b = blockIdx.x >> 1;
  const uint32_t role = cluster_rank();
  if (a.active && a.active[b] == 0) return;        // both blocks of the pair
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
  const int gid = tid / TPN, NG = NT / TPN, gl = tid % TPN;     // node group of this thread, lane in it
  const Layout& L = g.lay;
  const uint32_t sb = smem_u32(smem_lean);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_lean + L.bars);
  float* red = reinterpret_cast<float*>(smem_lean + L.red);
  float* gt = reinterpret_cast<float*>(smem_lean + L.gtile);
  const uint32_t tile_bytes = L.tile1 - L.tile0;
  auto tilep = [&](int buf) { return reinterpret_cast<float*>(smem_lean + L.tile0 + (uint32_t)buf * tile_bytes); };
  auto s_tile = [&](int buf) { return sb + L.tile0 + (uint32_t)buf * tile_bytes; };
  const uint32_t s_gt = sb + L.gtile, s_rowsum = sb + L.rowsum;
  const uint32_t s_node_in = sb + L.node_in, s_node_out = sb + L.node_out, s_flags = sb + L.nflags,
                 s_fw = sb + L.fw, s_in = sb + L.in_pack, s_out = sb + L.out_pack,
                 s_out_gidx = sb + L.out_gidx, s_in_gidx = sb + L.in_gidx, s_gw = sb + L.gw,
                 s_xch = sb + L.xch;
  // arc records / node records / arc -> weight index of the two sweeps of this rank
  const uint32_t s_ppack = role ? s_out : s_in, s_pnode = role ? s_node_out : s_node_in;
  const uint32_t s_spack = role ? s_in : s_out, s_snode = role ? s_node_in : s_node_out;
  const uint32_t s_sgidx = role ? s_in_gidx : s_out_gidx;

  Builder bld;
  bld.init(bp, b);
  const int N = bld.num_nodes();
  const int A = bld.num_slots();
  {
    Build bd{s_node_in, s_node_out, s_flags, s_fw, s_in, s_out, s_out_gidx, s_in_gidx, g.want_gw};
    for (int k = tid; k < g.aslots + 4; k += NT) {
      sts_u2(s_in + 8u * k, 0u, 0u);
      sts_u2(s_out + 8u * k, 0u, 0u);
    }
    __syncthreads();
    bld.build(bd);
    if (g.want_gw) {
      for (int k = tid; k < A; k += NT) sts_f(s_gw + 4u * k, 0.f);
      // both ranks add their half of the frames to the utterance's weight gradient: rank 0 clears
      // it here, before the first cluster barrier (shared gradients are cleared by the host)
      if (role == 0) bld.zero_weight_grad();
    }
  }
  const int T = a.T, C = a.C, Kt = a.Kt;
  const float* Eb = a.E + (size_t)b * T * C;
  const int ntiles = (T + Kt - 1) / Kt;            // >= 2 (launcher)
  const int nt0 = ntiles / 2;                       // tiles of rank 0 (all full)
  const int ntr = role ? ntiles - nt0 : nt0;        // tiles of this rank
  const int S0 = nt0 * Kt;
  const int S = role ? T - S0 : S0;                 // frame steps of this rank
  float* hist = a.hist + ((size_t)b * (T + 1) + (role ? S0 : 0)) * a.hist_stride;   // rows 0 .. S-1
  double* offP = a.offs + (size_t)b * a.offs_stride;                                // indexed by global tile
  auto gtile_of = [&](int j) { return role ? ntiles - 1 - j : j; };
  auto rows_of = [&](int gti) { return min(Kt, T - gti * Kt); };
  auto sbase_of = [&](int j) { return role ? (j == 0 ? 0 : rows_of(ntiles - 1) + (j - 1) * Kt) : j * Kt; };
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  uint32_t phase = 0u;
  __syncthreads();

  auto flag_of = [&](int v) {
    uint32_t f;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(f) : "r"(s_flags + (uint32_t)v));
    return f;
  };
  auto tile_tma_ok = [&](int gti) {
    const float* src = Eb + (size_t)gti * Kt * C;
    uint32_t bytes = (uint32_t)rows_of(gti) * C * 4u;
    return ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15u) == 0);
  };
  auto issue_tile = [&](int gti, int buf) {
    const float* src = Eb + (size_t)gti * Kt * C;
    int n = rows_of(gti) * C;
    if (tile_tma_ok(gti)) {
      if (tid == 0) {
        mbar_expect_tx(&bars[buf], (uint32_t)n * 4u);
        bulk_g2s(tilep(buf), src, (uint32_t)n * 4u, &bars[buf]);
      }
    } else {
      float* dstp = tilep(buf);
      for (int k = tid; k < n; k += NT) dstp[k] = __ldg(src + k);
    }
  };
  auto wait_tile = [&](int gti, int buf) {
    if (tile_tma_ok(gti)) {
      mbar_wait(&bars[buf], (phase >> buf) & 1u);
      phase ^= 1u << buf;
    }
  };

  // node -> thread assignment (see the single-block kernel)
  int vnode[NPT];
  if (Builder::kSort) {
    const uint32_t s_perm = sb + L.perm;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(red);
    if (tid < 34) cnt[tid] = 0u;
    __syncthreads();
    auto key_of = [&](int v) {
      const uint32_t bi = lds_u(s_node_in + 4u * v), bo = lds_u(s_node_out + 4u * v);
      // nodes with at most DEG arcs cost the same (the register path evaluates DEG slots): one
      // bucket, in which they keep the order of their ids -- neighbouring nodes gather from
      // neighbouring alpha rows, i.e. from different banks
      const uint32_t d = max(max((bi >> 16) - (bi & 0xffffu), (bo >> 16) - (bo & 0xffffu)), (uint32_t)DEG);
      return 31u - min(d, 31u);
    };
    for (int v = tid; v < N; v += NT) atomicAdd(&cnt[key_of(v) + 1], 1u);
    __syncthreads();
    if (tid == 0)
      for (int k = 1; k < 33; ++k) cnt[k] += cnt[k - 1];
    __syncthreads();
    for (int v = tid; v < N; v += NT) sts_u(s_perm + 4u * atomicAdd(&cnt[key_of(v)], 1u), (uint32_t)v);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int q = j * NG + ((j & 1) ? NG - 1 - gid : gid);
      vnode[j] = (q < N) ? (int)lds_u(s_perm + 4u * q) : -1;
    }
    __syncthreads();
  } else {
#pragma unroll
    for (int j = 0; j < NPT; ++j) vnode[j] = (gid + j * NG < N) ? gid + j * NG : -1;
  }
  auto slot_of = [&](int j) { return (Builder::kSort && (j & 1)) ? j * NG + NG - 1 - gid : j * NG + gid; };
  uint32_t hslot[NPT];     // history rows are addressed by 32-bit element offsets (row * stride + slot)
#pragma unroll
  for (int j = 0; j < NPT; ++j) hslot[j] = (uint32_t)slot_of(j);
  const uint32_t hstride = (uint32_t)a.hist_stride;

  // label carried by most arcs of the secondary direction
  uint32_t cstar = 0;
  if (a.gradE != nullptr) {
    for (int c = tid; c < C; c += NT) sts_u(s_gt + 4u * c, 0u);
    __syncthreads();
    for (int v = tid; v < N; v += NT) {
      const uint32_t be = lds_u(s_snode + 4u * v);
      for (uint32_t k = be & 0xffffu; k < (be >> 16); ++k) red_add_u(s_gt + (lds_u(s_spack + 8u * k) >> 16), 1u);
    }
    __syncthreads();
    uint32_t best = 0;
    for (int c = tid; c < C; c += NT) {
      uint32_t n = min(lds_u(s_gt + 4u * c), 0xffffu);
      best = max(best, (n << 16) | (uint32_t)(0xffff - c));
    }
    best = __reduce_max_sync(0xffffffffu, best);
    uint32_t* redu = reinterpret_cast<uint32_t*>(red);
    if (lane == 0) redu[tid >> 5] = best;
    __syncthreads();
    best = (lane < ((NT + 31) >> 5)) ? redu[lane] : 0u;
    best = __reduce_max_sync(0xffffffffu, best);
    cstar = 0xffffu - (best & 0xffffu);
    if (cstar >= (uint32_t)C) cstar = 0;
    cstar <<= 2;     // byte offset, like the label field of the arc records
    __syncthreads();
  }
  cluster_sync_all();     // the peer block runs: its shared memory may be written from here on

  // ------------------------------------------------------------- primary sweep
  uint32_t cur = sb + L.alpha0, nxt = sb + L.alpha1;
  double cumP = 0.0;
  auto init_val = [&](int v) {
    const uint32_t f = flag_of(v);
    return role ? ((f & 2u) ? lds_f(s_fw + 4u * v) : kNegInf) : ((f & 1u) ? 0.f : kNegInf);
  };
  for (int v = tid; v < N; v += NT) sts_f(cur + 4u * v, init_val(v));
#pragma unroll
  for (int j = 0; j < NPT; ++j)
    if (vnode[j] >= 0 && gl == 0) hist[hslot[j]] = init_val(vnode[j]);
  issue_tile(gtile_of(0), 0);
  uint32_t be_p[NPT];
#pragma unroll
  for (int j = 0; j < NPT; ++j) be_p[j] = (vnode[j] >= 0) ? lds_u(s_pnode + 4u * vnode[j]) : 0u;
  __syncthreads();
  for (int jt = 0; jt < ntr; ++jt) {
    const int buf = jt & 1, gti = gtile_of(jt);
    wait_tile(gti, buf);
    if (jt + 1 < ntr) issue_tile(gtile_of(jt + 1), buf ^ 1);
    const int rows = rows_of(gti), sbase = sbase_of(jt);
    if (jt % a.renorm_every == 0) {
      float pm = kNegInf;
      for (int v = tid; v < N; v += NT) pm = fmaxf(pm, lds_f(cur + 4u * v));
      const float mx = block_max(pm, red);
      if (mx != kNegInf && mx != -kNegInf && mx == mx) {
        for (int v = tid; v < N; v += NT) sts_f(cur + 4u * v, lds_f(cur + 4u * v) - mx);
        cumP += (double)mx;
      }
      __syncthreads();
    }
    if (tid == 0) offP[gti] = cumP;
    for (int r = 0; r < rows; ++r) {
      const int tt = role ? rows - 1 - r : r;
      const int s = sbase + r;
      const uint32_t Et = s_tile(buf) + 4u * (uint32_t)(tt * C);
      const uint32_t hrow = (uint32_t)(s + 1) * hstride;
      const bool keep = s + 1 < S;                   // R_S goes to the peer, not to the history
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        const int v = vnode[j];
        if (v >= 0) {
          const uint32_t k0 = be_p[j] & 0xffffu, ke = be_p[j] >> 16;
          uint2 rec[DEG];
#pragma unroll
          for (int d = 0; d < DEG; ++d) {
            // TPN == 1 reads up to DEG - 1 records past the list (padding slots of the layout)
            const uint32_t kk = k0 + gl + d * TPN;
            rec[d] = (TPN == 1 || kk < ke) ? lds_u2(s_ppack + 8u * kk) : make_uint2(0u, 0u);
          }
          float x[DEG];
#pragma unroll
          for (int d = 0; d < DEG; ++d) {
            const float av = lds_f(cur + (rec[d].x & 0xffffu)), ev = lds_f(Et + (rec[d].x >> 16));
            x[d] = (k0 + gl + d * TPN < ke) ? av + ev + __uint_as_float(rec[d].y) : kNegInf;
          }
          float m = x[0];
#pragma unroll
          for (int d = 1; d < DEG; ++d) m = fmaxf(m, x[d]);
          auto eval = [&](uint32_t k) {
            const uint2 rr = lds_u2(s_ppack + 8u * k);
            return lds_f(cur + (rr.x & 0xffffu)) + lds_f(Et + (rr.x >> 16)) + __uint_as_float(rr.y);
          };
          if (TAIL)
            for (uint32_t k = k0 + gl + DEG * TPN; k < ke; k += TPN) m = fmaxf(m, eval(k));
          if (TPN > 1) {
#pragma unroll
            for (int o = TPN / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          }
          float rv = kNegInf;
          if (m != kNegInf) {
            float sum = 0.f;
            const float ml = -m * kLog2e;
#pragma unroll
            for (int d = 0; d < DEG; ++d) sum += ex2_approx(fmaf(x[d], kLog2e, ml));
            if (TAIL)
              for (uint32_t k = k0 + gl + DEG * TPN; k < ke; k += TPN) sum += ex2_approx(fmaf(eval(k), kLog2e, ml));
            if (TPN > 1) {
#pragma unroll
              for (int o = TPN / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            }
            rv = m + __logf(sum);
          }
          if (gl == 0) {
            sts_f(nxt + 4u * v, rv);
            if (keep) hist[hrow + hslot[j]] = rv;
          }
        }
      }
      __syncthreads();
      const uint32_t tmp = cur; cur = nxt; nxt = tmp;
    }
  }

  // ------------------------------------------------------------- exchange, Z
  {
    const uint32_t peer = map_to_peer(s_xch, role ^ 1u);
    for (int v = tid; v < N; v += NT) st_peer_f(peer + 8u + 4u * v, lds_f(cur + 4u * v));
    if (tid == 0) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(cumP);
      st_peer_u2(peer, (uint32_t)bits, (uint32_t)(bits >> 32));
    }
  }
  cluster_sync_all();
  double cumQ;
  {
    const uint2 cb = lds_u2(s_xch);
    cumQ = __longlong_as_double((long long)(((unsigned long long)cb.y << 32) | cb.x));
  }
  float part = kNegInf;
  for (int v = tid; v < N; v += NT) part = log_add(part, lds_f(cur + 4u * v) + lds_f(s_xch + 8u + 4u * v));
  const float Zn = block_lse(part, red);
  const double Zd = (double)Zn + (cumP + cumQ);     // symmetric: both ranks get the same bits
  const float Z = (float)Zd;
  if (tid == 0 && role == 0) a.scores[b] = Z;
  const bool want_gE = a.gradE != nullptr;
  constexpr bool want_gW = GW;     // == (g.want_gw != 0), launcher
  if (!want_gE && !want_gW) return;
  const float gs = a.sign * (a.grad_scale ? a.grad_scale[b] : 1.f);
  float* gEb = want_gE ? a.gradE + (size_t)b * T * C : nullptr;
  const bool feasible = (Z != kNegInf) && (Z == Z) && (Z != -kNegInf);
  if (!feasible) {
    if (want_gE && !a.accumulate) {
      const size_t k0 = role ? (size_t)S0 * C : 0, k1 = role ? (size_t)T * C : (size_t)S0 * C;
      for (size_t k = k0 + tid; k < k1; k += NT) gEb[k] = 0.f;
    }
    return;    // weight gradients: the buffer was zeroed by the launcher
  }

  // ------------------------------------------------------------- secondary sweep
  __syncthreads();
  for (int v = tid; v < N; v += NT) sts_f(nxt + 4u * v, lds_f(s_xch + 8u + 4u * v));   // Q_S
  issue_tile(gtile_of(ntr - 1), (ntr - 1) & 1);
  float pr_next[NPT];
  uint32_t be_s[NPT];
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    pr_next[j] = (vnode[j] >= 0) ? hist[(uint32_t)(S - 1) * hstride + hslot[j]] : kNegInf;
    be_s[j] = (vnode[j] >= 0) ? lds_u(s_snode + 4u * vnode[j]) : 0u;
  }
  __syncthreads();
  for (int jt = ntr - 1; jt >= 0; --jt) {
    const int buf = jt & 1, gti = gtile_of(jt);
    wait_tile(gti, buf);
    if (jt > 0) issue_tile(gtile_of(jt - 1), buf ^ 1);
    const int rows = rows_of(gti), sbase = sbase_of(jt);
    if (want_gE) {
      for (int k = tid; k < rows * C; k += NT) sts_u(s_gt + 4u * k, 0u);
      if (tid < rows) sts_u(s_rowsum + 4u * tid, 0u);
    }
    if ((ntr - 1 - jt) % a.renorm_every == 0) {
      float pm = kNegInf;
      for (int v = tid; v < N; v += NT) pm = fmaxf(pm, lds_f(nxt + 4u * v));
      const float mx = block_max(pm, red);
      if (mx != kNegInf && mx != -kNegInf && mx == mx) {
        for (int v = tid; v < N; v += NT) sts_f(nxt + 4u * v, lds_f(nxt + 4u * v) - mx);
        cumQ += (double)mx;
      }
    }
    __syncthreads();
    // offset of the history row R_s: rows written during primary tile jt are relative to its
    // offset; the row a tile starts from belongs to the tile before (R_0: 0)
    const double off_first = (jt > 0) ? offP[gtile_of(jt - 1)] : 0.0;
    const double off_tile = offP[gti];
    for (int r = rows - 1; r >= 0; --r) {
      const int tt = role ? rows - 1 - r : r;
      const int s = sbase + r;
      const uint32_t Et = s_tile(buf) + 4u * (uint32_t)(tt * C);
      const uint32_t grow = s_gt + 4u * (uint32_t)(tt * C);
      const float dlt = (float)(((r == 0) ? off_first : off_tile) + cumQ - Zd);
      float pr[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        pr[j] = pr_next[j];
        pr_next[j] = (vnode[j] >= 0 && s > 0) ? hist[(uint32_t)(s - 1) * hstride + hslot[j]] : kNegInf;
      }
      uint32_t qstar = 0, qtot = 0;
      // pf: the posterior of the arc in the fixed-point unit of the tile (kFixOne = 2^30); an arc
      // at -inf (padding slot included) gives 0
      // warp_call: all 32 lanes are here together (TPN == 32, register slots).  The in-arcs of a
      // node of an n-gram transition graph all carry the node's own label: 32 atomics on one
      // address would serialise, so a warp whose non-zero posteriors share a label adds their
      // sum once
      auto post = [&](float pf, uint32_t rx, uint32_t k, bool warp_call) {
        const float p = pf * (1.f / kFixOne);      // used by the weight gradient only
        if (want_gE) {
          const uint32_t q = __float2uint_rn(pf);
          const uint32_t lab = rx >> 16;
          qtot += q;
          bool done = false;
          if (TPN == 32 && warp_call) {
            const uint32_t act = __ballot_sync(0xffffffffu, q != 0u);
            if (act == 0u) {
              done = true;
            } else {
              const int src = __ffs(act) - 1;
              const uint32_t lab0 = __shfl_sync(0xffffffffu, lab, src);
              if (__all_sync(0xffffffffu, q == 0u || lab == lab0)) {
                const uint32_t tot = __reduce_add_sync(0xffffffffu, q);    // posteriors of one frame: <= 2^30 in all
                if (lane == src) {
                  if (lab0 == cstar) qstar += tot;
                  else red_add_u(grow + lab0, tot);
                }
                done = true;
              }
            }
          }
          if (!done) {
            if (lab == cstar) qstar += q;
            else if (q != 0u) red_add_u(grow + lab, q);
          }
        }
        if (want_gW && p != 0.f) sts_f(s_gw + 4u * k, lds_f(s_gw + 4u * k) + p);
      };
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        const int u = vnode[j];
        if (u >= 0) {
          const uint32_t k0 = be_s[j] & 0xffffu, ke = be_s[j] >> 16;
          uint2 rec[DEG];
#pragma unroll
          for (int d = 0; d < DEG; ++d) {
            const uint32_t kk = k0 + gl + d * TPN;
            rec[d] = (TPN == 1 || kk < ke) ? lds_u2(s_spack + 8u * kk) : make_uint2(0u, 0u);
          }
          float x[DEG];
#pragma unroll
          for (int d = 0; d < DEG; ++d) {
            const float bv = lds_f(nxt + (rec[d].x & 0xffffu)), ev = lds_f(Et + (rec[d].x >> 16));
            x[d] = (k0 + gl + d * TPN < ke) ? ev + __uint_as_float(rec[d].y) + bv : kNegInf;
          }
          float m = x[0];
#pragma unroll
          for (int d = 1; d < DEG; ++d) m = fmaxf(m, x[d]);
          uint32_t rr = 0;
          auto eval = [&](uint32_t k, uint32_t& rx) {
            const uint2 q = lds_u2(s_spack + 8u * k);
            rx = q.x;
            return lds_f(Et + (q.x >> 16)) + __uint_as_float(q.y) + lds_f(nxt + (q.x & 0xffffu));
          };
          if (TAIL)
            for (uint32_t k = k0 + gl + DEG * TPN; k < ke; k += TPN) m = fmaxf(m, eval(k, rr));
          if (TPN > 1) {
#pragma unroll
            for (int o = TPN / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          }
          float rv = kNegInf;
          if (m != kNegInf) {
            float sum = 0.f;
            const float ml = -m * kLog2e;
#pragma unroll
            for (int d = 0; d < DEG; ++d) {
              x[d] = ex2_approx(fmaf(x[d], kLog2e, ml));     // exp(x - m): reused by the posterior below
              sum += x[d];
            }
            if (TAIL)
              for (uint32_t k = k0 + gl + DEG * TPN; k < ke; k += TPN) sum += ex2_approx(fmaf(eval(k, rr), kLog2e, ml));
            if (TPN > 1) {
#pragma unroll
              for (int o = TPN / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            }
            rv = m + __logf(sum);
            if (pr[j] != kNegInf) {
              // posterior * 2^30 = exp(x - m) * exp2((alpha + offsets - Z) * log2(e) + 30 + m * log2(e)):
              // one exp2 per node instead of one per arc
              const float scale = ex2_approx(fmaf(pr[j] + dlt, kLog2e, 30.f) - ml);
#pragma unroll
              for (int d = 0; d < DEG; ++d) post(x[d] * scale, rec[d].x, k0 + gl + d * TPN, true);
              if (TAIL)
                for (uint32_t k = k0 + gl + DEG * TPN; k < ke; k += TPN) {
                  const float xv = eval(k, rr);
                  post(ex2_approx(fmaf(xv, kLog2e, ml)) * scale, rr, k, false);
                }
            }
          }
          if (gl == 0) sts_f(cur + 4u * u, rv);
        }
      }
      if (want_gE) {
        __syncwarp();
        qstar = __reduce_add_sync(0xffffffffu, qstar);
        qtot = __reduce_add_sync(0xffffffffu, qtot);
        if (lane == 0) {
          if (qstar) red_add_u(grow + cstar, qstar);
          if (qtot) red_add_u(s_rowsum + 4u * (uint32_t)tt, qtot);
        }
      }
      __syncthreads();
      const uint32_t tmp = cur; cur = nxt; nxt = tmp;
    }
    if (want_gE) {
      const uint32_t* gtu = reinterpret_cast<const uint32_t*>(gt);
      for (int r = 0; r < rows; ++r) {
        const uint32_t rs = lds_u(s_rowsum + 4u * (uint32_t)r);
        const float f = rs ? gs / (float)rs : 0.f;
        for (int c = tid; c < C; c += NT) gt[r * C + c] = (float)gtu[r * C + c] * f;
      }
      float* dst = gEb + (size_t)gti * Kt * C;
      const int n = rows * C;
      const bool tma = !a.accumulate && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n & 3) == 0);
      if (tma) {
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
          bulk_s2g(dst, gt, (uint32_t)n * 4u);
          bulk_commit();
          bulk_wait_read<0>();
        }
      } else {
        __syncthreads();
        if (a.accumulate) {
          for (int k = tid; k < n; k += NT) dst[k] += gt[k];
        } else {
          for (int k = tid; k < n; k += NT) dst[k] = gt[k];
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) bulk_wait_all<0>();
  __syncthreads();
  if (role == 1) {
    // rank 1 ends with alpha_T (in `nxt`, relative to cumQ): posterior of ending in v
    for (int v = tid; v < N; v += NT) {
      const float av = lds_f(nxt + 4u * v);
      if ((flag_of(v) & 2u) && av != kNegInf)
        bld.add_final_grad(v, __expf((float)((double)av + (double)lds_f(s_fw + 4u * v) + cumQ - Zd)) * gs);
    }
  }
  bld.finish(s_gw, s_sgidx, s_snode, gs, g.want_gw, true);
}

}  // namespace lean
}  // namespace wfst
