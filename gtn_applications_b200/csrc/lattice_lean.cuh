// Lean variant of the log-semiring lattice kernel (lattice.cuh): same recursion, same
// numerics contract (float64 offsets, per-frame posterior normalisation, fixed-point
// posterior tile), same workspace layout — but the acceptor of the utterance is
// materialised ONCE into shared memory as packed arc records, and the per-frame loops
// run on 32-bit shared-memory addresses only.
//
//   in-arc  record (8 B): { src  | label << 16, weight }      grouped by destination
//   out-arc record (8 B): { dst  | label << 16, weight }      grouped by source
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
// classes < 65536, nodes <= 16 * 1024, everything must fit 227 KB of shared memory.
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

// byte offsets of the regions inside dynamic shared memory
struct Layout {
  uint32_t bars, red, rowsum, tile0, tile1, gtile, alpha0, alpha1, node_in, node_out, nflags, fw, perm,
      in_pack, out_pack, out_gidx, gw, total;
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
  l.gw = o; o += want_gw ? (uint32_t)aslots * 4u : 0u;
  l.total = (o + 15u) & ~15u;
  return l;
}

struct Args {
  LatticeArgs a;     // E, T, C, grad_scale, sign, scores, gradE, accumulate, hist, offs, Kt, npad, active
  int aslots;        // arc slots reserved per direction
  int want_gw;       // arc-weight gradients wanted
};

// what a builder sees while it materialises the acceptor
struct Build {
  uint32_t node_in, node_out, nflags, fw, in_pack, out_pack, out_gidx;
  int want_gw;
  __device__ __forceinline__ void node(int v, uint32_t in_beg, uint32_t in_end, uint32_t out_beg,
                                       uint32_t out_end, int start, int accept, float final_w) const {
    sts_u(node_in + 4u * v, in_beg | (in_end << 16));
    sts_u(node_out + 4u * v, out_beg | (out_end << 16));
    sts_f(fw + 4u * v, final_w);
    uint8_t f = (uint8_t)((start ? 1 : 0) | (accept ? 2 : 0));
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(nflags + (uint32_t)v), "r"((uint32_t)f) : "memory");
  }
  __device__ __forceinline__ void in_arc(uint32_t slot, int src, int label, float w) const {
    sts_u2(in_pack + 8u * slot, (uint32_t)src | ((uint32_t)label << 16), __float_as_uint(w));
  }
  __device__ __forceinline__ void out_arc(uint32_t slot, int dst, int label, float w, int gidx) const {
    sts_u2(out_pack + 8u * slot, (uint32_t)dst | ((uint32_t)label << 16), __float_as_uint(w));
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
  const Layout L = make_layout(a.Kt, a.C, a.npad, g.aslots, g.want_gw);
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

  Builder bld;
  bld.init(bp, b);
  const int N = bld.num_nodes();
  const int A = bld.num_slots();
  {
    Build bd{s_node_in, s_node_out, s_flags, s_fw, s_in, s_out, s_gidx, g.want_gw};
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
  const float* Eb = a.E + (size_t)b * T * C;
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
      for (uint32_t k = be & 0xffffu; k < (be >> 16); ++k) red_add_u(s_gt + 4u * (lds_u(s_out + 8u * k) >> 16), 1u);
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
            const float av = lds_f(cur + 4u * (rec[d].x & 0xffffu)), ev = lds_f(Et + 4u * (rec[d].x >> 16));
            x[d] = (k0 + d < ke) ? av + ev + __uint_as_float(rec[d].y) : kNegInf;
          }
          float m = x[0];
#pragma unroll
          for (int d = 1; d < DEG; ++d) m = fmaxf(m, x[d]);
          auto eval = [&](uint32_t k) {
            const uint2 r = lds_u2(s_in + 8u * k);
            return lds_f(cur + 4u * (r.x & 0xffffu)) + lds_f(Et + 4u * (r.x >> 16)) + __uint_as_float(r.y);
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
  float* gEb = want_gE ? a.gradE + (size_t)b * T * C : nullptr;
  const bool feasible = (Z != kNegInf) && (Z == Z) && (Z != -kNegInf);
  if (!feasible) {
    if (want_gE && !a.accumulate)
      for (size_t k = tid; k < (size_t)T * C; k += NT) gEb[k] = 0.f;
    bld.finish(s_gw, s_gidx, 0.f, g.want_gw);
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
          else if (q != 0u) red_add_u(grow + 4u * lab, q);
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
            const float bv = lds_f(nxt + 4u * (rec[d].x & 0xffffu)), ev = lds_f(Et + 4u * (rec[d].x >> 16));
            x[d] = (k0 + d < ke) ? ev + __uint_as_float(rec[d].y) + bv : kNegInf;
          }
          float m = x[0];
#pragma unroll
          for (int d = 1; d < DEG; ++d) m = fmaxf(m, x[d]);
          uint32_t rr = 0;
          auto eval = [&](uint32_t k, uint32_t& rx) {
            const uint2 r = lds_u2(s_out + 8u * k);
            rx = r.x;
            return lds_f(Et + 4u * (r.x >> 16)) + __uint_as_float(r.y) + lds_f(nxt + 4u * (r.x & 0xffffu));
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
          if (qstar) red_add_u(grow + 4u * cstar, qstar);
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
  bld.finish(s_gw, s_gidx, gs, g.want_gw);
}

}  // namespace lean
}  // namespace wfst
