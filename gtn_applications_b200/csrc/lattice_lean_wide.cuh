// "Wide" variant of the two-block cluster kernel of lattice_lean.cuh for degree-sorted (CSR)
// acceptors of at most 2048 nodes: 512 threads with 128 registers each, four nodes per thread.
// What changes against lattice_lean_pair_kernel (same contract, same workspace, same numerics):
//   * the arc records of a thread's nodes live in REGISTERS for a whole sweep (node offset, label
//     offset, weight; padding slots carry weight -inf), so a frame step reads no record from
//     shared memory and needs no predicate per arc;
//   * the nodes are handed out by decreasing degree, so the first node of every thread is one of
//     the 512 largest: it gets D0 = 8 register slots, the other three D = 4; arcs beyond the
//     register slots (none for the transducer's alignment graphs) take the tail loop;
//   * the frame step is written without control flow between the four nodes of a thread (all
//     gathers first, then the maxima, the exponentials, the logarithms), so that the loads of all
//     24 arcs are in flight together.
// Measured on cfg4 (B=64, T=1000, the reference's 1000 word pieces): see DESIGN.md §3.2.
#pragma once

#include "lattice_lean.cuh"

namespace wfst {

namespace lean {

template <class Builder, bool GW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1)
lattice_lean_wide_kernel(Args g, typename Builder::Params bp) {
  constexpr int NPT = 4;     // nodes per thread
  constexpr int D0 = 8;      // register slots of a thread's first (largest) node
  constexpr int D = 4;       // register slots of the other nodes
  constexpr int DEG = D;     // bucket limit of the degree sort
  extern __shared__ __align__(16) unsigned char smem_lean[];
  const LatticeArgs& a = g.a;
  const int b = blockIdx.x >> 1;
  const uint32_t role = cluster_rank();
  if (a.active && a.active[b] == 0) return;        // both blocks of the pair
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
      const int q = j * NT + ((j & 1) ? NT - 1 - tid : tid);
      vnode[j] = (q < N) ? (int)lds_u(s_perm + 4u * q) : -1;
    }
    __syncthreads();
  } else {
#pragma unroll
    for (int j = 0; j < NPT; ++j) vnode[j] = (tid + j * NT < N) ? tid + j * NT : -1;
  }
  auto slot_of = [&](int j) { return (Builder::kSort && (j & 1)) ? j * NT + NT - 1 - tid : j * NT + tid; };
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
    if (vnode[j] >= 0) hist[hslot[j]] = init_val(vnode[j]);
  issue_tile(gtile_of(0), 0);
  // register slots: D0 for node 0 of the thread, D for the others
  constexpr int NS = D0 + (NPT - 1) * D;
  auto sb_of = [](int j) { return j == 0 ? 0 : D0 + (j - 1) * D; };
  auto sc_of = [](int j) { return j == 0 ? D0 : D; };
  uint32_t rn[NS], rl[NS];     // byte offset of the neighbour's alpha / of the label's emission
  float rw[NS];                // arc weight, -inf in a padding slot: the arc evaluates to -inf
  uint32_t voff[NPT], k0r[NPT];
  bool tail_any = false;       // some node of this thread has more arcs than register slots
  auto load_records = [&](uint32_t s_pack, uint32_t s_node) {
    tail_any = false;
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const uint32_t be = (vnode[j] >= 0) ? lds_u(s_node + 4u * vnode[j]) : 0u;
      const uint32_t k0 = be & 0xffffu, ke = be >> 16;
      k0r[j] = k0;
      if (ke > k0 + (uint32_t)sc_of(j)) tail_any = true;
#pragma unroll
      for (int d = 0; d < D0; ++d)
        if (d < sc_of(j)) {
          const bool ok = k0 + d < ke;
          uint2 rec = make_uint2(0u, 0u);
          if (ok) rec = lds_u2(s_pack + 8u * (k0 + d));
          rn[sb_of(j) + d] = rec.x & 0xffffu;
          rl[sb_of(j) + d] = rec.x >> 16;
          rw[sb_of(j) + d] = ok ? __uint_as_float(rec.y) : kNegInf;
        }
    }
  };
#pragma unroll
  for (int j = 0; j < NPT; ++j) voff[j] = (vnode[j] >= 0) ? 4u * (uint32_t)vnode[j] : 0xffffffffu;
  // arcs beyond the register slots of node j: [k0 + slots, ke) of its node record
  auto tail_of = [&](int j, uint32_t s_node, uint32_t& kb, uint32_t& ke) {
    const uint32_t be = (vnode[j] >= 0) ? lds_u(s_node + 4u * vnode[j]) : 0u;
    kb = (be & 0xffffu) + (uint32_t)sc_of(j);
    ke = be >> 16;
  };
  load_records(s_ppack, s_pnode);
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
      float x[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) x[i] = lds_f(cur + rn[i]) + lds_f(Et + rl[i]) + rw[i];
      float m[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        m[j] = x[sb_of(j)];
#pragma unroll
        for (int d = 1; d < D0; ++d)
          if (d < sc_of(j)) m[j] = fmaxf(m[j], x[sb_of(j) + d]);
      }
      auto eval = [&](uint32_t k) {
        const uint2 rr = lds_u2(s_ppack + 8u * k);
        return lds_f(cur + (rr.x & 0xffffu)) + lds_f(Et + (rr.x >> 16)) + __uint_as_float(rr.y);
      };
      if (tail_any) {
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
          uint32_t kb, ke;
          tail_of(j, s_pnode, kb, ke);
          for (uint32_t k = kb; k < ke; ++k) m[j] = fmaxf(m[j], eval(k));
        }
      }
      float sum[NPT], ml[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        m[j] = (m[j] == kNegInf) ? 0.f : m[j];      // nothing arrives: every term is exp2(-inf) = 0
        ml[j] = -m[j] * kLog2e;
        sum[j] = 0.f;
#pragma unroll
        for (int d = 0; d < D0; ++d)
          if (d < sc_of(j)) sum[j] += ex2_approx(fmaf(x[sb_of(j) + d], kLog2e, ml[j]));
      }
      if (tail_any) {
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
          uint32_t kb, ke;
          tail_of(j, s_pnode, kb, ke);
          for (uint32_t k = kb; k < ke; ++k) sum[j] += ex2_approx(fmaf(eval(k), kLog2e, ml[j]));
        }
      }
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        const float rv = m[j] + __logf(sum[j]);      // log(0) = -inf
        if (vnode[j] >= 0) {
          sts_f(nxt + voff[j], rv);
          if (keep) hist[hrow + hslot[j]] = rv;
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
#pragma unroll
  for (int j = 0; j < NPT; ++j)
    pr_next[j] = (vnode[j] >= 0) ? hist[(uint32_t)(S - 1) * hstride + hslot[j]] : kNegInf;
  load_records(s_spack, s_snode);
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
      float x[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) x[i] = lds_f(Et + rl[i]) + rw[i] + lds_f(nxt + rn[i]);
      float m[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        m[j] = x[sb_of(j)];
#pragma unroll
        for (int d = 1; d < D0; ++d)
          if (d < sc_of(j)) m[j] = fmaxf(m[j], x[sb_of(j) + d]);
      }
      auto eval = [&](uint32_t k, uint32_t& rx) {
        const uint2 q = lds_u2(s_spack + 8u * k);
        rx = q.x;
        return lds_f(Et + (q.x >> 16)) + __uint_as_float(q.y) + lds_f(nxt + (q.x & 0xffffu));
      };
      uint32_t rr = 0;
      if (tail_any) {
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
          uint32_t kb, ke;
          tail_of(j, s_snode, kb, ke);
          for (uint32_t k = kb; k < ke; ++k) m[j] = fmaxf(m[j], eval(k, rr));
        }
      }
      float sum[NPT], ml[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        m[j] = (m[j] == kNegInf) ? 0.f : m[j];
        ml[j] = -m[j] * kLog2e;
        sum[j] = 0.f;
#pragma unroll
        for (int d = 0; d < D0; ++d)
          if (d < sc_of(j)) {
            // x becomes exp(x - m): the posterior below is this times one factor per node
            x[sb_of(j) + d] = ex2_approx(fmaf(x[sb_of(j) + d], kLog2e, ml[j]));
            sum[j] += x[sb_of(j) + d];
          }
      }
      if (tail_any) {
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
          uint32_t kb, ke;
          tail_of(j, s_snode, kb, ke);
          for (uint32_t k = kb; k < ke; ++k) sum[j] += ex2_approx(fmaf(eval(k, rr), kLog2e, ml[j]));
        }
      }
      // posteriors in the fixed-point unit of the tile (kFixOne = 2^30):
      //   exp(x + alpha + offsets - Z) * 2^30 = exp(x - m) * exp2((alpha + offsets - Z) * log2(e) + 30 + m * log2(e)),
      // the first factor is already there from the sum, the second is one exp2 per NODE instead
      // of one per arc; an arc at -inf (padding slots) has exp(x - m) = 0, a node alpha has not
      // reached has the factor exp2(-inf) = 0
      uint32_t qstar = 0, qtot = 0;
      auto post = [&](float pf, uint32_t lab, uint32_t k) {
        if (want_gE) {
          const uint32_t q = __float2uint_rn(pf);
          qtot += q;
          if (lab == cstar) qstar += q;
          else if (q != 0u) red_add_u(grow + lab, q);
        }
        if (want_gW && pf != 0.f) sts_f(s_gw + 4u * k, lds_f(s_gw + 4u * k) + pf * (1.f / kFixOne));
      };
      float scale[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        scale[j] = ex2_approx(fmaf(pr[j] + dlt, kLog2e, 30.f) - ml[j]);
#pragma unroll
        for (int d = 0; d < D0; ++d)
          if (d < sc_of(j)) post(x[sb_of(j) + d] * scale[j], rl[sb_of(j) + d], k0r[j] + d);
      }
      if (tail_any) {
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
          uint32_t kb, ke;
          tail_of(j, s_snode, kb, ke);
          for (uint32_t k = kb; k < ke; ++k) {
            const float xv = eval(k, rr);
            post(ex2_approx(fmaf(xv, kLog2e, ml[j])) * scale[j], rr >> 16, k);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NPT; ++j)
        if (vnode[j] >= 0) sts_f(cur + voff[j], m[j] + __logf(sum[j]));
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
