// Time-synchronous log-semiring forward/backward over (emissions x acceptor),
// one thread block per utterance.  This is the generic skeleton: what varies
// between criteria is only the acceptor ("Topo" policy) — a packed CSR graph, the
// closed-form CTC chain, the ASG force-alignment chain or the ASG full-connect
// graph.  It computes what the reference gets from
//   forward_score(intersect(linear_graph(T,C) with weights E_b, A_b))  + gtn.backward
// (criterions/ctc.py:49-51,78; asg.py:111-115,158; stc.py:85-86,113;
// transducer.py:283-290,321) without materialising the composed lattice:
//   alpha_0[q] = 0 (start) / -inf;  alpha_{t+1}[v] = LSE_{u->v} alpha_t[u] + E[t,c] + w
//   Z = LSE_{accept} alpha_T;  beta symmetric;
//   dZ/dE[t,c] = sum_{arcs labelled c} exp(alpha_t[u] + E[t,c] + w + beta_{t+1}[v] - Z)
//
// Data movement: the [T,C] emissions of the utterance are streamed HBM -> shared
// memory in tiles of Kt frames with 1-D bulk async copies (TMA engine, mbarrier
// completion), double buffered; alpha_t is kept in shared memory and spilled to a
// [T+1, N] history in HBM for the backward sweep; gradient rows are accumulated in a
// shared-memory tile and written back with a bulk async store (or coalesced
// stores when the tile is not 16-byte aligned / in accumulate mode).
#pragma once

#include "common.cuh"

namespace wfst {

constexpr float kFixOne = 1073741824.f;   // 2^30: fixed-point unit of the posterior tile

struct LatticeArgs {
  const float* E;        // [B, T, C]
  int T, C;
  const float* grad_scale;  // [B] or null
  float sign;            // multiplies grad_scale (e.g. -1 for loss = -Z)
  float* scores;         // [B] Z_b
  float* gradE;          // [B, T, C] or null
  int accumulate;        // add into gradE instead of overwriting
  float* hist;           // [B, T+1, hist_stride] alpha history
  int hist_stride;       // >= max nodes
  double* offs;          // [B, offs_stride] per tile: {cumulative alpha offset, per-frame drift}
  int offs_stride;       // >= 2 * number of tiles
  int renorm_every;      // renormalise alpha/beta every this many tiles
  int Kt;                // frames per tile (multiple of 4)
  int npad;              // smem floats reserved per alpha buffer
  int extra_floats;      // policy-owned smem floats (after the fixed regions)
  const int* active;     // optional [B]: blocks with active[b] == 0 return at once
  int e_mod;             // > 0 ("cross" launch of the single-block lean kernel): item b reads emission
                         // item b % e_mod and ADDS its gradient there atomically (several items share it)
};

// shared memory carve-up (all float-sized slots; base is 16B aligned)
struct SmemLayout {
  float* alpha0;
  float* alpha1;
  float* tile0;
  float* tile1;
  float* gtile;
  float* extra;
  uint64_t* bars;
  float* red;
};

__device__ __forceinline__ SmemLayout carve(float* base, const LatticeArgs& a) {
  SmemLayout s;
  size_t tile = (size_t)a.Kt * a.C;
  tile = (tile + 3) & ~(size_t)3;
  s.bars = reinterpret_cast<uint64_t*>(base);  // 2 barriers = 4 floats
  s.red = base + 4;                            // 36 floats (warp partials + broadcast)
  s.tile0 = base + 40;
  s.tile1 = s.tile0 + tile;
  s.gtile = s.tile1 + tile;
  s.alpha0 = s.gtile + tile;
  s.alpha1 = s.alpha0 + a.npad;
  s.extra = s.alpha1 + a.npad;
  return s;
}

inline size_t lattice_smem_bytes(int Kt, int C, int npad, int extra_floats) {
  size_t tile = ((size_t)Kt * C + 3) & ~(size_t)3;
  return sizeof(float) * (40 + 3 * tile + 2 * (size_t)npad + (size_t)extra_floats);
}

// block-wide LSE of per-thread partial (m, s) pairs
__device__ __forceinline__ float block_lse(float v, float* red) {
  // v is a per-thread log value (-inf allowed)
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  float m = warp_max(v);
  if (lane == 0) red[w] = m;
  __syncthreads();
  float bm = (lane < nw) ? red[lane] : kNegInf;
  bm = warp_max(bm);
  __syncthreads();
  float e = (bm == kNegInf || v == kNegInf) ? 0.f : __expf(v - bm);
  e = warp_sum(e);
  if (lane == 0) red[w] = e;
  __syncthreads();
  float bs = (lane < nw) ? red[lane] : 0.f;
  bs = warp_sum(bs);
  __syncthreads();
  return (bm == kNegInf) ? kNegInf : bm + __logf(bs);
}

// block-wide max (all threads get the result)
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  float m = warp_max(v);
  if (lane == 0) red[w] = m;
  __syncthreads();
  float bm = (lane < nw) ? red[lane] : kNegInf;
  bm = warp_max(bm);
  __syncthreads();
  return bm;
}

template <class Topo>
__global__ void __launch_bounds__(1024, 1) lattice_fwd_bwd_kernel(LatticeArgs a, typename Topo::Params tp) {
  extern __shared__ __align__(16) float smem_raw[];
  const int b = blockIdx.x;
  if (a.active && a.active[b] == 0) return;
  const int tid = threadIdx.x, NT = blockDim.x;
  SmemLayout sm = carve(smem_raw, a);
  Topo topo;
  topo.init(tp, b, sm.extra);  // may use all threads; ends with __syncthreads
  const int N = topo.num_nodes();
  const int T = a.T, C = a.C, Kt = a.Kt;
  const float* Eb = a.E + (size_t)b * T * C;
  float* hist = a.hist + (size_t)b * (T + 1) * a.hist_stride;
  double* offA = a.offs + (size_t)b * a.offs_stride;
  const int ntiles = (T + Kt - 1) / Kt;

  if (tid == 0) {
    mbar_init(&sm.bars[0], 1);
    mbar_init(&sm.bars[1], 1);
    fence_barrier_init();
  }
  uint32_t phase[2] = {0u, 0u};
  float* tiles[2] = {sm.tile0, sm.tile1};
  __syncthreads();

  auto tile_rows = [&](int i) { return min(Kt, T - i * Kt); };
  auto tile_tma_ok = [&](int i) {
    const float* src = Eb + (size_t)i * Kt * C;
    uint32_t bytes = (uint32_t)tile_rows(i) * C * 4u;
    return ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15u) == 0);
  };
  // all threads call; returns whether the consumer must wait on the mbarrier
  auto issue_tile = [&](int i, int buf) {
    const float* src = Eb + (size_t)i * Kt * C;
    int n = tile_rows(i) * C;
    if (tile_tma_ok(i)) {
      if (tid == 0) {
        mbar_expect_tx(&sm.bars[buf], (uint32_t)n * 4u);
        bulk_g2s(tiles[buf], src, (uint32_t)n * 4u, &sm.bars[buf]);
      }
    } else {
      for (int k = tid; k < n; k += NT) tiles[buf][k] = __ldg(src + k);
    }
  };
  auto wait_tile = [&](int i, int buf) {
    if (tile_tma_ok(i)) {
      mbar_wait(&sm.bars[buf], phase[buf]);
      phase[buf] ^= 1u;
    }
  };

  // ------------------------------------------------------------- forward
  float* cur = sm.alpha0;
  float* nxt = sm.alpha1;
  double cumA = 0.0;  // identical in every thread
  float dA = 0.f;     // per-frame drift compensation, re-estimated at every renormalisation
  for (int v = tid; v < N; v += NT) {
    float x = topo.is_start(v) ? 0.f : kNegInf;
    cur[v] = x;
    hist[v] = x;
  }
  if (ntiles > 0) issue_tile(0, 0);
  __syncthreads();
  for (int i = 0; i < ntiles; ++i) {
    const int buf = i & 1;
    wait_tile(i, buf);
    if (i + 1 < ntiles) issue_tile(i + 1, buf ^ 1);
    const int rows = tile_rows(i);
    // Renormalisation keeps |alpha| bounded for any T: every `renorm_every` tiles alpha
    // is re-centred on its maximum; the removed offset is carried in float64.  (dA/dB are
    // a per-frame drift term, currently always 0: measured on B200 it did not improve
    // accuracy because the posterior-relevant states are not the ones near the maximum.)  Hist row t (t >= 1), written at in-tile step k of tile i, is relative to
    // offA[2i] + (k+1) * offA[2i+1].
    if (i % a.renorm_every == 0) {
      float pm = kNegInf;
      for (int v = tid; v < N; v += NT) pm = fmaxf(pm, cur[v]);
      const float mx = block_max(pm, sm.red);
      if (mx != kNegInf && mx != -kNegInf && mx == mx) {
        for (int v = tid; v < N; v += NT) cur[v] -= mx;
        cumA += (double)mx;
      }
      __syncthreads();
    }
    if (tid == 0) { offA[2 * i] = cumA; offA[2 * i + 1] = (double)dA; }
    for (int tt = 0; tt < rows; ++tt) {
      const float* Et = tiles[buf] + tt * C;
      const int t = i * Kt + tt;
      float* hrow = hist + (size_t)(t + 1) * a.hist_stride;
      for (int v = tid; v < N; v += NT) {
        float m = kNegInf;
        topo.in_arcs(v, [&](int u, int lab, float w, int) {
          m = fmaxf(m, cur[u] + Et[lab] + w);
        });
        float r = kNegInf;
        if (m != kNegInf) {
          float s = 0.f;
          topo.in_arcs(v, [&](int u, int lab, float w, int) {
            s += __expf(cur[u] + Et[lab] + w - m);
          });
          r = m + __logf(s) - dA;
        }
        nxt[v] = r;
        hrow[v] = r;
      }
      cumA += (double)dA;
      __syncthreads();
      float* tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
  }

  // ------------------------------------------------------------- Z
  float part = kNegInf;
  for (int v = tid; v < N; v += NT)
    if (topo.is_accept(v)) part = log_add(part, cur[v] + topo.final_w(v));
  const float Zn = block_lse(part, sm.red);   // relative to cumA
  const double Zd = (double)Zn + cumA;
  const float Z = (float)Zd;
  if (tid == 0) a.scores[b] = Z;
  const bool want_gE = a.gradE != nullptr;
  const bool want_gW = topo.wants_weight_grad();
  if (!want_gE && !want_gW) return;
  const float gs = a.sign * (a.grad_scale ? a.grad_scale[b] : 1.f);
  float* gEb = want_gE ? a.gradE + (size_t)b * T * C : nullptr;
  const bool feasible = (Z != kNegInf) && (Z == Z) && (Z != -kNegInf);
  if (!feasible) {
    // no accepting path: gradient defined as zero (documented deviation: GTN yields NaNs)
    if (want_gE && !a.accumulate)
      for (size_t k = tid; k < (size_t)T * C; k += NT) gEb[k] = 0.f;
    topo.finish_weight_grad(0.f);
    return;
  }

  // ------------------------------------------------------------- backward
  // beta_{t+1} lives in `nxt`, beta_t is written to `cur`
  __syncthreads();
  for (int v = tid; v < N; v += NT) {
    const bool acc = topo.is_accept(v);
    nxt[v] = acc ? topo.final_w(v) : kNegInf;
    // posterior of ending in v = d Z / d final weight of v
    if (acc && cur[v] != kNegInf) topo.add_final_grad(v, __expf(cur[v] + topo.final_w(v) - Zn) * gs);
  }
  double cumB = 0.0;
  float dB = 0.f;
  if (ntiles > 0) issue_tile(ntiles - 1, (ntiles - 1) & 1);
  __syncthreads();
  for (int i = ntiles - 1; i >= 0; --i) {
    const int buf = i & 1;
    wait_tile(i, buf);
    if (i > 0) issue_tile(i - 1, buf ^ 1);
    const int rows = tile_rows(i);
    float* gt = sm.gtile;
    if (want_gE) {
      for (int k = tid; k < rows * C; k += NT) gt[k] = 0.f;
    }
    if ((ntiles - 1 - i) % a.renorm_every == 0) {
      float pm = kNegInf;
      for (int v = tid; v < N; v += NT) pm = fmaxf(pm, nxt[v]);
      const float mx = block_max(pm, sm.red);
      if (mx != kNegInf && mx != -kNegInf && mx == mx) {
        for (int v = tid; v < N; v += NT) nxt[v] -= mx;
        cumB += (double)mx;
      }
    }
    __syncthreads();
    // offset of alpha row t = i*Kt + tt: row i*Kt is the last row of tile i-1
    const double oa_first = (i > 0) ? offA[2 * (i - 1)] + (double)Kt * offA[2 * (i - 1) + 1] : 0.0;
    const double oa_base = offA[2 * i], oa_step = offA[2 * i + 1];
    for (int tt = rows - 1; tt >= 0; --tt) {
      const float* Et = tiles[buf] + tt * C;
      const int t = i * Kt + tt;
      const float* hrow = hist + (size_t)t * a.hist_stride;
      float* grow = gt + tt * C;
      // log-normaliser of this frame's posteriors, exact in float64
      const float dlt = (float)(((tt == 0) ? oa_first : oa_base + (double)tt * oa_step) + cumB - Zd);
      for (int u = tid; u < N; u += NT) {
        const float au = hrow[u];
        float m = kNegInf;
        topo.out_arcs(u, [&](int v, int lab, float w, int) {
          m = fmaxf(m, Et[lab] + w + nxt[v]);
        });
        float r = kNegInf;
        if (m != kNegInf) {
          float s = 0.f;
          const bool live = au != kNegInf;
          const float off = au + dlt;
          topo.out_arcs(u, [&](int v, int lab, float w, int arc) {
            float x = Et[lab] + w + nxt[v];
            s += __expf(x - m);
            if (live && x != kNegInf) {
              float p = __expf(x + off);
              // posteriors are accumulated in 2^-30 fixed point: shared-memory integer atomics
              // are native (float ones are compare-and-swap loops) and the sum does not depend
              // on the order of the arcs
              if (want_gE) atomicAdd(reinterpret_cast<unsigned int*>(grow) + lab, __float2uint_rn(p * kFixOne));
              topo.add_weight_grad(arc, p);
            }
          });
          r = m + __logf(s) - dB;
        }
        cur[u] = r;
      }
      cumB += (double)dB;
      __syncthreads();
      float* tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
    if (want_gE) {
      // Every accepting path crosses each frame exactly once, so the posteriors of a
      // frame sum to one.  Normalising each row by its own sum removes the common-mode
      // rounding error that alpha_t + beta_t - Z has picked up over the T steps.
      const int lane = tid & 31, wid = tid >> 5, nw = NT >> 5;
      for (int r = wid; r < rows; r += nw) {
        float* row = gt + r * C;
        const unsigned int* rowu = reinterpret_cast<const unsigned int*>(row);
        float rs = 0.f;
        for (int c = lane; c < C; c += 32) rs += (float)rowu[c];
        rs = warp_sum(rs);
        const float f = (rs > 0.f) ? gs / rs : 0.f;     // the fixed-point scale cancels
        for (int c = lane; c < C; c += 32) row[c] = (float)rowu[c] * f;
      }
      __syncthreads();
      float* dst = gEb + (size_t)i * Kt * C;
      const int n = rows * C;
      const bool tma = !a.accumulate && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n & 3) == 0);
      if (tma) {
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
          bulk_s2g(dst, gt, (uint32_t)n * 4u);
          bulk_commit();
          bulk_wait_read<0>();  // gtile is re-zeroed by the next tile
        }
      } else if (a.accumulate) {
        for (int k = tid; k < n; k += NT) dst[k] += gt[k];
      } else {
        for (int k = tid; k < n; k += NT) dst[k] = gt[k];
      }
      __syncthreads();
    }
  }
  if (tid == 0) bulk_wait_all<0>();
  topo.finish_weight_grad(gs);
}

}  // namespace wfst
