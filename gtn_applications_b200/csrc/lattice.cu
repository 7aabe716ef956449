// Acceptor policies for the generic lattice kernel (lattice.cuh) and the C ABI
// entry points built on it.  See include/wfst_b200.h for the contract of each
// entry point and the reference call sites it replaces.
#include <cstdlib>
#include "lattice_builders.cuh"

#include <cstring>

namespace wfst {

// ===========================================================================
// small finishing kernels
// ===========================================================================
// loss_b = sign * (za_b [- zb_b]); mean = sum_b loss_b * grad_scale[b] in fixed order
__global__ void finalize_loss_kernel(const float* za, const float* zb, float sign, int B,
                                     const float* grad_scale, float* loss, float* mean_loss) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float l = sign * (zb ? za[b] - zb[b] : za[b]);
    if (loss) loss[b] = l;
    acc += l * (grad_scale ? grad_scale[b] : 1.f);
  }
  // fixed-shape tree: deterministic for a given B and block size
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && mean_loss) *mean_loss = v;
  }
}

__global__ void scale_inplace_kernel(float* x, size_t n, const float* scale) {
  const float s = *scale;
  if (s == 1.f) return;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] *= s;
}

__global__ void add_inplace_kernel(float* __restrict__ x, const float* __restrict__ y, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    float4* x4 = reinterpret_cast<float4*>(x);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const size_t n4 = n / 4;
    for (size_t k = i; k < n4; k += stride) {
      float4 a = x4[k];
      const float4 b = __ldcs(y4 + k);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      x4[k] = a;
    }
    for (size_t k = n4 * 4 + i; k < n; k += stride) x[k] += y[k];
  } else {
    for (size_t k = i; k < n; k += stride) x[k] += y[k];
  }
}

// ===========================================================================
// host launch helpers
// ===========================================================================
static int pick_kt(int C) {
  int kt = (16384 / (4 * C)) & ~3;
  if (kt < 4) kt = 4;
  if (kt > 64) kt = 64;
  return kt;
}
static int pick_threads(int max_nodes) {
  int nt = (max_nodes + 31) / 32 * 32;
  if (nt < 64) nt = 64;
  if (nt > 1024) nt = 1024;
  return nt;
}

template <class Topo>
static int launch_lattice(LatticeArgs a, typename Topo::Params tp, int B, int max_nodes,
                          cudaStream_t st) {
  size_t smem = lattice_smem_bytes(a.Kt, a.C, a.npad, a.extra_floats);
  if (smem > 227 * 1024) {
    set_error("lattice needs %zu bytes of shared memory per block (max 232448): "
              "C=%d nodes=%d extra=%d", smem, a.C, max_nodes, a.extra_floats);
    return WFST_ERR_UNSUPPORTED;
  }
  auto kern = lattice_fwd_bwd_kernel<Topo>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<B, pick_threads(max_nodes), smem, st>>>(a, tp);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}


// Lean kernel when the acceptor fits its limits (lattice_lean.cuh); returns false otherwise.
static int g_force_generic_lattice = 0;   // test hook: 1 = never use the lean kernels, 2 = no pair (cluster) kernel, 3 = pair kernel whenever T allows, 4 = like 3 without the wide-register variant

template <class Builder, int NPT>
static int launch_lean_npt(const lean::Args& g, typename Builder::Params bp, int B, int nt, size_t smem,
                           cudaStream_t st) {
  auto kern = lean::lattice_lean_kernel<Builder, NPT>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<B, nt, smem, st>>>(g, bp);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

template <class Builder>
static bool try_launch_lean(LatticeArgs a, typename Builder::Params bp, int B, int max_nodes, int aslots,
                            int want_gw, cudaStream_t st, int* rc, bool single_only = false) {
  if (g_force_generic_lattice == 1) return false;
  if (max_nodes < 1 || max_nodes > 16 * 1024 || aslots > 65535 || a.C >= 16 * 1024) return false;
  int npt = (max_nodes + 1023) / 1024;
  npt = npt <= 4 ? npt : (npt <= 8 ? 8 : 16);
  // degree-sorted (CSR) acceptors of 1025..2048 nodes that take the cluster kernel: four nodes per
  // thread on 512 threads with 128 registers (lattice_lean_wide.cuh); test hook 4 switches it off
  const bool wide = Builder::kSort && max_nodes > 1024 && max_nodes <= 2048 && g_force_generic_lattice != 4;
  if (wide) npt = 4;
  int nt = ((max_nodes + npt - 1) / npt + 31) / 32 * 32;
  if (nt < 64) nt = 64;
  a.npad = (max_nodes + 3) & ~3;
  if (aslots < 1) aslots = 1;
  int kt = a.Kt;
  lean::Layout lay = lean::make_layout(kt, a.C, a.npad, aslots, want_gw);
  while (lay.total > 227u * 1024u && kt > 4) {
    kt -= 4;
    lay = lean::make_layout(kt, a.C, a.npad, aslots, want_gw);
  }
  if (lay.total > 227u * 1024u) return false;
  a.Kt = kt;
  a.renorm_every = (16 + kt - 1) / kt;
  lean::Args g{a, aslots, want_gw, lay};
  // two blocks per utterance that meet in the middle (half the dependent frame steps each) when
  // there are tiles to split and the single-block launch would leave the SMs short of warps
  // (measured: B=64 x 22 warps 5.4 -> 2.9 ms; B=256 x 12 warps, already issue-bound, 1.9 -> 2.1 ms)
  const int ntiles = (a.T + kt - 1) / kt;
  const bool starved = (long long)B * (nt / 32) <= 148LL * 16;
  if constexpr (Builder::kSort) {
    // small dense acceptors (n-gram transition graphs: 84 nodes x 81 arcs): a warp per node in the
    // cluster kernel instead of a thread per node, whatever the batch size (test hooks 2 and 4:
    // off)
    const bool dense = !single_only && max_nodes <= 32 * 16 && (long long)aslots >= 16LL * max_nodes;
    if (dense && ntiles >= 2 && g_force_generic_lattice != 2 && g_force_generic_lattice != 4) {
      const int ng = max_nodes < 32 ? max_nodes : 32;          // warps = node groups
      int np = (max_nodes + ng - 1) / ng;
      np = np <= 4 ? np : (np <= 8 ? 8 : 16);
      *rc = launch_lean_pair_wpn(g, bp, B, 32 * (ng < 2 ? 2 : ng), lay.total, np, st);
      return true;
    }
  }
  if (!single_only && g_force_generic_lattice != 2 && ntiles >= 2 && (starved || g_force_generic_lattice >= 3)) {
    if constexpr (Builder::kSort) {
      if (wide) {
        *rc = launch_lean_wide(g, bp, B, nt, lay.total, st);
        return true;
      }
    }
    *rc = launch_lean_pair<Builder>(g, bp, B, nt, lay.total, npt, st);
    return true;
  }
  if (wide) {      // single-block launch: back to the usual split
    npt = (max_nodes + 1023) / 1024;
    nt = ((max_nodes + npt - 1) / npt + 31) / 32 * 32;
  }
  switch (npt) {
    case 1: *rc = launch_lean_npt<Builder, 1>(g, bp, B, nt, lay.total, st); break;
    case 2: *rc = launch_lean_npt<Builder, 2>(g, bp, B, nt, lay.total, st); break;
    case 3: *rc = launch_lean_npt<Builder, 3>(g, bp, B, nt, lay.total, st); break;
    case 4: *rc = launch_lean_npt<Builder, 4>(g, bp, B, nt, lay.total, st); break;
    case 8: *rc = launch_lean_npt<Builder, 8>(g, bp, B, nt, lay.total, st); break;
    default: *rc = launch_lean_npt<Builder, 16>(g, bp, B, nt, lay.total, st); break;
  }
  return true;
}

int lattice_force_generic(int on) { int old = g_force_generic_lattice; g_force_generic_lattice = on; return old; }
int lattice_forced_mode() { return g_force_generic_lattice; }

static size_t hist_bytes(int B, int T, int stride) {
  return align_up((size_t)B * (T + 1) * stride * sizeof(float), 256);
}


// ---------------------------------------------------------------------------
// internal launchers used by capi.cu
// ---------------------------------------------------------------------------
// sized for the smallest tile (4 frames): the lean kernel may shrink Kt to fit shared memory
static int num_tiles(int T, int C) { (void)C; return 2 * ((T + 3) / 4 + 1); }

static LatticeArgs base_args(const float* E, int B, int T, int C, const float* grad_scale, float sign,
                             float* scores, float* gradE, int accumulate, float* hist,
                             int max_nodes, int extra_floats) {
  LatticeArgs a{};
  a.E = E; a.T = T; a.C = C; a.grad_scale = grad_scale; a.sign = sign;
  a.scores = scores; a.gradE = gradE; a.accumulate = accumulate;
  a.hist = hist; a.hist_stride = (max_nodes + 3) & ~3;
  a.Kt = pick_kt(C); a.npad = (max_nodes + 3) & ~3; a.extra_floats = extra_floats;
  // the per-tile offsets live right after the history
  a.offs = reinterpret_cast<double*>(reinterpret_cast<char*>(hist) +
                                     hist_bytes(B, T, (max_nodes + 3) & ~3));
  a.offs_stride = num_tiles(T, C);
  a.renorm_every = (16 + a.Kt - 1) / a.Kt;
  return a;
}

size_t lattice_hist_bytes(int B, int T, int C, int max_nodes) {
  return hist_bytes(B, T, (max_nodes + 3) & ~3) + align_up((size_t)B * num_tiles(T, C) * sizeof(double), 256);
}

int launch_ctc(const float* E, const int* targets, const int* offsets, int B, int T, int C,
               int blank, int max_target_len, const float* grad_scale, float* scores,
               float* gradE, float* hist, const int* active, cudaStream_t st) {
  int S = 2 * max_target_len + 1;
  LatticeArgs a = base_args(E, B, T, C, grad_scale, -1.f, scores, gradE, 0, hist, S, 2 * S + 2);
  a.active = active;
  CtcTopo::Params tp{targets, offsets, blank, C};
  int rc;
  if (try_launch_lean<CtcLean>(a, tp, B, S, 3 * S, 0, st, &rc)) return rc;
  return launch_lattice<CtcTopo>(a, tp, B, S, st);
}

int launch_csr(const float* E, int T, int C, const wfst_acceptor_batch_t& g, int shared,
               const float* grad_scale, float sign, float* scores, float* gradE, int accumulate,
               float* gradW, float* hist, cudaStream_t st) {
  LatticeArgs a = base_args(E, g.B, T, C, grad_scale, sign, scores, gradE, accumulate, hist,
                            g.max_nodes, gradW ? g.max_arcs : 0);
  CsrTopo::Params tp{g, gradW, shared};
  int rc;
  if (try_launch_lean<CsrLean>(a, tp, g.B, g.max_nodes, g.max_arcs, gradW ? 1 : 0, st, &rc)) return rc;
  return launch_lattice<CsrTopo>(a, tp, g.B, g.max_nodes, st);
}

// Every emission item against every graph of a packed batch in ONE launch of the single-block lean
// kernel: item k * Bw + w scores graph k against emission item w (scores, grad_scale [K, Bw]); the
// emission gradients of all graphs are ADDED into gradE [Bw, T, C] atomically, the arc-weight
// gradients into gradW [arcs of the batch] (both cleared by the caller).  WFST_ERR_UNSUPPORTED when
// the acceptors do not fit the lean kernel (the caller then launches graph by graph).
int launch_csr_cross(const float* E, int Bw, int T, int C, const wfst_acceptor_batch_t& g,
                     const float* grad_scale, float* scores, float* gradE, float* gradW, float* hist,
                     cudaStream_t st) {
  const int K = g.B;
  const long long items = (long long)K * Bw;
  if (items > 0x7fffffffLL) return WFST_ERR_UNSUPPORTED;
  LatticeArgs a = base_args(E, (int)items, T, C, grad_scale, 1.f, scores, gradE, 1, hist,
                            g.max_nodes, gradW ? g.max_arcs : 0);
  a.e_mod = Bw;
  CsrTopo::Params tp{g, gradW, 0, Bw};
  int rc;
  if (try_launch_lean<CsrLean>(a, tp, (int)items, g.max_nodes, g.max_arcs, gradW ? 1 : 0, st, &rc, true))
    return rc;
  return WFST_ERR_UNSUPPORTED;
}

int launch_asg_fal(const float* E, const float* tr, const int* targets, const int* offsets, int B,
                   int T, int C, int max_target_len, const float* grad_scale, float sign,
                   float* scores, float* gradE, int accumulate, float* gradTr, float* hist,
                   cudaStream_t st, const int* active) {
  LatticeArgs a = base_args(E, B, T, C, grad_scale, sign, scores, gradE, accumulate, hist,
                            max_target_len + 1, gradTr ? 2 * (C + 1) * C + 2 : 0);
  a.active = active;
  AsgFalTopo::Params tp{targets, offsets, tr, gradTr, C};
  int rc;
  if (try_launch_lean<AsgFalLean>(a, tp, B, max_target_len + 1, 2 * (max_target_len + 1), gradTr ? 1 : 0, st, &rc))
    return rc;
  return launch_lattice<AsgFalTopo>(a, tp, B, max_target_len + 1, st);
}

int launch_asg_fcc(const float* E, const float* tr, int B, int T, int C, const float* grad_scale,
                   float sign, float* scores, float* gradE, int accumulate, float* gradTr,
                   float* hist, cudaStream_t st) {
  LatticeArgs a = base_args(E, B, T, C, grad_scale, sign, scores, gradE, accumulate, hist, C + 1,
                            gradTr ? (C + 1) * C : 0);
  AsgFccTopo::Params tp{tr, gradTr, C};
  return launch_lattice<AsgFccTopo>(a, tp, B, C + 1, st);
}

int launch_finalize(const float* za, const float* zb, float sign, int B, const float* grad_scale,
                    float* loss, float* mean_loss, cudaStream_t st) {
  finalize_loss_kernel<<<1, 256, 0, st>>>(za, zb, sign, B, grad_scale, loss, mean_loss);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

int launch_add(float* x, const float* y, size_t n, cudaStream_t st) {
  if (n == 0) return WFST_OK;
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  add_inplace_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, n);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

int launch_scale(float* x, size_t n, const float* scale, cudaStream_t st) {
  if (n == 0) return WFST_OK;
  size_t blocks = (n + 4 * 256 - 1) / (4 * 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  scale_inplace_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, scale);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

}  // namespace wfst
