// Acceptor policies for the generic lattice kernel (lattice.cuh) and the C ABI
// entry points built on it.  See include/wfst_b200.h for the contract of each
// entry point and the reference call sites it replaces.
#include "lattice.cuh"

#include <cstring>

namespace wfst {

// ===========================================================================
// Packed CSR acceptor (STC, transducer alignments, anything built on the host)
// ===========================================================================
struct CsrTopo {
  struct Params {
    wfst_acceptor_batch_t g;
    float* gradW;   // [arcs] or null
    int shared;     // all utterances use graph 0; gradW accumulated atomically
  };
  const int* in_ptr; const int* in_src; const int* in_label; const int* in_arc;
  const int* out_ptr; const int* out_dst; const int* out_label; const int* out_arc;
  const uint8_t* flags; const float* w; float* gw; float* gradW;
  const float* fw; float* gradF;
  int N, A, shared;

  __device__ void init(const Params& p, int b, float* extra) {
    int gb = p.shared ? 0 : b;
    int nb = p.g.node_offsets[gb], ab = p.g.arc_offsets[gb];
    N = p.g.node_offsets[gb + 1] - nb;
    A = p.g.arc_offsets[gb + 1] - ab;
    in_ptr = p.g.in_ptr + nb + gb;  out_ptr = p.g.out_ptr + nb + gb;
    in_src = p.g.in_src + ab;  in_label = p.g.in_label + ab;  in_arc = p.g.in_arc + ab;
    out_dst = p.g.out_dst + ab; out_label = p.g.out_label + ab; out_arc = p.g.out_arc + ab;
    flags = p.g.node_flags + nb;
    w = p.g.weights ? p.g.weights + ab : nullptr;
    gradW = p.gradW ? p.gradW + ab : nullptr;
    fw = p.g.final_weights ? p.g.final_weights + nb : nullptr;
    gradF = p.g.grad_final_weights ? p.g.grad_final_weights + nb : nullptr;
    gw = extra;
    shared = p.shared;
    if (gradW) for (int k = threadIdx.x; k < A; k += blockDim.x) gw[k] = 0.f;
    __syncthreads();
  }
  __device__ int num_nodes() const { return N; }
  __device__ bool is_start(int v) const { return flags[v] & 1; }
  __device__ bool is_accept(int v) const { return flags[v] & 2; }
  __device__ float final_w(int v) const { return fw ? fw[v] : 0.f; }
  __device__ void add_final_grad(int v, float g) const {
    if (!gradF) return;
    if (shared) atomicAdd(&gradF[v], g); else gradF[v] = g;
  }
  __device__ bool wants_weight_grad() const { return gradW != nullptr; }
  template <class F>
  __device__ void in_arcs(int v, F f) const {
    for (int k = in_ptr[v], e = in_ptr[v + 1]; k < e; ++k) {
      int arc = in_arc[k];
      f(in_src[k], in_label[k], w ? w[arc] : 0.f, arc);
    }
  }
  template <class F>
  __device__ void out_arcs(int u, F f) const {
    for (int k = out_ptr[u], e = out_ptr[u + 1]; k < e; ++k) {
      int arc = out_arc[k];
      f(out_dst[k], out_label[k], w ? w[arc] : 0.f, arc);
    }
  }
  // every arc is owned by the thread that owns its source node: no atomics
  __device__ void add_weight_grad(int arc, float p) const { if (gradW) gw[arc] += p; }
  __device__ void finish_weight_grad(float gs) const {
    if (!gradW) return;
    __syncthreads();
    for (int k = threadIdx.x; k < A; k += blockDim.x) {
      if (shared) { if (gw[k] != 0.f) atomicAdd(&gradW[k], gw[k] * gs); }
      else gradW[k] = gw[k] * gs;
    }
  }
};

// ===========================================================================
// CTC chain in closed form (criterions/ctc.py:15-29): states s in [0, 2L],
// label(s) = blank (s even) / y[(s-1)/2]; in-arcs of s: s, s-1, and s-2 when s
// is odd, s > 1 and y differs from the previous label.  Start {0}; accept
// {2L, 2L-1}.
// ===========================================================================
struct CtcTopo {
  struct Params { const int* targets; const int* offsets; int blank; int C; };
  int* lab; int* skip; int S;
  __device__ void init(const Params& p, int b, float* extra) {
    const int* y = p.targets + p.offsets[b];
    int L = p.offsets[b + 1] - p.offsets[b];
    S = 2 * L + 1;
    lab = reinterpret_cast<int*>(extra);
    skip = lab + S;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      int k = (s - 1) >> 1;
      int l = (s & 1) ? y[k] : p.blank;
      l = min(max(l, 0), p.C - 1);
      lab[s] = l;
      skip[s] = ((s & 1) && s > 1 && y[k] != y[k - 1]) ? 1 : 0;
    }
    __syncthreads();
  }
  __device__ int num_nodes() const { return S; }
  __device__ bool is_start(int v) const { return v == 0; }
  __device__ bool is_accept(int v) const { return v == S - 1 || v == S - 2; }
  __device__ float final_w(int) const { return 0.f; }
  __device__ void add_final_grad(int, float) const {}
  __device__ bool wants_weight_grad() const { return false; }
  template <class F>
  __device__ void in_arcs(int s, F f) const {
    int l = lab[s];
    f(s, l, 0.f, -1);
    if (s > 0) f(s - 1, l, 0.f, -1);
    if (skip[s]) f(s - 2, l, 0.f, -1);
  }
  template <class F>
  __device__ void out_arcs(int u, F f) const {
    f(u, lab[u], 0.f, -1);
    if (u + 1 < S) f(u + 1, lab[u + 1], 0.f, -1);
    if (u + 2 < S && skip[u + 2]) f(u + 2, lab[u + 2], 0.f, -1);
  }
  __device__ void add_weight_grad(int, float) const {}
  __device__ void finish_weight_grad(float) const {}
};

// ===========================================================================
// ASG force-alignment o transitions (criterions/asg.py:71-81,111-113): nodes
// 0..L (0 start, L accept when L > 0); arc (l-1 -> l) and self loop (l -> l) both
// labelled y_l, weighted with the transition into y_l from the previous label
// (or from <s>).  transitions layout: asg.py:53-69.
// ===========================================================================
struct AsgFalTopo {
  struct Params { const int* targets; const int* offsets; const float* tr; float* gradTr; int C; };
  const int* y; const float* tr; float* gtr; float* gradTr; int L, C;
  __device__ void init(const Params& p, int b, float* extra) {
    y = p.targets + p.offsets[b];
    L = p.offsets[b + 1] - p.offsets[b];
    C = p.C; tr = p.tr; gradTr = p.gradTr;
    gtr = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(extra) + 7) & ~(uintptr_t)7);   // 64-bit slots
    if (gradTr) for (int k = threadIdx.x; k < 2 * (C + 1) * C; k += blockDim.x) gtr[k] = 0.f;
    __syncthreads();
  }
  __device__ int lbl(int k) const { return min(max(y[k], 0), C - 1); }
  __device__ int num_nodes() const { return L + 1; }
  __device__ bool is_start(int v) const { return v == 0; }
  __device__ bool is_accept(int v) const { return L > 0 && v == L; }
  __device__ float final_w(int) const { return 0.f; }
  __device__ void add_final_grad(int, float) const {}
  __device__ bool wants_weight_grad() const { return gradTr != nullptr; }
  template <class F>
  __device__ void in_arcs(int l, F f) const {
    if (l == 0) return;
    int cur = lbl(l - 1);
    int enter = (l == 1) ? cur : C + cur * C + lbl(l - 2);
    int loop = C + cur * C + cur;
    f(l - 1, cur, tr[enter], enter);
    f(l, cur, tr[loop], loop);
  }
  template <class F>
  __device__ void out_arcs(int u, F f) const {
    if (u < L) {
      int nx = lbl(u);
      int enter = (u == 0) ? nx : C + nx * C + lbl(u - 1);
      f(u + 1, nx, tr[enter], enter);
    }
    if (u >= 1) {
      int cur = lbl(u - 1);
      int loop = C + cur * C + cur;
      f(u, cur, tr[loop], loop);
    }
  }
  // several nodes share a transition: accumulate in 2^-32 fixed point (64-bit shared-memory
  // integer atomics are native; float ones are compare-and-swap loops) -- also order-independent
  __device__ void add_weight_grad(int idx, float p) const {
    if (gradTr) atomicAdd(reinterpret_cast<unsigned long long*>(gtr) + idx, (unsigned long long)__float2ull_rn(p * 4294967296.f));
  }
  __device__ void finish_weight_grad(float gs) const {
    if (!gradTr) return;
    __syncthreads();
    const unsigned long long* g64 = reinterpret_cast<const unsigned long long*>(gtr);
    for (int k = threadIdx.x; k < (C + 1) * C; k += blockDim.x)
      if (g64[k] != 0ull) atomicAdd(&gradTr[k], (float)((double)g64[k] * (1.0 / 4294967296.0)) * gs);
  }
};

// ===========================================================================
// ASG full-connect graph (criterions/asg.py:53-69,114): node 0 start, nodes
// 1..C accept; (0 -> i+1, label i, tr[0,i]); (j+1 -> i+1, label i, tr[1+i, j]).
// ===========================================================================
struct AsgFccTopo {
  struct Params { const float* tr; float* gradTr; int C; };
  const float* tr; float* gtr; float* gradTr; int C;
  __device__ void init(const Params& p, int, float* extra) {
    C = p.C; tr = p.tr; gradTr = p.gradTr; gtr = extra;
    if (gradTr) for (int k = threadIdx.x; k < (C + 1) * C; k += blockDim.x) gtr[k] = 0.f;
    __syncthreads();
  }
  __device__ int num_nodes() const { return C + 1; }
  __device__ bool is_start(int v) const { return v == 0; }
  __device__ bool is_accept(int v) const { return v > 0; }
  __device__ float final_w(int) const { return 0.f; }
  __device__ void add_final_grad(int, float) const {}
  __device__ bool wants_weight_grad() const { return gradTr != nullptr; }
  template <class F>
  __device__ void in_arcs(int v, F f) const {
    if (v == 0) return;
    int i = v - 1;
    f(0, i, tr[i], i);
    for (int j = 0; j < C; ++j) f(j + 1, i, tr[C + i * C + j], C + i * C + j);
  }
  template <class F>
  __device__ void out_arcs(int u, F f) const {
    if (u == 0) {
      for (int i = 0; i < C; ++i) f(i + 1, i, tr[i], i);
    } else {
      int j = u - 1;
      for (int i = 0; i < C; ++i) f(i + 1, i, tr[C + i * C + j], C + i * C + j);
    }
  }
  // arc (u -> *) is owned by the thread that owns u: plain accumulation
  __device__ void add_weight_grad(int idx, float p) const { if (gradTr) gtr[idx] += p; }
  __device__ void finish_weight_grad(float gs) const {
    if (!gradTr) return;
    __syncthreads();
    for (int k = threadIdx.x; k < (C + 1) * C; k += blockDim.x)
      if (gtr[k] != 0.f) atomicAdd(&gradTr[k], gtr[k] * gs);
  }
};

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

static size_t hist_bytes(int B, int T, int stride) {
  return align_up((size_t)B * (T + 1) * stride * sizeof(float), 256);
}


// ---------------------------------------------------------------------------
// internal launchers used by capi.cu
// ---------------------------------------------------------------------------
static int num_tiles(int T, int C) { int kt = pick_kt(C); return 2 * ((T + kt - 1) / kt + 1); }

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
  return launch_lattice<CtcTopo>(a, tp, B, S, st);
}

int launch_csr(const float* E, int T, int C, const wfst_acceptor_batch_t& g, int shared,
               const float* grad_scale, float sign, float* scores, float* gradE, int accumulate,
               float* gradW, float* hist, cudaStream_t st) {
  LatticeArgs a = base_args(E, g.B, T, C, grad_scale, sign, scores, gradE, accumulate, hist,
                            g.max_nodes, gradW ? g.max_arcs : 0);
  CsrTopo::Params tp{g, gradW, shared};
  return launch_lattice<CsrTopo>(a, tp, g.B, g.max_nodes, st);
}

int launch_asg_fal(const float* E, const float* tr, const int* targets, const int* offsets, int B,
                   int T, int C, int max_target_len, const float* grad_scale, float sign,
                   float* scores, float* gradE, int accumulate, float* gradTr, float* hist,
                   cudaStream_t st) {
  LatticeArgs a = base_args(E, B, T, C, grad_scale, sign, scores, gradE, accumulate, hist,
                            max_target_len + 1, gradTr ? 2 * (C + 1) * C + 2 : 0);
  AsgFalTopo::Params tp{targets, offsets, tr, gradTr, C};
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
