// Acceptor policies for the generic lattice kernel (lattice.cuh) and the C ABI
// entry points built on it.  See include/wfst_b200.h for the contract of each
// entry point and the reference call sites it replaces.
#include "lattice.cuh"
#include "lattice_lean.cuh"

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
// Builders for the lean kernel (lattice_lean.cuh): the same three acceptors,
// written once per utterance into shared memory as packed arc records.
// ===========================================================================
struct CsrLean {
  using Params = CsrTopo::Params;
  static constexpr int kDeg = 4;          // arcs per node held in registers; more are allowed
  static constexpr bool kTail = true;
  static constexpr bool kSort = true;     // irregular degrees: nodes handed to threads by degree
  const int* in_ptr; const int* in_src; const int* in_label; const int* in_arc;
  const int* out_ptr; const int* out_dst; const int* out_label; const int* out_arc;
  const uint8_t* flags; const float* w; float* gradW; const float* fw; float* gradF;
  int N, A, shared;
  __device__ void init(const Params& p, int b) {
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
    shared = p.shared;
  }
  __device__ int num_nodes() const { return N; }
  __device__ int num_slots() const { return A; }
  __device__ void build(const lean::Build& bd) {
    for (int v = threadIdx.x; v < N; v += blockDim.x)
      bd.node(v, in_ptr[v], in_ptr[v + 1], out_ptr[v], out_ptr[v + 1], flags[v] & 1, flags[v] & 2,
              fw ? fw[v] : 0.f);
    for (int k = threadIdx.x; k < A; k += blockDim.x) {
      const int ia = in_arc[k], oa = out_arc[k];
      bd.in_arc(k, in_src[k], in_label[k], w ? w[ia] : 0.f, ia);
      bd.out_arc(k, out_dst[k], out_label[k], w ? w[oa] : 0.f, oa);
    }
  }
  __device__ void add_final_grad(int v, float g) const {
    if (!gradF) return;
    if (shared) atomicAdd(&gradF[v], g); else gradF[v] = g;
  }
  __device__ void zero_weight_grad() const {
    if (gradW && !shared)
      for (int k = threadIdx.x; k < A; k += blockDim.x) gradW[k] = 0.f;
  }
  // every arc has exactly one slot per direction, owned by one thread: plain sums per utterance
  // (atomic: the buffer was zeroed and another block — other utterances of a shared graph, or
  // the other half of the frames in the pair kernel — adds to it as well)
  __device__ void finish(uint32_t s_gw, uint32_t s_gidx, uint32_t, float gs, int want, bool atomic) const {
    if (!want || !gradW) return;
    for (int k = threadIdx.x; k < A; k += blockDim.x) {
      const float v = lean::lds_f(s_gw + 4u * k) * gs;
      const int idx = (int)lean::lds_u(s_gidx + 4u * k);
      if (shared || atomic) { if (v != 0.f) atomicAdd(&gradW[idx], v); }
      else gradW[idx] = v;
    }
  }
};

struct CtcLean {
  using Params = CtcTopo::Params;
  static constexpr int kDeg = 3;          // self, previous, skip
  static constexpr bool kTail = false;
  static constexpr bool kSort = false;
  const int* y; int L, S, blank, C;
  __device__ void init(const Params& p, int b) {
    y = p.targets + p.offsets[b];
    L = p.offsets[b + 1] - p.offsets[b];
    S = 2 * L + 1; blank = p.blank; C = p.C;
  }
  __device__ int num_nodes() const { return S; }
  __device__ int num_slots() const { return 3 * S; }
  __device__ int lab(int s) const { return (s & 1) ? min(max(y[(s - 1) >> 1], 0), C - 1) : blank; }
  __device__ bool skip(int s) const { return (s & 1) && s > 1 && y[(s - 1) >> 1] != y[((s - 1) >> 1) - 1]; }
  __device__ void build(const lean::Build& bd) {
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      const int l = lab(s);
      uint32_t ni = 0, no = 0;
      bd.in_arc(3 * s + ni++, s, l, 0.f, -1);
      if (s > 0) bd.in_arc(3 * s + ni++, s - 1, l, 0.f, -1);
      if (skip(s)) bd.in_arc(3 * s + ni++, s - 2, l, 0.f, -1);
      bd.out_arc(3 * s + no++, s, l, 0.f, -1);
      if (s + 1 < S) bd.out_arc(3 * s + no++, s + 1, lab(s + 1), 0.f, -1);
      if (s + 2 < S && skip(s + 2)) bd.out_arc(3 * s + no++, s + 2, lab(s + 2), 0.f, -1);
      bd.node(s, 3 * s, 3 * s + ni, 3 * s, 3 * s + no, s == 0, s == S - 1 || s == S - 2, 0.f);
    }
  }
  __device__ void add_final_grad(int, float) const {}
  __device__ void zero_weight_grad() const {}
  __device__ void finish(uint32_t, uint32_t, uint32_t, float, int, bool) const {}
};

struct AsgFalLean {
  using Params = AsgFalTopo::Params;
  static constexpr int kDeg = 2;          // enter, self loop
  static constexpr bool kTail = false;
  static constexpr bool kSort = false;
  const int* y; const float* tr; float* gradTr; int L, C;
  __device__ void init(const Params& p, int b) {
    y = p.targets + p.offsets[b];
    L = p.offsets[b + 1] - p.offsets[b];
    C = p.C; tr = p.tr; gradTr = p.gradTr;
  }
  __device__ int lbl(int k) const { return min(max(y[k], 0), C - 1); }
  __device__ int num_nodes() const { return L + 1; }
  __device__ int num_slots() const { return 2 * (L + 1); }
  __device__ void build(const lean::Build& bd) {
    for (int l = threadIdx.x; l <= L; l += blockDim.x) {
      uint32_t ni = 0, no = 0;
      if (l >= 1) {
        const int cur = lbl(l - 1);
        const int enter = (l == 1) ? cur : C + cur * C + lbl(l - 2);
        const int loop = C + cur * C + cur;
        bd.in_arc(2 * l + ni++, l - 1, cur, tr[enter], enter);
        bd.in_arc(2 * l + ni++, l, cur, tr[loop], loop);
      }
      if (l < L) {
        const int nx = lbl(l);
        const int enter = (l == 0) ? nx : C + nx * C + lbl(l - 1);
        bd.out_arc(2 * l + no++, l + 1, nx, tr[enter], enter);
      }
      if (l >= 1) {
        const int cur = lbl(l - 1);
        const int loop = C + cur * C + cur;
        bd.out_arc(2 * l + no++, l, cur, tr[loop], loop);
      }
      bd.node(l, 2 * l, 2 * l + ni, 2 * l, 2 * l + no, l == 0, L > 0 && l == L, 0.f);
    }
  }
  __device__ void add_final_grad(int, float) const {}
  __device__ void zero_weight_grad() const {}   // shared by all utterances: cleared by the host
  // several arcs (and all utterances) share a transition: one atomic per arc and utterance
  __device__ void finish(uint32_t s_gw, uint32_t s_gidx, uint32_t s_node_rec, float gs, int want, bool) const {
    if (!want || !gradTr || gs == 0.f) return;
    for (int u = threadIdx.x; u <= L; u += blockDim.x) {
      const uint32_t be = lean::lds_u(s_node_rec + 4u * u);
      for (uint32_t k = be & 0xffffu; k < (be >> 16); ++k) {
        const float v = lean::lds_f(s_gw + 4u * k);
        if (v != 0.f) atomicAdd(&gradTr[lean::lds_u(s_gidx + 4u * k)], v * gs);
      }
    }
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
static int g_force_generic_lattice = 0;   // test hook: 1 = never use the lean kernels, 2 = no pair (cluster) kernel, 3 = pair kernel whenever T allows

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

template <class Builder, int NPT>
static int launch_lean_pair_npt(const lean::Args& g, typename Builder::Params bp, int B, int nt, size_t smem,
                                cudaStream_t st) {
  auto kern = lean::lattice_lean_pair_kernel<Builder, NPT>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<2 * B, nt, smem, st>>>(g, bp);      // clusters of two blocks (compile-time cluster dims)
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

template <class Builder>
static bool try_launch_lean(LatticeArgs a, typename Builder::Params bp, int B, int max_nodes, int aslots,
                            int want_gw, cudaStream_t st, int* rc) {
  if (g_force_generic_lattice == 1) return false;
  if (max_nodes < 1 || max_nodes > 16 * 1024 || aslots > 65535 || a.C > 65535) return false;
  int npt = (max_nodes + 1023) / 1024;
  npt = npt <= 4 ? npt : (npt <= 8 ? 8 : 16);
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
  lean::Args g{a, aslots, want_gw};
  // two blocks per utterance that meet in the middle (half the dependent frame steps each) when
  // there are tiles to split and the single-block launch would leave the SMs short of warps
  // (measured: B=64 x 22 warps 5.4 -> 2.9 ms; B=256 x 12 warps, already issue-bound, 1.9 -> 2.1 ms)
  const int ntiles = (a.T + kt - 1) / kt;
  const bool starved = (long long)B * (nt / 32) <= 148LL * 16;
  if (g_force_generic_lattice != 2 && ntiles >= 2 && (starved || g_force_generic_lattice == 3)) {
    switch (npt) {
      case 1: *rc = launch_lean_pair_npt<Builder, 1>(g, bp, B, nt, lay.total, st); break;
      case 2: *rc = launch_lean_pair_npt<Builder, 2>(g, bp, B, nt, lay.total, st); break;
      case 3: *rc = launch_lean_pair_npt<Builder, 3>(g, bp, B, nt, lay.total, st); break;
      case 4: *rc = launch_lean_pair_npt<Builder, 4>(g, bp, B, nt, lay.total, st); break;
      case 8: *rc = launch_lean_pair_npt<Builder, 8>(g, bp, B, nt, lay.total, st); break;
      default: *rc = launch_lean_pair_npt<Builder, 16>(g, bp, B, nt, lay.total, st); break;
    }
    return true;
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

int launch_asg_fal(const float* E, const float* tr, const int* targets, const int* offsets, int B,
                   int T, int C, int max_target_len, const float* grad_scale, float sign,
                   float* scores, float* gradE, int accumulate, float* gradTr, float* hist,
                   cudaStream_t st) {
  LatticeArgs a = base_args(E, B, T, C, grad_scale, sign, scores, gradE, accumulate, hist,
                            max_target_len + 1, gradTr ? 2 * (C + 1) * C + 2 : 0);
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
