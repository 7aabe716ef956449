// Acceptor policies of the generic lattice kernel (lattice.cuh) and builders of the lean kernels
// (lattice_lean.cuh), shared by the translation units that instantiate those kernels
// (lattice.cu: generic + single-block lean kernels; lattice_pair.cu: two-block cluster kernels —
// split so that the two compile in parallel).
#pragma once

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
    int graph_div;  // > 0 ("cross" launch, CsrLean only): item b uses graph b / graph_div, several items
                    // share a graph: gradW accumulated atomically, cleared by the host
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
    int gb = p.graph_div > 0 ? b / p.graph_div : (p.shared ? 0 : b);
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
    shared = p.shared || p.graph_div > 0;
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


// two-block cluster kernel of builder `Builder` for `npt` nodes per thread (lattice_pair.cu)
template <class Builder>
int launch_lean_pair(const lean::Args& g, typename Builder::Params bp, int B, int nt, size_t smem, int npt,
                     cudaStream_t st);

// cluster kernel with a warp per node for small dense packed CSR acceptors (lattice_pair_wpn.cu)
int launch_lean_pair_wpn(const lean::Args& g, CsrLean::Params bp, int B, int nt, size_t smem, int npt, cudaStream_t st);

// wide-register cluster kernel for degree-sorted acceptors of at most 2048 nodes (lattice_wide.cu)
int launch_lean_wide(const lean::Args& g, CsrLean::Params bp, int B, int nt, size_t smem, cudaStream_t st);

}  // namespace wfst
