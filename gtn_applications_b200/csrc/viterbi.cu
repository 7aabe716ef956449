// Tropical-semiring best path over (emissions x packed acceptor), one thread block per
// utterance — the decode counterpart of the lattice kernel: replaces
// gtn.viterbi_path(gtn.intersect(emissions, A)) (criterions/asg.py:225,
// criterions/transducer.py:215-221) without materialising the lattice.
//   delta_{t+1}[v] = max_{(u->v, c, w)} delta_t[u] + E[t,c] + w, back-pointer = the first
//   maximising in-arc in the node's in-list order (strict >, as GTN's traversal does).
#include "common.cuh"
#include "launchers.h"

namespace wfst {

struct ViterbiArgs {
  const float* E;
  int T, C;
  wfst_acceptor_batch_t g;
  int shared;
  float* scores;     // [B]
  int32_t* labels;   // [B, T] ilabel of the arc taken at each frame (-1 if no path)
  int32_t* arcs;     // [B, T] original arc index taken at each frame
  int32_t* bp;       // [B, T, stride] back-pointers (position in the in-list)
  int stride;
};

__global__ void __launch_bounds__(1024, 1) lattice_viterbi_kernel(ViterbiArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  const int gb = a.shared ? 0 : b;
  const int nb = a.g.node_offsets[gb], ab = a.g.arc_offsets[gb];
  const int N = a.g.node_offsets[gb + 1] - nb;
  const int* in_ptr = a.g.in_ptr + nb + gb;
  const int* in_src = a.g.in_src + ab;
  const int* in_label = a.g.in_label + ab;
  const int* in_arc = a.g.in_arc + ab;
  const uint8_t* flags = a.g.node_flags + nb;
  const float* w = a.g.weights ? a.g.weights + ab : nullptr;
  const float* fw = a.g.final_weights ? a.g.final_weights + nb : nullptr;
  const int T = a.T, C = a.C;
  const float* Eb = a.E + (size_t)b * T * C;
  int32_t* bp = a.bp + (size_t)b * T * a.stride;
  float* cur = sm;
  float* nxt = sm + a.stride;
  for (int v = tid; v < N; v += NT) cur[v] = (flags[v] & 1) ? 0.f : kNegInf;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const float* Et = Eb + (size_t)t * C;
    for (int v = tid; v < N; v += NT) {
      float best = kNegInf;
      int arg = -1;
      for (int k = in_ptr[v], e = in_ptr[v + 1]; k < e; ++k) {
        const float x = cur[in_src[k]] + __ldg(Et + in_label[k]) + (w ? w[in_arc[k]] : 0.f);
        if (x > best) { best = x; arg = k; }
      }
      nxt[v] = best;
      bp[(size_t)t * a.stride + v] = arg;
    }
    __syncthreads();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  if (tid == 0) {
    float best = kNegInf;
    int v = -1;
    for (int q = 0; q < N; ++q)
      if ((flags[q] & 2) && cur[q] + (fw ? fw[q] : 0.f) > best) { best = cur[q] + (fw ? fw[q] : 0.f); v = q; }
    a.scores[b] = best;
    for (int t = T - 1; t >= 0; --t) {
      int k = (v >= 0) ? bp[(size_t)t * a.stride + v] : -1;
      a.labels[(size_t)b * T + t] = (k >= 0) ? in_label[k] : -1;
      a.arcs[(size_t)b * T + t] = (k >= 0) ? in_arc[k] : -1;
      v = (k >= 0) ? in_src[k] : -1;
    }
  }
}

size_t viterbi_workspace_bytes(int B, int T, int max_nodes) {
  return align_up((size_t)B * T * ((max_nodes + 3) & ~3) * sizeof(int32_t), 256);
}

int launch_viterbi(const float* E, int B, int T, int C, const wfst_acceptor_batch_t& g, int shared,
                   float* scores, int32_t* labels, int32_t* arcs, void* workspace, cudaStream_t st) {
  ViterbiArgs a{};
  a.E = E; a.T = T; a.C = C; a.g = g; a.shared = shared;
  a.scores = scores; a.labels = labels; a.arcs = arcs;
  a.bp = (int32_t*)workspace;
  a.stride = (g.max_nodes + 3) & ~3;
  size_t smem = 2 * (size_t)a.stride * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("viterbi: acceptor with %d nodes does not fit in shared memory", g.max_nodes);
    return WFST_ERR_UNSUPPORTED;
  }
  WFST_CUDA_CHECK(cudaFuncSetAttribute(lattice_viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int nt = (g.max_nodes + 31) / 32 * 32;
  nt = nt < 64 ? 64 : (nt > 1024 ? 1024 : nt);
  lattice_viterbi_kernel<<<B, nt, smem, st>>>(a);
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

}  // namespace wfst
