// Two-block cluster lattice kernel with a WARP per node (lattice_lean.cuh: lattice_lean_pair_kernel,
// TPN = 32) for small dense packed CSR acceptors -- its own translation unit (build time).
#include "lattice_builders.cuh"

namespace wfst {

template <int NPT, bool GW>
static int launch_wpn_gw(const lean::Args& g, CsrLean::Params bp, int B, int nt, size_t smem, cudaStream_t st) {
  auto kern = lean::lattice_lean_pair_kernel<CsrLean, NPT, GW, 32>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<2 * B, nt, smem, st>>>(g, bp);      // clusters of two blocks (compile-time cluster dims)
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

template <int NPT>
static int launch_wpn_npt(const lean::Args& g, CsrLean::Params bp, int B, int nt, size_t smem, cudaStream_t st) {
  return g.want_gw ? launch_wpn_gw<NPT, true>(g, bp, B, nt, smem, st) : launch_wpn_gw<NPT, false>(g, bp, B, nt, smem, st);
}

// npt: nodes per warp (1, 2, 3, 4, 8 or 16); nt = 32 x the number of warps
int launch_lean_pair_wpn(const lean::Args& g, CsrLean::Params bp, int B, int nt, size_t smem, int npt, cudaStream_t st) {
  switch (npt) {
    case 1: return launch_wpn_npt<1>(g, bp, B, nt, smem, st);
    case 2: return launch_wpn_npt<2>(g, bp, B, nt, smem, st);
    case 3: return launch_wpn_npt<3>(g, bp, B, nt, smem, st);
    case 4: return launch_wpn_npt<4>(g, bp, B, nt, smem, st);
    case 8: return launch_wpn_npt<8>(g, bp, B, nt, smem, st);
    default: return launch_wpn_npt<16>(g, bp, B, nt, smem, st);
  }
}

}  // namespace wfst
