// Wide-register two-block cluster kernel (lattice_lean_wide.cuh) for packed CSR acceptors -- its own
// translation unit so that it compiles next to lattice.cu and lattice_pair.cu.
#include "lattice_builders.cuh"
#include "lattice_lean_wide.cuh"

namespace wfst {

template <bool GW>
static int launch_wide_gw(const lean::Args& g, CsrLean::Params bp, int B, int nt, size_t smem, cudaStream_t st) {
  auto kern = lean::lattice_lean_wide_kernel<CsrLean, GW>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<2 * B, nt, smem, st>>>(g, bp);      // clusters of two blocks (compile-time cluster dims)
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

int launch_lean_wide(const lean::Args& g, CsrLean::Params bp, int B, int nt, size_t smem, cudaStream_t st) {
  return g.want_gw ? launch_wide_gw<true>(g, bp, B, nt, smem, st) : launch_wide_gw<false>(g, bp, B, nt, smem, st);
}

}  // namespace wfst
