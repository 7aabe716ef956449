// Two-block cluster ("pair") lattice kernels (lattice_lean.cuh: lattice_lean_pair_kernel) for the
// three builders — their own translation unit so that they compile next to lattice.cu.
#include "lattice_builders.cuh"

namespace wfst {

template <class Builder, int NPT, bool GW>
static int launch_lean_pair_gw(const lean::Args& g, typename Builder::Params bp, int B, int nt, size_t smem,
                               cudaStream_t st) {
  auto kern = lean::lattice_lean_pair_kernel<Builder, NPT, GW>;
  WFST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<2 * B, nt, smem, st>>>(g, bp);      // clusters of two blocks (compile-time cluster dims)
  g_launches++;
  WFST_CUDA_CHECK(cudaGetLastError());
  return WFST_OK;
}

// arc-weight gradients are a compile-time switch of the kernel (the per-arc test in the posterior
// loop cost 3 % of the instructions of a launch that does not want them)
template <class Builder, int NPT>
static int launch_lean_pair_npt(const lean::Args& g, typename Builder::Params bp, int B, int nt, size_t smem,
                                cudaStream_t st) {
  return g.want_gw ? launch_lean_pair_gw<Builder, NPT, true>(g, bp, B, nt, smem, st)
                   : launch_lean_pair_gw<Builder, NPT, false>(g, bp, B, nt, smem, st);
}

template <class Builder>
int launch_lean_pair(const lean::Args& g, typename Builder::Params bp, int B, int nt, size_t smem, int npt,
                     cudaStream_t st) {
  switch (npt) {
    case 1: return launch_lean_pair_npt<Builder, 1>(g, bp, B, nt, smem, st);
    case 2: return launch_lean_pair_npt<Builder, 2>(g, bp, B, nt, smem, st);
    case 3: return launch_lean_pair_npt<Builder, 3>(g, bp, B, nt, smem, st);
    case 4: return launch_lean_pair_npt<Builder, 4>(g, bp, B, nt, smem, st);
    case 8: return launch_lean_pair_npt<Builder, 8>(g, bp, B, nt, smem, st);
    default: return launch_lean_pair_npt<Builder, 16>(g, bp, B, nt, smem, st);
  }
}

template int launch_lean_pair<CsrLean>(const lean::Args&, CsrLean::Params, int, int, size_t, int, cudaStream_t);
template int launch_lean_pair<CtcLean>(const lean::Args&, CtcLean::Params, int, int, size_t, int, cudaStream_t);
template int launch_lean_pair<AsgFalLean>(const lean::Args&, AsgFalLean::Params, int, int, size_t, int, cudaStream_t);

}  // namespace wfst
