// Internal host-side launchers (lattice.cu) used by the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/wfst_b200.h"

namespace wfst {
size_t lattice_hist_bytes(int B, int T, int C, int max_nodes);
// test hook: 1 = never use the shared-memory ("lean") lattice kernel; returns the old value
int lattice_force_generic(int on);
int lattice_forced_mode();
int launch_ctc(const float* E, const int* targets, const int* offsets, int B, int T, int C,
               int blank, int max_target_len, const float* grad_scale, float* scores,
               float* gradE, float* hist, const int* active, cudaStream_t st);
int launch_csr(const float* E, int T, int C, const wfst_acceptor_batch_t& g, int shared,
               const float* grad_scale, float sign, float* scores, float* gradE, int accumulate,
               float* gradW, float* hist, cudaStream_t st);
int launch_csr_cross(const float* E, int Bw, int T, int C, const wfst_acceptor_batch_t& g,
                     const float* grad_scale, float* scores, float* gradE, float* gradW, float* hist,
                     cudaStream_t st);
int launch_asg_fal(const float* E, const float* tr, const int* targets, const int* offsets, int B,
                   int T, int C, int max_target_len, const float* grad_scale, float sign,
                   float* scores, float* gradE, int accumulate, float* gradTr, float* hist,
                   cudaStream_t st, const int* active = nullptr);
// scaled-probability force-align chain (asg_fal_chain.cu); utterances it flags in *hazard_out are
// redone by launch_asg_fal(..., active = hazard)
bool asg_fal_chain_eligible(int T, int C, int max_target_len);
size_t asg_fal_chain_workspace_bytes(int B, int T, int max_target_len);
int launch_asg_fal_chain(const float* E, const float* tr, const int* targets, const int* offsets, int B,
                         int T, int C, int max_target_len, const float* grad_scale, float sign,
                         float* scores, float* gradE, float* gradTr, void* workspace, int** hazard_out,
                         cudaStream_t st);
int launch_asg_fcc(const float* E, const float* tr, int B, int T, int C, const float* grad_scale,
                   float sign, float* scores, float* gradE, int accumulate, float* gradTr,
                   float* hist, cudaStream_t st);
// dense full-connect ASG lattice, one warp per utterance (asg_dense.cu); C <= 32, T >= 1
bool asg_fcc_dense_eligible(int T, int C);
extern int g_asg_dense_single;   // test hook: 1 = never split an utterance over two warps
int launch_asg_fcc_dense(const float* E, const float* tr, int B, int T, int C, const float* grad_scale,
                         float sign, float* scores, float* gradE, int accumulate, float* gradTr,
                         float* hist, cudaStream_t st);
// best path through emissions x the bigram graph, one warp per utterance (asg_dense.cu)
bool asg_viterbi_dense_eligible(int T, int C);
int launch_asg_viterbi_dense(const float* E, const float* tr, int B, int T, int C, float* scores,
                             int32_t* labels, cudaStream_t st);
int launch_finalize(const float* za, const float* zb, float sign, int B, const float* grad_scale,
                    float* loss, float* mean_loss, cudaStream_t st);
int launch_scale(float* x, size_t n, const float* scale, cudaStream_t st);
// x[i] += y[i]
int launch_add(float* x, const float* y, size_t n, cudaStream_t st);
// fast CTC (ctc_fast.cu)
bool ctc_fast_eligible(int T, int C, int max_target_len);
size_t ctc_fast_workspace_bytes(int B, int T, int max_target_len);
int launch_ctc_fast(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                    int blank, int max_target_len, const float* grad_scale, float* z_out,
                    float* gradE, void* workspace, int** hazard_out, cudaStream_t st);
// paired fast CTC (ctc_pair.cu): two utterances per block, packed FP32 arithmetic
bool ctc_pair_eligible(int T, int C, int max_target_len);
bool ctc_pair_fused_eligible(int T, int C, int max_target_len);
size_t ctc_pair_workspace_bytes(int B, int T, int max_target_len);
// fused != 0: E holds raw logits, the kernel applies log_softmax over C itself and gradE
// receives d/d logits (criterions/ctc.py:107 followed by ctc.py:31-94)
int launch_ctc_pair(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                    int blank, int max_target_len, const float* grad_scale, float* z_out,
                    float* gradE, void* workspace, int** hazard_out, int fused, cudaStream_t st);
// chain-split scaled CTC (ctc_chain.cu): one utterance per block, both time directions packed
// in FP32 pairs, the chain split over W warps skewed by one 8-frame step
bool ctc_chain_eligible(int T, int C, int max_target_len);
int ctc_chain_force_config(int K, int W);
size_t ctc_chain_workspace_bytes(int B, int T, int max_target_len);
int launch_ctc_chain(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                     int blank, int max_target_len, const float* grad_scale, float* z_out,
                     float* gradE, void* workspace, int** hazard_out, cudaStream_t st);
// solo-chain scaled CTC (ctc_solo.cu): as ctc_chain.cu, but every time direction has its own warps
// and plain float arithmetic (shorter dependent chain per warp)
bool ctc_solo_eligible(int T, int C, int max_target_len);
int ctc_solo_force_config(int K, int W);
size_t ctc_solo_workspace_bytes(int B, int T, int max_target_len);
int launch_ctc_solo(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                    int blank, int max_target_len, const float* grad_scale, float* z_out,
                    float* gradE, void* workspace, int** hazard_out, cudaStream_t st);
// tick-scheduled chain CTC (ctc_tick.cu): ctc_solo.cu's roles and numerics, every hand-off between
// roles through one named barrier per 8-frame tick instead of per-resource mbarriers
bool ctc_tick_eligible(int T, int C, int max_target_len);
bool ctc_tick_two_per_sm(int T, int C, int max_target_len);
int ctc_tick_force_config(int K, int W);
size_t ctc_tick_workspace_bytes(int B, int T, int max_target_len);
int launch_ctc_tick(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                    int blank, int max_target_len, const float* grad_scale, float* z_out,
                    float* gradE, void* workspace, int** hazard_out, cudaStream_t st);
// float64 log-semiring CTC (ctc_exact.cu): recomputes the utterances the scaled kernels flag
bool ctc_exact_eligible(int T, int C, int max_target_len);
size_t ctc_exact_hist_bytes(int B, int T, int max_target_len);
int launch_ctc_exact(const float* E, const int* targets, const int* offsets, int B, int T, int C,
                     int blank, int max_target_len, const float* grad_scale, float* scores,
                     float* gradE, void* hist, const int* active, cudaStream_t st);
// log_softmax rows / its backward for the utterances with active[b] != 0 (lsm.cu): the
// fallback of the fused mode
int launch_lsm_rows(const float* x, const int* active, int B, int T, int C, float* out, cudaStream_t st);
int launch_lsm_backward(const float* lsm, const float* g, const int* active, int B, int T, int C,
                        float* out, cudaStream_t st);
// best path (viterbi.cu)
size_t viterbi_workspace_bytes(int B, int T, int max_nodes);
int launch_viterbi(const float* E, int B, int T, int C, const wfst_acceptor_batch_t& g, int shared,
                   float* scores, int32_t* labels, int32_t* arcs, void* workspace, cudaStream_t st);
}  // namespace wfst
