/*
 * wfst_b200.h — C ABI of libwfst_b200.so, the B200 (sm_100a) implementation of the
 * WFST sequence-criterion hot path of facebookresearch/gtn_applications.
 *
 * What it replaces.  In the reference every criterion computes, per utterance b,
 *     Z_b = forward_score(intersect(linear_graph(T, C) with weights E[b], A_b))
 * on the CPU inside the external GTN library and then calls gtn.backward to get
 * dZ_b/dE[b] (and dZ_b/d arc weights of A_b):
 *     criterions/ctc.py:40-51,78-81        (CTC: A_b = create_ctc_graph, ctc.py:15-29)
 *     criterions/asg.py:96-115,158-168     (ASG: force-align o transitions, and transitions)
 *     criterions/stc.py:74-86,113-115      (STC: create_stc_graph, stc.py:22-64)
 *     criterions/transducer.py:262-290,320-336 (alignment graph, optional transitions)
 * The entry points below are what a binding of that path would call instead of
 * gtn.linear_graph / set_weights / intersect / forward_score / backward: the
 * composed lattice is never materialised, the time-synchronous recursion
 *     alpha_{t+1}[v] = LSE_{(u->v, c, w) in A_b} alpha_t[u] + E[b,t,c] + w
 * and its reverse are run on the GPU, one thread block per utterance.
 *
 * Conventions.  Plain pointers and sizes only; no torch / C++ types.  Unless a
 * function name ends in _host, every pointer is a DEVICE pointer on the current
 * CUDA device and `stream` is a cudaStream_t passed as void* (NULL = the legacy
 * default stream); calls are asynchronous with respect to the host.  All scores
 * are float32, all indices int32, epsilon is not allowed in acceptors handed to
 * the lattice kernels.  Every function returns WFST_OK (0) or a negative error
 * code; wfst_last_error() returns a thread-local message for the last failure.
 * There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef WFST_B200_H
#define WFST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFST_OK 0
#define WFST_ERR_INVALID (-1)     /* bad argument (shape, null pointer, label out of range) */
#define WFST_ERR_CUDA (-2)        /* a CUDA runtime call failed */
#define WFST_ERR_UNSUPPORTED (-3) /* shape outside what the kernels handle */
#define WFST_ERR_WORKSPACE (-4)   /* workspace too small */

#define WFST_ABI_VERSION 2

#if defined(__GNUC__)
#define WFST_API __attribute__((visibility("default")))
#else
#define WFST_API
#endif

/* thread-local message of the last error returned on this thread ("" if none) */
WFST_API const char* wfst_last_error(void);
WFST_API int wfst_abi_version(void);
/* test hook: 1 = run CTC on the log-semiring lattice kernel only (and ASG full-connect on the
 * generic lattice kernel), 2 = skip the paired scaled-probability kernel (use the
 * single-utterance one), 3 = dense ASG full-connect kernel with one warp per utterance only
 * (no two-warp split), 4 = no chain-split CTC kernel, 5 = chain-split CTC kernel first
 * (default: the paired kernel where it is eligible, the chain-split one otherwise),
 * 6 = solo-chain CTC kernel first (csrc/ctc_solo.cu; never selected by default),
 * 0 = default (returns the old value) */
WFST_API int wfst_debug_force_generic_ctc(int on);
/* test / tuning hook: the chain-split CTC kernel (csrc/ctc_chain.cu) prefers the configuration
 * (K slots per lane, W warps per chain) for targets it can hold; (0, 0) = automatic */
WFST_API int wfst_debug_ctc_chain_config(int K, int W);
/* test hook: 1 = the acceptor lattice entry points (CSR, ASG force-align, CTC fallback) use the
 * generic global-memory kernel even when the acceptor fits the shared-memory ("lean") kernels,
 * 2 = the single-block lean kernel only (no two-block cluster kernel), 3 = the two-block
 * cluster kernel whenever T allows (by default only when the single-block launch would leave the
 * SMs short of warps), 4 = like 3 but without the wide-register variant of the cluster kernel
 * (packed CSR acceptors of 1025..2048 nodes, csrc/lattice_lean_wide.cuh), 0 = default (returns
 * the old value) */
WFST_API int wfst_debug_force_generic_lattice(int on);
/* test hook: copies the per-utterance fallback flags of the last CTC call that used
 * `workspace` to the host (1 = recomputed by the log-semiring kernel, -1 = fast path not used) */
WFST_API int wfst_debug_ctc_hazards(const void* workspace, int B, int T, int C, int max_target_len,
                                    int32_t* host_flags);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
WFST_API unsigned long long wfst_launch_count(void);

/* ------------------------------------------------------------------------
 * CTC — replaces CTCLossFunction.forward + .backward (criterions/ctc.py:31-94):
 * per-utterance create_ctc_graph (built on the device from the target), the
 * intersect with the emissions graph, forward_score, negate, and gtn.backward.
 *
 *   emissions       [B, T, C] float32 row-major (ctc.py:33,43-44: log_probs[b])
 *   targets         concatenated int32 labels; target_offsets [B+1]
 *   blank           blank label (ctc.py:32 blank_idx)
 *   max_target_len  max_b (target_offsets[b+1] - target_offsets[b]); sizes the
 *                   workspace and shared memory (labels must lie in [0, C))
 *   grad_scale      [B] or NULL(=1): multiplies utterance b's gradient; the
 *                   Function passes scale_b / B (ctc.py:53-56,81,87)
 *   loss            [B] out: -Z_b (unscaled; ctc.py:49-51).  +inf if no alignment.
 *   mean_loss       [1] out or NULL: sum_b loss_b * grad_scale[b]  (= ctc.py:68-69
 *                   when grad_scale = scale_b / B); deterministic order
 *   grad            [B, T, C] out or NULL: grad_scale[b] * d(-Z_b)/dE[b]
 *                   (= -grad_scale[b] * posterior; all zero for an infeasible b)
 * ---------------------------------------------------------------------- */
WFST_API size_t wfst_ctc_workspace_bytes(int B, int T, int C, int max_target_len);

WFST_API int wfst_ctc_forward_backward(const float* emissions, const int32_t* targets,
                              const int32_t* target_offsets, int B, int T, int C, int blank,
                              int max_target_len, const float* grad_scale, float* loss,
                              float* mean_loss, float* grad, void* workspace,
                              size_t workspace_bytes, void* stream);

/* CTC on raw logits — replaces CTC.forward's log_softmax (criterions/ctc.py:107) followed by
 * CTCLossFunction (ctc.py:31-94) and the softmax backward autograd runs after it:
 *   logits [B, T, C] float32 (16-byte aligned), grad [B, T, C] out or NULL:
 *   grad_scale[b] * (softmax(logits[b]) - posterior), i.e. d(loss_b)/d logits.
 * Everything else as wfst_ctc_forward_backward.  Only shapes for which
 * wfst_ctc_logits_supported() returns 1 are handled (WFST_ERR_UNSUPPORTED otherwise; the
 * caller then composes log_softmax with wfst_ctc_forward_backward). */
WFST_API int wfst_ctc_logits_supported(int B, int T, int C, int max_target_len);
WFST_API size_t wfst_ctc_logits_workspace_bytes(int B, int T, int C, int max_target_len);
WFST_API int wfst_ctc_logits_forward_backward(const float* logits, const int32_t* targets,
                                     const int32_t* target_offsets, int B, int T, int C, int blank,
                                     int max_target_len, const float* grad_scale, float* loss,
                                     float* mean_loss, float* grad, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* Same computation with HOST buffers: copies inputs to the device, runs the
 * kernels and copies loss / mean_loss (and grad when not NULL) back, then
 * synchronises.  This is the call a binding that owns host memory would make
 * (the reference hands GTN a host pointer, ctc.py:43-44). */
WFST_API int wfst_ctc_forward_backward_host(const float* emissions, const int32_t* targets,
                                   const int32_t* target_offsets, int B, int T, int C,
                                   int blank, const float* grad_scale, float* loss,
                                   float* mean_loss, float* grad);

/* ------------------------------------------------------------------------
 * Generic acceptor lattice — replaces forward_score(intersect(emissions, A_b))
 * + gtn.backward for arbitrary epsilon-free acceptors A_b (STC stc.py:85-86;
 * transducer alignments transducer.py:283; ASG force-align asg.py:111-113).
 * A batch of acceptors is passed packed: nodes and arcs of all utterances are
 * concatenated; arcs are listed twice, grouped by destination (in_*) and by
 * source (out_*), each entry carrying the ORIGINAL arc index (position in
 * `weights` / `grad_weights`, relative to arc_offsets[b]).
 * ---------------------------------------------------------------------- */
typedef struct {
  int32_t B;
  int32_t max_nodes;            /* max over b of nodes in A_b */
  int32_t max_arcs;             /* max over b of arcs in A_b  */
  const int32_t* node_offsets;  /* [B+1] */
  const int32_t* arc_offsets;   /* [B+1] */
  const uint8_t* node_flags;    /* [nodes] bit0 = start, bit1 = accept */
  const int32_t* in_ptr;        /* [nodes + B] per-utterance CSR (N_b+1 entries at node_offsets[b]+b) */
  const int32_t* in_src;        /* [arcs] local source node, grouped by destination */
  const int32_t* in_label;      /* [arcs] */
  const int32_t* in_arc;        /* [arcs] original (local) arc index */
  const int32_t* out_ptr;       /* [nodes + B] */
  const int32_t* out_dst;       /* [arcs] local destination node, grouped by source */
  const int32_t* out_label;     /* [arcs] */
  const int32_t* out_arc;       /* [arcs] */
  const float* weights;         /* [arcs] original order, or NULL (= 0) */
  /* ABI 2: final weights.  A path that ends in accept node v scores final_weights[v] on top
   * (NULL = 0).  This is how acceptors with epsilon arcs (n-gram </s> arcs, back-off arcs:
   * transducer.py:52-56, scripts/build_transitions.py) reach the kernels: the host folds every
   * epsilon path into the arc that follows it and every epsilon path into an accept node
   * into that node's final weight (gtn_applications_b200/epsilon.py). */
  const float* final_weights;   /* [nodes] or NULL */
  float* grad_final_weights;    /* [nodes] out or NULL: grad_scale[b] * dZ_b/d final_weights
                                   (summed over b when the graph is shared) */
} wfst_acceptor_batch_t;

WFST_API size_t wfst_lattice_workspace_bytes(int B, int T, int C, int total_nodes, int max_nodes);

/*   shared_graph  0: graphs->B == B, utterance b is scored against graph b;
 *                 1: graphs->B == 1 and every utterance is scored against graph 0
 *                    (transducer.py:287 em o transitions); grad_weights is then
 *                    the sum over b
 *   scores        [B] out: Z_b (-inf if no accepting path)
 *   grad_emissions[B, T, C] out or NULL: grad_scale[b] * dZ_b/dE[b]; when
 *                 `accumulate` != 0 it is added to what the buffer holds
 *   grad_weights  [arcs] out or NULL: grad_scale[b] * dZ_b/dw_a (original order) */
WFST_API int wfst_lattice_forward_backward(const float* emissions, int B, int T, int C,
                                  const wfst_acceptor_batch_t* graphs, int shared_graph,
                                  const float* grad_scale, float* scores, float* grad_emissions,
                                  int accumulate, float* grad_weights, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* The same emissions against K shared acceptors, one after the other on `stream` — replaces the
 * loops over lexicon entries and windows of ConvTransduce1DFunction.forward / .backward
 * (criterions/transducer.py:487-509, 529-556), which score every window against every kernel graph.
 *   graphs        K acceptor batches with B == 1 each (acceptor k is shared by all B items)
 *   grad_scale    [K, B] or NULL (= 1)
 *   scores        [K, B] out
 *   grad_emissions[B, T, C] or NULL: sum over k of grad_scale[k, b] * dZ_kb/dE[b] is ADDED to what
 *                 the buffer holds (the caller clears it)
 *   grad_weights  K pointers ([arcs of acceptor k] out, summed over b) or NULL
 *   workspace     >= wfst_lattice_workspace_bytes(B, T, C, 0, max over k of graphs[k].max_nodes) */
/* The same, in ONE launch: every emission item against every acceptor of a packed batch of K.
 *   graphs        one batch of K acceptors (no final weights)
 *   grad_scale    [K, B] or NULL (= 1);  scores [K, B] out
 *   grad_emissions[B, T, C] or NULL: ADDED to (float atomics; the caller clears it)
 *   grad_weights  [arcs of the batch] or NULL: ADDED to, summed over the B items (the caller clears it)
 *   workspace     >= wfst_lattice_workspace_bytes(K * B, T, C, 0, graphs->max_nodes)
 * Returns WFST_ERR_UNSUPPORTED (and launches nothing) when the acceptors do not fit the
 * shared-memory kernel; wfst_lattice_forward_backward_many takes any acceptor. */
WFST_API int wfst_lattice_forward_backward_cross(const float* emissions, int B, int T, int C,
                                  const wfst_acceptor_batch_t* graphs, const float* grad_scale,
                                  float* scores, float* grad_emissions, float* grad_weights,
                                  void* workspace, size_t workspace_bytes, void* stream);

WFST_API int wfst_lattice_forward_backward_many(const float* emissions, int B, int T, int C,
                                  const wfst_acceptor_batch_t* graphs, int K,
                                  const float* grad_scale, float* scores, float* grad_emissions,
                                  float* const* grad_weights, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * ASG — replaces ASGLossFunction.forward + .backward (criterions/asg.py:83-185):
 *   loss_b = Z(E_b o transitions) - Z((force_align(y_b) o transitions) o E_b)
 *   transitions [(C+1), C]: row 0 = start scores, transitions[1+i, j] = i after j
 *   (layout of create_transitions_graph, asg.py:53-69).
 *   grad_transitions [(C+1), C] out or NULL: sum_b grad_scale[b] * dloss_b/dtransitions
 *   (asg.py:175-179 with grad_scale = scale_b / B).
 * ---------------------------------------------------------------------- */
WFST_API size_t wfst_asg_workspace_bytes(int B, int T, int C, int max_target_len);

WFST_API int wfst_asg_forward_backward(const float* emissions, const float* transitions,
                              const int32_t* targets, const int32_t* target_offsets, int B,
                              int T, int C, int max_target_len, const float* grad_scale,
                              float* loss, float* mean_loss, float* grad_emissions,
                              float* grad_transitions, void* workspace, size_t workspace_bytes,
                              void* stream);

/* ------------------------------------------------------------------------
 * Host-side graphs — replaces the gtn.Graph objects and small-graph operations the
 * reference performs per utterance before scoring (criterions/transducer.py:23-123,
 * 260-281; criterions/stc.py:22-64; utils.py:261 gtn.load; tests use gtn.loadtxt):
 * gtn.Graph / add_node / add_arc / arc_sort / mark_arc_sorted / set_weights /
 * compose|intersect / remove / project_input|output / linear_graph / load / save.
 * Graphs live in the library and are named by int32 handles (>= 0); functions that
 * return a handle return a negative error code on failure.  Node and arc numbering
 * follows GTN's construction order (composition: breadth-first from the start pairs
 * after a backward co-reachability pass, arcs in matcher order), epsilon = -1.
 * ---------------------------------------------------------------------- */
WFST_API int32_t wfst_graph_create(int calc_grad);
WFST_API int wfst_graph_destroy(int32_t graph);
WFST_API int wfst_graph_destroy_many(const int32_t* graphs, int n);          /* invalid handles are skipped */
WFST_API int wfst_graph_add_node(int32_t graph, int start, int accept);      /* returns the node id */
WFST_API int wfst_graph_add_arc(int32_t graph, int src, int dst, int ilabel, int olabel,
                                float weight);                                /* returns the arc id */
WFST_API int wfst_graph_add_arcs(int32_t graph, int n, const int32_t* src, const int32_t* dst,
                                 const int32_t* ilabel, const int32_t* olabel, const float* weight);
WFST_API int wfst_graph_num_nodes(int32_t graph);
WFST_API int wfst_graph_num_arcs(int32_t graph);
WFST_API int wfst_graph_arc_sort(int32_t graph, int by_olabel);
WFST_API int wfst_graph_mark_arc_sorted(int32_t graph, int by_olabel);
WFST_API int wfst_graph_sorted_flags(int32_t graph);              /* bit0 ilabel-sorted, bit1 olabel-sorted */
WFST_API int wfst_graph_get_calc_grad(int32_t graph);
WFST_API int wfst_graph_set_calc_grad(int32_t graph, int calc_grad);
WFST_API int wfst_graph_set_weights(int32_t graph, const float* weights);     /* HOST pointer, num_arcs floats */
WFST_API int wfst_graph_get_weights(int32_t graph, float* weights);
WFST_API int wfst_graph_get_arcs(int32_t graph, int32_t* src, int32_t* dst, int32_t* ilabel,
                                 int32_t* olabel);
WFST_API int wfst_graph_get_node_flags(int32_t graph, uint8_t* flags);         /* bit0 start, bit1 accept */
WFST_API int wfst_graph_get_arc_order(int32_t graph, int incoming, int32_t* order);
WFST_API int wfst_graph_get_provenance(int32_t graph, int32_t* arc_first, int32_t* arc_second);
WFST_API int32_t wfst_graph_compose(int32_t first, int32_t second);           /* also gtn.intersect */
WFST_API int32_t wfst_graph_remove(int32_t graph, int ilabel, int olabel);
WFST_API int32_t wfst_graph_project(int32_t graph, int input);
WFST_API int32_t wfst_graph_linear(int M, int N, int calc_grad);
WFST_API int32_t wfst_graph_loadtxt(const char* path);
WFST_API int wfst_graph_savetxt(int32_t graph, const char* path);
WFST_API int32_t wfst_graph_load(const char* path);
WFST_API int wfst_graph_save(int32_t graph, const char* path);

/* STC acceptors of a whole batch — replaces STCLossFunction.create_stc_graph (criterions/stc.py:22-64)
 * called once per utterance: same node and arc order, then arc-sorted by input label.  targets are
 * token indices in [0, star_idx), concatenated; the <star> arcs weigh log_prob; out_handles [B]. */
WFST_API int wfst_stc_graphs(const int32_t* targets, const int32_t* target_offsets, int B, int star_idx,
                             float log_prob, int blank_idx, int32_t* out_handles);

/* Alignment acceptors of a whole batch (transducer.py:260-276), built on host threads:
 * project_input(remove(compose(tokens, remove(project_output(compose(chain(y_b), lexicon)))))).
 * targets are grapheme indices, concatenated; out_handles [B]. */
WFST_API int wfst_transducer_alignment_graphs(int32_t tokens, int32_t lexicon,
                                              const int32_t* targets,
                                              const int32_t* target_offsets, int B,
                                              int32_t* out_handles);
/* The graphs above depend on (tokens, lexicon, target) only and are kept, frozen (handles to them
 * reject modification), in an LRU bounded by `capacity` arcs + nodes (default 16 Mi, or
 * $WFST_ALIGN_CACHE_ARCS): a target met again (next epoch; every iteration of the reference's
 * benchmarks/transducer_benchmark.py) costs a lookup instead of two compositions.  This call
 * empties the cache, sets its capacity (0 = off, < 0 = unchanged) and returns the hit / miss
 * counters since the last call (pointers may be NULL). */
WFST_API int wfst_transducer_alignment_cache(long long capacity, unsigned long long* hits,
                                             unsigned long long* misses);

/* Transducer with a transition graph that has epsilon arcs (ngram > 1: the </s> arcs of
 * make_transitions_graph, transducer.py:52-56; back-off graphs loaded with gtn.load): for every
 * utterance intersect(transitions, aligns[b]) is formed with the epsilons in place (as
 * transducer.py:279-281 does), arc-sorted and folded into an epsilon-free acceptor with final
 * weights (gtn_applications_b200/epsilon.py describes the fold) on host threads.  out_handles [B]
 * receive the folded graphs (pack them with wfst_graph_pack); the returned fold object holds the
 * int64 index arrays that tie folded arcs / final-weight paths to the TRANSITION graph's arcs:
 * wfst_fold_sizes -> {folded arcs, final paths, nodes, arc tie entries, final tie entries};
 * wfst_fold_fill copies arc_seg / arc_src, fin_seg / fin_src, fin_node.  aligns == NULL (B = 1)
 * folds the transition graph itself. */
WFST_API int32_t wfst_fold_transitions_batch(int32_t transitions, const int32_t* aligns, int B,
                                             int32_t* out_handles);
WFST_API int wfst_fold_sizes(int32_t fold, int64_t* sizes);
WFST_API int wfst_fold_fill(int32_t fold, int64_t* arc_seg, int64_t* arc_src, int64_t* fin_seg,
                            int64_t* fin_src, int64_t* fin_node);
WFST_API int wfst_fold_destroy(int32_t fold);

/* Packs B host graphs into the arrays of wfst_acceptor_batch_t (HOST buffers sized from
 * wfst_graph_pack_sizes; the caller uploads them).  Arc lists keep their arc_sort order. */
/* Alignment -> tokens for a batch of best alignments — the host part of Transducer.viterbi
 * (criterions/transducer.py:223-233: per utterance compose(chain of frame labels, tokens),
 * viterbi_path, project_output, remove epsilons), on host threads.  `tokens` must be
 * ilabel-sorted (transducer.py:222).  labels [B, T] (host); out [B, T] (host) receives the tokens
 * of utterance b at out[b*T ..], out_counts [B] their number. */
WFST_API int wfst_transducer_decode_paths(int32_t tokens, const int32_t* labels, int B, int T,
                                          int32_t* out, int32_t* out_counts);
WFST_API int wfst_graph_pack_sizes(const int32_t* handles, int B, int32_t* total_nodes,
                                   int32_t* total_arcs, int32_t* max_nodes, int32_t* max_arcs,
                                   int32_t* has_epsilon);
WFST_API int wfst_graph_pack(const int32_t* handles, int B, int32_t* node_offsets,
                             int32_t* arc_offsets, uint8_t* node_flags, int32_t* in_ptr,
                             int32_t* in_src, int32_t* in_label, int32_t* in_arc, int32_t* out_ptr,
                             int32_t* out_dst, int32_t* out_label, int32_t* out_arc, float* weights);

/* ------------------------------------------------------------------------
 * Best path (tropical semiring) — replaces gtn.viterbi_path(gtn.intersect(emissions, A))
 * (criterions/asg.py:225, criterions/transducer.py:215-221).
 *   scores [B] best path score (-inf: none); labels / arcs [B, T]: ilabel and original arc
 *   index of the acceptor arc taken at each frame (-1 when there is no path).
 * ---------------------------------------------------------------------- */
WFST_API size_t wfst_lattice_viterbi_workspace_bytes(int B, int T, int max_nodes);
WFST_API int wfst_lattice_viterbi(const float* emissions, int B, int T, int C,
                                  const wfst_acceptor_batch_t* graphs, int shared_graph,
                                  float* scores, int32_t* labels, int32_t* arcs, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* Best path through emissions x the ASG bigram transition graph — ASG.viterbi
 * (criterions/asg.py:217-226: gtn.viterbi_path(gtn.intersect(g_emissions, g_transitions)) with
 * the graph of create_transitions_graph, asg.py:53-69) without building the graph: one warp per
 * utterance, transitions [(C+1), C] read directly.  Same result as wfst_lattice_viterbi on the
 * packed transition graph, ties included (lowest previous label, lowest final label).
 *   labels [B, T] out: label taken at each frame (-1: no path);  scores [B] out.
 * wfst_asg_viterbi_supported: 1 when the shape fits the dense kernel (C <= 32, T >= 1 and the
 * back-pointers, 32 bytes per frame, fit shared memory); otherwise the call returns
 * WFST_ERR_UNSUPPORTED and the caller uses wfst_lattice_viterbi. */
WFST_API int wfst_asg_viterbi_supported(int T, int C);
WFST_API int wfst_asg_viterbi(const float* emissions, const float* transitions, int B, int T, int C,
                              float* scores, int32_t* labels, void* stream);

/* Host-side best path of a small acyclic graph (the alignment -> token mapping of
 * Transducer.viterbi, transducer.py:222-229): returns a chain graph with the arcs of the
 * best path, first maximum in in-list order on ties (GTN's traversal order). */
WFST_API int32_t wfst_graph_viterbi_path(int32_t graph);

/* Host-side scoring of an acyclic graph: gtn.forward_score (tropical = 0) or gtn.viterbi_score
 * (tropical != 0) and, if arc_grad != NULL (num_arcs floats), the gradient gtn.backward would
 * leave on its arcs (arc posteriors / indicator of the best path).  This serves the
 * graph-building API (the reference's tests build expected values this way,
 * tests/transducer_test.py:218-273); batched scoring of emissions runs on the GPU. */
WFST_API int wfst_graph_score(int32_t graph, int tropical, float* score, float* arc_grad);

/* Multiplies x[0..n) in place by *scale (a device scalar); returns immediately on
 * the device when *scale == 1 (the common loss.backward() case), so autograd's
 * grad_output costs no pass over the [B,T,C] gradient (ctc.py:87, asg.py:174). */
WFST_API int wfst_scale_inplace(float* x, size_t n, const float* scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WFST_B200_H */
