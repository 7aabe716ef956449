"""Epsilon arcs for the time-synchronous lattice kernels.

The reference scores emissions against transition graphs that contain epsilon arcs — the
`</s>` arcs of `make_transitions_graph` for ngram > 1 (criterions/transducer.py:52-56) and
the back-off arcs of graphs built by scripts/build_transitions.py / loaded with `gtn.load`
(utils.py:261) — by letting GTN's `intersect` follow them between two frames.  The GPU
kernels advance one frame per arc, so the host folds the epsilons away first:

  * every path  u --eps*--> v --(label a)--> x  becomes one arc  u --a--> x  whose weight is
    the SUM of the weights along the path (parallel arcs are summed by the kernel's
    log-sum-exp, exactly as the separate lattice paths would be);
  * every path  u --eps*--> v  into an accept node makes u accepting with a FINAL weight
    (log-sum-exp over such paths; max for Viterbi) — `wfst_acceptor_batch_t.final_weights`.

Weights stay tied to the original arcs: the folded structure is computed once from the
topology (numpy), the weights of a call are gathered from the current parameters with
torch ops, and the gradients the kernel returns for folded arcs / final weights are
scattered back to the original arcs."""
import numpy as np
import torch

EPS = -1


class FoldedAcceptor:
    """Epsilon-free view of one acceptor.  Arrays (numpy): start, accept [N]; src, dst, label
    [A'] of the folded arcs; arc_prov_ptr [A'+1], arc_prov [..] original arc ids summed into
    folded arc k; fin_node [P], fin_prov_ptr [P+1], fin_prov [..] one entry per epsilon path
    (possibly empty) from node fin_node[p] into an original accept node."""

    def __init__(self, arrays):
        start = np.asarray(arrays["start"]).astype(bool)
        accept = np.asarray(arrays["accept"]).astype(bool)
        src = np.asarray(arrays["src"], dtype=np.int64)
        dst = np.asarray(arrays["dst"], dtype=np.int64)
        lab = np.asarray(arrays["ilabel"], dtype=np.int64)
        N, A = len(start), len(src)
        eps_out = [[] for _ in range(N)]
        emit_out = [[] for _ in range(N)]
        for a in range(A):
            (eps_out if lab[a] == EPS else emit_out)[src[a]].append(a)
        closures = {}

        def closure(u, depth=0):
            """all epsilon paths from u: list of (node, tuple(arc ids))"""
            if u in closures:
                return closures[u]
            if depth > N:
                raise ValueError("epsilon cycle in acceptor")
            out = [(u, ())]
            for a in eps_out[u]:
                for v, path in closure(int(dst[a]), depth + 1):
                    out.append((v, (a,) + path))
            closures[u] = out
            return out

        n_src, n_dst, n_lab, prov_ptr, prov = [], [], [], [0], []
        fin_node, fin_ptr, fin_prov = [], [0], []
        for u in range(N):
            for v, path in closure(u):
                for a in emit_out[v]:
                    n_src.append(u); n_dst.append(int(dst[a])); n_lab.append(int(lab[a]))
                    prov.extend(path); prov.append(a)
                    prov_ptr.append(len(prov))
                if accept[v]:
                    fin_node.append(u)
                    fin_prov.extend(path)
                    fin_ptr.append(len(fin_prov))
        self.num_nodes, self.num_orig_arcs = N, A
        self.start = start
        self.accept = np.zeros(N, dtype=bool)
        self.accept[np.asarray(fin_node, dtype=np.int64)] = True
        self.src = np.asarray(n_src, dtype=np.int32)
        self.dst = np.asarray(n_dst, dtype=np.int32)
        self.label = np.asarray(n_lab, dtype=np.int32)
        self.arc_prov_ptr = np.asarray(prov_ptr, dtype=np.int64)
        self.arc_prov = np.asarray(prov, dtype=np.int64)
        self.fin_node = np.asarray(fin_node, dtype=np.int64)
        self.fin_prov_ptr = np.asarray(fin_ptr, dtype=np.int64)
        self.fin_prov = np.asarray(fin_prov, dtype=np.int64)

    def graph_dict(self):
        """what packing.PackedAcceptors takes"""
        return {"start": self.start, "accept": self.accept, "src": self.src, "dst": self.dst,
                "label": self.label}


def has_epsilon(arrays):
    lab = np.asarray(arrays["ilabel"])
    return bool(lab.size and (lab == EPS).any())


class FoldedBatch:
    """Index tensors (on `device`) that tie the folded arcs / final weights of a batch of
    FoldedAcceptors to ONE vector of original parameters.  `orig_index[b]` maps the original
    arc ids of acceptor b into that vector (None = identity)."""

    def __init__(self, folded, device, orig_index=None):
        arc_seg, arc_src, fin_seg, fin_src, fin_node = [], [], [], [], []
        a0 = p0 = n0 = 0
        for b, f in enumerate(folded):
            remap = (lambda x: x) if orig_index is None or orig_index[b] is None else \
                (lambda x, m=np.asarray(orig_index[b], dtype=np.int64): m[x])
            counts = np.diff(f.arc_prov_ptr)
            arc_seg.append(np.repeat(np.arange(len(counts), dtype=np.int64) + a0, counts))
            arc_src.append(remap(f.arc_prov))
            fcounts = np.diff(f.fin_prov_ptr)
            fin_seg.append(np.repeat(np.arange(len(fcounts), dtype=np.int64) + p0, fcounts))
            fin_src.append(remap(f.fin_prov))
            fin_node.append(f.fin_node + n0)
            a0 += len(counts); p0 += len(fcounts); n0 += f.num_nodes
        cat = lambda xs: torch.from_numpy(np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)).to(device)  # noqa: E731
        self.num_arcs, self.num_paths, self.num_nodes = a0, p0, n0
        self.arc_seg, self.arc_src = cat(arc_seg), cat(arc_src)
        self.fin_seg, self.fin_src, self.fin_node = cat(fin_seg), cat(fin_src), cat(fin_node)

    @classmethod
    def from_fold(cls, fold, device):
        """from a fold object of the host library (wfst_fold_transitions_batch: the composition with the
        transition graph, the fold and these index arrays for the whole batch in C++ on host threads)"""
        from . import _lib
        L = _lib.lib()
        sizes = np.zeros(5, dtype=np.int64)
        _lib.check(L.wfst_fold_sizes(fold, sizes.ctypes.data))
        na, npaths, nn, ea, ef = (int(x) for x in sizes)
        # the five index arrays are filled into ONE pinned staging tensor and travel in one
        # asynchronous copy (five synchronous copies from pageable memory made the host wait for
        # the previous step's kernels: 1.9 ms per step of the n-gram transducers)
        cuts = np.cumsum([0, ea, ea, ef, ef, npaths])
        host = torch.empty(int(cuts[-1]), dtype=torch.int64, pin_memory=torch.cuda.is_available())
        base = host.data_ptr()
        _lib.check(L.wfst_fold_fill(fold, *(base + 8 * int(c) for c in cuts[:5])))
        dev = host.to(device, non_blocking=True)
        self = cls.__new__(cls)
        self.num_arcs, self.num_paths, self.num_nodes = na, npaths, nn
        self.arc_seg, self.arc_src, self.fin_seg, self.fin_src, self.fin_node = (
            dev[int(cuts[i]):int(cuts[i + 1])] for i in range(5))
        return self

    def weights(self, params, tropical=False):
        """(folded arc weights [A'], final weights [N] (-inf where not accepting), path weights [P])"""
        dev = params.device
        w = torch.zeros(self.num_arcs, dtype=torch.float32, device=dev)
        if self.arc_src.numel():
            w.index_add_(0, self.arc_seg, params[self.arc_src])
        pw = torch.zeros(self.num_paths, dtype=torch.float32, device=dev)
        if self.fin_src.numel():
            pw.index_add_(0, self.fin_seg, params[self.fin_src])
        fmax = torch.full((self.num_nodes,), float("-inf"), dtype=torch.float32, device=dev)
        if self.num_paths:
            fmax.scatter_reduce_(0, self.fin_node, pw, reduce="amax", include_self=True)
        if tropical:
            return w, fmax, pw
        fsum = torch.zeros(self.num_nodes, dtype=torch.float32, device=dev)
        if self.num_paths:
            fsum.index_add_(0, self.fin_node, torch.exp(pw - fmax[self.fin_node]))
        fw = torch.where(fsum > 0, fmax + torch.log(fsum.clamp_min(1e-38)), fmax)
        return w, fw, pw

    def scatter_grads(self, num_params, g_arcs, g_final, fw, pw):
        """d/d original parameters from d/d folded arcs and d/d final weights"""
        dev = g_arcs.device if g_arcs is not None else g_final.device
        g = torch.zeros(num_params, dtype=torch.float32, device=dev)
        if g_arcs is not None and self.arc_src.numel():
            g.index_add_(0, self.arc_src, g_arcs[self.arc_seg])
        if g_final is not None and self.fin_src.numel():
            share = torch.exp(pw - fw[self.fin_node])            # softmax over a node's paths
            gp = g_final[self.fin_node] * share
            g.index_add_(0, self.fin_src, gp[self.fin_seg])
        return g
