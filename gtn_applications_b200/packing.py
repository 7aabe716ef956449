"""Packs a batch of epsilon-free acceptors into the layout the lattice kernel
reads (wfst_acceptor_batch_t, include/wfst_b200.h): nodes and arcs of all
utterances concatenated, arcs listed once grouped by destination and once grouped
by source, each entry carrying the original arc index."""
import ctypes

import numpy as np
import torch

from . import _lib


class PackedAcceptors:
    """Device-resident packed batch; keeps the tensors alive for the C struct."""

    def __init__(self, graphs, device):
        """graphs: list of dicts with numpy arrays start, accept (bool/int [N]),
        src, dst, label (int [A]) and optional weight (float [A])."""
        B = len(graphs)
        node_off = np.zeros(B + 1, dtype=np.int32)
        arc_off = np.zeros(B + 1, dtype=np.int32)
        for b, g in enumerate(graphs):
            node_off[b + 1] = node_off[b] + len(g["start"])
            arc_off[b + 1] = arc_off[b] + len(g["src"])
        nn, na = int(node_off[-1]), int(arc_off[-1])
        flags = np.zeros(nn, dtype=np.uint8)
        in_ptr = np.zeros(nn + B, dtype=np.int32)
        out_ptr = np.zeros(nn + B, dtype=np.int32)
        cols = {k: np.zeros(na, dtype=np.int32) for k in
                ("in_src", "in_label", "in_arc", "out_dst", "out_label", "out_arc")}
        weights = np.zeros(na, dtype=np.float32)
        for b, g in enumerate(graphs):
            n0, a0 = int(node_off[b]), int(arc_off[b])
            N, A = len(g["start"]), len(g["src"])
            flags[n0:n0 + N] = (np.asarray(g["start"]).astype(np.uint8) & 1) | \
                ((np.asarray(g["accept"]).astype(np.uint8) & 1) << 1)
            src = np.asarray(g["src"], dtype=np.int64)
            dst = np.asarray(g["dst"], dtype=np.int64)
            lab = np.asarray(g["label"], dtype=np.int32)
            if A and lab.min() < 0:
                raise ValueError("epsilon arcs are not supported by the lattice kernel")
            if "weight" in g and g["weight"] is not None:
                weights[a0:a0 + A] = np.asarray(g["weight"], dtype=np.float32)
            by_dst = np.argsort(dst, kind="stable")
            by_src = np.argsort(src, kind="stable")
            cols["in_src"][a0:a0 + A] = src[by_dst]
            cols["in_label"][a0:a0 + A] = lab[by_dst]
            cols["in_arc"][a0:a0 + A] = by_dst
            cols["out_dst"][a0:a0 + A] = dst[by_src]
            cols["out_label"][a0:a0 + A] = lab[by_src]
            cols["out_arc"][a0:a0 + A] = by_src
            in_ptr[n0 + b + 1:n0 + b + N + 1] = np.cumsum(np.bincount(dst, minlength=N))
            out_ptr[n0 + b + 1:n0 + b + N + 1] = np.cumsum(np.bincount(src, minlength=N))
        self.B = B
        self.num_nodes = nn
        self.num_arcs = na
        self.arc_offsets_host = arc_off
        self.max_nodes = int(np.max(np.diff(node_off))) if B else 0
        self.max_arcs = int(np.max(np.diff(arc_off))) if B else 0
        up = lambda a: torch.from_numpy(a).to(device)  # noqa: E731
        self.t = {"node_offsets": up(node_off), "arc_offsets": up(arc_off), "node_flags": up(flags),
                  "in_ptr": up(in_ptr), "out_ptr": up(out_ptr), "weights": up(weights)}
        for k, v in cols.items():
            self.t[k] = up(v)

    def struct(self, weights=None, final_weights=None, grad_final_weights=None):
        s = _lib.AcceptorBatch()
        s.B, s.max_nodes, s.max_arcs = self.B, self.max_nodes, self.max_arcs
        for k in ("node_offsets", "arc_offsets", "node_flags", "in_ptr", "in_src", "in_label",
                  "in_arc", "out_ptr", "out_dst", "out_label", "out_arc"):
            setattr(s, k, self.t[k].data_ptr())
        w = self.t["weights"] if weights is None else weights
        s.weights = w.data_ptr()
        s.final_weights = final_weights.data_ptr() if final_weights is not None else None
        s.grad_final_weights = grad_final_weights.data_ptr() if grad_final_weights is not None else None
        return s
