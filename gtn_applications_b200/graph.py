"""Host-side graphs of the product: a thin Python face over the C++ container in
libwfst_b200.so (csrc/graph.cpp, handles via the C ABI in include/wfst_b200.h).

It covers the part of the ``gtn`` Python API the reference's criteria use to BUILD
graphs (SURVEY.md §8(b) Level 3): ``Graph`` with add_node / add_arc / arc_sort /
mark_arc_sorted / set_weights / num_arcs / num_nodes / weights_to_numpy /
labels_to_list / calc_grad, and the free functions compose / intersect / remove /
project_input / project_output / linear_graph / load / save / loadtxt / equal /
isomorphic.  Scoring (forward_score / backward / viterbi) is NOT here: that is what the
CUDA kernels do."""
import ctypes

import numpy as np

from . import _lib

epsilon = -1


def _L():
    return _lib.lib()


def _h(rc):
    """a returned handle / id, or raise with the library's message"""
    if rc < 0:
        _lib.check(rc)
    return rc


class Graph:
    """gtn.Graph(calc_grad=True) look-alike holding a handle to a C++ HostGraph."""

    def __init__(self, calc_grad=True, _handle=None):
        self._h = _h(_L().wfst_graph_create(1 if calc_grad else 0)) if _handle is None else _handle

    def __del__(self):
        try:
            if self._h is not None and _lib._lib is not None:
                _lib._lib.wfst_graph_destroy(self._h)
        except Exception:
            pass
        self._h = None

    # ---- construction
    def add_node(self, start=False, accept=False):
        return _h(_L().wfst_graph_add_node(self._h, int(bool(start)), int(bool(accept))))

    def add_arc(self, src_node, dst_node, ilabel, olabel=None, weight=0.0):
        if olabel is None:
            olabel = ilabel
        return _h(_L().wfst_graph_add_arc(self._h, int(src_node), int(dst_node), int(ilabel),
                                          int(olabel), float(weight)))

    def add_arcs(self, src, dst, ilabel, olabel=None, weight=None):
        """bulk add (numpy int32 arrays) — not in gtn; used by the criteria's builders"""
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        il = np.ascontiguousarray(ilabel, dtype=np.int32)
        ol = None if olabel is None else np.ascontiguousarray(olabel, dtype=np.int32)
        w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float32)
        _lib.check(_L().wfst_graph_add_arcs(
            self._h, len(src), src.ctypes.data, dst.ctypes.data, il.ctypes.data,
            None if ol is None else ol.ctypes.data, None if w is None else w.ctypes.data))

    def arc_sort(self, olabel=False):
        _lib.check(_L().wfst_graph_arc_sort(self._h, int(bool(olabel))))

    def mark_arc_sorted(self, olabel=False):
        _lib.check(_L().wfst_graph_mark_arc_sorted(self._h, int(bool(olabel))))

    def ilabel_sorted(self):
        return bool(_h(_L().wfst_graph_sorted_flags(self._h)) & 1)

    def olabel_sorted(self):
        return bool(_h(_L().wfst_graph_sorted_flags(self._h)) & 2)

    # ---- sizes / flags
    def num_nodes(self):
        return _h(_L().wfst_graph_num_nodes(self._h))

    def num_arcs(self):
        return _h(_L().wfst_graph_num_arcs(self._h))

    @property
    def calc_grad(self):
        return bool(_h(_L().wfst_graph_get_calc_grad(self._h)))

    @calc_grad.setter
    def calc_grad(self, v):
        _lib.check(_L().wfst_graph_set_calc_grad(self._h, int(bool(v))))

    # ---- weights
    def set_weights(self, weights):
        """An int is a raw host pointer to num_arcs float32 values (what the reference
        passes: tensor.data_ptr()); sequences / arrays are copied."""
        n = self.num_arcs()
        if isinstance(weights, int):
            _lib.check(_L().wfst_graph_set_weights(self._h, ctypes.c_void_p(weights)))
            return
        arr = np.ascontiguousarray(np.asarray(weights, dtype=np.float32).reshape(-1))
        if arr.size != n:
            raise ValueError("set_weights needs num_arcs (%d) values, got %d" % (n, arr.size))
        _lib.check(_L().wfst_graph_set_weights(self._h, arr.ctypes.data))

    def weights_to_numpy(self):
        out = np.empty(self.num_arcs(), dtype=np.float32)
        _lib.check(_L().wfst_graph_get_weights(self._h, out.ctypes.data))
        return out

    def weights_to_list(self):
        return self.weights_to_numpy().tolist()

    def item(self):
        if self.num_arcs() != 1:
            raise ValueError("item() needs a graph with exactly one arc")
        return float(self.weights_to_numpy()[0])

    # ---- structure dumps
    def arrays(self):
        """dict(start, accept, src, dst, ilabel, olabel, weight) as numpy arrays"""
        n, a = self.num_nodes(), self.num_arcs()
        flags = np.empty(n, dtype=np.uint8)
        cols = [np.empty(a, dtype=np.int32) for _ in range(4)]
        _lib.check(_L().wfst_graph_get_node_flags(self._h, flags.ctypes.data))
        _lib.check(_L().wfst_graph_get_arcs(self._h, *[c.ctypes.data for c in cols]))
        return {"start": (flags & 1).astype(np.int32), "accept": ((flags >> 1) & 1).astype(np.int32),
                "src": cols[0], "dst": cols[1], "ilabel": cols[2], "olabel": cols[3],
                "weight": self.weights_to_numpy().astype(np.float64)}

    def labels_to_list(self, ilabel=True):
        return self.arrays()["ilabel" if ilabel else "olabel"].tolist()

    def arc_order(self, incoming=False):
        out = np.empty(self.num_arcs(), dtype=np.int32)
        _lib.check(_L().wfst_graph_get_arc_order(self._h, int(bool(incoming)), out.ctypes.data))
        return out

    def provenance(self):
        a = self.num_arcs()
        p1, p2 = np.empty(a, dtype=np.int32), np.empty(a, dtype=np.int32)
        _lib.check(_L().wfst_graph_get_provenance(self._h, p1.ctypes.data, p2.ctypes.data))
        return p1, p2


# ---- free functions (gtn names) ---------------------------------------------------
def compose(first, second):
    return Graph(_handle=_h(_L().wfst_graph_compose(first._h, second._h)))


intersect = compose


def remove(g, ilabel=epsilon, olabel=None):
    if olabel is None:
        olabel = ilabel
    return Graph(_handle=_h(_L().wfst_graph_remove(g._h, int(ilabel), int(olabel))))


def project_input(g):
    return Graph(_handle=_h(_L().wfst_graph_project(g._h, 1)))


def project_output(g):
    return Graph(_handle=_h(_L().wfst_graph_project(g._h, 0)))


def linear_graph(M, N, *args, **kwargs):
    """linear_graph(M, N, device, calc_grad) / linear_graph(M, N, calc_grad) (ctc.py:40)"""
    calc_grad = kwargs.get("calc_grad", True)
    if len(args) == 1 and isinstance(args[0], bool):
        calc_grad = args[0]
    elif len(args) >= 2:
        calc_grad = args[1]
    return Graph(_handle=_h(_L().wfst_graph_linear(int(M), int(N), int(bool(calc_grad)))))


def viterbi_path(g):
    """host best path of a small acyclic graph (see wfst_graph_viterbi_path)"""
    return Graph(_handle=_h(_L().wfst_graph_viterbi_path(g._h)))


def loadtxt(path):
    return Graph(_handle=_h(_L().wfst_graph_loadtxt(str(path).encode())))


def savetxt(path, g):
    _lib.check(_L().wfst_graph_savetxt(g._h, str(path).encode()))


def load(path):
    return Graph(_handle=_h(_L().wfst_graph_load(str(path).encode())))


def save(path, g):
    _lib.check(_L().wfst_graph_save(g._h, str(path).encode()))


def equal(a, b):
    """same node flags and the same multiset of (src, dst, ilabel, olabel, weight)"""
    x, y = a.arrays(), b.arrays()
    if len(x["start"]) != len(y["start"]) or len(x["src"]) != len(y["src"]):
        return False
    if not (np.array_equal(x["start"], y["start"]) and np.array_equal(x["accept"], y["accept"])):
        return False
    key = lambda d: sorted(zip(d["src"].tolist(), d["dst"].tolist(), d["ilabel"].tolist(),  # noqa: E731
                               d["olabel"].tolist(), d["weight"].tolist()))
    return key(x) == key(y)


def isomorphic(a, b):
    """structure-preserving node bijection (backtracking from the start nodes; the graphs
    the reference compares this way are small)"""
    x, y = a.arrays(), b.arrays()
    n = len(x["start"])
    if n != len(y["start"]) or len(x["src"]) != len(y["src"]) or \
            x["start"].sum() != y["start"].sum() or x["accept"].sum() != y["accept"].sum():
        return False
    if n == 0:
        return True

    def out_lists(d):
        out = [[] for _ in range(n)]
        for s, t, i, o, w in zip(d["src"], d["dst"], d["ilabel"], d["olabel"], d["weight"]):
            out[s].append((int(t), int(i), int(o), float(w)))
        return out

    ox, oy = out_lists(x), out_lists(y)

    def match(u, v, mapping):
        if u in mapping:
            return mapping[u] == v
        if v in mapping.values() or x["start"][u] != y["start"][v] or x["accept"][u] != y["accept"][v] \
                or len(ox[u]) != len(oy[v]):
            return False
        mapping[u] = v
        used = set()
        for (t, i, o, w) in ox[u]:
            ok = False
            for k, (t2, i2, o2, w2) in enumerate(oy[v]):
                if k in used or (i, o, w) != (i2, o2, w2):
                    continue
                snap = dict(mapping)
                if match(t, t2, mapping):
                    used.add(k)
                    ok = True
                    break
                mapping.clear()
                mapping.update(snap)
            if not ok:
                del mapping[u]
                return False
        return True

    sx = [i for i in range(n) if x["start"][i]]
    sy = [i for i in range(n) if y["start"][i]]
    if not sx:
        return equal(a, b)
    return any(match(sx[0], s, {}) for s in sy)


# ---- batch packing for the lattice kernel --------------------------------------------
def pack_graphs(graphs, device):
    """list[Graph] -> packing.PackedAcceptors (device resident) via wfst_graph_pack"""
    B = len(graphs)
    return pack_handles((ctypes.c_int32 * B)(*[g._h for g in graphs]), B, device)


def destroy_handles(handles, B):
    """frees B host graphs that were never wrapped in Graph objects (one call, host threads)"""
    _lib.check(_L().wfst_graph_destroy_many(handles, B))


def pack_handles(handles, B, device):
    """ctypes int32 array of B graph handles -> packing.PackedAcceptors (device resident)"""
    import torch
    from .packing import PackedAcceptors
    tn, ta, mn, ma, eps = (ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(),
                           ctypes.c_int32())
    _lib.check(_L().wfst_graph_pack_sizes(handles, B, ctypes.byref(tn), ctypes.byref(ta),
                                          ctypes.byref(mn), ctypes.byref(ma), ctypes.byref(eps)))
    if eps.value:
        raise NotImplementedError("acceptors with epsilon arcs cannot be scored by the lattice kernel yet")
    tn, ta = tn.value, ta.value
    # one pinned int32 staging buffer -> one H2D copy
    sizes = {"node_offsets": B + 1, "arc_offsets": B + 1, "in_ptr": tn + B, "out_ptr": tn + B,
             "in_src": ta, "in_label": ta, "in_arc": ta, "out_dst": ta, "out_label": ta, "out_arc": ta,
             "weights": ta}
    total = sum(sizes.values())
    # allocated pinned (torch caches pinned blocks); .pin_memory() would allocate pageable memory
    # first and copy it
    ints = torch.empty(total, dtype=torch.int32, pin_memory=True)
    flags = torch.empty(max(tn, 1), dtype=torch.uint8, pin_memory=True)
    views, pos = {}, 0
    for k, n in sizes.items():
        views[k] = ints[pos:pos + n]
        pos += n
    ptr = lambda t: t.data_ptr()  # noqa: E731
    _lib.check(_L().wfst_graph_pack(
        handles, B, ptr(views["node_offsets"]), ptr(views["arc_offsets"]), ptr(flags),
        ptr(views["in_ptr"]), ptr(views["in_src"]), ptr(views["in_label"]), ptr(views["in_arc"]),
        ptr(views["out_ptr"]), ptr(views["out_dst"]), ptr(views["out_label"]), ptr(views["out_arc"]),
        ptr(views["weights"])))
    dev_ints = ints.to(device, non_blocking=True)
    dev_flags = flags.to(device, non_blocking=True)
    packed = PackedAcceptors.__new__(PackedAcceptors)
    packed.B, packed.num_arcs = B, ta
    packed.num_nodes = tn
    packed.max_nodes, packed.max_arcs = mn.value, ma.value
    packed.arc_offsets_host = views["arc_offsets"].numpy().copy()
    packed.t, pos = {"node_flags": dev_flags}, 0
    for k, n in sizes.items():
        t = dev_ints[pos:pos + n]
        packed.t[k] = t.view(torch.float32) if k == "weights" else t
        pos += n
    packed._keep = (ints, flags)
    return packed
