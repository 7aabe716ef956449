"""`import gtn` for the reference's callers (utils.py:10, criterions/*.py, tests/*.py), on the
product's host graph library (gtn_applications_b200/graph.py -> csrc/graph.cpp).

Graph building is the product API as is.  On top of it this module keeps the small autograd
tape the reference's tests rely on when they build expected values from plain graph ops
(tests/transducer_test.py:218-273: intersect -> forward_score -> subtract -> backward ->
emissions.grad()): scoring runs in wfst_graph_score (host, float64 accumulation), composition
gradients are scattered through the arc provenance the product's compose records.  This is the
host-side graph API for small graphs; the batched criteria never come through here."""
import ctypes
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from gtn_applications_b200 import _lib
from gtn_applications_b200 import graph as _pg

epsilon = _pg.epsilon


class Device:
    """gtn.Device(gtn.CPU): host graphs only (the GPU work happens inside the criteria)"""

    def __init__(self, kind=0, index=0):
        self.kind, self.index = kind, index


CPU = 0
CUDA = 1


class Graph(_pg.Graph):
    """gtn.Graph(calc_grad=True): a product host graph plus the tape fields"""

    def __init__(self, calc_grad=True, _handle=None):
        if isinstance(calc_grad, (Device,)) or (calc_grad in (CPU, CUDA) and not isinstance(calc_grad, bool)):
            calc_grad = True     # gtn.Graph(gtn.CPU) (tests/transducer_test.py:252): a device, not a flag
        super().__init__(calc_grad, _handle)
        self._inputs = ()
        self._grad_fn = None
        self._grad = None

    # ---- autograd surface
    def grad(self):
        if self._grad is None:
            raise RuntimeError("grad() before backward()")
        d = self.arrays()
        g = Graph(False)
        for s, a in zip(d["start"], d["accept"]):
            g.add_node(bool(s), bool(a))
        g.add_arcs(d["src"], d["dst"], d["ilabel"], d["olabel"], self._grad.astype(np.float32))
        return g

    def is_grad_available(self):
        return self._grad is not None

    def zero_grad(self):
        self._grad = None

    def _add_grad(self, delta):
        if not self.calc_grad:
            return
        delta = np.asarray(delta, dtype=np.float64).reshape(-1)
        self._grad = delta.copy() if self._grad is None else self._grad + delta


def _wrap(pg_graph, inputs=(), grad_fn=None, calc_grad=None):
    """a product graph (fresh handle) as a compat Graph on the tape"""
    h, pg_graph._h = pg_graph._h, None
    g = Graph.__new__(Graph)
    _pg.Graph.__init__(g, True, _handle=h)
    g._inputs, g._grad_fn, g._grad = tuple(inputs), grad_fn, None
    if calc_grad is None:
        calc_grad = any(i.calc_grad for i in inputs) if inputs else True
    g.calc_grad = calc_grad
    return g


def _scalar(value, inputs, grad_fn):
    g = Graph(any(i.calc_grad for i in inputs))
    g.add_node(True)
    g.add_node(False, True)
    g.add_arc(0, 1, 0, 0, float(value))
    g._inputs, g._grad_fn = tuple(inputs), grad_fn
    return g


# ---- graph ops ----------------------------------------------------------------------
def compose(first, second):
    out = _pg.compose(first, second)
    p1, p2 = out.provenance()

    def grad_fn(delta):
        for src, prov in ((first, p1), (second, p2)):
            if isinstance(src, Graph) and src.calc_grad:
                acc = np.zeros(src.num_arcs(), dtype=np.float64)
                ok = prov >= 0
                np.add.at(acc, prov[ok], delta[ok])
                src._add_grad(acc)

    return _wrap(out, [g for g in (first, second) if isinstance(g, Graph)], grad_fn)


intersect = compose


def _score(g, tropical):
    n = g.num_arcs()
    score = ctypes.c_float()
    arc_grad = np.zeros(max(n, 1), dtype=np.float32)
    _lib.check(_lib.lib().wfst_graph_score(g._h, int(tropical), ctypes.byref(score), arc_grad.ctypes.data))

    def grad_fn(delta):
        g._add_grad(float(delta[0]) * arc_grad[:n].astype(np.float64))

    return _scalar(score.value, [g] if isinstance(g, Graph) else [], grad_fn)


def forward_score(g):
    return _score(g, False)


def viterbi_score(g):
    return _score(g, True)


def viterbi_path(g):
    return _wrap(_pg.viterbi_path(g), calc_grad=False)


def negate(g):
    return _scalar(-g.item(), [g], lambda d: g._add_grad(-d))


def add(a, b):
    def grad_fn(d):
        a._add_grad(d)
        b._add_grad(d)
    return _scalar(a.item() + b.item(), [a, b], grad_fn)


def subtract(a, b):
    def grad_fn(d):
        a._add_grad(d)
        b._add_grad(-d)
    return _scalar(a.item() - b.item(), [a, b], grad_fn)


def backward(g, grad=None, retain_graph=False):
    """gtn.backward(g[, retain_graph]) / gtn.backward(g, grad, retain_graph)"""
    if isinstance(grad, bool):
        retain_graph, grad = grad, None
    seed = np.ones(g.num_arcs(), dtype=np.float64) if grad is None else \
        np.asarray(grad.weights_to_numpy(), dtype=np.float64)
    # reverse topological order of the tape below g
    order, seen = [], set()

    def visit(x):
        if id(x) in seen:
            return
        seen.add(id(x))
        for i in x._inputs:
            visit(i)
        order.append(x)

    visit(g)
    g._grad = seed
    for x in reversed(order):   # every consumer of a graph comes before the graph itself
        if x._grad_fn is not None and x._grad is not None:
            x._grad_fn(x._grad)
    if not retain_graph:
        for x in order:
            x._inputs, x._grad_fn = (), None


def remove(g, ilabel=epsilon, olabel=None):
    return _wrap(_pg.remove(g, ilabel, olabel), calc_grad=False)


def project_input(g):
    return _wrap(_pg.project_input(g), calc_grad=g.calc_grad)


def project_output(g):
    return _wrap(_pg.project_output(g), calc_grad=g.calc_grad)


def linear_graph(M, N, *args, **kwargs):
    return _wrap(_pg.linear_graph(M, N, *[a for a in args if not isinstance(a, (Device, Graph))], **kwargs))


def load(path):
    return _wrap(_pg.load(path))


def loadtxt(path):
    return _wrap(_pg.loadtxt(path))


save = _pg.save
savetxt = _pg.savetxt
equal = _pg.equal
isomorphic = _pg.isomorphic


def write_dot(g, path, isymbols=None, osymbols=None):
    d = g.arrays()
    with open(path, "w") as f:
        f.write("digraph FST {\n  rankdir = LR;\n")
        for n, (s, a) in enumerate(zip(d["start"], d["accept"])):
            f.write('  %d [label = "%d" shape = %s style = %s];\n' %
                    (n, n, "doublecircle" if a else "circle", "bold" if s else "solid"))
        lab = lambda t, x: "ε" if x == epsilon else (t[x] if t else str(x))  # noqa: E731
        for s, t, i, o, w in zip(d["src"], d["dst"], d["ilabel"], d["olabel"], d["weight"]):
            f.write('  %d -> %d [label = "%s:%s/%g"];\n' % (s, t, lab(isymbols, i), lab(osymbols, o), w))
        f.write("}\n")


_pool = None


def parallel_for(function, int_list):
    """gtn.parallel_for(fn, range(B)): host threads (the product's C++ calls release the GIL)"""
    global _pool
    items = list(int_list)
    if len(items) <= 1:
        for i in items:
            function(i)
        return
    if _pool is None:
        _pool = ThreadPoolExecutor()
    for r in list(_pool.map(function, items)):
        pass
