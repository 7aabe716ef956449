"""Helpers for the committed golden fixtures (tests/golden/*.npz, produced by
tests/golden/make_golden.py from the reference's own criterions)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def unpack(flat, off):
    return [flat[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def graph_of(z, prefix):
    keys = ("start", "accept", "src", "dst", "ilabel", "olabel", "weight")
    return {k: z[prefix + "_" + k] for k in keys}


def through_log_softmax(logits, grad_lp):
    """d loss / d logits given d loss / d log_softmax(logits) (numpy, float64)."""
    x = np.asarray(logits, dtype=np.float64)
    g = np.asarray(grad_lp, dtype=np.float64)
    m = x.max(-1, keepdims=True)
    p = np.exp(x - m)
    p /= p.sum(-1, keepdims=True)
    return g - p * g.sum(-1, keepdims=True)


def log_softmax(x):
    x = np.asarray(x, dtype=np.float64)
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))
