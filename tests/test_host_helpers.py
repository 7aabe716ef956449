"""Host-side helpers of the criterion modules (no GPU): batched replabel unpacking against the
per-sequence restatement of asg.py:35-49, ragged target flattening."""
import numpy as np
import torch

from gtn_applications_b200 import _runtime as rt
from gtn_applications_b200.criterions.asg import pack_replabels, unpack_replabels, unpack_replabels_batch


def test_unpack_replabels_batch_matches_per_sequence_unpacking():
    rng = np.random.default_rng(0)
    for R in (1, 2, 3):
        for _ in range(100):
            B = int(rng.integers(1, 6))
            counts = rng.integers(0, 12, B)
            seqs = [rng.integers(0, R + 4, c).tolist() for c in counts]
            flat = np.array([x for s in seqs for x in s], dtype=np.int32)
            got = unpack_replabels_batch(flat, counts, R)
            assert len(got) == B
            for s, g in zip(seqs, got):
                assert g.dtype == torch.int32
                assert unpack_replabels(s, R) == g.tolist()


def test_unpack_replabels_batch_inverts_pack_replabels():
    rng = np.random.default_rng(1)
    for R in (1, 2):
        seqs = [rng.integers(0, 5, n).tolist() for n in (0, 1, 7, 30)]
        packed = [pack_replabels(s, R) for s in seqs]
        flat = np.array([x for p in packed for x in p], dtype=np.int32)
        got = unpack_replabels_batch(flat, [len(p) for p in packed], R)
        assert [g.tolist() for g in got] == seqs


def test_flatten_targets_host_accepts_lists_tensors_and_matrices():
    lists = [[3, 1, 2], [], [5]]
    flat, offs = rt.flatten_targets_host(lists)
    assert flat.dtype == np.int32 and flat.tolist() == [3, 1, 2, 5] and offs.tolist() == [0, 3, 3, 4]
    flat, offs = rt.flatten_targets_host([torch.tensor(t, dtype=torch.int64) for t in lists])
    assert flat.dtype == np.int32 and flat.tolist() == [3, 1, 2, 5] and offs.tolist() == [0, 3, 3, 4]
    flat, offs = rt.flatten_targets_host(torch.tensor([[1, 2], [3, 4]]))
    assert flat.tolist() == [1, 2, 3, 4] and offs.tolist() == [0, 2, 4]
