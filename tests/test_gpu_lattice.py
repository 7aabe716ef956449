"""GPU parity: the generic acceptor lattice kernel (wfst_lattice_forward_backward,
packed CSR acceptors) against the float64 closed-form DP oracle on random
acceptors — scores, emission gradients and arc-weight gradients."""
import numpy as np
import pytest
import torch

import dp_numpy
from test_gpu_ctc import assert_close

pytestmark = pytest.mark.gpu


def random_acceptor(rng, N, A, C, weights=True):
    g = {"start": np.zeros(N, dtype=bool), "accept": np.zeros(N, dtype=bool),
         "src": rng.integers(0, N, A), "dst": rng.integers(0, N, A), "label": rng.integers(0, C, A)}
    g["start"][rng.integers(0, N, 2)] = True
    g["accept"][rng.integers(0, N, 3)] = True
    g["weight"] = rng.standard_normal(A).astype(np.float32) if weights else None
    return g


@pytest.mark.parametrize("B,T,C,N,A", [(3, 9, 5, 6, 20), (4, 33, 17, 40, 160), (2, 70, 1001, 300, 900),
                                       (5, 16, 8, 1, 3), (2, 21, 40, 2300, 7000), (2, 40, 6, 3, 40),
                                       (3, 300, 6, 40, 150), (2, 531, 17, 25, 90), (2, 129, 9, 12, 30),
                                       # 1025..2048 nodes: the wide-register cluster kernel (register slots
                                       # + tail arcs: random degrees reach 10 and more), several tiles
                                       (2, 150, 30, 1500, 5200), (2, 40, 1001, 1100, 3600), (2, 23, 300, 2048, 9000),
                                       # small dense acceptors (n-gram transition graphs): a warp per node in the
                                       # cluster kernel -- 1, 3 and 16 nodes per warp, lists longer than 32 x 4 arcs
                                       (2, 150, 30, 84, 6800), (3, 130, 12, 20, 500), (2, 70, 50, 300, 6000),
                                       (2, 150, 9, 12, 2400)])
def test_random_acceptors(B, T, C, N, A, lattice_kernel):
    from gtn_applications_b200.packing import PackedAcceptors
    from gtn_applications_b200.lattice import lattice_forward_backward
    rng = np.random.default_rng(B + T + C)
    graphs = [random_acceptor(rng, max(1, N - b), max(1, A - 3 * b), C) for b in range(B)]
    E = rng.standard_normal((B, T, C)).astype(np.float32)
    gs = rng.uniform(0.5, 2.0, B).astype(np.float32)
    packed = PackedAcceptors(graphs, "cuda")
    scores, gE, gW = lattice_forward_backward(torch.tensor(E, device="cuda"), packed,
                                              grad_scale=torch.tensor(gs, device="cuda"),
                                              want_grad_weights=True)
    scores, gE, gW = scores.cpu().numpy(), gE.cpu().numpy(), gW.cpu().numpy()
    for b, g in enumerate(graphs):
        Z, rE, rW = dp_numpy.acceptor_forward_backward(E[b], g["start"], g["accept"], g["src"], g["dst"],
                                                       g["label"], g["weight"])
        if not np.isfinite(Z):
            assert scores[b] == -np.inf and np.all(gE[b] == 0)
            continue
        assert abs(scores[b] - Z) <= 1e-4 * max(1.0, abs(Z))
        assert_close(gE[b], rE * gs[b])
        a0, a1 = packed.arc_offsets_host[b], packed.arc_offsets_host[b + 1]
        assert_close(gW[a0:a1], rW * gs[b])


def test_ctc_chain_through_csr_matches_closed_form_kernel(lattice_kernel):
    from gtn_applications_b200.packing import PackedAcceptors
    from gtn_applications_b200.lattice import lattice_forward_backward
    from gtn_applications_b200.criterions.ctc import CTCLoss
    rng = np.random.default_rng(5)
    B, T, C = 4, 60, 11
    tg = [rng.integers(0, C - 1, size=n).tolist() for n in (9, 0, 25, 14)]
    graphs = []
    for y in tg:
        st, ac, src, dst, lab, w = dp_numpy.ctc_acceptor(y, C - 1)
        graphs.append({"start": st, "accept": ac, "src": src, "dst": dst, "label": lab, "weight": None})
    E = torch.tensor(rng.standard_normal((B, T, C)), dtype=torch.float32, device="cuda")
    scores, gE, _ = lattice_forward_backward(E, PackedAcceptors(graphs, "cuda"))
    lp = E.clone().requires_grad_(True)
    loss = CTCLoss(lp, tg, C - 1, "none")
    loss.backward()
    assert abs(loss.item() + scores.mean().item()) <= 1e-5 * abs(loss.item())
    torch.testing.assert_close(-gE / B, lp.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("T", [12, 200])          # 200 frames: several tiles, the two-block kernel applies
def test_shared_graph_accumulates_weight_gradient(lattice_kernel, T):
    from gtn_applications_b200.packing import PackedAcceptors
    from gtn_applications_b200.lattice import lattice_forward_backward
    rng = np.random.default_rng(11)
    B, C = 6, 7
    g = random_acceptor(rng, 9, 40, C)
    E = rng.standard_normal((B, T, C)).astype(np.float32)
    packed = PackedAcceptors([g], "cuda")
    scores, gE, gW = lattice_forward_backward(torch.tensor(E, device="cuda"), packed,
                                              want_grad_weights=True, shared=True)
    want = np.zeros(40)
    for b in range(B):
        Z, rE, rW = dp_numpy.acceptor_forward_backward(E[b], g["start"], g["accept"], g["src"], g["dst"],
                                                       g["label"], g["weight"])
        assert abs(scores[b].item() - Z) <= 1e-4 * max(1.0, abs(Z))
        assert_close(gE[b].cpu().numpy(), rE)
        want += rW
    assert_close(gW.cpu().numpy(), want)


@pytest.mark.parametrize("B,T,C,N,A", [(3, 150, 6, 30, 110), (2, 300, 1001, 200, 700), (2, 150, 30, 1300, 4500)])
def test_final_weights_and_weight_gradients_agree_across_kernels(B, T, C, N, A):
    """Final weights (ABI 2) over many tiles: the two-block cluster kernel and the single-block
    shared-memory kernel against the generic kernel (itself checked against the reference's
    epsilon-graph fixtures in test_gpu_stc_transducer.py) — scores, emission, arc-weight and
    final-weight gradients."""
    from gtn_applications_b200 import _lib
    from gtn_applications_b200.packing import PackedAcceptors
    from gtn_applications_b200.lattice import lattice_forward_backward
    rng = np.random.default_rng(T + C)
    graphs = [random_acceptor(rng, N - b, A - 2 * b, C) for b in range(B)]
    E = torch.tensor(rng.standard_normal((B, T, C)).astype(np.float32), device="cuda")
    packed = PackedAcceptors(graphs, "cuda")
    fw = torch.tensor(rng.standard_normal(packed.num_nodes).astype(np.float32), device="cuda")
    gs = torch.tensor(rng.uniform(0.5, 2.0, B).astype(np.float32), device="cuda")
    out = {}
    for mode in (1, 2, 3, 4):
        old = _lib.lib().wfst_debug_force_generic_lattice(mode)
        try:
            out[mode] = [x.cpu().numpy() for x in lattice_forward_backward(
                E, packed, grad_scale=gs, want_grad_weights=True, final_weights=fw)]
        finally:
            _lib.lib().wfst_debug_force_generic_lattice(old)
    assert np.isfinite(out[1][0]).any()
    for mode in (2, 3, 4):
        for ref, got in zip(out[1], out[mode]):
            fin = np.isfinite(ref)
            assert np.array_equal(fin, np.isfinite(got))
            assert_close(got[fin], ref[fin])


def test_many_shared_acceptors_in_one_call_match_one_call_each(lattice_kernel):
    """wfst_lattice_forward_backward_many (every window against every kernel graph of
    ConvTransduce1D in one call) against one shared-graph call per acceptor: scores, the
    emission gradient summed over the acceptors, the arc-weight gradients of each."""
    import ctypes
    from gtn_applications_b200 import _lib, _runtime as rt
    from gtn_applications_b200.packing import PackedAcceptors
    from gtn_applications_b200.lattice import lattice_forward_backward
    rng = np.random.default_rng(3)
    B, T, C, K = 7, 20, 9, 4
    graphs = [random_acceptor(rng, 5 + 3 * k, 18 + 7 * k, C) for k in range(K)]
    E = torch.tensor(rng.standard_normal((B, T, C)).astype(np.float32), device="cuda")
    gs = torch.tensor(rng.uniform(0.5, 2.0, (K, B)).astype(np.float32), device="cuda")
    packed = [PackedAcceptors([g], "cuda") for g in graphs]
    want_s, want_w = [], []
    want_e = torch.zeros_like(E)
    for k, pk in enumerate(packed):
        s, _, w = lattice_forward_backward(E, pk, grad_scale=gs[k].contiguous(), want_grad_weights=True,
                                           shared=True, accumulate_into=want_e)
        want_s.append(s)
        want_w.append(w)
    L = _lib.lib()
    structs = (_lib.AcceptorBatch * K)(*[pk.struct() for pk in packed])
    narcs = [pk.num_arcs for pk in packed]
    flat = torch.empty(sum(narcs), dtype=torch.float32, device="cuda")
    ptrs = (ctypes.c_void_p * K)()
    pos = 0
    for k in range(K):
        ptrs[k] = flat.data_ptr() + 4 * pos
        pos += narcs[k]
    scores = torch.empty(K, B, dtype=torch.float32, device="cuda")
    got_e = torch.zeros_like(E)
    ws = rt.workspace(E.device, L.wfst_lattice_workspace_bytes(B, T, C, 0, max(pk.max_nodes for pk in packed)))
    _lib.check(L.wfst_lattice_forward_backward_many(
        E.data_ptr(), B, T, C, structs, K, gs.data_ptr(), scores.data_ptr(), got_e.data_ptr(), ptrs,
        ws.data_ptr(), ws.numel(), rt.stream_ptr(E.device)))
    torch.cuda.synchronize()
    assert torch.equal(scores, torch.stack(want_s))
    assert torch.equal(got_e, want_e)
    # summed over the items with float atomics: the order of the additions is not fixed
    assert_close(flat.cpu().numpy(), torch.cat(want_w).cpu().numpy())


def test_cross_launch_matches_one_call_per_acceptor():
    """wfst_lattice_forward_backward_cross (every emission item against every acceptor of a packed
    batch in ONE launch, emission and weight gradients added atomically) against one shared-graph
    call per acceptor; with the lean kernels switched off the entry refuses and launches nothing."""
    import ctypes
    from gtn_applications_b200 import _lib, _runtime as rt
    from gtn_applications_b200.packing import PackedAcceptors
    from gtn_applications_b200.lattice import lattice_forward_backward
    rng = np.random.default_rng(4)
    B, T, C, K = 9, 11, 9, 5
    graphs = [random_acceptor(rng, 4 + 2 * k, 12 + 6 * k, C) for k in range(K)]
    E = torch.tensor(rng.standard_normal((B, T, C)).astype(np.float32), device="cuda")
    gs = torch.tensor(rng.uniform(0.5, 2.0, (K, B)).astype(np.float32), device="cuda")
    want_s, want_w = [], []
    want_e = torch.zeros_like(E)
    for k, g in enumerate(graphs):
        s, _, w = lattice_forward_backward(E, PackedAcceptors([g], "cuda"), grad_scale=gs[k].contiguous(),
                                           want_grad_weights=True, shared=True, accumulate_into=want_e)
        want_s.append(s)
        want_w.append(w)
    L = _lib.lib()
    packed = PackedAcceptors(graphs, "cuda")
    st = packed.struct()
    scores = torch.empty(K, B, dtype=torch.float32, device="cuda")
    got_e = torch.zeros_like(E)
    got_w = torch.zeros(packed.num_arcs, dtype=torch.float32, device="cuda")
    ws = rt.workspace(E.device, L.wfst_lattice_workspace_bytes(K * B, T, C, 0, packed.max_nodes))
    args = (E.data_ptr(), B, T, C, ctypes.byref(st), gs.data_ptr(), scores.data_ptr(), got_e.data_ptr(),
            got_w.data_ptr(), ws.data_ptr(), ws.numel(), rt.stream_ptr(E.device))
    _lib.check(L.wfst_lattice_forward_backward_cross(*args))
    torch.cuda.synchronize()
    ws_, gs_ = torch.stack(want_s).cpu().numpy(), scores.cpu().numpy()
    fin = np.isfinite(ws_)
    assert np.array_equal(fin, np.isfinite(gs_)) and fin.any()
    assert_close(gs_[fin], ws_[fin])
    assert_close(got_e.cpu().numpy(), want_e.cpu().numpy())
    assert_close(got_w.cpu().numpy(), torch.cat(want_w).cpu().numpy())
    old = L.wfst_debug_force_generic_lattice(1)
    try:
        assert L.wfst_lattice_forward_backward_cross(*args) == -3
    finally:
        L.wfst_debug_force_generic_lattice(old)
