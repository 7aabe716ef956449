"""Generates tests/golden/*.npz by running the REFERENCE's own criterions
(/root/reference/criterions/*.py, imported unchanged) on the oracle `gtn` shim
(float32) with seeded inputs.  Build container only:

    python tests/golden/make_golden.py

The fixtures pin the oracle restatement (oracle/ref_criterions.py) and the CUDA
path on the GPU box, where /root/reference does not exist.  The literal
known-answer vectors of the reference's tests (gtn_ctc_test.py:48-80,
gtn_asg_test.py:25-105, gtn_stc_test.py:25-51, transducer_test.py:143-216) are
kept verbatim in tests/test_golden_literals.py instead.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _refload import load_reference  # noqa: E402

ctc, asg, stc, tr = load_reference()
import gtn  # noqa: E402  (oracle shim)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import ref_criterions as rc  # noqa: E402


def pack(targets):
    flat = np.array([t for tg in targets for t in tg], dtype=np.int32)
    off = np.zeros(len(targets) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(t) for t in targets])
    return flat, off


def graph_fields(prefix, g):
    return {prefix + "_" + k: v for k, v in rc.graph_arrays(g).items()}


def gen_ctc():
    out = {}
    cases = {
        # name: (B, T, C, target lengths, log_softmax?, reduction)
        "small_none": (3, 12, 6, [4, 0, 6], True, "none"),
        "small_mean": (3, 12, 6, [4, 0, 6], True, "mean"),
        "raw_mean": (4, 20, 9, [7, 1, 10, 3], False, "mean"),
        "repeats": (2, 15, 4, [6, 7], True, "none"),
        # BASELINE.json configs[0]: CTC B=4 T=150 C=28 L=20
        "cfg1_raw": (4, 150, 28, [20] * 4, False, "none"),
        "cfg1_lsm": (4, 150, 28, [20] * 4, True, "none"),
    }
    for name, (B, T, C, lens, lsm, red) in cases.items():
        torch.manual_seed(0)
        x = torch.randn(B, T, C)
        if name.startswith("cfg1"):
            tg = torch.randint(C - 2, (B, lens[0])).tolist()
        elif name == "repeats":
            tg = [[0, 0, 1, 1, 1, 2], [2, 2, 2, 0, 0, 1, 1]]
        else:
            tg = [torch.randint(C - 1, (n,)).tolist() for n in lens]
        e = (torch.log_softmax(x, 2) if lsm else x).detach().requires_grad_(True)
        loss = ctc.CTCLoss(e, tg, C - 1, red)
        loss.backward()
        flat, off = pack(tg)
        out[name + "_emissions"] = e.detach().numpy()
        out[name + "_targets"] = flat
        out[name + "_offsets"] = off
        out[name + "_blank"] = np.int32(C - 1)
        out[name + "_reduction"] = np.array(red)
        out[name + "_loss"] = np.float32(loss.item())
        out[name + "_grad"] = e.grad.numpy()
    g = ctc.CTCLossFunction.create_ctc_graph([3, 3, 1, 0, 0, 2], 5)
    out.update(graph_fields("graph", g))
    out["graph_in_order"] = np.array(g.in_order(), dtype=np.int32)
    out["graph_out_order"] = np.array(g.out_order(), dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "ctc.npz"), **out)


def gen_asg():
    out = {}
    for name, (B, T, C, lens, red) in {
        "small_none": (3, 10, 5, [3, 5, 1], "none"),
        "small_mean": (3, 10, 5, [3, 5, 1], "mean"),
        "mid_mean": (4, 40, 12, [9, 14, 2, 20], "mean"),
    }.items():
        torch.manual_seed(1)
        x = torch.randn(B, T, C, requires_grad=True)
        t = torch.randn(C + 1, C, requires_grad=True)
        tg = [torch.randint(C, (n,)).tolist() for n in lens]
        loss = asg.ASGLoss(x, t, tg, red)
        loss.backward()
        flat, off = pack(tg)
        out[name + "_emissions"] = x.detach().numpy()
        out[name + "_transitions"] = t.detach().numpy()
        out[name + "_targets"] = flat
        out[name + "_offsets"] = off
        out[name + "_reduction"] = np.array(red)
        out[name + "_loss"] = np.float32(loss.item())
        out[name + "_grad"] = x.grad.numpy()
        out[name + "_grad_transitions"] = t.grad.numpy()
    # module level: replabels + garbage (asg.py:191-209) and viterbi (asg.py:211-237)
    torch.manual_seed(2)
    crit = asg.ASG(4, num_replabels=2, use_garbage=True)
    crit.transitions.data = torch.randn_like(crit.transitions) * 0.5
    x = torch.randn(2, 14, crit.N, requires_grad=True)
    tg = [torch.tensor([0, 0, 0, 1, 2, 2]), torch.tensor([3, 3, 1])]
    loss = crit(x, tg)
    loss.backward()
    flat, off = pack([t.tolist() for t in tg])
    out["module_emissions"] = x.detach().numpy()
    out["module_transitions"] = crit.transitions.detach().numpy()
    out["module_targets"] = flat
    out["module_offsets"] = off
    out["module_loss"] = np.float32(loss.item())
    out["module_grad"] = x.grad.numpy()
    out["module_grad_transitions"] = crit.transitions.grad.numpy()
    vit = crit.viterbi(x.detach())
    vflat, voff = pack([v.tolist() for v in vit])
    out["module_viterbi"] = vflat
    out["module_viterbi_offsets"] = voff
    g = asg.ASGLossFunction.create_transitions_graph(torch.randn(4, 3))
    out.update(graph_fields("transgraph", g))
    g = asg.ASGLossFunction.create_force_align_graph([2, 0, 0, 1])
    out.update(graph_fields("falgraph", g))
    np.savez_compressed(os.path.join(HERE, "asg.npz"), **out)


def gen_stc():
    out = {}
    # Function level (stc.py:66-129)
    for name, (B, T, C, lens, prob, red) in {
        "fn_none": (3, 9, 5, [2, 0, 4], 0.6, "none"),
        "fn_mean": (2, 16, 6, [5, 3], 0.25, "mean"),
    }.items():
        torch.manual_seed(3)
        x = torch.randn(B, T, 2 * C, requires_grad=True)
        tg = [(1 + torch.randint(C - 1, (n,))).tolist() for n in lens]
        loss = stc.STCLoss(x, tg, prob, red)
        loss.backward()
        flat, off = pack(tg)
        out[name + "_emissions"] = x.detach().numpy()
        out[name + "_targets"] = flat
        out[name + "_offsets"] = off
        out[name + "_prob"] = np.float64(prob)
        out[name + "_reduction"] = np.array(red)
        out[name + "_loss"] = np.float32(loss.item())
        out[name + "_grad"] = x.grad.numpy()
    # Module level (stc.py:174-221): inputs are [T, B, C] log-probs
    torch.manual_seed(4)
    crit = stc.STC(0, p0=0.5, plast=0.9, thalf=3, reduction="mean")
    crit.train()
    logits = torch.randn(11, 3, 7, requires_grad=True)
    tg = [[3, 5], [1, 1, 6, 2], [4]]
    lp = torch.log_softmax(logits, 2)
    loss = crit(lp, tg)
    loss.backward()
    flat, off = pack(tg)
    out["module_logits"] = logits.detach().numpy()
    out["module_targets"] = flat
    out["module_offsets"] = off
    out["module_loss"] = np.float32(loss.item())
    out["module_grad_logits"] = logits.grad.numpy()
    g = stc.STCLossFunction.create_stc_graph([2, 1, 1], 4, 0.5)
    out.update(graph_fields("graph", g))
    np.savez_compressed(os.path.join(HERE, "stc.npz"), **out)


def gen_transducer():
    out = {}
    tokens = ["a", "b", "ab", "ba", "aba"]
    g2i = {"a": 0, "b": 1}
    tg = [[0, 1, 0], [1, 1, 0, 1, 0, 0], [0]]
    flat, off = pack(tg)
    out["wp_targets"] = flat
    out["wp_offsets"] = off
    for name, blank, rep in (("wp_none", "none", True), ("wp_opt", "optional", True),
                             ("wp_norep", "optional", False), ("wp_forced", "forced", True)):
        torch.manual_seed(5)
        C = len(tokens) + int(blank != "none")
        x = torch.randn(3, 13, C, requires_grad=True)
        crit = tr.Transducer(tokens, g2i, blank=blank, allow_repeats=rep, reduction="mean")
        loss = crit(x, tg)
        loss.backward()
        out[name + "_logits"] = x.detach().numpy()
        out[name + "_loss"] = np.float32(loss.item())
        out[name + "_grad_logits"] = x.grad.numpy()
        vit = crit.viterbi(x.detach())
        vf, vo = pack([v.tolist() for v in vit])
        out[name + "_viterbi"] = vf
        out[name + "_viterbi_offsets"] = vo
        out.update(graph_fields(name + "_tokens", crit.tokens))
        out.update(graph_fields(name + "_lexicon", crit.lexicon))
        # per-utterance alignment acceptor of utterance 1 (transducer.py:265-276)
        mine = rc.Transducer(gtn, tokens, g2i, blank=blank, allow_repeats=rep)
        mine.tokens.arc_sort(True)
        out.update(graph_fields(name + "_align1", mine.alignment_graph(tg[1])))
    # learned bigram transitions (ngram=2, epsilon </s> arcs; transducer.py:32-58)
    for name, ngram, blank, rep in (("ngram1", 1, "optional", False), ("ngram2", 2, "optional", False),
                                    ("ngram2_asg", 2, "none", True)):
        torch.manual_seed(6)
        N = 4
        toks = [(i,) for i in range(N)]
        gi = {i: i for i in range(N)}
        C = N + int(blank != "none")
        crit = tr.Transducer(toks, gi, ngram=ngram, blank=blank, allow_repeats=rep, reduction="mean")
        crit.transition_params.data = torch.randn_like(crit.transition_params) * 0.4
        x = torch.randn(2, 10, C, requires_grad=True)
        t2 = [[0, 1, 1, 3], [2, 0]]
        loss = crit(x, t2)
        loss.backward()
        f2, o2 = pack(t2)
        out[name + "_emissions"] = x.detach().numpy()
        out[name + "_targets"] = f2
        out[name + "_offsets"] = o2
        out[name + "_params"] = crit.transition_params.detach().numpy()
        out[name + "_loss"] = np.float32(loss.item())
        out[name + "_grad"] = x.grad.numpy()
        out[name + "_grad_params"] = crit.transition_params.grad.numpy()
        out.update(graph_fields(name + "_transitions", crit.transitions))
        vit = crit.viterbi(x.detach())
        vf, vo = pack([v.tolist() for v in vit])
        out[name + "_viterbi"] = vf
        out[name + "_viterbi_offsets"] = vo
    # loaded back-off transitions (tests/trans_backoff_test.txt; transducer_test.py:534-566)
    torch.manual_seed(7)
    N = 5
    toks = [(i,) for i in range(N)]
    gi = {i: i for i in range(N)}
    tgraph = gtn.loadtxt("/root/reference/tests/trans_backoff_test.txt")
    out.update(graph_fields("backoff_file", tgraph))
    crit = tr.Transducer(toks, gi, blank="optional", allow_repeats=False, transitions=tgraph)
    crit.transition_params.data = torch.randn_like(crit.transition_params) * 0.3
    x = torch.randn(2, 8, N + 1, requires_grad=True)
    t3 = [[0, 1, 0], [4, 4, 2]]
    loss = crit(x, t3)
    loss.backward()
    f3, o3 = pack(t3)
    out["backoff_emissions"] = x.detach().numpy()
    out["backoff_targets"] = f3
    out["backoff_offsets"] = o3
    out["backoff_params"] = crit.transition_params.detach().numpy()
    out["backoff_loss"] = np.float32(loss.item())
    out["backoff_grad"] = x.grad.numpy()
    out["backoff_grad_params"] = crit.transition_params.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "transducer.npz"), **out)


def gen_conv():
    """ConvTransduce1D (criterions/transducer.py:351-556): window x lexicon-kernel scores,
    their gradients w.r.t. the inputs and the (learned) kernel arc weights, forward-score
    and viterbi modes; plus the kernel graphs themselves."""
    out = {}
    lexicon = [(0, 0), (0, 1), (1, 0), (1, 1), (1,)]
    blank_idx, C = 2, 3
    for name, ks, stride, B, T, learn, vit, opt, spike, norm, scale in (
            ("fwd", 5, 3, 2, 8, False, False, True, False, "none", "none"),
            ("learn", 5, 2, 2, 9, True, False, True, False, "pre", "sqrt"),
            ("viterbi", 5, 3, 2, 8, True, True, True, False, "none", "none"),
            ("forced_spike", 7, 4, 1, 11, True, False, False, True, "post", "linear")):
        torch.manual_seed(11)
        layer = tr.ConvTransduce1D(lexicon, ks, stride, blank_idx, blank_optional=opt, learn_params=learn,
                                   scale=scale, normalize=norm, viterbi=vit, spike=spike)
        if learn:
            layer.kernel_params.data = torch.randn_like(layer.kernel_params) * 0.5
        x = torch.randn(B, T, C, requires_grad=True)
        y = layer(x)
        go = torch.randn_like(y)
        y.backward(go)
        out[name + "_inputs"] = x.detach().numpy()
        out[name + "_outputs"] = y.detach().numpy()
        out[name + "_grad_outputs"] = go.numpy()
        out[name + "_grad_inputs"] = x.grad.numpy()
        out[name + "_config"] = np.array([ks, stride, blank_idx, int(opt), int(learn), int(vit), int(spike)], dtype=np.int32)
        out[name + "_normalize"] = np.array(norm)
        out[name + "_scale"] = np.array(scale)
        if learn:
            out[name + "_params"] = layer.kernel_params.detach().numpy()
            out[name + "_grad_params"] = layer.kernel_params.grad.numpy()
    flat, off = pack([list(l) for l in lexicon])
    out["lexicon"] = flat
    out["lexicon_offsets"] = off
    for i, (tok, opt, spike) in enumerate((((0, 0), True, False), ((0, 1), False, False), ((0, 1), True, False),
                                           ((1, 0, 1), True, True), ((), True, False))):
        out["kg%d_token" % i] = np.array(tok, dtype=np.int32)
        out["kg%d_flags" % i] = np.array([int(opt), int(spike)], dtype=np.int32)
        out.update(graph_fields("kg%d" % i, tr.make_kernel_graph(list(tok), blank_idx, opt, spike=spike)))
    np.savez_compressed(os.path.join(HERE, "conv.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "conv":
        gen_conv()
        raise SystemExit(0)
    gen_ctc()
    gen_asg()
    gen_stc()
    gen_transducer()
    gen_conv()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
