"""The reference's own callers, unchanged, on the product ("train.py and
benchmarks/{ctc,asg,transducer}_benchmark.py call it unchanged"): verbatim copies of
benchmarks/*.py (tests/golden/ref_benchmarks/) are executed with gtn_applications_b200/compat/
on the path — `utils.CTCLoss`, `utils.ASGLoss`, `transducer.Transducer`, `gtn` all resolve to
the product; nothing under oracle/ is importable from these processes."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "ref_benchmarks")
COMPAT = os.path.join(ROOT, "gtn_applications_b200", "compat")


def run_benchmark(tmp_path, script, *argv, timeout=600):
    bdir = tmp_path / "benchmarks"
    bdir.mkdir(exist_ok=True)
    for f in ("ctc_benchmark.py", "asg_benchmark.py", "transducer_benchmark.py", "time_utils.py"):
        shutil.copy(os.path.join(FIX, f), bdir / f)
    shutil.copy(os.path.join(ROOT, "tests", "golden", "word_pieces_tokens_1000.txt"), bdir / "word_pieces_tokens_1000.txt")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([COMPAT, ROOT])
    # guard: the checker must not be reachable from the caller's process
    probe = subprocess.run([sys.executable, "-c", "import gtn, sys; print(gtn.__file__); "
                            "assert 'oracle' not in gtn.__file__; assert not any('oracle' in p for p in sys.path)"],
                           cwd=str(bdir), env=env, capture_output=True, text=True, timeout=120)
    assert probe.returncode == 0, probe.stderr
    r = subprocess.run([sys.executable, script, *[str(a) for a in argv]], cwd=str(bdir), env=env,
                       capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


@pytest.mark.gpu
def test_reference_ctc_benchmark_runs_unchanged(tmp_path):
    out = run_benchmark(tmp_path, "ctc_benchmark.py", 8)
    assert '"ctc fwd + bwd" took' in out


@pytest.mark.gpu
def test_reference_asg_benchmark_runs_unchanged(tmp_path):
    out = run_benchmark(tmp_path, "asg_benchmark.py", 8)
    assert '"asg fwd + bwd" took' in out


@pytest.mark.gpu
def test_reference_transducer_benchmark_runs_unchanged(tmp_path):
    out = run_benchmark(tmp_path, "transducer_benchmark.py", 2, timeout=1500)
    for name in ("word decomps fwd + bwd", "word decomps viterbi", "ctc fwd + bwd, ngram=2", "asg viterbi, ngram=2"):
        assert '"%s" took' % name in out, out
    rec = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(rec, exist_ok=True)
        with open(os.path.join(rec, "r2_ref_transducer_benchmark.txt"), "w") as f:
            f.write(out)
    except OSError:
        pass


def test_compat_gtn_scoring_ops_match_closed_forms():
    """forward_score / viterbi_score / subtract / backward on host graphs (what
    tests/transducer_test.py:218-273 builds its expected values from), against torch autograd."""
    import numpy as np
    import torch
    sys.path.insert(0, COMPAT)
    try:
        for m in [k for k in sys.modules if k == "gtn" or k.startswith("gtn.")]:
            del sys.modules[m]
        import gtn
        assert "compat" in gtn.__file__
        T, N = 6, 4
        torch.manual_seed(0)
        scores = torch.randn(1, T, N)
        al = gtn.Graph(False)
        al.add_node(True)
        lab = [0, 1, 0]
        for k, l in enumerate(lab):
            al.add_node(False, k == len(lab) - 1)
            al.add_arc(k, k + 1, l)
            al.add_arc(k + 1, k + 1, l)
        em = gtn.linear_graph(T, N, gtn.Graph(gtn.CPU), True)
        em.set_weights(scores.data_ptr())
        loss = gtn.subtract(gtn.forward_score(em), gtn.forward_score(gtn.intersect(em, al)))
        gtn.backward(loss)
        got = em.grad().weights_to_numpy().reshape(T, N)
        s = scores[0].double().requires_grad_(True)
        ninf = torch.tensor(-float("inf"), dtype=torch.double)
        a = [s[0, lab[0]], ninf, ninf]
        for t in range(1, T):
            a = [(a[k] if k == 0 else torch.logsumexp(torch.stack([a[k], a[k - 1]]), 0)) + s[t, lab[k]] for k in range(3)]
        want = torch.logsumexp(s, 1).sum() - a[2]
        want.backward()
        assert abs(loss.item() - want.item()) < 1e-5
        np.testing.assert_allclose(got, s.grad.numpy(), atol=1e-6)
        assert abs(gtn.viterbi_score(em).item() - scores[0].max(1).values.sum().item()) < 1e-5
        vp = gtn.viterbi_path(em)
        assert vp.labels_to_list() == scores[0].argmax(1).tolist()
        assert abs(gtn.negate(gtn.forward_score(em)).item() + torch.logsumexp(scores[0], 1).sum().item()) < 1e-5
    finally:
        sys.path.remove(COMPAT)
        for m in [k for k in sys.modules if k == "gtn" or k.startswith("gtn.")]:
            del sys.modules[m]
