#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: utterances/sec of CTC forward+backward at
B=256, T=1000, C=30 (L=176) on N B200s of one node.

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the CPU restatement of the GTN path

A step is one pass of the hot path (wfst_ctc_forward_backward: loss + [B,T,C]
gradient) over one batch of synthetic emissions already resident in HBM.  Every
rank owns B utterances (weak scaling); the only collective is one NCCL all-reduce of
the scalar loss per step.  Rank 0 prints ONE JSON line.

Timing: W >= 3 warm-up steps, then K steps bracketed by barrier + synchronize, timed
with CUDA events on the launching (current) stream, max over ranks.  The inputs
rotate over ROT distinct batches (ROT x 30.7 MB > the 126 MB L2), so no step finds its
inputs in L2.  Besides the contract's keys the line carries `ctc_module_on_logits`: the
step the reference's CTC module runs (log_softmax + loss + backward to the logits,
criterions/ctc.py:107) with the log-softmax fused into the kernel and as two steps.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, T, C, L)  — SURVEY.md §8 / BASELINE.md §2
    "ctc_cfg2": (256, 1000, 30, 176),
    "ctc_cfg1": (4, 150, 28, 20),
    "ctc_cfg5": (256, 1500, 80, 264),
}
METRIC = "utterances/sec CTC fwd+bwd (B=256,T=1000,C=30)"
ROT = 8


def metric_name(workload):
    B, T, C, L = WORKLOADS[workload]
    if workload == "ctc_cfg2":
        return METRIC
    return "utterances/sec CTC fwd+bwd (B=%d per GPU,T=%d,C=%d)" % (B, T, C)


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu summary."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed
    region runs (the profiling recipe's clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self._rec = threading.Event()
        self.ready = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            self.ready.set()
            while not self._stop_evt.is_set():
                if not self._rec.is_set():
                    time.sleep(0.001)
                    continue
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.004)
        except Exception as e:  # pragma: no cover
            self.reasons.add("nvml_unavailable:" + type(e).__name__)
            self.ready.set()

    def begin(self):
        self._rec.set()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def synth(workload, device, seed):
    """torch.manual_seed(seed); randn emissions -> log_softmax; randint(C-2) targets;
    blank C-1 (benchmarks/ctc_benchmark.py:22-24 with explicit seeds)."""
    import torch
    B, T, C, L = WORKLOADS[workload]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, C, generator=g)
    lp = torch.log_softmax(x, 2)
    tg = torch.randint(C - 2, (B, L), generator=g)
    return lp if device is None else lp.to(device), tg


# ------------------------------------------------------------------ other BASELINE configs
def reference_word_pieces():
    """benchmarks/transducer_benchmark.py:19-23 on the committed copy of the reference's token list"""
    with open(os.path.join(ROOT, "tests", "golden", "word_pieces_tokens_1000.txt"), "r") as fid:
        tokens = sorted([l.strip() for l in fid])
    graphemes = sorted(set(c for t in tokens for c in t))
    return tokens, {t: i for i, t in enumerate(graphemes)}


def roofline_of(alg_bytes, ms, traffic=None):
    peak, src = measured_peak_hbm()
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": src, "algorithmic_bytes_per_step": alg_bytes, "ms": ms}


def other_configs(dev):
    """BASELINE.json configs[2] (ASG B=256,T=1000,C=30,L=176), configs[3] (transducer on the
    reference's 1000-word-piece list, B=64, T=1000, 150 pieces per utterance) and the per-GPU shard
    of configs[4] (CTC B=256,T=1500,C=80,L=264): ms per fwd+bwd step through the Functions
    (time_utils.py:11-21 protocol + synchronise) and at the C ABI (CUDA events), each with its own
    roofline (algorithmic bytes 8*T*C per utterance)."""
    import random
    import torch
    from gtn_applications_b200 import _lib, _runtime as rt
    from gtn_applications_b200.criterions.asg import ASGLoss
    from gtn_applications_b200.criterions.transducer import Transducer
    L_ = _lib.lib()
    stream = torch.cuda.current_stream(dev)

    def timed(fn, n, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / n

    def event_timed(fn, n, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            fn()
        b.record(stream)
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / n

    out = {}
    g = torch.Generator().manual_seed(3)
    B, T, C, L = 256, 1000, 30, 176
    es = [torch.randn(B, T, C, generator=g).to(dev).requires_grad_(True) for _ in range(5)]   # 154 MB > L2
    tr = torch.randn(C + 1, C, generator=g).to(dev).requires_grad_(True)
    tgt = torch.randint(C, (B, L), generator=g)
    tg = [t.tolist()[0] for t in tgt.split(1)]          # list of lists, as benchmarks/asg_benchmark.py:23-24
    it = [0]

    def asg():
        e = es[it[0] % 5]
        it[0] += 1
        e.grad = None
        tr.grad = None
        ASGLoss(e, tr, tg, "mean").backward()

    s3 = timed(asg, 20)
    flat = tgt.reshape(-1).to(torch.int32).to(dev)
    off = (torch.arange(B + 1, dtype=torch.int32) * L).to(dev)
    gs = torch.full((B,), 1.0 / B, dtype=torch.float32, device=dev)
    o3 = torch.empty(B + 1, dtype=torch.float32, device=dev)
    ge, gt = torch.empty(B, T, C, device=dev), torch.empty(C + 1, C, device=dev)
    ws = rt.workspace(dev, L_.wfst_asg_workspace_bytes(B, T, C, L))
    trd = tr.detach()

    def asg_abi():
        ed = es[it[0] % 5].detach()
        it[0] += 1
        _lib.check(L_.wfst_asg_forward_backward(
            ed.data_ptr(), trd.data_ptr(), flat.data_ptr(), off.data_ptr(), B, T, C, L, gs.data_ptr(),
            o3.data_ptr(), o3[B:].data_ptr(), ge.data_ptr(), gt.data_ptr(), ws.data_ptr(), ws.numel(),
            stream.cuda_stream))

    a3 = event_timed(asg_abi, 20)
    alg3 = 8.0 * T * C * B + 8.0 * (C + 1) * C
    out["asg_cfg3"] = {"ms_per_step": s3 * 1e3, "utterances_per_s": B / s3, "abi_ms_per_step": a3,
                       "roofline": roofline_of(alg3, a3), "roofline_function_path": roofline_of(alg3, s3 * 1e3),
                       "what": "ASGLoss(inputs, transitions, list_of_lists, 'mean').backward() (benchmarks/"
                               "asg_benchmark.py:26-30), B=256 T=1000 C=30 L=176, randn emissions and "
                               "transitions; abi = wfst_asg_forward_backward, CUDA events"}
    del es, tr, ge
    # ---- configs[3]
    tokens, g2i = reference_word_pieces()
    rnd = random.Random(0)
    Bt, NP = 64, 150
    crit = Transducer(tokens, g2i, blank="optional", allow_repeats=False, reduction="mean")
    Ct = len(tokens) + 1
    x = torch.randn(Bt, T, Ct, generator=g).to(dev).requires_grad_(True)
    targets = [torch.tensor([g2i[l] for wp in (rnd.choice(tokens) for _ in range(NP)) for l in wp])
               for _ in range(Bt)]

    def tdc():
        x.grad = None
        crit(x, targets).backward()

    # the reference benchmark meets the same targets every iteration (transducer_benchmark.py:37-44,
    # time_utils.py): alignment graphs come from the LRU after the first step ("warm").  "cold":
    # the cache is emptied before every step, i.e. every step builds its 64 alignment graphs.
    def tdc_cold():
        from gtn_applications_b200.criterions import transducer as _tm
        _lib.check(L_.wfst_transducer_alignment_cache(-1, None, None))
        _tm._PACKED_LRU.clear()
        tdc()

    s4c = timed(tdc_cold, 5, warm=2)
    s4 = timed(tdc, 10, warm=2)
    alg4 = 8.0 * T * Ct * Bt
    out["transducer_cfg4"] = {"ms_per_step": s4 * 1e3, "utterances_per_s": Bt / s4,
                              "cold_ms_per_step": s4c * 1e3, "cold_utterances_per_s": Bt / s4c,
                              "roofline": roofline_of(alg4, s4 * 1e3, ncu_traffic("transducer_cfg4")),
                              "roofline_cold": roofline_of(alg4, s4c * 1e3, ncu_traffic("transducer_cfg4")),
                              "what": "Transducer module fwd+bwd (log_softmax + alignment graphs on host "
                                      "threads + lattice kernel), B=64 T=1000, the reference's %d word pieces "
                                      "(benchmarks/word_pieces_tokens_1000.txt), 150 pieces per utterance, blank "
                                      "optional, no repeats (transducer_benchmark.py:19-44); ms_per_step: targets "
                                      "repeat as in the reference benchmark (alignment graphs and the packed device batch from the LRU caches); "
                                      "cold_*: caches emptied before every step" % len(tokens)}
    del x
    # ---- n-gram transducers (benchmarks/transducer_benchmark.py:56-119: N=81 tokens, T=250, L=44, ngram=2;
    # there B=1 on CPU, here B=32): epsilon transition graph, composition + fold for the batch in the
    # host library (wfst_fold_transitions_batch), lattice kernels with final weights
    try:
        Nn, Tn, Ln, Bn = 81, 250, 44, 32
        toks = [(i,) for i in range(Nn)]
        gi = {i: i for i in range(Nn)}
        xn = torch.randn(Bn, Tn, Nn, generator=g).to(dev).requires_grad_(True)
        tgn = [t.squeeze() for t in torch.randint(Nn, size=(Bn, Ln), generator=g).split(1)]
        res = {}
        for name, kw in (("ngram_ctc", dict(ngram=2, blank="optional", allow_repeats=False, reduction="mean")),
                         ("ngram_asg", dict(ngram=2, reduction="mean"))):
            cn = Transducer(toks, gi, **kw).to(dev)

            def ng():
                xn.grad = None
                cn(xn, tgn).backward()

            res[name + "_ms_per_step"] = timed(ng, 5, warm=2) * 1e3
        res["what"] = ("Transducer(ngram=2) fwd+bwd through the module, B=32, T=250, 81 tokens, L=44 "
                       "(the shapes of transducer_benchmark.py:56-119, which runs B=1 on CPU)")
        out["transducer_ngram2"] = res
        del xn
    except Exception as exc:
        out["transducer_ngram2"] = {"error": repr(exc)[:200]}
    # ---- ConvTransduce1D: 500 lexicon entries of 1-4 letters against every window
    try:
        from gtn_applications_b200.criterions.transducer import ConvTransduce1D
        rl = random.Random(1)
        lex = [[rl.randrange(26) for _ in range(rl.randint(1, 4))] for _ in range(500)]
        conv = ConvTransduce1D(lex, kernel_size=9, stride=4, blank_idx=26).to(dev)
        xc = torch.randn(8, 200, 27, generator=g).to(dev).requires_grad_(True)

        def cv():
            xc.grad = None
            conv(torch.log_softmax(xc, 2)).sum().backward()

        out["conv_transduce1d"] = {
            "ms_per_step": timed(cv, 5, warm=2) * 1e3,
            "what": "ConvTransduce1D(500 lexicon entries of 1-4 letters, kernel_size=9, stride=4) fwd+bwd, B=8, "
                    "T=200, C=27 (384 windows x 500 kernel graphs; one launch per direction, "
                    "wfst_lattice_forward_backward_cross)"}
        del xc, conv
    except Exception as exc:
        out["conv_transduce1d"] = {"error": repr(exc)[:200]}
    # ---- STC: the module on [T, B, C] log-probabilities, acceptors of the batch from
    # the host library (wfst_stc_graphs), lattice kernel
    try:
        from gtn_applications_b200.criterions.stc import STC
        rs = random.Random(2)
        Bs, Ts, Cs, Ls = 32, 500, 30, 60
        stc = STC(0, 0.1, 0.9, 10000, "mean")
        xs = torch.randn(Ts, Bs, Cs, generator=g).to(dev).requires_grad_(True)
        tgs = [[rs.randrange(1, Cs) for _ in range(Ls)] for _ in range(Bs)]

        def st():
            xs.grad = None
            stc(torch.log_softmax(xs, 2), tgs).backward()

        out["stc"] = {"ms_per_step": timed(st, 10, warm=3) * 1e3,
                      "what": "STC(blank 0, p0 0.1, plast 0.9) fwd+bwd through the module, B=32, T=500, C=30, L=60 "
                              "(log_softmax + <star> features + STC acceptors of the batch built in the host library "
                              "+ lattice kernel)"}
        del xs
    except Exception as exc:
        out["stc"] = {"error": repr(exc)[:200]}
    # ---- configs[4], one GPU's shard
    B5, T5, C5, L5 = WORKLOADS["ctc_cfg5"]
    lp5, tg5 = synth("ctc_cfg5", dev, 7)
    flat5 = tg5.reshape(-1).to(torch.int32).to(dev)
    off5 = (torch.arange(B5 + 1, dtype=torch.int32) * L5).to(dev)
    gs5 = torch.full((B5,), 1.0 / B5, dtype=torch.float32, device=dev)
    o5 = torch.empty(B5 + 1, dtype=torch.float32, device=dev)
    g5 = torch.empty(B5, T5, C5, device=dev)
    ws5 = rt.workspace(dev, L_.wfst_ctc_workspace_bytes(B5, T5, C5, L5))

    def cfg5():
        _lib.check(L_.wfst_ctc_forward_backward(
            lp5.data_ptr(), flat5.data_ptr(), off5.data_ptr(), B5, T5, C5, C5 - 1, L5, gs5.data_ptr(),
            o5.data_ptr(), o5[B5:].data_ptr(), g5.data_ptr(), ws5.data_ptr(), ws5.numel(), stream.cuda_stream))

    a5 = event_timed(cfg5, 20)
    out["ctc_cfg5_shard"] = {"ms_per_step": a5, "utterances_per_s": B5 / (a5 * 1e-3),
                             "roofline": roofline_of(8.0 * T5 * C5 * B5, a5, ncu_traffic("ctc_cfg5")),
                             "what": "wfst_ctc_forward_backward on one GPU's shard of configs[4] (B=256 of 2048, "
                                     "T=1500, C=80, L=264), CUDA events; the 8-GPU run is `torchrun ... bench.py "
                                     "--gpus 8 --workload ctc_cfg5`"}
    return out


# ------------------------------------------------------------------ CPU baseline
def cpu_baseline_run(workload, budget_s, steps=None, warmup=1, full_batch=False):
    """Times the oracle — the C++ restatement of the GTN CPU algorithm (materialised
    intersect + Kahn forward_score + tape backward) driven like criterions/ctc.py:31-94,
    batch-parallel on all host cores like gtn.parallel_for.  With full_batch each pass is the
    workload's whole batch (the reference arm: same config as the GPU arm); otherwise a bounded
    sample of it (the cpu_baseline leg of the GPU line).  Returns utterances/sec and a description."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gtn
    import _gtn_oracle
    import ref_criterions as rc
    B, T, C, L = WORKLOADS[workload]
    cores = _gtn_oracle.pool_size()
    n = B if full_batch else min(B, max(cores, 8))
    lp, tg = synth(workload, None, 0)
    e = lp[:n].numpy()
    t = tg[:n].tolist()
    if full_batch and steps is not None:
        # keep the whole --steps K run within a few minutes: time one pass over a core-sized
        # sample first and shrink the per-step batch only if K full batches would not fit
        k = min(B, max(cores, 8))
        t0 = time.perf_counter()
        rc.ctc(gtn, e[:k], t[:k], C - 1, "none")
        est = (time.perf_counter() - t0) * B / k
        if est * (steps + warmup) > 360.0:
            n = max(k, int(B * 360.0 / (est * (steps + warmup))) // k * k)
            e, t = e[:n], t[:n]
    for _ in range(warmup):
        rc.ctc(gtn, e, t, C - 1, "none")
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        rc.ctc(gtn, e, t, C - 1, "none")
        times.append(time.perf_counter() - t0)
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() - t_start > budget_s or len(times) >= 50:
            break
    mean = sum(times) / len(times)
    return {
        "value": n / mean, "unit": "utterances/s", "cores": cores, "kind": "port",
        "sample": "%d utterances of %s (T=%d,C=%d,L=%d) x %d passes, fwd+bwd, "
                  "GTN-algorithm restatement (oracle/gtn_cpu.h) on %d threads"
                  % (n, workload, T, C, L, len(times), cores),
    }, mean, n


def torch_ctc_cpu(workload, budget_s=8.0):
    """The reference's other sanctioned CPU path (CTC(use_pt=True), criterions/ctc.py:109-121):
    log_softmax + torch.nn.functional.ctc_loss + backward on the host cores, whole batch."""
    import torch
    B, T, C, L = WORKLOADS[workload]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, C, generator=g).requires_grad_(True)
    tg = torch.randint(C - 2, (B, L), generator=g)

    def step():
        x.grad = None
        lp = torch.nn.functional.log_softmax(x, dim=2)
        loss = torch.nn.functional.ctc_loss(lp.permute(1, 0, 2), tg.reshape(-1), [T] * B, [L] * B, blank=C - 1,
                                            zero_infinity=True)
        loss.backward()

    step()
    times = []
    t_start = time.perf_counter()
    while time.perf_counter() - t_start < budget_s and len(times) < 10:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return {"value": B / mean, "unit": "utterances/s", "ms_per_step": mean * 1e3, "threads": torch.get_num_threads(),
            "what": "log_softmax + torch.nn.functional.ctc_loss + backward on the host (the reference's use_pt "
                    "path, ctc.py:109-121), B=%d T=%d C=%d L=%d, %d passes" % (B, T, C, L, len(times))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, mean, n = cpu_baseline_run(args.workload, 0, steps=args.steps, warmup=max(1, min(args.warmup, 2)),
                                     full_batch=True)
    B, T, C, L = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": base["value"], "unit": "utterances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s: CTC fwd+bwd B=%d T=%d C=%d L=%d; each step = %d utterances (%s) on the "
                               "host cores" % (args.workload, B, T, C, L, n,
                                               "the whole batch" if n == B else "a bounded sample of the batch"),
                   "same_config_as_gpu_arm": n == B},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "utterances/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_torch_cpu:
        try:
            line["torch_ctc_loss_cpu"] = torch_ctc_cpu(args.workload)
        except Exception as exc:
            line["torch_ctc_loss_cpu"] = {"error": repr(exc)[:200]}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from gtn_applications_b200 import _lib, _runtime as rt
    from gtn_applications_b200.criterions.ctc import CTCLoss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, C, L = WORKLOADS[args.workload]
    L_ = _lib.lib()

    # resident inputs: ROT distinct batches per rank
    batches = []
    for r in range(ROT):
        lp, tg = synth(args.workload, dev, 1000 * rank + r)
        flat = tg.reshape(-1).to(torch.int32).to(dev)
        off = (torch.arange(B + 1, dtype=torch.int32) * L).to(dev)
        batches.append((lp.contiguous(), flat, off, tg))
    gscale = torch.full((B,), 1.0 / (B * world), dtype=torch.float32, device=dev)
    NOUT = 4            # loss buffers in rotation: a step's all-reduce may still be in flight
    outs = [torch.empty(B + 1, dtype=torch.float32, device=dev) for _ in range(NOUT)]
    out = outs[0]
    pending = [None] * NOUT
    grad = torch.empty(B, T, C, dtype=torch.float32, device=dev)
    ws = rt.workspace(dev, L_.wfst_ctc_workspace_bytes(B, T, C, L))
    stream = torch.cuda.current_stream(dev)

    def step(i):
        lp, flat, off, _ = batches[i % ROT]
        o = outs[i % NOUT]
        if pending[i % NOUT] is not None:      # the all-reduce that last used this buffer
            pending[i % NOUT].wait()
            pending[i % NOUT] = None
        _lib.check(L_.wfst_ctc_forward_backward(
            lp.data_ptr(), flat.data_ptr(), off.data_ptr(), B, T, C, C - 1, L, gscale.data_ptr(),
            o.data_ptr(), o[B:].data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(),
            stream.cuda_stream))
        if world > 1:
            # the scalar loss reduce runs on NCCL's stream behind this step's kernel; the next
            # step's kernel does not wait for it (it is waited for when its buffer comes round
            # again and before the timed region ends)
            pending[i % NOUT] = dist.all_reduce(o[B:], op=dist.ReduceOp.SUM, async_op=True)

    def drain():
        for k in range(NOUT):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(timeout=5)
    launches0 = _lib.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    start.record(stream)
    for i in range(args.steps):
        evs[i][0].record(stream)
        step(i)
        evs[i][1].record(stream)
    drain()                      # every step's loss reduce is complete inside the timed region
    end.record(stream)
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    total_ms = start.elapsed_time(end)
    kern_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = float(t[0]), float(t[1])
    loss_value = float(outs[(args.steps - 1) % NOUT][B].item())

    # ---- e2e: the user-facing call with HOST inputs (pinned), copies inside the timed region.
    # Like a training loop with a prefetching loader, the copies of the next steps' emissions are
    # issued on a second stream before step i's loss is read back, so transfer and compute overlap;
    # every step still copies its own inputs and reads its own result inside the timed region.
    # emissions and targets ([B, L] int32) of 4 batches in pinned host memory
    host = [(b[0].cpu().pin_memory(), b[3].to(torch.int32).pin_memory()) for b in batches[:4]]
    e2e_steps = max(3, min(args.steps, 100))
    copy_stream = torch.cuda.Stream(dev)
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_losses = []

    def prefetch(i):
        lp_h, tg_h = host[i % len(host)]
        with torch.cuda.stream(copy_stream):
            # all of a step's inputs cross PCIe on the copy stream: a small copy on the compute
            # stream would queue behind the bulk copies in flight and hold the kernel back
            tg = tg_h.to(dev, non_blocking=True)
            lp_d = lp_h.to(dev, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return lp_d, tg, ready

    def e2e_run(n):
        # two batches in flight on the copy stream, like a loader with prefetch depth 2: the copy
        # engine then never waits for the Python side of a step (measured: issuing the next copy
        # only after the step's launches left a 0.2-0.3 ms bubble per step; raw pinned H2D of one
        # batch is 0.56 ms, tools/h2d_bandwidth.py)
        queue = [prefetch(k) for k in range(min(2, n))]
        pending = None
        for i in range(n):
            lp_d, tg, ready = queue.pop(0)
            if i + 2 < n:
                queue.append(prefetch(i + 2))
            stream.wait_event(ready)
            lp_d.record_stream(stream)
            tg.record_stream(stream)
            lp_d.requires_grad_(True)
            loss = CTCLoss(lp_d, tg, C - 1, "none")
            loss.backward()
            # device -> host read of this step's result into pinned memory; the host consumes it
            # one step later (a training loop that logs the loss with one step of lag), so the
            # Python side of step i+1 overlaps the kernel of step i
            hbuf = loss_host[i % 2]
            hbuf.copy_(loss.detach(), non_blocking=True)
            done = torch.cuda.Event()
            done.record(stream)
            if pending is not None:
                pending[1].synchronize()
                e2e_losses.append(float(pending[0]))
            pending = (hbuf, done)
        pending[1].synchronize()
        e2e_losses.append(float(pending[0]))

    e2e_run(4)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])

    # ---- fallback rate of the timed batches: utterances the scaled kernel flagged and the
    # log-semiring kernel recomputed (outside the timed region: one more step per batch + flag read)
    import numpy as np
    flagged = 0
    for r in range(ROT):
        step(r)
        drain()
        torch.cuda.synchronize(dev)
        fl = np.zeros(B, dtype=np.int32)
        _lib.check(L_.wfst_debug_ctc_hazards(ws.data_ptr(), B, T, C, L, fl.ctypes.data))
        flagged += int((fl > 0).sum())
    fallback_rate = flagged / float(ROT * B)

    # ---- steep emissions: log_softmax(s * randn) with targets unrelated to the scores is outside
    # float32's range for the scaled kernel at s >= 2; those utterances are recomputed by the
    # float64 log-semiring kernel (ctc_exact.cu).  What that costs, reported beside the headline.
    steep = None
    if world == 1 and not args.no_comparators:
        steep = {}
        for sc in (2.0, 4.0):
            g = torch.Generator(device="cpu").manual_seed(77)
            lp_s = torch.log_softmax(torch.randn(B, T, C, generator=g).to(dev) * sc, 2).contiguous()
            flat, off = batches[0][1], batches[0][2]

            def steep_step():
                _lib.check(L_.wfst_ctc_forward_backward(
                    lp_s.data_ptr(), flat.data_ptr(), off.data_ptr(), B, T, C, C - 1, L, gscale.data_ptr(),
                    out.data_ptr(), out[B:].data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(), stream.cuda_stream))
            for _ in range(3):
                steep_step()
            torch.cuda.synchronize(dev)
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(10):
                steep_step()
            b2.record(stream)
            torch.cuda.synchronize(dev)
            fl = np.zeros(B, dtype=np.int32)
            _lib.check(L_.wfst_debug_ctc_hazards(ws.data_ptr(), B, T, C, L, fl.ctypes.data))
            ms_s = a.elapsed_time(b2) / 10
            steep["scale_%g" % sc] = {"ms_per_step": ms_s, "utterances_per_s": B / (ms_s * 1e-3),
                                       "fallback_rate": float((fl > 0).mean())}
            del lp_s
        steep["what"] = ("wfst_ctc_forward_backward on log_softmax(scale * randn) emissions with unrelated targets "
                         "(same B, T, C, L): scaled kernel + float64 fallback kernel for the flagged utterances; CUDA events")

    # ---- function path: the reference benchmark's own call (benchmarks/ctc_benchmark.py:23-29,
    # time_utils.py:11-21): device-resident emissions, targets as a Python list of lists,
    # CTCLoss(inputs, tgt, N - 1).backward(), 5 warm-ups, wall clock over the iterations with a
    # synchronise at both ends (the path is asynchronous)
    fp_inputs = [b[0].clone().requires_grad_(True) for b in batches[:5]]     # 154 MB > L2
    fp_tgt = [t.tolist()[0] for t in batches[0][3].split(1)]

    def fp_func(i):
        x = fp_inputs[i % 5]
        x.grad = None
        op = CTCLoss(x, fp_tgt, C - 1)
        op.backward()

    for i in range(5):
        fp_func(i)
    torch.cuda.synchronize(dev)
    fp_iters = max(5, min(args.steps, 100))
    t0 = time.perf_counter()
    for i in range(fp_iters):
        fp_func(i)
    torch.cuda.synchronize(dev)
    fp_s = (time.perf_counter() - t0) / fp_iters
    t = torch.tensor([fp_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fp_s = float(t[0])
    del fp_inputs

    # ---- the existing GPU kernel on the same box: torch.nn.functional.ctc_loss, the reference's
    # use_pt path (criterions/ctc.py:109-121), forward + backward on the same shapes, CUDA events
    torch_gpu = None
    if world == 1 and not args.no_comparators:
        try:
            xs_t = [torch.randn(B, T, C, device=dev, generator=torch.Generator(device=dev).manual_seed(70 + r))
                    .requires_grad_(True) for r in range(5)]
            lps_t = [b[0].clone().requires_grad_(True) for b in batches[:5]]
            tgt_t = batches[0][3].to(dev)
            ilen, tlen = [T] * B, [L] * B

            def pt_logits(i):      # what CTC(use_pt=True).forward + backward run (ctc.py:107,114-121)
                x = xs_t[i % 5]
                x.grad = None
                lp = torch.nn.functional.log_softmax(x, dim=2)
                torch.nn.functional.ctc_loss(lp.permute(1, 0, 2), tgt_t, ilen, tlen, blank=C - 1,
                                             zero_infinity=True).backward()

            def pt_logprobs(i):    # the criterion alone on log-probabilities (the headline's inputs)
                lp = lps_t[i % 5]
                lp.grad = None
                torch.nn.functional.ctc_loss(lp.permute(1, 0, 2), tgt_t, ilen, tlen, blank=C - 1,
                                             zero_infinity=True).backward()

            def ev_time(fn, n):
                for i in range(3):
                    fn(i)
                torch.cuda.synchronize(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for i in range(n):
                    fn(i)
                b.record(stream)
                torch.cuda.synchronize(dev)
                return a.elapsed_time(b) / n

            n_pt = max(3, min(args.steps, 20))
            ms_lp, ms_x = ev_time(pt_logprobs, n_pt), ev_time(pt_logits, n_pt)
            torch_gpu = {"from_log_probs_ms_per_step": ms_lp, "from_logits_ms_per_step": ms_x,
                         "utterances_per_s_from_log_probs": B / (ms_lp * 1e-3),
                         "utterances_per_s_from_logits": B / (ms_x * 1e-3), "steps": n_pt,
                         "what": "torch.nn.functional.ctc_loss fwd+bwd on this GPU (native CUDA kernel; cuDNN's "
                                 "needs blank=0), reduction='mean', zero_infinity=True as criterions/ctc.py:114-121; "
                                 "from_logits adds log_softmax and its backward (ctc.py:107); CUDA events, inputs "
                                 "rotate over 5 batches"}
            del xs_t, lps_t
        except Exception as exc:
            torch_gpu = {"error": repr(exc)[:200]}
        torch.cuda.empty_cache()     # ctc_loss keeps [B, T, 2L+1] alpha / beta buffers (2 x 361 MB) cached

    # ---- CTC on logits (device resident), event-timed around the C-ABI calls: the fused entry
    # point against torch.log_softmax + wfst_ctc_forward_backward + the softmax backward autograd
    # would run (the step criterions/ctc.py:107 + CTCLoss + backward amount to)
    module_line = None
    if world == 1 and L_.wfst_ctc_logits_supported(B, T, C, L):
        xs = [torch.randn(B, T, C, device=dev, generator=torch.Generator(device=dev).manual_seed(50 + r))
              for r in range(ROT)]
        flat0, off0 = batches[0][1], batches[0][2]
        ws2 = rt.workspace(dev, L_.wfst_ctc_logits_workspace_bytes(B, T, C, L))

        def fused(i):
            _lib.check(L_.wfst_ctc_logits_forward_backward(
                xs[i % ROT].data_ptr(), flat0.data_ptr(), off0.data_ptr(), B, T, C, C - 1, L, gscale.data_ptr(),
                out.data_ptr(), out[B:].data_ptr(), grad.data_ptr(), ws2.data_ptr(), ws2.numel(), stream.cuda_stream))

        def two_step(i):
            lp = torch.log_softmax(xs[i % ROT], 2)
            _lib.check(L_.wfst_ctc_forward_backward(
                lp.data_ptr(), flat0.data_ptr(), off0.data_ptr(), B, T, C, C - 1, L, gscale.data_ptr(),
                out.data_ptr(), out[B:].data_ptr(), grad.data_ptr(), ws2.data_ptr(), ws2.numel(), stream.cuda_stream))
            return grad - torch.exp(lp) * grad.sum(2, keepdim=True)

        def time_it(fn, n):
            for i in range(3):
                fn(i)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(n):
                fn(i)
            b.record(stream)
            torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / n

        n_mod = max(3, min(args.steps, 50))
        f_ms, t_ms = time_it(fused, n_mod), time_it(two_step, n_mod)
        module_line = {"fused_ms_per_step": f_ms, "two_step_ms_per_step": t_ms,
                       "fused_utterances_per_s": B / (f_ms * 1e-3), "steps": n_mod,
                       "what": "CTC on [B,T,C] logits in HBM, CUDA events around the calls: "
                               "wfst_ctc_logits_forward_backward vs torch.log_softmax + "
                               "wfst_ctc_forward_backward + softmax backward (torch ops)"}

    # ---- the other BASELINE configs (ASG cfg3, transducer cfg4), a few steps each through the
    # Functions (wall clock incl. Python, CUDA-synchronised): reported beside the headline, not
    # part of it
    other = None
    if world == 1 and args.workload == "ctc_cfg2" and not args.no_other_configs:
        try:
            other = other_configs(dev)
        except Exception as exc:  # never let the side measurements break the bench line
            other = {"error": repr(exc)[:200]}

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        alg_bytes = 8.0 * T * C * B            # read E once + write grad once
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": metric_name(args.workload), "value": B * world * args.steps / (total_ms * 1e-3),
            "unit": "utterances/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "%s: CTC fwd+bwd (loss + [B,T,C] grad), B=%d per GPU, T=%d, C=%d, L=%d, "
                            "log_softmax(randn) emissions, blank=C-1" % (args.workload, B, T, C, L),
                "global_batch": B * world, "parallelism": "batch-sharded x%d, 1 scalar all-reduce/step" % world,
                "l2": "inputs rotate over %d resident batches (%.0f MB) > 126 MB L2" % (
                    ROT, ROT * B * T * C * 4 / 1e6),
                "loss": loss_value,
            },
            "clocks": clocks,
            "e2e": {"value": B * world * e2e_steps / e2e_s, "unit": "utterances/s",
                    "h2d_bytes_per_step": B * T * C * 4 + B * L * 4,
                    "d2h_bytes_per_step": 4,
                    "path": "CTCLoss(emissions, targets).backward() with both copied pinned host -> cuda on a copy stream every step (two steps in flight); every step's loss copied to pinned host memory and read by the host one step later"},
            "gpu_launches": int(launches),
            "fallback_rate": fallback_rate,
            "kernel": ("ctc_chain_kernel (csrc/ctc_chain.cu): one block per utterance, both time directions packed in "
                       "FP32 pairs, chain split over warps" if args.workload == "ctc_cfg5" else
                       "ctc_tick_kernel<6,2> (csrc/ctc_tick.cu): one block per utterance, one time direction per warp "
                       "set, chain split over 2 warps, roles (producer / live / recompute / reduce) advance one "
                       "8-frame step per tick behind one named barrier") +
                      "; flagged utterances -> ctc_exact_kernel (float64)",
            "function_path": {
                "ms_per_step": fp_s * 1e3, "value": B * world / fp_s, "unit": "utterances/s", "iterations": fp_iters,
                "what": "CTCLoss(inputs, list_of_lists, C-1).backward() exactly as benchmarks/ctc_benchmark.py:"
                        "23-29 with device-resident emissions (5 batches in rotation), time_utils.py protocol "
                        "(5 warm-ups, wall clock) + synchronise; includes walking the %d Python ints of the "
                        "targets every call (csrc/pytargets.c)" % (B * L)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": kern_ms},
        }
        if steep is not None:
            line["steep_emissions"] = steep
        if torch_gpu is not None:
            line["torch_ctc_loss_gpu"] = torch_gpu
        if module_line is not None:
            line["ctc_module_on_logits"] = module_line
        if other is not None:
            line["other_configs"] = other
        if world == 1 and not args.no_cpu_baseline:
            base, _, _ = cpu_baseline_run(args.workload, args.cpu_seconds)
            line["cpu_baseline"] = base
            if not args.no_torch_cpu:
                try:
                    line["torch_ctc_loss_cpu"] = torch_ctc_cpu(args.workload, 6.0)
                except Exception as exc:
                    line["torch_ctc_loss_cpu"] = {"error": repr(exc)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ctc_cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-torch-cpu", action="store_true")
    ap.add_argument("--no-comparators", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
