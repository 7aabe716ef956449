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
def other_configs(dev):
    """BASELINE.json configs[2] (ASG B=256,T=1000,C=30,L=176) and configs[3] (transducer, 1000 word
    pieces, B=64, T=1000, 150 pieces per utterance): ms per fwd+bwd step through the Functions."""
    import random
    import torch
    from gtn_applications_b200.criterions.asg import ASGLoss
    from gtn_applications_b200.criterions.transducer import Transducer

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / n

    g = torch.Generator().manual_seed(3)
    B, T, C, L = 256, 1000, 30, 176
    e = torch.randn(B, T, C, generator=g).to(dev).requires_grad_(True)
    tr = torch.randn(C + 1, C, generator=g).to(dev).requires_grad_(True)
    tg = [t for t in torch.randint(C, (B, L), generator=g)]

    def asg():
        e.grad = None
        tr.grad = None
        ASGLoss(e, tr, tg, "mean").backward()

    s3 = timed(asg, 10)
    out = {"asg_cfg3": {"ms_per_step": s3 * 1e3, "utterances_per_s": B / s3,
                        "what": "ASGLoss fwd+bwd, B=256 T=1000 C=30 L=176, randn emissions and transitions"}}
    del e, tr
    rnd = random.Random(0)
    letters = "abcdefghijklmnopqrstuvwxyz"
    pieces = sorted({"".join(rnd.choice(letters) for _ in range(rnd.randint(1, 4))) for _ in range(1400)})[:1000]
    pieces = sorted(set(pieces) | set(letters))
    g2i = {ch: i for i, ch in enumerate(letters)}
    Bt, NP = 64, 150
    crit = Transducer(pieces, g2i, blank="optional", allow_repeats=False, reduction="mean")
    x = torch.randn(Bt, T, len(pieces) + 1, generator=g).to(dev).requires_grad_(True)
    targets = [torch.tensor([g2i[c] for c in "".join(rnd.choice(pieces) for _ in range(NP))]) for _ in range(Bt)]

    def tdc():
        x.grad = None
        crit(x, targets).backward()

    s4 = timed(tdc, 3)
    out["transducer_cfg4"] = {"ms_per_step": s4 * 1e3, "utterances_per_s": Bt / s4,
                              "what": "Transducer module fwd+bwd (log_softmax + alignment graphs on host threads + "
                                      "lattice kernel), B=64 T=1000, %d word pieces (synthetic list), 150 pieces per "
                                      "utterance, blank optional, no repeats" % len(pieces)}
    return out


# ------------------------------------------------------------------ CPU baseline
def cpu_baseline_run(workload, budget_s, steps=None, warmup=1):
    """Times the oracle — the C++ restatement of the GTN CPU algorithm (materialised
    intersect + Kahn forward_score + tape backward) driven like criterions/ctc.py:31-94,
    batch-parallel on all host cores like gtn.parallel_for — on a bounded sample of the
    workload.  Returns utterances/sec and a description."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gtn
    import _gtn_oracle
    import ref_criterions as rc
    B, T, C, L = WORKLOADS[workload]
    cores = _gtn_oracle.pool_size()
    n = min(B, max(cores, 8))
    lp, tg = synth(workload, None, 0)
    e = lp[:n].numpy()
    t = tg[:n].tolist()
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, mean, n = cpu_baseline_run(args.workload, 0, steps=args.steps, warmup=max(1, min(args.warmup, 2)))
    B, T, C, L = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "utterances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s: CTC fwd+bwd B=%d T=%d C=%d L=%d; each step = %d utterances on the "
                               "host cores" % (args.workload, B, T, C, L, n)},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "utterances/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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
            "metric": METRIC, "value": B * world * args.steps / (total_ms * 1e-3),
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
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": kern_ms},
        }
        if module_line is not None:
            line["ctc_module_on_logits"] = module_line
        if other is not None:
            line["other_configs"] = other
        if world == 1 and not args.no_cpu_baseline:
            base, _, _ = cpu_baseline_run(args.workload, args.cpu_seconds)
            line["cpu_baseline"] = base
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
