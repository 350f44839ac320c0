#!/usr/bin/env python
"""bench.py - Glow fwd+bwd training throughput (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg1..cfg5]
                  [--precision auto|fp32|fp16x3|bf16x3|bf16] [--dp abi|explicit|torch]

One "step" = forward -> Gaussian NLL gradient (dZ = Z/B) -> memory-efficient backward (+ gradient average over the
ranks when N > 1); the optimizer is the caller's (Flux) and is excluded, as in SURVEY 8(d).  The default workload is
BASELINE configs[1] (cfg2): NetworkGlow(3, 256, L=3, K=16; split_scales) on 256x256x3, GLOBAL batch 64 sharded over
the N ranks (strong scaling).  The other BASELINE configurations are selectable with --config (cfg1 NetworkGlow
64x64x1, cfg3 NetworkConditionalGlow 64x64, cfg4 NetworkMultiScaleHINT 128x128x2, cfg5 NetworkGlow3D 64^3).
Prints ONE JSON line on rank 0.

Multi-GPU (--dp abi, the default): the library's own data-parallel plane (include/inb200.h): an ncclComm_t created
through the C ABI, ActNorm initialised from the statistics of the GLOBAL batch, every scale's gradients averaged on the
communicator's stream while the remaining scales of the backward pass still run.  --dp explicit = one
inb_allreduce_grads call after backward, --dp torch = torch.distributed.all_reduce on the flat gradient buffer.

`--impl reference` times the reference's algorithm on the host cores: the Julia reference cannot run in this image (no
Julia), so it is the torch-CPU oracle restatement (oracle/*.py, kind "port").
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "glow_fwd_bwd_samples_per_sec"
UNIT = "samples/s"

CONFIGS = {
    # BASELINE configs[1] (memory_usage_invertiblenetworks.jl:54-68)
    "cfg2": dict(kind="glow", n_in=3, nh=256, L=3, K=16, sp=(256, 256), gb=64, precision="fp16x3",
                 label="NetworkGlow(3,256,L=3,K=16;split_scales=true) on 256x256x3"),
    # BASELINE configs[0] (examples/networks/network_glow.jl:20-26)
    "cfg1": dict(kind="glow", n_in=1, nh=32, L=2, K=2, sp=(64, 64), gb=8, precision="fp16x3",
                 label="NetworkGlow(1,32,L=2,K=2) on 64x64x1"),
    # BASELINE configs[2] (amortized_glow_mnist_inpainting.jl:82-89 at the SURVEY 8d size)
    "cfg3": dict(kind="cglow", n_in=1, n_cond=1, nh=32, L=2, K=10, sp=(64, 64), gb=128, precision="fp16x3",
                 label="NetworkConditionalGlow(1,1,32,L=2,K=10;split_scales=true) on 64x64x1 + 64x64x1 condition"),
    # BASELINE configs[3] (BASELINE does not fix L, K, n_hidden; the reference's default block k1 = k2 = 3)
    "cfg4": dict(kind="hint", n_in=2, nh=128, L=2, K=4, sp=(128, 128), gb=32, precision="bf16x3", k2=3,
                 label="NetworkMultiScaleHINT(2,128,L=2,K=4;split_scales=true,k1=3,k2=3) on 128x128x2"),
    # BASELINE configs[4] (L = K = 2, n_hidden = 32 unless told otherwise)
    "cfg5": dict(kind="glow", n_in=1, nh=32, L=2, K=2, sp=(64, 64, 64), gb=8, precision="fp16x3",
                 label="NetworkGlow3D(1,32,L=2,K=2) on 64x64x64x1"),
}


def _round_half_even(c):
    return int(round(c / 2))


def rb_flops_per_sample(cfg):
    """F of SURVEY 8(d): FLOPs of one ResidualBlock.forward summed over the L*K flow steps (None for HINT, whose
    recursion visits a block several times)."""
    if cfg["kind"] == "hint":
        return None
    nd = len(cfg["sp"])
    f, taps = 2 ** nd, 3 ** nd
    c, cc, px, F = cfg["n_in"], cfg.get("n_cond", 0), math.prod(cfg["sp"]), 0
    for i in range(cfg["L"]):
        c *= f
        cc *= f
        px //= f
        k = _round_half_even(c)
        F += cfg["K"] * 2 * px * (taps * (c - k + cc) * cfg["nh"] + cfg["nh"] ** 2 + taps * cfg["nh"] * 2 * k)
        if i < cfg["L"] - 1:
            c //= 2
    return F


def elementwise_bytes_per_sample(cfg):
    """6 * sum(A) of SURVEY 8(d)."""
    nd = len(cfg["sp"])
    f = 2 ** nd
    c, px, A = cfg["n_in"], math.prod(cfg["sp"]), 0
    for i in range(cfg["L"]):
        c *= f
        px //= f
        A += cfg["K"] * c * px * 4
        if i < cfg["L"] - 1 and cfg["kind"] != "hint":
            c //= 2
    return 6 * A


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"], "tensor": p["bf16_tflops_sustained"],
                "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback"}


# ---------------------------------------------------------------------------------------------- reference arm
def oracle_step_fn(cfg, batch):
    """The reference algorithm (oracle port) for one training step of `cfg` on the host cores."""
    import torch
    if cfg["kind"] == "hint":
        from oracle import hint_oracle as H
        net = H.NetworkMultiScaleHINT(cfg["n_in"], cfg["nh"], cfg["L"], cfg["K"], split_scales=True,
                                      k2=cfg.get("k2", 3), p2=(cfg.get("k2", 3) - 1) // 2, seed=0)
        X = torch.rand(batch, cfg["n_in"], *cfg["sp"])
        return lambda: H.hint_train_step(net, X)
    from oracle import glow_oracle as O
    if cfg["kind"] == "cglow":
        net = O.NetworkConditionalGlow(cfg["n_in"], cfg["n_cond"], cfg["nh"], cfg["L"], cfg["K"], split_scales=True,
                                       seed=0, faithful=False)
        X, C = torch.rand(batch, cfg["n_in"], *cfg["sp"]), torch.rand(batch, cfg["n_cond"], *cfg["sp"])

        def step():
            ZX, ZC, ld = net.forward(X, C)
            for p in net.get_params():
                p.grad = None
            net.backward(ZX / batch, ZX, ZC)
        return step
    net = O.NetworkGlow(cfg["n_in"], cfg["nh"], cfg["L"], cfg["K"], split_scales=True, ndims=len(cfg["sp"]), seed=0,
                        faithful=False)
    X = torch.rand(batch, cfg["n_in"], *cfg["sp"])
    return lambda: O.glow_train_step(net, X)


ORACLE_NOTE = ("torch-CPU oracle restatement of the Julia reference (no Julia in the image); lenient stand-in: the "
               "Householder gradient uses the closed-form Gram contraction instead of the reference's per-sample "
               "mat_tens_i loop (faithful=False) and the convolutions go through torch/oneDNN, both faster than "
               "the Julia CPU path they stand for")


def time_oracle(cfg, batch, steps, warmup, threads):
    import torch
    torch.set_num_threads(threads)
    step = oracle_step_fn(cfg, batch)
    with torch.no_grad():
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def cpu_sample_batch(cfg_name, cfg):
    return 2 if cfg_name == "cfg2" else min(cfg["gb"], 8)


def run_reference(args, cfg, cfg_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_b = cpu_sample_batch(cfg_name, cfg)
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    if cfg_name == "cfg2":
        steps = min(steps, 3)  # bounded: each step is 2 samples of the full 48-step network on the CPU
    v, sec = time_oracle(cfg, sample_b, steps, warmup, threads)
    sample = f"{steps} step(s) of batch {sample_b} of the full {cfg_name} network ({cfg['label']})"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, cfg_name, sample_b, "fp32"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": ORACLE_NOTE},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(cfg, name, global_batch, precision):
    big = name == "cfg2"
    return {"workload": f"{name}: {cfg['label']}", "global_batch": global_batch, "precision": precision,
            "l2": "activations per step >> 126 MB L2 (no flush needed)" if big else
                  "a 192 MB buffer is written between timed steps (L2 flush)"}


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on fd 1 at
    communicator creation), so fd 1 is pointed at stderr for the whole run and the line goes to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# ---------------------------------------------------------------------------------------------- our arm
class Workload:
    """Network + synthetic inputs of one BASELINE configuration behind one step interface."""

    def __init__(self, cfg, precision, B, rank, dev):
        import torch
        import inb200
        self.cfg, self.B, self.kind = cfg, B, cfg["kind"]
        gen = torch.Generator().manual_seed(1234 + rank)
        if self.kind == "hint":
            k2 = cfg.get("k2", 3)
            self.G = inb200.NetworkMultiScaleHINT(cfg["n_in"], cfg["nh"], cfg["L"], cfg["K"], split_scales=True, k2=k2,
                                                  p2=(k2 - 1) // 2, precision=precision, seed=0, device=dev)
        elif self.kind == "cglow":
            self.G = inb200.NetworkConditionalGlow(cfg["n_in"], cfg["n_cond"], cfg["nh"], cfg["L"], cfg["K"],
                                                   split_scales=True, precision=precision, seed=0, device=dev)
        else:
            self.G = inb200.NetworkGlow(cfg["n_in"], cfg["nh"], cfg["L"], cfg["K"], split_scales=True,
                                        ndims=len(cfg["sp"]), precision=precision, seed=0, device=dev)
        self.host = [torch.rand(B, cfg["n_in"], *cfg["sp"], generator=gen).pin_memory()]
        if self.kind == "cglow":
            self.host.append(torch.rand(B, cfg["n_cond"], *cfg["sp"], generator=gen).pin_memory())
        self.dev_in = [t.to(dev) for t in self.host]
        self.h2d_bytes = sum(t.numel() * 4 for t in self.host)

    def step(self, inputs=None):
        """forward -> NLL gradient -> backward; returns (nll, logdet) device scalars"""
        import inb200
        x = self.dev_in if inputs is None else inputs
        if self.kind == "cglow":
            ZX, ZC, ld = self.G.forward(x[0], x[1])
            nll, dZ = inb200.nll_grad(ZX, self.B)
            self.G.backward(dZ, ZX, ZC)
        else:
            Z, ld = self.G.forward(x[0])
            nll, dZ = inb200.nll_grad(Z, self.B)
            self.G.backward(dZ, Z)
        inb200.clear_grad(self.G)
        return nll, ld


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("INB_PRECISION", "auto"))
    ap.add_argument("--config", default="cfg2", choices=list(CONFIGS))
    ap.add_argument("--global-batch", type=int, default=0)
    ap.add_argument("--dp", default="abi", choices=["abi", "explicit", "torch"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg, args.config)

    import torch
    import inb200
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert inb200.lib.load().inb_device_ok() == 1

    precision = cfg["precision"] if args.precision == "auto" else args.precision
    gb = args.global_batch or cfg["gb"]
    assert gb % world == 0, "global batch must divide over the ranks"
    B = gb // world
    steps, warmup = args.steps, max(args.warmup, 3)

    W = Workload(cfg, precision, B, rank, dev)
    G = W.G
    comm = None
    dp_mode = args.dp if (world > 1 and W.kind != "hint") else ("torch" if world > 1 else "none")
    if world > 1 and dp_mode in ("abi", "explicit"):
        comm = inb200.dp.Communicator.from_dist()
    if dp_mode == "abi":
        # ActNorm from the statistics of the GLOBAL batch inside the first forward; gradients averaged per scale on the
        # communicator's stream inside backward
        inb200.dp.attach(G, comm)
        W.step()
    elif world > 1:
        # data-dependent ActNorm init on rank 0's shard, then identical parameters everywhere
        if rank == 0:
            W.step()
        inb200.dp.broadcast_params(G.flat_params, src=0)
        G._mark_initialized()
    small = args.config != "cfg2"
    flush = torch.empty(48 * 1024 * 1024, device=dev) if small else None  # 192 MB > 126 MB L2

    def step_device(inputs=None):
        if flush is not None:
            flush.zero_()
        out = W.step(inputs)
        if dp_mode == "explicit":
            inb200.dp.allreduce_grads_abi(G, comm)
        elif dp_mode == "torch":
            inb200.dp.allreduce_grads(G.flat_grads)
        return out

    # End-to-end loop (what a training script does): every step's batch comes from pinned host memory and the
    # step's loss goes back to the host, which waits for it before the next step.  The input copy of step k+1
    # is issued on a copy stream while step k computes (two device buffers), like any prefetching data loader;
    # all K copies and K loss reads happen inside the timed region.
    loss_host = torch.empty(2).pin_memory()
    Xbuf = [W.dev_in, [torch.empty_like(t) for t in W.dev_in]]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"k": 0, "pending": False}

    def prefetch(buf):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[buf])
            for d, h in zip(Xbuf[buf], W.host):
                d.copy_(h, non_blocking=True)
            copied[buf].record(copy_stream)

    def step_e2e(last=False):
        k = e2e_state["k"]
        buf = k & 1
        if not e2e_state["pending"]:
            prefetch(buf)
        cur = torch.cuda.current_stream()
        cur.wait_event(copied[buf])
        nll, ld = step_device(Xbuf[buf])
        consumed[buf].record(cur)
        if not last:
            prefetch(buf ^ 1)
        e2e_state["pending"] = not last
        loss_host[0:1].copy_(nll.reshape(1), non_blocking=True)
        loss_host[1:2].copy_(ld.reshape(1), non_blocking=True)
        cur.synchronize()
        e2e_state["k"] = k + 1
        return float(loss_host[0] - loss_host[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, mark_last=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            if mark_last:
                fn(last=(i == n - 1))
            else:
                fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(warmup):
        step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = inb200.lib.launch_count()
    ms = timed(step_device, steps)
    launches = inb200.lib.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    ms_flush = 0.0
    if flush is not None:  # the L2 flush between steps is not part of the step
        ms_flush = timed(lambda: flush.zero_(), steps)
        ms = max(ms - ms_flush, 1e-6)

    # communication: the raw time of the gradient average alone, and what of it is exposed in the step
    comm_ms = comm_exposed_ms = None
    if world > 1:
        if dp_mode == "abi":
            raw = timed(lambda: inb200.dp.allreduce_grads_abi(G, comm), steps)
            inb200.dp.attach(G, None)
            for _ in range(2):
                W.step()
            ms_local = timed(lambda: (flush.zero_() if flush is not None else None, W.step()), steps) - ms_flush
            inb200.dp.attach(G, comm)
            W.step()
            W.step()
        else:
            raw = timed((lambda: inb200.dp.allreduce_grads_abi(G, comm)) if dp_mode == "explicit" else
                        (lambda: inb200.dp.allreduce_grads(G.flat_grads)), steps)
            ms_local = timed(lambda: (flush.zero_() if flush is not None else None, W.step()), steps) - ms_flush
        comm_ms = raw / steps
        comm_exposed_ms = (ms - ms_local) / steps

    # per-family CUDA-event timing on the launching stream for the roofline object (a second pass of
    # the same steps; the value above is measured without the extra events)
    prof = None
    if rank == 0:
        L_ = inb200.lib.load()
        L_.inb_prof_reset()
        L_.inb_prof_enable(1)
    ms_prof = timed(step_device, steps)
    if rank == 0:
        L_.inb_prof_enable(0)
        prof = inb200.lib.prof_table()

    # host-side cost of enqueuing one step (no synchronisation inside): when it approaches the device time the
    # per-step synchronisation of the end-to-end loop exposes it
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_device()
    host_enqueue_ms = (time.perf_counter() - t0) * 1e3 / steps
    barrier()

    step_e2e(last=True)
    step_e2e(last=True)  # both input buffers have been through the graph cache once
    g0 = G.graph_stats() if hasattr(G, "graph_stats") else {}
    ms_e2e = timed(step_e2e, steps, mark_last=True) - ms_flush
    g1 = G.graph_stats() if hasattr(G, "graph_stats") else {}
    f_last = step_e2e(last=True)
    # host -> device bandwidth of this box for the step's input (explains e2e - value when the link is slow)
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(3):
        for d, h in zip(Xbuf[1], W.host):
            d.copy_(h, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * W.h2d_bytes / (h0.elapsed_time(h1) / 1e3) / 1e9

    if rank != 0:
        if comm is not None:
            comm.destroy()
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = gb * steps / (ms / 1e3)
    e2e = gb * steps / (ms_e2e / 1e3)
    F = rb_flops_per_sample(cfg)
    terms = 3 if precision in ("bf16x3", "fp16x3") else 1
    # dominant kernel family by device time
    roof = None
    if prof:
        top = max((r for r in prof if r["name"] != "grad_finish"), key=lambda r: r["ms"])
        per_launch_ms = top["ms"] / max(top["scopes"], 1)
        tensor_bound = top["name"] in ("conv_simt", "wgrad_simt", "conv_tc", "wgrad_tc")
        if tensor_bound:
            executed = top["flops"] / (top["ms"] / 1e3) / 1e12
            # algorithmic flops of the family on this rank: the fused chain runs three of the four RB passes
            # of memory-efficient training (forward, recompute, data gradient) = 3F per sample, the weight gradients F
            alg = None
            if F is not None and top["name"] in ("conv_tc", "conv_simt"):
                alg = 3 * F * B * steps
            elif F is not None and top["name"] in ("wgrad_tc", "wgrad_simt"):
                alg = F * B * steps
            on_tc = top["name"] in ("conv_tc", "wgrad_tc")
            achieved = (alg / (top["ms"] / 1e3) / 1e12) if alg else executed / (terms if on_tc else 1)
            roof = {"bound": "tensor", "kernel": top["name"], "achieved": achieved, "peak": pk["tensor"],
                    "unit": "TFLOP/s", "frac": achieved / pk["tensor"], "traffic": None,
                    "achieved_is": "ALGORITHMIC flops of the family (3F per sample for the fused chain, SURVEY 8d) / its "
                                   "device time, against the measured dense bf16 peak",
                    "executed_mma_tflops": executed if on_tc else None,
                    "frac_executed_mma": (executed / pk["tensor"]) if on_tc else None,
                    "executed_is": f"issued tensor-core flops: padded tiles x {terms} products per term ({precision}); "
                                   "= tensor-pipe occupancy, not useful work",
                    "mode_peak_tflops": pk["tensor"] / terms if on_tc else None,
                    "frac_of_mode_peak": (achieved / (pk["tensor"] / terms)) if on_tc else None,
                    "peak_source": f"{pk['src']} bf16 sustained (kernel timed inside a long step); kind::f16 with half "
                                   "operands runs at the same rate"}
        else:
            achieved = top["bytes"] / (top["ms"] / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": top["name"], "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": achieved / pk["hbm"], "traffic": None, "peak_source": pk["src"]}
        # dram bytes per launch of that kernel from the committed ncu --set full capture
        for fn in ("r02_traffic.json", "r01_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", fn)) as fh:
                    tr = json.load(fh)
                if top["name"] in tr:
                    roof["traffic"] = tr[top["name"]].get("traffic_bytes_per_launch_mean")
                    roof["traffic_source"] = tr["source"]
                    break
            except Exception:
                pass
        roof["off_critical_path"] = ["grad_finish"]  # side-lane kernels: their event pairs also time the waiting on the lane
        roof["launches"] = top["scopes"]
        roof["avg_launch_ms"] = per_launch_ms
        roof["share_of_step"] = top["ms"] / ms_prof
        roof["families"] = {r["name"]: {"ms_per_step": r["ms"] / steps, "launches_per_step": r["launches"] / steps,
                                        "tflops": (r["flops"] / (r["ms"] / 1e3) / 1e12) if r["ms"] and r["flops"] else None,
                                        "gbs": (r["bytes"] / (r["ms"] / 1e3) / 1e9) if r["ms"] and r["bytes"] else None,
                                        "frac_of_hbm_peak": (r["bytes"] / (r["ms"] / 1e3) / 1e9 / pk["hbm"])
                                        if r["ms"] and r["bytes"] and not r["flops"] else None}
                            for r in sorted(prof, key=lambda r: -r["ms"])}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "bf16x3(f32-equivalent)", "fp16x3": "fp16x3(f32-equivalent)",
                  "bf16": "bf16"}[precision],
        "data": "synthetic", "config": workload_config(cfg, args.config, gb, precision),
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": W.h2d_bytes * world,
                "d2h_bytes_per_step": 8 * world, "ms_per_step": ms_e2e / steps, "loss": f_last,
                "h2d_gbs_measured": h2d_gbs, "input_prefetch": "copy of step k+1 overlaps step k (copy stream)"},
        "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
        "graphs_in_e2e_region": {k: g1[k] - g0[k] for k in g1},
        "roofline": roof,
    }
    if F is not None:
        useful = 4 * F * value / 1e12  # whole job
        line["model_flops"] = {"rb_forward_gflop_per_sample": F / 1e9, "fwd_bwd_gflop_per_sample": 4 * F / 1e9,
                               "useful_tflops": useful, "useful_tflops_per_gpu": useful / world,
                               "frac_of_tensor_peak": useful / world / pk["tensor"],
                               "elementwise_mb_per_sample": elementwise_bytes_per_sample(cfg) / 1e6}
    if world > 1:
        line["dp"] = {"mode": dp_mode, "comm_ms_per_step": comm_ms, "comm_exposed_ms_per_step": comm_exposed_ms,
                      "grad_bytes": int(G.flat_grads.numel() * 4),
                      "what": "comm_ms = the gradient average alone (one explicit all-reduce of the flat buffer, CUDA "
                              "events, max over ranks); exposed = step with averaging - step without",
                      "actnorm_init": "global batch (per-layer all-reduce of sums)" if dp_mode == "abi"
                                      else "rank 0's shard, broadcast"}
        if comm is not None:
            line["dp"]["comm"] = comm.info()
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sb = cpu_sample_batch(args.config, cfg)
        nst = 4 if args.config == "cfg2" else 10
        v, sec = time_oracle(cfg, sb, nst, 0 if args.config == "cfg2" else 1, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{nst} steps of batch {sb} of the full {args.config} network on the host "
                                          f"cores ({threads} threads, {sec * nst:.1f} s)",
                                "note": ORACLE_NOTE}
    emit(line)
    if comm is not None:
        comm.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
