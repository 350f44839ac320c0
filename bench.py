#!/usr/bin/env python
"""bench.py - Glow fwd+bwd training throughput (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16x3|bf16]

One "step" = forward -> Gaussian NLL gradient (dZ = Z/B) -> memory-efficient backward (+ gradient
all-reduce when N > 1); the optimizer is the caller's (Flux) and is excluded, as in SURVEY 8(d).
Workload = BASELINE configs[1]: NetworkGlow(3, 256, L=3, K=16; split_scales) on 256x256x3, GLOBAL batch
64 sharded over the N ranks (strong scaling).  Prints ONE JSON line on rank 0.

`--impl reference` times the reference's algorithm on the host cores: the Julia reference cannot run in
this image (no Julia), so it is the torch-CPU oracle restatement (oracle/glow_oracle.py, kind "port").
"""
import argparse
import json
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
    # name: (n_in, n_hidden, L, K, (ny, nx), global batch)
    "cfg2": (3, 256, 3, 16, (256, 256), 64),   # BASELINE configs[1] (memory_usage_invertiblenetworks.jl:54-68)
    "cfg1": (1, 32, 2, 2, (64, 64), 8),        # BASELINE configs[0] (examples/networks/network_glow.jl)
}


def rb_flops_per_sample(n_in, nh, L, K, sp):
    """F of SURVEY 8(d): FLOPs of one ResidualBlock.forward summed over the L*K flow steps."""
    c, px, F = n_in, sp[0] * sp[1], 0
    for i in range(L):
        c *= 4
        px //= 4
        k = c // 2
        F += K * 2 * px * (9 * (c - k) * nh + nh * nh + 9 * nh * 2 * k)
        if i < L - 1:
            c //= 2
    return F


def elementwise_bytes_per_sample(n_in, L, K, sp):
    """6 * sum(A) of SURVEY 8(d)."""
    c, px, A = n_in, sp[0] * sp[1], 0
    for i in range(L):
        c *= 4
        px //= 4
        A += K * c * px * 4
        if i < L - 1:
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


def time_oracle(cfg, batch, steps, warmup, threads):
    """The reference algorithm (oracle port) on the host cores: samples/s over `steps` steps of `batch`."""
    import torch
    from oracle import glow_oracle as O
    n_in, nh, L, K, sp, _ = cfg
    torch.set_num_threads(threads)
    G = O.NetworkGlow(n_in, nh, L, K, split_scales=True, seed=0, faithful=False)
    X = torch.rand(batch, n_in, *sp)
    with torch.no_grad():
        for _ in range(warmup):
            O.glow_train_step(G, X)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.glow_train_step(G, X)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference(args, cfg, cfg_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_b = 2 if cfg_name == "cfg2" else cfg[5]
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    if cfg_name == "cfg2":
        steps = min(steps, 3)  # bounded: each step is 2 samples of the full 48-step network on the CPU
    v, sec = time_oracle(cfg, sample_b, steps, warmup, threads)
    sample = f"{steps} step(s) of batch {sample_b} of the full {cfg_name} network (L={cfg[2]}, K={cfg[3]}, n_hidden={cfg[1]})"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, cfg_name, sample_b, "fp32"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "torch-CPU oracle restatement of the Julia reference (no Julia in the image)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(cfg, name, global_batch, precision):
    n_in, nh, L, K, sp, _ = cfg
    return {"workload": f"{name}: NetworkGlow({n_in},{nh},L={L},K={K};split_scales=true) on {sp[1]}x{sp[0]}x{n_in}",
            "global_batch": global_batch, "precision": precision,
            "l2": "activations per step >> 126 MB L2 (no flush needed)" if name == "cfg2" else "L2-resident (latency bound)"}


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


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("INB_PRECISION", "bf16x3"))
    ap.add_argument("--config", default="cfg2", choices=list(CONFIGS))
    ap.add_argument("--global-batch", type=int, default=0)
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

    n_in, nh, L, K, sp, gb = cfg
    gb = args.global_batch or gb
    assert gb % world == 0, "global batch must divide over the ranks"
    B = gb // world
    steps, warmup = args.steps, max(args.warmup, 3)

    G = inb200.NetworkGlow(n_in, nh, L, K, split_scales=True, precision=args.precision, seed=0, device=dev)
    gen = torch.Generator().manual_seed(1234 + rank)
    X_host = torch.rand(B, n_in, *sp, generator=gen).pin_memory()
    X = X_host.to(dev)
    # data-dependent ActNorm init on rank 0's shard, then identical parameters everywhere
    if rank == 0:
        G.forward(X)
    if world > 1:
        inb200.dp.broadcast_params(G.flat_params, src=0)
        G._mark_initialized()

    def step_device(Xin=None):
        Z, ld = G.forward(X if Xin is None else Xin)
        nll, dZ = inb200.nll_grad(Z, B)
        G.backward(dZ, Z)
        inb200.clear_grad(G)
        if world > 1:
            inb200.dp.allreduce_grads(G.flat_grads)
        return nll, ld

    # End-to-end loop (what a training script does): every step's batch comes from pinned host memory and the
    # step's loss goes back to the host, which waits for it before the next step.  The input copy of step k+1
    # is issued on a copy stream while step k computes (two device buffers), like any prefetching data loader;
    # all K copies and K loss reads happen inside the timed region.
    loss_host = torch.empty(2).pin_memory()
    Xbuf = [X, torch.empty_like(X)]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"k": 0, "pending": False}

    def prefetch(buf):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[buf])
            Xbuf[buf].copy_(X_host, non_blocking=True)
            copied[buf].record(copy_stream)

    def step_e2e(last=False):
        k = e2e_state["k"]
        buf = k & 1
        if not e2e_state["pending"]:
            prefetch(buf)
        cur = torch.cuda.current_stream()
        cur.wait_event(copied[buf])
        Z, ld = G.forward(Xbuf[buf])
        consumed[buf].record(cur)  # backward recomputes X from Z: the input buffer is free after forward
        if not last:
            prefetch(buf ^ 1)
        e2e_state["pending"] = not last
        nll, dZ = inb200.nll_grad(Z, B)
        G.backward(dZ, Z)
        inb200.clear_grad(G)
        if world > 1:
            inb200.dp.allreduce_grads(G.flat_grads)
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
    g0 = G.graph_stats()
    ms_e2e = timed(step_e2e, steps, mark_last=True)
    g1 = G.graph_stats()
    f_last = step_e2e(last=True)
    # host -> device bandwidth of this box for the step's input (explains e2e - value when the link is slow)
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(3):
        Xbuf[1].copy_(X_host, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * X_host.numel() * 4 / (h0.elapsed_time(h1) / 1e3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = gb * steps / (ms / 1e3)
    e2e = gb * steps / (ms_e2e / 1e3)
    F = rb_flops_per_sample(n_in, nh, L, K, sp)
    # dominant kernel family by device time
    roof = None
    if prof:
        top = max((r for r in prof if r["name"] != "grad_finish"), key=lambda r: r["ms"])
        per_launch_ms = top["ms"] / max(top["scopes"], 1)
        tensor_bound = top["name"] in ("conv_simt", "wgrad_simt", "conv_tc", "wgrad_tc")
        if tensor_bound:
            achieved = top["flops"] / (top["ms"] / 1e3) / 1e12
            roof = {"bound": "tensor", "kernel": top["name"], "achieved": achieved, "peak": pk["tensor"],
                    "unit": "TFLOP/s", "frac": achieved / pk["tensor"], "traffic": None,
                    "peak_source": f"{pk['src']} bf16 sustained (kernel timed inside a long step)"}
        else:
            achieved = top["bytes"] / (top["ms"] / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": top["name"], "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": achieved / pk["hbm"], "traffic": None, "peak_source": pk["src"]}
        # dram bytes per launch of that kernel from the committed ncu --set full capture (profiles/r01_traffic.json)
        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
                tr = json.load(fh)
            roof["traffic"] = tr[top["name"]].get("traffic_bytes_per_launch_mean")
            roof["traffic_source"] = tr["source"]
        except Exception:
            pass
        roof["off_critical_path"] = ["grad_finish"]  # side-lane kernels: their event pairs also time the waiting on the lane
        roof["launches"] = top["scopes"]
        roof["avg_launch_ms"] = per_launch_ms
        roof["share_of_step"] = top["ms"] / ms_prof
        roof["families"] = {r["name"]: {"ms_per_step": r["ms"] / steps, "launches_per_step": r["launches"] / steps,
                                        "tflops": (r["flops"] / (r["ms"] / 1e3) / 1e12) if r["ms"] and r["flops"] else None,
                                        "gbs": (r["bytes"] / (r["ms"] / 1e3) / 1e9) if r["ms"] and r["bytes"] else None}
                            for r in sorted(prof, key=lambda r: -r["ms"])}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "bf16x3(f32-equivalent)", "bf16": "bf16"}[args.precision],
        "data": "synthetic", "config": workload_config(cfg, args.config, gb, args.precision),
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": X_host.numel() * 4 * world,
                "d2h_bytes_per_step": 8 * world, "ms_per_step": ms_e2e / steps, "loss": f_last,
                "h2d_gbs_measured": h2d_gbs, "input_prefetch": "copy of step k+1 overlaps step k (copy stream)"},
        "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
        "graphs_in_e2e_region": {k: g1[k] - g0[k] for k in g1},
        "roofline": roof,
        "model_flops": {"rb_forward_gflop_per_sample": F / 1e9, "fwd_bwd_gflop_per_sample": 4 * F / 1e9,
                        "useful_tflops": 4 * F * value / 1e12, "frac_of_tensor_peak": 4 * F * value / 1e12 / pk["tensor"],
                        "elementwise_mb_per_sample": elementwise_bytes_per_sample(n_in, L, K, sp) / 1e6},
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sb = 2 if args.config == "cfg2" else gb
        nst = 4 if args.config == "cfg2" else 20
        v, sec = time_oracle(cfg, sb, nst, 0 if args.config == "cfg2" else 1, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{nst} steps of batch {sb} of the full {args.config} network on the host "
                                          f"cores (torch-CPU oracle restating the Julia reference, {threads} threads, "
                                          f"{sec * nst:.1f} s)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
