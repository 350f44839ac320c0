"""Staged probe of the data-parallel plane on two (or more) GPUs with per-rank progress logs in gpurun_out/ - every
stage is logged before and after, and faulthandler dumps the Python stack of a rank that sits in one call for more
than 45 s, so a hang names its call instead of eating the time limit.
    gpurun --gpus 2 -- 'timeout 150 python scripts/dp_probe.py 2'"""
import faulthandler
import os
import socket
import sys
import time

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port):
    sys.path.insert(0, ROOT)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", f"dp_probe_rank{rank}.log"), "w", buffering=1)
    faulthandler.enable(log)

    def say(msg):
        log.write(f"[{time.time() % 1000:8.2f}] {msg}\n")
        faulthandler.cancel_dump_traceback_later()
        faulthandler.dump_traceback_later(45, file=log, exit=True)

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    say("init gloo")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import inb200
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    torch.zeros(1, device=dev)
    say("communicator")
    comm = inb200.dp.Communicator.from_dist()
    say(f"communicator ok {comm.info()}")
    prec = os.environ.get("INB_PRECISION", "fp32")
    G = inb200.NetworkGlow(3, 128, 3, 2, split_scales=True, precision=prec, seed=1, device=dev)
    Xg = torch.rand(4 * world // 2, 3, 64, 64, generator=torch.Generator().manual_seed(7))
    lo, hi = inb200.dp.shard_bounds(Xg.shape[0], rank, world)
    X = Xg[lo:hi].to(dev)
    B = X.shape[0]
    stages = os.environ.get("DP_STAGES", "explicit,init,direct,graph").split(",")
    if "init" in stages:
        inb200.dp.attach(G, comm)
    say("forward (init)")
    Z, ld = G.forward(X)
    torch.cuda.synchronize()
    say("forward ok")
    if "explicit" in stages:
        inb200.dp.attach(G, None)
        nll, dZ = inb200.nll_grad(Z, B)
        G.backward(dZ, Z)
        torch.cuda.synchronize()
        say("local backward ok; explicit all-reduce")
        inb200.dp.allreduce_grads_abi(G, comm)
        torch.cuda.synchronize()
        say(f"explicit all-reduce ok {comm.info()}")
        inb200.clear_grad(G)
    inb200.dp.attach(G, comm)
    for it in range(4):
        say(f"attached backward {it}")
        nll, dZ = inb200.nll_grad(Z, B)
        G.backward(dZ, Z)
        torch.cuda.synchronize()
        say(f"attached backward {it} ok graph {G.graph_stats()} comm {comm.info()}")
        inb200.clear_grad(G)
    say("barrier")
    dist.barrier()
    say("destroy")
    comm.destroy()
    dist.destroy_process_group()
    say("done")
    faulthandler.cancel_dump_traceback_later()


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(worker, args=(n, port), nprocs=n, join=True)
    print("dp_probe finished")
