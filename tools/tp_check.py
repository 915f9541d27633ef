"""Tensor-parallel decoder tail on N GPUs (one process per GPU): parity against the unsharded chain and timing.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py [--m 1]

o / down are row-sharded (K / N per rank, 128-channel boundaries: reference plugin.py:33-50), gate_up / qkv column-sharded;
the two all-reduces of the block happen INSIDE the chain launch (block sums written into every rank's peer buffer over
NVLink, added in rank order).  Checks: every rank's residual stream is bit-identical; outputs match the unsharded chain
within 1e-3 (another fp32 summation order); then times the TP chain against the same shards run as single linears +
NCCL all-reduce (what the reference does through vLLM).
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from paroquant_b200 import chain  # noqa: E402
from paroquant_b200.checkpoint import make_synthetic_activations, make_synthetic_layer  # noqa: E402
from paroquant_b200.linear import ParoLinearKernel  # noqa: E402
from paroquant_b200.parallel import shard_columns, shard_rows  # noqa: E402

H, KV, I = 4096, 1024, 14336
SH = {"o": (H, [H], "row"), "gate_up": (H, [I, I], "col"), "down": (I, [H], "row"), "qkv": (H, [H, KV, KV], "col")}


def err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", dest="m", type=int, default=1)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--small", action="store_true", help="hidden 1024 / inter 2048 (quick parity run)")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    M, dt = a.m, torch.bfloat16
    sh = SH if not a.small else {"o": (1024, [1024], "row"), "gate_up": (1024, [2048, 2048], "col"), "down": (2048, [1024], "row"),
                                 "qkv": (1024, [1024, 256, 256], "col")}
    hid, inter = sh["o"][0], sh["down"][0]

    # ---- parity on one layer: full layer everywhere (same seeds), sharded by rank
    full = {n: make_synthetic_layer(K, p, seed=40 + i, device=dev) for i, (n, (K, p, _)) in enumerate(sh.items())}
    shard = {n: (shard_rows(full[n], rank, world) if sh[n][2] == "row" else shard_columns(full[n], rank, world)) for n in sh}
    kf = {n: ParoLinearKernel.from_buffers(full[n], dt, check_pairs=False) for n in sh}
    ks = {n: ParoLinearKernel.from_buffers(shard[n], dt, check_pairs=False) for n in sh}
    attn = make_synthetic_activations(M, hid, seed=1, device=dev, dtype=dt)
    resid = make_synthetic_activations(M, hid, seed=2, device=dev, dtype=dt)
    g = torch.Generator(device="cpu").manual_seed(3)
    w1 = (1 + 0.1 * torch.randn(hid, generator=g)).to(dt).to(dev)
    w2 = (1 + 0.1 * torch.randn(hid, generator=g)).to(dt).to(dev)
    chf, bf = chain.decoder_tail(kf["o"], kf["gate_up"], kf["down"], kf["qkv"], attn_out=attn, residual=resid, post_attn_norm=w1, next_input_norm=w2)
    attn_sh = attn[:, rank * hid // world:(rank + 1) * hid // world].contiguous()
    chs, bs = chain.decoder_tail(ks["o"], ks["gate_up"], ks["down"], ks["qkv"], attn_out=attn_sh, residual=resid, post_attn_norm=w1,
                                 next_input_norm=w2, tensor_parallel=True)
    chf()
    for _ in range(3):      # several launches: the peer buffers alternate with the launch parity
        chs()
    torch.cuda.synchronize()
    dist.barrier()
    how = chs._tp[0][1].how if chs._tp else "-"
    e1, e2 = err(bs["residual_mid"], bf["residual_mid"]), err(bs["residual_out"], bf["residual_out"])
    cols = torch.cat([torch.arange(n0 + rank * n // world, n0 + (rank + 1) * n // world) for n0, n in ((0, inter), (inter, inter))]).to(dev)
    e3 = err(bs["mlp_act"], bf["mlp_act"][:, cols])
    qp = sh["qkv"][1]
    qcols, n0 = [], 0
    for n in qp:
        qcols.append(torch.arange(n0 + rank * n // world, n0 + (rank + 1) * n // world))
        n0 += n
    e4 = err(bs["qkv"], bf["qkv"][:, torch.cat(qcols).to(dev)])
    gathered = [torch.empty_like(bs["residual_out"]) for _ in range(world)]
    dist.all_gather(gathered, bs["residual_out"])
    same = all(torch.equal(t, gathered[0]) for t in gathered)
    ok = e1 < 1e-3 and e2 < 2e-3 and e3 < 2e-3 and e4 < 2e-3 and same
    print(f"[rank {rank}] peer memory: {how}; residual_mid {e1:.2e} residual_out {e2:.2e} mlp_act {e3:.2e} qkv {e4:.2e}; "
          f"ranks bit-identical: {same} -> {'OK' if ok else 'FAIL'}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if int(flag):
        dist.barrier()
        os._exit(1)
    if a.small:
        dist.barrier()
        os._exit(0)

    # ---- timing: `layers` distinct weight sets (sharded directly), TP chain vs single linears + NCCL
    del kf, full, chf
    torch.cuda.empty_cache()

    def shard_shape(n):
        K, p, kind = sh[n]
        return (K // world, p) if kind == "row" else (K, [v // world for v in p])

    layers = [{n: ParoLinearKernel.from_buffers(make_synthetic_layer(*shard_shape(n), seed=700 + 8 * li + i + 100 * rank, device=dev), dt,
                                                check_pairs=False, max_m=M) for i, n in enumerate(sh)} for li in range(a.layers)]
    chains = [chain.decoder_tail(l["o"], l["gate_up"], l["down"], l["qkv"], attn_out=attn_sh, residual=resid, post_attn_norm=w1, next_input_norm=w2,
                                 tensor_parallel=True)[0] for l in layers]
    xs = {n: make_synthetic_activations(M, shard_shape(n)[0], seed=5 + i, device=dev, dtype=dt) for i, n in enumerate(sh)}
    ys = {n: torch.empty(M, sum(shard_shape(n)[1]), dtype=dt, device=dev) for n in sh}

    def run_chain():
        for c in chains:
            c()

    def run_nccl():
        for l in layers:
            for n in sh:
                l[n].forward_into(xs[n], ys[n])
                if sh[n][2] == "row":
                    dist.all_reduce(ys[n])

    for name, fn in (("tp chain (fused sum)", run_chain), ("linears + NCCL", run_nccl)):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        for _ in range(3):
            gr.replay()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            gr.replay()
        e1_.record()
        torch.cuda.synchronize()
        us = torch.tensor([e0.elapsed_time(e1_) * 1e3 / (a.reps * a.layers)], device=dev)
        dist.all_reduce(us, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name:22s} TP={world} M={M}: {float(us):8.2f} us / layer -> {M * 1e6 / (32 * float(us)):7.0f} tok/s for 32 layers", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
