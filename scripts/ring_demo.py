#!/usr/bin/env python
"""The NCCL sample-shard ring on N GPUs (launch with torchrun, one rank per GPU).

    torchrun --nproc-per-node N scripts/ring_demo.py --check          parity of every block against the oracle (small)
    torchrun --nproc-per-node N scripts/ring_demo.py --shard 12544 --length 2900000 --steps 1
                                                                      BASELINE configs[3] at N = 8 (100,352 samples)
Rank 0 prints one JSON line: pairwise base comparisons / s of the whole job (max over ranks, CUDA events).
Never a bench.py value; kept under profiles/ as the evidence for the ring row of SURVEY.md section 8e."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ring  # noqa: E402
from ccphylo_b200 import api, synth  # noqa: E402
import synth_torch  # noqa: E402

_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shard", type=int, default=256)
    ap.add_argument("--length", type=int, default=128 * 40 + 7)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    S, L = a.shard, a.length
    assert S % 256 == 0
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    # every rank generates ITS shard only (seeded by the rank): no rank ever holds the whole set
    seqs, masks = synth_torch.make_packed_torch(S, L, seed=1000 + rank, device=dev, snp=0.01 if a.check else synth.SNP_RATE)
    st = ring.RankState(rank, world, S, L, seqs, masks, dev)
    st.ctx.set_stream(stream.cuda_stream)
    pinned = {}
    cells = [0]
    checks = []

    def on_block(g, hi, lo, row0, D, N, dn):
        cells[0] += D.numel()
        key = (hi == lo)
        if key not in pinned or pinned[key][0].numel() < D.numel():
            pinned[key] = (torch.empty(D.numel(), dtype=torch.float64).pin_memory(),
                           torch.empty(D.numel(), dtype=torch.float64).pin_memory())
        pinned[key][0][:D.numel()].copy_(D.reshape(-1), non_blocking=True)
        pinned[key][1][:N.numel()].copy_(N.reshape(-1), non_blocking=True)
        if a.check:
            torch.cuda.synchronize()
            checks.append((hi, lo, row0, dn, D.cpu().numpy().copy(), N.cpu().numpy().copy()))

    def one_pass():
        cells[0] = 0
        ring.run_nccl(st, on_block, norm=0, min_length=1, min_cov=0.5)

    for _ in range(a.warmup):
        checks.clear()
        one_pass()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        checks.clear()
        one_pass()
    e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / max(a.steps, 1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tot = torch.tensor([cells[0]], device=dev, dtype=torch.int64)
    dist.all_reduce(tot)
    ok = True
    if a.check:
        import oracle
        # every rank needs all shards for the oracle: gather the packed rows (small sizes only)
        gs = [torch.empty_like(seqs) for _ in range(world)]
        gm = [torch.empty_like(masks) for _ in range(world)]
        dist.all_gather(gs, seqs)
        dist.all_gather(gm, masks)
        hs = [x.cpu().numpy().view(np.uint64) for x in gs]
        hm = [x.cpu().numpy().view(np.uint32) for x in gm]
        for hi, lo, row0, dn, D, N in checks:
            if hi == lo:
                Do, No, _ = oracle.fsa_cmp_pair(hs[hi], hm[hi], np.ones(S, np.uint8), L)
                ok &= bool(np.array_equal(D, Do) and np.array_equal(N, No))
            else:
                both_s = np.concatenate([hs[lo], hs[hi]])
                both_m = np.concatenate([hm[lo], hm[hi]])
                Do, No, _ = oracle.fsa_cmp_pair(both_s, both_m, np.ones(2 * S, np.uint8), L)
                for r in range(D.shape[0]):
                    rr = S + row0 + r
                    o = rr * (rr - 1) // 2
                    ok &= bool(np.array_equal(D[r], Do[o:o + S]) and np.array_equal(N[r], No[o:o + S]))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    if rank == 0:
        n = world * S
        pairs = n * (n - 1) // 2
        line = {"what": "sample-shard ring over NCCL (ccphylo_b200/ring.py)", "n_gpus": world, "samples": n, "shard": S,
                "length": L, "cells_computed": int(tot.item()), "cells_strict_lower_triangle": pairs,
                "ms_per_pass": float(ms.item()), "base_cmp_per_s": pairs * L / (float(ms.item()) * 1e-3),
                "parity_vs_oracle": ok if a.check else None, "kernel": st.ctx.last_kernel}
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    st.ctx.close()
    dist.destroy_process_group()
    if a.check and not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
