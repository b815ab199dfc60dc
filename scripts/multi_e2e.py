#!/usr/bin/env python
"""End-to-end time of ONE library call on the in-process multi-GPU context (ccg_init_multi): host rows of the
whole alignment in pinned memory -> ccg_fsa_cmp_thread_out -> host D / N.  This is what `ccphylo-b200 dist` gets.

    python scripts/multi_e2e.py --gpus 8 [--samples 10000 --length 5000000 --steps 3]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--samples", type=int, default=10000)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", type=int, default=96, help="samples of the oracle spot check")
    args = ap.parse_args()
    import torch
    from ccphylo_b200 import api, synth
    import synth_torch  # noqa: E402

    n, length = args.samples, args.length
    W = api.words(length)
    L = api.load()
    hs_ptr, hm_ptr = L.ccg_host_alloc(n * W * 8), L.ccg_host_alloc(n * W * 4)
    ncell = api.cells(n)
    hD_ptr, hN_ptr = L.ccg_host_alloc(ncell * 8), L.ccg_host_alloc(ncell * 8)
    assert hs_ptr and hm_ptr and hD_ptr and hN_ptr
    hs = np.ctypeslib.as_array(C.cast(hs_ptr, C.POINTER(C.c_uint64)), shape=(n, W))
    hm = np.ctypeslib.as_array(C.cast(hm_ptr, C.POINTER(C.c_uint32)), shape=(n, W))
    dev = torch.device("cuda", 0)
    t0 = time.time()
    blk = 1000
    for k in range(0, n, blk):                                   # generate on the GPU block by block, park on the host
        m = min(blk, n - k)
        s, mk = synth_torch.make_packed_torch(m, length, seed=2 + k, device=dev)
        torch.from_numpy(hs[k:k + m].view(np.int64)).copy_(s)
        torch.from_numpy(hm[k:k + m].view(np.int32)).copy_(mk)
    torch.cuda.synchronize()
    del s, mk
    torch.cuda.empty_cache()
    print(f"generated {n} x {length} in {time.time() - t0:.1f}s", file=sys.stderr)
    sp = (C.c_void_p * n)(*[hs_ptr + k * W * 8 for k in range(n)])
    mp = (C.c_void_p * n)(*[hm_ptr + k * W * 4 for k in range(n)])
    include = np.ones(n, np.uint8)
    dn, ginc = C.c_int(0), C.c_uint(0)
    ctx = api.Context(multi=args.gpus)
    times = []
    for it in range(args.steps + 2):
        t = time.perf_counter()
        rc = L.ccg_fsa_cmp_thread_out(ctx._h, 1, hD_ptr, hN_ptr, 8, 1.0, n, length, sp, include.ctypes.data, mp, 0, 1, 0.5, 0,
                                      C.byref(dn), C.byref(ginc))
        dt = time.perf_counter() - t
        if rc:
            raise SystemExit(L.ccg_last_error(ctx._h).decode())
        if it >= 2:
            times.append(dt)
    hD = np.ctypeslib.as_array(C.cast(hD_ptr, C.POINTER(C.c_double)), shape=(ncell,))
    hN = np.ctypeslib.as_array(C.cast(hN_ptr, C.POINTER(C.c_double)), shape=(ncell,))
    import bench
    import oracle
    ids = bench.parity_sample_ids(n, args.check)
    cell_idx, _ = bench.packed_index(ids)
    Do, No, _ = oracle.fsa_cmp_pair(hs[ids], hm[ids], np.ones(len(ids), np.uint8), length)
    ok = bool(np.array_equal(hD[cell_idx], Do) and np.array_equal(hN[cell_idx], No))
    gp, active = ctx.multi_gpus()
    ms = float(np.mean(times)) * 1e3
    print(json.dumps({"gpus": gp, "active": active, "samples": n, "length": length, "e2e_ms": ms, "wall": "host clock around the call",
                      "base_cmp_per_s": ncell * length / (ms * 1e-3), "kernel": ctx.last_kernel,
                      "parity_vs_oracle": ok, "cells_checked": int(len(cell_idx))}))
    ctx.close()
    if not ok:
        raise SystemExit("multi_e2e: result differs from the oracle")


if __name__ == "__main__":
    main()
