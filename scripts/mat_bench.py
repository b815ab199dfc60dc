#!/usr/bin/env python
"""Throughput of the count-matrix (.mat) path, BASELINE configs[4] shape: n samples x L positions, all pairs,
-d cos and -d chi2.  Prints position pairs / s on the GPU and for the CPU oracle port (1 core, compute only --
the reference additionally re-inflates and re-parses sample j's file for every cell).  Not a bench.py value."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from ccphylo_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rng = np.random.default_rng(1)
ref = rng.integers(0, 4, size=L)
base = []
for k in range(8):                                           # 8 distinct samples, cycled over the n slots
    c = rng.poisson(0.3, size=(L, 6)).astype(np.uint16)
    call = np.where(rng.random(L) < 0.01, (ref + rng.integers(1, 4, size=L)) % 4, ref)
    c[np.arange(L), call] += rng.poisson(40, size=L).astype(np.uint16)
    base.append((c, c.astype(np.uint32).sum(axis=1).astype(np.uint32)))
ctx = api.Context(0)
t0 = time.perf_counter()
ctx.mat_set_problem(n, L)
for i in range(n):
    ctx.mat_put_sample(i, *base[i % 8])
ctx.sync()
t_up = time.perf_counter() - t0
out = {"samples": n, "positions": L, "upload_s": t_up, "position_pairs": n * (n - 1) // 2 * L}
for method in ("cos", "chi2", "l1"):
    ctx.mat_run(None, method=method)
    t0 = time.perf_counter()
    D, N, dn, rows = ctx.mat_run(None, method=method)
    dt = time.perf_counter() - t0
    out[method] = {"seconds": dt, "kernel_ms": ctx.last_compare_ms(), "position_pairs_per_s": out["position_pairs"] / dt,
                   "kernel": ctx.last_kernel}
ctx.close()
# CPU port on a small slice (compute only)
import oracle  # noqa: E402
ns, ls = 8, min(L, 200000)
counts = np.stack([base[i][0][:ls] for i in range(ns)])
totals = np.stack([base[i][1][:ls] for i in range(ns)])
t0 = time.perf_counter()
oracle.mat_matrix(counts, totals, np.full(ns, ls, np.int32), None, method="cos")
dt = time.perf_counter() - t0
out["cpu_port_cos_position_pairs_per_s_1core"] = ns * (ns - 1) // 2 * ls / dt
print(json.dumps(out))
