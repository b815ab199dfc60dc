#!/usr/bin/env python
"""Where the end-to-end time of ccg_fsa_cmp_thread_out goes at bench scale: plain H2D rate of the
pinned rows (flat and 2-D slab copies), then the call itself with and without K-slab streaming."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ccphylo_b200 import api, synth  # noqa: E402
import synth_torch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
length = 5_000_000
dev = torch.device("cuda", 0)
W = api.words(length)
seqs_t, masks_t = synth_torch.make_packed_torch(n, length, seed=2, device=dev)
L = api.load()
hs_ptr = L.ccg_host_alloc(n * W * 8)
hm_ptr = L.ccg_host_alloc(n * W * 4)
nc = api.cells(n)
hD_ptr = L.ccg_host_alloc(nc * 8)
hN_ptr = L.ccg_host_alloc(nc * 8)
hs = torch.from_numpy(np.ctypeslib.as_array(C.cast(hs_ptr, C.POINTER(C.c_int64)), shape=(n, W)))
hm = torch.from_numpy(np.ctypeslib.as_array(C.cast(hm_ptr, C.POINTER(C.c_int32)), shape=(n, W)))
hs.copy_(seqs_t)
hm.copy_(masks_t)
torch.cuda.synchronize()
# plain H2D rates
for name, fn in (("flat", lambda: seqs_t.copy_(hs, non_blocking=True)),
                 ("2-D slab (1/8 of every row)", lambda: seqs_t[:, :W // 8].copy_(hs[:, :W // 8], non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    nbytes = n * W * 8 if name == "flat" else n * (W // 8) * 8
    print(f"H2D {name}: {nbytes / dt / 1e9:.1f} GB/s", file=sys.stderr)
del seqs_t, masks_t
torch.cuda.empty_cache()
sp = (C.c_void_p * n)(*[hs_ptr + k * W * 8 for k in range(n)])
mp = (C.c_void_p * n)(*[hm_ptr + k * W * 4 for k in range(n)])
include = np.ones(n, dtype=np.uint8)
for mode in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("stream", "stream4", "stream16")):
    if mode == "nostream":
        os.environ["CCG_STREAM_MIN_CHUNKS"] = "0"
    if mode.startswith("stream") and mode != "stream":
        os.environ["CCG_FEED_SLABS"] = mode[6:]
    ctx = api.Context(0)
    dn, gi = C.c_int(0), C.c_uint(0)
    for k in range(3):
        t = time.perf_counter()
        rc = L.ccg_fsa_cmp_thread_out(ctx._h, 1, hD_ptr, hN_ptr, 8, 1.0, n, length, sp, include.ctypes.data, mp, 0, 1, 0.5,
                                      0, C.byref(dn), C.byref(gi))
        dt = time.perf_counter() - t
        assert rc == 0, L.ccg_last_error(ctx._h)
        print(f"{mode} call {k}: {dt * 1e3:.1f} ms  compare {ctx.last_compare_ms():.1f} ms  [{ctx.last_kernel}]", file=sys.stderr)
    ctx.close()
