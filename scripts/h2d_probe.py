#!/usr/bin/env python
"""Host -> device copy rates of this box: flat cudaMemcpyAsync against the strided cudaMemcpy2DAsync slab copies the
K-slab streaming of host rows issues (feed_slab in ccg_api.cu), one and two streams.  Explains e2e numbers."""
import ctypes as C
import sys
import time

import torch

rt = C.CDLL("libcudart.so.12")
n, W = 10000, 156250
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
host = torch.empty((n, W), dtype=torch.int64).pin_memory()
dev = torch.empty((n, W // 16 + 8), dtype=torch.int64, device="cuda")
flat = torch.empty(n * W // 16, dtype=torch.int64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()


def timed(fn, nbytes, label):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"{label}: {nbytes / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms)", file=sys.stderr)


wslab = W // 16                      # words of one of 16 slabs
nb = n * wslab * 8
timed(lambda: rt.cudaMemcpyAsync(flat.data_ptr(), host.data_ptr(), nb, 1, s1.cuda_stream), nb, "flat, one stream")
timed(lambda: rt.cudaMemcpy2DAsync(dev.data_ptr(), dev.stride(0) * 8, host.data_ptr(), W * 8, wslab * 8, n, 1, s1.cuda_stream), nb,
      f"2-D slab ({wslab * 8} B of every {W * 8} B row), one stream")


def two():
    h = n // 2
    rt.cudaMemcpy2DAsync(dev.data_ptr(), dev.stride(0) * 8, host.data_ptr(), W * 8, wslab * 8, h, 1, s1.cuda_stream)
    rt.cudaMemcpy2DAsync(dev[h:].data_ptr(), dev.stride(0) * 8, host[h:].data_ptr(), W * 8, wslab * 8, n - h, 1, s2.cuda_stream)


timed(two, nb, "2-D slab, two streams (half of the rows each)")


def batches():
    # what feed_slab does: batches of rows that fit 256 MB of staging, alternating streams
    per = max(32, (256 << 20) // (wslab * 12))
    k, q = 0, 0
    while k < n:
        m = min(per, n - k)
        rt.cudaMemcpy2DAsync(dev[k:].data_ptr(), dev.stride(0) * 8, host[k:].data_ptr(), W * 8, wslab * 8, m, 1, (s1, s2)[q].cuda_stream)
        k += m
        q ^= 1


timed(batches, nb, "2-D slab in 256 MB batches, alternating streams")
