#!/usr/bin/env python
"""The kernels behind -P / -y / -V at a moderate size, one call each, with their device times: for ncu captures
(`ncu -k regex:'k_pairdist_proxi|k_motif_mask|k_variants|k_sample_proxi'`) and quick timing.  Never a bench value.

    python scripts/f4_kernels.py [--samples 2048] [--length 1000000] [--proxi 10] [--list 128]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ccphylo_b200 import api, synth  # noqa: E402
import synth_torch  # noqa: E402

DAM_DCM = [[4, 17, 8, 2], [4, 1, 24, 2], [2, 18, 9, 4, 4], [2, 4, 9, 20, 4]]      # gAtc, gaTc, cCwgg, cgwGg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=2048)
    ap.add_argument("--length", type=int, default=1_000_000)
    ap.add_argument("--proxi", type=int, default=10)
    ap.add_argument("--list", type=int, default=128, help="samples whose pairs are listed by -V")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, L = a.samples, a.length
    seqs, masks = synth_torch.make_packed_torch(n, L, seed=2, device=dev)
    ctx = api.Context(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_problem(n, L, pair=True)
    ctx.put_samples_packed_dev(seqs.data_ptr(), masks.data_ptr(), n, seqs.stride(0))
    nc = api.cells(n)
    d_D = torch.zeros(nc, dtype=torch.float64, device=dev)
    d_N = torch.zeros(nc, dtype=torch.float64, device=dev)
    out = {}

    def timed(name, fn, units, unit):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record(stream)
        r = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = {"ms": ms, "wall_ms": (time.perf_counter() - t0) * 1e3, "rate": units / (ms * 1e-3), "unit": unit}
        print(f"{name}: {ms:.2f} ms  {units / (ms * 1e-3):.3e} {unit}", file=sys.stderr, flush=True)
        return r

    # -V first (needs the unmasked planes): the pairs among the first --list samples
    inc = np.zeros(n, np.uint8)
    inc[:a.list] = 1
    lists = timed("k_variants (count + write, incl. D2H of the lists)", lambda: ctx.list_variants(pair=True, include=inc),
                  api.cells(a.list) * L, "base-cmp/s")
    out["variants_listed"] = int(sum(len(v) for _, v in lists))
    # -P: per-sample builder (count only), then the pair kernel
    ctx.set_proximity(a.proxi)
    timed("k_sample_proxi", lambda: ctx.sample_proximity(0, n, apply=False), n * L, "bases/s")
    for k in range(2):
        timed(f"k_pairdist_proxi run {k}", lambda: ctx.run_pair_dev(d_D.data_ptr(), d_N.data_ptr(), norm=0, min_length=1,
                                                                     min_cov=0.5, elem_size=8), nc * L, "base-cmp/s")
    out["kernel"] = ctx.last_kernel
    ctx.set_proximity(0)
    # -y last (it changes the planes)
    ctx.set_motifs(DAM_DCM)
    timed("k_motif_mask + k_motif_remask", lambda: ctx.mask_motifs(0, n), n * L, "bases/s")
    ctx.set_motifs([])
    import json
    print(json.dumps({"what": "kernels behind -P / -y / -V (scripts/f4_kernels.py)", "samples": n, "length": L, "proxi": a.proxi,
                      **out}))
    ctx.close()


if __name__ == "__main__":
    main()
