#!/usr/bin/env python
"""One hot-path step at bench scale, for ncu captures and knob experiments.

    python scripts/one_step.py [--samples 10000] [--length 5000000] [--steps 2] [--kernel auto]
Prints step / compare / GEMM-phase milliseconds per step on stderr.  Never a bench value."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ccphylo_b200 import api, synth  # noqa: E402
import synth_torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=10000)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--scratch-gb", type=float, default=0.0)
    ap.add_argument("--proxi", type=int, default=0, help="-P: run k_pairdist_proxi instead of the plain compare")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    seqs, masks = synth_torch.make_packed_torch(a.samples, a.length, seed=2, device=dev)
    ctx = api.Context(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_kernel({"auto": api.KERNEL_AUTO, "popc": api.KERNEL_POPC, "umma": api.KERNEL_UMMA,
                    "fused": api.KERNEL_FUSED}[a.kernel])
    if a.scratch_gb:
        ctx.set_scratch_limit(int(a.scratch_gb * 2 ** 30))
    ctx.set_proximity(a.proxi)
    ctx.set_problem(a.samples, a.length, pair=True)
    nc = api.cells(a.samples)
    d_D = torch.zeros(nc, dtype=torch.float64, device=dev)
    d_N = torch.zeros(nc, dtype=torch.float64, device=dev)
    for k in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.put_samples_packed_dev(seqs.data_ptr(), masks.data_ptr(), a.samples, seqs.stride(0))
        ctx.run_pair_dev(d_D.data_ptr(), d_N.data_ptr(), norm=0, min_length=1, min_cov=0.5, elem_size=8)
        e1.record(stream)
        torch.cuda.synchronize()
        print(f"step {k}: {e0.elapsed_time(e1):.2f} ms, compare {ctx.last_compare_ms():.2f} ms, "
              f"gemm phase {ctx.last_phase_ms(1):.2f} ms  [{ctx.last_kernel}] "
              f"{api.cells(a.samples) * a.length / (ctx.last_compare_ms() * 1e-3):.3e} base-cmp/s (compare)  "
              f"KSLICES={os.environ.get('CCG_KSLICES')} SERIAL={os.environ.get('CCG_EXPAND_SERIAL')}",
              file=sys.stderr, flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
