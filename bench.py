#!/usr/bin/env python
"""bench.py -- pairwise base comparisons/s of the `ccphylo dist` hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of synthetic KMA-consensus samples:
encode (reference packed words -> device bit planes) + all-vs-all compare + fused epilogue
(packed lower-triangular D and N as doubles).  Workload: the configuration BASELINE.json's
metric is quoted on -- 10,000 samples x 5 Mbp, pair mode, -n (configs[2]; it fits one B200:
19 GB of bit planes, the int8 operand panel is processed in K slabs).  The job is the same
at every N ("strong").  At N > 1 the ALIGNMENT axis is cut (K split): every GPU holds 1/N of
the bases of every sample, runs all macro tiles on its slice, and the int32 partial sums are
added by the GPU that owns a matrix row, through peer pointers over NVLink, inside its
epilogue kernel (ccphylo_b200/csrc/ccg_group.cu).  `--scaling weak` instead grows the sample
count with sqrt(N).  `--workload ring|mat` measures BASELINE configs[3] / configs[4]
(bench_workloads.py).

`value`  : inputs resident in HBM (reference packed format) when the timed region starts.
`e2e`    : the same metric through the reference-facing C-ABI call with HOST buffers
           (pinned rows in, host matrices out; H2D + D2H inside the timed region).
`--impl reference`: the UNMODIFIED reference's fsaCmpThreadOut (oracle/_ref, all host
           threads) on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "pairwise base comparisons/sec"
UNIT = "base-cmp/s"
LENGTH = 5_000_000
BASE_SAMPLES = 10000
OPS_PER_BASECMP = 8          # tensor ops (2 per MAC) of the K=4L tetrahedral+mask contraction


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Library chatter (e.g. NCCL's version banner) goes to fd 1; the contract is ONE JSON line on stdout.
# Keep the real stdout aside, point fd 1 at stderr, and write the result line to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def samples_for(world, base, scaling="strong"):
    if scaling == "strong" or world == 1:
        return base
    n = int(round(base * math.sqrt(world)))
    return max(64, (n // 64) * 64)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (profiling guide recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1] or [r for _, r in self.rows[-3:]]
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference's hot path on the host cores
# --------------------------------------------------------------------------------------
def cpu_reference_rate(seqs, masks, length, cores, budget_s, steps=1, warmup=0):
    """Runs the reference (oracle/_ref) -- or the oracle port if the reference .so is absent --
    on the first n_s samples, n_s sized so one step is about budget_s seconds.
    Returns (rate, kind, sample description, seconds per step, D, N)."""
    import oracle

    kind = "reference" if oracle.have_ref() else "port"
    n_all = seqs.shape[0]

    def run(ns, tnum):
        inc = np.ones(ns, dtype=np.uint8)
        t0 = time.perf_counter()
        if kind == "reference":
            D, N, dn = oracle.ref_fsa_cmp(seqs[:ns], masks[:ns], inc, length, pair=True, tnum=tnum, min_length=1,
                                          min_cov=0.5)
        else:
            D, N, dn = oracle.fsa_cmp_pair(seqs[:ns], masks[:ns], inc, length, min_length=1, min_cov=0.5)
        return time.perf_counter() - t0, D, N

    threads = cores if kind == "reference" else 1
    probe = min(n_all, max(8, int(math.sqrt(2 * threads * 4)) + 1))
    t_probe, _, _ = run(probe, threads)
    rate = probe * (probe - 1) / 2 * length / max(t_probe, 1e-9)
    pairs = budget_s * rate / length
    ns = int(min(n_all, max(probe, (1 + math.sqrt(1 + 8 * pairs)) / 2)))
    times = []
    D = N = None
    for k in range(warmup + steps):
        t, D, N = run(ns, threads)
        if k >= warmup:
            times.append(t)
    sec = float(np.mean(times))
    value = ns * (ns - 1) / 2 * length / sec
    what = (f"first {ns} of the workload's samples x {length} bp, pair mode, "
            f"{'fsaCmpThreadOut(cmpairFsaThrd) -t ' + str(threads) if kind == 'reference' else 'oracle port, 1 thread'}")
    return value, kind, what, sec, threads, ns, D, N


def make_host_workload(n, length, seed):
    """Synthetic workload on the host (numpy) for the reference arm when no GPU generated it."""
    import oracle
    from ccphylo_b200 import synth
    import synth_torch  # noqa: E402

    codes = synth.make_codes(n, length, seed=seed)
    return oracle.encode_samples(codes)[:2]


def reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = samples_for(world, args.samples, args.scaling)
    # the bounded sample never needs more than a few hundred samples: generate just those
    ns_cap = min(n, 64 + 8 * cores)
    seqs, masks = make_host_workload(ns_cap, args.length, seed=2)
    # the whole --steps / --warmup run must end within a few minutes: share about two of them between the steps
    budget = min(args.cpu_budget, 120.0 / max(1, args.steps + args.warmup))
    value, kind, what, sec, threads, ns, _, _ = cpu_reference_rate(seqs, masks, args.length, cores, budget,
                                                                   steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u64 words / u32 counters (scalar bit loops)", "data": "synthetic",
        "config": {"workload": f"{n} samples x {args.length} bp all-vs-all D+N (pair mode, -f 3 -n), "
                               f"bounded sample: {what}"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def parity_sample_ids(n, count=128, seed=12345):
    """Sample slots for the parity check, stratified over the 256-sample row blocks so that the pairs among them
    fall into EVERY macro tile of the lower triangle (>= 2 samples per block when count allows it)."""
    rng = np.random.default_rng(seed)
    blocks = (n + 255) // 256
    per = max(2, count // blocks)
    ids = []
    for b in range(blocks):
        lo, hi = b * 256, min(n, b * 256 + 256)
        k = min(per, hi - lo)
        ids.extend(rng.choice(np.arange(lo, hi), size=k, replace=False).tolist())
    return np.array(sorted(ids), dtype=np.int64)


def packed_index(ids):
    """Packed lower-triangular cell of every pair (a > b) among `ids`, in the order of a matrix over just those samples."""
    r, c = np.tril_indices(len(ids), -1)
    hi, lo = ids[r], ids[c]
    return hi * (hi - 1) // 2 + lo, hi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fasta", choices=["fasta", "ring", "mat"],
                    help="fasta: BASELINE configs[2], the metric's own configuration (default); ring: configs[3]; mat: configs[4]")
    ap.add_argument("--samples", type=int, default=None, help="samples (at N=1 when --scaling weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--length", type=int, default=None)
    ap.add_argument("--kernel", default="auto", choices=["auto", "popc", "umma", "fused"])
    ap.add_argument("--method", default="cos", help="--workload mat: the -d method")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work per reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--parity-samples", type=int, default=128)
    ap.add_argument("--copy-rows", action="store_true", help="device-resident arm: ccg_put_samples_packed_dev (copy into the "
                    "bit-plane store) instead of lending the rows")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}")

    if args.workload != "fasta":
        import bench_workloads
        return bench_workloads.main(args, rank, world, local_rank, emit, log, ClockSampler, measured_peaks)
    if args.samples is None:
        args.samples = BASE_SAMPLES
    if args.length is None:
        args.length = LENGTH
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        log("note: fewer than 3 warm-up steps requested; using 3")
        args.warmup = 3

    import torch
    import torch.distributed as dist

    from ccphylo_b200 import api, synth
    import synth_torch  # noqa: E402

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    # ---- the workload: n samples x `length` bp; at N > 1 the ALIGNMENT is cut (K split): rank r holds the bases
    # [b_r, b_r+1) of every sample and synthesises just that slice (the alignment is the concatenation) ----
    n, length = samples_for(world, args.samples, args.scaling), args.length
    slices = api.group_slices(length, world)
    b0, b1 = slices[rank], slices[rank + 1]
    len_r = b1 - b0
    if len_r < 256:
        raise SystemExit(f"bench.py: {length} bp cannot be cut into {world} slices of at least 256 bp")
    W = api.words(len_r)
    t_gen = time.time()
    seqs_t, masks_t = synth_torch.make_packed_torch(n, len_r, seed=2 + 1000 * rank, device=dev)
    torch.cuda.synchronize()
    log(f"[rank {rank}] generated {n} x {len_r} bp (bases {b0}..{b1} of {length}) in {time.time() - t_gen:.1f}s")

    ctx = api.Context(local_rank)
    # a real (non-default) stream: the library launches on it and the torch events below see the work
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    use_umma = args.kernel in ("auto", "umma", "fused")   # AUTO resolves to a tensor kernel at bench sizes
    ctx.set_kernel({"auto": api.KERNEL_AUTO, "popc": api.KERNEL_POPC, "umma": api.KERNEL_UMMA,
                    "fused": api.KERNEL_FUSED}[args.kernel])
    if world > 1:
        # K-split group, one process per GPU: exchange the CUDA IPC handles of the accumulator windows
        handle = ctx.group_export(n)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        ctx.group_join(rank, world, handles)
        ctx.group_set_alignment(length)
    ctx.set_problem(n, len_r, pair=True)
    ncell = api.cells(n)
    d_D = torch.zeros(ncell, dtype=torch.float64, device=dev)
    d_N = torch.zeros(ncell, dtype=torch.float64, device=dev)
    row_blk = api.group_row_block()                      # matrix rows are owned in blocks, dealt round-robin over the ranks
    total_basecmp = float(ncell) * length

    def step():
        # the packed rows stay where they are in HBM: the library is lent them and expands its operands straight from
        # the reference's words (ccg_put_samples_packed_dev_borrowed); --copy-rows takes the copying upload instead
        if args.copy_rows:
            ctx.put_samples_packed_dev(seqs_t.data_ptr(), masks_t.data_ptr(), n, seqs_t.stride(0))
        else:
            ctx.put_samples_packed_dev_borrowed(seqs_t.data_ptr(), masks_t.data_ptr(), n, seqs_t.stride(0))
        ctx.run_pair_dev(d_D.data_ptr(), d_N.data_ptr(), norm=0, min_length=1, min_cov=0.5, elem_size=8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    t1 = time.time()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t0, t1)
    launches = ctx.launches - launches0
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    value = total_basecmp / (ms_step * 1e-3)

    # ---- dominant kernel: average launch duration measured live (library events on this stream) ----
    kern_ms, expand_ms, compare_ms = [], [], []
    for _ in range(3):
        step()
        torch.cuda.synchronize()
        compare_ms.append(ctx.last_compare_ms())
        if use_umma:
            if ctx.last_phase_ms(0) >= 0:
                expand_ms.append(ctx.last_phase_ms(0))
            kern_ms.append(ctx.last_phase_ms(1))
        else:
            kern_ms.append(ctx.last_compare_ms())
    kern_ms = float(np.mean(kern_ms))
    compare_ms = float(np.mean(compare_ms))
    barrier()
    peaks, peaks_src = measured_peaks()
    int8_2x_bf16 = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    # the tensor pipe's own rate on this GPU under this box's power cap: a loads-free loop of the kernel's MMA kind
    # and shape (tcgen05, cta_group::2, 256x256), run about as long as the GEMM phase (sustained) and for a few ms
    # (burst).  MEASURED_PEAKS.json only holds bf16; 4 x / 2 x that figure is reported beside.
    i8_burst = i8_sustained = fp4_burst = fp4_sustained = None
    fp4_inexact = None
    is_fp4 = "mxf4" in ctx.last_kernel
    if use_umma:
        sustain_ms = max(50.0, min(kern_ms, 1000.0))
        i8_burst = ctx.measure_i8_peak(10.0)
        i8_sustained = ctx.measure_i8_peak(sustain_ms)
        fp4_burst, fp4_inexact = ctx.measure_fp4_peak(10.0)
        fp4_sustained, _ = ctx.measure_fp4_peak(sustain_ms)
        torch.cuda.synchronize()
        if is_fp4 and fp4_inexact != 0:
            raise SystemExit("bench.py: the kind::mxf4 accumulators are not exact on this device -- number withheld")
    own_peak = fp4_sustained if is_fp4 else i8_sustained
    pipe_peak = own_peak if (own_peak and own_peak > 0) else int8_2x_bf16
    # algorithmic work of this rank's GEMM launches: the USEFUL pairwise base comparisons (cells of the strict lower
    # triangle) x the bases of this rank's slice x 8 tensor ops
    my_basecmp = float(ncell) * len_r
    achieved = OPS_PER_BASECMP * my_basecmp / (kern_ms * 1e-3) / 1e12
    step_frac = OPS_PER_BASECMP * total_basecmp / (ms_step * 1e-3) / 1e12 / (world * pipe_peak)
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of this workload
    traffic = None
    tnote = None
    for tname in ("r02_umma2_mxf4_10k_traffic.json", "r01_umma2_mxf4_10k_traffic.json") if is_fp4 else ("r01_umma2_10k_traffic.json",):
        tpath = os.path.join(ROOT, "profiles", tname)
        if use_umma and world == 1 and os.path.exists(tpath) and traffic is None:
            with open(tpath) as f:
                tj = json.load(f)
            if tj["samples"] == n and tj["length"] == length and tj["kernel"] in ctx.last_kernel and \
                    tj.get("operands", "i8") == ("mxf4" if is_fp4 else "i8"):
                traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
                # a capture that cut the K axis into more launches than this run: scale to this run's launches per step
                nslabs_now = int(ctx.last_kernel.split("slabs=")[1].split()[0]) if "slabs=" in ctx.last_kernel else 1
                traffic = traffic * tj.get("launches_per_step_in_capture", nslabs_now) / max(nslabs_now, 1)
                tnote = f"DRAM bytes per GEMM launch from the committed ncu --set full capture profiles/{tname}"
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": pipe_peak,
        "unit": "TOP/s (e2m1 x e2m1 -> f32, kind::mxf4)" if is_fp4 else "TOP/s (int8)",
        "frac": achieved / pipe_peak, "traffic": traffic, "traffic_note": tnote,
        "kernel": ctx.last_kernel,
        "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / ms_step,
        "compare_phase_ms": compare_ms,
        "expand_ms": float(np.mean(expand_ms)) if expand_ms else None,
        "step_frac_all_gpus": step_frac,
        "step_frac_note": "8 ops x all base comparisons / (step time x n_gpus x peak): the whole step, every GPU, against the pipe",
        "peak_source": ("own loads-free tcgen05 microbenchmark of the kernel's MMA kind and shape "
                        "(ccg_measure_fp4_peak / ccg_measure_i8_peak), sustained: run as long as the GEMM phase, same "
                        "power cap" if own_peak and own_peak > 0 else
                        f"2 x bf16_tflops_sustained of {peaks_src} MEASURED_PEAKS.json"),
        "peak_fp4_burst": fp4_burst, "peak_fp4_sustained": fp4_sustained, "fp4_inexact_elements_at_1.6e7": fp4_inexact,
        "peak_i8_burst": i8_burst, "peak_i8_sustained": i8_sustained,
        "peak_2x_bf16_sustained": int8_2x_bf16, "frac_of_4x_bf16_sustained": achieved / (2 * int8_2x_bf16),
        "peaks_file": peaks_src,
        "algorithmic": f"{OPS_PER_BASECMP} tensor ops (4 MACs) per pairwise base comparison (K=4L contraction), "
                       f"useful cells only (strict lower triangle); rank 0's launches cover {len_r} of {length} bases",
    }

    if rank == 0:
        log(f"[rank 0] device-resident: {ms_step:.2f} ms/step, value {value:.4g} {UNIT}, kernel {kern_ms:.2f} ms, "
            f"frac {achieved / pipe_peak:.3f}, {ctx.last_kernel}")

    # ---- parity (outside the timed region): >= 1000 cells spread over EVERY macro tile (and, at N > 1, over every
    # rank's rows) against the oracle on the same samples.  The rows of those samples are gathered on rank 0. ----
    ids = parity_sample_ids(n, args.parity_samples)
    idx_t = torch.from_numpy(ids).to(dev)
    sub_s = seqs_t[idx_t].cpu().numpy().view(np.uint64)
    sub_m = masks_t[idx_t].cpu().numpy().view(np.uint32)
    cell_idx, cell_row = packed_index(ids)
    own = (cell_row // row_blk) % world == rank
    cidx_t = torch.from_numpy(cell_idx).to(dev)
    own_t = torch.from_numpy(own).to(dev)
    gD = torch.where(own_t, d_D[cidx_t], torch.zeros((), dtype=torch.float64, device=dev))
    gN = torch.where(own_t, d_N[cidx_t], torch.zeros((), dtype=torch.float64, device=dev))
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (sub_s, sub_m))
        dist.all_reduce(gD)
        dist.all_reduce(gN)
        sub_s = np.concatenate([p[0] for p in parts], axis=1)
        sub_m = np.concatenate([p[1] for p in parts], axis=1)
    parity = None
    if rank == 0:
        import oracle
        assert sub_s.shape[1] == api.words(length), (sub_s.shape, api.words(length))
        Do, No, _ = oracle.fsa_cmp_pair(sub_s, sub_m, np.ones(len(ids), np.uint8), length)
        parity = bool(np.array_equal(gD.cpu().numpy(), Do) and np.array_equal(gN.cpu().numpy(), No))
        tiles_hit = len({(int(a) // 256, int(b) // 256) for a, b in zip(cell_row, ids[np.tril_indices(len(ids), -1)[1]])})
        if not parity:
            raise SystemExit("bench.py: GPU result differs from the oracle -- number withheld")

    # ---- e2e: host buffers through the reference-facing C-ABI: ONE call per rank at every N.
    # ccg_fsa_cmp_thread_out(host row pointers in, host matrices out).  At N > 1 the rank's context is a member of
    # the K-split group: its rows are its slice of the alignment (pinned host memory), the library streams them
    # slab by slab under the GEMM, synchronises with the peers on the device, and copies the span of cells it owns
    # into the host matrices. ----
    e2e = None
    if not args.no_e2e:
        L = api.load()
        import ctypes as C
        row_s, row_m = W * 8, W * 4
        hs_ptr = L.ccg_host_alloc(n * row_s)
        hm_ptr = L.ccg_host_alloc(n * row_m)
        hD_ptr = L.ccg_host_alloc(max(ncell, 1) * 8)
        hN_ptr = L.ccg_host_alloc(max(ncell, 1) * 8)
        if not (hs_ptr and hm_ptr and hD_ptr and hN_ptr):
            raise SystemExit("bench.py: pinned host allocation failed")
        hs = np.ctypeslib.as_array(C.cast(hs_ptr, C.POINTER(C.c_uint64)), shape=(n, W))
        hm = np.ctypeslib.as_array(C.cast(hm_ptr, C.POINTER(C.c_uint32)), shape=(n, W))
        hs_t = torch.from_numpy(hs.view(np.int64))
        hm_t = torch.from_numpy(hm.view(np.int32))
        hs_t.copy_(seqs_t)                      # device -> pinned host, no pageable intermediate
        hm_t.copy_(masks_t)
        torch.cuda.synchronize()
        include = np.ones(n, dtype=np.uint8)
        dn, ginc = C.c_int(0), C.c_uint(0)
        sp = (C.c_void_p * n)(*[hs_ptr + k * row_s for k in range(n)])
        mp = (C.c_void_p * n)(*[hm_ptr + k * row_m for k in range(n)])

        def e2e_step():
            rc = L.ccg_fsa_cmp_thread_out(ctx._h, 1, hD_ptr, hN_ptr, 8, 1.0, n, len_r, sp, include.ctypes.data, mp,
                                          0, 1, 0.5, 0, C.byref(dn), C.byref(ginc))
            if rc:
                raise SystemExit("ccg_fsa_cmp_thread_out failed: " + L.ccg_last_error(ctx._h).decode())
        call = "ccg_fsa_cmp_thread_out(ctx, pair=1, host rows in pinned memory, host D/N out)" + \
               (f", one call per rank on its slice of the alignment (K-split group of {world})" if world > 1 else "")

        barrier()                 # the ranks leave the host allocations at different times: start the first call together
        # the floor under e2e: the same bytes as plain pinned-host -> device copies, all ranks at once
        evc0, evc1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evc0.record(stream)
        seqs_t.copy_(hs_t, non_blocking=True)
        masks_t.copy_(hm_t, non_blocking=True)
        evc1.record(stream)
        barrier()
        h2d_floor_ms = evc0.elapsed_time(evc1)
        if world > 1:
            tt = torch.tensor([h2d_floor_ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            h2d_floor_ms = float(tt.item())
        for _ in range(2):
            e2e_step()
        barrier()
        e_steps = max(2, min(args.steps, 5))
        ev0.record(stream)
        for _ in range(e_steps):
            e2e_step()
        ev1.record(stream)
        barrier()
        e_ms = ev0.elapsed_time(ev1) / e_steps
        if world > 1:
            tt = torch.tensor([e_ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_ms = float(tt.item())
        hD = np.ctypeslib.as_array(C.cast(hD_ptr, C.POINTER(C.c_double)), shape=(max(ncell, 1),))
        hN = np.ctypeslib.as_array(C.cast(hN_ptr, C.POINTER(C.c_double)), shape=(max(ncell, 1),))
        # the host-path result must equal the device-path result on the cells of the row blocks this rank owns
        same, owned = True, 0
        dD, dN = d_D.cpu().numpy(), d_N.cpu().numpy()
        for lo, hi in api.group_owned_blocks(n, rank, world):
            c_lo, c_hi = lo * (lo - 1) // 2, hi * (hi - 1) // 2
            same = same and bool(np.array_equal(hD[c_lo:c_hi], dD[c_lo:c_hi]) and np.array_equal(hN[c_lo:c_hi], dN[c_lo:c_hi]))
            owned += c_hi - c_lo
        del dD, dN
        same = same and owned == api.group_cells(n, rank, world) and owned > 0
        if world > 1:
            st = torch.tensor([1 if same else 0, owned], device=dev, dtype=torch.int64)
            dist.all_reduce(st)
            same = int(st[0].item()) == world and int(st[1].item()) == ncell      # the owned blocks tile the triangle
        if not same:
            raise SystemExit("bench.py: host-path result differs from the device-path result")
        e2e = {"value": total_basecmp / (e_ms * 1e-3), "unit": UNIT, "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(n * (api.words(length) * 12)), "d2h_bytes_per_step": int(2 * ncell * 8),
               "steps": e_steps, "call": call, "matches_device_path": same,
               "h2d_floor_ms": h2d_floor_ms,
               "h2d_floor_note": "the step's input bytes as plain cudaMemcpyAsync copies from the same pinned buffers, all ranks at "
                                 "once (max over ranks): what PCIe and the host memory of this box allow; the e2e step cannot be faster"}
        for p in (hs_ptr, hm_ptr, hD_ptr, hN_ptr):
            L.ccg_host_free(p)

    # ---- cpu baseline: the unmodified reference on this box's host cores (rank 0, N=1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ns_cap = min(n, 64 + 8 * cores)
        hs = seqs_t[:ns_cap].cpu().numpy().view(np.uint64)
        hm = masks_t[:ns_cap].cpu().numpy().view(np.uint32)
        v, kind, what, sec, threads, ns, Dc, Nc = cpu_reference_rate(hs, hm, length, cores, args.cpu_budget)
        same = bool(np.array_equal(Dc, d_D[:api.cells(ns)].cpu().numpy())
                    and np.array_equal(Nc, d_N[:api.cells(ns)].cpu().numpy()))
        if not same:
            raise SystemExit("bench.py: reference CPU result differs from the GPU result")
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": what,
                        "seconds": sec, "matches_gpu_bit_exact": same}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None,
            "dtype": (("e2m1 operands (+1/-1/0, exact), f32 accumulate of exact integers (tcgen05 kind::mxf4), int32 "
                       "split-K sums, f64 epilogue" if is_fp4 else
                       "int8 operands, int32 accumulate (tcgen05 kind::i8), f64 epilogue") if use_umma
                      else "u32 bit-planes (LOP3+POPC), u32 counters, f64 epilogue"),
            "data": "synthetic",
            "config": {"workload": f"{n} samples x {length} bp all-vs-all distance + inclusion matrix "
                                   f"(pair mode, -f 3 -n), " + ("BASELINE configs[2], the metric's own configuration, same job at every N"
                                                             if args.scaling == "strong" else
                                                             "sample count grown with sqrt(N)"),
                       "samples": n, "length": length, "pairs": ncell,
                       "partition": ("one GPU" if world == 1 else
                                     f"K split: each of the {world} GPUs holds 1/{world} of the alignment of every sample and runs all "
                                     f"macro tiles on it; int32 partial sums reduced by the owner of a matrix row through peer "
                                     f"pointers over NVLink inside the epilogue kernel (no NCCL collective on the data path)"),
                       "l2": "inputs (%.2f GB of planes per GPU) larger than the 126 MB L2; no explicit flush" %
                             (n * W * 12 / 1e9),
                       "step": ("encode (packed words -> bit planes) + operand expansion + compare + epilogue" if args.copy_rows else
                                "operand expansion straight from the resident packed words (rows lent to the library) + compare + "
                                "epilogue")},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(launches), "parity_vs_oracle": parity,
            "parity_cells_checked": int(len(cell_idx)), "parity_tiles_hit": int(tiles_hit),
            "parity_tiles_total": int(((n + 255) // 256) * ((n + 255) // 256 + 1) // 2),
        }
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
