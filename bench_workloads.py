"""bench.py --workload ring | mat: BASELINE configs[3] and configs[4], measured to the same contract as the default arm.

ring  configs[3]: 100,000 samples x 2.9 Mbp (S. aureus length) on 8 GPUs -- 12,500 samples per GPU at other N
      ("weak").  The sample set is too large for one GPU (109 GB of bit planes) and its n x n accumulators too large
      for any (80 GB), so every GPU holds a SEQUENCE SHARD -- bases [b_r, b_r+1) of every sample -- and the lower
      triangle is run in windows of macro-tile rows; each window's int32 partial sums are reduced over NVLink by the
      owners of its rows inside the epilogue kernel (ccphylo_b200/csrc/ccg_group.cu).  The shards never move: what
      the north star sketches as a ring exchange of shards is replaced by a reduction of partial sums, 35 GB per
      GPU over NVLink per run instead of 4 x 13.6 GB of shard traffic plus an idle half step.  (The NCCL ring of
      round 1 is kept as scripts/ring_demo.py + scripts/ring.py for comparison.)  Results stay in compact per-rank
      buffers (ccg_group_set_output): a rank holds the rows it owns.
mat   configs[4]: 2,000 KMA count matrices (.mat), -d method, positions cut over the GPUs.
"""
import ctypes as C
import json
import math
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC = "pairwise base comparisons/sec"
UNIT = "base-cmp/s"
OPS_PER_BASECMP = 8


def ring_parity_ids(n, seed=4242):
    """one sample per 256-row block (every off-diagonal macro tile gets a pair), a second one in every 8th block"""
    rng = np.random.default_rng(seed)
    ids = []
    for b in range((n + 255) // 256):
        lo, hi = b * 256, min(n, b * 256 + 256)
        k = min(2 if b % 8 == 0 else 1, hi - lo)
        ids.extend(rng.choice(np.arange(lo, hi), size=k, replace=False).tolist())
    return np.array(sorted(ids), dtype=np.int64)


def compact_row_base(n, rank, world, blk):
    """offset of every owned row in a rank's compact result buffer (all samples included): ccg_group_set_output"""
    base = np.full(n, -1, dtype=np.int64)
    rows = np.concatenate([np.arange(b * blk, min(n, b * blk + blk)) for b in range(rank, (n + blk - 1) // blk, world)]
                          or [np.zeros(0, np.int64)]).astype(np.int64)
    base[rows] = np.concatenate([[0], np.cumsum(rows)[:-1]]) if len(rows) else []
    return base


def ring_main(args, rank, world, local_rank, emit, log, ClockSampler, measured_peaks):
    import torch
    import torch.distributed as dist

    import bench
    from ccphylo_b200 import api, synth
    import synth_torch  # noqa: E402

    if args.impl == "reference":
        if rank == 0:
            args.samples = args.samples or 12_500 * max(world, 1)
            args.length = args.length or 2_900_000
            bench.reference_arm(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n = args.samples or 12_500 * world
    length = args.length or 2_900_000
    slices = api.group_slices(length, world)
    b0, b1 = slices[rank], slices[rank + 1]
    len_r = b1 - b0
    W = api.words(len_r)
    t_gen = time.time()
    seqs_t, masks_t = synth_torch.make_packed_torch(n, len_r, seed=4 + 1000 * rank, device=dev)
    torch.cuda.synchronize()
    log(f"[rank {rank}] generated {n} x {len_r} bp (bases {b0}..{b1} of {length}) in {time.time() - t_gen:.1f}s")

    ctx = api.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_kernel(api.KERNEL_UMMA)
    blk = api.group_row_block()
    if world > 1:
        handle = ctx.group_export(n)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        ctx.group_join(rank, world, handles)
        ctx.group_set_alignment(length)
        ctx.group_set_output(True)
    ctx.set_problem(n, len_r, pair=True)
    ncell = api.cells(n)
    own_cells = api.group_cells(n, rank, world) if world > 1 else ncell
    d_D = torch.zeros(max(own_cells, 1), dtype=torch.float64, device=dev)
    d_N = torch.zeros(max(own_cells, 1), dtype=torch.float64, device=dev)
    total_basecmp = float(ncell) * length

    def step():
        ctx.put_samples_packed_dev_borrowed(seqs_t.data_ptr(), masks_t.data_ptr(), n, seqs_t.stride(0))   # rows lent, read in place
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
    kern_ms = ctx.last_phase_ms(1)
    expand_ms = ctx.last_phase_ms(0)
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    value = total_basecmp / (ms_step * 1e-3)
    peaks, peaks_src = measured_peaks()
    sustain_ms = max(50.0, min(kern_ms, 1000.0))
    fp4_burst, fp4_inexact = ctx.measure_fp4_peak(10.0)
    fp4_sustained, _ = ctx.measure_fp4_peak(sustain_ms)
    torch.cuda.synchronize()
    if fp4_inexact != 0:
        raise SystemExit("bench.py: the kind::mxf4 accumulators are not exact on this device -- number withheld")
    achieved = OPS_PER_BASECMP * float(ncell) * len_r / (kern_ms * 1e-3) / 1e12
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": fp4_sustained, "unit": "TOP/s (e2m1 x e2m1 -> f32, kind::mxf4)",
        "frac": achieved / fp4_sustained, "traffic": None, "kernel": ctx.last_kernel, "kernel_ms": kern_ms,
        "kernel_share_of_step": kern_ms / ms_step, "expand_ms": expand_ms,
        "kernel_ms_note": "all windows of the run: GEMM launches + per-window barrier and peer-reduce epilogue (library events on the stream)",
        "step_frac_all_gpus": OPS_PER_BASECMP * total_basecmp / (ms_step * 1e-3) / 1e12 / (world * fp4_sustained),
        "peak_source": "own loads-free tcgen05 kind::mxf4 loop of the kernel's MMA shape, run as long as the compare phase (ccg_measure_fp4_peak)",
        "peak_fp4_burst": fp4_burst, "peak_2x_bf16_sustained": 2.0 * float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])),
        "peaks_file": peaks_src,
        "algorithmic": f"{OPS_PER_BASECMP} tensor ops per pairwise base comparison, useful cells only; rank 0's launches cover {len_r} of {length} bases",
    }

    if rank == 0:
        log(f"[rank 0] device-resident: {ms_step:.2f} ms/step, value {value:.4g} {UNIT}, compare phase {kern_ms:.2f} ms, "
            f"frac {achieved / fp4_sustained:.3f}, {ctx.last_kernel}")

    # ---- parity: pairs spread over every off-diagonal macro tile (and every 8th diagonal one), values read from the
    # owners' compact buffers, against the oracle on the gathered rows ----
    ids = ring_parity_ids(n)
    idx_t = torch.from_numpy(ids).to(dev)
    sub_s = seqs_t[idx_t].cpu().numpy().view(np.uint64)
    sub_m = masks_t[idx_t].cpu().numpy().view(np.uint32)
    r, c = np.tril_indices(len(ids), -1)
    hi, lo = ids[r], ids[c]
    if world > 1:
        base = compact_row_base(n, rank, world, blk)
        own = (hi // blk) % world == rank
        off = np.where(own, base[hi] + lo, 0)
    else:
        own = np.ones(len(hi), bool)
        off = hi * (hi - 1) // 2 + lo
    off_t = torch.from_numpy(off).to(dev)
    own_t = torch.from_numpy(own).to(dev)
    zero = torch.zeros((), dtype=torch.float64, device=dev)
    gD = torch.where(own_t, d_D[off_t], zero)
    gN = torch.where(own_t, d_N[off_t], zero)
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
        t_or = time.time()
        # raw integer counts from the oracle on all host cores, then the pair epilogue of fsacmpthrd.c:419-434 (norm 0)
        mo, no = oracle.raw_pair_matrix(sub_s, sub_m, length, nthreads=min(os.cpu_count() or 8, 64))
        min_len = max(1, int(0.5 * length))
        Do = np.where(no >= min_len, mo.astype(np.float64), -1.0)
        No = no.astype(np.float64)
        parity = bool(np.array_equal(gD.cpu().numpy(), Do) and np.array_equal(gN.cpu().numpy(), No))
        log(f"[rank 0] oracle on {len(ids)} samples / {len(hi)} cells in {time.time() - t_or:.1f}s: parity {parity}")
        if not parity:
            raise SystemExit("bench.py: GPU result differs from the oracle -- number withheld")

    # ---- e2e: one call per rank, host rows of its sequence shard in pinned memory, its rows of D / N out ----
    e2e = None
    if not args.no_e2e:
        L = api.load()
        row_s, row_m = W * 8, W * 4
        hs_ptr, hm_ptr = L.ccg_host_alloc(n * row_s), L.ccg_host_alloc(n * row_m)
        hD_ptr, hN_ptr = L.ccg_host_alloc(max(own_cells, 1) * 8), L.ccg_host_alloc(max(own_cells, 1) * 8)
        if not (hs_ptr and hm_ptr and hD_ptr and hN_ptr):
            raise SystemExit("bench.py: pinned host allocation failed")
        hs = np.ctypeslib.as_array(C.cast(hs_ptr, C.POINTER(C.c_uint64)), shape=(n, W))
        hm = np.ctypeslib.as_array(C.cast(hm_ptr, C.POINTER(C.c_uint32)), shape=(n, W))
        torch.from_numpy(hs.view(np.int64)).copy_(seqs_t)
        torch.from_numpy(hm.view(np.int32)).copy_(masks_t)
        torch.cuda.synchronize()
        include = np.ones(n, dtype=np.uint8)
        dn, ginc = C.c_int(0), C.c_uint(0)
        sp = (C.c_void_p * n)(*[hs_ptr + k * row_s for k in range(n)])
        mp = (C.c_void_p * n)(*[hm_ptr + k * row_m for k in range(n)])

        def e2e_step():
            rc = L.ccg_fsa_cmp_thread_out(ctx._h, 1, hD_ptr, hN_ptr, 8, 1.0, n, len_r, sp, include.ctypes.data, mp, 0, 1, 0.5, 0,
                                          C.byref(dn), C.byref(ginc))
            if rc:
                raise SystemExit("ccg_fsa_cmp_thread_out failed: " + L.ccg_last_error(ctx._h).decode())
        barrier()                 # pinning tens of GB takes the ranks different times: start the first call together
        e2e_step()
        barrier()
        e_steps = max(1, min(args.steps, 2))
        ev0.record(stream)
        for _ in range(e_steps):
            e2e_step()
        ev1.record(stream)
        barrier()
        e_ms = ev0.elapsed_time(ev1) / e_steps
        hD = np.ctypeslib.as_array(C.cast(hD_ptr, C.POINTER(C.c_double)), shape=(max(own_cells, 1),))
        same = bool(np.array_equal(hD[:own_cells], d_D[:own_cells].cpu().numpy()))
        if world > 1:
            tt = torch.tensor([e_ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_ms = float(tt.item())
            st = torch.tensor([1 if same else 0], device=dev, dtype=torch.int64)
            dist.all_reduce(st)
            same = int(st.item()) == world
        if not same:
            raise SystemExit("bench.py: host-path result differs from the device-path result")
        e2e = {"value": total_basecmp / (e_ms * 1e-3), "unit": UNIT, "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(n * api.words(length) * 12), "d2h_bytes_per_step": int(2 * ncell * 8), "steps": e_steps,
               "call": "ccg_fsa_cmp_thread_out(ctx, pair=1, host rows of the rank's sequence shard in pinned memory, the rank's rows of "
                       "D / N out), one call per rank", "matches_device_path": same}
        for p in (hs_ptr, hm_ptr, hD_ptr, hN_ptr):
            L.ccg_host_free(p)

    if rank == 0:
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "e2m1 operands (+1/-1/0, exact), f32 accumulate of exact integers (tcgen05 kind::mxf4), int32 split-K sums, f64 epilogue",
            "data": "synthetic",
            "config": {"workload": f"{n} samples x {length} bp all-vs-all distance + inclusion matrix (pair mode, -f 3 -n): BASELINE "
                                   f"configs[3] is 100,000 samples on 8 GPUs; 12,500 samples per GPU at other N",
                       "samples": n, "length": length, "pairs": ncell,
                       "partition": (f"sequence shards: each of the {world} GPUs holds 1/{world} of the bases of every sample; windows of "
                                     f"macro-tile rows, int32 partial sums reduced by the row owners through peer pointers over NVLink "
                                     f"inside the epilogue kernel; no shard ever moves" if world > 1 else "one GPU"),
                       "l2": "inputs far larger than the 126 MB L2; no explicit flush",
                       "step": "operand expansion straight from the resident packed words + windows of (GEMM, peer reduce + epilogue)"},
            "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
            "parity_vs_oracle": parity, "parity_cells_checked": int(len(hi)), "parity_samples": int(len(ids)),
        })
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main(args, rank, world, local_rank, emit, log, ClockSampler, measured_peaks):
    if args.workload == "ring":
        return ring_main(args, rank, world, local_rank, emit, log, ClockSampler, measured_peaks)
    import bench_mat
    return bench_mat.mat_main(args, rank, world, local_rank, emit, log, ClockSampler, measured_peaks)
