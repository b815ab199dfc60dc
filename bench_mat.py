"""bench.py --workload mat: BASELINE configs[4] -- 2,000 KMA count matrices (.mat), all pairs, -d <method>, -r template.

The reference's path: ltdMatrixThrd (ltdmatrixthrd.c:376) -> cmpMats (matcmp.c:448) with the -d method's per-position
vector distance; for every cell (i, j) it re-opens, inflates and re-parses sample j's file.  Here every sample is
uploaded once (12 bytes per position) and all pairs are computed by k_matdist (ccphylo_b200/csrc/k_matdist.cu).

metric  pairwise position comparisons / s: one per-position distance between the count vectors of two samples.
step    value: the samples resident in HBM -> all pairs -> D, N and rowsInc on the host (the run call's own copies).
        e2e  : host count rows (the reference's in-memory format, 12 B per position) -> upload -> run -> host D / N.
N > 1   the POSITION axis is cut (the K split of the FASTA path): every rank holds a slice of every sample and
        computes raw per-pair sums over it (ccg_mat_run_partial); the sums are added over the ranks (one small NCCL
        all-reduce: 12 bytes per cell) and rank 0 applies the gates on the host (ccg_mat_finalize_host).
cpu_baseline / --impl reference: the UNMODIFIED reference binary (oracle/_ref/ccphylo dist -d <method> -t <cores>) on
        plain-text .mat files of a bounded sample of the same workload, its output compared with the GPU's.
"""
import ctypes as C
import json
import os
import shutil
import subprocess
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC = "pairwise position comparisons/sec (.mat count vectors)"
UNIT = "position-pairs/s"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ccphylo")
DISTINCT = 8


def make_samples(length, seed=5):
    """DISTINCT synthetic samples (cycled over the slots): called base ~ Poisson(40), errors Poisson(0.3) per other
    symbol, 1 % SNPs, 1 % low-depth rows; returns (ref bases, [counts6 (L, 6) u16])."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, size=length).astype(np.uint8)
    out = []
    for _ in range(DISTINCT):
        c = rng.poisson(0.3, size=(length, 6)).astype(np.uint16)
        call = np.where(rng.random(length) < 0.01, (ref + rng.integers(1, 4, size=length)) % 4, ref)
        depth = rng.poisson(40, size=length)
        depth = np.where(rng.random(length) < 0.01, rng.integers(0, 10, size=length), depth)
        c[np.arange(length), call] += depth.astype(np.uint16)
        out.append(np.ascontiguousarray(c))
    return ref, out


def reference_rate(ref, samples, length, method, cores, budget_s, steps=1, warmup=0):
    """Times the unmodified reference binary on ns samples x ls positions of the workload (sized for ~budget_s).
    Returns (rate, sample description, seconds, ns, ls, D cells as printed)."""
    import oracle

    if not os.path.exists(REF_BIN):
        return None
    # the reference manages ~1e7 position pairs / s / core including its O(n^2) re-parsing: size the sample for the budget
    ls = int(min(length, 200_000))
    pairs_wanted = budget_s * 1.0e7 * max(cores, 1) / ls
    ns = int(min(64, max(8, (1 + (1 + 8 * pairs_wanted) ** 0.5) / 2)))
    td = tempfile.mkdtemp(prefix="ccphylo_mat_")
    try:
        files = []
        for k in range(ns):
            path = os.path.join(td, f"s{k:03d}.mat")
            oracle.write_mat(path, "tmpl", ref[:ls], samples[k % DISTINCT][:ls])
            files.append(path)
        out = os.path.join(td, "ref.phy")
        cmd = [REF_BIN, "dist", "-r", "tmpl", "-d", method, "-t", str(cores), "-o", out, "-i"] + files
        times = []
        for k in range(warmup + steps):
            t0 = time.perf_counter()
            p = subprocess.run(cmd, capture_output=True, text=True)
            dt = time.perf_counter() - t0
            if p.returncode != 0:
                raise SystemExit("reference binary failed: " + p.stderr[-500:])
            if k >= warmup:
                times.append(dt)
        cells = []
        with open(out) as f:
            for line in f.read().split("\n")[2:]:
                cells.extend(float(x) for x in line.split("\t")[1:] if x)
        sec = float(np.mean(times))
        what = (f"first {ns} samples x first {ls} positions of the workload as plain-text .mat files, "
                f"ccphylo dist -d {method} -r tmpl -t {cores} (wall time of the whole command: parse + compare + print)")
        return ns * (ns - 1) / 2 * ls / sec, what, sec, ns, ls, np.array(cells)
    finally:
        shutil.rmtree(td, ignore_errors=True)


def mat_main(args, rank, world, local_rank, emit, log, ClockSampler, measured_peaks):
    n = args.samples or 2000
    length = args.length or 1_000_000
    method = args.method
    cores = os.cpu_count() or 1
    if args.impl == "reference":
        if rank != 0:
            return
        ref, samples = make_samples(min(length, 200_000))
        budget = min(args.cpu_budget, 120.0 / max(1, args.steps + args.warmup))
        r = reference_rate(ref, samples, length, method, cores, budget, steps=args.steps, warmup=args.warmup)
        if r is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref/ccphylo was not built (needs /root/reference at build time)"})
            return
        v, what, sec, ns, ls, _ = r
        emit({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "u16 counts, f64 per-position distances and sums", "data": "synthetic",
              "config": {"workload": f"{n} count matrices x {length} positions, all pairs, -d {method}; bounded sample: {what}"},
              "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": what},
              "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return

    import torch
    import torch.distributed as dist

    from ccphylo_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    bounds = [length if g == world else (length * g // world) // 32 * 32 for g in range(world + 1)]
    p0, p1 = bounds[rank], bounds[rank + 1]
    len_r = p1 - p0
    t_gen = time.time()
    ref, samples = make_samples(length)
    # this rank's slice of every distinct sample in PINNED host memory: the e2e arm's uploads are real asynchronous
    # copies (the contract's "from pinned host memory"), not staged through the driver's bounce buffer
    mine = []
    for s_ in samples:
        t_ = torch.from_numpy(np.ascontiguousarray(s_[p0:p1])).pin_memory()
        mine.append(t_.numpy())
    log(f"[rank {rank}] {DISTINCT} distinct samples x {length} positions generated in {time.time() - t_gen:.1f}s; slice {p0}..{p1}")
    L = api.load()
    ctx = api.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    mid, order = api.mat_method(method)
    ncell = api.cells(n)
    total_pp = float(ncell) * length
    lens = np.full(n, length, np.int32)
    include = np.ones(n, np.uint8)
    hD = np.zeros(ncell, np.float64)
    hN = np.zeros(ncell, np.float64)
    h_rows = np.zeros(ncell, np.uint32)
    p_dist = np.zeros(ncell, np.float64)
    p_rows = np.zeros(ncell, np.uint32)
    dn = C.c_int(0)

    def upload():
        ctx.mat_set_problem(n, len_r)
        for i in range(n):
            ctx.mat_put_sample(i, mine[i % DISTINCT])
        ctx.sync()

    def run():
        if world == 1:
            rc = L.ccg_mat_run(ctx._h, include.ctypes.data, mid, order, 0.05, 0, 15, 1, 0.5, 8, 1.0, hD.ctypes.data, hN.ctypes.data,
                               C.byref(dn), h_rows.ctypes.data)
            if rc:
                raise SystemExit("ccg_mat_run failed: " + L.ccg_last_error(ctx._h).decode())
            return
        rc = L.ccg_mat_run_partial(ctx._h, include.ctypes.data, mid, order, 0.05, 15, p_dist.ctypes.data, p_rows.ctypes.data, C.byref(dn))
        if rc:
            raise SystemExit("ccg_mat_run_partial failed: " + L.ccg_last_error(ctx._h).decode())
        td = torch.from_numpy(p_dist).to(dev, non_blocking=True)
        tr = torch.from_numpy(p_rows.view(np.int32)).to(dev, non_blocking=True)
        dist.all_reduce(td)
        dist.all_reduce(tr)
        if rank == 0:
            sd = td.cpu().numpy()
            sr = tr.cpu().numpy().view(np.uint32)
            rc = L.ccg_mat_finalize_host(n, include.ctypes.data, lens.ctypes.data, sd.ctypes.data, sr.ctypes.data, 0, 1, 0.5, 8, 1.0,
                                         hD.ctypes.data, hN.ctypes.data, h_rows.ctypes.data, None)
            if rc:
                raise SystemExit("ccg_mat_finalize_host failed")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_up = time.time()
    upload()
    log(f"[rank {rank}] uploaded {n} x {len_r} positions in {time.time() - t_up:.1f}s")
    for _ in range(args.warmup):
        run()
    barrier()
    launches0 = ctx.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record(stream)
    kms = []
    for _ in range(args.steps):
        run()
        kms.append(ctx.last_compare_ms())
    ev1.record(stream)
    barrier()
    t1 = time.time()
    ms_step = (t1 - t0) * 1e3 / args.steps                        # host clock: the run calls copy their results to the host and return
    clocks = sampler.stop(t0, t1)
    launches = ctx.launches - launches0
    if world > 1:
        tt = torch.tensor([ms_step], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_step = float(tt.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    value = total_pp / (ms_step * 1e-3)
    kern_ms = float(np.mean(kms))
    peaks, peaks_src = measured_peaks()
    # algorithmic bytes of k_matdist: every (32 x 32 tile, block of positions) reads 2 x 32 samples x 12 B per position
    tiles = ((n + 31) // 32) * ((n + 31) // 32 + 1) // 2
    alg_bytes = float(tiles) * 64 * 12 * len_r
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    hbm = float(peaks["hbm_gbs"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_matdist_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("samples") == n and tj.get("positions") == len_r and tj.get("method") == method:
            traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
        "kernel": ctx.last_kernel, "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / ms_step,
        "position_pairs_per_s_kernel": float(ncell) * len_r / (kern_ms * 1e-3),
        "note": "the contract's two bounds do not describe this kernel: 12 B per sample and position serve 32 pairs, so HBM is idle "
                "(frac << 1) and there is no contraction for the tensor pipe; it is bound by the CUDA cores -- one fp64 divide, a "
                "handful of fp64 adds / multiplies and ~30 integer instructions per position pair (pipe utilisation from ncu: "
                "profiles/r02_matdist_ncu.txt)",
        "peaks_file": peaks_src,
        "algorithmic": "2 x 32 samples x 12 B per position and 32 x 32 tile of pairs",
    }
    if rank == 0:
        log(f"[rank 0] {ms_step:.1f} ms/step, kernel {kern_ms:.1f} ms, {value:.4g} {UNIT}")

    # ---- parity: cells spread over the whole matrix against the oracle at FULL length (N exact, D within 1e-6 relative:
    # the north-star tolerance for floating point), plus -- on one GPU -- a whole small matrix ----
    parity = None
    cpu_baseline = None
    if rank == 0:
        import oracle
        rng = np.random.default_rng(77)
        checked = 0
        parity = True
        cache = {}
        for _ in range(24):
            i = int(rng.integers(1, n))
            j = int(rng.integers(0, i))
            key = (i % DISTINCT, j % DISTINCT)
            if key not in cache:
                pair = np.stack([samples[key[1]], samples[key[0]]])           # the earlier sample first
                Do, No, _ = oracle.mat_matrix(pair, pair.astype(np.uint32).sum(axis=2).astype(np.uint32), np.full(2, length, np.int32), None,
                                              method=method)
                cache[key] = (float(Do[0]), float(No[0]))
            d_o, n_o = cache[key]
            c = i * (i - 1) // 2 + j
            parity = parity and bool(hN[c] == n_o and h_rows[c] == n_o and abs(hD[c] - d_o) <= 1e-6 * abs(d_o) + 1e-12)
            checked += 1
        if world == 1:
            ns, ls = min(n, 24), min(length, 100_000)
            counts = np.stack([samples[i % DISTINCT][:ls] for i in range(ns)])
            c2 = api.Context(local_rank)
            c2.mat_set_problem(ns, ls)
            for i in range(ns):
                c2.mat_put_sample(i, counts[i])
            Dg, Ng, _, _ = c2.mat_run(None, method=method)
            c2.close()
            Do, No, _ = oracle.mat_matrix(counts, counts.astype(np.uint32).sum(axis=2).astype(np.uint32), np.full(ns, ls, np.int32), None,
                                          method=method)
            parity = parity and bool(np.array_equal(Ng, No) and np.all(np.abs(Dg - Do) <= 1e-6 * np.abs(Do) + 1e-12))
        if not parity:
            raise SystemExit("bench.py: .mat result differs from the oracle -- number withheld")

    # ---- e2e: host count rows -> upload -> run -> host D / N (rank r's rows are its slice of every sample) ----
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.time()
        upload()
        run()
        barrier()
        e_ms = (time.time() - t0) * 1e3
        if world > 1:
            tt = torch.tensor([e_ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_ms = float(tt.item())
        e2e = {"value": total_pp / (e_ms * 1e-3), "unit": UNIT, "ms_per_step": e_ms, "h2d_bytes_per_step": int(n) * length * 12,
               "d2h_bytes_per_step": int(ncell) * 20, "steps": 1,
               "call": "ccg_mat_set_problem + n x ccg_mat_put_sample (host rows, 12 B per position) + ccg_mat_run" +
                       (" (per rank: its positions, ccg_mat_run_partial; sums all-reduced; ccg_mat_finalize_host on rank 0)" if world > 1 else "")}

    # ---- cpu baseline: the reference binary on a bounded sample (rank 0, N = 1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = reference_rate(ref, samples, length, method, cores, args.cpu_budget)
        if r is not None:
            v, what, sec, ns2, ls2, cells_ref = r
            c2 = api.Context(local_rank)
            c2.mat_set_problem(ns2, ls2)
            for i in range(ns2):
                c2.mat_put_sample(i, samples[i % DISTINCT][:ls2])
            Dg, _, _, _ = c2.mat_run(None, method=method)
            c2.close()
            agree = bool(len(cells_ref) == len(Dg) and np.all(np.abs(Dg - cells_ref) <= 1e-6 * np.abs(cells_ref) + 6e-10))
            if not agree:
                raise SystemExit("bench.py: the reference binary's matrix differs from the GPU's beyond 1e-6")
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": what, "seconds": sec,
                            "matches_gpu_within_1e-6": agree}

    if rank == 0:
        emit({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "u16 counts, int64 dot products, f64 per-position distances and sums", "data": "synthetic",
              "config": {"workload": f"{n} KMA count matrices (.mat) x {length} positions, all pairs, -d {method}, depth gate 15 "
                                     f"(BASELINE configs[4]); {DISTINCT} distinct synthetic samples cycled over the slots -- the work per "
                                     f"position pair does not depend on the values",
                         "samples": n, "positions": length, "pairs": ncell, "method": method,
                         "partition": "one GPU" if world == 1 else f"position axis cut over {world} GPUs, raw sums all-reduced (NCCL, 12 B per cell)",
                         "l2": "the count store (%.1f GB per GPU) is larger than the 126 MB L2; no explicit flush" % (n * len_r * 12 / 1e9),
                         "step": "all pairs over resident samples + D / N / rowsInc copied to the host"},
              "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
              "parity_vs_oracle": parity})
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
