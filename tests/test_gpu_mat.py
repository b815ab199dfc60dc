"""GPU parity of the count-matrix (.mat) path: k_matdist through the C-ABI (ccg_mat_*) against the
CPU oracle (oracle/mat_oracle.c, pinned to the reference's golden text) for all 18 -d methods.
Bar: the inclusion counts N (rowsInc) and the -1 / 0 cells of pairs without sufficient overlap are
exact; distances are within 1e-6 relative (north star) -- the GPU adds the per-position terms per
K slice, the reference strictly left to right, and pow / erf differ from glibc in the last ulps."""
import numpy as np
import pytest

import helpers
import oracle
from ccphylo_b200 import api

pytestmark = pytest.mark.gpu
REL_TOL = 1e-6
ALL = ["cos", "z", "chi2", "nchi2", "c", "nc", "p", "np", "bc", "nbc", "l1", "l2", "linf", "l3", "nl1", "nl2", "nlinf",
       "nl3"]


def random_counts(n, length, seed, depth=40, low=0.02):
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, size=length)
    counts = rng.poisson(0.3, size=(n, length, 6)).astype(np.uint16)
    for i in range(n):
        call = np.where(rng.random(length) < 0.02, (ref + rng.integers(1, 4, size=length)) % 4, ref)
        d = rng.poisson(depth, size=length)
        d = np.where(rng.random(length) < low, rng.integers(0, 6, size=length), d)
        counts[i, np.arange(length), call] += d.astype(np.uint16)
    totals = counts.astype(np.uint32).sum(axis=2).astype(np.uint32)
    return counts, totals


def close(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= REL_TOL * np.abs(b) + 1e-12)


@pytest.fixture(scope="module")
def ctx(built):
    c = api.Context()
    yield c
    c.close()


@pytest.mark.parametrize("method", ALL)
def test_all_methods_against_the_oracle(ctx, method):
    n, length = 37, 2000 + 13
    counts, totals = random_counts(n, length, seed=ALL.index(method) + 5)
    lens = np.full(n, length, np.int32)
    include = np.ones(n, np.uint8)
    include[[4, 20]] = 0
    ctx.mat_set_problem(n, length)
    for i in range(n):
        ctx.mat_put_sample(i, counts[i], totals[i])
    for norm in (0, 1000):
        D, N, dn, rows = ctx.mat_run(include, method=method, norm=norm)
        Do, No, dno = oracle.mat_matrix(counts, totals, lens, include, method=method, norm=norm)
        assert dn == dno == n - 2
        assert np.array_equal(N, No)
        assert np.array_equal(rows, No.astype(np.uint32))
        assert close(D, Do), float(np.max(np.abs(D - Do) / np.maximum(np.abs(Do), 1e-300)))
    assert "k_matdist" in ctx.last_kernel


def test_cos_kernel_deep_and_identical_samples(ctx):
    # k_matdist_cos: stages whose counts stay below 20,725 take the int32 dot product and the two-fma division, deeper
    # stages the literal path (64-bit sums, a real division); identical samples give distances of (almost) zero; counts
    # stay below 46,341, where the reference's int products start to wrap (matcmp.c:429-437)
    n, length = 70, 1500 + 7
    counts, totals = random_counts(n, length, seed=99, depth=300)
    rng = np.random.default_rng(5)
    for i in range(0, n, 3):
        s = int(rng.integers(0, length - 200))
        counts[i, s:s + 150, :4] = rng.integers(15000, 46000, size=(150, 4)).astype(np.uint16)
    counts[11] = counts[10]
    counts[40] = counts[10]
    counts[5, 300:330] = 0                                   # positions without a single read: c1 = 0, never counted
    totals = counts.astype(np.uint32).sum(axis=2).astype(np.uint32)
    lens = np.full(n, length, np.int32)
    lens[7] = 900
    counts[7, 900:] = 0
    totals[7, 900:] = 0
    ctx.mat_set_problem(n, length)
    for i in range(n):
        ctx.mat_put_sample(i, counts[i, :lens[i]], totals[i, :lens[i]])
    for kw in (dict(min_depth=15), dict(min_depth=0, min_cov=0.0, norm=1000), dict(min_depth=100, min_cov=0.3)):
        D, N, dn, rows = ctx.mat_run(None, method="cos", **kw)
        Do, No, dno = oracle.mat_matrix(counts, totals, lens, None, method="cos", **kw)
        assert dn == dno and np.array_equal(N, No), kw
        assert close(D, Do), (kw, float(np.max(np.abs(D - Do))))
    assert "k_matdist_cos" in ctx.last_kernel


def test_depth_gate_overlap_gate_and_short_samples(ctx):
    n, length = 18, 900
    counts, totals = random_counts(n, length, seed=99, low=0.3)
    lens = np.full(n, length, np.int32)
    # two samples whose covered halves are disjoint: D = -1, N = 0 for that pair
    counts[5, : length // 2 + 40] = 0
    counts[9, length // 2 - 40:] = 0
    lens[12] = 700                                         # a shorter template instance (rows beyond are absent)
    counts[12, 700:] = 0
    totals = counts.astype(np.uint32).sum(axis=2).astype(np.uint32)
    ctx.mat_set_problem(n, length)
    for i in range(n):
        ctx.mat_put_sample(i, counts[i, :lens[i]], totals[i, :lens[i]])
    for kw in (dict(min_depth=15, min_cov=0.3), dict(min_depth=30, min_cov=0.1, norm=1000000), dict(min_depth=0, min_cov=0.0)):
        D, N, dn, rows = ctx.mat_run(None, method="cos", **kw)
        Do, No, dno = oracle.mat_matrix(counts, totals, lens, None, method="cos", **kw)
        assert dn == dno == n
        assert np.array_equal(N, No)
        assert np.array_equal(D == -1.0, Do == -1.0)
        assert close(D, Do)
    assert (Do == -1.0).any()


@pytest.mark.parametrize("elem,scale", [(4, 1.0), (2, 10.0), (1, 0.1)])
def test_cell_types(ctx, elem, scale):
    n, length = 20, 640
    counts, totals = random_counts(n, length, seed=7)
    ctx.mat_set_problem(n, length)
    for i in range(n):
        ctx.mat_put_sample(i, counts[i], totals[i])
    D8, N8, dn, _ = ctx.mat_run(None, method="l1")
    D, N, _, _ = ctx.mat_run(None, method="l1", elem_size=elem, byte_scale=scale)
    if elem == 4:
        assert np.array_equal(D, D8.astype(np.float32)) and np.array_equal(N, N8.astype(np.float32))
    else:
        # dtouc(value, 0.5) truncated into the cell (bytescale.h:22); l1 distances are integers
        mask = (1 << (8 * elem)) - 1
        assert np.array_equal(D.astype(np.int64), (D8 * scale + 0.5).astype(np.int64) & mask)
        assert np.array_equal(N.astype(np.int64), (N8 * scale + 0.5).astype(np.int64) & mask)


G = helpers.load_golden("mat_dist.json")
FILES = [c for c in G["cases"] if c["mode"] == "files"]


@pytest.mark.parametrize("case", FILES, ids=lambda c: c["name"])
def test_golden_cases_through_the_abi(ctx, case):
    """The reference's own numbers: parse the fixture's .mat text, gate the samples as
    ltdMatrixThrd does, run the GPU path, compare with the printed matrices."""
    from test_mat_oracle_golden import mat_args
    o = mat_args(case["args"])
    parsed = [helpers.parse_mat(G["pool"][k], case["template"]) for k in case["text_ids"]]
    keep = [pm for pm in parsed if pm is not None and helpers.mat_sample_gate(pm[1], o["min_depth"], o["min_length"], o["min_cov"])]
    lmax = max(len(t) for _, t in keep)
    ctx.mat_set_problem(len(keep), lmax)
    for k, (c, t) in enumerate(keep):
        ctx.mat_put_sample(k, c, t)
    D, N, dn, _ = ctx.mat_run(None, method=o["method"], norm=o["norm"], min_depth=o["min_depth"], min_length=o["min_length"],
                              min_cov=o["min_cov"])
    (names, dref), = helpers.parse_phy(case["phy"])
    (_, nref), = helpers.parse_phy(case["num"])
    assert dn == len(names)
    assert np.array_equal(N, np.array(nref))
    # the text carries `precision` decimals
    tol = 0.51 * 10.0 ** (-o["precision"])
    assert np.all(np.abs(D - np.array(dref)) <= tol + REL_TOL * np.abs(np.array(dref)))


def test_rank_partition_covers_every_cell_once(built):
    n, length = 70, 800
    counts, totals = random_counts(n, length, seed=3)
    lens = np.full(n, length, np.int32)
    Do, No, _ = oracle.mat_matrix(counts, totals, lens, None, method="chi2")
    world = 3
    Dsum, Nsum, hits = np.zeros_like(Do), np.zeros_like(No), np.zeros_like(No)
    for r in range(world):
        with api.Context() as c:
            c.set_partition(r, world)
            c.mat_set_problem(n, length)
            for i in range(n):
                c.mat_put_sample(i, counts[i], totals[i])
            D, N, dn, _ = c.mat_run(None, method="chi2")
            Dsum += D
            Nsum += N
            hits += (N != 0)
    assert np.array_equal(Nsum, No) and hits.max() == 1
    assert close(Dsum, Do)


# ---- the 12-byte store: totals recomputed on the device, a side plane for depths above 65,535 ----
def test_row_totals_that_are_not_the_sum_of_the_stored_counts(ctx):
    # the reference keeps 16 bits of a count but the whole depth in the row total (matparse.c:246-258): a depth of
    # 70,000 is stored as 4,464 with total 70,000 + the rest.  Such samples carry their totals separately.
    n, length = 9, 700
    counts, totals = random_counts(n, length, seed=7)
    lens = np.full(n, length, np.int32)
    totals[3, 100:140] += 65536                              # what a count of 65,536 + c leaves behind
    totals[6, ::50] += 3 * 65536
    ctx.mat_set_problem(n, length)
    for i in range(n):
        ctx.mat_put_sample(i, counts[i], totals[i])
    for method in ("cos", "nchi2", "bc"):
        D, N, dn, rows = ctx.mat_run(None, method=method, min_depth=30)
        Do, No, dno = oracle.mat_matrix(counts, totals, lens, None, method=method, min_depth=30)
        assert np.array_equal(N, No) and close(D, Do), method
    # putting an ordinary sample into the slot again removes the override
    totals[3] = counts[3].astype(np.uint32).sum(axis=1)
    ctx.mat_put_sample(3, counts[3], totals[3])
    D, N, dn, rows = ctx.mat_run(None, method="nchi2", min_depth=30)
    Do, No, dno = oracle.mat_matrix(counts, totals, lens, None, method="nchi2", min_depth=30)
    assert np.array_equal(N, No) and close(D, Do)


# ---- position split over several GPUs (members on the devices that exist; all on GPU 0 on a 1-GPU box) ----
@pytest.mark.parametrize("members", [2, 3])
@pytest.mark.parametrize("method", ["cos", "chi2", "nl2"])
def test_position_split_over_members(built, monkeypatch, members, method):
    import torch
    ngpu = max(torch.cuda.device_count(), 1)
    monkeypatch.setenv("CCG_MULTI_FORCE", "1")
    n, length = 70, 3000 + 7
    counts, totals = random_counts(n, length, seed=31 + members)
    lens = np.full(n, length, np.int32)
    lens[5], lens[40] = 1500, 2999                          # shorter instances of the template: some members hold none of 5's tail
    include = np.ones(n, np.uint8)
    include[11] = 0
    c = api.Context(multi=[g % ngpu for g in range(members)])
    try:
        c.mat_set_problem(n, length)
        for i in range(n):
            c.mat_put_sample(i, counts[i, :lens[i]], totals[i, :lens[i]])
        for kw in (dict(norm=0), dict(norm=1000, min_depth=20, min_cov=0.4)):
            D, N, dn, rows = c.mat_run(include, method=method, **kw)
            cz = counts.copy()
            for i in range(n):
                cz[i, lens[i]:] = 0
            Do, No, dno = oracle.mat_matrix(cz, cz.astype(np.uint32).sum(axis=2).astype(np.uint32), lens, include, method=method, **kw)
            assert dn == dno == n - 1
            assert np.array_equal(N, No) and np.array_equal(rows, No.astype(np.uint32))
            assert close(D, Do)
        assert "position split" in c.last_kernel and c.multi_gpus()[0] == members
    finally:
        c.close()


def test_partial_sums_and_host_epilogue(ctx):
    # what a rank of a one-process-per-GPU position split does: raw sums from the device, the tail of cmpMats on the host
    import ctypes as C
    n, length = 33, 1200
    counts, totals = random_counts(n, length, seed=3, low=0.2)
    lens = np.full(n, length, np.int32)
    include = np.ones(n, np.uint8)
    include[0] = 0
    ctx.mat_set_problem(n, length)
    for i in range(n):
        ctx.mat_put_sample(i, counts[i], totals[i])
    L = api.load()
    cells = api.cells(n - 1)
    for elem, scale in ((8, 1.0), (4, 1.0), (2, 100.0)):
        D, N, dn, rows = ctx.mat_run(include, method="cos", norm=1000, min_cov=0.6, elem_size=elem, byte_scale=scale)
        dist = np.zeros(cells, np.float64)
        rr = np.zeros(cells, np.uint32)
        dnp = C.c_int(0)
        mid, order = api.mat_method("cos")
        assert L.ccg_mat_run_partial(ctx._h, include.ctypes.data, mid, order, 0.05, 15, dist.ctypes.data, rr.ctypes.data, C.byref(dnp)) == 0
        D2 = np.zeros(cells, api.ELEM_DTYPE[elem])
        N2 = np.zeros(cells, api.ELEM_DTYPE[elem])
        r2 = np.zeros(cells, np.uint32)
        assert L.ccg_mat_finalize_host(n, include.ctypes.data, lens.ctypes.data, dist.ctypes.data, rr.ctypes.data, 1000, 1, 0.6, elem, scale,
                                       D2.ctypes.data, N2.ctypes.data, r2.ctypes.data, None) == 0
        assert dnp.value == dn and np.array_equal(D2.view(np.uint8), D.view(np.uint8)) and np.array_equal(N2, N) and np.array_equal(r2, rows)
