"""GPU parity of the tensor-core path (tcgen05 int8 contraction, CCG_KERNEL_UMMA): same
bit-exact bar as the LOP3+POPC path -- integer counts, every cell type, shared-mask mode,
the rank partition, and multi-slab K processing -- against the CPU oracle and against the
POPC kernel on the device."""
import numpy as np
import pytest

import helpers
import oracle
from ccphylo_b200 import api, synth
import synth_torch  # noqa: E402

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def _set(n, length, seed, **kw):
    kw.setdefault("snp", 0.02)
    kw.setdefault("nrun", 0.05)
    codes = synth.make_codes(n, length, seed=seed, **kw)
    seqs, masks, inc = oracle.encode_samples(codes)
    return codes, seqs, masks, inc


KNAME = {api.KERNEL_UMMA: "k_pairdist_umma", api.KERNEL_FUSED: "k_pairdist_fused"}


# the tensor path in its three forms: e2m1 operands on kind::mxf4 (the default), int8 operands on kind::i8
# (CCG_I8=1, read when the context is created), and the fused int8 kernel
@pytest.fixture(scope="module", params=[(api.KERNEL_UMMA, "0"), (api.KERNEL_UMMA, "1"), (api.KERNEL_FUSED, "0")],
                ids=["umma-mxf4", "umma-i8", "fused"])
def ctx(built, request):
    import os
    kind, i8 = request.param
    os.environ["CCG_I8"] = i8
    try:
        c = api.Context()
    finally:
        del os.environ["CCG_I8"]
    c.set_kernel(kind)
    c.kind = kind
    c.bytes_per_chunk_slot = 512 if (i8 == "1" or kind == api.KERNEL_FUSED) else 256
    yield c
    c.close()


@pytest.mark.parametrize("n,length", [(2, 1), (3, 31), (5, 128), (64, 129), (129, 1000), (257, 4099), (300, 20000 + 17),
                                      (513, 3000)])
def test_umma_raw_counts_bit_exact(ctx, n, length):
    codes, seqs, masks, inc = _set(n, length, seed=n * 31 + length)
    ctx.set_problem(n, length, pair=True)
    ctx.put_samples_packed(seqs, masks)
    D, N, dn = ctx.run_pair(min_length=0, min_cov=0.0)
    assert dn == n
    assert KNAME[ctx.kind] in ctx.last_kernel
    mism, ninc = ctx.raw_counts(dn)
    mo, no = oracle.raw_pair_matrix(seqs, masks, length)
    assert np.array_equal(ninc, no)
    assert np.array_equal(mism, mo)
    assert np.array_equal(N, no.astype(np.float64))
    assert np.array_equal(D, mo.astype(np.float64))


@pytest.mark.parametrize("elem,scale", [(8, 1.0), (4, 1.0), (2, 10.0), (1, 0.01)])
@pytest.mark.parametrize("norm", [0, 1000000])
def test_umma_pair_epilogue_with_exclusions(ctx, elem, scale, norm):
    n, length = 140, 3001
    codes, seqs, masks, inc = _set(n, length, seed=elem + norm % 97)
    codes[7, :] = 4
    codes[130, : length - 100] = 4
    seqs, masks, inc = oracle.encode_samples(codes)
    min_len = int(0.5 * length)
    include = (inc >= min_len).astype(np.uint8)
    D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=norm, min_length=min_len,
                                         min_cov=0.5, elem_size=elem, byte_scale=scale, ctx=ctx)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=norm, min_length=min_len, min_cov=0.5,
                                      elem_size=elem, byte_scale=scale)
    assert dn == dno == n - 2
    assert np.array_equal(_bits(D), _bits(Do))
    assert np.array_equal(_bits(N), _bits(No))


@pytest.mark.parametrize("norm", [0, 1000])
def test_umma_global_mode(ctx, norm):
    n, length = 150, 5003
    codes, seqs, masks, inc = _set(n, length, seed=23, nrun=0.002)
    include = np.ones(n, dtype=np.uint8)
    include[[3, 129]] = 0
    gmask = oracle.global_mask(codes, include)
    D, _, dn, ginc = api.fsa_cmp_thread_out(seqs, include, gmask.reshape(1, -1), length, pair=False, norm=norm, ctx=ctx)
    Do, dno, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=norm)
    assert dn == dno == n - 2 and ginc == ginco and ginc > length // 4
    assert np.array_equal(_bits(D), _bits(Do))
    assert float(D.max()) > 0


POOL, CASES = helpers.golden_cases()
SOME = [c for c in CASES if c["name"].startswith(("c1_", "c6_", "rand_L129", "rand_L777"))]


@pytest.mark.parametrize("case", SOME, ids=[c["name"] for c in SOME])
def test_umma_golden_text(ctx, case):
    def backend(prep):
        o = prep["opts"]
        return api.fsa_cmp_thread_out(prep["seqs"], prep["include"], prep["masks"] if prep["pair"] else prep["gmask"],
                                      prep["L"], pair=prep["pair"], norm=o["norm"], min_length=prep["min_len"],
                                      min_cov=o["min_cov"], elem_size=o["elem"], byte_scale=o["scale"], ctx=ctx)

    phy, num, err = helpers.replay(case, POOL, backend)
    assert err == case["stderr"] and phy == case["phy"] and num == case["num"]


@pytest.mark.parametrize("world", [2, 3])
def test_umma_partition(ctx, world):
    n, length = 600, 2048 + 5
    codes, seqs, masks, inc = _set(n, length, seed=world)
    mo, no = oracle.raw_pair_matrix(seqs, masks, length)
    ctx.set_problem(n, length, pair=True)
    ctx.put_samples_packed(seqs, masks)
    total_D = np.zeros(api.cells(n))
    total_N = np.zeros(api.cells(n))
    try:
        for r in range(world):
            ctx.set_partition(r, world)
            D, N, dn = ctx.run_pair(min_length=0, min_cov=0.0)
            assert np.count_nonzero(N) == api.partition_cells(n, r, world)
            assert not np.any((total_N != 0) & (N != 0))
            total_D += D
            total_N += N
    finally:
        ctx.set_partition(0, 1)
    assert np.array_equal(total_N, no.astype(np.float64)) and np.array_equal(total_D, mo.astype(np.float64))


def test_umma_multi_slab_and_kslices(ctx):
    """Tiny scratch budget forces several K slabs; long K forces several K slices per tile."""
    n, length = 200, 300000 + 77
    codes, seqs, masks, inc = _set(n, length, seed=4242, snp=0.001, nrun=0.01)
    ctx.set_problem(n, length, pair=True)
    ctx.put_samples_packed(seqs, masks)
    mo, no = oracle.raw_pair_matrix(seqs, masks, length)
    if ctx.kind == api.KERNEL_UMMA:
        try:
            ctx.set_scratch_limit(2 * 256 * ctx.bytes_per_chunk_slot * 350)      # two 350-chunk slab buffers -> 7 slabs
            D, N, dn = ctx.run_pair(min_length=0, min_cov=0.0)
            assert "slabs=7" in ctx.last_kernel, ctx.last_kernel
            mism, ninc = ctx.raw_counts(dn)
            assert np.array_equal(mism, mo) and np.array_equal(ninc, no)
        finally:
            ctx.set_scratch_limit(0)
    D, N, dn = ctx.run_pair(min_length=0, min_cov=0.0)
    assert "kslices=1 " not in ctx.last_kernel + " "
    mism, ninc = ctx.raw_counts(dn)
    assert np.array_equal(mism, mo) and np.array_equal(ninc, no)


def test_umma_equals_popc_on_device(built):
    import torch

    n, length = 700, 400_000
    seqs_t, masks_t = synth_torch.make_packed_torch(n, length, seed=5, device="cuda")
    torch.cuda.synchronize()
    out = {}
    for kind in (api.KERNEL_POPC, api.KERNEL_UMMA, api.KERNEL_FUSED):
        with api.Context() as c:
            c.set_kernel(kind)
            c.set_problem(n, length, pair=True)
            c.put_samples_packed_dev(seqs_t.data_ptr(), masks_t.data_ptr(), n, seqs_t.stride(0))
            D, N, dn = c.run_pair(norm=1000000)
            out[kind] = (D, N, c.raw_counts(dn))
    a = out[api.KERNEL_POPC]
    for kind in (api.KERNEL_UMMA, api.KERNEL_FUSED):
        b = out[kind]
        assert np.array_equal(a[2][0], b[2][0]) and np.array_equal(a[2][1], b[2][1])
        assert np.array_equal(_bits(a[0]), _bits(b[0])) and np.array_equal(_bits(a[1]), _bits(b[1]))
    assert a[2][0].max() > 0


@pytest.mark.parametrize("kind", [api.KERNEL_POPC, api.KERNEL_UMMA], ids=["popc", "umma"])
def test_partition_set_before_upload_only_touches_needed_rows(built, kind):
    """One process per GPU sets its partition FIRST: rows that none of its macro tiles reads are
    registered (they still count in the compaction) but never uploaded / expanded."""
    n, length, world = 900, 4096 + 31, 4
    codes, seqs, masks, inc = _set(n, length, seed=77)
    include = np.ones(n, dtype=np.uint8)
    include[[5, 400]] = 0
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=1000, min_length=0, min_cov=0.0)
    total_D = np.zeros(api.cells(dno))
    total_N = np.zeros(api.cells(dno))
    for r in range(world):
        with api.Context() as c:
            c.set_kernel(kind)
            c.set_partition(r, world)
            c.set_problem(n, length, pair=True)
            c.put_samples_packed(seqs, masks)
            D, N, dn = c.run_pair(include=include, norm=1000, min_length=0, min_cov=0.0)
            assert dn == dno
            assert not np.any((total_N != 0) & (N != 0))
            total_D += D
            total_N += N
    assert np.array_equal(total_N, No) and np.array_equal(total_D, Do)


def test_partition_change_after_upload_is_refused(built):
    n, length = 900, 2048
    codes, seqs, masks, inc = _set(n, length, seed=78)
    with api.Context() as c:
        c.set_partition(3, 4)
        c.set_problem(n, length, pair=True)
        c.put_samples_packed(seqs, masks)
        c.set_partition(0, 1)
        with pytest.raises(api.CcgError) as e:
            c.run_pair()
        assert e.value.code == 3 and "ccg_set_partition before" in str(e.value)


@pytest.mark.parametrize("pair", [True, False], ids=["pair", "shared-mask"])
@pytest.mark.parametrize("contiguous", [True, False], ids=["2d-copies", "row-copies"])
def test_streamed_host_rows_match_resident_upload(built, monkeypatch, pair, contiguous):
    """ccg_fsa_cmp_thread_out on the tensor path streams the host rows K slab by K slab (upload of
    slab s+1 under the GEMM of slab s); results must equal the oracle bit for bit, for rows in one
    allocation (2-D copies) and for separately allocated rows (row-by-row copies)."""
    monkeypatch.setenv("CCG_STREAM_MIN_CHUNKS", "32")
    n, length = 260, 128 * 150 + 77
    codes, seqs, masks, inc = _set(n, length, seed=5, nrun=0.004 if not pair else 0.05)
    include = np.ones(n, dtype=np.uint8)
    include[[7, 200]] = 0
    rows = None
    if not contiguous:
        # every row in its own buffer, so the row pointers are at irregular distances
        rs = [np.array(seqs[k], dtype=np.uint64, copy=True) for k in range(n)]
        rm = [np.array(masks[k], dtype=np.uint32, copy=True) for k in range(n)]
        rows = (rs, rm)
    with api.Context() as c:
        c.set_kernel(api.KERNEL_UMMA)
        if pair:
            D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=1000, ctx=c, rows=rows)
            Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=1000)
            assert dn == dno == n - 2
            assert np.array_equal(_bits(N), _bits(No))
            assert np.array_equal(_bits(D), _bits(Do))
            assert np.array_equal(c.inc_counts()[include == 1], inc[include == 1])
        else:
            gmask = oracle.global_mask(codes, include)
            D, _, dn, ginc = api.fsa_cmp_thread_out(seqs, include, gmask.reshape(1, -1), length, pair=False, norm=1000,
                                                    ctx=c, rows=rows)
            Do, dno, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=1000)
            assert dn == dno == n - 2 and ginc == ginco
            assert np.array_equal(_bits(D), _bits(Do))
        assert "slabs=" in c.last_kernel and int(c.last_kernel.split("slabs=")[1]) > 1


# ---- the f32 accumulators of kind::mxf4 at their exactness bound ----
# An S work item adds 3 per matching base: identical, fully known samples drive every accumulator element to
# 3 x (bases of the K slice), the largest magnitude the path can produce.  The host keeps a slice at or below
# CCG_FP4_MAX_PAIRS = 20,000 chunk pairs (5.12 Mbp): 3 * 256 * 20000 = 15,360,000 < 2^24.  The guard must hold
# when the experiment knob CCG_KSLICES asks for fewer slices, and on the automatic path for longer genomes.
@pytest.mark.parametrize("length,kslices_env,min_slices", [(5_120_000, "1", 1), (5_120_000 + 256, "1", 2), (12_000_000, None, 3)],
                         ids=["5.12Mbp-one-slice-at-the-bound", "one-pair-over-forces-two-slices", "12Mbp-auto"])
def test_mxf4_exact_at_the_accumulator_bound(built, length, kslices_env, min_slices):
    import os
    n = 40
    rng = np.random.default_rng(length % 1009)
    row = rng.integers(0, 4, size=length, dtype=np.uint8)
    codes = np.broadcast_to(row, (n, length)).copy()
    codes[1, ::7] = (codes[1, ::7] + 1) & 3            # one sample that differs in every 7th base
    codes[2, 1000:2000] = 4                            # one with unknown bases
    seqs, masks, inc = oracle.encode_samples(codes)
    del codes
    if kslices_env is not None:
        os.environ["CCG_KSLICES"] = kslices_env
    try:
        c = api.Context()
    finally:
        os.environ.pop("CCG_KSLICES", None)
    try:
        c.set_kernel(api.KERNEL_UMMA)
        c.set_problem(n, length, pair=True)
        c.put_samples_packed(seqs, masks)
        D, N, dn = c.run_pair(min_length=0, min_cov=0.0)
        kern = c.last_kernel
        assert "mxf4" in kern
        ks = int(kern.split("kslices=")[1].split()[0])
        slabs = int(kern.split("slabs=")[1].split()[0])
        assert ks * slabs >= min_slices, kern
        pairs_per_item = -(-(-(-length // 256)) // (ks * slabs))
        assert pairs_per_item <= 20000, kern
    finally:
        c.close()
    mo, no = oracle.raw_pair_matrix(seqs, masks, length)
    assert dn == n
    assert np.array_equal(N, no.astype(np.float64))
    assert np.array_equal(D, mo.astype(np.float64))
    assert (mo == 0).sum() >= (n - 2) * (n - 3) // 2 and no.max() == length


# ---- rows LENT to the library: the tensor path expands straight from the reference's packed words ----
@pytest.mark.parametrize("n,length", [(300, 128 * 70 + 5), (513, 20000 + 17), (260, 255)])
@pytest.mark.parametrize("pair", [True, False], ids=["pair", "shared-mask"])
def test_lent_rows_expand_without_the_plane_store(built, n, length, pair):
    import torch

    codes, seqs, masks, inc = _set(n, length, seed=n + length, nrun=0.004 if not pair else 0.05)
    include = np.ones(n, np.uint8)
    include[[3, n - 1]] = 0
    gmask = oracle.global_mask(codes, include)
    d_seqs = torch.from_numpy(seqs.view(np.int64)).cuda()
    d_masks = torch.from_numpy(masks.view(np.int32)).cuda()
    torch.cuda.synchronize()
    with api.Context() as c:
        c.set_kernel(api.KERNEL_UMMA)
        c.set_problem(n, length, pair=pair)
        if not pair:
            c.put_global_mask(gmask)
        c.put_samples_packed_dev_borrowed(d_seqs.data_ptr(), d_masks.data_ptr() if pair else None, n, d_seqs.stride(0))
        launches0 = c.launches
        if pair:
            D, N, dn = c.run_pair(include=include, norm=1000)
            Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=1000)
            assert np.array_equal(_bits(N), _bits(No))
        else:
            D, dn, ginc = c.run_global(include=include, norm=1000)
            Do, dno, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=1000)
            assert ginc == ginco
        assert dn == dno == n - 2 and np.array_equal(_bits(D), _bits(Do))
        assert "mxf4" in c.last_kernel
        # no k_repack launch: expansion(s) + GEMM + finalize only
        assert c.launches - launches0 <= 4
        if pair:
            # something that needs the planes builds them from the lent rows: per-sample counts, then the POPC kernel
            assert np.array_equal(c.inc_counts(), inc)
            c.set_kernel(api.KERNEL_POPC)
            D2, N2, dn2 = c.run_pair(include=include, norm=1000)
            assert np.array_equal(_bits(D2), _bits(Do)) and np.array_equal(_bits(N2), _bits(No))
            # and a fresh loan after that is read directly again
            c.set_kernel(api.KERNEL_UMMA)
            c.put_samples_packed_dev_borrowed(d_seqs.data_ptr(), d_masks.data_ptr(), n, d_seqs.stride(0))
            D3, N3, dn3 = c.run_pair(include=include, norm=1000)
            assert np.array_equal(_bits(D3), _bits(Do)) and np.array_equal(_bits(N3), _bits(No))


# ---- every extent of the last macro-tile row: thin (transposed, MMA N = 16 .. 96) up to 96 valid rows, full tiles above ----
@pytest.mark.parametrize("v", [15, 16, 17, 31, 32, 33, 47, 48, 49, 64, 80, 81, 95, 96, 97, 128, 200, 255])
def test_last_tile_row_extents(built, v):
    n = (512 if v % 2 else 256) + v
    length = 128 * 9 + 77
    codes, seqs, masks, inc = _set(n, length, seed=1000 + v, snp=0.05, nrun=0.08)
    include = np.ones(n, np.uint8)
    include[[1, n - 2]] = 0
    with api.Context() as c:
        c.set_kernel(api.KERNEL_UMMA)
        c.set_problem(n, length, pair=True)
        c.put_samples_packed(seqs, masks)
        D, N, dn = c.run_pair(min_length=0, min_cov=0.0)
        assert "mxf4" in c.last_kernel and dn == n
        mism, ninc = c.raw_counts(dn)
        mo, no = oracle.raw_pair_matrix(seqs, masks, length)
        assert np.array_equal(ninc, no) and np.array_equal(mism, mo)
        D, N, dn = c.run_pair(include=include, norm=1000000)
        Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=1000000)
        assert dn == dno == n - 2
        assert np.array_equal(_bits(D), _bits(Do)) and np.array_equal(_bits(N), _bits(No))
