"""GPU parity: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs and against the reference's golden vectors.  Integer counts and every cell
type must be BIT-EXACT (the epilogue is IEEE round-to-nearest without FMA contraction, so
floating-point cells are bit-exact too; the 1e-6 relative tolerance of the north star is
asserted as well, explicitly, where doubles are involved)."""
import numpy as np
import pytest

import helpers
import oracle
from ccphylo_b200 import api, synth
import synth_torch  # noqa: E402

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6            # north-star tolerance for floating-point distances


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def _set(n, length, seed, **kw):
    kw.setdefault("snp", 0.02)
    kw.setdefault("nrun", 0.05)
    codes = synth.make_codes(n, length, seed=seed, **kw)
    seqs, masks, inc = oracle.encode_samples(codes)
    return codes, seqs, masks, inc


@pytest.fixture(scope="module")
def ctx(built):
    c = api.Context()
    yield c
    c.close()


POOL, CASES = helpers.golden_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_text_through_the_abi(built, case):
    """Byte-identical .phy / .num / stderr text to the reference binary's."""
    phy, num, err = helpers.replay(case, POOL, helpers.gpu_backend)
    assert err == case["stderr"]
    assert phy == case["phy"]
    assert num == case["num"]


@pytest.mark.parametrize("n,length", [(2, 1), (3, 31), (5, 32), (64, 33), (65, 127), (70, 128), (129, 129),
                                      (200, 4099), (257, 16384 + 17)])
def test_raw_counts_bit_exact(ctx, n, length):
    codes, seqs, masks, inc = _set(n, length, seed=n * 31 + length)
    ctx.set_problem(n, length, pair=True)
    ctx.put_samples_packed(seqs, masks)
    D, N, dn = ctx.run_pair(min_length=0, min_cov=0.0)
    assert dn == n
    mism, ninc = ctx.raw_counts(dn)
    mo, no = oracle.raw_pair_matrix(seqs, masks, length)
    assert np.array_equal(mism, mo)
    assert np.array_equal(ninc, no)
    assert np.array_equal(N, no.astype(np.float64))
    assert np.array_equal(ctx.inc_counts(), inc.astype(np.uint32))
    assert "k_pairdist" in ctx.last_kernel and ctx.launches > 0


@pytest.mark.parametrize("elem,scale", [(8, 1.0), (4, 1.0), (2, 10.0), (2, 100.0), (1, 0.01), (1, 1.0)])
@pytest.mark.parametrize("norm", [0, 1000, 1000000])
def test_pair_epilogue_all_cell_types(ctx, elem, scale, norm):
    n, length = 90, 3001
    codes, seqs, masks, inc = _set(n, length, seed=elem + norm % 97)
    codes[7, :] = 4
    codes[40, : length - 100] = 4
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
    if elem == 8:
        np.testing.assert_allclose(D, Do, rtol=REL_TOL, atol=0.0)


@pytest.mark.parametrize("elem", [8, 4])
def test_zero_over_zero_cells_print_like_the_reference(ctx, elem):
    # -L 0 -C 0 with a pair that shares no included position and a normalisation weight: the reference divides 0 by 0.
    # Its NaN has the sign bit set in both cell types ("-nan" in the Phylip text); the epilogue writes the same bits.
    n, length = 6, 700
    codes = synth.make_codes(n, length, seed=31, snp=0.02, nrun=0.0)
    codes[1, :400] = 4
    codes[2, 400:] = 4
    seqs, masks, inc = oracle.encode_samples(codes)
    include = np.ones(n, dtype=np.uint8)
    D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=1000, min_length=0, min_cov=0.0,
                                         elem_size=elem, ctx=ctx)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=1000, min_length=0, min_cov=0.0, elem_size=elem)
    assert dn == dno == n
    assert np.isnan(Do).sum() == 1 and np.signbit(Do[np.isnan(Do)]).all()
    assert np.array_equal(_bits(D), _bits(Do))
    assert np.array_equal(_bits(N), _bits(No))


def test_pair_gate_writes_minus_one(ctx):
    n, length = 20, 2000
    codes = synth.make_codes(n, length, seed=5, snp=0.02, nrun=0.0)
    codes[1, :1200] = 4
    codes[2, 900:] = 4
    seqs, masks, inc = oracle.encode_samples(codes)
    include = np.ones(n, dtype=np.uint8)
    for elem, scale in ((8, 1.0), (4, 1.0), (2, 10.0), (1, 0.1)):
        D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=100, min_length=500,
                                             min_cov=0.0, elem_size=elem, byte_scale=scale, ctx=ctx)
        Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=100, min_length=500, min_cov=0.0,
                                          elem_size=elem, byte_scale=scale)
        assert dn == dno == n
        assert np.array_equal(_bits(D), _bits(Do))
        assert np.array_equal(_bits(N), _bits(No))
    D8, _, _, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=100, min_length=500,
                                         min_cov=0.0, ctx=ctx)
    assert (D8 == -1.0).sum() >= 1


@pytest.mark.parametrize("elem,scale", [(8, 1.0), (4, 1.0), (2, 10.0), (1, 0.5)])
@pytest.mark.parametrize("norm", [0, 1000])
def test_global_mode(ctx, elem, scale, norm):
    n, length = 75, 5003
    codes, seqs, masks, inc = _set(n, length, seed=11 + elem, nrun=0.004)
    include = np.ones(n, dtype=np.uint8)
    include[[3, 64]] = 0                      # intended semantics: excluded samples are skipped
    gmask = oracle.global_mask(codes, include)
    D, _, dn, ginc = api.fsa_cmp_thread_out(seqs, include, gmask.reshape(1, -1), length, pair=False, norm=norm,
                                            elem_size=elem, byte_scale=scale, ctx=ctx)
    Do, dno, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=norm, elem_size=elem, byte_scale=scale)
    assert dn == dno == n - 2 and ginc == ginco and ginc > length // 4
    assert np.array_equal(_bits(D), _bits(Do))
    assert float(D.max()) > 0


@pytest.mark.parametrize("kernel", [api.KERNEL_POPC, api.KERNEL_UMMA], ids=["popc", "umma"])
def test_global_mask_applied_to_a_pair_store(built, kernel):
    """ccg_apply_global_mask: samples streamed into a pair-mode store first (the dist driver's
    load loop, cdist.c:55-168), the shared mask ANDed in afterwards, then cmpFsaThrd semantics."""
    n, length = 140, 4000 + 21
    codes, seqs, masks, inc = _set(n, length, seed=77, nrun=0.003)
    include = np.ones(n, dtype=np.uint8)
    gmask = oracle.global_mask(codes, include)
    with api.Context() as c:
        c.set_kernel(kernel)
        c.set_problem(n, length, pair=True)
        c.put_samples_packed(seqs, masks)
        c.apply_global_mask(gmask)
        D, dn, ginc = c.run_global(include, norm=1000)
        # and with the mask built on the device from the uploaded samples' own masks
        c.set_problem(n, length, pair=True)
        c.put_samples_packed(seqs, masks)
        assert c.build_global_mask(include) == int(sum(bin(int(w)).count("1") for w in gmask))
        D2, dn2, ginc2 = c.run_global(include, norm=1000)
        assert dn2 == dn and ginc2 == ginc and np.array_equal(_bits(D2), _bits(D))
    Do, dno, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=1000)
    assert dn == dno == n and ginc == ginco
    assert np.array_equal(_bits(D), _bits(Do))
    assert float(D.max()) > 0


def test_codes_upload_matches_packed_upload(ctx):
    """Device-side qseq2nibble + initIncPos + getIncPos + getNpos (ccg_put_sample_codes)."""
    n, length = 66, 1000 + 29
    codes, seqs, masks, inc = _set(n, length, seed=77)
    ctx.set_problem(n, length, pair=True)
    for i in range(n):
        ctx.put_sample_codes(i, codes[i])
    assert np.array_equal(ctx.inc_counts(), inc.astype(np.uint32))
    D, N, dn = ctx.run_pair(norm=1000)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, np.ones(n, np.uint8), length, norm=1000)
    assert dn == dno
    assert np.array_equal(_bits(D), _bits(Do)) and np.array_equal(_bits(N), _bits(No))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_tile_partition_ranks_sum_to_whole(ctx, world):
    """Each rank fills only its own lower-triangular tile blocks; the union is the full matrix."""
    n, length = 300, 2048 + 5
    codes, seqs, masks, inc = _set(n, length, seed=world)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, np.ones(n, np.uint8), length, norm=1000, min_length=0, min_cov=0.0)
    ctx.set_problem(n, length, pair=True)
    ctx.put_samples_packed(seqs, masks)
    total_D = np.zeros(api.cells(n))
    total_N = np.zeros(api.cells(n))
    owned = 0
    try:
        for r in range(world):
            ctx.set_partition(r, world)
            D, N, dn = ctx.run_pair(norm=1000, min_length=0, min_cov=0.0)
            assert dn == n
            assert np.count_nonzero(N) == api.partition_cells(n, r, world)
            assert not np.any((total_N != 0) & (N != 0)), "two ranks wrote the same cell"
            total_D += D
            total_N += N
            owned += np.count_nonzero(N)
    finally:
        ctx.set_partition(0, 1)
    assert owned == api.cells(n)
    assert np.array_equal(total_N, No) and np.array_equal(total_D, Do)


def test_ksplit_and_single_slice_agree(ctx):
    """Long sequences are K-split across CTAs (integer RED.ADD + ticket); must equal the oracle."""
    n, length = 64, 400000 + 3
    codes, seqs, masks, inc = _set(n, length, seed=4242, snp=0.001, nrun=0.01)
    ctx.set_problem(n, length, pair=True)
    ctx.put_samples_packed(seqs, masks)
    D, N, dn = ctx.run_pair(norm=0, min_length=0, min_cov=0.0)
    assert "ksplit=" in ctx.last_kernel and not ctx.last_kernel.endswith("ksplit=1")
    mism, ninc = ctx.raw_counts(dn)
    mo, no = oracle.raw_pair_matrix(seqs, masks, length)
    assert np.array_equal(mism, mo) and np.array_equal(ninc, no)


def test_size_independent_properties_large(ctx):
    """A size the oracle cannot enumerate quickly: check algebraic properties instead --
    (1) counts over [0, L) = counts over [0, L/2) + counts over [L/2, L)   (linearity in K)
    (2) permuting the samples permutes the matrix
    (3) N[i][j] <= min(inc_i, inc_j) and mism <= N
    (4) a sample of spot-checked cells equals the oracle."""
    import torch

    n, length = 512, 1_000_000
    half = (length // 2 // 32) * 32
    seqs_t, masks_t = synth_torch.make_packed_torch(n, length, seed=9, device="cuda")
    W = seqs_t.shape[1]

    def run(s_t, m_t, L):
        ctx.set_problem(n, L, pair=True)
        ctx.put_samples_packed_dev(s_t.data_ptr(), m_t.data_ptr(), n, s_t.stride(0))
        D, N, dn = ctx.run_pair(min_length=0, min_cov=0.0)
        return ctx.raw_counts(dn)

    torch.cuda.synchronize()
    m_all, n_all = run(seqs_t, masks_t, length)
    wa = half // 32
    lo_s, lo_m = seqs_t[:, :wa].contiguous(), masks_t[:, :wa].contiguous()
    hi_s, hi_m = seqs_t[:, wa:].contiguous(), masks_t[:, wa:].contiguous()
    m_lo, n_lo = run(lo_s, lo_m, half)
    m_hi, n_hi = run(hi_s, hi_m, length - half)
    assert np.array_equal(m_all, m_lo + m_hi) and np.array_equal(n_all, n_lo + n_hi)

    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    m_p, n_p = run(seqs_t[perm].contiguous(), masks_t[perm].contiguous(), length)
    full = helpers.full_from_packed(m_all, n)
    full_p = helpers.full_from_packed(m_p, n)
    p = perm.cpu().numpy()
    assert np.array_equal(full_p, full[np.ix_(p, p)])

    inc = ctx.inc_counts()          # of the permuted upload
    Nfull = helpers.full_from_packed(n_p, n)
    assert (Nfull <= np.minimum.outer(inc, inc) + np.eye(n, dtype=np.uint32) * 0).all()
    assert (full_p <= Nfull).all()

    seqs = seqs_t.cpu().numpy().view(np.uint64)
    masks = masks_t.cpu().numpy().view(np.uint32)
    rng = np.random.default_rng(0)
    for _ in range(40):
        i, j = sorted(rng.choice(n, size=2, replace=False))[::-1]
        mo, no = oracle.raw_pair_matrix(np.stack([seqs[j], seqs[i]]), np.stack([masks[j], masks[i]]), length, 1)
        assert full[i, j] == mo[0] and helpers.full_from_packed(n_all, n)[i, j] == no[0]


def test_empty_and_degenerate_inputs(ctx):
    # a single included sample: Dn = 1, nothing to compare
    codes, seqs, masks, inc = _set(3, 100, seed=1)
    include = np.array([0, 1, 0], dtype=np.uint8)
    D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, 100, pair=True, ctx=ctx)
    assert dn == 1 and len(D) == 0
    # nothing included
    D, N, dn, _ = api.fsa_cmp_thread_out(seqs, np.zeros(3, np.uint8), masks, 100, pair=True, ctx=ctx)
    assert dn == 0 and len(D) == 0
    # a call order the device cannot serve is a loud refusal, never a silent fallback: variant lists (-V) under
    # proximity masking (-P) compare the whole packed words, which a packed upload made BEFORE -P was set has already
    # ANDed with the rows' masks
    ctx.set_problem(3, 100, pair=True)
    ctx.put_samples_packed(seqs, masks)
    ctx.set_proximity(3)
    try:
        with pytest.raises(api.CcgError) as e:
            ctx.list_variants(pair=True)
        assert e.value.code == 3 and "ccg_set_proximity before" in str(e.value)
    finally:
        ctx.set_proximity(0)


def test_temporary_context_path(built):
    """ctx = NULL: the one-call drop-in creates and destroys its own context."""
    codes, seqs, masks, inc = _set(10, 500, seed=2)
    include = np.ones(10, dtype=np.uint8)
    D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, 500, pair=True, norm=10)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, 500, norm=10)
    assert dn == dno and np.array_equal(D, Do) and np.array_equal(N, No)
