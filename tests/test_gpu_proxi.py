"""GPU parity of -P proximity masking: k_pairdist_proxi against maskProxi + fsacmpair
(fsacmp.c:355-485,587-633), k_sample_proxi against getIncPos / getIncPosInsig / getIncPosInsigPrune with proxi > 0
(fsacmp.c:181-353), both through the C-ABI and against the oracle, which tests/test_oracle_vs_reference.py pins to
the reference's own functions.  Counts and cells are BIT-EXACT."""
import numpy as np
import pytest

import oracle
from ccphylo_b200 import api, synth

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def _codes(n, length, seed, snp=0.03):
    codes = synth.make_codes(n, length, seed=seed, snp=snp, nrun=0.03, lower=0.02, gap=0.01)
    return codes


@pytest.fixture(scope="module")
def ctx(built):
    c = api.Context()
    yield c
    c.set_proximity(0)
    c.close()


@pytest.mark.parametrize("proxi", [1, 2, 7, 31, 32, 33, 100, 1000, 50000])
@pytest.mark.parametrize("n,length", [(2, 2), (5, 33), (64, 128), (65, 129), (70, 1000), (130, 4099), (200, 16384 + 17)])
def test_pair_counts_under_proximity_mask(ctx, n, length, proxi):
    codes = _codes(n, length, seed=n + length + proxi)
    seqs, masks, inc = oracle.encode_samples(codes, proxi=proxi)
    include = np.ones(n, dtype=np.uint8)
    D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=0, min_length=0, min_cov=0.0,
                                         proxi=proxi, ctx=ctx)
    assert dn == n and "k_pairdist_proxi" in ctx.last_kernel
    mism, ninc = ctx.raw_counts(dn)
    k = 0
    differs = False
    for i in range(1, n):
        for j in range(i):
            if n <= 70 or (i * 131 + j) % 17 == 0:
                want = oracle.pair_counts_proxi(seqs[i], seqs[j], masks[i], masks[j], length, proxi)
                assert (int(mism[k]), int(ninc[k])) == want, (i, j)
                differs |= want != oracle.pair_counts_proxi(seqs[i], seqs[j], masks[i], masks[j], length, 0)
            k += 1
    assert differs or length < 128
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=0, min_length=0, min_cov=0.0, proxi=proxi)
    assert np.array_equal(_bits(D), _bits(Do)) and np.array_equal(_bits(N), _bits(No))


@pytest.mark.parametrize("elem,scale", [(8, 1.0), (4, 1.0), (2, 10.0), (1, 0.5)])
@pytest.mark.parametrize("proxi", [3, 40])
def test_pair_epilogue_with_proximity(ctx, elem, scale, proxi):
    n, length = 90, 3001
    codes = _codes(n, length, seed=elem * 5 + proxi)
    codes[7, :] = 4
    codes[40, : length - 100] = 4
    seqs, masks, inc = oracle.encode_samples(codes, proxi=proxi)
    min_len = int(0.5 * length)
    include = (inc >= min_len).astype(np.uint8)
    D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=1000, min_length=min_len,
                                         min_cov=0.5, proxi=proxi, elem_size=elem, byte_scale=scale, ctx=ctx)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=1000, min_length=min_len, min_cov=0.5,
                                      elem_size=elem, byte_scale=scale, proxi=proxi)
    assert dn == dno and 2 <= n - dn
    assert np.array_equal(_bits(D), _bits(Do))
    assert np.array_equal(_bits(N), _bits(No))


@pytest.mark.parametrize("snp_only", [False, True])
@pytest.mark.parametrize("proxi", [1, 5, 32, 64, 300, 40000])
def test_per_sample_builder_on_device(ctx, proxi, snp_only):
    """codes in, device-side getIncPos(seq, seq, proxi) (cdist.c:91), counts and the pair run that follows"""
    n, length = 37, 40000 + 13
    variant = 1 if snp_only else 0
    codes = synth.make_codes(n, length, seed=proxi, snp=0.03, nrun=0.004, lower=0.002, gap=0.001)
    codes[5, 100:39000] = 4
    seqs, masks, inc = oracle.encode_samples(codes, proxi=proxi, variant=variant)
    ctx.set_proximity(proxi, snp_only)
    try:
        ctx.set_problem(n, length, pair=True)
        for i in range(n):
            ctx.put_sample_codes(i, codes[i])
            ctx.sync()
        counted = ctx.sample_proximity(0, n, apply=False)
        assert np.array_equal(counted, inc.astype(np.uint32))
        plain = ctx.inc_counts()
        assert snp_only or proxi < 32 or (plain > counted).any()
        # one slot at a time, as the host driver does, then all again (idempotent)
        for i in range(n):
            assert ctx.sample_proximity(i, 1, apply=True)[0] == inc[i]
        assert np.array_equal(ctx.sample_proximity(0, n, apply=True), inc.astype(np.uint32))
        assert np.array_equal(ctx.inc_counts(), inc.astype(np.uint32))
        min_len = 1200
        include = (inc >= min_len).astype(np.uint8)
        D, N, dn = ctx.run_pair(include, norm=100, min_length=min_len, min_cov=0.0)
    finally:
        ctx.set_proximity(0)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=100, min_length=min_len, min_cov=0.0, proxi=proxi)
    assert dn == dno == int(include.sum()) and include[5] == 0
    assert dn >= n - 1 or proxi > 1000
    assert np.array_equal(_bits(D), _bits(Do)) and np.array_equal(_bits(N), _bits(No))


@pytest.mark.parametrize("snp_only", [False, True])
@pytest.mark.parametrize("proxi", [2, 33, 500, 70000])
def test_shared_mask_with_proximity(ctx, proxi, snp_only):
    """cdist.c:101-112 with -P: every included sample clears the runs between its close events against the first
    included sample from the shared mask; cmpFsaThrd then ignores proxi"""
    n, length = 40, 70000 + 5
    variant = 1 if snp_only else 0
    codes = _codes(n, length, seed=proxi + 1, snp=0.002)
    include = np.ones(n, dtype=np.uint8)
    include[[0, 9]] = 0
    gmask = oracle.global_mask(codes, include, proxi=proxi, variant=variant)
    plain = oracle.global_mask(codes, include)
    assert (gmask != plain).any()
    seqs, masks, inc = oracle.encode_samples(codes)
    ctx.set_proximity(proxi, snp_only)
    try:
        ctx.set_problem(n, length, pair=True)
        for i in range(n):
            if include[i]:
                ctx.put_sample_codes(i, codes[i])
                ctx.sync()
        ginc = ctx.build_global_mask(include)
        D, dn, ginc2 = ctx.run_global(include, norm=1000)
    finally:
        ctx.set_proximity(0)
    Do, dno, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=1000)
    assert ginc == ginc2 == ginco
    assert dn == dno == n - 2
    assert np.array_equal(_bits(D), _bits(Do))


def test_proximity_at_bench_scale_properties(ctx):
    """size-independent properties at a larger size: proxi = 0 equals the plain path, mismatches never grow and
    included counts never grow with proxi, and a proxi above the alignment length leaves one SNP per pair"""
    n, length = 256, 200000
    codes = _codes(n, length, seed=3, snp=0.001)
    seqs, masks, inc = oracle.encode_samples(codes)
    include = np.ones(n, dtype=np.uint8)
    prev = None
    for proxi in (0, 10, 1000, length + 1):
        api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=0, min_length=0, min_cov=0.0, proxi=proxi,
                               ctx=ctx)
        mism, ninc = ctx.raw_counts(n)
        mism, ninc = mism.copy(), ninc.copy()
        if proxi == 0:
            mo, no = oracle.raw_pair_matrix(seqs, masks, length)
            assert np.array_equal(mism, mo) and np.array_equal(ninc, no)
        else:
            assert (mism <= prev[0]).all() and (ninc <= prev[1]).all()
        if proxi > length:
            assert set(np.unique(mism)) <= {0, 1}
        prev = (mism, ninc)
