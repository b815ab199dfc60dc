"""The CPU oracle against the golden vectors produced by the unmodified reference binary
(tests/golden/fasta_dist.json, generator scripts/make_golden.py; includes SURVEY.md
Appendix C #1, #2, #6, #10, #11).  Byte-identical .phy / .num / stderr text."""
import numpy as np
import pytest

import helpers

POOL, CASES = helpers.golden_cases()


def test_fixture_is_populated():
    assert len(CASES) >= 60
    names = {c["name"] for c in CASES}
    assert {"c1_pair", "c1_pair_W", "c1_global", "c6_excluded_pair", "c1_short_W", "c1_byte_W"} <= names


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_reference_text(built, case):
    phy, num, err = helpers.replay(case, POOL, helpers.oracle_backend)
    assert err == case["stderr"]
    assert phy == case["phy"]
    assert num == case["num"]


def test_appendix_c13_packing_bit_order(built):
    """35-base string C + 30xA + TGNA."""
    import oracle

    codes = oracle.translate(b"C" + b"A" * 30 + b"TGNA", flag=1)
    assert len(codes) == 35
    seq, unknown = oracle.pack(codes)
    mask, inc = oracle.known_mask(codes)
    assert unknown == 1 and inc == 34
    assert [int(x) for x in seq] == [0x4000000000000003, 0x8000000000000000]
    assert [int(x) for x in mask] == [0xFFFFFFFF, 0xA0000000]


def test_appendix_c12_formulations(built):
    """one-hot M.M^T - X.X^T and tetrahedral (3 M.M^T - T.T^T)/4 equal the
    oracle's integer counts (this is the algebra the tensor-core kernel relies on)."""
    import oracle
    from ccphylo_b200 import synth

    n, L = 12, 1000
    codes = synth.make_codes(n, L, seed=3, snp=0.05, nrun=0.05)
    seqs, masks, _ = oracle.encode_samples(codes)
    mism, ninc = oracle.raw_pair_matrix(seqs, masks, L, nthreads=2)
    known = (codes != 4).astype(np.int64)
    onehot = np.stack([(codes == b).astype(np.int64) for b in range(4)], axis=2).reshape(n, -1)
    tet = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1], [0, 0, 0]], dtype=np.int64)[codes].reshape(n, -1)
    inc_m = known @ known.T
    mism_onehot = inc_m - onehot @ onehot.T
    S = tet @ tet.T
    mism_tet = (3 * inc_m - S) // 4
    assert ((3 * inc_m - S) % 4 == 0).all()
    k = 0
    for r in range(1, n):
        for c in range(r):
            assert mism[k] == mism_onehot[r, c] == mism_tet[r, c]
            assert ninc[k] == inc_m[r, c]
            k += 1
