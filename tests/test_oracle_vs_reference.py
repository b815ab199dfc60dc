"""Pin the CPU oracle against the UNMODIFIED reference called in-process through
oracle/_ref/libccphylo_ref.so (reference objects + oracle/ref_shim.c): the real
get2BitTable / qseq2nibble / initIncPos / getIncPos / getNpos / maskProxi / fsacmpair and
the real fsaCmpThreadOut fan-out with cmpairFsaThrd / cmpFsaThrd.  Bit-exact."""
import os

import numpy as np
import pytest

import oracle
from ccphylo_b200 import synth

pytestmark = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")

ALPHABET = np.frombuffer(b"ACGTUNacgtun-RYSWKMBDHVXryswkmbdhvx\n\r *0Zz", dtype=np.uint8)


@pytest.mark.parametrize("flag", [1, 3, 9, 11])
@pytest.mark.parametrize("length", [0, 1, 31, 32, 33, 64, 95, 1000, 4097])
def test_encode_matches_reference(built, flag, length):
    rng = np.random.default_rng(1000 * flag + length)
    data = ALPHABET[rng.integers(0, len(ALPHABET), size=length)].tobytes()
    codes_r, seq_r, mask_r, unk_r, inc_r = oracle.ref_encode(data, flag=flag)
    codes = oracle.translate(data, flag)
    assert np.array_equal(codes, codes_r)
    seq, unk = oracle.pack(codes)
    mask, inc = oracle.known_mask(codes)
    assert unk == unk_r and inc == inc_r
    assert np.array_equal(seq, seq_r)
    assert np.array_equal(mask, mask_r)


def _random_set(n, length, seed, all_n=()):
    codes = synth.make_codes(n, length, seed=seed, snp=0.03, nrun=0.08, lower=0.02, gap=0.01)
    for k in all_n:
        codes[k, :] = 4
    return codes


@pytest.mark.parametrize("elem,scale", [(8, 1.0), (4, 1.0), (2, 10.0), (2, 100.0), (1, 0.01), (1, 1.0)])
@pytest.mark.parametrize("norm", [0, 1000, 1000000])
def test_pair_mode_matches_reference(built, elem, scale, norm):
    n, length = 13, 2085
    codes = _random_set(n, length, seed=elem * 7 + norm % 13, all_n=(0, 5))
    seqs, masks, inc = oracle.encode_samples(codes)
    min_len = int(0.5 * length)
    include = (inc >= min_len).astype(np.uint8)
    for tnum in (1, 3):
        Dr, Nr, dnr = oracle.ref_fsa_cmp(seqs, masks, include, length, pair=True, tnum=tnum, norm=norm,
                                         min_length=min_len, min_cov=0.5, elem_size=elem, byte_scale=scale)
        Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=norm, min_length=min_len, min_cov=0.5,
                                          elem_size=elem, byte_scale=scale)
        assert dnr == dno == n - 2
        assert np.array_equal(Dr.view(np.uint8), Do.view(np.uint8))
        assert np.array_equal(Nr.view(np.uint8), No.view(np.uint8))


def test_pair_gate_minus_one(built):
    """Pairs whose joint inclusion falls below the gate get -1 in D and keep their count in N."""
    n, length = 6, 640
    codes = _random_set(n, length, seed=5)
    codes[1, :300] = 4
    codes[2, 250:] = 4          # 1 & 2 overlap in 0 known positions... 250..300 unknown in both
    seqs, masks, inc = oracle.encode_samples(codes)
    include = np.ones(n, dtype=np.uint8)
    for elem, scale in ((8, 1.0), (4, 1.0), (2, 10.0), (1, 0.1)):
        Dr, Nr, dnr = oracle.ref_fsa_cmp(seqs, masks, include, length, pair=True, norm=100, min_length=200,
                                         min_cov=0.0, elem_size=elem, byte_scale=scale)
        Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=100, min_length=200, min_cov=0.0,
                                          elem_size=elem, byte_scale=scale)
        assert dnr == dno == n
        assert np.array_equal(Dr.view(np.uint8), Do.view(np.uint8))
        assert np.array_equal(Nr.view(np.uint8), No.view(np.uint8))
    D8, _, _ = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=100, min_length=200, min_cov=0.0)
    assert (D8 == -1.0).any()


@pytest.mark.parametrize("elem,scale", [(8, 1.0), (4, 1.0), (2, 10.0), (1, 0.5)])
@pytest.mark.parametrize("norm", [0, 1000])
def test_global_mode_matches_reference_without_exclusions(built, elem, scale, norm):
    n, length = 11, 3333
    codes = _random_set(n, length, seed=17 + elem)
    seqs, _, _ = oracle.encode_samples(codes)
    include = np.ones(n, dtype=np.uint8)
    gmask = oracle.global_mask(codes, include)
    for tnum in (1, 4):
        Dr, _, dnr = oracle.ref_fsa_cmp(seqs, gmask, include, length, pair=False, tnum=tnum, norm=norm,
                                        elem_size=elem, byte_scale=scale)
        Do, dno, ginc = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=norm, elem_size=elem,
                                              byte_scale=scale)
        assert dnr == dno == n
        assert ginc == oracle.lib().orc_mask_count(gmask, length)
        assert np.array_equal(Dr.view(np.uint8), Do.view(np.uint8))


def test_global_mask_accumulation_matches_reference(built):
    n, length = 7, 1999
    codes = _random_set(n, length, seed=99)
    include = np.ones(n, dtype=np.uint8)
    g = oracle.global_mask(codes, include)
    R = oracle.ref()
    W = oracle.words(length)
    gr = np.zeros(W + 1, dtype=np.uint32)
    R.refshim_init_mask.argtypes = [np.ctypeslib.ndpointer(np.uint32), __import__("ctypes").c_int]
    R.refshim_init_mask.restype = None
    R.refshim_init_mask(gr, length)
    ref0 = np.ascontiguousarray(codes[0]).copy()
    R.refshim_and_known(gr, ref0.copy(), ref0.copy(), length, 0)
    for i in range(1, n):
        R.refshim_and_known(gr, np.ascontiguousarray(codes[i]).copy(), ref0.copy(), length, 0)
    assert np.array_equal(g, gr[:W])


# ---------------------------------------------------------------------------------------------
# -P proximity masking (fsacmp.c:181-485): per-sample builders and the per-pair maskProxi
# ---------------------------------------------------------------------------------------------
def _proxi_set(n, length, seed):
    """clustered SNPs and unknowns so that every proxi has close and distant neighbours; the last column is kept
    equal and known (a SNP at len-1 makes the reference write past its scratch array, fsacmp.c:395-421)"""
    codes = synth.make_codes(n, length, seed=seed, snp=0.04, nrun=0.03, lower=0.03, gap=0.01)
    if length:
        codes[:, -1] = 2
    return codes


@pytest.mark.parametrize("flag", [1, 8, 32])
@pytest.mark.parametrize("proxi", [1, 2, 5, 31, 32, 33, 100, 5000])
def test_per_sample_builders_with_proximity(built, flag, proxi):
    variant = oracle.variant_of(flag)
    for length in (1, 31, 32, 64, 97, 1500):
        codes = _proxi_set(3, length, seed=proxi * 7 + length)
        # pair mode: seq against itself (cdist.c:91); shared-mask mode: seq against ref on the running mask (:111)
        for s, r in ((0, 0), (1, 0), (2, 0)):
            want = oracle.full_mask(length).copy()
            pad = np.zeros(len(want) + 1, np.uint32)
            pad[:len(want)] = want
            oracle.ref_inc_pos(pad, codes[s], codes[r], proxi, flag)
            got = oracle.full_mask(length).copy()
            oracle.inc_pos(got, codes[s], codes[r], proxi, variant)
            assert np.array_equal(got, pad[:len(want)]), (length, s, r)
            assert pad[len(want)] == 0


@pytest.mark.parametrize("flag,builder", [(1, 0), (8, 1), (32, 2)], ids=["getIncPos", "getIncPosInsig", "getIncPosInsigPrune"])
@pytest.mark.parametrize("proxi", [0, 1, 2, 5, 31, 33, 100, 5000])
def test_trim_pass_on_the_iupac_alphabet(built, flag, builder, proxi):
    """orc_trim_pass (what `trim` does per sample, on getIupacBitTable's codes: 0-3 bases, 4 unknown, 5 gap, 6-15
    ambiguity letters, +16 soft-masked) against the reference's own getIncPos / getIncPosInsig / getIncPosInsigPrune:
    the sample against itself, and a later sample against the stored reference sample (soft flags stripped)"""
    rng = np.random.default_rng(flag * 1000 + proxi)
    for length in (1, 31, 32, 64, 97, 1500):
        base = rng.integers(0, 4, size=length).astype(np.uint8)

        def sample():
            c = base.copy()
            sub = rng.random(length) < 0.05
            c[sub] = rng.integers(0, 16, size=int(sub.sum()))                 # substitutions, unknowns, gaps, ambiguity letters
            soft = (rng.random(length) < 0.05) & (c != 4)
            c[soft] |= 16
            if length:
                c[-1] &= 3
            return c

        first, later = sample(), sample()
        # the sample against itself (trim.c:201: always getIncPos)
        want = np.zeros(oracle.words(length) + 1, np.uint32)
        want[:-1] = oracle.full_mask(length)
        oracle.ref_inc_pos(want, first, first, proxi, 1)
        got = oracle.full_mask(length).copy()
        stored = first.copy()
        oracle.trim_pass(got, stored, None, proxi, 0)
        assert np.array_equal(got, want[:-1]) and want[-1] == 0, (length, "self")
        assert np.array_equal(stored, first & 15)
        # a later sample against it, on the running mask (trim.c:176-177)
        want2 = want.copy()
        oracle.ref_inc_pos(want2, later, stored, proxi, flag)
        got2 = got.copy()
        oracle.trim_pass(got2, later.copy(), stored, proxi, builder)
        assert np.array_equal(got2, want2[:-1]) and want2[-1] == 0, (length, "against the reference sample")


@pytest.mark.parametrize("proxi", [1, 2, 3, 7, 31, 32, 33, 64, 200, 100000])
def test_pair_counts_with_proximity(built, proxi):
    for length in (2, 33, 64, 96, 127, 1000, 3001):
        codes = _proxi_set(6, length, seed=proxi + length)
        seqs, masks, _ = oracle.encode_samples(codes, proxi=proxi)
        some = False
        for i in range(1, 6):
            for j in range(i):
                want = oracle.ref_pair(seqs[i], seqs[j], masks[i], masks[j], length, proxi)
                got = oracle.pair_counts_proxi(seqs[i], seqs[j], masks[i], masks[j], length, proxi)
                assert got == want, (length, i, j)
                plain = oracle.pair_counts_proxi(seqs[i], seqs[j], masks[i], masks[j], length, 0)
                some |= plain != got
        assert some or length < 64, "the inputs never triggered proximity masking"


@pytest.mark.parametrize("elem,scale", [(8, 1.0), (4, 1.0), (2, 10.0), (1, 0.5)])
@pytest.mark.parametrize("proxi", [3, 40])
def test_pair_mode_with_proximity_matches_reference(built, elem, scale, proxi):
    n, length = 12, 2085
    codes = _proxi_set(n, length, seed=elem + proxi)
    codes[4, :] = 4
    seqs, masks, inc = oracle.encode_samples(codes, proxi=proxi)
    min_len = int(0.5 * length)
    include = (inc >= min_len).astype(np.uint8)
    for tnum in (1, 3):
        Dr, Nr, dnr = oracle.ref_fsa_cmp(seqs, masks, include, length, pair=True, tnum=tnum, norm=1000,
                                         min_length=min_len, min_cov=0.5, proxi=proxi, elem_size=elem, byte_scale=scale)
        Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=1000, min_length=min_len, min_cov=0.5,
                                          elem_size=elem, byte_scale=scale, proxi=proxi)
        assert dnr == dno
        assert np.array_equal(Dr.view(np.uint8), Do.view(np.uint8))
        assert np.array_equal(Nr.view(np.uint8), No.view(np.uint8))


@pytest.mark.parametrize("proxi", [1, 2, 3, 7, 31, 32, 33, 64, 200, 100000])
def test_pair_mask_with_proximity_and_the_variants_under_it(built, proxi):
    """-V with -P (fsacmpthrd.c:410-414): maskProxi's mask itself, word for word, and fsacmpairint's lines under it"""
    cleared = 0
    for length in (1, 31, 32, 33, 64, 95, 128, 129, 700, 4099):
        codes = _proxi_set(5, length, seed=3 * proxi + length)
        seqs, masks, _ = oracle.encode_samples(codes, proxi=proxi)
        for i in range(1, 5):
            for j in range(i):
                want = oracle.ref_mask_proxi(seqs[i], seqs[j], masks[i], masks[j], length, proxi)
                got = oracle.pair_mask_proxi(seqs[i], seqs[j], masks[i], masks[j], length, proxi)
                assert np.array_equal(got, want), (length, i, j)
                cleared += int(np.bitwise_count((masks[i] & masks[j]) ^ got).sum())
                text, r = oracle.ref_variants(True, i, j, seqs[i], seqs[j], want, length)
                lst = oracle.list_variants(seqs[i], seqs[j], got, length)
                assert oracle.variant_text(i, j, lst) == text and (r >> 32) == len(lst)
    assert cleared > 0


# ---------------------------------------------------------------------------------------------
# -V variant listing (fsacmp.c:635-737)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("length", [1, 31, 32, 33, 64, 97, 1500, 4099])
def test_variant_listing_matches_reference(built, length):
    codes = synth.make_codes(5, length, seed=length, snp=0.05, nrun=0.05, lower=0.03, gap=0.01)
    seqs, masks, _ = oracle.encode_samples(codes)
    gmask = oracle.global_mask(codes, np.ones(5, np.uint8))
    total = 0
    for i in range(1, 5):
        for j in range(i):
            pm = masks[i] & masks[j]
            text, r = oracle.ref_variants(True, i, j, seqs[i], seqs[j], pm, length)
            got = oracle.list_variants(seqs[i], seqs[j], pm, length)
            assert oracle.variant_text(i, j, got) == text
            assert (r >> 32) == len(got)
            total += len(got)
            text, r = oracle.ref_variants(False, i, j, seqs[i], seqs[j], gmask, length)
            got = oracle.list_variants(seqs[i], seqs[j], gmask, length)
            assert oracle.variant_text(i, j, got) == text and r == len(got)
    assert total > 0 or length < 31


# ---------------------------------------------------------------------------------------------
# -y methylation motif masking (meth.c:52-159, methparse.c:27-296)
# ---------------------------------------------------------------------------------------------
MOTIF_FILES = [">dam\ngAtc\n", ">dam\ngAtc\n>dcm\ncCwgg\n>x\nrgATcnny\n", "gatC\n>multi line\ncC\nwg\ng\n>odd chars\nGA-NT.C\n",
               ">long\nacgtacgtAcgtacgtacgtacgTacgtacgt\n>three\ngAn\n>iupac\nRYSWKMBDHVN\n"]


@pytest.mark.parametrize("which", range(len(MOTIF_FILES)))
def test_motif_masking_matches_reference(built, tmp_path, which):
    path = str(tmp_path / "motifs.fsa")
    with open(path, "w") as f:
        f.write(MOTIF_FILES[which])
    # all but the first file hold motifs on which the reference reads past an array (see oracle.parse_motifs):
    # as_built reproduces what this container's build of it does; on the first file both readings agree
    motifs = oracle.parse_motifs(MOTIF_FILES[which], as_built=True)
    assert motifs and all(len(m) <= 32 for m in motifs)
    assert (motifs == oracle.parse_motifs(MOTIF_FILES[which])) == (which == 0)
    hits = 0
    for length in (1, 4, 31, 32, 33, 64, 65, 700, 4099):
        codes = synth.make_codes(3, length, seed=length + which, snp=0.05, nrun=0.05, lower=0.03, gap=0.01)
        if length >= 700:
            codes[1, 100:132] = np.tile([0, 1, 2, 3], 8)            # the 32-mer of the last file
        seqs, masks, _ = oracle.encode_samples(codes)
        for s in range(3):
            W = oracle.words(length)
            want = np.zeros(W + 2, np.uint32)
            want[:W] = oracle.full_mask(length)
            n_ref = oracle.ref_mask_motifs(path, seqs[s], want, length)
            got = oracle.full_mask(length).copy()
            n = oracle.mask_motifs(seqs[s], got, length, motifs)
            assert n == n_ref, (length, s)
            assert np.array_equal(got, want[:W]), (length, s)
            assert want[W] == 0 and want[W + 1] == 0
            hits += n
    assert hits > 0


# ---------------------------------------------------------------------------------------------
# -a: one row against an existing matrix (cmpFsaRowThrd fsacmpthrd.c:482-580), through the reference binary
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flag,proxi", [(3, 0), (3, 9), (3, 200), (35, 50), (11, 4)])
def test_row_restatement_matches_reference_binary(built, tmp_path, flag, proxi):
    import subprocess
    ref_bin = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ccphylo")
    n, length = 7, 6000 + 11
    rows = synth.make_ascii(n + 1, length, seed=flag + proxi, snp=0.01, nrun=0.02)
    rows[3, : length - 2000] = ord("N")
    td = str(tmp_path)
    for i in range(n + 1):
        synth.write_fasta(os.path.join(td, f"s{i}.fsa"), rows[i], header="ref", width=60)
    files = [os.path.join(td, f"s{i}.fsa") for i in range(n)]
    common = ["-f", str(flag), "-P", str(proxi), "-W", "1000"]
    p = subprocess.run([ref_bin, "dist", "-r", "ref", "-C", "0.0", "-i"] + files + common + ["-o", "m.phy", "-n", "m.num"],
                       cwd=td, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    p = subprocess.run([ref_bin, "dist", "-r", "ref", "-a", os.path.join(td, f"s{n}.fsa"), "-i", files[0]] + common +
                       ["-o", "m.phy", "-n", "m.num"], cwd=td, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    want_d = [float(x) for x in open(os.path.join(td, "m.phy")).read().splitlines()[-1].split("\t")[1:]]
    want_n = [float(x) for x in open(os.path.join(td, "m.num")).read().splitlines()[-1].split("\t")[1:]]
    codes = np.stack([oracle.translate(rows[i].tobytes(), flag) for i in range(n + 1)])
    variant = oracle.variant_of(flag)
    seqs, masks, _ = oracle.encode_samples(codes)
    own = oracle.full_mask(length).copy()
    oracle.inc_pos(own, codes[n], codes[n], proxi, variant)
    masks[n] = own
    D, N = oracle.fsa_cmp_row(seqs, masks, n, length, norm=1000, min_length=1, min_cov=0.5, proxi=proxi, variant=variant,
                              codes=codes)
    assert len(want_d) == n and want_d[3] == -1
    assert np.array_equal(N, np.array(want_n))
    assert np.all(np.abs(D - np.array(want_d)) <= 0.5000001e-9 + 1e-12 * np.abs(D))
