"""GPU parity of `dist -a` (one more sample against an existing matrix): ccg_run_row against
cmpFsaRowThrd (fsacmpthrd.c:482-580) and ccg_mat_run_row against cmpMatRowThrd (ltdmatrixthrd.c:111-181) through
the C-ABI and the oracle, and the host driver's -a against the reference binary: the updated .phy / .num files
byte for byte for FASTA inputs, within the printed precision for .mat inputs."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers
import oracle
from ccphylo_b200 import api, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.environ.get("CCPHYLO_TEST_BIN") or os.path.join(ROOT, "ccphylo_b200", "bin", "ccphylo-b200")   # (tests/csrc/mock_ccg.c gives a CPU driver)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ccphylo")


@pytest.fixture(scope="module")
def ctx(built):
    c = api.Context()
    yield c
    c.close()


@pytest.mark.parametrize("kernel", [api.KERNEL_POPC, api.KERNEL_UMMA, api.KERNEL_AUTO])
@pytest.mark.parametrize("n,length", [(1, 40), (5, 33), (70, 1000), (255, 4099), (256, 4099), (300, 20000 + 3), (600, 9000)])
def test_fasta_row_against_the_oracle(ctx, n, length, kernel):
    codes = synth.make_codes(n + 1, length, seed=n + length, snp=0.02, nrun=0.05)
    if n > 3:
        codes[2, : length - length // 4] = 4          # a column sample that fails the overlap gate
    seqs, masks, inc = oracle.encode_samples(codes)
    ctx.set_kernel(kernel)
    try:
        ctx.set_problem(n + 1, length, pair=True)
        ctx.put_samples_packed(seqs, masks)
        for norm in (0, 1000000):
            D, N = ctx.run_row(n, norm=norm, min_length=1, min_cov=0.5)
            Do, No = oracle.fsa_cmp_row(seqs, masks, n, length, norm=norm, min_length=1, min_cov=0.5)
            assert np.array_equal(D.view(np.uint8), Do.view(np.uint8))
            assert np.array_equal(N.view(np.uint8), No.view(np.uint8))
        if n > 3:
            assert D[2] == -1.0 and N[2] == 0.0
        # -V with -a: the row's variant lists
        got = ctx.list_variants(row=n)
        want = [((n, j), v) for j in range(n) for v in [oracle.list_variants(seqs[n], seqs[j], masks[n] & masks[j], length)] if v]
        assert got == want
        # the full matrix still comes out right on the same store afterwards (the row mode leaves no state behind)
        Df, Nf, dn = ctx.run_pair(norm=1000, min_length=1, min_cov=0.5)
        Dw, Nw, dnw = oracle.fsa_cmp_pair(seqs, masks, np.ones(n + 1, np.uint8), length, norm=1000, min_length=1, min_cov=0.5)
        assert dn == dnw and np.array_equal(Df.view(np.uint8), Dw.view(np.uint8)) and np.array_equal(Nf, Nw)
    finally:
        ctx.set_kernel(api.KERNEL_AUTO)


@pytest.mark.parametrize("snp_only", [False, True])
@pytest.mark.parametrize("proxi", [1, 6, 33, 400])
@pytest.mark.parametrize("n,length", [(3, 97), (40, 4099), (300, 9000 + 1)])
def test_fasta_row_with_proximity_against_the_oracle(ctx, n, length, proxi, snp_only):
    """-a with -P: the pair's mask is the new sample's own mask put through the per-sample builder against each
    column sample (not maskProxi)"""
    variant = 1 if snp_only else 0
    codes = synth.make_codes(n + 1, length, seed=n + length + proxi, snp=0.02, nrun=0.01, lower=0.01, gap=0.005)
    if n > 3:
        codes[2, : length - length // 4] = 4
    seqs, masks, inc = oracle.encode_samples(codes)                                   # columns: known positions
    own = oracle.full_mask(length).copy()
    oracle.inc_pos(own, codes[n], codes[n], proxi, variant)                           # includeadd
    masks_o = masks.copy()
    masks_o[n] = own
    ctx.set_proximity(proxi, snp_only)
    try:
        ctx.set_problem(n + 1, length, pair=True)
        for i in range(n + 1):
            ctx.put_sample_codes(i, codes[i])
            ctx.sync()
        before = ctx.inc_counts().copy()
        for norm in (0, 1000):
            D, N = ctx.run_row(n, norm=norm, min_length=1, min_cov=0.5)
            Do, No = oracle.fsa_cmp_row(seqs, masks_o, n, length, norm=norm, min_length=1, min_cov=0.5, proxi=proxi,
                                        variant=variant, codes=codes)
            assert np.array_equal(D.view(np.uint8), Do.view(np.uint8))
            assert np.array_equal(N.view(np.uint8), No.view(np.uint8))
        assert "k_row_proxi" in ctx.last_kernel
        # the new sample's planes were put back as uploaded
        assert np.array_equal(ctx.inc_counts(), before)
        # -V with -a and -P: fsacmpairint walks the mask the builder leaves for each column sample (fsacmpthrd.c:545-553)
        got = ctx.list_variants(row=n)
        want = []
        for j in range(n):
            pm = own.copy()
            oracle.inc_pos(pm, codes[j], codes[n], proxi, variant)
            v = oracle.list_variants(seqs[n], seqs[j], pm, length)
            if v:
                want.append(((n, j), v))
        assert got == want
        assert np.array_equal(ctx.inc_counts(), before)
        D2, N2 = ctx.run_row(n, norm=1000, min_length=1, min_cov=0.5)
        assert np.array_equal(D2.view(np.uint8), Do.view(np.uint8))
    finally:
        ctx.set_proximity(0)
    plain, _ = oracle.fsa_cmp_row(seqs, masks, n, length, norm=1000, min_length=1, min_cov=0.5)
    assert proxi < 33 or length < 1000 or not np.array_equal(plain, Do)
    D0, N0 = ctx.run_row(n, norm=1000, min_length=1, min_cov=0.5)
    assert np.array_equal(D0, plain)


@pytest.mark.parametrize("method", ["cos", "chi2", "bc", "z", "l1", "nl2"])
def test_mat_row_against_the_oracle(ctx, method):
    from test_gpu_mat import random_counts, close
    n, length = 41, 3000 + 7
    counts, totals = random_counts(n + 1, length, seed=17)
    lens = np.full(n + 1, length, np.int32)
    ctx.mat_set_problem(n + 1, length)
    for i in range(n + 1):
        ctx.mat_put_sample(i, counts[i], totals[i])
    for norm in (0, 1000):
        D, N, rows = ctx.mat_run_row(n, method=method, norm=norm)
        Do, No, dno = oracle.mat_matrix(counts, totals, lens, np.ones(n + 1, np.uint8), method=method, norm=norm)
        row = slice(n * (n - 1) // 2, n * (n - 1) // 2 + n)
        assert np.array_equal(N, No[row]) and np.array_equal(rows, No[row].astype(np.uint32))
        assert close(D, Do[row])


def _run(cmd, cwd):
    return subprocess.run(cmd, capture_output=True, text=True, cwd=cwd, timeout=300)


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
@pytest.mark.parametrize("args", [["-f", "3"], ["-f", "3", "-W", "1000000"], ["-f", "2", "-W", "1000"], ["-f", "11", "-x", "4"],
                                  ["-f", "3", "-P", "9", "-W", "1000"], ["-f", "35", "-P", "50"]],
                         ids=["relaxed", "normalised", "strict-names", "insig-precision", "proximity", "proximity-prune"])
def test_cli_add_fasta_row_against_the_reference_binary(built, tmp_path, args):
    """matrix of n samples by the reference, then -a with one more sample by the reference and by this driver on
    copies of the same files: both must leave the same bytes in the .phy and the .num"""
    n, length = 9, 12000 + 5
    rows = synth.make_ascii(n + 1, length, seed=21, snp=0.01, nrun=0.02)
    rows[3, : length - 4000] = ord("N")        # passes no pair gate with the default -C 0.5: "No sufficient overlap"
    dirs = {}
    for tag in ("reference", "driver"):
        d = tmp_path / tag
        d.mkdir()
        for i in range(n + 1):
            synth.write_fasta(str(d / f"s{i}.fsa"), rows[i], header="ref", width=60)
        dirs[tag] = str(d)
    d = dirs["reference"]
    files = [os.path.join(d, f"s{i}.fsa") for i in range(n)]
    p = _run([REF_BIN, "dist", "-r", "ref", "-C", "0.0", "-i"] + files + args + ["-o", "m.phy", "-n", "m.num"], d)
    assert p.returncode == 0, p.stderr
    shutil.copy(os.path.join(d, "m.phy"), os.path.join(dirs["driver"], "m.phy"))
    shutil.copy(os.path.join(d, "m.num"), os.path.join(dirs["driver"], "m.num"))
    outs = {}
    for tag, exe in (("reference", REF_BIN), ("driver", BIN)):
        d = dirs[tag]
        p = _run([exe, "dist", "-r", "ref", "-a", os.path.join(d, f"s{n}.fsa"), "-i", os.path.join(d, "s0.fsa"), "-t", "3"] + args +
                 ["-o", "m.phy", "-n", "m.num"], d)
        assert p.returncode == 0, p.stderr
        outs[tag] = (open(os.path.join(d, "m.phy")).read(), open(os.path.join(d, "m.num")).read(),
                     sorted(p.stderr.replace(d + "/", "").splitlines()))
    assert outs["reference"][0].startswith("%10d\n" % (n + 1)) and "\t-1" in outs["reference"][0].splitlines()[-1]
    assert outs["driver"] == outs["reference"]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
@pytest.mark.parametrize("extra", [[], ["-P", "8"], ["-P", "120"]], ids=["plain", "P8", "P120"])
def test_cli_add_row_with_variant_file_against_the_reference_binary(built, tmp_path, extra):
    """-a with -V: the new row's variant lists are appended to the file (fsacmpthrd.c:632-637, :552-553); with -P under
    the mask the per-sample builder leaves for each column sample"""
    n, length = 6, 8000 + 3
    rows = synth.make_ascii(n + 1, length, seed=33 + len(extra) * 7, snp=0.01, nrun=0.01)
    outs = {}
    for tag, exe in (("reference", REF_BIN), ("driver", BIN)):
        d = tmp_path / tag
        d.mkdir()
        d = str(d)
        for i in range(n + 1):
            synth.write_fasta(os.path.join(d, f"s{i}.fsa"), rows[i], header="ref", width=60)
        files = [os.path.join(d, f"s{i}.fsa") for i in range(n)]
        p = _run([REF_BIN, "dist", "-r", "ref", "-f", "3", "-t", "1", "-V", "v.txt"] + extra + ["-i"] + files + ["-o", "m.phy", "-n", "m.num"], d)
        assert p.returncode == 0, p.stderr
        p = _run([exe, "dist", "-r", "ref", "-f", "3", "-t", "1", "-V", "v.txt"] + extra + ["-a", os.path.join(d, f"s{n}.fsa"), "-i", files[0],
                  "-o", "m.phy", "-n", "m.num"], d)
        assert p.returncode == 0, p.stderr
        outs[tag] = (open(os.path.join(d, "v.txt")).read(), open(os.path.join(d, "m.phy")).read(), open(os.path.join(d, "m.num")).read())
    assert ("(%d, 0)\t" % n) in outs["reference"][0] and outs["reference"][0].startswith("(1, 0)\t")
    assert outs["driver"] == outs["reference"]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
@pytest.mark.parametrize("method", ["cos", "chi2", "nbc"])
def test_cli_add_mat_row_against_the_reference_binary(built, tmp_path, method):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from make_golden_mat import random_sample
    rng = np.random.default_rng(5)
    length, n = 1500, 6
    ref = "".join("ACGT"[k] for k in rng.integers(0, 4, size=length))
    dirs = {}
    texts = [helpers.mat_text("tmpl", ref, random_sample(rng, ref, 40, 0.3, 0.02, 0.02)) for _ in range(n + 1)]
    for tag in ("reference", "driver"):
        d = tmp_path / tag
        d.mkdir()
        for i in range(n + 1):
            (d / f"s{i}.mat").write_text(texts[i])
        dirs[tag] = str(d)
    d = dirs["reference"]
    files = [os.path.join(d, f"s{i}.mat") for i in range(n)]
    common = ["-d", method, "-W", "1000"]
    p = _run([REF_BIN, "dist", "-r", "tmpl", "-i"] + files + common + ["-o", "m.phy", "-n", "m.num"], d)
    assert p.returncode == 0, p.stderr
    shutil.copy(os.path.join(d, "m.phy"), os.path.join(dirs["driver"], "m.phy"))
    shutil.copy(os.path.join(d, "m.num"), os.path.join(dirs["driver"], "m.num"))
    outs = {}
    for tag, exe in (("reference", REF_BIN), ("driver", BIN)):
        d = dirs[tag]
        p = _run([exe, "dist", "-r", "tmpl", "-a", os.path.join(d, f"s{n}.mat"), "-i", os.path.join(d, "s0.mat")] + common +
                 ["-o", "m.phy", "-n", "m.num"], d)
        assert p.returncode == 0, p.stderr
        outs[tag] = (open(os.path.join(d, "m.phy")).read(), open(os.path.join(d, "m.num")).read())
    assert outs["driver"][1] == outs["reference"][1]
    got, want = helpers.parse_phy(outs["driver"][0]), helpers.parse_phy(outs["reference"][0])
    assert got[0][0] == want[0][0] and len(want[0][0]) == n + 1
    g, w = np.array(got[0][1]), np.array(want[0][1])
    assert np.all(np.abs(g - w) <= 1.01e-9 + 1e-6 * np.abs(w))
    # the rows of the existing matrix were not touched
    assert outs["driver"][0].splitlines()[1:n + 1] == outs["reference"][0].splitlines()[1:n + 1]
