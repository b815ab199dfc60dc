"""GPU parity of -y / --methylation_motifs: ccg_mask_motifs against the oracle's restatement
of maskMotifs (meth.c:52-159; pinned to the reference in tests/test_oracle_vs_reference.py) on the same motif
lists, and the host driver's -y against the reference binary (which also covers the driver's motif-file parser,
reverse complements and the reference's as-built handling of ambiguity letters, see host/motifs.c)."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from ccphylo_b200 import api, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.environ.get("CCPHYLO_TEST_BIN") or os.path.join(ROOT, "ccphylo_b200", "bin", "ccphylo-b200")   # (tests/csrc/mock_ccg.c gives a CPU driver)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ccphylo")

FILES = [">dam\ngAtc\n", ">dam\ngAtc\n>dcm\ncCwgg\n>x\nrgATcnny\n", "gatC\n>multi line\ncC\nwg\ng\n>odd chars\nGA-NT.C\n",
         ">long\nacgtacgtAcgtacgtacgtacgTacgtacgt\n>three\ngAn\n>iupac\nRYSWKMBDHVN\n>short\nAc\n"]


@pytest.fixture(scope="module")
def ctx(built):
    c = api.Context()
    yield c
    c.set_motifs([])
    c.close()


@pytest.mark.parametrize("as_built", [False, True])
@pytest.mark.parametrize("which", range(len(FILES)))
@pytest.mark.parametrize("n,length", [(3, 1), (3, 4), (5, 31), (5, 32), (7, 33), (9, 127), (33, 128), (40, 129), (70, 4099),
                                      (130, 20000 + 7)])
def test_motif_masks_against_the_oracle(ctx, n, length, which, as_built):
    motifs = oracle.parse_motifs(FILES[which], as_built=as_built)
    codes = synth.make_codes(n, length, seed=n + length + which, snp=0.05, nrun=0.03, lower=0.02, gap=0.01)
    if length >= 4099:
        codes[1, 100:132] = np.tile([0, 1, 2, 3], 8)            # the 32-mer, across a word boundary
        codes[2, 4064:4096] = np.tile([0, 1, 2, 3], 8)          # and across a 128-base chunk boundary
        codes[0, length - 4:] = [2, 0, 3, 1]                    # gatc as the very last bases
    seqs, masks, inc = oracle.encode_samples(codes)
    want_masks = masks.copy()
    hits = 0
    for i in range(n):
        hits += oracle.mask_motifs(seqs[i], want_masks[i], length, motifs)
    want_inc = np.array([oracle.lib().orc_mask_count(want_masks[i], length) for i in range(n)], dtype=np.uint32)
    ctx.set_motifs(motifs)
    try:
        ctx.set_problem(n, length, pair=True)
        for i in range(n):
            ctx.put_sample_codes(i, codes[i])
            ctx.sync()
        got_inc = np.concatenate([ctx.mask_motifs(0, n // 2), ctx.mask_motifs(n // 2, n - n // 2)])
        assert np.array_equal(got_inc, want_inc)
        assert np.array_equal(ctx.inc_counts(), want_inc)
        D, N, dn = ctx.run_pair(norm=1000, min_length=0, min_cov=0.0)
    finally:
        ctx.set_motifs([])
    assert hits > 0 or length < 200
    include = np.ones(n, np.uint8)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, want_masks, include, length, norm=1000, min_length=0, min_cov=0.0)
    assert dn == dno
    assert np.array_equal(N, No) and np.array_equal(D.view(np.uint8), Do.view(np.uint8))


# (with the two-letter motif every sample would be trimmed away)
CLI_FILES = FILES[:3] + [FILES[3].replace(">short\nAc\n", "")]


def _run(cmd, cwd):
    return subprocess.run(cmd, capture_output=True, text=True, cwd=cwd, timeout=300)


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
@pytest.mark.parametrize("which", range(len(CLI_FILES)))
@pytest.mark.parametrize("flag", ["3", "1"], ids=["pair", "shared-mask"])
@pytest.mark.parametrize("msa", [False, True], ids=["files", "msa"])
def test_cli_motif_masking_against_the_reference_binary(built, tmp_path, msa, flag, which):
    td = str(tmp_path)
    n, length = 10, 15000 + 9
    rows = synth.make_ascii(n, length, seed=which + 3, snp=0.01, nrun=0.004)
    if flag == "3":
        rows[6, 20:] = ord("N")
    mot = os.path.join(td, "motifs.fsa")
    with open(mot, "w") as f:
        f.write(CLI_FILES[which])
    if msa:
        path = os.path.join(td, "aln.fsa")
        with open(path, "wb") as f:
            for i in range(n):
                f.write(b">s%d\n" % i)
                for s0 in range(0, length, 60):
                    f.write(rows[i, s0:s0 + 60].tobytes() + b"\n")
        inputs = ["-i", path]
    else:
        files = []
        for i in range(n):
            fp = os.path.join(td, f"s{i:02d}.fsa")
            synth.write_fasta(fp, rows[i], header="ref", width=60)
            files.append(fp)
        inputs = ["-r", "ref", "-i"] + files
    outs = {}
    for tag, exe in (("reference", REF_BIN), ("driver", BIN)):
        phy, num = os.path.join(td, tag + ".phy"), os.path.join(td, tag + ".num")
        p = _run([exe, "dist", "-f", flag, "-y", mot, "-W", "1000", "-t", "3", "-o", phy, "-n", num] + inputs, td)
        assert p.returncode == 0, p.stderr[-2000:]
        outs[tag] = (open(phy).read(), open(num).read(), p.stderr)
    plain = _run([REF_BIN, "dist", "-f", flag, "-W", "1000", "-o", os.path.join(td, "plain.phy")] + inputs, td)
    assert open(os.path.join(td, "plain.phy")).read() != outs["reference"][0] or plain.stderr != outs["reference"][2]
    assert outs["driver"] == outs["reference"]


# ---- -y together with -V and with -P: the reference runs them, so does the driver ----
def _inputs(td, msa, rows):
    n, length = rows.shape
    if msa:
        path = os.path.join(td, "aln.fsa")
        with open(path, "wb") as f:
            for i in range(n):
                f.write(b">s%d\n" % i)
                for s0 in range(0, length, 60):
                    f.write(rows[i, s0:s0 + 60].tobytes() + b"\n")
        return ["-i", path]
    files = []
    for i in range(n):
        fp = os.path.join(td, f"s{i:02d}.fsa")
        synth.write_fasta(fp, rows[i], header="ref", width=60)
        files.append(fp)
    return ["-r", "ref", "-i"] + files


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
@pytest.mark.parametrize("flag", ["3", "1"], ids=["pair", "shared-mask"])
@pytest.mark.parametrize("msa", [False, True], ids=["files", "msa"])
def test_cli_motifs_with_variant_listing(built, tmp_path, msa, flag):
    # maskMotifs changes the inclusion masks only; fsacmpairint / fsacmprint then compare the UNCHANGED packed words
    # (fsacmp.c:646-737): the device keeps the code planes untouched until the listing is done
    td = str(tmp_path)
    n, length = 9, 6000 + 11
    rows = synth.make_ascii(n, length, seed=17, snp=0.01, nrun=0.004)
    mot = os.path.join(td, "motifs.fsa")
    with open(mot, "w") as f:
        f.write(CLI_FILES[0])
    inputs = _inputs(td, msa, rows)
    outs = {}
    for tag, exe in (("reference", REF_BIN), ("driver", BIN)):
        phy, num, var = (os.path.join(td, tag + ext) for ext in (".phy", ".num", ".var"))
        p = _run([exe, "dist", "-f", flag, "-y", mot, "-V", var, "-W", "1000", "-t", "1", "-o", phy, "-n", num] + inputs, td)
        assert p.returncode == 0, p.stderr[-2000:]
        outs[tag] = (open(phy).read(), open(num).read(), open(var).read(), p.stderr)
    assert len(outs["reference"][2]) > 100
    assert outs["driver"] == outs["reference"]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
@pytest.mark.parametrize("flag,proxi", [("3", "7"), ("3", "40"), ("11", "12"), ("35", "9"), ("1", "7"), ("1", "60"), ("9", "12"), ("33", "9")])
@pytest.mark.parametrize("msa", [False, True], ids=["files", "msa"])
def test_cli_motifs_with_proximity(built, tmp_path, msa, flag, proxi):
    # pair mode, cdist.c:90-91: maskMotifs, then getIncPosPtr(includes[i], seq, seq, proxi); both only clear mask bits.
    # shared-mask mode, cdist.c:109-111 / :137-138: the motif sites narrow the shared mask, the proximity events are
    # defined on the sequences -- ccg_build_global_mask applies the motifs after its proximity pass
    td = str(tmp_path)
    n, length = 9, 6000 + 11
    rows = synth.make_ascii(n, length, seed=23, snp=0.02, nrun=0.01, gap=0.004)
    mot = os.path.join(td, "motifs.fsa")
    with open(mot, "w") as f:
        f.write(CLI_FILES[0])
    inputs = _inputs(td, msa, rows)
    outs = {}
    for tag, exe in (("reference", REF_BIN), ("driver", BIN)):
        phy, num = os.path.join(td, tag + ".phy"), os.path.join(td, tag + ".num")
        p = _run([exe, "dist", "-f", flag, "-y", mot, "-P", proxi, "-W", "1000", "-t", "2", "-o", phy, "-n", num] + inputs, td)
        assert p.returncode == 0, p.stderr[-2000:]
        outs[tag] = (open(phy).read(), open(num).read(), p.stderr)
    assert outs["driver"] == outs["reference"]
