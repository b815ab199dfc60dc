"""`ccphylo-b200 trim` (host/trim_main.c + csrc/k_trim.cu behind ccg_trim_*) against the reference
binary's `ccphylo trim` (trim.c:77-260) on the same files: trimmed FASTA and stderr byte for byte, for every output flag,
-P, -y, -L / -C exclusions, multi-file (-r) and multi-record inputs, plain and gzip."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from ccphylo_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.environ.get("CCPHYLO_TEST_BIN") or os.path.join(ROOT, "ccphylo_b200", "bin", "ccphylo-b200")   # (tests/csrc/mock_ccg.c gives a CPU driver)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ccphylo")
needs_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")


def _run(cmd, cwd):
    return subprocess.run(cmd, capture_output=True, cwd=cwd, timeout=300)


def _rows(n, length, seed, soft_ok=True, ambiguity=True):
    """alignment rows with everything trim's alphabet knows: SNP clusters, N runs, gaps, ambiguity letters and (where the
    reference handles them without reading past its letter table) soft-masked stretches"""
    rng = np.random.default_rng(seed)
    rows = synth.make_ascii(n, length, seed=seed, snp=0.02, nrun=0.01, lower=0.0, gap=0.01)
    for i in range(n):
        if ambiguity:
            at = rng.random(length) < 0.01
            rows[i, at] = rng.choice(np.frombuffer(b"RYSWKMBDHV", dtype=np.uint8), size=int(at.sum()))
        if soft_ok:
            for _ in range(max(1, length // 400)):
                s = int(rng.integers(0, length))
                e = min(length, s + int(rng.integers(1, 40)))
                seg = rows[i, s:e]
                letters = (seg >= ord("A")) & (seg <= ord("Z")) & (seg != ord("N"))
                seg[letters] |= 0x20
    return rows


def _write_files(td, rows, header="ref", gz=False):
    files = []
    for i in range(rows.shape[0]):
        fp = os.path.join(td, f"s{i:02d}.fsa")
        synth.write_fasta(fp, rows[i], header=header, width=60)
        if gz:
            with open(fp, "rb") as f, gzip.open(fp + ".gz", "wb") as g:
                g.write(f.read())
            fp += ".gz"
        files.append(fp)
    return files


def _write_msa(td, rows, names=None):
    path = os.path.join(td, "aln.fsa")
    with open(path, "wb") as f:
        for i in range(rows.shape[0]):
            f.write(b">" + (names[i] if names else b"rec%d" % i) + b"\n")
            for s0 in range(0, rows.shape[1], 70):
                f.write(rows[i, s0:s0 + 70].tobytes() + b"\n")
    return path


def _both(td, args):
    outs = {}
    for tag, exe in (("reference", REF_BIN), ("driver", BIN)):
        out = os.path.join(td, tag + ".out")
        p = _run([exe, "trim", "-o", out] + args, td)
        assert p.returncode == 0, p.stderr[-2000:]
        outs[tag] = (open(out, "rb").read(), p.stderr)
    return outs


# soft-masked input next to an unknown reference base, or under -f 8 without -f 1 / -f 4, makes the reference index past
# its 16-letter table (trim.c:40,50,60-64): those inputs carry no lower-case letters
@needs_ref
@pytest.mark.parametrize("flag", [0, 1, 4, 5, 12, 13, 16, 17, 20, 32, 33, 36, 48, 9])
@pytest.mark.parametrize("proxi", [0, 7, 300])
def test_shared_mask_files_against_the_reference_binary(built, tmp_path, flag, proxi):
    td = str(tmp_path)
    n, length = 7, 6000 + 13
    soft_ok = not (flag & 8) or bool(flag & 5)
    rows = _rows(n, length, seed=flag * 5 + proxi, soft_ok=False)
    if soft_ok:
        # lower-case stretches only where the reference sample (the first file) has a known base
        soft = _rows(n, length, seed=flag * 5 + proxi, soft_ok=True)
        known_in_ref = (rows[0] != ord("N")) & (rows[0] != ord("n"))
        rows = np.where(known_in_ref[None, :], soft, rows)
        rows[0] = soft[0]
    rows[3, 40:] = ord("N")                       # a sample below the coverage threshold
    files = _write_files(td, rows, gz=(flag == 4))
    # (with -P 300 little is left of a sample: a low coverage threshold keeps the reference from running out of samples,
    # which it answers with a null-pointer read, trim.c:237)
    args = ["-f", str(flag), "-r", "ref"] + (["-P", str(proxi), "-C", "0.5"] if proxi else []) + ["-i"] + files
    outs = _both(td, args)
    assert outs["reference"][0].count(b">") >= n - 2
    assert outs["driver"] == outs["reference"]


@needs_ref
@pytest.mark.parametrize("flag", [2, 3, 6, 7, 34])
@pytest.mark.parametrize("proxi", [0, 11])
def test_pairwise_flag_against_the_reference_binary(built, tmp_path, flag, proxi):
    td = str(tmp_path)
    rows = _rows(6, 4000 + 31, seed=flag + proxi)
    files = _write_files(td, rows)
    args = ["-f", str(flag), "-r", "ref", "-C", "10"] + (["-P", str(proxi)] if proxi else []) + ["-i"] + files
    outs = _both(td, args)
    assert outs["reference"][0].count(b">") == 6
    assert outs["driver"] == outs["reference"]


@needs_ref
@pytest.mark.parametrize("flag,proxi", [(0, 0), (1, 5), (4, 0), (16, 0), (20, 9), (2, 0)])
def test_multi_record_input_against_the_reference_binary(built, tmp_path, flag, proxi):
    td = str(tmp_path)
    rows = _rows(9, 3000 + 7, seed=flag + 3 * proxi + 1, soft_ok=False)
    path = _write_msa(td, rows, names=[b"rec %d extra words" % i for i in range(9)])
    args = ["-f", str(flag)] + (["-P", str(proxi)] if proxi else []) + ["-i", path]
    outs = _both(td, args)
    assert outs["reference"][0].count(b">") == 9
    assert outs["driver"] == outs["reference"]


@needs_ref
def test_multi_record_input_with_a_dropped_record(built, tmp_path):
    # a record below the threshold in the middle: the reference's name table and slot array go out of step (see
    # host/trim_main.c); the driver prints what the reference prints
    td = str(tmp_path)
    rows = _rows(7, 2000, seed=77, soft_ok=False)
    rows[2, 100:] = ord("N")
    path = _write_msa(td, rows)
    outs = _both(td, ["-i", path])
    assert outs["driver"] == outs["reference"]


@needs_ref
@pytest.mark.parametrize("flag", [0, 4, 2, 6])
def test_motif_masking_against_the_reference_binary(built, tmp_path, flag):
    td = str(tmp_path)
    rows = _rows(5, 5000 + 3, seed=flag + 200, soft_ok=False, ambiguity=(flag & 4) == 0)
    # plant dam / dcm sites
    for i in range(5):
        for s in range(50 + 7 * i, 4900, 331):
            rows[i, s:s + 4] = np.frombuffer(b"GATC", dtype=np.uint8)
            rows[i, s + 100:s + 105] = np.frombuffer(b"CCAGG" if (s // 331) % 2 else b"CCTGG", dtype=np.uint8)
    files = _write_files(td, rows)
    motifs = os.path.join(td, "motifs.fsa")
    with open(motifs, "w") as f:
        f.write(">dam\ngAtc\n>dcm\ncCwgg\n")
    args = ["-f", str(flag), "-y", motifs, "-r", "ref", "-i"] + files
    outs = _both(td, args)
    plain = _both(td, ["-f", str(flag), "-r", "ref", "-i"] + files)
    assert outs["driver"] == outs["reference"]
    assert outs["reference"][0] != plain["reference"][0], "the motifs never matched"


def test_trim_usage(built, tmp_path):
    p = _run([BIN, "trim", "-h"], str(tmp_path))
    assert p.returncode == 0 and p.stdout.startswith(b"#ccphylo-b200 trim")
    p = _run([BIN, "trim", "-F"], str(tmp_path))
    assert b"Pairwise comparison" in p.stdout
    a = tmp_path / "a.txt"
    a.write_text("no fasta here\n")
    p = _run([BIN, "trim", "-i", str(a)], str(tmp_path))
    assert p.returncode == 1 and p.stderr == b'"%s" is not fasta.\n' % str(a).encode()
