"""GPU parity of -V / --nucleotide_variations: ccg_list_variants against the oracle's
restatement of fsacmpairint / fsacmprint (fsacmp.c:646-737, pinned to the reference in
tests/test_oracle_vs_reference.py), and the host driver's -V output against the reference binary's, line for line."""
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


@pytest.fixture(scope="module")
def ctx(built):
    c = api.Context()
    yield c
    c.close()


@pytest.mark.parametrize("n,length", [(2, 1), (3, 33), (9, 128), (40, 4099), (130, 3000), (700, 700)])
def test_pair_mode_lists(ctx, n, length):
    codes = synth.make_codes(n, length, seed=n * 7 + length, snp=0.03, nrun=0.05, lower=0.02, gap=0.01)
    if n > 4:
        codes[3, :] = 4
    seqs, masks, inc = oracle.encode_samples(codes)
    include = (inc > 0).astype(np.uint8) if length > 1 else np.ones(n, np.uint8)
    ctx.set_problem(n, length, pair=True)
    ctx.put_samples_packed(seqs, masks)
    got = ctx.list_variants(pair=True, include=include)
    want = []
    for i in range(1, n):
        for j in range(i):
            if include[i] and include[j]:
                v = oracle.list_variants(seqs[i], seqs[j], masks[i] & masks[j], length)
                if v:
                    want.append(((i, j), v))
    assert got == want
    # the counts of the run on the same store agree with the lists
    D, N, dn = ctx.run_pair(include, min_length=0, min_cov=0.0)
    mism, _ = ctx.raw_counts(dn)
    assert int(mism.sum()) == sum(len(v) for _, v in got)


@pytest.mark.parametrize("proxi", [1, 3, 33, 200, 100000])
@pytest.mark.parametrize("n,length", [(2, 1), (3, 33), (9, 129), (40, 4099), (150, 3000)])
def test_pair_mode_lists_with_proximity(ctx, n, length, proxi):
    """-V with -P: fsacmpairint walks maskProxi's per-pair mask (fsacmpthrd.c:410-414); the labels depend on every word
    of that mask"""
    codes = synth.make_codes(n, length, seed=n * 11 + length + proxi, snp=0.04, nrun=0.03, lower=0.02, gap=0.01)
    seqs, masks, inc = oracle.encode_samples(codes, proxi=proxi)
    include = (inc > 0).astype(np.uint8) if length > 1 else np.ones(n, np.uint8)
    ctx.set_proximity(proxi)
    try:
        ctx.set_problem(n, length, pair=True)
        ctx.put_samples_packed(seqs, masks)
        got = ctx.list_variants(pair=True, include=include)
        want, cleared = [], 0
        for i in range(1, n):
            for j in range(i):
                if include[i] and include[j]:
                    pm = oracle.pair_mask_proxi(seqs[i], seqs[j], masks[i], masks[j], length, proxi)
                    cleared += int(np.bitwise_count(pm ^ (masks[i] & masks[j])).sum())
                    v = oracle.list_variants(seqs[i], seqs[j], pm, length)
                    if v:
                        want.append(((i, j), v))
        assert got == want
        assert cleared > 0 or length < 64
        # the counts of the run on the same store agree with the lists
        D, N, dn = ctx.run_pair(include, min_length=0, min_cov=0.0)
        mism, _ = ctx.raw_counts(dn)
        assert int(mism.sum()) == sum(len(v) for _, v in got)
    finally:
        ctx.set_proximity(0)


@pytest.mark.parametrize("n,length", [(4, 97), (30, 5003), (260, 2000)])
def test_shared_mask_lists(ctx, n, length):
    rare = n > 100           # many samples: keep the shared mask from running empty
    codes = synth.make_codes(n, length, seed=n + length, snp=0.03, nrun=0.0002 if rare else 0.004,
                             lower=0.0002 if rare else 0.02, gap=0.0001 if rare else 0.01)
    seqs, masks, inc = oracle.encode_samples(codes)
    include = np.ones(n, np.uint8)
    if n > 5:
        include[2] = 0
    gmask = oracle.global_mask(codes, include)
    ctx.set_problem(n, length, pair=True)
    for i in range(n):
        ctx.put_sample_codes(i, codes[i])
        ctx.sync()
    ginc = ctx.build_global_mask(include)
    got = ctx.list_variants(pair=False, include=include)
    want = []
    for i in range(1, n):
        for j in range(i):
            if include[i] and include[j]:
                v = oracle.list_variants(seqs[i], seqs[j], gmask, length)
                if v:
                    want.append(((i, j), v))
    assert got == want and len(want) > 0
    # the run that follows applies the mask and gives the distances of the lists
    D, dn, ginc2 = ctx.run_global(include)
    Do, dno, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length)
    assert ginc == ginc2 == ginco and np.array_equal(D, Do)
    assert int(D.sum()) == sum(len(v) for _, v in got)
    with pytest.raises(api.CcgError):
        ctx.list_variants(pair=False, include=include)       # the planes are masked now


def _run(cmd, cwd):
    return subprocess.run(cmd, capture_output=True, text=True, cwd=cwd, timeout=300)


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
@pytest.mark.parametrize("flag,proxi", [("3", 0), ("1", 0), ("11", 0), ("3", 5), ("3", 300), ("35", 40), ("1", 12), ("11", 9)])
@pytest.mark.parametrize("msa", [False, True], ids=["files", "msa"])
def test_cli_variant_file_against_the_reference_binary(built, tmp_path, msa, flag, proxi):
    td = str(tmp_path)
    n, length = 12, 9000 + 3
    rows = synth.make_ascii(n, length, seed=int(flag) + 40 + proxi, snp=0.01, nrun=0.01)
    if int(flag) & 2:
        rows[4, 30:] = ord("N")          # excluded sample: the listed sample numbers are file indices (files) / kept records (msa)
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
        phy, num, var = (os.path.join(td, tag + e) for e in (".phy", ".num", ".var"))
        prox = ["-P", str(proxi)] if proxi else []
        p = _run([exe, "dist", "-f", flag, "-t", "1", "-V", var, "-o", phy, "-n", num] + prox + inputs, td)
        assert p.returncode == 0, p.stderr[-2000:]
        outs[tag] = (open(var).read(), open(phy).read(), p.stderr)
        # variants and matrix into the same file: the lists come first
        both = os.path.join(td, tag + ".both")
        p = _run([exe, "dist", "-f", flag, "-t", "1", "-V", both, "-o", both] + prox + inputs, td)
        # (the reference closes that FILE twice and aborts after everything has been written, dist.c:293-298)
        assert p.returncode == 0 or exe == REF_BIN, p.stderr[-2000:]
        outs[tag] += (open(both).read(),)
    assert outs["reference"][0].count("\n") > 100
    assert outs["driver"] == outs["reference"]


# (with a CPU driver from tests/csrc/mock_ccg.c there is no device for the bound reference either)
REF_GPU = os.environ.get("CCPHYLO_TEST_REF_GPU") or os.path.join(ROOT, "oracle", "_ref", "ccphylo_gpu" if not os.environ.get("CCPHYLO_TEST_BIN") else "ccphylo_gpu.absent")


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(REF_GPU)), reason="oracle/_ref was not built (needs /root/reference)")
@pytest.mark.parametrize("flag,proxi", [("3", 0), ("11", 0), ("1", 0), ("3", 6), ("35", 150), ("1", 6)])
def test_reference_bound_to_the_library_lists_variants_on_the_device(built, tmp_path, flag, proxi):
    # the UNMODIFIED reference linked against libccphylo_gpu.so through integration/fsacmpgpu.c: -V in pair mode comes
    # from ccg_list_variants (in shared-mask mode the stub leaves it on the reference's own code); same bytes either way
    td = str(tmp_path)
    n, length = 10, 7000 + 5
    rows = synth.make_ascii(n, length, seed=int(flag) + 90 + proxi, snp=0.01, nrun=0.01)
    files = []
    for i in range(n):
        fp = os.path.join(td, f"s{i:02d}.fsa")
        synth.write_fasta(fp, rows[i], header="ref", width=60)
        files.append(fp)
    outs = {}
    for tag, exe in (("reference", REF_BIN), ("bound", REF_GPU)):
        phy, num, var = (os.path.join(td, tag + e) for e in (".phy", ".num", ".var"))
        p = _run([exe, "dist", "-f", flag, "-t", "1", "-V", var, "-o", phy, "-n", num] + (["-P", str(proxi)] if proxi else [])
                 + ["-r", "ref", "-i"] + files, td)
        assert p.returncode == 0, p.stderr[-2000:]
        outs[tag] = (open(var).read(), open(phy).read(), open(num).read(), p.stderr)
    assert outs["reference"][0].count("\n") > 50
    assert outs["bound"] == outs["reference"]
