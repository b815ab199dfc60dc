"""The .mat oracle against the reference BINARY on random count matrices (scripts/fuzz_cli.py's generator: every -d
method, -E / -C / -L gates with excluded samples and pairs without overlap, -W, -x, -l, -f, another template in
front): the oracle sums in the reference's order, so the Phylip text must match byte for byte.  Runs where the
reference was compiled."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import fuzz_cli  # noqa: E402
import oracle  # noqa: E402
import test_mat_oracle_golden as G  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.exists(fuzz_cli.REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")


@pytest.mark.parametrize("idx", range(150))
def test_mat_oracle_prints_what_the_reference_binary_prints(built, tmp_path, monkeypatch, idx):
    case = fuzz_cli.make_mat_case(11, idx)
    cmd, ref = fuzz_cli.run_mat(case, fuzz_cli.REF_BIN, str(tmp_path))        # (the reference side always runs -t 1)
    if ref["rc"] < 0:
        pytest.skip("the reference binary died or hung on this input (rc %d)" % ref["rc"])
    assert ref["rc"] == 0, ref["stderr"]
    args = list(case["args"])
    alpha = 0.05
    if "-l" in args:
        k = args.index("-l")
        alpha = float(args[k + 1])
        del args[k:k + 2]
    o = G.mat_args(args)
    real = oracle.mat_matrix
    monkeypatch.setattr(oracle, "mat_matrix", lambda *a, **kw: real(*a, alpha=alpha, **kw))
    names = ["%c.mat" % (ord("a") + k) + (".gz" if case["mode"] == "mat_gz" else "") for k in range(case["n"])]
    phy, num, err = G.expected_text(names, case["texts"], "tmpl", o)
    # the threaded loop names the ROW sample of a pair without overlap by its compact row number (ltdmatrixthrd.c:320)
    kept = [ln.split("\t")[0].strip() for ln in phy.split("\n")[2 if o["flag"] & 4 else 1:] if ln]
    fixed = []
    for ln in err.splitlines():
        if ln.startswith("No sufficient overlap between samples:"):
            head, row, col = ln.split("\t")
            ln = "\t".join((head, names[kept.index(row)], col))
        fixed.append(ln)
    assert sorted(fixed + [""]) == sorted(ref["stderr"].decode().split("\n"))
    assert phy.encode() == (ref["phy"] or b"")
    assert num.encode() == (ref["num"] or b"")
