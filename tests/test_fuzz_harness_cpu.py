"""scripts/fuzz_cli.py without a GPU: the generator's command lines are valid for the reference binary (it runs them
against itself, `--self`), cases are reproducible from (seed, index), and the .mat cell comparison accepts what the
documented tolerance allows and nothing else."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import fuzz_cli  # noqa: E402

needs_ref = pytest.mark.skipif(not os.path.exists(fuzz_cli.REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")


@needs_ref
def test_reference_against_itself(tmp_path):
    ran = 0
    for idx in range(24):
        r = fuzz_cli.check(fuzz_cli.make_case(5, idx), str(tmp_path), True)
        assert r["verdict"] == "ok", r
        ran += r["rc"][0] == 0
    for idx in range(12):
        r = fuzz_cli.check_mat(fuzz_cli.make_mat_case(5, idx), str(tmp_path), True)
        assert r["verdict"] == "ok", r
        ran += r["rc"][0] == 0
    assert ran >= 30                      # the generator mostly produces runs that succeed, not usage errors


def test_cases_are_reproducible():
    a, b = fuzz_cli.make_case(3, 17), fuzz_cli.make_case(3, 17)
    assert a["args"] == b["args"] and a["flag"] == b["flag"] and all((x == y).all() for x, y in zip(a["rows"], b["rows"]))
    assert fuzz_cli.make_case(3, 18)["rows"][0].tobytes() != a["rows"][0].tobytes() or fuzz_cli.make_case(3, 18)["args"] != a["args"]
    m, k = fuzz_cli.make_mat_case(3, 4), fuzz_cli.make_mat_case(3, 4)
    assert m["texts"] == k["texts"] and m["args"] == k["args"]
    big = fuzz_cli.make_case(3, 17, big=True)
    assert big["tool"] == "dist" and 192 <= big["n"] <= 330 and 8192 <= big["length"] <= 20000 and not big["variants"]


def test_mat_cell_comparison():
    ref = b"         3\na.mat\nb.mat\t0.123456789\nc.mat\t-1\t1000.5\n"
    assert fuzz_cli.cells_close(ref, ref, 9)
    assert fuzz_cli.cells_close(ref, ref.replace(b"0.123456789", b"0.123456790"), 9)          # last printed digit
    assert fuzz_cli.cells_close(ref, ref.replace(b"1000.5", b"1000.5009"), 9)                 # 1e-6 relative
    assert not fuzz_cli.cells_close(ref, ref.replace(b"1000.5", b"1000.51"), 9)
    assert not fuzz_cli.cells_close(ref, ref.replace(b"-1", b"0"), 9)
    assert not fuzz_cli.cells_close(ref, ref.replace(b"b.mat", b"x.mat"), 9)
    assert not fuzz_cli.cells_close(ref, ref + b"d.mat\t1\t2\t3\n", 9)
    assert fuzz_cli.cells_close(None, None, 9) and not fuzz_cli.cells_close(ref, None, 9)
