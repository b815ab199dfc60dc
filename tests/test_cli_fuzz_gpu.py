"""A fixed selection of scripts/fuzz_cli.py's random command lines, replayed on the GPU box: the reference binary
(oracle/_ref/ccphylo, host cores) and `ccphylo-b200` must leave the same bytes (FASTA inputs: .phy, .num, -V listing,
trim output, sorted stderr, return code) or the same cells within 1e-6 (.mat inputs).  The whole sweep (hundreds of
cases per seed) is run with the script itself; its summaries are under profiles/."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import fuzz_cli  # noqa: E402

needs_ref = pytest.mark.skipif(not os.path.exists(fuzz_cli.REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")

# (seed, index): a spread over dist / trim, files / MSA / gz, -P, -y, -V, cell types
FASTA = [(1, k) for k in (0, 2, 3, 7, 11, 18, 37, 47)] + [(2, k) for k in (3, 9)]
# seed 1 / 9, 12, 24, 42: an excluded file in front of a pair without sufficient overlap (the threaded loop names the row
# sample by its compact row number, ltdmatrixthrd.c:320)
MAT = [(1, k) for k in (0, 3, 9, 12, 24, 42)] + [(2, 1)]


@needs_ref
@pytest.mark.parametrize("seed,idx", FASTA, ids=["s%d_%d" % c for c in FASTA])
def test_fasta_command_lines(built, tmp_path, seed, idx):
    r = fuzz_cli.check(fuzz_cli.make_case(seed, idx), str(tmp_path), False)
    assert r["verdict"] in ("ok", "ref_crash", "known_divergence_3"), r


@needs_ref
@pytest.mark.parametrize("seed,idx", MAT, ids=["s%d_%d" % c for c in MAT])
def test_mat_command_lines(built, tmp_path, seed, idx):
    r = fuzz_cli.check_mat(fuzz_cli.make_mat_case(seed, idx), str(tmp_path), False)
    assert r["verdict"] in ("ok", "ref_crash"), r
