"""Command lines that end before any device work -- usage errors, bad values, the -F / -D help texts -- printed by
`ccphylo-b200 dist` / `trim` and by the reference binary: same stdout, same stderr, same exit code.  (The -h text names
this driver and is not compared.)  Runs where the reference was compiled."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ccphylo_b200", "bin", "ccphylo-b200")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ccphylo")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")

DIST = ["-F", "-D", "--flag_help", "--distance_help", "-Z", "-pZq", "--bogus", "-f 3 -Q", "-x", "-f", "-W", "-t", "-L", "-P", "-o", "-n",
        "-y", "-V", "-S", "-T", "-a", "-x abc -i r1.fsa", "-C abc -i r1.fsa", "-C 150 -i r1.fsa", "-C -5 -i r1.fsa", "-d ln -i r1.fsa",
        "-d lx -i r1.fsa", "-d bogus -i r1.fsa", "-s abc -i r1.fsa", "-ps abc -i r1.fsa", "--short_precision abc -i r1.fsa",
        "--short_precision=abc -i r1.fsa", "-b abc -i r1.fsa", "--byte_precision abc -i r1.fsa", "-s 0 -i r1.fsa", "-b 0 -i r1.fsa",
        "-s -1 -i r1.fsa", "-l -1 -i r1.fsa", "-t abc -i r1.fsa", "-W abc -i r1.fsa", "-f abc -i r1.fsa", "-S ab -i r1.fsa",
        "-S -i r1.fsa", "-L abc -i r1.fsa", "-P abc -i r1.fsa", "-E abc -i r1.fsa", "--min_cov=abc -i r1.fsa", "-i nothere.fsa",
        "-i r1.fsa r2.fsa", "r1.fsa r2.fsa", "-i r1.fsa r2.fsa -r", "-r ref -i nothere1 nothere2"]
TRIM = ["-F", "-Z", "-pZq", "--bogus", "-x", "-t 2", "-C abc -i r1.fsa", "-f abc -i r1.fsa", "-L abc -i r1.fsa", "-P abc -i r1.fsa"]


def _both(tool, args, cwd):
    out = []
    for exe in (REF_BIN, BIN):
        p = subprocess.run([exe, tool] + args.split(), capture_output=True, cwd=cwd, stdin=subprocess.DEVNULL, timeout=60)
        out.append((p.returncode, p.stdout, p.stderr))
    return out


@pytest.fixture()
def files(tmp_path):
    (tmp_path / "r1.fsa").write_text(">ref\nACGT\n")
    (tmp_path / "r2.fsa").write_text(">ref\nACGA\n")
    return str(tmp_path)


@pytest.mark.parametrize("args", DIST)
def test_dist_usage_errors_and_help_texts(built, files, args):
    ref, drv = _both("dist", args, files)
    assert ref[0] >= 0 and b"CUDA" not in drv[2]
    assert drv == ref


@pytest.mark.parametrize("args", TRIM)
def test_trim_usage_errors_and_help_texts(built, files, args):
    ref, drv = _both("trim", args, files)
    assert ref[0] >= 0 and b"CUDA" not in drv[2]
    assert drv == ref
