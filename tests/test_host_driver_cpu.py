"""The HOST driver end to end without a GPU: ccphylo_b200/host/*.c linked with tests/csrc/mock_ccg.c (the C-ABI answered
by the oracle on the CPU -- test infrastructure, see its header) and compiled with -fsanitize=address,undefined, then
driven by scripts/fuzz_cli.py's random command lines against the reference binary: option scanning, FASTA / MSA / gz /
.mat parsing, the parser pool, gates and messages, -P / -y / -V plumbing, the Phylip writer and the file-backed matrices
must print the reference's bytes and trip no sanitizer.  (What the mock does not stand in for -- shared-mask
mode with -P and -y together -- is reported as unsupported and skipped; that runs on the GPU box, tests/test_cli_fuzz_gpu.py.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import fuzz_cli  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.exists(fuzz_cli.REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
HOST = os.path.join(ROOT, "ccphylo_b200", "host")
SRCS = ["dist_main.c", "trim_main.c", "dist_mat.c", "cmdline.c", "phy_writer.c", "fsa_reader.c", "mat_reader.c", "ordered_pool.c",
        "motifs.c", "phy_update.c"]


@pytest.fixture(scope="module")
def mock_driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("mock") / "ccphylo-b200-mock")
    cmd = ["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=gnu99", "-Wall",
           "-Wno-unused-parameter", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"), "-I", HOST, "-o", exe]
    cmd += [os.path.join(HOST, s) for s in SRCS]
    cmd += [os.path.join(ROOT, "tests", "csrc", "mock_ccg.c"), os.path.join(ROOT, "oracle", "fsa_oracle.c"),
            os.path.join(ROOT, "oracle", "mat_oracle.c"), "-lz", "-lpthread", "-lm"]
    subprocess.run(cmd, check=True)
    old, fuzz_cli.BIN = fuzz_cli.BIN, exe
    os.environ["ASAN_OPTIONS"] = "detect_leaks=0"          # the driver exits without freeing its tables, as the reference does
    yield exe
    fuzz_cli.BIN = old


def test_fasta_command_lines_under_the_sanitizers(mock_driver, tmp_path):
    ok = 0
    for idx in range(160):
        r = fuzz_cli.check(fuzz_cli.make_case(21, idx), str(tmp_path), False)
        assert r["verdict"] in ("ok", "ref_crash", "known_divergence_3", "known_trim_soft_letters", "unsupported"), r
        ok += r["verdict"] == "ok"
    assert ok >= 120


def test_mat_command_lines_under_the_sanitizers(mock_driver, tmp_path):
    ok = 0
    for idx in range(160):
        r = fuzz_cli.check_mat(fuzz_cli.make_mat_case(21, idx), str(tmp_path), False)
        assert r["verdict"] in ("ok", "ref_crash"), r
        ok += r["verdict"] == "ok"
    assert ok >= 150


def test_union_command_lines_under_the_sanitizers(mock_driver, tmp_path):
    """count matrices behind a union file: one block per template row over its own subset of the files; file 0's gate line
    comes when the first later sample that passes streams it (ltdmatrix.c:157-158), not at all if none does"""
    for idx in range(160):
        r = fuzz_cli.check_mat(fuzz_cli.make_union_case(21, idx), str(tmp_path), False)
        assert r["verdict"] in ("ok", "ref_crash"), r
    for idx in (111, 120, 313, 523):                      # the cases that showed it
        r = fuzz_cli.check_mat(fuzz_cli.make_union_case(1, idx), str(tmp_path), False)
        assert r["verdict"] == "ok", r


def test_add_row_command_lines_under_the_sanitizers(mock_driver, tmp_path):
    """-a: the reference builds a matrix, then one more FASTA sample is added to copies of it by the reference and by the
    driver (Phylip re-reading, name lookup, the row's gates, -P, -V appended to the listing)"""
    ok = 0
    for idx in range(120):
        r = fuzz_cli.check_add(21, idx, str(tmp_path), False)
        assert r["verdict"] in ("ok", "ref_crash", "no_matrix"), r
        ok += r["verdict"] == "ok"
    assert ok >= 90


def test_add_mat_row_command_lines_under_the_sanitizers(mock_driver, tmp_path):
    """-a on .mat input: one more count matrix against a matrix the reference built"""
    ok = 0
    for idx in range(100):
        r = fuzz_cli.check_add_mat(21, idx, str(tmp_path), False)
        assert r["verdict"] in ("ok", "ref_crash", "no_matrix"), r
        ok += r["verdict"] == "ok"
    assert ok >= 60


def test_no_overlap_line_only_when_a_gate_can_fail(mock_driver, tmp_path):
    """-L 0 -C 0: a pair without one comparable position passes the gate (cell 0/0 = -nan, N = 0) and the reference prints no
    'No sufficient overlap' line (found by this sweep: seed 5, cases 26 and 993)"""
    for idx in (26, 993):
        r = fuzz_cli.check_mat(fuzz_cli.make_mat_case(5, idx), str(tmp_path), False)
        assert r["verdict"] == "ok", r


def _bound_reference(mock_driver):
    """the UNMODIFIED reference with fsaCmpThreadOut bound through integration/fsacmpgpu.c (as oracle/Makefile's ref_gpu does)
    -- to the mock instead of libccphylo_gpu.so"""
    import glob
    exe = os.path.join(os.path.dirname(mock_driver), "ccphylo_gpu_mock")
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    objs = [o for o in sorted(glob.glob(os.path.join(ref_dir, "obj", "*.o"))) if os.path.basename(o) != "cdist.o"]
    cdist = os.path.join(os.path.dirname(mock_driver), "cdist_gpu.o")
    subprocess.run(["gcc", "-w", "-O3", "-std=c99", "-DfsaCmpThreadOut=fsaCmpGpuOut", "-c", "-o", cdist, "/root/reference/cdist.c"], check=True)
    subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=gnu99", "-w",
                    "-I", "/root/reference", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"), "-o", exe,
                    "/root/reference/main.c"] + objs + [cdist, os.path.join(ROOT, "integration", "fsacmpgpu.c"),
                    os.path.join(ROOT, "tests", "csrc", "mock_ccg.c"), os.path.join(ROOT, "oracle", "fsa_oracle.c"),
                    os.path.join(ROOT, "oracle", "mat_oracle.c"), "-lm", "-lpthread", "-lz"], check=True)
    return exe


def test_the_gpu_box_cli_tests_on_the_cpu_driver(mock_driver):
    """the command-line tests of the GPU suite (golden FASTA / .mat / union text, MSA plain and gz, -P, -y, -V, -a, -H,
    long options, the union | dist | tree pipe) with the CPU driver in the place of ccphylo-b200: CCPHYLO_TEST_BIN"""
    env = dict(os.environ, CCPHYLO_TEST_BIN=mock_driver, ASAN_OPTIONS="detect_leaks=0")
    if os.path.exists("/root/reference/cdist.c") and os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "obj")):
        env["CCPHYLO_TEST_REF_GPU"] = _bound_reference(mock_driver)      # the integration stub under the sanitizers as well
    keep = ("(golden or cli or against_the_reference_binary or pipe or msa or option or file_backed or gz_input or trim or record or bound) "
            "and not several_gpus and not config1 and not motifs_with_proximity")
    p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x", "-k", keep, "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "test_cli_gpu.py"), os.path.join(ROOT, "tests", "test_gpu_addrow.py"),
                        os.path.join(ROOT, "tests", "test_gpu_motifs.py"), os.path.join(ROOT, "tests", "test_gpu_variants.py"),
                        os.path.join(ROOT, "tests", "test_gpu_trim.py")],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:]
    assert " passed" in p.stdout and int(p.stdout.rsplit(" passed", 1)[0].split()[-1]) >= 200, p.stdout[-500:]
