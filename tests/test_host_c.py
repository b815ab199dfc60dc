"""Host-side C of the driver, on CPU: the threaded Phylip writer prints exactly what a per-cell fprintf loop in the
reference's format prints (all four cell types, strict and relaxed names, comment line), and the driver's option
scanner accepts the reference's command-line dialect."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ccphylo_b200", "host")
BIN = os.path.join(ROOT, "ccphylo_b200", "bin", "ccphylo-b200")


def test_threaded_phylip_writer_matches_per_cell_fprintf(tmp_path):
    exe = str(tmp_path / "phy_writer_test")
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "csrc", "phy_writer_test.c"),
                    os.path.join(HOST, "phy_writer.c"), "-lpthread", "-lm"], check=True)
    for n in ("9", "1200"):
        p = subprocess.run([exe, n, str(tmp_path)], capture_output=True, text=True)
        assert p.returncode == 0 and p.stdout.strip() == "OK", p.stderr


def test_proximity_arithmetic_of_the_kernels_on_the_host(built, tmp_path):
    """ccphylo_b200/csrc/proxi_core.h (what k_pairdist_proxi / k_sample_proxi execute per word) compiled for the
    host and compared with the oracle's maskProxi / getIncPos* restatement"""
    exe = str(tmp_path / "proxi_core_test")
    odir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-I", odir, "-I", os.path.join(ROOT, "ccphylo_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "csrc", "proxi_core_test.cpp"), "-L", odir, "-loracle",
                    "-Wl,-rpath," + odir], check=True)
    for seed in ("1", "2"):
        p = subprocess.run([exe, seed], capture_output=True, text=True)
        assert p.returncode == 0 and p.stdout.startswith("OK "), p.stdout + p.stderr


def test_motif_file_parser_matches_the_oracle(built, tmp_path):
    """host/motifs.c (the driver's -y parser: IUPAC sets, methylation sites, the reference's reverse complement of
    odd-length motifs, its as-built padding of ambiguity letters and CCPHYLO_MOTIF_STRICT) against oracle.parse_motifs,
    which tests/test_oracle_vs_reference.py pins to the reference's getMethMotifs + maskMotifs"""
    import oracle
    exe = str(tmp_path / "motifs_test")
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-Wall", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "csrc", "motifs_test.c"),
                    os.path.join(HOST, "motifs.c"), os.path.join(HOST, "fsa_reader.c"), "-lz"], check=True)
    files = [">dam\ngAtc\n", ">dam\ngAtc\n>dcm\ncCwgg\n>x\nrgATcnny\n", "gatC\n>multi line\ncC\nwg\ng\n>odd chars\nGA-NT.C\r\n",
             ">long\nacgtacgtAcgtacgtacgtacgTacgtacgt\n>three\ngAn\n>iupac\nRYSWKMBDHVN\n>u\nUu\n>empty\n>x\nXx\n"]
    for k, text in enumerate(files):
        path = str(tmp_path / f"m{k}.fsa")
        with open(path, "w") as f:
            f.write(text)
        for strict in (False, True):
            env = dict(os.environ)
            env.pop("CCPHYLO_MOTIF_STRICT", None)
            if strict:
                env["CCPHYLO_MOTIF_STRICT"] = "1"
            p = subprocess.run([exe, path], capture_output=True, text=True, env=env)
            assert p.returncode == 0, p.stderr
            got = [[int(x) for x in line.split()] for line in p.stdout.splitlines()]
            assert got == oracle.parse_motifs(text, as_built=not strict), (k, strict)
    path = str(tmp_path / "toolong.fsa")
    with open(path, "w") as f:
        f.write(">m\n" + "acgt" * 8 + "a\n")
    p = subprocess.run([exe, path], capture_output=True, text=True)
    assert p.returncode == 1 and "more than 32 positions" in p.stderr
    p = subprocess.run([exe, str(tmp_path / "missing.fsa")], capture_output=True, text=True)
    assert p.returncode == 1 and "Filename:" in p.stderr


def test_option_scanner_dialect(built, tmp_path):
    def run(*args):
        return subprocess.run([BIN, "dist"] + list(args), capture_output=True, text=True, cwd=str(tmp_path))

    assert run("-h").stdout.startswith("#ccphylo-b200 dist")
    assert "Relaxed Phylip" in run("-F").stdout and "Relaxed Phylip" in run("--flag_help").stdout
    assert "# cos:" in run("-D").stdout
    assert run("--bogus").stderr == 'Unknown argument or option: "--bogus"\n'
    assert run("-Z").stderr == 'Unknown argument or option: "-Z"\n'
    assert run("-x").stderr == "Missing argument at x.\n"
    assert run("-x", "abc").stderr == "Invalid value parsed at x.\n"
    assert run("--print_precision=abc").stderr == "Invalid value parsed at print_precision.\n"
    assert run("-d", "foo").stderr == 'Invalid value parsed at "-d".\n'
    assert run("-d", "l2x").stderr == 'Invalid value parsed at "-d ln".\n'
    assert run("-C", "150").stderr == 'Invalid value parsed at "--min_cov".\n'
    assert run("-s", "0").stderr == 'Invalid value parsed at "--short_precision".\n'
    # bundled flags, attached values, `-s` without a value before another option: all parse; with two FASTA
    # files the run then stops at the device (this container has none) or at the refused option
    a = tmp_path / "a.fsa"
    a.write_text(">ref\nACGT\n")
    p = run("-pf3", "-s", "-W", "100", "-rref", "-i", str(a), str(a), "-P", "2", "-y", "m.txt")
    assert p.returncode == 1 and "-y / --methylation_motifs together with -P / --proximity is not available on the GPU path" in p.stderr
    p = run("-r", "ref", "-a", "x", "-V", "v.txt", "-P", "2", str(a), str(a))
    assert p.returncode == 1 and "-V / --nucleotide_variations together with -P / --proximity" in p.stderr
    # -a reads the existing matrix before it needs the device: a multi-matrix file is refused as the reference does
    m = tmp_path / "m.phy"
    m.write_text("%10d\na.fsa\nb.fsa\t1\n%10d\na.fsa\nb.fsa\t2\n" % (2, 2))
    p = run("-r", "ref", "-a", str(a), "-i", str(a), "-o", str(m))
    assert p.returncode == 1 and p.stderr == "Cannot update a multi distance phylip file.\n"


def test_integration_stub_is_the_documented_one_and_compiles(tmp_path):
    import re
    import pytest
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```c\n(.*?)```", md, flags=re.S)
    stub = open(os.path.join(ROOT, "integration", "fsacmpgpu.c")).read()
    assert any(b == stub for b in blocks), "integration/fsacmpgpu.c and the stub printed in INTEGRATION.md differ"
    if not os.path.exists("/root/reference/fsacmpthrd.h"):
        pytest.skip("the reference headers are only present in the authoring container")
    subprocess.run(["gcc", "-std=gnu99", "-Wall", "-Werror", "-c", "-I", "/root/reference", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "integration", "fsacmpgpu.c"), "-o", str(tmp_path / "stub.o")], check=True)
