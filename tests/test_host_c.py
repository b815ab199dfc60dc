"""Host-side C of the driver, on CPU: the threaded Phylip writer prints exactly what a per-cell fprintf loop in the
reference's format prints (all four cell types, strict and relaxed names, comment line), and the driver's option
scanner accepts the reference's command-line dialect."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ccphylo_b200", "host")
# CCB_TEST_CFLAGS="-fsanitize=address,undefined -g": the same unit tests under the sanitizers
EXTRA = os.environ.get("CCB_TEST_CFLAGS", "").split()
BIN = os.path.join(ROOT, "ccphylo_b200", "bin", "ccphylo-b200")


def test_threaded_phylip_writer_matches_per_cell_fprintf(tmp_path):
    exe = str(tmp_path / "phy_writer_test")
    subprocess.run(["gcc", *EXTRA, "-O2", "-std=gnu99", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "csrc", "phy_writer_test.c"),
                    os.path.join(HOST, "phy_writer.c"), "-lpthread", "-lm"], check=True)
    for n in ("9", "1200"):
        p = subprocess.run([exe, n, str(tmp_path)], capture_output=True, text=True)
        assert p.returncode == 0 and p.stdout.strip() == "OK", p.stderr
    # the writer's own "%.*f" (exact 128-bit arithmetic, ties to even) against snprintf: 7 million values, precision 0 .. 18
    p = subprocess.run([exe, "fixed", "600000"], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.startswith("OK "), p.stdout + p.stderr


def test_proximity_arithmetic_of_the_kernels_on_the_host(built, tmp_path):
    """ccphylo_b200/csrc/proxi_core.h (what k_pairdist_proxi / k_sample_proxi execute per word) compiled for the
    host and compared with the oracle's maskProxi / getIncPos* restatement"""
    exe = str(tmp_path / "proxi_core_test")
    odir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", *EXTRA, "-O2", "-std=c++17", "-Wall", "-I", odir, "-I", os.path.join(ROOT, "ccphylo_b200", "csrc"), "-o", exe,
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
    subprocess.run(["gcc", *EXTRA, "-O2", "-std=gnu99", "-Wall", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "csrc", "motifs_test.c"),
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


def test_phylip_update_matches_the_reference(built, tmp_path):
    """host/phy_update.c (-a: names of the existing matrix, row append) against the reference's own getSizePhy /
    getFilenamesPhy / printphyUpdate called in-process: relaxed and strict names, a comment line (which both
    overwrite with the new count), quoted names, precision, a second matrix in the file"""
    import shutil
    import numpy as np
    import oracle
    if not oracle.have_ref():
        import pytest
        pytest.skip("oracle/_ref was not built (needs /root/reference)")
    exe = str(tmp_path / "phy_update_test")
    subprocess.run(["gcc", *EXTRA, "-O2", "-std=gnu99", "-Wall", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "csrc", "phy_update_test.c"),
                    os.path.join(HOST, "phy_update.c"), os.path.join(HOST, "fsa_reader.c"), "-lz"], check=True)
    relaxed = "%10d\nsample_a.fsa\ns_b\t12\nanother name.fsa\t3.500000000\t-1\n" % 3
    strict = "%10d\n%-10.10s\n%-10.10s\t12\n%-10.10s\t3\t4\n" % (3, "s0.fsa", "a_long_name_cut", "x")
    comment = "#tmpl\n" + relaxed
    two = relaxed + relaxed
    cases = {"relaxed": relaxed, "strict": strict, "comment": comment, "two": two, "one": "%10d\nonly\n" % 1}
    for tag, text in cases.items():
        path = str(tmp_path / (tag + ".phy"))
        with open(path, "w") as f:
            f.write(text)
        n_ref, names_ref = oracle.ref_phy_names(path, "dir/sub/")
        p = subprocess.run([exe, "names", path, "dir/sub/", "\t"], capture_output=True, text=True)
        assert p.returncode == 0
        lines = p.stdout.splitlines()
        if tag == "two":
            assert n_ref == -2 and lines == ["-1"] and p.stderr == "Cannot update a multi distance phylip file.\n"
            continue
        if tag == "one":
            assert n_ref == -1 and lines == ["0"] and p.stderr == "Malformatted phylip file, name on row: 1\n"
            continue
        assert lines[0] == "1" and n_ref == len(lines) - 1
        assert lines[1:] == names_ref, tag
        # append a row with both
        n = n_ref + 1
        row = np.array([0.0, 17.0, 2.123456789123, -1.0][: n - 1])
        for flag, precision, name in ((1, 9, "path/to/new sample.fsa"), (0, 4, "'quoted_and_long_name.fsa'"), (1, 2, '"q.fsa"')):
            a, b = str(tmp_path / "a.phy"), str(tmp_path / "b.phy")
            shutil.copy(path, a)
            shutil.copy(path, b)
            oracle.ref_phy_update(a, n, name, row, flag, precision)
            p = subprocess.run([exe, "append", b, str(n), name, str(flag), str(precision)] + [repr(float(x)) for x in row],
                               capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
            assert open(a, "rb").read() == open(b, "rb").read(), (tag, flag, precision, name)


def test_ordered_parse_pool(tmp_path):
    """host/ordered_pool.c: results in job order, slots not reused before release, look-ahead bounded by the window,
    more threads than window slots or jobs"""
    exe = str(tmp_path / "ordered_pool_test")
    subprocess.run(["gcc", *EXTRA, "-O2", "-std=gnu99", "-Wall", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "csrc", "ordered_pool_test.c"),
                    os.path.join(HOST, "ordered_pool.c"), "-lpthread"], check=True)
    for args in (("300", "8", "10"), ("50", "16", "3"), ("5", "8", "10"), ("100", "1", "1"), ("64", "4", "4")):
        p = subprocess.run([exe] + list(args), capture_output=True, text=True, timeout=60)
        assert p.returncode == 0 and p.stdout.strip() == "OK", (args, p.stdout)


def _readers_exe(tmp_path):
    exe = str(tmp_path / "readers_test")
    subprocess.run(["gcc", *EXTRA, "-O2", "-std=gnu99", "-Wall", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "csrc", "readers_test.c"),
                    os.path.join(HOST, "fsa_reader.c"), os.path.join(HOST, "mat_reader.c"), "-lz"], check=True)
    return exe


def test_fasta_reader_matches_the_oracle_translation(built, tmp_path):
    """host/fsa_reader.c (gzip-transparent buffer, header / sequence scanning of seqparse.c:128-248, the byte -> code
    table of fsacmp.c:32-91 for every -f variant) against oracle.translate, itself pinned to the reference's table"""
    import gzip
    import numpy as np
    import oracle
    exe = _readers_exe(tmp_path)
    rng = np.random.default_rng(4)
    alphabet = np.frombuffer(b"ACGTacgtNn-RYKMSWBDHVryUuXx*. 1\r", dtype=np.uint8)
    records = []
    for k in range(6):
        n = [0, 1, 59, 60, 61, 5000][k]
        seq = alphabet[rng.integers(0, len(alphabet), size=n)].tobytes().replace(b">", b"A")
        records.append((b"rec%d some text \t " % k if k % 2 else b"rec%d" % k, seq))
    text = b""
    for hdr, seq in records:
        text += b">" + hdr + b"\n"
        for s0 in range(0, len(seq), 60):
            text += seq[s0:s0 + 60] + b"\n"
    text += b">last header without sequence"
    plain = str(tmp_path / "a.fsa")
    with open(plain, "wb") as f:
        f.write(text)
    with gzip.open(plain + ".gz", "wb") as f:
        f.write(text)
    for flag in (1, 9, 33, 41):
        outs = []
        for path in (plain, plain + ".gz"):
            p = subprocess.run([exe, "fsa", str(flag), path], capture_output=True)
            assert p.returncode == 0
            outs.append(p.stdout)
        lines = outs[0].split(b"\n")
        assert lines[0] == b"first=62 plain=1" and outs[1].split(b"\n")[0] == b"first=62 plain=0"
        assert outs[0].split(b"\n")[1:] == outs[1].split(b"\n")[1:]
        k = 1
        offset = 0
        for hdr, seq in records:
            name, at = lines[k].rsplit(b" @", 1)
            assert name == b">" + hdr.rstrip() and int(at) == offset
            want = oracle.translate(seq.replace(b"\n", b""), flag)
            assert lines[k + 1] == bytes(want + ord("0")), (flag, hdr)
            k += 2
            offset += 1 + len(hdr) + 1 + len(seq) + (len(seq) + 59) // 60
        # a header that runs into the end of the file is no record (FileBuffgetFsaHeader returns 0, seqparse.c:128-160)
        assert lines[k:] == [b""]


def test_mat_reader_matches_a_plain_parse(built, tmp_path):
    """host/mat_reader.c against helpers.parse_mat (the storage order and totals of matparse.c:213-259): template
    selection among several, insertion rows dropped, gz input, a template that is absent"""
    import gzip
    import numpy as np
    import helpers
    exe = _readers_exe(tmp_path)
    rng = np.random.default_rng(7)
    blocks = []
    # "big": more than the reader's 1 MiB window, so rows are cut by the window's end (in the plain and in the gz file
    # at different rows), with 1 .. 5-digit counts; rows within 72 bytes of a window's end take the line-by-line parser
    for name, rows in (("other template", 40), ("tmpl", 300), ("big", 150000), ("tail", 5)):
        ref = "".join("ACGT-"[k] for k in rng.integers(0, 5, size=rows))
        counts = rng.integers(0, 70, size=(rows, 6))
        if name == "big":
            counts = counts * rng.choice([1, 1, 10, 900], size=(rows, 1))
        counts[rng.random(rows) < 0.1] = 0
        blocks.append((name, ref, counts))
    text = "".join(helpers.mat_text(nm, ref, c) for nm, ref, c in blocks)
    path = str(tmp_path / "s.mat")
    with open(path, "w") as f:
        f.write(text)
    with gzip.open(path + ".gz", "wt") as f:
        f.write(text)
    for target in ("tmpl", "other template", "big", "tail", "absent"):
        for min_depth in (1, 15):
            for q in (path, path + ".gz"):
                p = subprocess.run([exe, "mat", str(min_depth), q, target], capture_output=True, text=True)
                assert p.returncode == 0
                lines = p.stdout.splitlines()
                want = helpers.parse_mat(text, target)
                if want is None:
                    assert lines == ["status=0"]
                    continue
                counts, totals = want
                assert lines[0] == "status=1"
                assert lines[1] == "%d %d" % (len(totals), int((totals >= min_depth).sum()))
                got = np.array([[int(x) for x in ln.split()] for ln in lines[2:]], dtype=np.int64).reshape(-1, 7)
                assert np.array_equal(got[:, :6], counts.astype(np.int64)) and np.array_equal(got[:, 6], totals.astype(np.int64))


def test_option_scanner_dialect(built, tmp_path):
    def run(*args):
        return subprocess.run([BIN, "dist"] + list(args), capture_output=True, text=True, cwd=str(tmp_path))

    assert run("-h").stdout.startswith("#ccphylo-b200 dist")
    assert "Relaxed Phylip" in run("-F").stdout and "Relaxed Phylip" in run("--flag_help").stdout
    assert "# cos:" in run("-D").stdout
    assert run("--bogus").stderr == 'Unknown argument or option: "--bogus"\n'
    assert run("-Z").stderr == 'Unknown argument or option: "Z"\n'          # the letter alone, as the reference prints it
    assert run("-pZq").stderr == 'Unknown argument or option: "Z"\n'
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
    p = run("-pf1", "-s", "-W", "100", "-rref", "-i", str(a), str(a), "-P", "2", "-y", "m.txt")
    assert p.returncode != 0 and "m.txt" in p.stderr          # the motif file is the first thing the run opens
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
