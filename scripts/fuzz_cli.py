#!/usr/bin/env python
"""Differential fuzzer of the command line: random small FASTA sets and random `dist` / `trim` option
combinations go through the UNMODIFIED reference binary (oracle/_ref/ccphylo, host cores) and through
`ccphylo-b200` (GPU); return code, stdout, stderr, .phy, .num and the -V listing must be the same bytes.

Harness only (it executes oracle/_ref): the committed tests replay hand-picked option sets, this sweeps the
combinations nobody thought of.  Every case is reproducible from (seed, index): `--only K` re-runs case K,
`--self` compares the reference with itself (generator check on a box without a GPU).  Mismatching cases are
written to <out>/case<K>/ with their inputs, both command lines and both outputs.

Modes: default FASTA (`dist` / `trim`), --big (tensor-kernel sizes), --noise (FASTA layout variety), --mat (.mat files),
--mat --union (union files, a third of them on stdin), --add / --add --mat (-a against a matrix the reference built).
--bin PATH swaps the driver binary: tests/test_host_driver_cpu.py links host/*.c with tests/csrc/mock_ccg.c (the C-ABI
answered by the oracle) so the same sweeps run on a box without a GPU, under the sanitizers.

Known reference defects the generator stays away from: #3 shared-mask mode with excluded
samples (only -C 0 -L 1 there), #8 a gzip file as first input without flag 16, #16 unmapped bytes.
"""
import argparse
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ccphylo")
BIN = os.path.join(ROOT, "ccphylo_b200", "bin", "ccphylo-b200")

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
LOWER = np.frombuffer(b"acgt", dtype=np.uint8)
ODD = np.frombuffer(b"NNNN-RYSWKMBDHVXn", dtype=np.uint8)
LENGTHS = [1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 300, 777, 1000, 2049, 4100, 9000]


def make_rows(rng, n, length):
    snp = float(rng.choice([0.0, 0.002, 0.02, 0.15]))
    unk = float(rng.choice([0.0, 0.01, 0.08]))
    low = float(rng.choice([0.0, 0.0, 0.03]))
    ref = rng.integers(0, 4, size=length)
    rows = []
    for i in range(n):
        code = ref.copy()
        if i and rng.random() < 0.1:
            rows.append(rows[int(rng.integers(0, i))].copy())      # an identical sample
            continue
        sub = rng.random(length) < snp
        code[sub] = (code[sub] + rng.integers(1, 4, size=int(sub.sum()))) & 3
        row = BASES[code]
        u = rng.random(length) < unk
        row = np.where(u, ODD[rng.integers(0, len(ODD), size=length)], row)
        lo = rng.random(length) < low
        row = np.where(lo & ~u, LOWER[code], row)
        if length >= 16 and rng.random() < 0.3:                     # a run of unknowns
            a = int(rng.integers(0, length - 4))
            b = min(length, a + int(rng.integers(1, max(2, length // 3))))
            row[a:b] = ord("N")
        rows.append(row.astype(np.uint8))
    return rows


def make_motifs(rng):
    letters = "ACGTACGTACGTNRYSWKMBDHV"
    out = []
    for m in range(int(rng.integers(1, 5))):
        ln = int(rng.integers(1, 9))
        s = "".join(letters[int(k)] for k in rng.integers(0, len(letters), size=ln))
        out.append(f">m{m}\n{s}\n")
    return "".join(out)


def make_case(seed, idx, big=False):
    rng = np.random.default_rng([seed, idx])
    tool = "trim" if rng.random() < 0.15 else "dist"
    n = int(rng.integers(2, 25))
    length = int(rng.choice(LENGTHS)) if rng.random() < 0.7 else int(rng.integers(1, 6000))
    if big:                 # sizes at which the driver takes the tensor-core kernel (from 192 samples x 8192 bases on)
        tool, n, length = "dist", int(rng.integers(192, 331)), int(rng.integers(8192, 20001))
    rows = make_rows(rng, n, length)
    pair = rng.random() < 0.7
    flag = (2 if pair else 0) | int(rng.choice([0, 1])) | int(rng.choice([0, 0, 4])) | int(rng.choice([0, 8])) | int(rng.choice([0, 0, 32]))
    mode = str(rng.choice(["files", "files", "msa", "msa", "msa_gz", "files_gz"]))
    if tool == "trim":
        # trim's own flag table (trim -F): 1 hard mask, 2 pairwise, 4 mask gaps / ambiguous, 8 unmask soft-masked,
        # 16 pseudo alignment (not with 2), 32 prune without insignificant bases; plain files only (App. B #8)
        flag = (2 if pair else int(rng.choice([0, 16]))) | int(rng.choice([0, 1])) | int(rng.choice([0, 4])) | int(rng.choice([0, 8])) | int(rng.choice([0, 0, 32]))
        mode = mode.replace("_gz", "")
    elif mode.endswith("_gz"):
        flag |= 16
    # (the reference's trim dies with SIGSEGV whenever a sample is excluded: keep that rare there)
    if pair and length >= 8 and rng.random() < (0.25 if tool == "dist" else 0.03):
        rows[int(rng.integers(0, n))][: length - int(rng.integers(0, max(1, length // 3)))] = ord("N")    # fails the coverage gate
    args = []
    if pair:
        c = rng.choice(["", "0", "30", "90"] if tool == "dist" else ["0", "0", "0", "20"])
        if c:
            args += ["-C", str(c)]
        if rng.random() < 0.3:
            args += ["-L", str(int(rng.integers(0, length + 2)))]
    else:
        args += ["-C", "0"]
        if rng.random() < 0.3:
            args += ["-L", "1"]
    proxi = int(rng.choice([0, 0, 0, 1, 2, 5, 33, 200]))
    if proxi:
        args += ["-P", str(proxi)]
    motifs = make_motifs(rng) if rng.random() < 0.25 else None
    case = {"idx": idx, "tool": tool, "n": n, "length": length, "mode": mode, "flag": flag, "motifs": motifs, "rows": rows,
            "variants": False, "out_stdout": False, "num": "o.num"}
    if tool == "dist":
        w = rng.choice(["", "7", "1000", "1000000"])
        if w:
            args += ["-W", str(w)]
        cell = rng.choice(["", "", "-p", "-s", "-b"])
        if cell == "-p":
            args += ["-p"]
        elif cell in ("-s", "-b"):
            args += [str(cell)]
            if rng.random() < 0.7:
                args += [str(rng.choice(["0.001", "0.5", "1", "10", "100"]))]
        if rng.random() < 0.15:
            args += ["-x", str(int(rng.integers(0, 13)))]
        if rng.random() < 0.1:
            args += ["-H"]
        case["variants"] = bool(rng.random() < 0.3) and not big      # (a listing of 50,000 pairs is all print)
        case["out_stdout"] = bool(rng.random() < 0.15)
        case["num"] = str(rng.choice(["o.num", "o.num", "", "o.phy"]))
    threads = 1 if case["variants"] else int(rng.integers(1, 6))
    if tool == "dist":
        args += ["-t", str(threads)]
    case["args"] = args
    return case


NOISE = False        # --noise: FASTA layout variety (line widths, junk bytes the parsers drop, other records around the target)


def _record(f, header, row, rng):
    """one FASTA record; with NOISE: any line width, digits / blanks / '*' inside the lines (dropped by the parser,
    seqparse.c:217-231), a missing final newline"""
    f.write(b">" + header + b"\n")
    if rng is None:
        for s in range(0, len(row), 60):
            f.write(row[s:s + 60].tobytes() + b"\n")
        return
    width = int(rng.choice([1, 7, 60, 61, 80, 1 << 30]))
    junk = rng.random() < 0.4
    body = []
    for s in range(0, len(row), width):
        line = row[s:s + width].tobytes()
        if junk and len(line) > 2 and rng.random() < 0.3:
            k = int(rng.integers(1, len(line)))
            line = line[:k] + bytes(rng.choice(np.frombuffer(b" 0123456789*\t", dtype=np.uint8), size=int(rng.integers(1, 4)))) + line[k:]
        body.append(line)
    text = b"\n".join(body) + (b"" if rng.random() < 0.2 else b"\n")
    f.write(text)
    if not text.endswith(b"\n") :
        return "open"
    return None


def write_inputs(case, d):
    rows, mode = case["rows"], case["mode"]
    opener = (lambda p: gzip.open(p, "wb")) if mode.endswith("_gz") else (lambda p: open(p, "wb"))
    ext = ".fsa.gz" if mode.endswith("_gz") else ".fsa"
    rng = np.random.default_rng([case["idx"], len(rows), 3]) if NOISE else None
    if mode.startswith("msa"):
        path = os.path.join(d, "aln" + ext)
        with opener(path) as f:
            for i, row in enumerate(rows):
                if _record(f, b"s%d" % i, row, rng) == "open" and i + 1 < len(rows):
                    f.write(b"\n")
        inputs = ["-i", path]
    else:
        files = []
        for i, row in enumerate(rows):
            path = os.path.join(d, f"s{i:02d}{ext}")
            with opener(path) as f:
                if rng is not None and rng.random() < 0.3:            # another record in front of the target
                    _record(f, b"other_contig", row[: max(1, len(row) // 3)][::-1].copy(), None)
                state = _record(f, b"ref", row, rng)
                if rng is not None and rng.random() < 0.3:            # and one behind it
                    if state == "open":
                        f.write(b"\n")
                    _record(f, b"ref2", row[: max(1, len(row) // 2)].copy(), None)
            files.append(path)
        inputs = ["-r", "ref", "-i"] + files
    if case["motifs"]:
        with open(os.path.join(d, "motifs.fsa"), "w") as f:
            f.write(case["motifs"])
    return inputs


def command(case, exe, d):
    inputs = write_inputs(case, d)
    cmd = [exe, case["tool"], "-f", str(case["flag"])] + case["args"]
    if case["motifs"]:
        cmd += ["-y", os.path.join(d, "motifs.fsa")]
    if case["tool"] == "dist":
        if case["variants"]:
            cmd += ["-V", os.path.join(d, "v.txt")]
        if not case["out_stdout"]:
            cmd += ["-o", os.path.join(d, "o.phy")]
        if case["num"]:
            cmd += ["-n", os.path.join(d, case["num"])]
    else:
        cmd += ["-o", os.path.join(d, "o.fsa")]
    return cmd + inputs


def run_one(case, exe, d):
    os.makedirs(d, exist_ok=True)
    cmd = command(case, exe, d)
    try:
        p = subprocess.run(cmd, capture_output=True, cwd=d, timeout=20)
        rc, out, err = p.returncode, p.stdout, p.stderr
    except subprocess.TimeoutExpired:
        rc, out, err = -999, b"", b"timeout"
    # (with one sample left the reference asks for 0 worker threads, fails to start 2^19 of them and says so: not compared)
    junk = (b"Error: 11 (", b"Will continue with ")
    norm = lambda b: b"\n".join(ln for ln in b.replace(d.encode() + b"/", b"").split(b"\n") if not ln.startswith(junk))

    def text(name, sort=False):
        q = os.path.join(d, name)
        if not os.path.exists(q):
            return None
        b = open(q, "rb").read()
        return b"\n".join(sorted(b.split(b"\n"))) if sort else b
    res = {"rc": rc, "stdout": norm(out), "stderr": b"\n".join(sorted(norm(err).split(b"\n"))),
           "phy": text("o.phy"), "num": text("o.num"), "variants": text("v.txt", sort=True), "trim": text("o.fsa")}
    return cmd, res


METHODS = ["cos", "z", "chi2", "nchi2", "c", "nc", "p", "np", "bc", "nbc", "l1", "l2", "linf", "l3", "nl1", "nl2", "nlinf", "nl3"]


def make_mat_case(seed, idx):
    """.mat inputs (KMA count matrices): every -d method, depth / coverage gates, gz, another template in front.  No
    '-' reference rows (App. B #12), depths far below 46,340 (#15)."""
    rng = np.random.default_rng([seed, idx, 7])
    n = int(rng.integers(2, 11))
    length = int(rng.choice([1, 2, 31, 32, 33, 64, 100, 257, 600])) if rng.random() < 0.6 else int(rng.integers(1, 700))
    ref = "".join("ACGT"[k] for k in rng.integers(0, 4, size=length))
    depth = float(rng.choice([8, 20, 40, 200]))
    err = float(rng.choice([0.0, 0.3, 2.0]))
    low = float(rng.choice([0.0, 0.02, 0.3]))
    snp = float(rng.choice([0.0, 0.02, 0.2]))
    texts = []
    other = None
    if rng.random() < 0.3:
        oref = "".join("ACGT"[k] for k in rng.integers(0, 4, size=int(rng.integers(1, 50))))
        other = (oref, rng)
    for k in range(n):
        if k and rng.random() < 0.1:
            texts.append(texts[int(rng.integers(0, k))])
            continue
        lines = []
        if other:
            lines.append("#other_template")
            for b in other[0]:
                lines.append(b + "\t" + "\t".join(str(int(x)) for x in rng.poisson(10, size=6)))
            lines.append("")
        lines.append("#tmpl")
        lowk = 0.9 if rng.random() < 0.1 else low            # a sample mostly below the depth gate
        call = np.array(["ACGT".index(b) for b in ref])
        sub = rng.random(length) < snp
        call[sub] = (call[sub] + rng.integers(1, 4, size=int(sub.sum()))) % 4
        d = rng.poisson(depth, size=length)
        lo = rng.random(length) < lowk
        d[lo] = rng.integers(0, 6, size=int(lo.sum()))
        c = rng.poisson(err, size=(length, 6)) if err else np.zeros((length, 6), dtype=np.int64)
        c[:, 4:] = c[:, 4:] // 3
        c[np.arange(length), call] += d
        for b, row in zip(ref, c):
            lines.append(b + "\t" + "\t".join(str(int(x)) for x in row))
        texts.append("\n".join(lines) + "\n\n")
    args = ["-d", str(rng.choice(METHODS))]
    e = rng.choice(["", "", "1", "5", "30"])
    if e:
        args += ["-E", str(e)]
    cc = rng.choice(["", "0", "30", "90"])
    if cc:
        args += ["-C", str(cc)]
    if rng.random() < 0.25:
        args += ["-L", str(int(rng.integers(0, length + 2)))]
    w = rng.choice(["", "", "1000", "1000000"])
    if w:
        args += ["-W", str(w)]
    if rng.random() < 0.15:
        args += ["-x", str(int(rng.integers(0, 13)))]
    if rng.random() < 0.2:
        args += ["-l", str(rng.choice(["0.05", "0.01", "0.5"]))]
    flag = int(rng.choice([0, 1])) | int(rng.choice([0, 0, 4])) | int(rng.choice([0, 2]))
    args += ["-f", str(flag), "-t", str(int(rng.integers(1, 5)))]
    return {"idx": idx, "tool": "dist", "n": n, "length": length, "mode": "mat_gz" if rng.random() < 0.4 else "mat", "flag": flag,
            "motifs": None, "variants": False, "texts": texts, "args": args}


def make_union_case(seed, idx):
    """the same count matrices behind a `union` file (dist.c:204-266: the file names the KMA result files, `dist` turns
    *.res into *.mat.gz; one Phylip block per template row, each over its own subset of the samples)"""
    case = make_mat_case(seed, idx)
    rng = np.random.default_rng([seed, idx, 9])
    n = case["n"]
    templates = ["tmpl"] + (["other_template"] if "#other_template" in case["texts"][0] else [])
    rows = []
    for t in templates:
        k = n if rng.random() < 0.5 else int(rng.integers(1, n + 1))
        members = sorted(int(x) for x in rng.choice(n, size=k, replace=False))
        rows.append((t, members))
    if len(rows) == 2 and rng.random() < 0.5:
        rows.reverse()
    # (with -L 0 -C 0 a pair without a comparable position gives 0/0; the union twin then fails its `-1.0 <= dist` test on
    # the NaN, drops the COLUMN sample from there on and leaves the packed rows out of step, ltdmatrix.c:163-177: not a
    # behaviour to reproduce -- the gates stay on here)
    a = case["args"]
    if "-L" in a and a[a.index("-L") + 1] == "0":
        a[a.index("-L") + 1] = "1"
    case["mode"] = "union"
    case["union_rows"] = rows
    case["stdin"] = bool(rng.random() < 0.4)          # `ccphylo union ... | ccphylo dist`: no -i, the union text on stdin
    return case


def run_mat(case, exe, d):
    os.makedirs(d, exist_ok=True)
    if case["mode"] == "union":
        names = ["%c.mat.gz" % (ord("a") + k) for k in range(case["n"])]
        for nm, text in zip(names, case["texts"]):
            with gzip.open(os.path.join(d, nm), "wt") as f:
                f.write(text)
        union = "%d\t" % case["n"] + "\t".join(os.path.join(d, nm.replace(".mat.gz", ".res")) for nm in names) + "\n"
        for t, members in case["union_rows"]:
            union += t + "\t%d\t" % len(members) + "\t".join(str(k) for k in members) + "\n"
        with open(os.path.join(d, "in.union"), "w") as f:
            f.write(union)
        src = [] if case.get("stdin") else ["-i", os.path.join(d, "in.union")]
        cmd = [exe, "dist"] + src + list(case["args"]) + ["-o", os.path.join(d, "o.phy"), "-n", os.path.join(d, "o.num")]
        return cmd, _run_mat_cmd(cmd, d, os.path.join(d, "in.union") if case.get("stdin") else None)
    files = []
    for k, text in enumerate(case["texts"]):
        path = os.path.join(d, "%c.mat" % (ord("a") + k) + (".gz" if case["mode"] == "mat_gz" else ""))
        with (gzip.open(path, "wt") if case["mode"] == "mat_gz" else open(path, "w")) as f:
            f.write(text)
        files.append(path)
    args = list(case["args"])
    if exe == REF_BIN:
        args[args.index("-t") + 1] = "1"        # the reference's threaded .mat loop hangs now and then (seen with -t 2 .. 4)
    cmd = [exe, "dist", "-r", "tmpl", "-i"] + files + args + ["-o", os.path.join(d, "o.phy"), "-n", os.path.join(d, "o.num")]
    return cmd, _run_mat_cmd(cmd, d)


def _run_mat_cmd(cmd, d, stdin_path=None):
    try:
        with open(stdin_path or os.devnull, "rb") as fin:
            p = subprocess.run(cmd, capture_output=True, cwd=d, timeout=20, stdin=fin)
        rc, out, err = p.returncode, p.stdout, p.stderr
    except subprocess.TimeoutExpired:
        rc, out, err = -999, b"", b"timeout"
    norm = lambda b: b.replace(d.encode() + b"/", b"")
    rd = lambda q: open(os.path.join(d, q), "rb").read() if os.path.exists(os.path.join(d, q)) else None
    return {"rc": rc, "stdout": norm(out), "stderr": b"\n".join(sorted(norm(err).split(b"\n"))), "phy": rd("o.phy"), "num": rd("o.num")}


def cells_close(a, b, digits):
    """Phylip text of a float matrix: same layout and names, cells within 1e-6 relative (the documented bar of the .mat
    path: the sum over positions is not taken in the reference's order) plus one unit of the printed precision."""
    if a is None or b is None:
        return a == b
    la, lb = a.split(b"\n"), b.split(b"\n")
    if len(la) != len(lb):
        return False
    for x, y in zip(la, lb):
        fx, fy = x.split(b"\t"), y.split(b"\t")
        if len(fx) != len(fy) or fx[0] != fy[0]:
            return False
        for u, v in zip(fx[1:], fy[1:]):
            if u == v:
                continue
            try:
                p, q = float(u), float(v)
            except ValueError:
                return False
            if p != p and q != q:
                continue
            if not abs(p - q) <= 1e-6 * max(abs(p), abs(q)) + 1.01 * 10.0 ** -digits:
                return False
    return True


def check_mat(case, keep_dir, self_check):
    base = tempfile.mkdtemp(prefix="fuzzm%d_" % case["idx"])
    try:
        rcmd, ref = run_mat(case, REF_BIN, os.path.join(base, "reference"))
        dcmd, drv = run_mat(case, REF_BIN if self_check else BIN, os.path.join(base, "driver"))
        digits = 9
        if "-x" in case["args"]:
            digits = int(case["args"][case["args"].index("-x") + 1])
        diff = [k for k in ("rc", "stdout", "stderr", "num") if ref[k] != drv[k]]
        if not cells_close(ref["phy"], drv["phy"], digits):
            diff.append("phy")
        verdict = "ok" if not diff else ("ref_crash" if ref["rc"] < 0 else "MISMATCH")
        if b"unsupported by the CPU mock" in drv["stderr"]:
            verdict = "unsupported"
        if verdict == "MISMATCH" and keep_dir:
            dst = os.path.join(keep_dir, "matcase%d" % case["idx"])
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(base, dst)
            with open(os.path.join(dst, "case.json"), "w") as f:
                json.dump({"idx": case["idx"], "diff": diff, "reference_cmd": rcmd, "driver_cmd": dcmd, "reference_rc": ref["rc"],
                           "driver_rc": drv["rc"], "reference_stderr": ref["stderr"].decode(errors="replace")[-2000:],
                           "driver_stderr": drv["stderr"].decode(errors="replace")[-2000:]}, f, indent=1)
        return {"idx": case["idx"], "verdict": verdict, "diff": diff, "rc": [ref["rc"], drv["rc"]],
                "what": "mat n=%d L=%d %s %s" % (case["n"], case["length"], case["mode"], " ".join(case["args"]))}
    finally:
        shutil.rmtree(base, ignore_errors=True)


def check(case, keep_dir, self_check):
    base = tempfile.mkdtemp(prefix="fuzz%d_" % case["idx"])
    try:
        rcmd, ref = run_one(case, REF_BIN, os.path.join(base, "reference"))
        dcmd, drv = run_one(case, REF_BIN if self_check else BIN, os.path.join(base, "driver"))
        diff = [k for k in ref if ref[k] != drv[k]]
        # a reference that crashed (signal) or hung (20 s) proves nothing either way
        verdict = "ok" if not diff else ("ref_crash" if ref["rc"] < 0 else "MISMATCH")
        if b"unsupported by the CPU mock" in drv["stderr"]:
            verdict = "unsupported"
        # App. B #3: shared-mask mode with an excluded sample -- the reference compares the wrong pairs (the excluded
        # sample's words among them); the driver follows the intended semantics, documented in DESIGN.md
        if verdict == "MISMATCH" and case["tool"] == "dist" and not case["flag"] & 2 and b"# Excluded:" in ref["stderr"]:
            verdict = "known_divergence_3"
        # trim -f 8 without -f 1 / -f 4: the reference leaves the soft flag on the letters it prints and indexes past its
        # 16-letter table (trim.c:40,50,60-64): bytes from beyond `bases`; the driver prints the letter (host/trim_main.c)
        # (the same beside an unknown base of the reference sample whatever the flags say: only -f 1 / -f 4 have no soft letters)
        if (verdict == "MISMATCH" and case["tool"] == "trim" and diff == ["trim"] and not case["flag"] & 5
                and len(ref["trim"]) == len(drv["trim"])
                and all(65 <= (y & ~32) <= 90 for x, y in zip(ref["trim"], drv["trim"]) if x != y)):
            verdict = "known_trim_soft_letters"
        if verdict == "MISMATCH" and keep_dir:
            dst = os.path.join(keep_dir, "case%d" % case["idx"])
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(base, dst)
            with open(os.path.join(dst, "case.json"), "w") as f:
                json.dump({"idx": case["idx"], "diff": diff, "reference_cmd": rcmd, "driver_cmd": dcmd,
                           "reference_rc": ref["rc"], "driver_rc": drv["rc"],
                           "reference_stderr": ref["stderr"].decode(errors="replace")[-2000:],
                           "driver_stderr": drv["stderr"].decode(errors="replace")[-2000:]}, f, indent=1)
        return {"idx": case["idx"], "verdict": verdict, "diff": diff, "rc": [ref["rc"], drv["rc"]],
                "what": "%s n=%d L=%d %s -f %d %s%s%s" % (case["tool"], case["n"], case["length"], case["mode"], case["flag"],
                                                         " ".join(case["args"]), " -y" if case["motifs"] else "",
                                                         " -V" if case["variants"] else "")}
    finally:
        shutil.rmtree(base, ignore_errors=True)


def check_add(seed, idx, keep_dir, self_check):
    """-a: the reference builds the matrix of the first n - 1 files, then the last file is added to copies of it by the
    reference and by the driver (add2Matrix dist.c:331-411): the .phy, the .num, the -V listing and sorted stderr must agree"""
    case = make_case(seed, idx)
    rng = np.random.default_rng([seed, idx, 5])
    n = max(3, case["n"])
    case["rows"] = (case["rows"] * 2)[:n]
    args = [a for a in case["args"]]
    for opt, has_value in (("-H", False), ("-x", True), ("-t", True)):
        while opt in args:
            k = args.index(opt)
            del args[k:k + (2 if has_value else 1)]
    for opt in ("-p", "-s", "-b"):                      # the added row is doubles (dist.c:386-391)
        while opt in args:
            k = args.index(opt)
            del args[k:k + (2 if k + 1 < len(args) and not args[k + 1].startswith("-") else 1)]
    flag = (case["flag"] | 2) & ~(4 | 16)
    variants = bool(rng.random() < 0.3)
    threads = 1 if variants else int(rng.integers(1, 4))
    base = tempfile.mkdtemp(prefix="fuzza%d_" % idx)
    try:
        res = {}
        for tag, exe in (("reference", REF_BIN), ("driver", REF_BIN if self_check else BIN)):
            d = os.path.join(base, tag)
            os.makedirs(d)
            files = []
            for i, row in enumerate(case["rows"]):
                files.append(os.path.join(d, "s%02d.fsa" % i))
                with open(files[-1], "wb") as f:
                    _record(f, b"ref", row, None)
            common = ["-r", "ref", "-f", str(flag)] + args + (["-V", "v.txt"] if variants else [])
            p0 = subprocess.run([REF_BIN, "dist"] + common + ["-t", "1", "-i"] + files[:-1] + ["-o", "m.phy", "-n", "m.num"], capture_output=True,
                                cwd=d, timeout=60)
            if p0.returncode != 0 or not os.path.exists(os.path.join(d, "m.phy")) or os.path.getsize(os.path.join(d, "m.phy")) == 0:
                return {"idx": idx, "verdict": "ref_crash" if p0.returncode < 0 else "no_matrix", "diff": [], "rc": [p0.returncode, 0], "what": "add"}
            cmd = [exe, "dist"] + common + ["-t", str(threads), "-a", files[-1], "-i", files[0],
                                            "-o", "m.phy", "-n", "m.num"]
            try:
                p = subprocess.run(cmd, capture_output=True, cwd=d, timeout=60)
                rc, err = p.returncode, p.stderr
            except subprocess.TimeoutExpired:
                rc, err = -999, b"timeout"
            rd = lambda q: open(os.path.join(d, q), "rb").read() if os.path.exists(os.path.join(d, q)) else None
            junk = (b"Error: 11 (", b"Will continue with ")
            lines = sorted(ln for ln in err.replace(d.encode() + b"/", b"").split(b"\n") if not ln.startswith(junk))
            v = rd("v.txt")
            res[tag] = {"rc": rc, "stderr": lines, "phy": rd("m.phy"), "num": rd("m.num"), "variants": sorted(v.split(b"\n")) if v else v, "cmd": cmd}
        ref, drv = res["reference"], res["driver"]
        diff = [k for k in ("rc", "stderr", "phy", "num", "variants") if ref[k] != drv[k]]
        verdict = "ok" if not diff else ("ref_crash" if ref["rc"] < 0 else "MISMATCH")
        if any(b"unsupported by the CPU mock" in ln for ln in drv["stderr"]):
            verdict = "unsupported"
        if verdict == "MISMATCH" and keep_dir:
            dst = os.path.join(keep_dir, "addcase%d" % idx)
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(base, dst)
            with open(os.path.join(dst, "case.json"), "w") as f:
                json.dump({"idx": idx, "diff": diff, "reference_cmd": ref["cmd"], "driver_cmd": drv["cmd"], "reference_rc": ref["rc"],
                           "driver_rc": drv["rc"], "reference_stderr": b"\n".join(ref["stderr"]).decode(errors="replace")[-2000:],
                           "driver_stderr": b"\n".join(drv["stderr"]).decode(errors="replace")[-2000:]}, f, indent=1)
        return {"idx": idx, "verdict": verdict, "diff": diff, "rc": [ref["rc"], drv["rc"]],
                "what": "add n=%d L=%d -f %d %s%s" % (n, case["length"], flag, " ".join(args), " -V" if variants else "")}
    finally:
        shutil.rmtree(base, ignore_errors=True)


def check_add_mat(seed, idx, keep_dir, self_check):
    """-a on .mat input: the reference builds the matrix of the first n - 1 count matrices, the last one is added to copies
    of it by the reference and by the driver (ltdRow_get ltdmatrix.c:205); new row within 1e-6, old rows untouched"""
    case = make_mat_case(seed, idx)
    n = max(3, case["n"])
    texts = (case["texts"] * 2)[:n]
    args = list(case["args"])
    for opt in ("-x", "-t"):
        if opt in args:
            k = args.index(opt)
            del args[k:k + 2]
    base = tempfile.mkdtemp(prefix="fuzzam%d_" % idx)
    try:
        res = {}
        for tag, exe in (("reference", REF_BIN), ("driver", REF_BIN if self_check else BIN)):
            d = os.path.join(base, tag)
            os.makedirs(d)
            files = []
            for k, text in enumerate(texts):
                files.append(os.path.join(d, "%c.mat" % (ord("a") + k)))
                with open(files[-1], "w") as f:
                    f.write(text)
            p0 = subprocess.run([REF_BIN, "dist", "-r", "tmpl", "-t", "1", "-i"] + files[:-1] + args + ["-o", "m.phy", "-n", "m.num"],
                                capture_output=True, cwd=d, timeout=60)
            if p0.returncode != 0 or not os.path.exists(os.path.join(d, "m.phy")) or os.path.getsize(os.path.join(d, "m.phy")) == 0:
                return {"idx": idx, "verdict": "ref_crash" if p0.returncode < 0 else "no_matrix", "diff": [], "rc": [p0.returncode, 0], "what": "add mat"}
            cmd = [exe, "dist", "-r", "tmpl", "-t", "1", "-a", files[-1], "-i", files[0]] + args + ["-o", "m.phy", "-n", "m.num"]
            try:
                p = subprocess.run(cmd, capture_output=True, cwd=d, timeout=60)
                rc, err = p.returncode, p.stderr
            except subprocess.TimeoutExpired:
                rc, err = -999, b"timeout"
            rd = lambda q: open(os.path.join(d, q), "rb").read() if os.path.exists(os.path.join(d, q)) else None
            res[tag] = {"rc": rc, "stderr": sorted(err.replace(d.encode() + b"/", b"").split(b"\n")), "phy": rd("m.phy"), "num": rd("m.num"), "cmd": cmd}
        ref, drv = res["reference"], res["driver"]
        diff = [k for k in ("rc", "stderr", "num") if ref[k] != drv[k]]
        if not cells_close(ref["phy"], drv["phy"], 9):
            diff.append("phy")
        verdict = "ok" if not diff else ("ref_crash" if ref["rc"] < 0 else "MISMATCH")
        if any(b"unsupported by the CPU mock" in ln for ln in drv["stderr"]):
            verdict = "unsupported"
        if verdict == "MISMATCH" and keep_dir:
            dst = os.path.join(keep_dir, "addmatcase%d" % idx)
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(base, dst)
            with open(os.path.join(dst, "case.json"), "w") as f:
                json.dump({"idx": idx, "diff": diff, "reference_cmd": ref["cmd"], "driver_cmd": drv["cmd"], "reference_rc": ref["rc"],
                           "driver_rc": drv["rc"], "reference_stderr": b"\n".join(ref["stderr"]).decode(errors="replace")[-2000:],
                           "driver_stderr": b"\n".join(drv["stderr"]).decode(errors="replace")[-2000:]}, f, indent=1)
        return {"idx": idx, "verdict": verdict, "diff": diff, "rc": [ref["rc"], drv["rc"]],
                "what": "add mat n=%d L=%d %s" % (n, case["length"], " ".join(args))}
    finally:
        shutil.rmtree(base, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--budget-s", type=float, default=0.0, help="stop handing out cases after this many seconds")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fuzz"))
    ap.add_argument("--self", dest="self_check", action="store_true")
    ap.add_argument("--bin", default=None, help="the driver binary (default ccphylo_b200/bin/ccphylo-b200); tests/csrc/mock_ccg.c "
                    "gives one that runs on the CPU")
    ap.add_argument("--big", action="store_true", help="192 .. 330 samples x 8 .. 20 kbp: the tensor-core kernel behind the command line")
    ap.add_argument("--add", action="store_true", help="-a: one more FASTA sample against a matrix the reference built")
    ap.add_argument("--noise", action="store_true", help="FASTA layout variety: line widths, junk bytes, records around the target")
    ap.add_argument("--union", action="store_true", help="with --mat: the count matrices behind a union file")
    ap.add_argument("--mat", action="store_true", help=".mat inputs (cells compared within 1e-6 relative) instead of FASTA")
    a = ap.parse_args()
    if a.bin:
        global BIN
        BIN = a.bin
    if a.noise:
        global NOISE
        NOISE = True
    os.makedirs(a.out, exist_ok=True)
    idxs = [a.only] if a.only >= 0 else list(range(a.first, a.first + a.cases))
    t0 = time.time()

    def job(i):
        if a.budget_s and time.time() - t0 > a.budget_s:
            return None
        if a.add:
            return check_add_mat(a.seed, i, a.out, a.self_check) if a.mat else check_add(a.seed, i, a.out, a.self_check)
        if a.mat:
            return check_mat(make_union_case(a.seed, i) if a.union else make_mat_case(a.seed, i), a.out, a.self_check)
        return check(make_case(a.seed, i, a.big), a.out, a.self_check)
    with ThreadPoolExecutor(max_workers=a.workers) as ex:
        results = [r for r in ex.map(job, idxs) if r]
    bad = [r for r in results if r["verdict"] == "MISMATCH"]
    summary = {"seed": a.seed, "cases_run": len(results), "ok": sum(r["verdict"] == "ok" for r in results),
               "reference_crashed": sum(r["verdict"] == "ref_crash" for r in results), "mismatch": len(bad),
               "known_divergence_3": sum(r["verdict"] == "known_divergence_3" for r in results),
               "unsupported_by_mock": sum(r["verdict"] == "unsupported" for r in results),
               "known_trim_soft_letters": sum(r["verdict"] == "known_trim_soft_letters" for r in results), "nonzero_rc_both": sum(r["verdict"] == "ok" and r["rc"][0] != 0 for r in results),
               "seconds": round(time.time() - t0, 1), "self_check": a.self_check, "mismatches": bad,
               "ref_crashes": [r for r in results if r["verdict"] == "ref_crash"]}
    with open(os.path.join(a.out, ("summary_add_mat_seed%d.json" if a.add and a.mat else "summary_add_seed%d.json" if a.add else "summary_mat_seed%d.json" if a.mat else "summary_big_seed%d.json" if a.big else "summary_seed%d.json") % a.seed), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k not in ("mismatches", "ref_crashes")}))
    for r in bad[:40]:
        print("MISMATCH", r["idx"], r["diff"], r["rc"], r["what"])
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
