"""The host driver `ccphylo-b200 dist` end to end on the GPU box: the reference binary's golden
.phy / .num / stderr text must come out byte for byte for FASTA inputs (multi-file, MSA, every
cell type, shared-mask mode), and within the printed precision for .mat inputs (files, gz, union)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.environ.get("CCPHYLO_TEST_BIN") or os.path.join(ROOT, "ccphylo_b200", "bin", "ccphylo-b200")   # (tests/csrc/mock_ccg.c gives a CPU driver)

POOL, ALL_CASES = helpers.golden_cases()
# every invocation pays a CUDA context start (~2 s): replay a representative third of the fixture here; the
# whole fixture goes through the C-ABI in test_gpu_parity.py
CASES = [c for c in ALL_CASES if c["name"].startswith(("c1_", "c6_", "c7_", "rand_L33_", "rand_L4100_"))]


def run(cmd, cwd):
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=cwd, timeout=300)
    return p


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_fasta_golden_byte_for_byte(built, tmp_path, case):
    td = str(tmp_path)
    seqs = [POOL[k] for k in case["seq_ids"]]
    if case["msa"]:
        path = os.path.join(td, "msa.fsa")
        with open(path, "w") as f:
            for nm, s in zip(case["names"], seqs):
                f.write(f">{nm}\n{s}\n")
        cmd = [BIN, "dist", "-i", path]
    else:
        files = []
        for nm, s in zip(case["names"], seqs):
            path = os.path.join(td, nm)
            with open(path, "w") as f:
                f.write(f">ref\n{s}\n")
            files.append(path)
        cmd = [BIN, "dist", "-r", "ref", "-i"] + files
    phy, num = os.path.join(td, "o.phy"), os.path.join(td, "o.num")
    p = run(cmd + case["args"] + ["-o", phy, "-n", num, "-t", "3"], td)
    assert p.returncode == case["returncode"], p.stderr
    assert open(phy).read() == case["phy"]
    assert open(num).read() == case["num"]
    assert p.stderr.replace(td + "/", "") == case["stderr"]


# (with a CPU driver from tests/csrc/mock_ccg.c there is no device for the bound reference either)
REF_GPU = os.environ.get("CCPHYLO_TEST_REF_GPU") or os.path.join(ROOT, "oracle", "_ref", "ccphylo_gpu" if not os.environ.get("CCPHYLO_TEST_BIN") else "ccphylo_gpu.absent")
BOUND = [c for c in ALL_CASES if c["name"].startswith(("c1_pair", "c1_global", "c1_float", "c1_short_W", "c6_", "c7_",
                                                       "rand_L129_", "rand_L4100_pair"))]


@pytest.mark.skipif(not os.path.exists(REF_GPU), reason="oracle/_ref/ccphylo_gpu was not built (needs /root/reference)")
@pytest.mark.parametrize("case", BOUND, ids=[c["name"] for c in BOUND])
def test_reference_bound_to_the_gpu_library(built, tmp_path, case):
    """The drop-in claim, literally: the unmodified reference binary with fsaCmpThreadOut bound to
    libccphylo_gpu.so through integration/fsacmpgpu.c (oracle/Makefile, target ref_gpu) prints what the
    original prints."""
    td = str(tmp_path)
    seqs = [POOL[k] for k in case["seq_ids"]]
    if case["msa"]:
        path = os.path.join(td, "msa.fsa")
        with open(path, "w") as f:
            for nm, s in zip(case["names"], seqs):
                f.write(f">{nm}\n{s}\n")
        cmd = [REF_GPU, "dist", "-i", path]
    else:
        files = []
        for nm, s in zip(case["names"], seqs):
            path = os.path.join(td, nm)
            with open(path, "w") as f:
                f.write(f">ref\n{s}\n")
            files.append(path)
        cmd = [REF_GPU, "dist", "-r", "ref", "-i"] + files
    phy, num = os.path.join(td, "o.phy"), os.path.join(td, "o.num")
    p = run(cmd + case["args"] + ["-o", phy, "-n", num], td)
    assert p.returncode == case["returncode"], p.stderr
    assert open(phy).read() == case["phy"]
    assert open(num).read() == case["num"]
    assert p.stderr.replace(td + "/", "") == case["stderr"]


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ccphylo")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
def test_config1_64_samples_5mbp_against_the_reference_binary(built, tmp_path):
    """BASELINE configs[0] at full size: 64 synthetic 5 Mbp KMA consensus files, `dist -r ref -f 3 -n`.  The
    unmodified reference binary runs on the host cores; this driver and the reference bound to the GPU library
    must print the same bytes (D, N and the stderr inclusion lines)."""
    import time
    from ccphylo_b200 import synth
    td = str(tmp_path)
    n, length = 64, 5_000_000
    rows = synth.make_ascii(n, length, seed=1)
    files = []
    for i in range(n):
        path = os.path.join(td, f"s{i:02d}.fsa")
        synth.write_fasta(path, rows[i], header="ref", width=60)
        files.append(path)
    del rows
    outs = {}
    for tag, exe in (("reference", REF_BIN), ("driver", BIN), ("bound", REF_GPU)):
        if not os.path.exists(exe):
            continue
        phy, num = os.path.join(td, tag + ".phy"), os.path.join(td, tag + ".num")
        t0 = time.perf_counter()
        p = run([exe, "dist", "-r", "ref", "-f", "3", "-t", str(os.cpu_count() or 1), "-i"] + files + ["-o", phy, "-n", num], td)
        dt = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr[-2000:]
        outs[tag] = (open(phy).read(), open(num).read(), p.stderr, dt)
        print(f"config 1 ({n} x {length} bp) {tag}: {dt:.2f} s wall")
    assert len(outs["reference"][0]) > 2000
    for tag in outs:
        assert outs[tag][0] == outs["reference"][0], tag + ": .phy differs from the reference binary's"
        assert outs[tag][1] == outs["reference"][1], tag + ": .num differs from the reference binary's"
        assert outs[tag][2] == outs["reference"][2], tag + ": stderr differs from the reference binary's"


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
@pytest.mark.parametrize("flag", ["3", "1"], ids=["pair", "shared-mask"])
def test_msa_parallel_and_gz_against_the_reference_binary(built, tmp_path, flag):
    """MSA input (one multi-FASTA alignment): the plain file is parsed by -t threads seeking to their records,
    the gz file sequentially; both must print what the reference binary prints (D block, then -- in pair mode --
    the N block in the same stream, cdist.c:364-369), including an excluded all-N record."""
    from ccphylo_b200 import synth
    td = str(tmp_path)
    n, length = 45, 30000 + 11
    rows = synth.make_ascii(n, length, seed=9, snp=0.01, nrun=0.02)
    if flag == "3":
        rows[7, :] = ord("N")                      # excluded by the coverage gate (shared-mask parity needs no exclusions)
    plain = os.path.join(td, "aln.fsa")
    with open(plain, "wb") as f:
        for i in range(n):
            f.write(b">sample_%d some description\n" % i)
            for s0 in range(0, length, 70):
                f.write(rows[i, s0:s0 + 70].tobytes() + b"\n")
    gz = plain + ".gz"
    with open(plain, "rb") as f, gzip.open(gz, "wb") as g:
        g.write(f.read())
    ref = run([REF_BIN, "dist", "-i", plain, "-f", flag, "-o", os.path.join(td, "ref.phy"), "-n", os.path.join(td, "ref.num")], td)
    assert ref.returncode == 0, ref.stderr
    want = open(os.path.join(td, "ref.phy")).read()
    assert len(want) > 1000
    for tag, path, threads in (("plain-t7", plain, "7"), ("plain-t1", plain, "1"), ("gz-t7", gz, "7")):
        p = run([BIN, "dist", "-i", path, "-f", flag, "-t", threads, "-o", os.path.join(td, tag + ".phy"), "-n",
                 os.path.join(td, tag + ".num")], td)
        assert p.returncode == 0, p.stderr
        assert open(os.path.join(td, tag + ".phy")).read() == want, tag
        assert p.stderr == ref.stderr, tag


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
@pytest.mark.parametrize("proxi", ["4", "75"])
@pytest.mark.parametrize("flag", ["3", "1", "11", "9", "35", "33"])
@pytest.mark.parametrize("msa", [False, True], ids=["files", "msa"])
def test_proximity_against_the_reference_binary(built, tmp_path, msa, flag, proxi):
    """-P on the device: the per-sample builder (getIncPos / getIncPosInsig / getIncPosInsigPrune by -f 8 / 32),
    maskProxi per pair (-f 2) and the shared-mask accumulation print what the reference binary prints -- .phy,
    .num and the `# Included` counts on stderr -- from this driver and from the reference bound to the library."""
    from ccphylo_b200 import synth
    td = str(tmp_path)
    n, length = 21, 20000 + 7
    rows = synth.make_ascii(n, length, seed=int(flag) * 100 + int(proxi), snp=0.01, nrun=0.02)
    rows[:, -1] = ord("A")          # a SNP in the last column makes the reference's maskProxi write out of bounds
    if int(flag) & 2:
        rows[5, 50:] = ord("N")     # excluded by the coverage gate (the shared-mask mode mishandles exclusions, App. B #3)
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
    for tag, exe in (("reference", REF_BIN), ("driver", BIN), ("bound", REF_GPU)):
        if not os.path.exists(exe):
            continue
        phy, num = os.path.join(td, tag + ".phy"), os.path.join(td, tag + ".num")
        p = run([exe, "dist", "-f", flag, "-P", proxi, "-W", "1000", "-t", "3", "-o", phy, "-n", num] + inputs, td)
        assert p.returncode == 0, p.stderr[-2000:]
        outs[tag] = (open(phy).read(), open(num).read(), p.stderr)
    plain = run([REF_BIN, "dist", "-f", flag, "-W", "1000", "-o", os.path.join(td, "plain.phy")] + inputs, td)
    assert open(os.path.join(td, "plain.phy")).read() != outs["reference"][0], "-P changed nothing: weak inputs"
    for tag in outs:
        assert outs[tag][0] == outs["reference"][0], tag + ": .phy differs from the reference binary's"
        assert outs[tag][1] == outs["reference"][1], tag + ": .num differs from the reference binary's"
        assert outs[tag][2] == outs["reference"][2], tag + ": stderr differs from the reference binary's"


def test_fasta_gz_input_stdout_and_long_options(built, tmp_path):
    case = next(c for c in CASES if c["name"] == "c1_pair_W")
    td = str(tmp_path)
    files = []
    for nm, k in zip(case["names"], case["seq_ids"]):
        path = os.path.join(td, nm + ".gz")
        with gzip.open(path, "wt") as f:
            f.write(">other\nACGT\n>ref\n" + "\n".join(POOL[k][s:s + 10] for s in range(0, len(POOL[k]), 10)) + "\n")
        files.append(path)
    p = run([BIN, "dist", "--reference", "ref", "--flag=3", "--normalization_weight", "1000000", "--input"] + files, td)
    assert p.returncode == 0, p.stderr
    assert p.stdout == case["phy"].replace("s0\n", "s0.gz\n").replace("s1\t", "s1.gz\t").replace("s2\t", "s2.gz\t").replace("s3\t", "s3.gz\t")


def test_file_backed_matrices(built, tmp_path):
    """-H / --mmap with and without -T (matrix.c:116 ltdMatrixMinit, tmp.c:27 tmpF): the matrices live in unlinked
    temporary files; the output is the one of the in-memory run and nothing is left behind"""
    case = next(c for c in CASES if c["name"] == "c1_pair_W")
    td = str(tmp_path)
    files = []
    for nm, k in zip(case["names"], case["seq_ids"]):
        path = os.path.join(td, nm)
        with open(path, "w") as f:
            f.write(">ref\n" + "\n".join(POOL[k][s:s + 60] for s in range(0, len(POOL[k]), 60)) + "\n")
        files.append(path)
    os.mkdir(os.path.join(td, "scratch"))
    base = [BIN, "dist", "-r", "ref", "-f", "3", "-W", "1000000", "-n", "-"]
    plain = run(base + ["-i"] + files, td)
    assert plain.returncode == 0 and plain.stdout.startswith(case["phy"])
    for extra in (["-H"], ["-H", "-T", "scratch/"], ["--mmap", "--tmp", "scratch/pre"], ["-p", "-H"], ["-b", "1", "-H", "-T", "scratch/"]):
        a = run(base + extra + ["-i"] + files, td)
        b = run(base + [x for x in extra if x not in ("-H", "--mmap")] + ["-i"] + files, td)
        assert a.returncode == 0 and a.stdout == b.stdout and a.stderr == b.stderr, extra
    assert os.listdir(os.path.join(td, "scratch")) == []
    # .mat inputs go through the same allocation
    G_ = helpers.load_golden("mat_dist.json")
    mc = next(c for c in G_["cases"] if c["name"] == "rand_cos")
    paths = []
    for nm, k in zip(mc["names"], mc["text_ids"]):
        path = os.path.join(td, nm)
        with open(path, "w") as f:
            f.write(G_["pool"][k])
        paths.append(path)
    cmd = [BIN, "dist", "-r", mc["template"], "-i"] + paths + mc["args"]
    a = run(cmd + ["-H", "-T", "scratch/"], td)
    b = run(cmd, td)
    assert a.returncode == 0 and len(a.stdout) > 50 and a.stdout == b.stdout and a.stderr == b.stderr
    assert os.listdir(os.path.join(td, "scratch")) == []


def test_refused_options_and_errors(built, tmp_path):
    td = str(tmp_path)
    a = os.path.join(td, "a.fsa")
    with open(a, "w") as f:
        f.write(">ref\nACGT\n")
    p = run([BIN, "dist", "-r", "ref", "-i", a, os.path.join(td, "missing.fsa")], td)
    assert p.returncode != 0
    p = run([BIN, "dist", "--nope"], td)
    assert p.returncode == 1 and p.stderr == 'Unknown argument or option: "--nope"\n'


G = helpers.load_golden("mat_dist.json")


def _cmp_phy(text, ref_text, precision, exact):
    got, ref = helpers.parse_phy(text), helpers.parse_phy(ref_text)
    assert len(got) == len(ref)
    for (gn, gc), (rn, rc) in zip(got, ref):
        assert gn == rn
        gc, rc = np.array(gc), np.array(rc)
        if exact:
            assert np.array_equal(gc, rc)
        else:
            assert np.all(np.abs(gc - rc) <= 1.01 * 10.0 ** (-precision) + 1e-6 * np.abs(rc))


MAT_CLI = [c for c in G["cases"] if c["name"] in ("c8_cos", "c8_cos_W", "c8_cos_gz", "c8_cos_f5", "c8_chi2_x3", "rand_cos", "rand_z",
                                                   "rand_nl3", "rand_cos_W_E30", "rand_bc_t3_gz", "rand_l2_C90",
                                                   "overlap_fail", "union_cos_f5", "union_chi2")]


@pytest.mark.parametrize("case", MAT_CLI, ids=lambda c: c["name"])
def test_mat_golden(built, tmp_path, case):
    from test_mat_oracle_golden import mat_args
    td = str(tmp_path)
    o = mat_args(case["args"])
    texts = [G["pool"][k] for k in case["text_ids"]]
    phy, num = os.path.join(td, "o.phy"), os.path.join(td, "o.num")
    if case["mode"] == "files":
        files = []
        for nm, text in zip(case["names"], texts):
            path = os.path.join(td, nm + (".gz" if case["gz"] else ""))
            with (gzip.open(path, "wt") if case["gz"] else open(path, "w")) as f:
                f.write(text)
            files.append(path)
        cmd = [BIN, "dist", "-r", case["template"], "-i"] + files
    else:
        for nm, text in zip(case["names"], texts):
            with gzip.open(os.path.join(td, nm), "wt") as f:
                f.write(text)
        upath = os.path.join(td, "in.union")
        with open(upath, "w") as f:
            f.write(case["union"].replace("@TD@/", td + "/"))
        cmd = [BIN, "dist", "-i", upath]
    p = run(cmd + case["args"] + ["-o", phy, "-n", num], td)
    assert p.returncode == case["returncode"], p.stderr
    _cmp_phy(open(phy).read(), case["phy"], o["precision"], exact=False)
    _cmp_phy(open(num).read(), case["num"], o["precision"], exact=True)
    assert sorted(p.stderr.replace(td + "/", "").splitlines()) == sorted(case["stderr"].splitlines())
    # the comment lines of -f 4 come through as well
    assert [l for l in open(phy).read().splitlines() if l.startswith("#")] == [l for l in case["phy"].splitlines() if l.startswith("#")]


# ---- multi-GPU: `ccphylo-b200 dist` opens ccg_init_multi for plain FASTA runs; the output must not depend on the
# number of GPUs.  CCG_MULTI_DEVICES puts the members on the devices that exist (all on GPU 0 on a 1-GPU box). ----
@pytest.mark.parametrize("flag", ["3", "1"], ids=["pair", "shared-mask"])
def test_msa_on_several_gpus_is_byte_identical_to_one(built, tmp_path, flag):
    import torch
    from ccphylo_b200 import synth

    td = str(tmp_path)
    n, length = 230, 9000 + 7
    rows = synth.make_ascii(n, length, seed=42, snp=0.01, nrun=0.02)
    rows[17, :] = ord("N")                                        # an excluded record
    path = os.path.join(td, "msa.fsa")
    with open(path, "wb") as f:
        for k in range(n):
            f.write(b">smp%d\n" % k)
            for s in range(0, length, 70):
                f.write(rows[k, s:s + 70].tobytes() + b"\n")
    outs = {}
    ngpu = max(torch.cuda.device_count(), 1)
    for members in (1, 2, 5):
        env = dict(os.environ, CCPHYLO_GPUS=str(members), CCG_MULTI_FORCE="1", CCPHYLO_GPU_STATS="1",
                   CCG_MULTI_DEVICES=",".join(str(g % ngpu) for g in range(members)))
        phy, num = os.path.join(td, f"o{members}.phy"), os.path.join(td, f"o{members}.num")
        p = subprocess.run([BIN, "dist", "-i", path, "-f", flag, "-W", "1000000", "-o", phy, "-n", num, "-t", "4"], capture_output=True,
                           text=True, cwd=td, timeout=300, env=env)
        assert p.returncode == 0, p.stderr[-2000:]
        stats = [ln for ln in p.stderr.splitlines() if ln.startswith("# gpu-stats") and "GPU(s)" in ln]
        assert stats and (f"{members} of {members} GPU(s)" in stats[0] if members > 1 else "1 of 1 GPU(s)" in stats[0]), p.stderr
        if members > 1:
            assert "K split" in stats[0] and "k_finalize_group" in stats[0]
        text = "\n".join(ln for ln in p.stderr.splitlines() if not ln.startswith("# gpu-stats"))
        outs[members] = (open(phy).read(), open(num).read(), text)
    assert outs[1][0].count("\n") >= n and "# Excluded:\tsmp17" in outs[1][2]
    assert outs[2] == outs[1] and outs[5] == outs[1]


# ---- the pipe the north star names: ccphylo union | ccphylo dist | ccphylo tree, with OUR dist in the middle ----
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ccphylo")
NEWICK_C9 = "((a.mat.gz:0.003971940,d.mat.gz:1.000283475):0.991239454,c.mat.gz:0.965967545,b.mat.gz:0.011964387);"


def _write_c8_inputs(td):
    """four KMA count matrices of template `tmpl` (reference (ACGT)x5) and their .res files"""
    ref = "ACGT" * 5
    spec = {"a": (30, {}), "b": (40, {3: "A"}), "c": (25, {3: "A", 10: "T", 11: "N"}), "d": (20, {0: "-", 5: "C"})}
    hdr = ("#Template\tScore\tExpected\tTemplate_length\tTemplate_Identity\tTemplate_Coverage\tQuery_Identity\tQuery_Coverage\t"
           "Depth\tq_value\tp_value\n")
    for name, (depth, variants) in spec.items():
        rows = []
        for p, r in enumerate(ref):
            b = variants.get(p, r)
            c = [0] * 6
            c["ACGTN-".index(b)] = depth
            if b in "ACGT":
                c[("ACGT".index(b) + 1) % 4] += p % 3
            rows.append(r + "\t" + "\t".join(map(str, c)))
        with gzip.open(os.path.join(td, name + ".mat.gz"), "wt") as f:
            f.write("#tmpl\n" + "\n".join(rows) + "\n\n")
        with open(os.path.join(td, name + ".res"), "w") as f:
            f.write(hdr + "tmpl\t1000\t10\t20\t100.00\t100.00\t100.00\t100.00\t%d.00\t500.00\t1.0e-26\n" % depth)
    return [name + ".res" for name in spec]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")
def test_union_dist_tree_pipe_with_the_gpu_dist_in_the_middle(built, tmp_path):
    td = str(tmp_path)
    res = _write_c8_inputs(td)
    union = subprocess.run([REF_BIN, "union", "-i"] + res, capture_output=True, cwd=td, timeout=60)
    assert union.returncode == 0 and union.stdout.startswith(b"4\ta.res")
    trees = {}
    for who, binary in (("reference", REF_BIN), ("gpu", BIN)):
        dist = subprocess.run([binary, "dist", "-f", "5"], input=union.stdout, capture_output=True, cwd=td, timeout=300)
        assert dist.returncode == 0, dist.stderr.decode()[-1000:]
        tree = subprocess.run([REF_BIN, "tree"], input=dist.stdout, capture_output=True, cwd=td, timeout=60)
        assert tree.returncode == 0, tree.stderr.decode()[-1000:]
        trees[who] = (dist.stdout.decode(), tree.stdout.decode())
    assert trees["reference"][1].strip() == ">tmpl" + NEWICK_C9
    assert trees["gpu"][1] == trees["reference"][1]                       # the same Newick through our dist
    assert trees["gpu"][0] == trees["reference"][0]                       # and the same Phylip text in between
