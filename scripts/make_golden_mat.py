#!/usr/bin/env python
"""Generate tests/golden/mat_dist.json by running the UNMODIFIED reference binary
(oracle/_ref/ccphylo) on small KMA count matrices (.mat): every -d method, -W, -E, gz input,
an excluded low-depth sample, a pair without sufficient overlap, and the union input mode.
Runs only in the authoring container; the GPU box replays the committed fixture."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ccphylo")
OUT = os.path.join(ROOT, "tests", "golden")
METHODS = ["cos", "z", "chi2", "nchi2", "c", "nc", "p", "np", "bc", "nbc", "l1", "l2", "linf", "l3", "nl1", "nl2", "nlinf",
           "nl3"]


def survey_sample(ref, depth, variants):
    """depth on the called base plus (p mod 3) on the next base (cyclic) for ACGT calls."""
    rows = []
    for p, b in enumerate(ref):
        call = variants.get(p, b)
        c = [0] * 6                                   # file order A C G T N -
        idx = "ACGTN-".index(call)
        c[idx] = depth
        if call in "ACGT":
            c[(idx + 1) % 4] += p % 3
        rows.append(c)
    return rows


def random_sample(rng, ref, depth, err, low_frac, snp):
    rows = []
    for p, b in enumerate(ref):
        call = "ACGT".index(b)
        if rng.random() < snp:
            call = (call + int(rng.integers(1, 4))) % 4
        d = int(rng.poisson(depth))
        if rng.random() < low_frac:
            d = int(rng.integers(0, 6))
        c = [int(rng.poisson(err)) for _ in range(4)] + [int(rng.poisson(err / 3)), int(rng.poisson(err / 3))]
        c[call] += d
        rows.append(c)
    return rows


def run_case(name, template, ref, samples, args, gz=False, other=None, names=None):
    names = names or ["%c.mat" % (ord("a") + k) for k in range(len(samples))]
    texts = []
    with tempfile.TemporaryDirectory() as td:
        files = []
        for nm, rows in zip(names, samples):
            text = ""
            if other:                                   # another template in front: the loader must skip it
                text += helpers.mat_text(other[0], other[1], other[2])
            text += helpers.mat_text(template, ref, rows)
            texts.append(text)
            path = os.path.join(td, nm + (".gz" if gz else ""))
            if gz:
                with gzip.open(path, "wt") as f:
                    f.write(text)
            else:
                with open(path, "w") as f:
                    f.write(text)
            files.append(path)
        phy, num = os.path.join(td, "o.phy"), os.path.join(td, "o.num")
        cmd = [REF_BIN, "dist", "-r", template, "-i"] + files + list(args) + ["-o", phy, "-n", num]
        p = subprocess.run(cmd, capture_output=True, text=True)
        rd = lambda q: open(q).read() if os.path.exists(q) else ""
        return {"name": name, "mode": "files", "template": template, "names": names, "gz": gz, "texts": texts,
                "args": list(args), "returncode": p.returncode, "phy": rd(phy), "num": rd(num),
                "stderr": p.stderr.replace(td + "/", "")}


def run_union(name, templates, refs, per_template_samples, args):
    """Union input: header 'N\\tfile...', rows 'template\\tnum\\tidx...'; sample files hold all templates."""
    n = len(per_template_samples[0])
    names = ["%c.mat.gz" % (ord("a") + k) for k in range(n)]
    with tempfile.TemporaryDirectory() as td:
        texts = []
        for k in range(n):
            text = "".join(helpers.mat_text(t, r, s[k]) for t, r, s in zip(templates, refs, per_template_samples))
            texts.append(text)
            with gzip.open(os.path.join(td, names[k]), "wt") as f:
                f.write(text)
        # the union file names the KMA result files (*.res); dist rewrites the suffix (dist.c:223-250)
        union = "%d\t" % n + "\t".join(os.path.join(td, nm.replace(".mat.gz", ".res")) for nm in names) + "\n"
        for t in templates:
            union += t + "\t%d\t" % n + "\t".join(str(k) for k in range(n)) + "\n"
        upath = os.path.join(td, "in.union")
        with open(upath, "w") as f:
            f.write(union)
        phy, num = os.path.join(td, "o.phy"), os.path.join(td, "o.num")
        cmd = [REF_BIN, "dist", "-i", upath] + list(args) + ["-o", phy, "-n", num]
        p = subprocess.run(cmd, capture_output=True, text=True)
        rd = lambda q: open(q).read() if os.path.exists(q) else ""
        return {"name": name, "mode": "union", "templates": templates, "names": names, "texts": texts,
                "union": union.replace(td + "/", "@TD@/"), "args": list(args), "returncode": p.returncode,
                "phy": rd(phy), "num": rd(num), "stderr": p.stderr.replace(td + "/", "")}


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = []
    # ---- SURVEY App. C #8 (without the '-' reference row of sample d: the reference corrupts those, App. B #12) ----
    ref = "ACGT" * 5
    s = [survey_sample(ref, 30, {}), survey_sample(ref, 40, {3: "A"}), survey_sample(ref, 25, {3: "A", 10: "T", 11: "N"}),
         survey_sample(ref, 20, {5: "C"})]
    for m in METHODS:
        cases.append(run_case(f"c8_{m}", "tmpl", ref, s, ["-d", m]))
    cases.append(run_case("c8_cos_W", "tmpl", ref, s, ["-d", "cos", "-W", "1000"]))
    cases.append(run_case("c8_cos_gz", "tmpl", ref, s, ["-d", "cos"], gz=True))
    cases.append(run_case("c8_cos_f5", "tmpl", ref, s, ["-d", "cos", "-f", "5"]))
    cases.append(run_case("c8_chi2_x3", "tmpl", ref, s, ["-d", "chi2", "-x", "3"]))
    # ---- randomised: 7 samples x 333 positions, another template in front, one low-depth (excluded) sample ----
    rng = np.random.default_rng(424242)
    L = 333
    ref = "".join("ACGT"[k] for k in rng.integers(0, 4, size=L))
    oref = "".join("ACGT"[k] for k in rng.integers(0, 4, size=40))
    other = ("other_template", oref, random_sample(rng, oref, 30, 0.2, 0.0, 0.0))
    samples = [random_sample(rng, ref, 40, 0.3, 0.01, 0.02) for _ in range(7)]
    samples[3] = random_sample(rng, ref, 40, 0.3, 0.8, 0.02)          # mostly below the depth gate
    for m in METHODS:
        cases.append(run_case(f"rand_{m}", "tmpl_1", ref, samples, ["-d", m], other=other))
    cases.append(run_case("rand_cos_W_E30", "tmpl_1", ref, samples, ["-d", "cos", "-W", "1000000", "-E", "30"], other=other))
    cases.append(run_case("rand_bc_t3_gz", "tmpl_1", ref, samples, ["-d", "bc", "-t", "3"], gz=True, other=other))
    cases.append(run_case("rand_l2_C90", "tmpl_1", ref, samples, ["-d", "l2", "-C", "90"], other=other))
    # ---- a pair without sufficient overlap: two samples whose well-covered halves are disjoint ----
    half_a = random_sample(rng, ref, 40, 0.3, 0.0, 0.0)
    half_b = random_sample(rng, ref, 40, 0.3, 0.0, 0.0)
    for p in range(L):
        if p < int(L * 0.45):
            half_b[p] = [1, 0, 0, 0, 0, 0]
        elif p >= int(L * 0.55):
            half_a[p] = [0, 1, 0, 0, 0, 0]
    cases.append(run_case("overlap_fail", "tmpl_1", ref, [samples[0], half_a, half_b, samples[1]], ["-d", "cos", "-C", "50"]))
    # ---- union input, two templates ----
    ref2 = "".join("ACGT"[k] for k in rng.integers(0, 4, size=120))
    s1 = [random_sample(rng, ref, 35, 0.3, 0.01, 0.02) for _ in range(4)]
    s2 = [random_sample(rng, ref2, 35, 0.3, 0.01, 0.05) for _ in range(4)]
    cases.append(run_union("union_cos_f5", ["tmpl_1", "tmpl_2"], [ref, ref2], [s1, s2], ["-f", "5"]))
    cases.append(run_union("union_chi2", ["tmpl_1", "tmpl_2"], [ref, ref2], [s1, s2], ["-d", "chi2"]))
    # every distinct sample file once; the cases refer to them by index
    pool, index = [], {}
    for c in cases:
        ids = []
        for t in c.pop("texts"):
            if t not in index:
                index[t] = len(pool)
                pool.append(t)
            ids.append(index[t])
        c["text_ids"] = ids
    with open(os.path.join(OUT, "mat_dist.json"), "w") as f:
        json.dump({"generator": "scripts/make_golden_mat.py", "reference": "ccphylo v0.8.5 (oracle/_ref/ccphylo)",
                   "pool": pool, "cases": cases}, f, indent=0)
    for c in cases:
        if c["returncode"] != 0 or not c["phy"]:
            print("note:", c["name"], "rc", c["returncode"], "stderr:", c["stderr"][:200].replace("\n", " | "))
    print(f"{len(cases)} cases written")


if __name__ == "__main__":
    main()
