#!/usr/bin/env python
"""Generate tests/golden/*.json by running the UNMODIFIED reference binary
(oracle/_ref/ccphylo, built by oracle/Makefile from /root/reference).

Runs only in the authoring container (the GPU box has no /root/reference and
only replays the committed fixtures).  Each case records the input sequences,
the `ccphylo dist` arguments, and the reference's .phy / .num / stderr text.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ccphylo")
OUT = os.path.join(ROOT, "tests", "golden")


def run_case(name, seqs, args, names=None, msa=False):
    names = names or [f"s{i}" for i in range(len(seqs))]
    with tempfile.TemporaryDirectory() as td:
        files = []
        if msa:
            path = os.path.join(td, "msa.fsa")
            with open(path, "w") as f:
                for nm, s in zip(names, seqs):
                    f.write(f">{nm}\n{s}\n")
            files = [path]
            cmd = [REF_BIN, "dist", "-i", path]
        else:
            for nm, s in zip(names, seqs):
                path = os.path.join(td, nm)
                with open(path, "w") as f:
                    f.write(f">ref\n{s}\n")
                files.append(path)
            cmd = [REF_BIN, "dist", "-r", "ref", "-i"] + files
        phy, num = os.path.join(td, "o.phy"), os.path.join(td, "o.num")
        cmd += list(args) + ["-o", phy, "-n", num]
        p = subprocess.run(cmd, capture_output=True, text=True)
        text = lambda q: open(q).read() if os.path.exists(q) else ""
        return {"name": name, "seqs": list(seqs), "names": names, "args": list(args), "msa": msa,
                "returncode": p.returncode, "phy": text(phy), "num": text(num),
                "stderr": p.stderr.replace(td + "/", "")}


def rand_seq(rng, length, ref=None, snp=0.05, unk=0.05, low=0.03):
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    base = rng.integers(0, 4, size=length) if ref is None else ref.copy()
    sub = rng.random(length) < snp
    base[sub] = (base[sub] + rng.integers(1, 4, size=int(sub.sum()))) & 3
    row = alphabet[base]
    odd = np.frombuffer(b"N-RYSWKMBDHVXn", dtype=np.uint8)
    u = rng.random(length) < unk
    row = np.where(u, odd[rng.integers(0, len(odd), size=length)], row)
    lo = rng.random(length) < low
    row = np.where(lo & ~u, np.frombuffer(b"acgt", dtype=np.uint8)[base], row)
    return row.tobytes().decode(), base


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = []
    # ---- SURVEY.md Appendix C #1, #2, #10, #11 inputs ----
    s0 = "ACGT" * 10 + "AC"
    s2 = "ACGAACGTACGTACGTNCGTACGTACGTACGTACGTACGTAA"
    s3 = "TCGAACGTACGTaCGTNCGTACGTACGT-CGTACGTACGTAG"
    c1 = [s0, s0, s2, s3]
    for nm, args in [("c1_pair", ["-f", "3"]), ("c1_pair_W", ["-f", "3", "-W", "1000000"]),
                     ("c1_pair_lower", ["-f", "11"]), ("c1_global", ["-f", "1"]),
                     ("c1_global_W", ["-f", "1", "-W", "1000"]),
                     ("c1_pair_L41", ["-f", "3", "-L", "41"]),
                     ("c1_float_W", ["-f", "3", "-W", "1000000", "-p"]),
                     ("c1_short_W", ["-f", "3", "-W", "1000000", "-s", "100"]),
                     ("c1_byte_W", ["-f", "3", "-W", "1000000", "-b", "0.001"]),
                     ("c1_short", ["-f", "3", "-s", "100"]),
                     ("c1_pair_C99", ["-f", "3", "-C", "99"])]:
        cases.append(run_case(nm, c1, args))
    # ---- Appendix C #6: excluded sample ----
    a = "A" * 64
    b = "N" * 64
    c = a[:5] + "C" + a[6:]
    d = a[:5] + "G" + a[6:9] + "T" + a[10:]
    cases.append(run_case("c6_excluded_pair", [a, b, c, d], ["-f", "3"], names=["xa", "xb", "xc", "xd"]))
    # ---- Appendix C #7: MSA mode prints D then N into the .phy stream ----
    cases.append(run_case("c7_msa_pair", c1, ["-f", "3"], names=["smp0", "smp1", "smp2", "smp3"], msa=True))
    # ---- randomised sets, tail lengths around the 32- and 128-base boundaries ----
    rng = np.random.default_rng(20261018)
    for length in (1, 31, 32, 33, 127, 128, 129, 777, 4100):
        n = 9
        _, ref = rand_seq(rng, length, snp=0.0, unk=0.0, low=0.0)
        seqs = [rand_seq(rng, length, ref=ref)[0] for _ in range(n)]
        if length >= 64:
            seqs[4] = "N" * (length - 3) + seqs[4][-3:]          # excluded by the 50 % coverage gate
        for tag, args in [("pair", ["-f", "3"]), ("pairW", ["-f", "3", "-W", "100000"]),
                          ("pair_low", ["-f", "11"]), ("pair_float", ["-f", "3", "-W", "1000", "-p"]),
                          ("pair_short", ["-f", "3", "-W", "1000", "-s", "10"])]:
            cases.append(run_case(f"rand_L{length}_{tag}", seqs, args))
        # global mode parity only without exclusions (App. B #3)
        gseqs = [s for k, s in enumerate(seqs) if k != 4]
        for tag, args in [("global", ["-f", "1", "-C", "0"]), ("globalW", ["-f", "1", "-W", "1000", "-C", "0"])]:
            cases.append(run_case(f"rand_L{length}_{tag}", gseqs, args))
    # store every distinct sequence once; cases refer to them by index
    pool, index = [], {}
    for cs in cases:
        ids = []
        for sq in cs.pop("seqs"):
            if sq not in index:
                index[sq] = len(pool)
                pool.append(sq)
            ids.append(index[sq])
        cs["seq_ids"] = ids
    with open(os.path.join(OUT, "fasta_dist.json"), "w") as f:
        json.dump({"generator": "scripts/make_golden.py", "reference": "ccphylo v0.8.5 (oracle/_ref/ccphylo)",
                   "pool": pool, "cases": cases}, f, indent=0)
    print(f"{len(cases)} cases written")


if __name__ == "__main__":
    main()
