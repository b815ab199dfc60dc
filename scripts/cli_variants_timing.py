#!/usr/bin/env python
"""-V (variant listing) through the command line, the reference binary beside the driver on the same files: wall time of
the whole command (parse + compare + list + print) and byte equality of the listing (against the reference at -t 1, whose
order is the deterministic one).  Never a bench value; the numbers go to profiles/.

    python scripts/cli_variants_timing.py [--samples 128] [--length 1000000] [--proxi 0] [--out file.json]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccphylo_b200 import synth  # noqa: E402

BIN = os.path.join(ROOT, "ccphylo_b200", "bin", "ccphylo-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "ccphylo")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=128)
    ap.add_argument("--length", type=int, default=1_000_000)
    ap.add_argument("--proxi", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    n, L = a.samples, a.length
    threads = os.cpu_count() or 1
    res = {"what": "dist -f 3 -V through the command line, whole-command wall time", "samples": n, "length": L, "proxi": a.proxi,
           "pairs": n * (n - 1) // 2, "host_threads": threads}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        rows = synth.make_ascii(n, L, seed=11, snp=0.001, nrun=0.002)
        files = []
        for i in range(n):
            fp = os.path.join(td, f"s{i:03d}.fsa")
            synth.write_fasta(fp, rows[i], header="ref", width=60)
            files.append(fp)
        prox = ["-P", str(a.proxi)] if a.proxi else []

        def run(tag, exe, t):
            var, phy, num = (os.path.join(td, tag + e) for e in (".var", ".phy", ".num"))
            t0 = time.perf_counter()
            p = subprocess.run([exe, "dist", "-f", "3", "-t", str(t), "-V", var, "-o", phy, "-n", num] + prox + ["-r", "ref", "-i"] + files,
                               capture_output=True, text=True)
            dt = time.perf_counter() - t0
            assert p.returncode == 0, p.stderr[-2000:]
            return dt, var, phy

        t_drv, v_drv, p_drv = run("driver", BIN, threads)
        t_drv2, _, _ = run("driver", BIN, threads)
        res["driver_s"] = min(t_drv, t_drv2)
        lines = sum(1 for _ in open(v_drv, "rb"))
        res["variant_lines"] = lines
        res["listing_bytes"] = os.path.getsize(v_drv)
        if os.path.exists(REF):
            t_ref, v_ref, p_ref = run("reference_mt", REF, threads)
            res["reference_s"] = t_ref
            res["reference_threads"] = threads
            t_ref1, v_ref1, p_ref1 = run("reference_t1", REF, 1)
            res["reference_t1_s"] = t_ref1
            res["listing_identical_to_reference_t1"] = open(v_ref1, "rb").read() == open(v_drv, "rb").read()
            res["phy_identical"] = open(p_ref1, "rb").read() == open(p_drv, "rb").read()
            res["base_cmp_per_s"] = {"driver": res["pairs"] * L / res["driver_s"], "reference": res["pairs"] * L / t_ref}
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            f.write(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
