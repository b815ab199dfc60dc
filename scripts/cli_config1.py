#!/usr/bin/env python
"""Wall time of the host driver on BASELINE configs[0]/[1]-shaped inputs (n files x 5 Mbp FASTA), with phase timings.

    python scripts/cli_config1.py [n] [length] [options]    options: "plain" (default), "P" (-P 10), "y" (-y dam/dcm motifs)
With the reference binary present (n <= 256) both run every option set and the outputs are compared byte for byte."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccphylo_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
length = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
with tempfile.TemporaryDirectory() as td:
    files = []
    base = synth.make_ascii(min(n, 16), length, seed=1)
    for i in range(n):
        path = os.path.join(td, f"s{i:04d}.fsa")
        synth.write_fasta(path, base[i % len(base)], header="ref", width=60)
        files.append(path)
    env = dict(os.environ, CCPHYLO_GPU_STATS="1")
    motifs = os.path.join(td, "motifs.fsa")
    with open(motifs, "w") as f:
        f.write(">dam\ngAtc\n>dcm\ncCwgg\n")
    sets = {"plain": [], "P": ["-P", "10"], "y": ["-y", motifs]}
    for tag in (sys.argv[3:] or ["plain"]):
        outs = {}
        for exe in ("ccphylo_b200/bin/ccphylo-b200", "oracle/_ref/ccphylo"):
            if not os.path.exists(os.path.join(ROOT, exe)) or (n > 256 and "oracle" in exe):
                continue
            phy, num = os.path.join(td, "o.phy"), os.path.join(td, "o.num")
            t0 = time.perf_counter()
            p = subprocess.run([os.path.join(ROOT, exe), "dist", "-r", "ref", "-f", "3", "-t", str(os.cpu_count()), "-i"] + files +
                               sets[tag] + ["-o", phy, "-n", num], capture_output=True, text=True, env=env)
            dt = time.perf_counter() - t0
            print(f"[{tag}] {exe}: rc={p.returncode} {dt:.2f} s wall, n={n}")
            print("".join(l + "\n" for l in p.stderr.splitlines() if "gpu-stats" in l or "rror" in l), end="")
            outs[exe] = (open(phy).read(), open(num).read(), "".join(l + "\n" for l in p.stderr.splitlines() if "gpu-stats" not in l))
        if len(outs) == 2:
            a, b = outs.values()
            print(f"[{tag}] outputs byte-identical: phy {a[0] == b[0]}, num {a[1] == b[1]}, stderr {a[2] == b[2]}")
