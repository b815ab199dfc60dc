"""The CPU oracle against the reference BINARY on random inputs (the golden fixture holds 76 hand-picked cases; this
sweeps random sample sets, gates and cell types each run of the suite): `oracle/_ref/ccphylo dist` prints the .phy /
.num / stderr text, helpers.replay renders the oracle's result the same way, and the bytes must agree.  Runs where the
reference was compiled (this container); the GPU box only has the committed fixtures."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import fuzz_cli  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.exists(fuzz_cli.REF_BIN), reason="oracle/_ref/ccphylo was not built (needs /root/reference)")


def _case(idx):
    rng = np.random.default_rng([77, idx])
    n = int(rng.integers(2, 14))
    length = int(rng.choice(fuzz_cli.LENGTHS[:17])) if rng.random() < 0.6 else int(rng.integers(1, 3000))
    rows = fuzz_cli.make_rows(rng, n, length)
    pair = rng.random() < 0.7
    flag = (2 if pair else 0) | 1 | int(rng.choice([0, 8]))
    args = ["-f", str(flag)]
    if pair:
        if length >= 8 and rng.random() < 0.3:
            rows[int(rng.integers(0, n))][: length - int(rng.integers(0, max(1, length // 3)))] = ord("N")
        c = rng.choice(["", "0", "30", "90"])
        if c:
            args += ["-C", str(c)]
        if rng.random() < 0.3:
            args += ["-L", str(int(rng.integers(0, length + 2)))]
    else:
        args += ["-C", "0"]                   # shared-mask parity only without exclusions (App. B #3)
        for r in rows:                        # ... so every sample keeps at least one known base
            if not np.isin(r, np.frombuffer(b"ACGT" + (b"acgt" if flag & 8 else b""), dtype=np.uint8)).any():
                r[0] = ord("A")
    w = rng.choice(["", "7", "1000", "1000000"])
    if w:
        args += ["-W", str(w)]
    cell = rng.choice(["", "", "-p", "-s", "-b"])
    if cell == "-p":
        args += ["-p"]
    elif cell:
        args += [str(cell), str(rng.choice(["0.001", "0.5", "1", "10", "100"]))]
    return rows, args


@pytest.mark.parametrize("idx", range(200))
def test_oracle_prints_what_the_reference_binary_prints(built, tmp_path, idx):
    rows, args = _case(idx)
    td = str(tmp_path)
    names, files = [], []
    for i, row in enumerate(rows):
        names.append("s%02d.fsa" % i)
        files.append(os.path.join(td, names[-1]))
        with open(files[-1], "wb") as f:
            f.write(b">ref\n" + row.tobytes() + b"\n")
    phy, num = os.path.join(td, "o.phy"), os.path.join(td, "o.num")
    p = subprocess.run([fuzz_cli.REF_BIN, "dist", "-r", "ref", "-i"] + files + args + ["-o", phy, "-n", num], capture_output=True,
                       text=True, timeout=60)
    if p.returncode < 0:
        pytest.skip("the reference binary died with signal %d on this input" % -p.returncode)
    assert p.returncode == 0, p.stderr
    pool = [r.tobytes().decode() for r in rows]
    case = {"args": args, "seq_ids": list(range(len(rows))), "names": names}
    ophy, onum, oerr = helpers.replay(case, pool, helpers.oracle_backend)
    junk = ("Error: 11 (", "Will continue with ", "Adjustning number of nodes")
    ref_err = "".join(ln + "\n" for ln in p.stderr.replace(td + "/", "").split("\n")[:-1] if not ln.startswith(junk))
    assert oerr == ref_err
    assert ophy == open(phy).read()
    assert onum == open(num).read()
