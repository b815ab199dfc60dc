"""The count-matrix (.mat) oracle (oracle/mat_oracle.c) against the golden text the unmodified
reference binary printed (tests/golden/mat_dist.json, scripts/make_golden_mat.py): every -d method,
-W, -E, -C, -x, gz input, an excluded sample, a pair without sufficient overlap, union input.
The oracle sums in the reference's order, so the Phylip text must match byte for byte."""
import numpy as np
import pytest

import helpers
import oracle

G = helpers.load_golden("mat_dist.json")
POOL, CASES = G["pool"], G["cases"]


def mat_args(args):
    o = {"method": "cos", "norm": 0, "min_depth": 15, "min_length": 1, "min_cov": 0.5, "precision": 9, "flag": 1}
    k = 0
    while k < len(args):
        a, v = args[k], args[k + 1]
        if a == "-d":
            o["method"] = v
        elif a == "-W":
            o["norm"] = int(v)
        elif a == "-E":
            o["min_depth"] = int(float(v))
        elif a == "-C":
            o["min_cov"] = float(v) / 100
        elif a == "-L":
            o["min_length"] = int(v)
        elif a == "-x":
            o["precision"] = int(v)
        elif a == "-f":
            o["flag"] = int(v)
        elif a == "-t":
            pass
        else:
            raise ValueError(a)
        k += 2
    return o


def expected_text(names, texts, template, o, threaded=True):
    """Replay of ltdMatrixThrd / ltdMatrix_get with the oracle as the compare step -> (phy, num, stderr)."""
    parsed = [helpers.parse_mat(t, template) for t in texts]
    err, keep = [], []
    for nm, pm in zip(names, parsed):
        if pm is None:
            err.append(f'Template ("{template}") is not included in:\t{nm}\n')
        elif not helpers.mat_sample_gate(pm[1], o["min_depth"], o["min_length"], o["min_cov"]):
            err.append(f'Template ("{template}") did not exceed threshold for inclusion:\t{nm}\n')
        else:
            keep.append((nm, pm))
    if len(keep) < 2:
        return "", "", "".join(err)
    lmax = max(len(pm[1]) for _, pm in keep)
    counts = np.zeros((len(keep), lmax, 6), np.uint16)
    totals = np.zeros((len(keep), lmax), np.uint32)
    lens = np.zeros(len(keep), np.int32)
    for k, (_, (c, t)) in enumerate(keep):
        counts[k, :len(t)], totals[k, :len(t)], lens[k] = c, t, len(t)
    D, N, dn = oracle.mat_matrix(counts, totals, lens, method=o["method"], norm=o["norm"], min_depth=o["min_depth"],
                                 min_length=o["min_length"], min_cov=o["min_cov"])
    kept = [nm for nm, _ in keep]
    k = 0
    for r in range(dn):
        for c in range(r):
            if D[k] == -1.0 and N[k] == 0:
                sep = "\t" if threaded else ", "
                err.append(f"No sufficient overlap between samples:\t{kept[r]}{sep}{kept[c]}\n")
            k += 1
    phy = helpers.format_phy_cells(kept, D, o["precision"], template, o["flag"])
    num = helpers.format_phy_cells(kept, N, o["precision"], template, o["flag"])
    return phy, num, "".join(err)


@pytest.mark.parametrize("case", [c for c in CASES if c["mode"] == "files"], ids=lambda c: c["name"])
def test_files_mode_text(case):
    o = mat_args(case["args"])
    texts = [POOL[k] for k in case["text_ids"]]
    names = [nm + (".gz" if case["gz"] else "") for nm in case["names"]]
    phy, num, err = expected_text(names, texts, case["template"], o)
    assert case["returncode"] == 0
    assert phy == case["phy"]
    assert num == case["num"]
    assert sorted(err.splitlines()) == sorted(case["stderr"].splitlines())


@pytest.mark.parametrize("case", [c for c in CASES if c["mode"] == "union"], ids=lambda c: c["name"])
def test_union_mode_text(case):
    o = mat_args(case["args"])
    texts = [POOL[k] for k in case["text_ids"]]
    phy = num = ""
    for t in case["templates"]:
        p, q, _ = expected_text(case["names"], texts, t, o, threaded=False)
        phy += p
        num += q
    assert phy == case["phy"]
    assert num == case["num"]
