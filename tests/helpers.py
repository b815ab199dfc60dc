"""Shared test helpers: Phylip parsing, packed-layout unpacking, golden loading."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def parse_phy(text):
    """Parse concatenated lower-triangular Phylip blocks as the reference prints them
    (phy.c:59-123): returns a list of (names, packed_cells list of float)."""
    lines = text.split("\n")
    blocks, k = [], 0
    while k < len(lines):
        line = lines[k].strip()
        if not line or line.startswith("#"):
            k += 1
            continue
        n = int(line)
        names, cells = [], []
        for r in range(n):
            parts = lines[k + 1 + r].split("\t")
            names.append(parts[0])
            cells.extend(float(x) for x in parts[1:] if x != "")
        blocks.append((names, cells))
        k += n + 1
    return blocks


def full_from_packed(cells, dn):
    M = np.zeros((dn, dn), dtype=np.asarray(cells).dtype)
    k = 0
    for r in range(1, dn):
        for c in range(r):
            M[r, c] = M[c, r] = cells[k]
            k += 1
    return M


# ----------------------------------------------------------------------------
# golden-case replay: host-side flow of ltdFsaMatrix_get (cdist.c:36-194) using
# the ORACLE for translate / pack / mask, and a pluggable backend for the hot
# path (the oracle itself, the unmodified reference, or the CUDA library).
# ----------------------------------------------------------------------------
def parse_args(args):
    o = {"flag": 1, "norm": 0, "min_len": 1, "min_cov": 0.5, "elem": 8, "scale": 1.0}
    k = 0
    while k < len(args):
        a = args[k]
        if a == "-f":
            o["flag"] = int(args[k + 1]); k += 2
        elif a == "-W":
            o["norm"] = int(args[k + 1]); k += 2
        elif a == "-L":
            o["min_len"] = int(args[k + 1]); k += 2
        elif a == "-C":
            o["min_cov"] = float(args[k + 1]) / 100.0; k += 2      # dist.c: stored /100
        elif a == "-p":
            o["elem"] = 4; k += 1
        elif a == "-s":
            o["elem"] = 2; o["scale"] = float(args[k + 1]); k += 2
        elif a == "-b":
            o["elem"] = 1; o["scale"] = float(args[k + 1]); k += 2
        else:
            raise ValueError(a)
    return o


def prepare_case(case, pool):
    """Returns dict with packed inputs, include flags, thresholds and the stderr the reference prints."""
    import oracle

    o = parse_args(case["args"])
    seqs_txt = [pool[k] for k in case["seq_ids"]]
    codes = [oracle.translate(s.encode(), o["flag"]) for s in seqs_txt]
    L = len(codes[0])
    assert all(len(c) == L for c in codes)
    W = oracle.words(L)
    n = len(codes)
    min_len = o["min_len"]
    if min_len < o["min_cov"] * L:
        min_len = int(o["min_cov"] * L)                                # cdist.c:116, truncating
    pair = bool(o["flag"] & 2)
    seqs = np.zeros((n, max(W, 1)), dtype=np.uint64)
    masks = np.zeros((n, max(W, 1)), dtype=np.uint32)
    include = np.zeros(n, dtype=np.uint8)
    log = []
    for i, c in enumerate(codes):
        seqs[i, :W], unknown = oracle.pack(c)
        masks[i, :W], inc = oracle.known_mask(c)
        assert inc == L - unknown
        include[i] = 1 if inc >= min_len else 0
        log.append(f"# {'Included' if include[i] else 'Excluded'}:\t{case['names'][i]}\t( {inc} / {L} )\n")
    gmask = None
    if not pair and include.any():
        gmask = np.zeros((1, max(W, 1)), dtype=np.uint32)
        gmask[0, :W] = oracle.global_mask(np.stack(codes), include)
    return {"opts": o, "L": L, "n": n, "pair": pair, "seqs": seqs, "masks": masks, "include": include,
            "min_len": min_len, "gmask": gmask, "log": "".join(log), "names": case["names"]}


def fmt_cell(v, precision=9):
    v = float(v)
    if v == int(v):
        return "%d" % int(v)
    return "%.*f" % (precision, v)


def format_phy(names, include, cells, dn, elem, scale):
    """printphy (phy.c:59-123) with format flag bit 1 (relaxed names)."""
    out = ["%10d\n" % dn]
    k = 0
    r = 0
    for i, nm in enumerate(names):
        if not include[i]:
            continue
        row = [nm]
        for _ in range(r):
            v = cells[k]
            k += 1
            if elem <= 2:
                v = float(v) / scale                                     # uctod, bytescale.h:23
            row.append(fmt_cell(v))
        out.append("\t".join(row) + "\n")
        r += 1
    return "".join(out)


def replay(case, pool, backend):
    """backend(prep) -> (D, N or None, dn, global_inc or None).  Returns (phy, num, stderr) text."""
    prep = prepare_case(case, pool)
    o = prep["opts"]
    D, N, dn, ginc = backend(prep)
    err = prep["log"]
    if not prep["pair"]:
        err += f"# {ginc} / {prep['L']} bases included in distance matrix.\n"
    phy = format_phy(prep["names"], prep["include"], D, dn, o["elem"], o["scale"]) if dn > 1 else ""
    num = ""
    if prep["pair"] and dn > 1:
        num = format_phy(prep["names"], prep["include"], N, dn, o["elem"], o["scale"])
    return phy, num, err


def oracle_backend(prep):
    import oracle

    o = prep["opts"]
    if prep["pair"]:
        D, N, dn = oracle.fsa_cmp_pair(prep["seqs"], prep["masks"], prep["include"], prep["L"], norm=o["norm"],
                                       min_length=prep["min_len"], min_cov=o["min_cov"], elem_size=o["elem"],
                                       byte_scale=o["scale"])
        return D, N, dn, None
    D, dn, ginc = oracle.fsa_cmp_global(prep["seqs"], prep["gmask"][0], prep["include"], prep["L"], norm=o["norm"],
                                        elem_size=o["elem"], byte_scale=o["scale"])
    return D, None, dn, ginc


def gpu_backend(prep):
    from ccphylo_b200 import api

    o = prep["opts"]
    D, N, dn, ginc = api.fsa_cmp_thread_out(prep["seqs"], prep["include"],
                                            prep["masks"] if prep["pair"] else prep["gmask"], prep["L"],
                                            pair=prep["pair"], norm=o["norm"], min_length=prep["min_len"],
                                            min_cov=o["min_cov"], elem_size=o["elem"], byte_scale=o["scale"])
    return D, N, dn, ginc


def golden_cases(skip_buggy_global=True):
    g = load_golden("fasta_dist.json")
    out = []
    for cs in g["cases"]:
        if cs.get("msa"):
            continue
        glob = not (parse_args(cs["args"])["flag"] & 2)
        if skip_buggy_global and glob and "# Excluded" in cs["stderr"]:
            continue                      # reference picks wrong pairs here
        out.append(cs)
    return g["pool"], out


# ----------------------------------------------------------------------------
# KMA count matrices (.mat): text <-> arrays, and the reference's gates
# ----------------------------------------------------------------------------
def mat_text(template, ref, counts_file_order):
    """'#template' block: rows ref\tA\tC\tG\tT\tN\t- , blank line at the end."""
    rows = [f"#{template}"]
    for b, c in zip(ref, counts_file_order):
        rows.append(b + "\t" + "\t".join(str(int(x)) for x in c))
    return "\n".join(rows) + "\n\n"


def parse_mat(text, template):
    """-> (counts (L, 6) u16 in storage order A,C,G,T,-,N; totals (L,) u32) of the template's
    non-insertion rows, or None if the template is absent."""
    lines = text.split("\n")
    try:
        k = lines.index("#" + template) + 1
    except ValueError:
        return None
    counts, totals = [], []
    while k < len(lines) and lines[k] and not lines[k].startswith("#"):
        f = lines[k].split("\t")
        v = [int(x) for x in f[1:7]]
        if f[0] != "-":
            counts.append([v[0], v[1], v[2], v[3], v[5], v[4]])
            totals.append(sum(v))
        k += 1
    return np.array(counts, dtype=np.uint16).reshape(-1, 6), np.array(totals, dtype=np.uint32)


def mat_sample_gate(totals, min_depth, min_length, min_cov):
    """ltdmatrixthrd.c:455-458 / :523-526: is the sample included?"""
    n_nucs = int((totals >= min_depth).sum())
    return not (n_nucs < min_length or n_nucs < min_cov * len(totals))


def format_phy_cells(names, cells, precision=9, comment=None, flags=1):
    """printphy (phy.c:59-123) for double cells."""
    out = []
    if flags & 4:
        out.append(f"#{comment}")
    out.append("%10d" % len(names))
    k = 0
    for r, nm in enumerate(names):
        row = [nm if flags & 1 else "%-10.10s" % nm]
        for _ in range(r):
            d = float(cells[k])
            k += 1
            row.append("%d" % int(d) if d == int(d) else "%.*f" % (precision, d))
        out.append("\t".join(row))
    return "\n".join(out) + "\n"
