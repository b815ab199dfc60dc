"""Synthetic KMA-consensus alignments.

A uniform ACGT reference genome; each sample is the reference with i.i.d.
substitutions (rate 1e-3), `N` runs (about 1 % of positions, blocks of 64),
0.2 % lowercase ("insignificant") calls and 0.05 % `-`.  ``make_codes`` (numpy)
feeds the CPU-side tests; ``synth_torch.make_packed_torch`` (repo root, harness
only) builds the reference's packed in-memory format directly on a torch device
for the full-size bench workloads, where host generation would take minutes.
"""
import numpy as np

SNP_RATE = 1e-3
NRUN_BLOCK = 64
NRUN_RATE = 0.008          # fraction of 64-base blocks that are all-N
LOWER_RATE = 2e-3
GAP_RATE = 5e-4

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_LOWER = np.frombuffer(b"acgt", dtype=np.uint8)


def make_ascii(n, length, seed=1, snp=SNP_RATE, nrun=NRUN_RATE, lower=LOWER_RATE, gap=GAP_RATE):
    """(n, length) uint8 array of FASTA sequence bytes (no newlines)."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, size=length, dtype=np.uint8)
    out = np.empty((n, length), dtype=np.uint8)
    nblk = (length + NRUN_BLOCK - 1) // NRUN_BLOCK
    for i in range(n):
        code = ref.copy()
        sub = rng.random(length) < snp
        k = int(sub.sum())
        code[sub] = (code[sub] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
        row = _BASES[code]
        low = rng.random(length) < lower
        row = np.where(low, _LOWER[code], row)
        row = np.where(rng.random(length) < gap, np.uint8(ord("-")), row)
        blk = rng.random(nblk) < nrun
        row = np.where(np.repeat(blk, NRUN_BLOCK)[:length], np.uint8(ord("N")), row)
        out[i] = row
    return out


_TABLE = np.full(256, 32, dtype=np.uint8)
for _c, _v in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3), (b"U", 3)):
    _TABLE[_c[0]] = _v
for _c in b"N-RYSWKMBDHVXryswkmbdhvxacgtun":
    _TABLE[_c] = 4


def make_codes(n, length, seed=1, **kw):
    """(n, length) uint8 translated codes 0..4 (default flag: lowercase = unknown)."""
    return _TABLE[make_ascii(n, length, seed, **kw)]


def write_fasta(path, row, header="ref", width=60):
    with open(path, "wb") as f:
        f.write(b">" + header.encode() + b"\n")
        for s in range(0, len(row), width):
            f.write(row[s:s + width].tobytes() + b"\n")
