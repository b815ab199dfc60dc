"""Synthetic KMA-consensus alignments.

A uniform ACGT reference genome; each sample is the reference with i.i.d.
substitutions (rate 1e-3), `N` runs (about 1 % of positions, blocks of 64),
0.2 % lowercase ("insignificant") calls and 0.05 % `-`.  ``make_codes`` (numpy)
feeds the CPU-side tests; ``make_packed_torch`` builds the reference's packed
in-memory format (qseqs.c:60 / fsacmp.c:164 layout) directly on a torch device
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


def make_packed_torch(n, length, seed, device, snp=SNP_RATE, nrun=NRUN_RATE, scatter=LOWER_RATE + GAP_RATE,
                      out_seqs=None, out_masks=None):
    """Reference packed format built on a torch device.

    Returns (seqs int64 (n, W), masks int32 (n, W)): the bit patterns of the
    reference's u64 / u32 words (two's complement views).
    """
    import torch

    W = (length >> 5) + (1 if length & 31 else 0)
    Lp = W * 32
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ref = torch.randint(0, 4, (Lp,), generator=g, device=device, dtype=torch.int64)
    seqs = out_seqs if out_seqs is not None else torch.empty((n, W), dtype=torch.int64, device=device)
    masks = out_masks if out_masks is not None else torch.empty((n, W), dtype=torch.int32, device=device)
    sh_code = (62 - 2 * torch.arange(32, device=device, dtype=torch.int64)).view(1, 32)
    sh_mask = (31 - torch.arange(32, device=device, dtype=torch.int64)).view(1, 32)
    valid = (torch.arange(Lp, device=device) < length)
    nblk = (Lp + NRUN_BLOCK - 1) // NRUN_BLOCK
    for i in range(n):
        r = torch.rand(Lp, generator=g, device=device)
        code = torch.where(r < snp, (ref + 1 + (r * 3e6).long() % 3) & 3, ref)
        known = (torch.rand(Lp, generator=g, device=device) >= scatter) & valid
        blk = torch.rand(nblk, generator=g, device=device) < nrun
        known &= ~blk.repeat_interleave(NRUN_BLOCK)[:Lp]
        k64 = known.long()
        seqs[i] = ((code * k64).view(W, 32) << sh_code).sum(dim=1)
        masks[i] = (k64.view(W, 32) << sh_mask).sum(dim=1).to(torch.int32)
    return seqs, masks
