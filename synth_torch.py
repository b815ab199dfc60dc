"""Bench / test harness only: the synthetic workload of ccphylo_b200/synth.py generated directly on a torch device in
the reference's packed in-memory format (qseqs.c:60 / fsacmp.c:164 layout), for the full-size bench workloads where
host generation would take minutes.  Not part of the product package (which imports neither torch nor the oracle)."""
from ccphylo_b200.synth import GAP_RATE, LOWER_RATE, NRUN_BLOCK, NRUN_RATE, SNP_RATE


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
