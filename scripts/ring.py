"""HARNESS, not product: the NCCL sample-shard ring of round 1, kept to compare against the library's own answer to
sets larger than one GPU (sequence shards + row windows, ccphylo_b200/csrc/ccg_group.cu; bench.py --workload ring:
2.17 s per 100,000 x 2.9 Mbp step on 8 B200 against 4.99 s for this ring).  It drives the C-ABI block by block
(ccg_set_tile_window) and moves the shards with torch.distributed.

Sample-shard ring for sets whose bit planes do not fit one GPU's HBM (BASELINE configs[3]:
100,000 samples x 2.9 Mbp = 109 GB of planes).

Every rank owns one shard of S consecutive samples (S a multiple of the macro-tile edge).  The
all-vs-all lower triangle is the set of shard blocks (hi, lo), lo <= hi.  Step 0: every rank
computes its diagonal block.  Step s = 1..: every rank holds the shard of rank (g - s) mod N as
its visitor -- it was passed one hop round the ring (NCCL send/recv over NVLink, issued BEFORE
the block of the current step is computed so the transfer hides behind it) -- and computes the
block of the two shards.  After floor(N/2) steps every unordered pair of shards has met; for even
N the last step's pairs meet at both ends, so the two owners split the block's rows.

On the device a block is one run of the ordinary pair kernel: the lower-index shard in sample
slots [0, S), the higher-index one in [S, 2S), restricted to rows [S, 2S) x columns [0, S) with
ccg_set_tile_window -- no new kernel, bit-identical cells.

The transport is pluggable: `NcclTransport` (one process per GPU, torch.distributed) or
`LoopbackTransport` (all ranks emulated in one process on one GPU: what the single-GPU parity
test drives; the schedule, the slot placement and the block extraction are the same code).
"""
import numpy as np

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccphylo_b200 import api  # noqa: E402


def shard_size(n, world, edge=256):
    """Samples per shard: ceil(n / world) rounded up to the macro-tile edge."""
    s = (n + world - 1) // world
    return ((s + edge - 1) // edge) * edge


def schedule(world):
    """[(step, shift, split)]: at `step` rank g meets shard (g - shift) % world; split = the block is
    shared with the partner rank (even world, last step) and each computes half of its rows."""
    out = [(0, 0, False)]
    for s in range(1, world // 2 + 1):
        if 2 * s == world:
            out.append((s, s, True))
        elif 2 * s < world:
            out.append((s, s, False))
    return out


def block_rows(g, h, shard, split, edge=256):
    """Rows of block (max(g,h), min(g,h)) that rank g computes: (row0, nrows) inside the block."""
    if not split:
        return 0, shard
    half = ((shard // 2 + edge - 1) // edge) * edge
    half = min(half, shard)
    return (0, half) if g < h else (half, shard - half)


class LoopbackTransport:
    """All ranks in one process: `rotate` hands every emulated rank the visitor of its left neighbour."""

    def __init__(self, world):
        self.world = world

    def rotate(self, visitors):
        return [visitors[(g - 1) % self.world] for g in range(self.world)]


class NcclTransport:
    """One process per GPU: send the current visitor to rank+1, receive the next one from rank-1."""

    def __init__(self, rank, world):
        import torch.distributed as dist
        self.dist, self.rank, self.world = dist, rank, world

    def start(self, send_bufs, recv_bufs):
        d = self.dist
        ops = []
        for sb, rb in zip(send_bufs, recv_bufs):
            ops.append(d.P2POp(d.isend, sb, (self.rank + 1) % self.world))
            ops.append(d.P2POp(d.irecv, rb, (self.rank - 1) % self.world))
        return d.batch_isend_irecv(ops)

    @staticmethod
    def wait(works):
        for w in works:
            w.wait()


class RankState:
    """What one rank keeps: its context, its shard (packed rows on the device) and result buffers."""

    def __init__(self, rank, world, shard, length, seqs, masks, device, kernel=api.KERNEL_AUTO, ctx=None):
        import torch
        self.rank, self.world, self.shard, self.length = rank, world, shard, length
        self.seqs, self.masks = seqs, masks                       # (shard, W) int64 / int32 device tensors
        self.ctx = ctx or api.Context(device.index if device.index is not None else 0)
        self.ctx.set_kernel(kernel)
        self.ctx.set_problem(2 * shard, length, pair=True)
        ncell = api.cells(2 * shard)
        self.d_D = torch.zeros(ncell, dtype=torch.float64, device=device)
        self.d_N = torch.zeros(ncell, dtype=torch.float64, device=device)

    def compute_block(self, other_rank, other_seqs, other_masks, split, include_lo=None, include_hi=None, **ep):
        """Runs block (hi, lo) of this rank and the visiting shard; returns (hi, lo, row0, D, N) with D, N
        (nrows, shard) float64 device tensors of the rows this rank is responsible for."""
        import torch
        g, h, S = self.rank, other_rank, self.shard
        ctx = self.ctx
        ctx.set_problem(2 * S, self.length, pair=True)
        if h == g:
            ctx.set_tile_window(0, S, 0, S)
            ctx.put_samples_packed_dev(self.seqs.data_ptr(), self.masks.data_ptr(), S, self.seqs.stride(0), first=0)
            inc = np.zeros(2 * S, np.uint8)
            inc[:S] = 1 if include_lo is None else include_lo
            dn = ctx.run_pair_dev(self.d_D.data_ptr(), self.d_N.data_ptr(), include=inc, **ep)
            k = api.cells(dn)
            return g, g, 0, self.d_D[:k], self.d_N[:k], dn      # packed triangle of the diagonal block
        lo_is_mine = g < h
        lo_s, lo_m = (self.seqs, self.masks) if lo_is_mine else (other_seqs, other_masks)
        hi_s, hi_m = (other_seqs, other_masks) if lo_is_mine else (self.seqs, self.masks)
        row0, nrows = block_rows(g, h, S, split)
        if nrows == 0:                                           # the partner's half covers the whole (one macro tile) block
            empty = self.d_D[:0].view(0, S)
            return max(g, h), min(g, h), row0, empty, empty, None
        ctx.set_tile_window(S + row0, S + row0 + nrows, 0, S)
        ctx.put_samples_packed_dev(lo_s.data_ptr(), lo_m.data_ptr(), S, lo_s.stride(0), first=0)
        ctx.put_samples_packed_dev(hi_s.data_ptr(), hi_m.data_ptr(), S, hi_s.stride(0), first=S)
        if include_lo is not None or include_hi is not None:
            raise NotImplementedError("per-sample exclusion inside a ring block: exclude before sharding")
        ctx.run_pair_dev(self.d_D.data_ptr(), self.d_N.data_ptr(), **ep)
        # rows S+row0 .. of the packed triangle over 2S samples, first S columns of each
        r = torch.arange(S + row0, S + row0 + nrows, device=self.d_D.device, dtype=torch.int64)
        idx = (r * (r - 1) // 2).view(-1, 1) + torch.arange(S, device=self.d_D.device, dtype=torch.int64).view(1, -1)
        return max(g, h), min(g, h), row0, self.d_D[idx], self.d_N[idx], None


def run_loopback(states, on_block, **ep):
    """Drive all emulated ranks of `states` through the ring on one device."""
    world = len(states)
    tr = LoopbackTransport(world)
    visitors = [(st.rank, st.seqs, st.masks) for st in states]
    prev_shift = 0
    for step, shift, split in schedule(world):
        for _ in range(shift - prev_shift):
            visitors = tr.rotate(visitors)
        prev_shift = shift
        for st, (vr, vs, vm) in zip(states, visitors):
            assert vr == (st.rank - shift) % world
            on_block(st.rank, *st.compute_block(vr, vs, vm, split, **ep))


def run_nccl(state, on_block, **ep):
    """One rank of the NCCL ring (call on every rank).  The visitor for step s+1 is in flight while
    the block of step s is computed."""
    import torch
    g, world = state.rank, state.world
    tr = NcclTransport(g, world)
    steps = schedule(world)
    bufs = [(state.seqs, state.masks), (torch.empty_like(state.seqs), torch.empty_like(state.masks)),
            (torch.empty_like(state.seqs), torch.empty_like(state.masks))]
    cur = 0                                                      # bufs[cur] = visitor of the current step
    for k, (step, shift, split) in enumerate(steps):
        works, nxt = None, None
        if k + 1 < len(steps):
            # pass the current visitor on and receive the next one while this block is computed
            nxt = 1 if cur != 1 else 2
            works = tr.start(bufs[cur], bufs[nxt])
        vr = (g - shift) % world
        on_block(g, *state.compute_block(vr, bufs[cur][0], bufs[cur][1], split, **ep))
        if works is not None:
            tr.wait(works)
            cur = nxt


def assemble(blocks, n_shards, shard):
    """Host-side check helper: {(hi, lo, row0): (D, N)} -> dense (n, n) lower-triangular arrays."""
    n = n_shards * shard
    D = np.zeros((n, n))
    N = np.zeros((n, n))
    for (hi, lo, row0, dn), (d, m) in blocks.items():
        if hi == lo:
            k = 0
            for r in range(1, dn):
                D[hi * shard + r, lo * shard:lo * shard + r] = d[k:k + r]
                N[hi * shard + r, lo * shard:lo * shard + r] = m[k:k + r]
                k += r
        else:
            D[hi * shard + row0:hi * shard + row0 + d.shape[0], lo * shard:(lo + 1) * shard] = d
            N[hi * shard + row0:hi * shard + row0 + d.shape[0], lo * shard:(lo + 1) * shard] = m
    return D, N
