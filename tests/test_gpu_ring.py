"""Sample-shard ring (ccphylo_b200/ring.py) on ONE GPU: all ranks emulated in one process with the
loopback transport.  The schedule, the slot placement, the tile window and the block extraction are
the code the NCCL ring runs (scripts/ring_demo.py drives that on 2-8 GPUs); every cell of the
assembled matrix must equal the oracle's all-vs-all result bit for bit."""
import numpy as np
import pytest
import torch

import oracle
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import ring  # noqa: E402  (scripts/ring.py: the NCCL sample-shard ring harness)
from ccphylo_b200 import api, synth  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("kernel", [api.KERNEL_POPC, api.KERNEL_UMMA], ids=["popc", "umma"])
def test_ring_blocks_assemble_to_the_full_matrix(built, world, kernel):
    S, length = 512 if world % 2 == 0 else 256, 128 * 9 + 50
    n = world * S
    codes = synth.make_codes(n, length, seed=world * 7 + kernel, snp=0.02, nrun=0.05)
    seqs, masks, _ = oracle.encode_samples(codes)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    blocks = {}
    with torch.cuda.stream(stream):
        states = []
        for g in range(world):
            s = torch.from_numpy(seqs[g * S:(g + 1) * S].view(np.int64)).to(dev)
            m = torch.from_numpy(masks[g * S:(g + 1) * S].view(np.int32)).to(dev)
            st = ring.RankState(g, world, S, length, s, m, dev, kernel=kernel)
            st.ctx.set_stream(stream.cuda_stream)
            states.append(st)

        def on_block(rank, hi, lo, row0, D, N, dn):
            torch.cuda.synchronize()
            key = (hi, lo, row0, dn)
            if D.numel() == 0:
                return
            assert key not in blocks, "a block was computed twice"
            blocks[key] = (D.cpu().numpy().copy(), N.cpu().numpy().copy())

        ring.run_loopback(states, on_block, norm=1000000, min_length=1, min_cov=0.5)
        for st in states:
            st.ctx.close()
    D, N = ring.assemble(blocks, world, S)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, np.ones(n, np.uint8), length, norm=1000000)
    assert dno == n
    k = 0
    for r in range(1, n):
        assert np.array_equal(N[r, :r], No[k:k + r]), f"inclusion counts differ in row {r}"
        assert np.array_equal(D[r, :r].view(np.uint64), Do[k:k + r].view(np.uint64)), f"distances differ in row {r}"
        k += r
    # every unordered pair of shards met exactly once (split blocks: two halves)
    met = {}
    for (hi, lo, row0, dn), (d, _) in blocks.items():
        met[(hi, lo)] = met.get((hi, lo), 0) + (S if hi == lo else d.shape[0])
    assert met == {(hi, lo): S for hi in range(world) for lo in range(hi + 1)}
