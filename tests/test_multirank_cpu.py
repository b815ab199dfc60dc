"""N>1 host-side logic on CPU: world_size-2 gloo ranks deal the lower-triangular tile blocks
between them with no overlap and no gap (the path has no data-path collective; the only
collective here is the test's own checksum)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ccphylo_b200 import api

    BM = api.load().ccg_tile_rows()
    rows = (n + BM - 1) // BM
    ids = {}
    for tm in range(rows):
        for tn in range(tm + 1):
            ids[(tm, tn)] = len(ids)
    owner = torch.zeros(len(ids), dtype=torch.int64)
    for t in api.partition_tiles(n, rank, world):
        owner[ids[t]] += 1
    cells = torch.tensor([api.partition_cells(n, rank, world)], dtype=torch.int64)
    dist.all_reduce(owner)
    dist.all_reduce(cells)
    if rank == 0:
        out.put((owner.tolist(), int(cells.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 1408])
def test_two_ranks_cover_the_triangle_exactly_once(built, n):
    import bench
    from ccphylo_b200 import api

    assert bench.samples_for(1, 1000) == 1000 and bench.samples_for(2, 1000) == 1000
    assert bench.samples_for(2, 1000, "weak") == 1408
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    owner, cells = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(v == 1 for v in owner)
    assert cells == api.cells(n)


# ---- sample-shard ring: the schedule and the NCCL-style transport, on gloo with CPU tensors ----
def _ring_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ccphylo_b200 import ring

    tr = ring.NcclTransport(rank, world)            # the transport only needs torch.distributed P2P
    cur = torch.full((4,), float(rank))
    met = []
    steps = ring.schedule(world)
    for k, (step, shift, split) in enumerate(steps):
        nxt = torch.empty(4)
        works = tr.start([cur], [nxt]) if k + 1 < len(steps) else None
        met.append((int(cur[0].item()), shift, split, ring.block_rows(rank, int(cur[0].item()), 512, split)))
        if works is not None:
            tr.wait(works)
            cur = nxt
    out.put((rank, met))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_ring_schedule_visits_every_shard_pair_once(world):
    from ccphylo_b200 import ring

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = {}
    for g, met in got.items():
        for visitor, shift, split, (row0, nrows) in met:
            assert visitor == (g - shift) % world            # the ring delivered the right shard at this step
            key = (max(g, visitor), min(g, visitor))
            rows.setdefault(key, []).append((row0, nrows))
    assert set(rows) == {(hi, lo) for hi in range(world) for lo in range(hi + 1)}
    for key, parts in rows.items():
        parts.sort()
        assert parts[0][0] == 0 and sum(p[1] for p in parts) == 512, (key, parts)
        assert all(a[0] + a[1] == b[0] for a, b in zip(parts, parts[1:])), (key, parts)
    assert ring.shard_size(100000, 8) == 12544 and ring.shard_size(1000, 4) == 256
