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
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import ring

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
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import ring

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


# ---- K-split group: host arithmetic of the row ownership and of the alignment slices, on gloo ranks ----
def _group_worker(rank, world, port, n, length, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ccphylo_b200 import api

    L = api.load()
    rows = torch.zeros(n, dtype=torch.int64)
    cells = 0
    for lo, hi in api.group_owned_blocks(n, rank, world):
        rows[lo:hi] += 1
        cells += hi * (hi - 1) // 2 - lo * (lo - 1) // 2
        assert all(L.ccg_group_row_owner(i, world) == rank for i in (lo, hi - 1))
    assert cells == api.group_cells(n, rank, world)
    cells = torch.tensor([cells], dtype=torch.int64)
    sl = api.group_slices(length, world)
    bases = torch.zeros(length, dtype=torch.int64)
    bases[sl[rank]:sl[rank + 1]] += 1
    worst = torch.tensor([float(cells.item())])
    dist.all_reduce(rows)
    dist.all_reduce(cells)
    dist.all_reduce(bases)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((rows.tolist(), int(cells.item()), bases.min().item(), bases.max().item(), float(worst.item()), sl))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,length", [(2, 1000, 5000), (2, 10000, 5_000_000), (3, 257, 77777)])
def test_group_rows_and_slices_partition_the_job(built, world, n, length):
    from ccphylo_b200 import api

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_group_worker, args=(r, world, port, n, length, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows, cells, bmin, bmax, worst, sl = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(v == 1 for v in rows)                       # every matrix row has exactly one owner
    assert cells == api.cells(n)                           # the owners' cells tile the triangle
    assert bmin == 1 and bmax == 1                         # every base of the alignment is in exactly one slice
    assert all(b % 256 == 0 for b in sl[:-1]) and sl[-1] == length
    assert worst <= api.cells(n) / world + 64 * n          # round-robin blocks of 64 rows: within a block row of equal


def test_group_ownership_edge_cases(built):
    from ccphylo_b200 import api

    assert api.group_row_block() == 64
    for n in (0, 1, 2, 63, 64, 65, 255, 10000, 100000):
        for world in (1, 2, 5, 8, 16):
            blocks = sorted(b for r in range(world) for b in api.group_owned_blocks(n, r, world))
            assert sum(hi - lo for lo, hi in blocks) == n
            assert all(x[1] == y[0] for x, y in zip(blocks, blocks[1:]))
            assert sum(api.group_cells(n, r, world) for r in range(world)) == api.cells(n)
    assert api.group_slices(5_000_000, 8)[1] == 624896 and api.group_slices(1000, 3) == [0, 256, 512, 1000]
