"""GPU parity of the multi-GPU K split (ccg_group.cu): every member owns a slice of the alignment, runs the whole
lower triangle on it, and the owner of a matrix row adds the members' int32 partial sums through peer pointers
inside its epilogue kernel.  Integer split sums are exact, so every result must equal the single-device result
and the oracle BIT FOR BIT -- pair mode, shared-mask mode, every cell type, exclusions, the one-call drop-in and
the staged calls the host driver uses.

On a single-GPU box the members are contexts on the same device (ccg_init_multi_devices with a repeated id):
the same barrier, the same peer-pointer epilogue, the same host fan-out.  With two or more GPUs the same tests
also run across real devices, and one process per GPU (CUDA IPC handles, torchrun) is covered through bench.py."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle
from ccphylo_b200 import api, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def _set(n, length, seed, **kw):
    kw.setdefault("snp", 0.02)
    kw.setdefault("nrun", 0.05)
    codes = synth.make_codes(n, length, seed=seed, **kw)
    seqs, masks, inc = oracle.encode_samples(codes)
    return codes, seqs, masks, inc


def _devices(members):
    """member -> device: real devices first, then wrapped round (one GPU: all on device 0)."""
    return [g % max(NGPU, 1) for g in range(members)]


@pytest.fixture(scope="module", params=[2, 3, 4], ids=lambda m: f"{m}members")
def multi(built, request):
    os.environ["CCG_MULTI_FORCE"] = "1"          # split whatever the size: the tests are small
    try:
        c = api.Context(multi=_devices(request.param))
    finally:
        del os.environ["CCG_MULTI_FORCE"]
    c.members = request.param
    yield c
    c.close()


@pytest.mark.parametrize("n,length", [(2, 600), (70, 1024), (257, 4099), (300, 20000 + 17), (513, 3000)])
def test_group_pair_counts_bit_exact(multi, n, length):
    codes, seqs, masks, inc = _set(n, length, seed=n * 7 + length)
    include = np.ones(n, np.uint8)
    D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, min_length=0, min_cov=0.0, ctx=multi)
    total, active = multi.multi_gpus()
    assert total == multi.members and active == min(multi.members, length // 256)
    if active > 1:
        assert "K split" in multi.last_kernel and "k_finalize_group" in multi.last_kernel
    mo, no = oracle.raw_pair_matrix(seqs, masks, length)
    assert dn == n
    assert np.array_equal(N, no.astype(np.float64))
    assert np.array_equal(D, mo.astype(np.float64))


@pytest.mark.parametrize("elem,scale", [(8, 1.0), (4, 1.0), (2, 10.0), (1, 0.01)])
@pytest.mark.parametrize("norm", [0, 1000000])
def test_group_pair_epilogue_with_exclusions(multi, elem, scale, norm):
    n, length = 300, 6001
    codes, seqs, masks, inc = _set(n, length, seed=elem + norm % 97)
    codes[7, :] = 4
    codes[130, : length - 100] = 4
    codes[299, :] = 4
    seqs, masks, inc = oracle.encode_samples(codes)
    min_len = int(0.5 * length)
    include = (inc >= min_len).astype(np.uint8)
    D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=norm, min_length=min_len,
                                         min_cov=0.5, elem_size=elem, byte_scale=scale, ctx=multi)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=norm, min_length=min_len, min_cov=0.5,
                                      elem_size=elem, byte_scale=scale)
    assert dn == dno == n - 3
    assert np.array_equal(N.view(np.uint8), No.view(np.uint8))
    assert np.array_equal(D.view(np.uint8), Do.view(np.uint8))


def test_group_min_cov_gate_uses_the_whole_alignment(multi):
    # a pair whose overlap is below minCov * L of the WHOLE alignment must read -1 (fsacmpthrd.c:292, :430) although
    # it passes minCov * (slice length) on every member
    n, length = 200, 8192
    codes, seqs, masks, inc = _set(n, length, seed=5)
    codes[3, : int(0.45 * length)] = 4
    codes[9, int(0.55 * length):] = 4
    seqs, masks, inc = oracle.encode_samples(codes)
    include = np.ones(n, np.uint8)
    D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=1000, min_length=1, min_cov=0.3, ctx=multi)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=1000, min_length=1, min_cov=0.3)
    assert (Do == -1.0).sum() > 0
    assert np.array_equal(D.view(np.uint64), Do.view(np.uint64)) and np.array_equal(N, No)


@pytest.mark.parametrize("norm", [0, 1000])
def test_group_shared_mask_mode(multi, norm):
    n, length = 260, 128 * 40 + 77
    codes, seqs, masks, inc = _set(n, length, seed=21, nrun=0.01)
    include = np.ones(n, np.uint8)
    gmask = oracle.global_mask(codes, include)
    D, _, dn, ginc = api.fsa_cmp_thread_out(seqs, include, gmask.reshape(1, -1), length, pair=False, norm=norm, ctx=multi)
    Do, dno, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=norm)
    assert dn == dno and ginc == ginco
    assert np.array_equal(D.view(np.uint64), Do.view(np.uint64))
    assert D.max() > 0


def test_group_staged_calls_of_the_host_driver(multi):
    # what ccphylo-b200 dist does: codes per sample -> device encode, per-sample counts back, gate, run
    n, length = 210, 5000 + 13
    codes, seqs, masks, inc = _set(n, length, seed=77)
    codes[11, :] = 4
    seqs, masks, inc = oracle.encode_samples(codes)
    multi.set_problem(n, length, pair=True)
    for k in range(n):
        multi.put_sample_codes(k, codes[k])
    got_inc = multi.inc_counts()
    assert np.array_equal(got_inc, inc)
    min_len = int(0.5 * length)
    include = (got_inc >= min_len).astype(np.uint8)
    D, N, dn = multi.run_pair(include=include, norm=1000000, min_length=1, min_cov=0.5)
    Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=1000000, min_length=1, min_cov=0.5)
    assert dn == dno == n - 1
    assert np.array_equal(N, No) and np.array_equal(D.view(np.uint64), Do.view(np.uint64))
    # shared-mask mode on the same store: the mask is built on the members, its count summed (cdist.c:101-112)
    ginc = multi.build_global_mask(include)
    Dg, dng, ginc2 = multi.run_global(include=include, norm=1000)
    gmask = oracle.global_mask(codes, include)
    Dgo, dngo, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=1000)
    assert ginc == ginc2 == ginco and dng == dngo
    assert np.array_equal(Dg.view(np.uint64), Dgo.view(np.uint64))


def test_group_repeated_runs_alternate_the_accumulator_buffers(multi):
    # the two accumulator buffers are used in turn and the barrier epochs advance: five different problems in a row
    for k in range(5):
        n, length = 200 + 16 * k, 3000 + 300 * k
        codes, seqs, masks, inc = _set(n, length, seed=100 + k)
        D, N, dn, _ = api.fsa_cmp_thread_out(seqs, np.ones(n, np.uint8), masks, length, pair=True, min_length=0, min_cov=0.0, ctx=multi)
        mo, no = oracle.raw_pair_matrix(seqs, masks, length)
        assert np.array_equal(N, no.astype(np.float64)) and np.array_equal(D, mo.astype(np.float64))


@pytest.mark.parametrize("members,n,length,tile_rows", [(3, 700, 3000, 1), (2, 1100, 2600, 2), (4, 513, 5000 + 3, 1)])
@pytest.mark.parametrize("pair", [True, False], ids=["pair", "shared-mask"])
def test_group_windows_of_macro_tile_rows(built, monkeypatch, members, n, length, tile_rows, pair):
    # BASELINE configs[3] in small: the n x n accumulators do not fit the peer window, so the triangle is run in
    # windows of `tile_rows` macro-tile rows, each reduced over the members and finalised before its buffer is reused
    n_pad = (n + 255) // 256 * 256
    monkeypatch.setenv("CCG_MULTI_FORCE", "1")
    monkeypatch.setenv("CCG_GROUP_WINDOW_BYTES", str(tile_rows * 2 * 2 * 256 * n_pad * 4))
    codes, seqs, masks, inc = _set(n, length, seed=n + length, nrun=0.02)
    codes[5, :] = 4
    seqs, masks, inc = oracle.encode_samples(codes)
    include = (inc >= length // 2).astype(np.uint8)
    c = api.Context(multi=_devices(members))
    try:
        if pair:
            D, N, dn, _ = api.fsa_cmp_thread_out(seqs, include, masks, length, pair=True, norm=1000000, min_length=1, min_cov=0.5, ctx=c)
            Do, No, dno = oracle.fsa_cmp_pair(seqs, masks, include, length, norm=1000000, min_length=1, min_cov=0.5)
            assert np.array_equal(N, No)
        else:
            gmask = oracle.global_mask(codes, include)
            D, _, dn, ginc = api.fsa_cmp_thread_out(seqs, include, gmask.reshape(1, -1), length, pair=False, norm=1000, ctx=c)
            Do, dno, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=1000)
            assert ginc == ginco
        kern = c.last_kernel
        assert f"windows={n_pad // (256 * tile_rows) + (1 if n_pad % (256 * tile_rows) else 0)}" in kern, kern
        assert dn == dno == n - 1
        assert np.array_equal(D.view(np.uint64), Do.view(np.uint64))
    finally:
        c.close()


def test_small_or_special_problems_stay_on_one_member(built):
    c = api.Context(multi=_devices(2))
    try:
        n, length = 96, 2000                      # below every threshold of multi_choose_active
        codes, seqs, masks, inc = _set(n, length, seed=3)
        D, N, dn, _ = api.fsa_cmp_thread_out(seqs, np.ones(n, np.uint8), masks, length, pair=True, min_length=0, min_cov=0.0, ctx=c)
        assert c.multi_gpus() == (2, 1)
        assert c.multi_contexts() == 1            # the second device's context was never started
        mo, no = oracle.raw_pair_matrix(seqs, masks, length)
        assert np.array_equal(N, no.astype(np.float64)) and np.array_equal(D, mo.astype(np.float64))
        # -P: sequential along the alignment -> member 0 alone, same result as a single-device context
        D1, N1, dn1, _ = api.fsa_cmp_thread_out(seqs, np.ones(n, np.uint8), masks, length, pair=True, min_length=0, min_cov=0.0,
                                                proxi=5, ctx=c)
        with api.Context() as s:
            D2, N2, dn2, _ = api.fsa_cmp_thread_out(seqs, np.ones(n, np.uint8), masks, length, pair=True, min_length=0,
                                                    min_cov=0.0, proxi=5, ctx=s)
        assert np.array_equal(D1, D2) and np.array_equal(N1, N2)
        assert c.multi_contexts() == 1
        # a problem worth splitting starts the other member, and the handle goes back to one member afterwards
        os.environ["CCG_MULTI_FORCE"] = "1"
        try:
            c2 = api.Context(multi=_devices(3))
        finally:
            del os.environ["CCG_MULTI_FORCE"]
        try:
            assert c2.multi_contexts() == 1
            n, length = 40, 3000
            codes, seqs, masks, inc = _set(n, length, seed=4)
            D, N, dn, _ = api.fsa_cmp_thread_out(seqs, np.ones(n, np.uint8), masks, length, pair=True, min_length=0, min_cov=0.0, ctx=c2)
            assert c2.multi_gpus() == (3, 3) and c2.multi_contexts() == 3
            mo, no = oracle.raw_pair_matrix(seqs, masks, length)
            assert np.array_equal(N, no.astype(np.float64)) and np.array_equal(D, mo.astype(np.float64))
        finally:
            c2.close()
    finally:
        c.close()


def test_split_problem_refuses_single_device_calls(multi):
    n, length = 200, 4096
    codes, seqs, masks, inc = _set(n, length, seed=9)
    api.fsa_cmp_thread_out(seqs, np.ones(n, np.uint8), masks, length, pair=True, ctx=multi)
    if multi.multi_gpus()[1] > 1:
        with pytest.raises(api.CcgError) as e:
            multi.raw_counts(n)
        assert e.value.code == 5
        with pytest.raises(api.CcgError):
            multi.run_row(5)


def torchrun(n, script, *args, port=29640):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script)] + list(args)
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)


@pytest.mark.skipif(NGPU < 2, reason="needs at least 2 GPUs")
def test_bench_k_split_one_process_per_gpu(built):
    # one process per GPU, CUDA IPC handles exchanged through torch.distributed: the driver's launch line
    n = 4 if NGPU >= 4 else 2
    p = torchrun(n, "bench.py", "--gpus", str(n), "--steps", "2", "--warmup", "3", "--samples", "1024", "--length", "600000",
                 "--no-cpu-baseline", port=29641)
    assert p.returncode == 0, p.stderr[-3000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == n and line["parity_vs_oracle"] is True and line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["parity_cells_checked"] >= 1000
