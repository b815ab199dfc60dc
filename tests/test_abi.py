"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares;
device entry points fail loudly (no CPU fallback) when no B200 is present."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        syms |= set(re.findall(r"\b(ccg_[a-z0-9_]+)\s*\(", text))
    return syms


def test_library_exports_every_declared_symbol(built):
    from ccphylo_b200 import api

    L = api.load()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in sorted(syms):
        assert hasattr(L, s), f"{s} declared in include/ but not exported"
    assert syms == set(api.EXPORTS), syms ^ set(api.EXPORTS)


def test_every_entry_point_cites_the_reference(built):
    text = open(os.path.join(ROOT, "include", "ccphylo_gpu.h")).read()
    for needle in ("fsacmpthrd.h:49", "fsacmpthrd.c:261", "fsacmpthrd.c:108", "qseqs.c:60", "fsacmp.c:164",
                   "cdist.c:181", "matrix.c:32", "matcmp.c:448", "ltdmatrixthrd.c:376", "dist.c:738"):
        assert needle in text


def test_no_cpu_fallback_without_device(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ccphylo_b200 import api

    with pytest.raises(api.CcgError) as e:
        api.Context()
    assert e.value.code == 1          # CCG_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "ccphylo_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
            text = open(path, errors="ignore").read()
            assert "import oracle" not in text and "liboracle" not in text and "fsa_oracle" not in text \
                and "mat_oracle" not in text, path


def test_partition_helpers_are_host_only(built):
    from ccphylo_b200 import api

    L = api.load()
    BM, BN = L.ccg_tile_rows(), L.ccg_tile_cols()
    assert (BM, BN) == (256, 256)
    for n in (1, 2, 63, 64, 65, 129, 1000, 2816):
        rows = (n + BM - 1) // BM
        expect = [(tm, tn) for tm in range(rows) for tn in range(tm + 1)] if n >= 2 else []
        for world in (1, 2, 3, 8):
            cells = [api.partition_cells(n, r, world) for r in range(world)]
            assert sum(cells) == api.cells(n)
            tiles = [t for r in range(world) for t in api.partition_tiles(n, r, world)]
            assert sorted(tiles) == expect
