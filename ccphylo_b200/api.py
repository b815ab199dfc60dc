"""ctypes binding of the C-ABI in include/ccphylo_gpu.h.

This is the host-side mirror used by the tests and the bench.  It adds no
arithmetic of its own: every result comes from the CUDA library, and loading
fails loudly when the library is missing (there is no CPU fallback).

Reference interface mirrored: ``fsaCmpThreadOut`` (fsacmpthrd.h:49) with its
two workers ``cmpairFsaThrd`` (pair mode, fsacmpthrd.c:261) and ``cmpFsaThrd``
(shared-mask mode, fsacmpthrd.c:108).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libccphylo_gpu.so")

ELEM_DTYPE = {8: np.float64, 4: np.float32, 2: np.uint16, 1: np.uint8}
KERNEL_AUTO, KERNEL_POPC, KERNEL_UMMA, KERNEL_FUSED = 0, 1, 2, 3

EXPORTS = [
    "ccg_strerror", "ccg_last_error", "ccg_init", "ccg_destroy", "ccg_set_stream", "ccg_set_kernel", "ccg_sync",
    "ccg_set_partition", "ccg_set_tile_window", "ccg_tile_rows", "ccg_tile_cols", "ccg_partition_cells", "ccg_partition_tiles",
    "ccg_set_scratch_limit", "ccg_set_problem", "ccg_put_global_mask", "ccg_apply_global_mask", "ccg_build_global_mask",
    "ccg_put_samples_packed",
    "ccg_put_samples_packed_dev", "ccg_put_samples_packed_dev_borrowed", "ccg_put_sample_codes", "ccg_get_inc_counts", "ccg_run_pair", "ccg_run_global",
    "ccg_run_pair_dev", "ccg_run_global_dev", "ccg_get_raw_counts", "ccg_fsa_cmp_thread_out", "ccg_host_alloc",
    "ccg_host_free", "ccg_launch_count", "ccg_last_kernel", "ccg_last_compare_ms", "ccg_last_phase_ms",
    "ccg_measure_i8_peak", "ccg_measure_fp4_peak", "ccg_mat_set_problem", "ccg_mat_put_sample", "ccg_mat_run",
    "ccg_set_proximity", "ccg_sample_proximity", "ccg_run_row", "ccg_mat_run_row", "ccg_list_variants", "ccg_set_motifs", "ccg_mask_motifs", "ccg_list_variants_row",
    "ccg_init_multi", "ccg_init_multi_devices", "ccg_multi_gpus", "ccg_multi_contexts", "ccg_group_export", "ccg_group_join", "ccg_group_leave",
    "ccg_mat_run_partial", "ccg_mat_finalize_host", "ccg_group_set_alignment", "ccg_group_set_output", "ccg_group_row_block", "ccg_group_row_owner", "ccg_group_cells",
    "ccg_sample_count_masked", "ccg_trim_begin", "ccg_trim_sample", "ccg_trim_keep_reference", "ccg_trim_get_mask", "ccg_trim_end",
]
GROUP_HANDLE_BYTES = 128

MAT_METHODS = ["cos", "z", "chi2", "nchi2", "c", "nc", "p", "np", "bc", "nbc", "l1", "l2", "linf", "ln", "nl1", "nl2",
               "nlinf", "nln"]          # CCG_MAT_* ids, include/ccphylo_gpu.h


def mat_method(name):
    """'-d' name -> (CCG_MAT_* id, order), same precedence as the dist driver (dist.c:738-786)."""
    if name in MAT_METHODS and name not in ("ln", "nln"):
        return MAT_METHODS.index(name), 0
    if name.startswith("l"):
        return MAT_METHODS.index("ln"), int(name[1:])
    if name.startswith("nl"):
        return MAT_METHODS.index("nln"), int(name[2:])
    raise ValueError(name)


VARIANT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_size_t)


class CcgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ccphylo_gpu error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libccphylo_gpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA library is the only implementation; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i, u, d, ll = C.c_void_p, C.c_int, C.c_uint, C.c_double, C.c_longlong
    L.ccg_strerror.restype = C.c_char_p
    L.ccg_strerror.argtypes = [i]
    L.ccg_last_error.restype = C.c_char_p
    L.ccg_last_error.argtypes = [vp]
    L.ccg_init.argtypes = [C.POINTER(vp), i]
    L.ccg_destroy.restype = None
    L.ccg_destroy.argtypes = [vp]
    L.ccg_set_stream.argtypes = [vp, vp]
    L.ccg_set_kernel.argtypes = [vp, i]
    L.ccg_sync.argtypes = [vp]
    L.ccg_set_partition.argtypes = [vp, i, i]
    L.ccg_set_tile_window.argtypes = [vp, i, i, i, i]
    L.ccg_tile_rows.restype = i
    L.ccg_tile_rows.argtypes = []
    L.ccg_tile_cols.restype = i
    L.ccg_tile_cols.argtypes = []
    L.ccg_set_scratch_limit.argtypes = [vp, C.c_size_t]
    L.ccg_partition_cells.restype = ll
    L.ccg_partition_cells.argtypes = [i, i, i]
    L.ccg_partition_tiles.restype = ll
    L.ccg_partition_tiles.argtypes = [i, i, i, vp, vp, ll]
    L.ccg_set_problem.argtypes = [vp, i, i, i]
    L.ccg_put_global_mask.argtypes = [vp, vp]
    L.ccg_apply_global_mask.argtypes = [vp, vp]
    L.ccg_build_global_mask.argtypes = [vp, vp, vp]
    L.ccg_put_samples_packed.argtypes = [vp, i, i, vp, vp]
    L.ccg_put_samples_packed_dev.argtypes = [vp, i, i, vp, vp, C.c_long]
    L.ccg_put_samples_packed_dev_borrowed.argtypes = [vp, i, i, vp, vp, C.c_long]
    L.ccg_put_sample_codes.argtypes = [vp, i, vp]
    L.ccg_get_inc_counts.argtypes = [vp, vp]
    L.ccg_set_motifs.argtypes = [vp, i, vp, vp]
    L.ccg_mask_motifs.argtypes = [vp, i, i, vp]
    L.ccg_set_proximity.argtypes = [vp, u, i]
    L.ccg_sample_proximity.argtypes = [vp, i, i, i, vp]
    L.ccg_run_pair.argtypes = [vp, vp, u, u, d, i, d, vp, vp, C.POINTER(i)]
    L.ccg_run_global.argtypes = [vp, vp, u, i, d, vp, C.POINTER(i), C.POINTER(u)]
    L.ccg_run_pair_dev.argtypes = [vp, vp, u, u, d, i, d, vp, vp, C.POINTER(i)]
    L.ccg_run_global_dev.argtypes = [vp, vp, u, i, d, vp, C.POINTER(i), C.POINTER(u)]
    L.ccg_get_raw_counts.argtypes = [vp, vp, vp]
    L.ccg_list_variants.argtypes = [vp, i, vp, VARIANT_FN, vp]
    L.ccg_list_variants_row.argtypes = [vp, i, VARIANT_FN, vp]
    L.ccg_run_row.argtypes = [vp, i, u, u, d, vp, vp, C.POINTER(i)]
    L.ccg_fsa_cmp_thread_out.argtypes = [vp, i, vp, vp, i, d, i, i, vp, vp, vp, u, u, d, u, C.POINTER(i), C.POINTER(u)]
    L.ccg_host_alloc.restype = vp
    L.ccg_host_alloc.argtypes = [C.c_size_t]
    L.ccg_host_free.restype = None
    L.ccg_host_free.argtypes = [vp]
    L.ccg_launch_count.restype = ll
    L.ccg_launch_count.argtypes = [vp]
    L.ccg_last_kernel.restype = C.c_char_p
    L.ccg_last_kernel.argtypes = [vp]
    L.ccg_last_compare_ms.restype = C.c_float
    L.ccg_last_compare_ms.argtypes = [vp]
    L.ccg_last_phase_ms.restype = C.c_float
    L.ccg_last_phase_ms.argtypes = [vp, i]
    L.ccg_mat_set_problem.argtypes = [vp, i, i]
    L.ccg_mat_put_sample.argtypes = [vp, i, vp, vp, i]
    L.ccg_mat_run.argtypes = [vp, vp, i, C.c_uint, C.c_double, C.c_uint, C.c_uint, C.c_uint, C.c_double, i, C.c_double,
                              vp, vp, vp, vp]
    L.ccg_mat_run_partial.argtypes = [vp, vp, i, C.c_uint, C.c_double, C.c_uint, vp, vp, vp]
    L.ccg_mat_finalize_host.argtypes = [i, vp, vp, vp, vp, C.c_uint, C.c_uint, C.c_double, i, C.c_double, vp, vp, vp, vp]
    L.ccg_mat_run_row.argtypes = [vp, i, i, C.c_uint, C.c_double, C.c_uint, C.c_uint, C.c_uint, C.c_double, vp, vp, vp]
    L.ccg_sample_count_masked.argtypes = [vp, i, vp]
    L.ccg_trim_begin.argtypes = [vp, i, C.c_uint]
    L.ccg_trim_sample.argtypes = [vp, vp, vp, i, i, vp]
    L.ccg_trim_keep_reference.argtypes = [vp]
    L.ccg_trim_get_mask.argtypes = [vp, i, vp, vp, vp]
    L.ccg_trim_end.argtypes = [vp]
    L.ccg_init_multi.argtypes = [C.POINTER(vp), i]
    L.ccg_init_multi_devices.argtypes = [C.POINTER(vp), i, vp]
    L.ccg_multi_gpus.argtypes = [vp, C.POINTER(i)]
    L.ccg_multi_contexts.argtypes = [vp]
    L.ccg_group_export.argtypes = [vp, i, vp]
    L.ccg_group_join.argtypes = [vp, i, i, vp]
    L.ccg_group_leave.argtypes = [vp]
    L.ccg_group_set_alignment.argtypes = [vp, ll, u]
    L.ccg_group_set_output.argtypes = [vp, i]
    L.ccg_group_row_block.argtypes = []
    L.ccg_group_row_owner.argtypes = [i, i]
    L.ccg_group_cells.restype = ll
    L.ccg_group_cells.argtypes = [i, i, i]
    L.ccg_measure_fp4_peak.restype = C.c_double
    L.ccg_measure_fp4_peak.argtypes = [vp, C.c_double, C.c_double, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    L.ccg_measure_i8_peak.restype = C.c_double
    L.ccg_measure_i8_peak.argtypes = [vp, C.c_double]
    _lib = L
    return L


def partition_cells(n, rank, world):
    """Cells of the n-sample lower triangle owned by `rank` of `world` (host only, no device needed)."""
    return load().ccg_partition_cells(n, rank, world)


def partition_tiles(n, rank, world):
    """(tm, tn) macro tiles owned by `rank` of `world` (host only, no device needed)."""
    L = load()
    k = L.ccg_partition_tiles(n, rank, world, None, None, 0)
    ti = np.zeros(max(k, 1), dtype=np.int32)
    tj = np.zeros(max(k, 1), dtype=np.int32)
    L.ccg_partition_tiles(n, rank, world, ti.ctypes.data, tj.ctypes.data, k)
    return list(zip(ti[:k].tolist(), tj[:k].tolist()))


def group_row_block():
    """Matrix rows are owned in blocks of this many rows, dealt round-robin over the members of a K-split group."""
    return load().ccg_group_row_block()


def group_owned_blocks(n, rank, world):
    """[(row_lo, row_hi)] of the row blocks of an n-sample matrix that member `rank` of `world` finalises (host only)."""
    blk = group_row_block()
    return [(b * blk, min(n, b * blk + blk)) for b in range(rank, (n + blk - 1) // blk, world)]


def group_cells(n, rank, world):
    """Packed cells of an n-sample triangle (all samples included) that member `rank` owns."""
    return load().ccg_group_cells(n, rank, world)


def group_slices(length, world):
    """First base of every member's slice of the alignment (multiples of 256) plus the end: what ccg_init_multi
    uses, and what one-process-per-GPU callers should use so that the slices partition the alignment."""
    return [length if g == world else (length * g // world) // 256 * 256 for g in range(world + 1)]


def words(length):
    return (length >> 5) + (1 if length & 31 else 0)


def cells(dn):
    return dn * (dn - 1) // 2 if dn > 1 else 0


def _row_ptrs(arr2d, skip=None):
    """Array of row pointers into a C-contiguous 2-D numpy array (NULL where skip[i])."""
    n = arr2d.shape[0]
    ptrs = (C.c_void_p * max(n, 1))()
    base, stride = arr2d.ctypes.data, arr2d.strides[0]
    for k in range(n):
        ptrs[k] = None if (skip is not None and skip[k]) else base + k * stride
    return ptrs


class Context:
    """One GPU context (one per process / GPU), wrapping ``ccg_ctx``."""

    def __init__(self, device=-1, multi=None):
        """device: CUDA device of a single-GPU context.  multi = N (0 = all visible) or a list of device ids makes
        an in-process multi-GPU context (ccg_init_multi / ccg_init_multi_devices) instead."""
        self._L = load()
        self._h = C.c_void_p()
        if multi is None:
            rc = self._L.ccg_init(C.byref(self._h), device)
        elif isinstance(multi, int):
            rc = self._L.ccg_init_multi(C.byref(self._h), multi)
        else:
            devs = np.ascontiguousarray(multi, dtype=np.int32)
            rc = self._L.ccg_init_multi_devices(C.byref(self._h), len(devs), devs.ctypes.data)
        if rc:
            raise CcgError(rc, self._L.ccg_last_error(None).decode())
        self.n = self.len = 0
        self.pair = True

    # ---- multi-GPU ----
    def multi_gpus(self):
        """(member GPUs, members working on the current problem)"""
        a = C.c_int(1)
        return self._L.ccg_multi_gpus(self._h, C.byref(a)), a.value

    def multi_contexts(self):
        """member device contexts started so far (member 0 at once, the others when a problem is first split)"""
        return self._L.ccg_multi_contexts(self._h)

    def group_export(self, max_samples):
        buf = (C.c_char * GROUP_HANDLE_BYTES)()
        self._ck(self._L.ccg_group_export(self._h, max_samples, buf))
        return bytes(buf)

    def group_join(self, rank, world, handles):
        blob = b"".join(handles)
        assert len(blob) == world * GROUP_HANDLE_BYTES
        self._ck(self._L.ccg_group_join(self._h, rank, world, blob))

    def group_leave(self):
        self._ck(self._L.ccg_group_leave(self._h))

    def group_set_output(self, compact):
        self._ck(self._L.ccg_group_set_output(self._h, 1 if compact else 0))

    def group_set_alignment(self, total_len, global_inc=0):
        self._ck(self._L.ccg_group_set_alignment(self._h, total_len, global_inc))


    def _ck(self, rc):
        if rc:
            msg = self._L.ccg_last_error(self._h).decode() or self._L.ccg_strerror(rc).decode()
            raise CcgError(rc, msg)

    def close(self):
        if self._h:
            self._L.ccg_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- configuration ----
    def set_stream(self, cuda_stream_ptr):
        self._ck(self._L.ccg_set_stream(self._h, cuda_stream_ptr))

    def set_kernel(self, kernel):
        self._ck(self._L.ccg_set_kernel(self._h, kernel))

    def set_partition(self, rank, world):
        self._ck(self._L.ccg_set_partition(self._h, rank, world))

    def set_tile_window(self, row_lo, row_hi=0, col_lo=0, col_hi=0):
        """Restrict the run to rows [row_lo, row_hi) x columns [col_lo, col_hi); row_lo < 0 removes the window."""
        self._ck(self._L.ccg_set_tile_window(self._h, row_lo, row_hi, col_lo, col_hi))

    def set_scratch_limit(self, nbytes):
        self._ck(self._L.ccg_set_scratch_limit(self._h, nbytes))

    def sync(self):
        self._ck(self._L.ccg_sync(self._h))

    def sample_count_masked(self, slot):
        """getNpos of a shared-mask reference candidate after maskMotifs + getIncPosPtr(seq, seq, proxi); store unchanged"""
        inc = C.c_uint(0)
        self._ck(self._L.ccg_sample_count_masked(self._h, slot, C.byref(inc)))
        return inc.value

    def set_proximity(self, proxi, snp_events_only=False):
        """-P proxi; snp_events_only selects the event definition of -f 8 / -f 32 (dist.c:802-806)."""
        self._ck(self._L.ccg_set_proximity(self._h, proxi, 1 if snp_events_only else 0))

    def set_problem(self, n, length, pair=True):
        self._ck(self._L.ccg_set_problem(self._h, n, length, 1 if pair else 0))
        self.n, self.len, self.pair = n, length, pair

    # ---- uploads ----
    def put_global_mask(self, mask):
        mask = np.ascontiguousarray(mask, dtype=np.uint32)
        assert mask.size >= words(self.len)
        self._ck(self._L.ccg_put_global_mask(self._h, mask.ctypes.data))

    def apply_global_mask(self, mask):
        mask = np.ascontiguousarray(mask, dtype=np.uint32)
        assert mask.size >= words(self.len)
        self._ck(self._L.ccg_apply_global_mask(self._h, mask.ctypes.data))

    def build_global_mask(self, include=None):
        inc = None if include is None else np.ascontiguousarray(include, dtype=np.uint8)
        g = C.c_uint(0)
        self._ck(self._L.ccg_build_global_mask(self._h, None if inc is None else inc.ctypes.data, C.byref(g)))
        return g.value

    def put_samples_packed(self, seqs, masks=None, first=0, skip=None):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint64)
        sp = _row_ptrs(seqs, skip)
        mp = None
        if masks is not None:
            masks = np.ascontiguousarray(masks, dtype=np.uint32)
            mp = _row_ptrs(masks, skip)
        self._ck(self._L.ccg_put_samples_packed(self._h, first, seqs.shape[0], sp, mp))

    def put_samples_packed_dev(self, d_seqs_ptr, d_masks_ptr, count, wstride, first=0):
        self._ck(self._L.ccg_put_samples_packed_dev(self._h, first, count, d_seqs_ptr, d_masks_ptr, wstride))

    def put_samples_packed_dev_borrowed(self, d_seqs_ptr, d_masks_ptr, count, wstride, first=0):
        """the rows are lent, not copied: keep them valid and unchanged until the run has finished"""
        self._ck(self._L.ccg_put_samples_packed_dev_borrowed(self._h, first, count, d_seqs_ptr, d_masks_ptr, wstride))

    def put_sample_codes(self, idx, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        assert codes.size == self.len
        self._ck(self._L.ccg_put_sample_codes(self._h, idx, codes.ctypes.data))

    def set_motifs(self, motifs):
        """-y: motifs = [[set | 16 * methylated, ...], ...] as getMethMotifs builds them (motif, reverse complement, ...)."""
        lens = np.array([len(m) for m in motifs], dtype=np.int32)
        sets = np.array([v for m in motifs for v in m], dtype=np.uint8)
        self._ck(self._L.ccg_set_motifs(self._h, len(motifs), lens.ctypes.data if len(motifs) else None,
                                        sets.ctypes.data if len(motifs) else None))

    def mask_motifs(self, first=0, count=None):
        """maskMotifs (meth.c:141) on slots [first, first+count) -> their included counts afterwards."""
        count = self.n - first if count is None else count
        out = np.zeros(max(count, 1), dtype=np.uint32)
        self._ck(self._L.ccg_mask_motifs(self._h, first, count, out.ctypes.data))
        return out[:count]

    def sample_proximity(self, first=0, count=None, apply=True):
        """getIncPosPtr(includes[i], seq, seq, proxi) of cdist.c:91 on slots [first, first+count) -> their counts."""
        count = self.n - first if count is None else count
        out = np.zeros(max(count, 1), dtype=np.uint32)
        self._ck(self._L.ccg_sample_proximity(self._h, first, count, 1 if apply else 0, out.ctypes.data))
        return out[:count]

    def inc_counts(self):
        out = np.zeros(max(self.n, 1), dtype=np.uint32)
        self._ck(self._L.ccg_get_inc_counts(self._h, out.ctypes.data))
        return out[:self.n]

    # ---- runs ----
    def run_pair(self, include=None, norm=0, min_length=1, min_cov=0.5, elem_size=8, byte_scale=1.0, want_n=True):
        dt = ELEM_DTYPE[elem_size]
        D = np.zeros(max(cells(self.n), 1), dtype=dt)
        N = np.zeros(max(cells(self.n), 1), dtype=dt) if want_n else None
        inc = None if include is None else np.ascontiguousarray(include, dtype=np.uint8)
        dn = C.c_int(0)
        self._ck(self._L.ccg_run_pair(self._h, None if inc is None else inc.ctypes.data, norm, min_length, min_cov,
                                      elem_size, byte_scale, D.ctypes.data, N.ctypes.data if want_n else None,
                                      C.byref(dn)))
        k = cells(dn.value)
        return D[:k], (N[:k] if want_n else None), dn.value

    def run_global(self, include=None, norm=0, elem_size=8, byte_scale=1.0):
        D = np.zeros(max(cells(self.n), 1), dtype=ELEM_DTYPE[elem_size])
        inc = None if include is None else np.ascontiguousarray(include, dtype=np.uint8)
        dn, ginc = C.c_int(0), C.c_uint(0)
        self._ck(self._L.ccg_run_global(self._h, None if inc is None else inc.ctypes.data, norm, elem_size,
                                        byte_scale, D.ctypes.data, C.byref(dn), C.byref(ginc)))
        return D[:cells(dn.value)], dn.value, ginc.value

    def run_pair_dev(self, d_D_ptr, d_N_ptr, include=None, norm=0, min_length=1, min_cov=0.5, elem_size=8,
                     byte_scale=1.0):
        inc = None if include is None else np.ascontiguousarray(include, dtype=np.uint8)
        dn = C.c_int(0)
        self._ck(self._L.ccg_run_pair_dev(self._h, None if inc is None else inc.ctypes.data, norm, min_length,
                                          min_cov, elem_size, byte_scale, d_D_ptr, d_N_ptr, C.byref(dn)))
        return dn.value

    def run_global_dev(self, d_D_ptr, include=None, norm=0, elem_size=8, byte_scale=1.0):
        inc = None if include is None else np.ascontiguousarray(include, dtype=np.uint8)
        dn, ginc = C.c_int(0), C.c_uint(0)
        self._ck(self._L.ccg_run_global_dev(self._h, None if inc is None else inc.ctypes.data, norm, elem_size,
                                            byte_scale, d_D_ptr, C.byref(dn), C.byref(ginc)))
        return dn.value, ginc.value

    def run_row(self, row_slot, norm=0, min_length=1, min_cov=0.5, want_n=True):
        """cmpFsaRowThrd (fsacmpthrd.c:482): slot row_slot against every uploaded slot below it -> (D, N) doubles."""
        D = np.zeros(max(row_slot, 1), dtype=np.float64)
        N = np.zeros(max(row_slot, 1), dtype=np.float64) if want_n else None
        cols = C.c_int(0)
        self._ck(self._L.ccg_run_row(self._h, row_slot, norm, min_length, min_cov, D.ctypes.data,
                                     N.ctypes.data if want_n else None, C.byref(cols)))
        return D[:cols.value], (N[:cols.value] if want_n else None)

    def list_variants(self, pair=True, include=None, row=None):
        """-V: [((sample_i, sample_j), [(label, code_i, code_j), ...]), ...] in the order the callback delivers them;
        row = slot restricts the listing to that sample against the slots below it (-V with -a)."""
        out = []

        def take(user, si, sj, ptr, count):
            v = np.ctypeslib.as_array(ptr, shape=(count,))
            out.append(((si, sj), list(zip((v >> np.uint64(4)).tolist(), ((v >> np.uint64(2)) & np.uint64(3)).tolist(),
                                           (v & np.uint64(3)).tolist()))))
            return 0
        if row is not None:
            self._ck(self._L.ccg_list_variants_row(self._h, row, VARIANT_FN(take), None))
            return out
        inc = None if include is None else np.ascontiguousarray(include, dtype=np.uint8)
        self._ck(self._L.ccg_list_variants(self._h, 1 if pair else 0, None if inc is None else inc.ctypes.data,
                                           VARIANT_FN(take), None))
        return out

    def raw_counts(self, dn):
        mism = np.zeros(max(cells(dn), 1), dtype=np.uint32)
        ninc = np.zeros(max(cells(dn), 1), dtype=np.uint32)
        self._ck(self._L.ccg_get_raw_counts(self._h, mism.ctypes.data, ninc.ctypes.data))
        return mism[:cells(dn)], ninc[:cells(dn)]

    # ---- count-matrix (.mat) path ----
    def mat_set_problem(self, n, max_len):
        self._ck(self._L.ccg_mat_set_problem(self._h, n, max_len))
        self.mat_n = n

    def mat_put_sample(self, idx, counts6, totals=None):
        counts6 = np.ascontiguousarray(counts6, dtype=np.uint16).reshape(-1, 6)
        tp = None
        if totals is not None:
            totals = np.ascontiguousarray(totals, dtype=np.uint32)
            tp = totals.ctypes.data
        self._ck(self._L.ccg_mat_put_sample(self._h, idx, counts6.ctypes.data, tp, counts6.shape[0]))

    def mat_run(self, include=None, method="cos", alpha=0.05, norm=0, min_depth=15, min_length=1, min_cov=0.5,
                elem_size=8, byte_scale=1.0):
        mid, order = mat_method(method)
        n = self.mat_n
        dt = ELEM_DTYPE[elem_size]
        D = np.zeros(max(cells(n), 1), dtype=dt)
        N = np.zeros(max(cells(n), 1), dtype=dt)
        rows = np.zeros(max(cells(n), 1), dtype=np.uint32)
        inc = None if include is None else np.ascontiguousarray(include, dtype=np.uint8)
        dn = C.c_int(0)
        self._ck(self._L.ccg_mat_run(self._h, None if inc is None else inc.ctypes.data, mid, order, alpha, norm, min_depth,
                                     min_length, min_cov, elem_size, byte_scale, D.ctypes.data, N.ctypes.data,
                                     C.byref(dn), rows.ctypes.data))
        k = cells(dn.value)
        return D[:k], N[:k], dn.value, rows[:k]

    def mat_run_row(self, row_slot, method="cos", alpha=0.05, norm=0, min_depth=15, min_length=1, min_cov=0.5):
        """cmpMatRowThrd (ltdmatrixthrd.c:111): slot row_slot against the slots below it -> (D, N, rows) per column."""
        mid, order = mat_method(method)
        D = np.zeros(max(row_slot, 1), dtype=np.float64)
        N = np.zeros(max(row_slot, 1), dtype=np.float64)
        rows = np.zeros(max(row_slot, 1), dtype=np.uint32)
        self._ck(self._L.ccg_mat_run_row(self._h, row_slot, mid, order, alpha, norm, min_depth, min_length, min_cov,
                                         D.ctypes.data, N.ctypes.data, rows.ctypes.data))
        return D[:row_slot], N[:row_slot], rows[:row_slot]

    # ---- introspection ----
    @property
    def launches(self):
        return self._L.ccg_launch_count(self._h)

    @property
    def last_kernel(self):
        return self._L.ccg_last_kernel(self._h).decode()

    def last_compare_ms(self):
        return self._L.ccg_last_compare_ms(self._h)

    def last_phase_ms(self, phase):
        return self._L.ccg_last_phase_ms(self._h, phase)

    def measure_fp4_peak(self, target_ms=20.0, check_sum=1.6e7):
        """(e2m1 TOP/s of a loads-free tcgen05 kind::mxf4 loop on every CTA pair, accumulator elements that
        differed from the exact integer after summing +1/-1 products up to about check_sum)."""
        bad = C.c_longlong(-1)
        tops = self._L.ccg_measure_fp4_peak(self._h, target_ms, check_sum, C.byref(bad), None)
        return tops, bad.value

    def measure_i8_peak(self, target_ms=20.0):
        """int8 TOP/s of a loads-free tcgen05 kind::i8 loop on every CTA pair (roofline denominator)."""
        return self._L.ccg_measure_i8_peak(self._h, target_ms)


def fsa_cmp_thread_out(seqs, include, includes, length, pair=True, norm=0, min_length=1, min_cov=0.5, proxi=0,
                       elem_size=8, byte_scale=1.0, want_n=True, ctx=None, rows=None):
    """Drop-in for the reference's ``fsaCmpThreadOut`` call (cdist.c:181/184).

    seqs (n, W) u64 and includes (n, W) u32 -- or (1, W) in shared-mask mode --
    are HOST arrays in the reference's packed formats; include is (n,) u8.
    Returns (D, N, Dn, global_inc): packed lower-triangular cells over the
    included samples.  rows = (list of per-sample u64 arrays, list of per-sample
    u32 arrays) passes separately allocated rows, as the reference holds them
    (dist.c:143-154), instead of the rows of the 2-D arrays.
    """
    L = load()
    seqs = np.ascontiguousarray(seqs, dtype=np.uint64)
    includes = np.ascontiguousarray(includes, dtype=np.uint32)
    n = seqs.shape[0]
    include = np.ascontiguousarray(include, dtype=np.uint8)
    dt = ELEM_DTYPE[elem_size]
    D = np.zeros(max(cells(n), 1), dtype=dt)
    N = np.zeros(max(cells(n), 1), dtype=dt) if (want_n and pair) else None
    sp = _row_ptrs(seqs)
    if rows is not None:
        for k in range(n):
            sp[k] = rows[0][k].ctypes.data
    if pair:
        mp = _row_ptrs(includes)
        if rows is not None:
            for k in range(n):
                mp[k] = rows[1][k].ctypes.data
    else:
        mp = (C.c_void_p * max(n, 1))()
        for k in range(max(n, 1)):
            mp[k] = includes.ctypes.data
    dn, ginc = C.c_int(0), C.c_uint(0)
    rc = L.ccg_fsa_cmp_thread_out(ctx._h if ctx else None, 1 if pair else 0, D.ctypes.data,
                                  N.ctypes.data if N is not None else None, elem_size, byte_scale, n, length, sp,
                                  include.ctypes.data, mp, norm, min_length, min_cov, proxi, C.byref(dn),
                                  C.byref(ginc))
    if rc:
        msg = L.ccg_last_error(ctx._h if ctx else None).decode() or L.ccg_strerror(rc).decode()
        raise CcgError(rc, msg)
    if ctx is not None:
        ctx.n, ctx.len, ctx.pair = n, length, bool(pair)     # the call declared this problem on ctx
    k = cells(dn.value)
    return D[:k], (N[:k] if N is not None else None), dn.value, ginc.value
