"""ctypes bindings for the CPU oracle (liboracle.so) and, when built, the
unmodified reference (oracle/_ref/libccphylo_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by ccphylo_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libccphylo_ref.so")
REF_BIN = os.path.join(_HERE, "_ref", "ccphylo")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")

ELEM_DTYPE = {8: np.float64, 4: np.float32, 2: np.uint16, 1: np.uint8}


def build(ref=True):
    """Compile the checker (and the reference when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", _HERE] + ([] if ref else ["liboracle.so"]), check=True)


def words(length):
    return (length >> 5) + (1 if length & 31 else 0)


def cells(dn):
    return dn * (dn - 1) // 2 if dn > 1 else 0


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.orc_translate.restype = C.c_long
        L.orc_translate.argtypes = [_u8p, C.c_long, C.c_uint, _u8p]
        L.orc_pack.restype = C.c_int
        L.orc_pack.argtypes = [_u8p, C.c_int, _u64p]
        L.orc_known_mask.restype = C.c_int
        L.orc_known_mask.argtypes = [_u8p, C.c_int, _u32p]
        L.orc_and_known.restype = None
        L.orc_and_known.argtypes = [_u32p, _u8p, _u8p, C.c_int]
        L.orc_inc_pos.restype = None
        L.orc_inc_pos.argtypes = [_u32p, _u8p, _u8p, C.c_int, C.c_uint, C.c_int]
        L.orc_trim_pass.restype = None
        L.orc_trim_pass.argtypes = [_u32p, _u8p, C.c_void_p, C.c_int, C.c_uint, C.c_int, C.c_void_p]
        L.orc_pair_counts_proxi.restype = None
        L.orc_pair_counts_proxi.argtypes = [_u64p, _u64p, _u32p, _u32p, C.c_int, C.c_uint, C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_uint32)]
        L.orc_mask_proxi.restype = None
        L.orc_mask_proxi.argtypes = [_u64p, _u64p, _u32p, _u32p, C.c_int, C.c_uint, _u32p]
        L.orc_fsa_cmp_pair_proxi.restype = C.c_int
        L.orc_fsa_cmp_pair_proxi.argtypes = [C.c_int, C.c_int, _u64p, C.c_long, _u8p, _u32p, C.c_uint, C.c_uint,
                                             C.c_double, C.c_uint, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_mask_motifs.restype = C.c_long
        L.orc_mask_motifs.argtypes = [_u64p, _u32p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_list_variants.restype = C.c_long
        L.orc_list_variants.argtypes = [_u64p, _u64p, _u32p, C.c_int, C.c_void_p, C.c_long]
        L.orc_mask_count.restype = C.c_int
        L.orc_mask_count.argtypes = [_u32p, C.c_int]
        L.orc_raw_pair_matrix.restype = None
        L.orc_raw_pair_matrix.argtypes = [C.c_int, C.c_int, _u64p, _u32p, C.c_long, _u32p, _u32p, C.c_int]
        L.orc_fsa_cmp_pair.restype = C.c_int
        L.orc_fsa_cmp_pair.argtypes = [C.c_int, C.c_int, _u64p, C.c_long, _u8p, _u32p, C.c_uint, C.c_uint,
                                       C.c_double, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_fsa_cmp_global.restype = C.c_int
        L.orc_fsa_cmp_global.argtypes = [C.c_int, C.c_int, _u64p, C.c_long, _u8p, _u32p, C.c_uint, C.c_int,
                                         C.c_double, C.c_void_p, C.POINTER(C.c_uint)]
        _u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
        _i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        _f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.orc_mat_matrix.restype = C.c_int
        L.orc_mat_matrix.argtypes = [C.c_int, C.c_uint, C.c_double, C.c_int, C.c_long, _u16p, _u32p, _i32p, _u8p,
                                     C.c_uint, C.c_uint, C.c_uint, C.c_double, _f64p, _f64p]
        _lib = L
    return _lib


# ----------------------------------------------------------------------------
# count-matrix (.mat) path: mat_oracle.c
# ----------------------------------------------------------------------------
MAT_METHODS = ["cos", "z", "chi2", "nchi2", "c", "nc", "p", "np", "bc", "nbc", "l1", "l2", "linf", "ln", "nl1", "nl2",
               "nlinf", "nln"]


def write_mat(path, name, refbases, counts6):
    """one sample's KMA .mat text file (bench / test infrastructure for runs of the reference binary)"""
    L = lib()
    L.orc_write_mat.restype = C.c_int
    L.orc_write_mat.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_long]
    refbases = np.ascontiguousarray(refbases, dtype=np.uint8)
    counts6 = np.ascontiguousarray(counts6, dtype=np.uint16)
    if L.orc_write_mat(path.encode(), name.encode(), refbases.ctypes.data, counts6.ctypes.data, counts6.shape[0]) != 0:
        raise OSError(f"cannot write {path}")


def mat_method(name):
    """'-d' name -> (method id, order) with the reference's precedence (dist.c:738-786)."""
    if name in MAT_METHODS and name not in ("ln", "nln"):
        return MAT_METHODS.index(name), 0
    if name.startswith("l"):
        return MAT_METHODS.index("ln"), int(name[1:])
    if name.startswith("nl"):
        return MAT_METHODS.index("nln"), int(name[2:])
    raise ValueError(name)


def mat_matrix(counts, totals, lens, include=None, method="cos", alpha=0.05, norm=0, min_depth=15, min_length=1,
               min_cov=0.5):
    """counts (n, L, 6) u16 [A,C,G,T,-,N], totals (n, L) u32, lens (n,) -> (D, N, Dn) packed doubles
    over the included samples (cmpMats + ltdMatrixThrd); D = -1 / N = 0 without sufficient overlap."""
    counts = np.ascontiguousarray(counts, dtype=np.uint16)
    totals = np.ascontiguousarray(totals, dtype=np.uint32)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    n, lmax = totals.shape
    inc = np.ones(n, np.uint8) if include is None else np.ascontiguousarray(include, dtype=np.uint8)
    mid, order = mat_method(method)
    D = np.zeros(max(cells(n), 1))
    N = np.zeros(max(cells(n), 1))
    dn = lib().orc_mat_matrix(mid, order, alpha, n, lmax, counts, totals, lens, inc, norm, min_depth, min_length,
                              min_cov, D, N)
    if dn < 0:
        raise RuntimeError("a sample fails its own inclusion gate inside a pair (the reference exits there)")
    return D[:cells(dn)], N[:cells(dn)], dn


# ----------------------------------------------------------------------------
# oracle (our restatement)
# ----------------------------------------------------------------------------
def translate(data: bytes, flag=1):
    buf = np.frombuffer(data, dtype=np.uint8)
    buf = np.ascontiguousarray(buf)
    codes = np.empty(max(len(buf), 1), dtype=np.uint8)
    n = lib().orc_translate(buf if len(buf) else np.zeros(1, np.uint8), len(buf), flag, codes)
    return codes[:n].copy()


def pack(codes):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    W = words(len(codes))
    seq = np.zeros(max(W, 1), dtype=np.uint64)
    unknown = lib().orc_pack(codes if len(codes) else np.zeros(1, np.uint8), len(codes), seq)
    return seq[:W], unknown


def known_mask(codes):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    W = words(len(codes))
    mask = np.zeros(max(W, 1), dtype=np.uint32)
    inc = lib().orc_known_mask(codes if len(codes) else np.zeros(1, np.uint8), len(codes), mask)
    return mask[:W], inc


def variant_of(flag):
    """-f bits 32 / 8 select getIncPosInsigPrune / getIncPosInsig (dist.c:802-806): events are SNPs only."""
    return 1 if flag & (32 | 8) else 0


def inc_pos(mask, seq_codes, ref_codes, proxi=0, variant=0):
    """getIncPosPtr(mask, seq, ref, proxi) on an existing mask (in place)."""
    seq_codes = np.ascontiguousarray(seq_codes, dtype=np.uint8)
    ref_codes = np.ascontiguousarray(ref_codes, dtype=np.uint8)
    if len(seq_codes):
        lib().orc_inc_pos(mask, seq_codes, ref_codes, len(seq_codes), proxi, variant)
    return mask


def trim_pass(mask, seq_codes, ref_codes=None, proxi=0, builder=0, columns=None):
    """`ccphylo trim`: one sample's pass over `mask` on trim's alphabet (orc_trim_pass; in place).  seq_codes is modified
    as the reference modifies it (soft flags stripped); ref_codes None = the sample against itself."""
    assert seq_codes.dtype == np.uint8 and seq_codes.flags.c_contiguous
    ref_ptr = None if ref_codes is None else np.ascontiguousarray(ref_codes, dtype=np.uint8).ctypes.data
    col_ptr = None if columns is None else columns.ctypes.data
    if len(seq_codes):
        lib().orc_trim_pass(mask, seq_codes, ref_ptr, len(seq_codes), proxi, builder, col_ptr)
    return mask


def full_mask(length):
    """initIncPos (fsacmp.c:164)."""
    W = words(length)
    m = np.full(max(W, 1), 0xFFFFFFFF, dtype=np.uint32)
    if length & 31:
        m[W - 1] = (0xFFFFFFFF << (32 - (length & 31))) & 0xFFFFFFFF
    return m[:W]


def encode_samples(codes2d, proxi=0, variant=0):
    """codes2d: (n, L) u8 codes -> (seqs (n,W) u64, masks (n,W) u32, inc (n,) i32); pair mode
    (cdist.c:88-92: initIncPos, qseq2nibble, getIncPosPtr(seq, seq, proxi), getNpos)."""
    n, L = codes2d.shape
    W = words(L)
    seqs = np.zeros((n, W), dtype=np.uint64)
    masks = np.zeros((n, W), dtype=np.uint32)
    inc = np.zeros(n, dtype=np.int32)
    for i in range(n):
        seqs[i], _ = pack(codes2d[i])
        if proxi:
            m = full_mask(L).copy()
            inc_pos(m, codes2d[i], codes2d[i], proxi, variant)
            masks[i] = m
            inc[i] = lib().orc_mask_count(m, L) if W else 0
        else:
            masks[i], inc[i] = known_mask(codes2d[i])
    return seqs, masks, inc


def global_mask(codes2d, include, proxi=0, variant=0):
    """cdist.c:86-112 accumulation: first included sample is `ref`."""
    n, L = codes2d.shape
    inc_idx = [i for i in range(n) if include[i]]
    ref = np.ascontiguousarray(codes2d[inc_idx[0]])
    if proxi:
        mask = full_mask(L).copy()
        inc_pos(mask, ref, ref, proxi, variant)
        for i in inc_idx[1:]:
            inc_pos(mask, codes2d[i], ref, proxi, variant)
        return mask
    mask, _ = known_mask(ref)
    mask = mask.copy()
    for i in inc_idx[1:]:
        lib().orc_and_known(mask, np.ascontiguousarray(codes2d[i]), ref, L)
    return mask


def pair_counts_proxi(seq_i, seq_j, inc_i, inc_j, length, proxi):
    m, n = C.c_uint32(0), C.c_uint32(0)
    lib().orc_pair_counts_proxi(np.ascontiguousarray(seq_i), np.ascontiguousarray(seq_j), np.ascontiguousarray(inc_i),
                                np.ascontiguousarray(inc_j), length, proxi, C.byref(m), C.byref(n))
    return m.value, n.value


def pair_mask_proxi(seq_i, seq_j, inc_i, inc_j, length, proxi):
    """maskProxi: the pair's mask the counts (and the -V walk, fsacmpthrd.c:410-414) are taken under"""
    out = np.zeros(max(words(length), 1), dtype=np.uint32)
    lib().orc_mask_proxi(np.ascontiguousarray(seq_i), np.ascontiguousarray(seq_j), np.ascontiguousarray(inc_i),
                         np.ascontiguousarray(inc_j), length, proxi, out)
    return out[:words(length)]


def fsa_cmp_row(seqs, masks, row, length, norm=0, min_length=1, min_cov=0.5, proxi=0, variant=0, codes=None):
    """cmpFsaRowThrd (fsacmpthrd.c:482-580): sample `row` against samples 0..row-1.  The pair mask is the new
    sample's own mask (includeadd, built :627-628 with its own builder) put through the per-sample builder against
    the column sample (:545-546: getIncPosPtr(includeseq, seq, ref, proxi)), counts by fsacmpair (:555); cell rule
    :559-569: D = norm ? mism * norm / inc : mism (norm is a double there), D = -1 and N = 0 below the gate.
    masks[row] is includeadd; with proxi > 0 the translated codes are needed for the builder's events."""
    gate = max(min_length, int(min_cov * length)) if min_length < min_cov * length else min_length
    D = np.zeros(row, dtype=np.float64)
    N = np.zeros(row, dtype=np.float64)
    ones = np.full(len(masks[row]), 0xFFFFFFFF, dtype=np.uint32)
    for j in range(row):
        if proxi:
            pm = np.ascontiguousarray(masks[row]).copy()
            inc_pos(pm, codes[j], codes[row], proxi, variant)
            m, inc = pair_counts_proxi(seqs[row], seqs[j], pm, ones, length, 0)
        else:
            m, inc = pair_counts_proxi(seqs[row], seqs[j], masks[row], masks[j], length, 0)
        if gate <= inc:
            D[j] = (np.float64(m) * np.float64(norm) / np.float64(inc)) if norm else np.float64(m)
            N[j] = inc
        else:
            D[j] = -1.0
            N[j] = 0.0
    return D, N


_IUPAC = {"a": 1, "c": 2, "g": 4, "t": 8, "u": 8, "r": 5, "y": 10, "s": 6, "w": 9, "k": 12, "m": 3, "b": 14, "d": 13, "h": 11,
          "v": 7, "x": 15, "n": 15}


def parse_motifs(text, as_built=False):
    """methparse.c:27-81,105-175,268-296: FASTA-like motif file -> [(sets, ...)], each motif followed by its reverse
    complement; lower case = plain position, upper case = methylation site (bit 4); anything else is dropped.

    as_built: qseq2methMotif pads the alternative words of a position that has fewer bases than the motif's most
    ambiguous one with `bases[*seq & 31]` (methparse.c:231-236); for an upper-case letter that index is >= 16 and
    runs past the 16-entry array -- undefined behaviour.  In the binary built here (gcc -O3, oracle/Makefile) the
    32-entry `nums` follows `bases` on the stack, so the padding base is popcount(set): such a position also
    accepts the base whose code equals the number of bases its letter stands for.  as_built=True reproduces that
    (pinned in tests/test_oracle_vs_reference.py); the default is the evident intention, which is also what the
    CUDA path implements."""
    motifs, cur = [], None
    lines = text.split("\n")
    for k, line in enumerate(lines):
        if line.startswith(">") and (cur is None or True):
            if cur:
                motifs.append(cur)
            cur = []
            continue
        if cur is None:
            cur = []
        for ch in line:
            if ch.lower() in _IUPAC and ch not in "-.":
                cur.append(_IUPAC[ch.lower()] | (16 if ch.isupper() else 0))
    if cur:
        motifs.append(cur)
    out = []
    for m in motifs:
        rc = []
        for v in reversed(m):
            s = v & 15
            comp = ((s & 1) << 3) | ((s & 2) << 1) | ((s & 4) >> 1) | ((s & 8) >> 3)
            rc.append(comp | (v & 16))
        if len(m) & 1:
            # strrcMeth (methparse.c:83-103) swaps and complements pairs from both ends, then "complements the
            # middle" through a pointer that still stands on the last LEFT element: that one is complemented a
            # second time (back to the letter it was swapped with) and the middle letter is never complemented.
            # (A one-letter motif makes it write in front of the buffer: not restated.)
            mid = len(m) >> 1
            rc[mid] = m[mid]
            if mid >= 1:
                rc[mid - 1] = m[len(m) - mid]
        for mm in (m, rc):
            if as_built:
                num = max(bin(v & 15).count("1") for v in mm)
                mm = [v | (1 << bin(v & 15).count("1")) if (v & 16) and bin(v & 15).count("1") < num else v for v in mm]
            out.append(mm)
    return out


def mask_motifs(seq, mask, length, motifs):
    """maskMotifs (meth.c:141) in place on `mask`; returns the number of matches"""
    lens = np.array([len(m) for m in motifs], dtype=np.int32)
    sets = np.array([v for m in motifs for v in m], dtype=np.uint8)
    if length == 0 or len(motifs) == 0:
        return 0
    return lib().orc_mask_motifs(np.ascontiguousarray(seq), mask, length, len(motifs), lens.ctypes.data, sets.ctypes.data)


def list_variants(seq_i, seq_j, mask, length):
    """-V: [(label, code_i, code_j), ...] of one pair under `mask`, as fsacmpairint / fsacmprint print them."""
    seq_i, seq_j, mask = np.ascontiguousarray(seq_i), np.ascontiguousarray(seq_j), np.ascontiguousarray(mask)
    k = lib().orc_list_variants(seq_i, seq_j, mask, length, None, 0)
    out = np.zeros(max(k, 1), dtype=np.uint64)
    lib().orc_list_variants(seq_i, seq_j, mask, length, out.ctypes.data, k)
    return [(int(v >> 4), int((v >> 2) & 3), int(v & 3)) for v in out[:k]]


def variant_text(si, sj, variants):
    """printDiff (fsacmp.c:635-644)"""
    return "".join("(%d, %d)\t%c%d%c\n" % (si, sj, "ACGT"[a], pos, "ACGT"[b]) for pos, a, b in variants)


def raw_pair_matrix(seqs, masks, length, nthreads=8):
    n, W = seqs.shape
    mism = np.zeros(max(cells(n), 1), dtype=np.uint32)
    ninc = np.zeros(max(cells(n), 1), dtype=np.uint32)
    lib().orc_raw_pair_matrix(n, length, seqs, masks, W, mism, ninc, nthreads)
    return mism[:cells(n)], ninc[:cells(n)]


def fsa_cmp_pair(seqs, masks, include, length, norm=0, min_length=1, min_cov=0.5, elem_size=8, byte_scale=1.0,
                 want_n=True, proxi=0):
    n, W = seqs.shape
    include = np.ascontiguousarray(include, dtype=np.uint8)
    dt = ELEM_DTYPE[elem_size]
    D = np.zeros(max(cells(n), 1), dtype=dt)
    N = np.zeros(max(cells(n), 1), dtype=dt)
    dn = lib().orc_fsa_cmp_pair_proxi(n, length, np.ascontiguousarray(seqs), W, include, np.ascontiguousarray(masks),
                                      norm, min_length, min_cov, proxi, elem_size, byte_scale, D.ctypes.data,
                                      N.ctypes.data if want_n else None)
    return D[:cells(dn)], (N[:cells(dn)] if want_n else None), dn


def fsa_cmp_global(seqs, mask, include, length, norm=0, elem_size=8, byte_scale=1.0):
    n, W = seqs.shape
    include = np.ascontiguousarray(include, dtype=np.uint8)
    D = np.zeros(max(cells(n), 1), dtype=ELEM_DTYPE[elem_size])
    ginc = C.c_uint(0)
    dn = lib().orc_fsa_cmp_global(n, length, np.ascontiguousarray(seqs), W, include,
                                  np.ascontiguousarray(mask), norm, elem_size, byte_scale, D.ctypes.data,
                                  C.byref(ginc))
    return D[:cells(dn)], dn, ginc.value


# ----------------------------------------------------------------------------
# the unmodified reference (oracle/_ref), when present
# ----------------------------------------------------------------------------
_ref = None


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        R = C.CDLL(REF_SO)
        R.refshim_encode.restype = C.c_long
        R.refshim_encode.argtypes = [_u8p, C.c_long, C.c_uint, C.c_uint, _u8p, _u64p, _u32p,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
        R.refshim_and_known.restype = None
        R.refshim_and_known.argtypes = [_u32p, _u8p, _u8p, C.c_int, C.c_uint]
        R.refshim_inc_pos.restype = None
        R.refshim_inc_pos.argtypes = [_u32p, _u8p, _u8p, C.c_int, C.c_uint, C.c_uint]
        R.refshim_phy_names.restype = C.c_int
        R.refshim_phy_names.argtypes = [C.c_char_p, C.c_char_p, C.c_char, C.c_char_p, C.c_long]
        R.refshim_phy_update.restype = None
        R.refshim_phy_update.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_double), C.c_uint, C.c_int]
        R.refshim_mask_motifs.restype = C.c_int
        R.refshim_mask_motifs.argtypes = [C.c_char_p, _u64p, _u32p, C.c_int]
        R.refshim_variants.restype = C.c_uint64
        R.refshim_variants.argtypes = [C.c_int, C.c_int, C.c_int, _u64p, _u64p, _u32p, C.c_int, C.c_char_p, C.c_long]
        R.refshim_mask_proxi.restype = None
        R.refshim_mask_proxi.argtypes = [_u64p, _u64p, _u32p, _u32p, C.c_int, C.c_uint, _u32p]
        R.refshim_pair.restype = C.c_uint64
        R.refshim_pair.argtypes = [_u64p, _u64p, _u32p, _u32p, C.c_int, C.c_uint]
        R.refshim_fsa_cmp.restype = C.c_int
        R.refshim_fsa_cmp.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u64p, C.c_long, _u8p, _u32p,
                                      C.c_uint, C.c_uint, C.c_double, C.c_uint, C.c_int, C.c_double,
                                      C.c_void_p, C.c_void_p]
        _ref = R
    return _ref


def ref_encode(data: bytes, flag=1, proxi=0):
    buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8))
    nb = len(buf)
    codes = np.zeros(nb + 1, dtype=np.uint8)
    W = nb // 32 + 2
    seq = np.zeros(W, dtype=np.uint64)
    mask = np.zeros(W, dtype=np.uint32)
    unk, inc = C.c_int(0), C.c_int(0)
    L = ref().refshim_encode(buf if nb else np.zeros(1, np.uint8), nb, flag, proxi, codes, seq, mask,
                             C.byref(unk), C.byref(inc))
    return codes[:L].copy(), seq[:words(L)].copy(), mask[:words(L)].copy(), unk.value, inc.value


def ref_inc_pos(mask, seq_codes, ref_codes, proxi=0, flag=1):
    """the reference's getIncPos / getIncPosInsig / getIncPosInsigPrune (by -f flag) on `mask`, in place"""
    seq_codes = np.ascontiguousarray(seq_codes, dtype=np.uint8).copy()
    ref_codes = np.ascontiguousarray(ref_codes, dtype=np.uint8).copy()
    ref().refshim_inc_pos(mask, seq_codes, ref_codes, len(seq_codes), proxi, flag)
    return mask


def ref_pair(seq_i, seq_j, inc_i, inc_j, length, proxi=0):
    """maskProxi + fsacmpair of the reference -> (mismatches, included)"""
    W = words(length)

    def pad(a, dt):
        b = np.zeros(W + 2, dtype=dt)
        b[:W] = a
        return b
    r = ref().refshim_pair(pad(seq_i, np.uint64), pad(seq_j, np.uint64), pad(inc_i, np.uint32), pad(inc_j, np.uint32),
                           length, proxi)
    return int(r >> 32), int(r & 0xFFFFFFFF)


def ref_mask_proxi(seq_i, seq_j, inc_i, inc_j, length, proxi):
    """maskProxi of the reference -> the pair's mask"""
    W = words(length)

    def pad(a, dt):
        b = np.zeros(W + 2, dtype=dt)
        b[:W] = a
        return b
    out = np.zeros(W + 2, dtype=np.uint32)
    ref().refshim_mask_proxi(pad(seq_i, np.uint64), pad(seq_j, np.uint64), pad(inc_i, np.uint32), pad(inc_j, np.uint32),
                             length, proxi, out)
    return out[:W]


def ref_phy_names(phy_path, directory, sep="\t"):
    """getSizePhy + getFilenamesPhy of the reference -> (n or error code, [names])"""
    buf = C.create_string_buffer(1 << 20)
    n = ref().refshim_phy_names(phy_path.encode(), directory.encode(), sep.encode(), buf, len(buf))
    return n, buf.value.decode().split("\n")[:-1]


def ref_phy_update(phy_path, n, name, row, flag=1, precision=9):
    """printphyUpdate of the reference on the file"""
    row = np.ascontiguousarray(row, dtype=np.float64)
    ref().refshim_phy_update(phy_path.encode(), n, C.create_string_buffer(name.encode()), row.ctypes.data_as(C.POINTER(C.c_double)),
                             flag, precision)


def ref_mask_motifs(motif_path, seq, mask, length):
    """the reference's getMethMotifs + maskMotifs on one packed sequence; mask (W + 2 words) is updated in place"""
    W = words(length)
    s = np.zeros(W + 2, dtype=np.uint64)
    s[:W] = seq
    return ref().refshim_mask_motifs(motif_path.encode(), s, mask, length)


def ref_variants(pair, si, sj, seq_i, seq_j, mask, length):
    """the text the reference's fsacmpairint (pair) / fsacmprint prints for one pair, and its return value"""
    W = words(length)

    def pad(a, dt):
        b = np.zeros(W + 2, dtype=dt)
        b[:W] = a
        return b
    buf = C.create_string_buffer(64 * (length + 1))
    r = ref().refshim_variants(1 if pair else 0, si, sj, pad(seq_i, np.uint64), pad(seq_j, np.uint64), pad(mask, np.uint32),
                               length, buf, len(buf))
    return buf.value.decode(), int(r)


def ref_fsa_cmp(seqs, masks, include, length, pair=True, tnum=1, norm=0, min_length=1, min_cov=0.5, proxi=0,
                elem_size=8, byte_scale=1.0, want_n=True):
    n, W = seqs.shape
    include = np.ascontiguousarray(include, dtype=np.uint8).copy()
    dt = ELEM_DTYPE[elem_size]
    D = np.zeros(max(cells(n), 1), dtype=dt)
    N = np.zeros(max(cells(n), 1), dtype=dt)
    masks = np.ascontiguousarray(masks)
    # the reference scratch-reads one word past ceil(len/32) when len%32==0 (fsacmpthrd.c:336 sizes len/32+1)
    seqs_p = np.zeros((n, W + 1), dtype=np.uint64)
    seqs_p[:, :W] = seqs
    if pair:
        masks_p = np.zeros((n, W + 1), dtype=np.uint32)
        masks_p[:, :W] = masks
    else:
        masks_p = np.zeros((1, W + 1), dtype=np.uint32)
        masks_p[0, :W] = masks.reshape(-1)[:W]
    dn = ref().refshim_fsa_cmp(tnum, 1 if pair else 0, n, length, seqs_p, W + 1, include, masks_p, norm,
                               min_length, min_cov, proxi, elem_size, byte_scale, D.ctypes.data,
                               N.ctypes.data if (want_n and pair) else None)
    return D[:cells(dn)], (N[:cells(dn)] if (want_n and pair) else None), dn
