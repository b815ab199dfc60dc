/*
 * ccphylo_gpu.h -- C-ABI of the B200 (sm_100a) implementation of ccphylo's
 * `dist` hot path: the all-vs-all pairwise nucleotide distance matrix D and
 * the inclusion-count matrix N over KMA consensus alignments.
 *
 * This is the drop-in boundary.  The reference has no FFI; its seam is the
 * fan-out call
 *
 *     fsaCmpThreadOut(tnum, &cmpairFsaThrd | &cmpFsaThrd, D, N, n, len, seqs,
 *                     include, includes, norm, minLength, minCov, ...)
 *                                  reference fsacmpthrd.h:49, fsacmpthrd.c:76
 *
 * called at cdist.c:181,184 (multi-file) and cdist.c:351,354 (MSA).  Every
 * entry point below cites the reference interface it replaces (paths relative
 * to the reference tree).  Plain C types only; all buffers are caller-owned.
 * There is no CPU fallback: when no sm_100 device is usable every call
 * returns an error and the caller must stop.
 *
 * Data formats at the boundary (exactly the reference's in-memory formats):
 *   seqs[i]      ceil(len/32) x u64, 32 bases per word, base p in bits
 *                63-2(p%32)..62-2(p%32), unknown packed as 00, tail
 *                left-aligned                      (qseqs.c:60 qseq2nibble)
 *   includes[i]  ceil(len/32) x u32, base p <-> bit 31-(p%32); bits >= len
 *                clear                    (fsacmp.c:164 initIncPos, :181 getIncPos)
 *   include[i]   0 = sample excluded (compacted out of D and N)
 *   D, N         packed strict lower triangle over the INCLUDED samples in
 *                input order, row-major, row r at r(r-1)/2, cells of
 *                elem_size bytes: 8 double | 4 float | 2 u16 | 1 u8
 *                                                (matrix.c:32 ltdMatrixInit)
 */
#ifndef CCPHYLO_GPU_H
#define CCPHYLO_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ccg_ctx ccg_ctx;

enum {
	CCG_OK = 0,
	CCG_ERR_NO_DEVICE = 1,   /* no CUDA device / not sm_100: there is no CPU fallback */
	CCG_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed: see ccg_last_error */
	CCG_ERR_ARG = 3,         /* invalid argument or call order */
	CCG_ERR_NOMEM = 4,       /* host or device allocation failed */
	CCG_ERR_UNSUPPORTED = 5  /* an option combination the device path does not take (e.g. the device-built global mask on a rank partition) */
};

/* kernel selection for ccg_set_kernel */
enum {
	CCG_KERNEL_AUTO = 0,
	CCG_KERNEL_POPC = 1,     /* bit-sliced LOP3+POPC path on the INT pipe */
	CCG_KERNEL_UMMA = 2,     /* contraction on tcgen05 tensor cores (CTA pairs; e2m1 operands on kind::mxf4, or int8 on kind::i8
	                          * with CCG_I8=1), operands expanded to an HBM panel */
	CCG_KERNEL_FUSED = 3     /* same contraction, operands expanded from the bit planes inside the CTA */
};

const char *ccg_strerror(int code);
/* last detailed message of this context (or of the failed ccg_init when ctx is NULL) */
const char *ccg_last_error(const ccg_ctx *ctx);

/* Create a context on CUDA device `device` (-1 = the current device).  Replaces
 * the thread fan-out set-up of fsaCmpThreadOut (fsacmpthrd.c:76-96). */
int ccg_init(ccg_ctx **ctx, int device);
void ccg_destroy(ccg_ctx *ctx);

/* Launch everything on the caller's cudaStream_t (NULL = the context's own
 * stream).  Lets a host that already owns a stream order and time the work.
 * The new stream is ordered after everything still queued on the old one. */
int ccg_set_stream(ccg_ctx *ctx, void *cuda_stream);
int ccg_set_kernel(ccg_ctx *ctx, int kernel);
/* Block until all work queued by this context has finished. */
int ccg_sync(ccg_ctx *ctx);

/* ---------------------------------------------------------------------------
 * Multi-GPU.  The reference's fan-out is one C call that spreads the pair loop over `tnum` threads
 * (fsaCmpThreadOut fsacmpthrd.c:76-106).  Here the same call spreads over the GPUs of one box by cutting
 * the ALIGNMENT axis ("K split"): member g of `world` GPUs holds the bases [b_g, b_g+1) of every sample
 * (b_g multiples of 256), runs the whole lower triangle on its slice, and the int32 partial sums are
 * added by the member that owns a matrix row -- over NVLink, fused with the epilogue.  Integer split
 * sums are exact, so the result is bit-identical to a single-GPU run.
 *
 * In one process: ccg_init_multi returns ONE handle that takes the calls below like a single-device
 * context (ccg_set_problem .. ccg_put_sample_codes / ccg_put_samples_packed / ccg_put_global_mask ..
 * ccg_run_pair / ccg_run_global, ccg_fsa_cmp_thread_out, ccg_get_inc_counts, ccg_build_global_mask).
 * A problem worth splitting (tensor path, >= 128 kbp per GPU, no -P / -y) runs on all members, one
 * host thread per device; anything else runs on member 0 alone; calls that need a whole sample on one
 * device (-P, -y, -V, -a, device-pointer uploads) return CCG_ERR_UNSUPPORTED while a problem is split.
 * ngpus <= 0: every visible device.  ccg_multi_gpus returns the member count, *active the number
 * working on the current problem.  Only member 0's device context is started by ccg_init_multi; the others start
 * (in parallel) when the first problem is split, so a small job on an 8-GPU box pays for one context, as the
 * reference's small jobs pay for no threads (fsacmpthrd.c:84-89).  ccg_multi_contexts: how many have been started. */
int ccg_init_multi(ccg_ctx **ctx, int ngpus);
int ccg_init_multi_devices(ccg_ctx **ctx, int ngpus, const int *devices);   /* devices may repeat (tests on one GPU) */
int ccg_multi_gpus(const ccg_ctx *ctx, int *active);
int ccg_multi_contexts(const ccg_ctx *ctx);

/* One process per GPU (torchrun, MPI): every rank creates its context with ccg_init, exports a handle
 * to the peer window that holds its accumulators (sized for max_samples sample slots), the ranks
 * exchange the CCG_GROUP_HANDLE_BYTES blobs (any host channel), and every rank joins with all of them
 * in rank order.  From then on the context's problem is THIS RANK'S SLICE of the alignment:
 * ccg_set_problem(n, slice length) and the put / run calls work on the slice, every rank passes the
 * same include[] and epilogue arguments, the run calls synchronise the ranks on the device, and a rank
 * writes only the cells of the matrix rows it owns (ccg_group_row_owner; host and device outputs outside
 * them are left untouched).  The int32 partial sums live in a peer-readable window of at most 8 GiB per
 * rank: a sample set whose n x n accumulators exceed it (BASELINE configs[3]: 100,000 samples) is run in
 * windows of whole macro-tile rows, each reduced and finalised before the next one starts, so the sequence
 * shards never move between the GPUs.
 * ccg_group_set_alignment: the length of the WHOLE alignment (the minCov gate, fsacmpthrd.c:292) and,
 * for shared-mask runs, getNpos of the whole global mask (fsacmpthrd.c:164-176); call it before a run.
 * Not available on such a context: -P, ccg_run_row, ccg_list_variants, ccg_get_raw_counts,
 * ccg_set_partition, ccg_set_tile_window.  A rank that never reaches the run makes the others fail
 * after 120 s (CCG_GROUP_TIMEOUT_S; CCG_ERR_CUDA): start the runs of all ranks together. */
#define CCG_GROUP_HANDLE_BYTES 128
int ccg_group_export(ccg_ctx *ctx, int max_samples, void *handle);
int ccg_group_join(ccg_ctx *ctx, int rank, int world, const void *handles);
int ccg_group_leave(ccg_ctx *ctx);
int ccg_group_set_alignment(ccg_ctx *ctx, long long total_len, unsigned global_inc);
/* compact != 0: the D / N buffers of the run calls (host or device) hold ONLY this rank's cells -- the rows of
 * its row blocks over the included samples, back to back: compact row r of an owned block starts where the
 * previous owned included row ended and has r cells.  For sample sets whose full matrices are too large to
 * exist once per rank (100,000 samples: 40 GB per matrix).  Default 0: full-size buffers, packed addressing. */
int ccg_group_set_output(ccg_ctx *ctx, int compact);
/* pure host arithmetic: rows are owned in blocks of ccg_group_row_block() = 64, row i by rank (i / 64) % world;
 * ccg_group_cells = packed cells of an n-sample triangle (all included) a rank owns */
int ccg_group_row_block(void);
int ccg_group_row_owner(int row, int world);
long long ccg_group_cells(int n, int rank, int world);

/* The older deal, kept for the sample-shard ring: this context computes only the lower-triangular
 * tile blocks dealt to `rank` of `world` (one process per GPU, no data-path collective).  Cells of
 * other ranks are left untouched in device outputs and read back as zero in host outputs.
 * Default rank 0 of 1. */
int ccg_set_partition(ccg_ctx *ctx, int rank, int world);
/* Restrict the run to the cells with row_lo <= i < row_hi and col_lo <= j < col_hi (sample slot
 * indices; row_lo and col_lo must be multiples of ccg_tile_rows() / ccg_tile_cols(); the window
 * is widened to whole macro tiles).  row_lo < 0 removes the window.  This is how the
 * sample-shard ring (sets larger than one GPU's HBM) computes one shard-by-shard block per
 * step: the resident shard in slots [0, S), the visiting one in [S, 2S), window rows [S, 2S) x
 * columns [0, S).  Cells outside the window are left untouched. */
int ccg_set_tile_window(ccg_ctx *ctx, int row_lo, int row_hi, int col_lo, int col_hi);

/* Pure host helpers (no device needed) describing that deal: the lower
 * triangle is cut into macro tiles of ccg_tile_rows() x ccg_tile_cols()
 * samples (tm, tn <= tm), ordered along a Z-order curve that is cut into
 * `world` contiguous runs of equal estimated cost (tiles + the row blocks
 * they read); rank r owns run r.
 * ccg_partition_cells: number of (r,c) cells of an n-sample matrix owned by
 * `rank`.  ccg_partition_tiles: writes up to `cap` owned tiles to tm[] / tn[]
 * and returns how many the rank owns. */
int ccg_tile_rows(void);
int ccg_tile_cols(void);
long long ccg_partition_cells(int n, int rank, int world);
long long ccg_partition_tiles(int n, int rank, int world, int *tm, int *tn, long long cap);

/* Upper bound in bytes for the tensor-core kernel's int8 operand panel (the
 * K axis is processed in slabs that fit); 0 = default (70 % of the free device memory). */
int ccg_set_scratch_limit(ccg_ctx *ctx, size_t bytes);

/* Declare the sample set: n sample slots of len bases.  pair_mode != 0 is
 * `-f` bit 2 (per-pair inclusion, cmpairFsaThrd); 0 is the shared-mask mode
 * (cmpFsaThrd).  Allocates the device-resident sample store. */
int ccg_set_problem(ccg_ctx *ctx, int n, int len, int pair_mode);

/* Shared-mask mode only: the global mask includes[0] (cdist.c:101-112), must
 * be set before the samples are put.  Host pointer. */
int ccg_put_global_mask(ccg_ctx *ctx, const uint32_t *mask);

/* Shared-mask mode for a store declared with pair_mode != 0 and already filled: ANDs every
 * sample with the global mask (host pointer, ceil(len/32) words).  For callers that stream
 * samples to the device before the global mask is complete (cdist.c:86-112 builds it while
 * loading).  Afterwards ccg_run_global[_dev] may be used on this store. */
int ccg_apply_global_mask(ccg_ctx *ctx, const uint32_t *mask);

/* Same, with the global mask built on the device: the AND of the inclusion masks of all
 * uploaded samples with include[i] != 0 (NULL = all), which is what cdist.c:86-112 accumulates
 * in includes[0] while loading.  *global_inc receives its getNpos (fsacmpthrd.c:164). */
int ccg_build_global_mask(ccg_ctx *ctx, const unsigned char *include, unsigned *global_inc);

/* Upload samples [first, first+count) in the reference's packed format from
 * HOST memory; seqs[k] / includes[k] are row pointers exactly as the
 * reference holds them (dist.c:143-154).  includes may be NULL in
 * shared-mask mode.  A NULL row pointer leaves that slot empty (excluded).
 * Source buffers of ALL host-pointer ccg_put_* calls: pageable rows are staged
 * before the call returns; rows in pinned memory (ccg_host_alloc) are read
 * asynchronously and must stay untouched until ccg_sync or a run call returns. */
int ccg_put_samples_packed(ccg_ctx *ctx, int first, int count,
                           const uint64_t *const *seqs,
                           const uint32_t *const *includes);

/* Same, from DEVICE memory already holding count x wstride words
 * (row-major).  d_masks may be NULL in shared-mask mode. */
int ccg_put_samples_packed_dev(ccg_ctx *ctx, int first, int count,
                               const uint64_t *d_seqs, const uint32_t *d_masks,
                               long wstride);

/* Same, but the rows are LENT instead of copied: they must stay valid and unchanged until the run that uses them
 * has finished (ccg_sync or a host-output run call), and the library reads them during that run.  The tensor path
 * then expands its operands straight from the packed words and skips building the bit-plane store (12 B read +
 * 12 B written per 32 bases and sample); any call that needs the planes (-P, -y, -V, the POPC kernel, the
 * per-sample counts) builds them from the lent rows first.  One lent range at a time. */
int ccg_put_samples_packed_dev_borrowed(ccg_ctx *ctx, int first, int count,
                                        const uint64_t *d_seqs, const uint32_t *d_masks,
                                        long wstride);

/* Upload one sample as translated codes 0..4 (len bytes, host); the device
 * performs qseq2nibble (qseqs.c:60), initIncPos + getIncPos(seq, seq, 0)
 * (fsacmp.c:164,181) and getNpos (:487).  Pair mode only. */
int ccg_put_sample_codes(ccg_ctx *ctx, int idx, const unsigned char *codes);

/* -P proxi (dist.c:712, "minimum proximity between SNPs"): set before the samples are put.
 * snp_events_only != 0 selects the event definition of getIncPosInsig / getIncPosInsigPrune
 * (-f 8 / -f 32, dist.c:802-806: only positions where both sequences are known and differ),
 * 0 that of getIncPos (fsacmp.c:181: differing or unknown positions).  With proxi > 0
 *   - ccg_run_pair[_dev] performs maskProxi per pair (fsacmp.c:355-485, fsacmpthrd.c:410)
 *     before counting, bug-compatible with the reference;
 *   - ccg_build_global_mask additionally clears, for every included sample, the runs
 *     between close events against the first included sample (cdist.c:111,138);
 *   - ccg_sample_proximity applies the per-sample builder (cdist.c:91).
 * proxi = 0 (the default) switches all of it off. */
int ccg_set_proximity(ccg_ctx *ctx, unsigned proxi, int snp_events_only);

/* Pair-mode store: getIncPosPtr(includes[i], seq, seq, proxi) (cdist.c:91,138) for the
 * uploaded slots [first, first+count): positions between two unknown positions at most
 * proxi apart are removed from the sample's own mask.  apply == 0 only counts.
 * inc_out[k] (may be NULL) receives getNpos of slot first+k's mask after the masking. */
int ccg_sample_proximity(ccg_ctx *ctx, int first, int count, int apply, unsigned *inc_out);

/* Shared-mask mode with -y AND -P: the reference narrows the shared mask by the motif sites (maskMotifs cdist.c:109,137)
 * and by the proximity runs, whose events are defined on the sequences (getIncPosPtr cdist.c:111,138).  The motif sites
 * must therefore not be in the samples' own masks when ccg_build_global_mask runs its proximity pass: with motifs set
 * and no ccg_mask_motifs call on the problem, ccg_build_global_mask applies the motif masking itself, after that pass.
 * ccg_sample_count_masked returns what the reference's inclusion test sees for a REFERENCE candidate (cdist.c:137-140:
 * getNpos after maskMotifs and getIncPosPtr(includes, seq, seq, proxi)) and leaves the store as it was. */
int ccg_sample_count_masked(ccg_ctx *ctx, int slot, unsigned *inc_out);

/* -y / --methylation_motifs: the motif list getMethMotifs (methparse.c:268-296) builds -- every
 * motif of the file followed by its reverse complement (as strrcMeth :83-103 produces it).  Motif
 * m has lens[m] (1 .. 32) positions whose codes follow each other in `sets`: bits 0..3 = the
 * bases A, C, G, T the position's IUPAC letter accepts, bit 4 = methylation site (an upper-case
 * letter).  Host pointers, copied.  nmotifs = 0 removes the list. */
int ccg_set_motifs(ccg_ctx *ctx, int nmotifs, const int *lens, const unsigned char *sets);

/* maskMotifs (meth.c:141-159, called cdist.c:90,109,137) for the uploaded slots
 * [first, first+count) of a pair-mode store, once, right after their upload: wherever a motif
 * matches the sample's packed sequence (unknown bases read as A) the methylation sites leave the
 * sample's mask.  inc_out[k] (may be NULL) receives getNpos of slot first+k afterwards. */
int ccg_mask_motifs(ccg_ctx *ctx, int first, int count, unsigned *inc_out);

/* Per-slot included-position counts (getNpos of each sample's own mask,
 * fsacmp.c:487; cdist.c:91).  out has n entries. */
int ccg_get_inc_counts(ccg_ctx *ctx, unsigned *out);

/* Pair mode: replaces fsaCmpThreadOut(tnum, &cmpairFsaThrd, ...)
 * (cdist.c:181, fsacmpthrd.c:261-480).  include has n flags (NULL = all
 * uploaded slots).  minLength is re-maxed with minCov*len as fsacmpthrd.c:292
 * does.  D and N (N may be NULL) are HOST buffers of Dn(Dn-1)/2 cells of
 * elem_size bytes; byteScale is the reference's global ByteScale
 * (bytescale.c:22) used for elem_size 2 and 1.  *Dn receives the number of
 * included samples (D->n). */
int ccg_run_pair(ccg_ctx *ctx, const unsigned char *include, unsigned norm,
                 unsigned minLength, double minCov, int elem_size,
                 double byteScale, void *D, void *N, int *Dn);

/* Shared-mask mode: replaces fsaCmpThreadOut(tnum, &cmpFsaThrd, ...)
 * (cdist.c:184, fsacmpthrd.c:108-259).  *global_inc receives getNpos of the
 * global mask (the caller prints the "# inc / len bases included" line,
 * fsacmpthrd.c:165).  Only included samples are compared (the reference's
 * own pair selection is wrong when a sample is excluded, fsacmpthrd.c:194). */
int ccg_run_global(ccg_ctx *ctx, const unsigned char *include, unsigned norm,
                   int elem_size, double byteScale, void *D, int *Dn,
                   unsigned *global_inc);

/* Device-resident variants: D / N are DEVICE buffers, nothing is copied to
 * the host and the call only enqueues work on the context's stream. */
int ccg_run_pair_dev(ccg_ctx *ctx, const unsigned char *include, unsigned norm,
                     unsigned minLength, double minCov, int elem_size,
                     double byteScale, void *d_D, void *d_N, int *Dn);
int ccg_run_global_dev(ccg_ctx *ctx, const unsigned char *include,
                       unsigned norm, int elem_size, double byteScale,
                       void *d_D, int *Dn, unsigned *global_inc);

/* One row against an existing matrix (-a): replaces fsaCmpThreadOut(tnum, &cmpFsaRowThrd, ...)
 * (fsacmpthrd.c:653, worker :482-580).  The new sample sits in slot row_slot of a pair-mode
 * store, the samples of the existing matrix in the slots below it.  D[k] / N[k] (HOST, doubles as
 * in the reference; N may be NULL) receive the cell against the k-th uploaded slot below
 * row_slot: N = positions known in both, D = norm ? mismatches * norm / N : mismatches, and
 * D = -1, N = 0 where N < max(minLength, minCov * len) (the caller prints the reference's
 * "No sufficient overlap with sample" line for those).  *cols receives the number of cells.
 * With ccg_set_proximity(proxi > 0) the pair's mask is what cmpFsaRowThrd builds (:545-546): the
 * new sample's own mask after its builder, put through getIncPosPtr(…, seq, ref, proxi) against the
 * column sample; the new sample is uploaded as it is (no ccg_sample_proximity on its slot). */
int ccg_run_row(ccg_ctx *ctx, int row_slot, unsigned norm, unsigned minLength, double minCov,
                double *D, double *N, int *cols);

/* -V / --nucleotide_variations: what fsaCmpThreadOut prints into `diffile` while it compares --
 * fsacmpairint (fsacmp.c:685-737, pair != 0) or fsacmprint (:646-683, shared mask), one
 * "(i, j)\t<base_i><label><base_j>" line per variant (printDiff :635-644).  fn is called once
 * per compared pair that has variants, rows ascending and columns ascending within a row (the
 * order of the reference run with -t 1): sample_i > sample_j are slot numbers,
 * variants[k] = label << 4 | code_i << 2 | code_j (codes 0..3 = ACGT; label is the reference's
 * running position counter, see csrc/k_variants.cu).  A non-zero return of fn stops the listing.
 * Pair mode: a pair-mode store as ccg_run_pair takes it.  Shared-mask mode: after
 * ccg_build_global_mask and BEFORE ccg_run_global.  Not with -P. */
typedef int (*ccg_variant_fn)(void *user, int sample_i, int sample_j, const uint64_t *variants, size_t count);
int ccg_list_variants(ccg_ctx *ctx, int pair, const unsigned char *include, ccg_variant_fn fn, void *user);
/* The same for the row of ccg_run_row (-V with -a: fsacmpairint(diffile, n, j, ...) in cmpFsaRowThrd,
 * fsacmpthrd.c:552-553): the pairs (row_slot, j) of the uploaded slots j below row_slot. */
int ccg_list_variants_row(ccg_ctx *ctx, int row_slot, ccg_variant_fn fn, void *user);

/* Raw integer results of the last pair-mode run for included samples:
 * mismatch counts and inclusion counts as u32, same packed layout.  HOST
 * buffers; either may be NULL. */
int ccg_get_raw_counts(ccg_ctx *ctx, uint32_t *mism, uint32_t *ninc);

/* One-call drop-in with the argument list of fsaCmpThreadOut
 * (fsacmpthrd.h:49): pair != 0 selects cmpairFsaThrd, else cmpFsaThrd.
 * Host row pointers in, host matrices out.  proxi > 0: the
 * caller's masks already went through getIncPos (cdist.c:91), the call performs
 * maskProxi per pair (pair mode; cmpFsaThrd ignores proxi).  ctx may be NULL (a temporary context on
 * the current device is used). */
int ccg_fsa_cmp_thread_out(ccg_ctx *ctx, int pair, void *D, void *N,
                           int elem_size, double byteScale, int n, int len,
                           const uint64_t *const *seqs,
                           const unsigned char *include,
                           const uint32_t *const *includes, unsigned norm,
                           unsigned minLength, double minCov, unsigned proxi,
                           int *Dn, unsigned *global_inc);

/* Pinned host memory for sample rows / result matrices (faster H2D / D2H). */
void *ccg_host_alloc(size_t bytes);
void ccg_host_free(void *p);

/* Introspection for tests and the bench: kernel launches issued by this
 * context so far, and the name of the compare kernel used by the last run. */
long long ccg_launch_count(const ccg_ctx *ctx);
const char *ccg_last_kernel(const ccg_ctx *ctx);
/* device time (ms, CUDA events on the context's stream) of the compare kernel
 * of the last run; < 0 if unavailable */
float ccg_last_compare_ms(ccg_ctx *ctx);
/* tensor-core path only: device time (ms) of the first K slab's operand
 * expansion (phase 0) or int8 GEMM launch (phase 1); < 0 if unavailable */
float ccg_last_phase_ms(ccg_ctx *ctx, int phase);

/* ---------------------------------------------------------------------------
 * KMA count-matrix (.mat) inputs: replaces ltdMatrixThrd (ltdmatrixthrd.c:376, called at
 * dist.c:168) / ltdMatrix_get (ltdmatrix.c:32, dist.c:264) with cmpMats (matcmp.c:448) and the
 * -d method table of dist.c:738-786.  Every sample's template is uploaded ONCE (the reference
 * re-reads sample j's file for every cell (i, j)).
 * ------------------------------------------------------------------------- */
enum {
	CCG_MAT_COS = 0, CCG_MAT_Z, CCG_MAT_CHI2, CCG_MAT_NCHI2, CCG_MAT_C, CCG_MAT_NC, CCG_MAT_P, CCG_MAT_NP,
	CCG_MAT_BC, CCG_MAT_NBC, CCG_MAT_L1, CCG_MAT_L2, CCG_MAT_LINF, CCG_MAT_LN, CCG_MAT_NL1, CCG_MAT_NL2,
	CCG_MAT_NLINF, CCG_MAT_NLN, CCG_MAT_METHODS
};

/* n sample slots of at most max_len positions (rows of the template whose reference base is
 * not '-': insertion rows are dropped by the caller, as stripMat matcmp.c:27 intends). */
int ccg_mat_set_problem(ccg_ctx *ctx, int n, int max_len);
/* counts6: len x 6 u16 in the reference's order A, C, G, T, -, N (matparse.c:254-259);
 * totals: len x u32 row totals, or NULL to sum the six counts.  Host pointers.  The device keeps the 12 bytes of
 * counts per position; the totals of a sample whose totals are NOT the sum of its (16-bit) counts -- a depth above
 * 65,535, which the reference truncates in the count but not in the total -- go to a side plane. */
int ccg_mat_put_sample(ccg_ctx *ctx, int idx, const uint16_t *counts6, const uint32_t *totals, int len);
/* All pairs of the samples with include[i] != 0 (NULL = all slots).  method is a CCG_MAT_* id,
 * order the n of l<n> / nl<n>, alpha the -l level of `z`.  D, N (N may be NULL): HOST buffers of
 * Dn(Dn-1)/2 cells of elem_size bytes; a pair without sufficient overlap (matcmp.c:485) gets
 * D = -1, N = 0.  rows_inc (may be NULL): u32 rowsInc of every cell, 0 for such pairs, so the
 * caller can print the reference's "No sufficient overlap" lines. */
int ccg_mat_run(ccg_ctx *ctx, const unsigned char *include, int method, unsigned order, double alpha,
                unsigned norm, unsigned minDepth, unsigned minLength, double minCov, int elem_size,
                double byteScale, void *D, void *N, int *Dn, uint32_t *rows_inc);

/* Position split over several GPUs (the K split of the FASTA path; a multi-GPU context does this behind ccg_mat_run).
 * A member context holds the positions [p_g, p_g+1) of every sample (ccg_mat_set_problem / ccg_mat_put_sample with the
 * slice lengths).  ccg_mat_run_partial returns its raw per-pair sums, packed over the included samples: dist[cell] the
 * fp64 sum of the per-position distances, rows[cell] the positions that counted (rowsInc, matcmp.c:470-481).  The
 * caller adds the members' sums and finishes with ccg_mat_finalize_host (pure host code, no device): the tail of
 * cmpMats (matcmp.c:483-494) and the cell formats, lens[] = the whole length of every sample slot. */
int ccg_mat_run_partial(ccg_ctx *ctx, const unsigned char *include, int method, unsigned order, double alpha,
                        unsigned minDepth, double *dist, uint32_t *rows, int *Dn);
int ccg_mat_finalize_host(int n, const unsigned char *include, const int *lens, const double *dist, const uint32_t *rows,
                          unsigned norm, unsigned minLength, double minCov, int elem_size, double byteScale, void *D, void *N,
                          uint32_t *rows_inc, int *Dn);

/* One row against an existing matrix (-a on .mat input): replaces matCmpThreadOut(tnum,
 * &cmpMatRowThrd, ...) (ltdmatrixthrd.c:605, worker :111-181).  The new sample sits in slot
 * row_slot, the samples of the existing matrix in the slots below.  D, N (may be NULL), rows_inc
 * (may be NULL): HOST buffers of row_slot doubles / u32, same cell rules as ccg_mat_run. */
int ccg_mat_run_row(ccg_ctx *ctx, int row_slot, int method, unsigned order, double alpha, unsigned norm,
                    unsigned minDepth, unsigned minLength, double minCov, double *D, double *N,
                    uint32_t *rows_inc);

/* ---------------------------------------------------------------------------------------------
 * `ccphylo trim` (fsaTrim, trim.c:77-260): the inclusion masks of the trimmed alignment.  trim has no pairwise stage;
 * what it computes per sample is the mask work the `dist` front end also does, on trim's own alphabet: the translated
 * code bytes of getIupacBitTable (fsacmp.c:93-162; 0-3 bases, 4 unknown, 5 gap, 6-15 ambiguity letters, +16 = soft-masked
 * input) or of get2BitTable (flag 4).  One job at a time per context:
 *
 *   ccg_set_motifs(ctx, ...)                      optional, before ccg_trim_begin (-y)
 *   ccg_trim_begin(ctx, len, proxi)               buffers for samples of len positions, -P proxi
 *   ccg_trim_sample(ctx, codes, nibbles, 0, 0, &inc)
 *        a sample on its own mask: initIncPos + maskMotifs + getIncPos(includes, seq, seq, proxi) + getNpos
 *        (trim.c:197-203; the pairwise flag sends every sample this way).  codes: len translated bytes (host);
 *        nibbles: the reference's packed words of the sample (qseq2nibble qseqs.c:60, ceil(len/32) u64), needed only
 *        when motifs are set; *inc receives getNpos of the mask.
 *   ccg_trim_keep_reference(ctx)                  the sample just processed becomes `ref` (trim.c:213-216): its stored
 *                                                 bytes (soft flags stripped) are what later samples are compared with,
 *                                                 its mask is the shared mask from here on
 *   ccg_trim_sample(ctx, codes, nibbles, 1, builder, 0)
 *        a later sample narrows the shared mask: maskMotifs + getIncPosPtr(includes, seq, ref, proxi) (trim.c:176-177);
 *        builder 0 = getIncPos, 1 = getIncPosInsig (flag 8), 2 = getIncPosInsigPrune (flag 32).  Also notes the
 *        columns where the sample's stored bytes differ from the reference sample's (pseudoAlnPrune fsacmp.c:504).
 *   ccg_trim_get_mask(ctx, variable_columns_only, mask, &inc, &var)
 *        the mask (ceil(len/32) words, bit 31-(p%32) of word p/32 = position p), *inc = getNpos of it, *var = the
 *        number of its positions that also vary between samples; with variable_columns_only the returned mask is
 *        restricted to those (flag 16).
 *   ccg_trim_end(ctx)
 * --------------------------------------------------------------------------------------------- */
int ccg_trim_begin(ccg_ctx *ctx, int len, unsigned proxi);
int ccg_trim_sample(ccg_ctx *ctx, const unsigned char *codes, const uint64_t *nibbles, int against_ref, int builder,
                    unsigned *inc_out);
int ccg_trim_keep_reference(ccg_ctx *ctx);
int ccg_trim_get_mask(ccg_ctx *ctx, int variable_columns_only, uint32_t *mask_out, unsigned *inc_out, unsigned *var_out);
int ccg_trim_end(ccg_ctx *ctx);

/* Roofline denominator for the tensor-core kernel: runs a loads-free loop of the kernel's own
 * MMA shape (tcgen05 kind::i8, cta_group::2, 256 x 256 x 32, operands static in shared memory)
 * on every CTA pair for about target_ms and returns the rate in int8 TOP/s (2 ops per MAC);
 * < 0 on failure.  A few ms measures the burst rate, a second the power-capped sustained one. */
double ccg_measure_i8_peak(ccg_ctx *ctx, double target_ms);
/* The same for the default operand format: e2m1 on tcgen05 kind::mxf4 (block scales 1.0, f32
 * accumulators).  Also accumulates +1 / -1 products up to about check_sum (< 2^24) and reports in
 * *inexact how many accumulator elements then differed from the exact integer (0 is what the
 * tensor path relies on); info, if not NULL, receives {dot of the test pattern, D[0][0], iterations}. */
double ccg_measure_fp4_peak(ccg_ctx *ctx, double target_ms, double check_sum, long long *inexact, int *info);

#ifdef __cplusplus
}
#endif
#endif
