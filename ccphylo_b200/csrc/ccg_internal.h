/*
 * ccg_internal.h -- shared declarations of the CUDA side (not part of the ABI).
 *
 * Device-resident sample store ("planes"), one per context:
 *
 *     planes[chunk][plane][slot][4]   u32
 *
 *   chunk  = 128 consecutive bases (4 words of 32 bases)
 *   plane  = 0: high bit of the 2-bit code, 1: low bit, 2: inclusion mask
 *            (pair mode: 3 planes; shared-mask mode: 2 planes, the global mask
 *             is folded into the code planes at encode time)
 *   slot   = sample index, padded to a multiple of CCG_SLOT_PAD
 *   bit 31-(p%32) of word (p/32)%4 <-> base p (the reference's mask bit order,
 *   fsacmp.c:164).  Code bits of masked positions are cleared, so a plane word
 *   of an absent / padded slot is all-zero and contributes nothing.
 *
 * One (chunk, plane) row of 64 slots is 1 KiB contiguous, every thread access
 * is a 128-bit vector, and a [KC chunks][planes][64 slots][4] box is a single
 * TMA tile.
 *
 * Work partition (all compare kernels): the lower triangle is cut into macro
 * tiles of CCG_UMMA_BM rows x CCG_UMMA_BN columns (tm, tn <= tm), ordered along
 * a Z-order curve that is cut into `world` contiguous cost-balanced runs.
 */
#ifndef CCG_INTERNAL_H
#define CCG_INTERNAL_H

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ccphylo_gpu.h"

#define CCG_TILE 64          /* samples per tile edge (popc kernel) */
#define CCG_CHUNK_WORDS 4    /* u32 words per plane per chunk */
#define CCG_CHUNK_BASES 128
#define CCG_UMMA_BM 256      /* macro tile rows    (tcgen05 M = 256 over a CTA pair: 2 x 128 TMEM lanes) */
#define CCG_UMMA_BN 256      /* macro tile columns (tcgen05 N) */
#define CCG_SLOT_PAD 256     /* slots are padded to a multiple of this */
#define CCG_FP4_MAX_PAIRS 20000   /* chunk pairs per kind::mxf4 work item: 3 * 256 * 20000 < 2^24, every partial sum exact in f32 */

struct EpilogueParams {
	int mode;              /* 0 pair (fsacmpthrd.c:419-475), 1 global (fsacmpthrd.c:247-255) */
	int elem_size;         /* 8, 4, 2, 1 */
	unsigned norm;
	unsigned minLength;    /* final gate (already maxed with minCov*len) */
	double byteScale;
	double nFactor;        /* global mode: norm ? norm/inc : 1.0 */
	void *D;               /* device, packed over included samples */
	void *N;               /* device or NULL */
	const int *rank;       /* slot -> compact index, -1 = excluded */
	const long long *row_base;   /* NULL: cell = r (r - 1) / 2 + c.  Else cell = row_base[r] + c (a K-split member's compact
	                              * buffer of the rows it owns) */
	int row_plus1;         /* 0 = the packed triangle.  r + 1: only the cells of compact row r are written, cell = column, and
	                        * a cell that fails the gate gets N = 0 (cmpFsaRowThrd fsacmpthrd.c:560-570; doubles only) */
};

struct PopcParams {
	int chunks;            /* ceil(words/4) */
	int ksplit;            /* K slices per tile */
	int chunks_per_split;  /* multiple of KC */
	int ntiles;            /* 64x64 tiles owned by this rank */
	const int2 *tiles;     /* device: (ti, tj) */
	uint32_t *acc;         /* [ntiles][2][TILE*TILE] raw mism / ninc */
	unsigned *tickets;     /* [ntiles] */
	EpilogueParams ep;
};

struct ProxiParams {
	int ntiles;            /* 64x64 tiles owned by this rank */
	const int2 *tiles;     /* device: (ti, tj) */
	uint32_t *acc;         /* [ntiles][2][TILE*TILE] mismatches / included under the proximity mask */
	EpilogueParams ep;
};

struct VariantParams {
	int ncells;                         /* cells of this batch */
	long long cell0;                    /* first packed cell (over included samples) */
	const int *slot_of_rank;            /* device: compact index -> slot */
	unsigned *counts;                   /* count pass: variants per cell */
	const unsigned long long *offsets;  /* write pass: running offsets of the batch's cells (offsets[0] = base) */
	unsigned long long *entries;        /* write pass: label << 4 | code_i << 2 | code_j */
	uint32_t *pair_mask;                /* -V with -P: the batch's per-pair masks, [word][cell of the batch] (0 = none) */
	long long pair_mask_stride;         /* cells per word row of pair_mask */
	const uint4 *row_raw;               /* -V with -a and -P: the row sample's planes as uploaded, [chunk][h, l, m] (0 = the store's) */
};

struct UmmaParams {
	int ntiles;            /* macro tiles owned by this rank */
	const int2 *tiles;     /* device: (tm, tn) */
	int kslices;           /* K slices per tile within the slab */
	int chunks_per_slice;
	int slab_chunks;       /* chunks resident in X */
	int *C_S, *C_I;        /* dense [n_pad][ldc] int32 accumulators */
	int ldc;
	int row_base;          /* first 128-byte row of this slab's buffer inside the panel allocation */
	unsigned *sync;        /* lock-step epoch counters [rounds][epochs_per_item] (zeroed), or NULL */
	int epochs_per_item;
	unsigned *resident;    /* every CTA adds 1 once it holds its SM (gates the overlapped expansion), or NULL */
	int no_mask_items;     /* e2m1 panel, shared-mask mode: the inclusion count is a constant, no I items */
	int fp4;               /* e2m1 panel (kind::mxf4): slab_chunks / chunks_per_slice count chunk PAIRS, S and I are separate items */
	int single;            /* tiles are 128 x 256 and run on the single-CTA kernel (CCG_UMMA1=1, experiments) */
	int thin_tm, thin_n, thin_valid;   /* thin items (k_pairdist_umma2): tile row, MMA N, valid rows; thin_tm < 0 = none */
	long long watchdog;    /* cycles an mbarrier wait may take before the kernel traps; 0 = no limit (set by ccg_launch_umma) */
};

/* ---- K-split group: every member GPU owns a slice of the alignment (all samples, 1/world of the bases), runs
 * the whole lower triangle on it, and the members' int32 partial sums are added up by the OWNER of a matrix row
 * through peer-mapped pointers (NVLink), fused with the epilogue -- ccg_group.cu ---- */
#define CCG_GROUP_MAX 16
#define CCG_GROUP_HDR_BYTES 4096
#define CCG_GROUP_ROW_BLOCK 64              /* matrix rows are owned in blocks of this many, dealt round-robin */

struct GroupHeader {                        /* start of every member's peer window */
	unsigned arrive[CCG_GROUP_MAX];         /* arrive[p]: last barrier epoch member p announced to this member */
	int iconst_from[2][CCG_GROUP_MAX];      /* shared-mask mode: member p's constant inclusion count, per C buffer */
};

struct GroupBarrierParams {
	GroupHeader *hdr[CCG_GROUP_MAX];
	int rank, world, buf, i_const;
	unsigned epoch;
	int spin;                               /* 0: publish only (members on one device synchronise on the host) */
	unsigned long long timeout_ns;
};

struct GroupFinalizeParams {
	const int *C[CCG_GROUP_MAX];            /* every member's current accumulator buffer: S plane, I plane behind it */
	const GroupHeader *own;
	size_t plane;                           /* ints per plane (window rows * ldc) */
	int world, rank, ldc, n, pair_mode, buf;
	int row_lo, row_hi;                     /* rows of this window; the member finalises those it owns */
};

struct ccg_multi;

struct ccg_ctx {
	int device;
	int sm_count;
	cudaStream_t own_stream, stream;
	cudaStream_t aux_stream;   /* operand expansion of slab s+1 runs here, under the GEMM of slab s */
	cudaEvent_t ev_fork, ev_launch, ev_x[2], ev_g[2];
	cudaEvent_t ev_switch;     /* ccg_set_stream: orders the new stream after the work queued on the old one */
	int kernel_choice;
	int rank, world;
	int win_on, win[4];            /* macro-tile window [tm_lo, tm_hi) x [tn_lo, tn_hi), ccg_set_tile_window */
	int min_slabs;                  /* cut the K axis into at least this many slabs (expansion overlapped with the GEMM) */
	int use_i8;                     /* CCG_I8=1: int8 operands (kind::i8) instead of the default e2m1 panel (kind::mxf4) */
	long long watchdog_cycles;      /* CCG_WATCHDOG_S (default 2 s at 2 GHz; 0 = off, for profiler replays) */
	int dbg_kslices, dbg_serial, dbg_nolock, dbg_umma1;   /* CCG_KSLICES / CCG_EXPAND_SERIAL / CCG_NOLOCK / CCG_UMMA1 overrides (experiments only) */

	int motif_n, motif_nsets;       /* -y: motifs (with their reverse complements) set by ccg_set_motifs */
	int *d_motif_lens;
	unsigned char *d_motif_sets;
	int codes_upload_masked;        /* a packed upload ANDed the code planes with the rows' masks (no -P set at the time): a
	                                 * -V listing under -P would not see the words the reference compares */
	int motif_applied;              /* ccg_mask_motifs has run on some slot of this problem */
	int remask_pending;             /* -y: mask planes changed by ccg_mask_motifs, code planes not yet re-masked (done before a run) */
	int row_slot1;                  /* ccg_run_row: 1 + the slot whose row is being computed, 0 otherwise */
	unsigned proxi;                 /* -P: minimum distance between SNPs (0 = no proximity masking), ccg_set_proximity */
	int proxi_snp_only;             /* events of the per-sample builder: 0 getIncPos, 1 getIncPosInsig / getIncPosInsigPrune */

	int n, len, pair_mode;
	int words, chunks, n_pad, nplanes;
	uint32_t *d_planes;
	size_t planes_bytes;
	uint32_t *d_gmask;         /* shared-mask mode: [words] */
	unsigned global_inc;
	int global_applied;        /* pair-mode store with a global mask (ccg_apply_global_mask / ccg_build_global_mask) */
	int global_pending;        /* ... built on the device but not yet ANDed into the planes (done by the first shared-mask run;
	                            * ccg_list_variants needs the unmasked code planes) */
	unsigned *d_inc;           /* [n_pad] per-slot included counts */
	unsigned char *present;    /* host [n_pad]: slot holds a sample of the current problem */
	unsigned char *need;       /* host [n_pad/128]: row block touched by a macro tile this rank owns */
	unsigned char *have;       /* host [n_pad/128]: row block whose present slots were really uploaded */

	/* rows lent by the caller (ccg_put_samples_packed_dev_borrowed): read straight by the tensor path's expansion;
	 * the plane store of those slots is built only when something else needs it (materialize_borrowed) */
	const uint64_t *bor_seqs;
	const uint32_t *bor_masks;
	long bor_wstride;
	int bor_first, bor_count, bor_pending;
	int planes_stale;          /* the last run streamed host rows straight into the operand panel: the plane store was not built */
	int *d_unfed;              /* streamed runs: slots of read row blocks that no batch writes (their panel rows are zeroed) */
	int n_unfed;
	size_t unfed_cap;

	void *d_stage;             /* staging for host rows */
	size_t stage_bytes;

	/* K-slab streaming of host rows (ccg_fsa_cmp_thread_out on the tensor path): the rows of slab
	 * s+1 cross PCIe on copy_stream while the GEMM of slab s runs */
	int stream_min_chunks;                 /* stream when the alignment has at least this many chunks (0 = never) */
	int dbg_feed_slabs;                    /* CCG_FEED_SLABS override (experiments) */
	int dbg_feed_planes;                   /* CCG_FEED_PLANES=1: streamed rows go through the plane store (the older path) */
	cudaStream_t copy_stream[2];
	cudaEvent_t ev_up[2], ev_main;
	const uint64_t *const *feed_seqs;      /* host row pointers, or NULL when nothing is being streamed */
	const uint32_t *const *feed_masks;     /* NULL in shared-mask mode */
	void *d_stage2[2];                     /* the streaming path's own staging buffers, one per copy stream */
	size_t stage2_bytes[2];

	int *d_rank;               /* [n_pad] */
	int *h_rank;               /* host mirror of the last run */
	int last_Dn;
	int2 *d_tiles;             /* tile list of the last run */
	size_t tiles_cap;
	int last_ntiles;
	int last_i_const;          /* shared-mask mode: the constant inclusion count of the last tensor-path run */
	int last_kernel_kind;      /* CCG_KERNEL_POPC / CCG_KERNEL_UMMA of the last run */
	uint32_t *d_acc;           /* popc: tile-major raw counts */
	size_t acc_bytes;
	unsigned *d_tickets;
	size_t tickets_count;
	void *d_out_D, *d_out_N;   /* device result buffers for the host-output API */
	size_t out_bytes;

	/* tensor-core path */
	int8_t *d_X;               /* operand panel: 2 slab buffers of [n_pad/128][x_chunks*4][128][128 B] */
	size_t x_bytes;            /* whole allocation */
	size_t x_buf_bytes;        /* one slab buffer */
	int x_chunks;              /* chunks per slab */
	int *d_C;                  /* 2 x [n_pad][n_pad] int32 */
	size_t c_bytes;
	size_t x_budget;           /* max bytes for d_X (0 = default) */
	int last_gemm_ctas;        /* CTAs of the last GEMM launch */
	unsigned *d_resident;      /* running count of GEMM CTAs that became resident during this run */
	void *fn_wait_value;       /* cuStreamWaitValue32, or NULL when stream memory operations are unavailable */
	int max_pairs;             /* CTA pairs of k_pairdist_umma2 that can be resident at once (0 = not queried) */
	unsigned *d_sync;          /* lock-step epoch counters of the persistent GEMM */
	size_t sync_cap;

	CUtensorMap tmap;          /* planes as a 4-D tensor, box [4 chunks][planes][64 slots][4] (POPC kernel) */
	CUtensorMap tmap_pl;       /* same tensor, box [1 chunk][planes][128 slots][4] (fused tensor kernel) */
	CUtensorMap tmap_x;        /* operand panel as a 2-D tensor */
	CUtensorMap tmap_thin;     /* same tensor, box of tmap_thin_rows rows: a CTA's half of the B operand of a thin item */
	int tmap_thin_rows, tmap_thin_valid, dbg_nothin;
	int tmap_valid;

	/* count-matrix (.mat) path, k_matdist.cu */
	void *mat_counts;          /* [mat_npad][3][mat_lpad] u32: A|C<<16, G|T<<16, -|N<<16 */
	void *mat_tot_over;        /* [mat_npad][mat_lpad] u32 row totals of samples with counts above 65,535, else NULL */
	int *mat_lens, *mat_hlens; /* device / host: rows per slot */
	int *mat_rank;
	int mat_n, mat_npad;
	long long mat_lpad;
	void *mat_stage;           /* device staging for one sample in the upload format (len x 6 u16) */
	double *mat_part_dist;
	unsigned *mat_part_rows, *mat_rows;
	size_t mat_part_cap;
	void *mat_out_D, *mat_out_N, *mat_tiles;   /* device results and tile list of a run, kept and grown between runs */
	size_t mat_cells_cap, mat_tiles_cap;

	/* `trim` (k_trim.cu): translated code bytes of the current and of the reference sample, the (shared or own) mask,
	 * the event words of the proximity pass, the columns where some sample differs from the first one */
	unsigned char *trim_cur, *trim_ref;
	uint32_t *trim_mask, *trim_events, *trim_columns, *trim_ones;
	unsigned *trim_count;
	int trim_len, trim_words, trim_has_ref;     /* trim_len < 0: no trim job */
	unsigned trim_proxi;

	/* K-split group membership (grp_world > 1), ccg_group.cu */
	int grp_world, grp_rank;
	void *grp_win[CCG_GROUP_MAX];          /* peer windows: header + 2 accumulator buffers; [grp_rank] is this member's own */
	unsigned char grp_opened[CCG_GROUP_MAX];   /* mapped with cudaIpcOpenMemHandle (closed on leave) */
	void *grp_own_win;                     /* this context's exported window (cudaMalloc) */
	size_t grp_win_bytes;
	int grp_npad_max;                      /* slots the window's accumulators were sized for */
	unsigned grp_epoch;                    /* barriers passed so far */
	int grp_buf;                           /* accumulator buffer of the next run (the two are used in turn) */
	long long grp_total_len;               /* length of the whole alignment: the minCov gate (fsacmpthrd.c:292) */
	unsigned grp_global_inc;               /* shared-mask mode: getNpos of the whole global mask */
	size_t grp_acc_bytes;                  /* accumulator bytes of a window (the smallest of the members' exports) */
	long long *d_row_base, *h_row_base;    /* group runs into host matrices: compact offset of every owned row */
	size_t row_base_cap;
	int grp_compact;                       /* ccg_group_set_output: D / N of the run calls hold only this member's rows */
	const long long *ep_row_base;          /* what the next run's epilogue uses as EpilogueParams::row_base */
	struct ccg_multi *multi;               /* leader of an in-process multi-GPU context (ccg_init_multi) */
	int grp_same_device;                   /* another member of this process sits on the same device: host-side barrier */
	void *grp_host_barrier;                /* members of one process: host rendezvous before the device barrier (ccg_group.cu) */

	cudaEvent_t ev0, ev1;
	int ev_valid;
	cudaEvent_t ev_phase[4];   /* tensor-core path, first slab: expand begin/end, GEMM begin/end */
	int phase_valid;
	long long launches;
	char last_kernel[96];
	char err[512];
};

/* ---- launchers implemented in the kernel translation units ---- */

/* k_encode.cu */
cudaError_t ccg_launch_repack(ccg_ctx *ctx, int first, int count, const uint64_t *d_seqs,
                              const uint32_t *d_masks, long wstride);
cudaError_t ccg_launch_repack_range(ccg_ctx *ctx, cudaStream_t stream, int first, int count, const uint64_t *d_seqs,
                                    const uint32_t *d_masks, long wstride, int chunk0, int nch);
cudaError_t ccg_launch_repack_direct(ccg_ctx *ctx, cudaStream_t stream, int first, int count, const uint64_t *d_seqs,
                                     const uint32_t *d_masks, long wstride, int chunk0, int nch);
cudaError_t ccg_launch_encode_codes(ccg_ctx *ctx, int first, int count, const unsigned char *d_codes,
                                    long stride);
cudaError_t ccg_launch_gather_raw(ccg_ctx *ctx, uint32_t *d_mism, uint32_t *d_ninc);
cudaError_t ccg_launch_apply_global_mask(ccg_ctx *ctx);
cudaError_t ccg_launch_build_global_mask(ccg_ctx *ctx, const unsigned char *d_use, unsigned *d_inc, int and_into);

/* k_pairdist_popc.cu */
cudaError_t ccg_launch_popc(ccg_ctx *ctx, const PopcParams &p);
int ccg_popc_kc(void);

/* k_pairdist_umma.cu */
cudaError_t ccg_launch_expand(ccg_ctx *ctx, cudaStream_t stream, int8_t *X, int chunk0, int nchunks, int bounded);
cudaError_t ccg_launch_expand_fp4(ccg_ctx *ctx, cudaStream_t stream, int8_t *X, int chunk0, int npairs, int bounded);
cudaError_t ccg_launch_expand_rows(ccg_ctx *ctx, cudaStream_t stream, int8_t *X, int first, int count, const uint64_t *d_seqs,
                                   const uint32_t *d_masks, long wstride, int chunk0, int npairs);
cudaError_t ccg_launch_zero_panel_rows(ccg_ctx *ctx, cudaStream_t stream, int8_t *X, const int *d_slots, int nslots, int npairs);
cudaError_t ccg_launch_umma(ccg_ctx *ctx, const UmmaParams &p);
int ccg_umma_pair_slots(ccg_ctx *ctx);
cudaError_t ccg_make_thin_tmap(ccg_ctx *ctx, int rows);
cudaError_t ccg_launch_finalize_umma(ccg_ctx *ctx, const UmmaParams &p, const EpilogueParams &ep, int i_const);
cudaError_t ccg_launch_gather_raw_dense(ccg_ctx *ctx, int i_const, uint32_t *d_mism, uint32_t *d_ninc);

/* k_proxi.cu */
cudaError_t ccg_launch_sample_proxi(ccg_ctx *ctx, int vs_ref, int ref_slot, const unsigned char *d_use, int apply,
                                    unsigned *d_cleared);
cudaError_t ccg_launch_count_mask(ccg_ctx *ctx, unsigned *d_count);
cudaError_t ccg_launch_pair_proxi(ccg_ctx *ctx, const ProxiParams &p);
cudaError_t ccg_launch_row_planes(ccg_ctx *ctx, int slot, void *d_buf, int restore);
cudaError_t ccg_launch_row_proxi(ccg_ctx *ctx, int row_slot, const void *d_rowraw, const EpilogueParams &ep);

/* k_motif.cu */
cudaError_t ccg_launch_motif_mask(ccg_ctx *ctx, int first, int count, unsigned *d_removed);
cudaError_t ccg_launch_remask_all(ccg_ctx *ctx);

/* k_variants.cu */
cudaError_t ccg_launch_variants(ccg_ctx *ctx, const VariantParams &p, int write, int shared_mask);
cudaError_t ccg_launch_pair_proxi_mask(ccg_ctx *ctx, const VariantParams &p);
cudaError_t ccg_launch_row_proxi_mask(ccg_ctx *ctx, const VariantParams &p, int row_slot);

/* k_matdist.cu */
void ccg_mat_free(ccg_ctx *ctx);
void ccg_trim_free(ccg_ctx *ctx);

/* ccg_group.cu */
void ccg_set_err(ccg_ctx *ctx, const char *fmt, ...);
extern "C" long long ccg_group_cells(int n, int rank, int world);
void ccg_group_release(ccg_ctx *ctx);
int ccg_group_window_rows(ccg_ctx *ctx);
int ccg_group_accumulators(ccg_ctx *ctx, int row0, int **C_S, int **C_I, void **clear, size_t *bytes);
int ccg_group_finalize(ccg_ctx *ctx, const EpilogueParams &ep, int i_const, int row0, int row1);
void ccg_multi_destroy(ccg_ctx *ctx);
int ccg_multi_set_problem(ccg_ctx *lead, int n, int len, int pair_mode);
int ccg_multi_set_kernel(ccg_ctx *lead, int kernel);
int ccg_multi_sync(ccg_ctx *lead);
void ccg_multi_note_special(ccg_ctx *lead, int bit, int on);
ccg_ctx *ccg_multi_solo(ccg_ctx *lead, const char *what, int *rc);
int ccg_multi_forwarded(ccg_ctx *lead, int rc);
int ccg_multi_put_global_mask(ccg_ctx *lead, const uint32_t *mask, int apply);
int ccg_multi_build_global_mask(ccg_ctx *lead, const unsigned char *include, unsigned *global_inc);
int ccg_multi_put_samples_packed(ccg_ctx *lead, int first, int count, const uint64_t *const *seqs, const uint32_t *const *includes);
int ccg_multi_put_sample_codes(ccg_ctx *lead, int idx, const unsigned char *codes);
int ccg_multi_get_inc_counts(ccg_ctx *lead, unsigned *out);
int ccg_multi_run(ccg_ctx *lead, int pair, const unsigned char *include, unsigned norm, unsigned minLength, double minCov,
                  int elem_size, double byteScale, void *D, void *N, int *Dn, unsigned *global_inc);
int ccg_multi_fsa_cmp_thread_out(ccg_ctx *lead, int pair, void *D, void *N, int elem_size, double byteScale, int n, int len,
                                 const uint64_t *const *seqs, const unsigned char *include, const uint32_t *const *includes,
                                 unsigned norm, unsigned minLength, double minCov, unsigned proxi, int *Dn, unsigned *global_inc);
long long ccg_multi_launch_count(const ccg_ctx *lead);
const char *ccg_multi_last_kernel(const ccg_ctx *lead);
float ccg_multi_last_compare_ms(ccg_ctx *lead);
ccg_ctx *ccg_multi_member(ccg_ctx *lead, int g);
int ccg_multi_mat_set_problem(ccg_ctx *lead, int n, int max_len);
int ccg_multi_mat_put_sample(ccg_ctx *lead, int idx, const uint16_t *counts6, const uint32_t *totals, int len);
int ccg_multi_mat_run(ccg_ctx *lead, const unsigned char *include, int method, unsigned order, double alpha, unsigned norm,
                      unsigned minDepth, unsigned minLength, double minCov, int elem_size, double byteScale, void *D, void *N, int *Dn_out,
                      uint32_t *rows_inc);
ccg_ctx *ccg_multi_mat_solo(ccg_ctx *lead);

/* a call that only one device can take: forwarded to member 0 of a multi-GPU context unless the problem is split */
#define CCG_MULTI_SOLO(ctx, what, call)                      \
	if((ctx) && (ctx)->multi) {                              \
		int rc__ = 0;                                        \
		ccg_ctx *m0 = ccg_multi_solo((ctx), what, &rc__);    \
		if(!m0) return rc__;                                 \
		return ccg_multi_forwarded((ctx), (call));           \
	}

/* k_pairdist_fused.cu */
cudaError_t ccg_launch_fused(ccg_ctx *ctx, const UmmaParams &p);

#endif
