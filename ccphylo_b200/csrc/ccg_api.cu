/*
 * ccg_api.cu -- the C-ABI of include/ccphylo_gpu.h: context, device sample
 * store, uploads, and the run entry points that replace the reference's
 * fsaCmpThreadOut fan-out (fsacmpthrd.c:76) and its two workers
 * cmpairFsaThrd (:261) / cmpFsaThrd (:108).
 *
 * Host logic only; the arithmetic lives in k_encode.cu / k_pairdist_*.cu.
 * No CPU fallback: every failure is reported to the caller.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "ccg_internal.h"

static char g_init_err[512] = "";
static void update_need(ccg_ctx *ctx);
static int materialize_borrowed(ccg_ctx *ctx);
static int planes_usable(ccg_ctx *ctx);
#define NEED_PLANES(ctx)                          \
	do {                                          \
		int rc__ = materialize_borrowed(ctx);     \
		if(!rc__) rc__ = planes_usable(ctx);      \
		if(rc__) return rc__;                     \
	} while(0)

static void set_err(ccg_ctx *ctx, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(ctx ? ctx->err : g_init_err, 512, fmt, ap);
	va_end(ap);
}

void ccg_set_err(ccg_ctx *ctx, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(ctx ? ctx->err : g_init_err, 512, fmt, ap);
	va_end(ap);
}

#define CK(ctx, call)                                                                              \
	do {                                                                                           \
		cudaError_t e__ = (call);                                                                  \
		if(e__ != cudaSuccess) {                                                                   \
			set_err(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
			return CCG_ERR_CUDA;                                                                   \
		}                                                                                          \
	} while(0)

extern "C" const char *ccg_strerror(int code) {
	switch(code) {
		case CCG_OK: return "ok";
		case CCG_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (there is no CPU fallback)";
		case CCG_ERR_CUDA: return "CUDA call failed";
		case CCG_ERR_ARG: return "invalid argument or call order";
		case CCG_ERR_NOMEM: return "out of memory";
		case CCG_ERR_UNSUPPORTED: return "unsupported on the GPU path";
		default: return "unknown error";
	}
}

extern "C" const char *ccg_last_error(const ccg_ctx *ctx) { return ctx ? ctx->err : g_init_err; }

extern "C" int ccg_init(ccg_ctx **out, int device) {
	int count = 0;
	if(!out) return CCG_ERR_ARG;
	*out = 0;
	if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
		set_err(0, "no CUDA device visible: %s", cudaGetErrorString(cudaGetLastError()));
		return CCG_ERR_NO_DEVICE;
	}
	if(device < 0 && cudaGetDevice(&device) != cudaSuccess) device = 0;
	if(device >= count) {
		set_err(0, "device %d requested, %d visible", device, count);
		return CCG_ERR_NO_DEVICE;
	}
	cudaDeviceProp prop;
	memset(&prop, 0, sizeof(prop));
	if(cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
		set_err(0, "cudaGetDeviceProperties(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError()));
		return CCG_ERR_NO_DEVICE;
	}
	if(prop.major != 10) {
		set_err(0, "device %d is compute capability %d.%d; this library is built for sm_100a only", device, prop.major,
		        prop.minor);
		return CCG_ERR_NO_DEVICE;
	}
	ccg_ctx *ctx = (ccg_ctx *) calloc(1, sizeof(ccg_ctx));
	if(ctx) ctx->trim_len = -1;
	if(!ctx) return CCG_ERR_NOMEM;
	ctx->device = device;
	ctx->sm_count = prop.multiProcessorCount;
	ctx->world = 1;
	if(cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
	   cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
		set_err(0, "context set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
		free(ctx);
		return CCG_ERR_CUDA;
	}
	ctx->stream = ctx->own_stream;
	{
		/* the fork / join ordering of run_umma rests on these: a failed creation must not degrade to stream 0 */
		cudaError_t e = cudaSuccess;
		for(int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaEventCreate(&ctx->ev_phase[k]);
		if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
		for(int k = 0; k < 2 && e == cudaSuccess; ++k) {
			e = cudaStreamCreateWithFlags(&ctx->copy_stream[k], cudaStreamNonBlocking);
			if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_up[k], cudaEventDisableTiming);
			if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_x[k], cudaEventDisableTiming);
			if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_g[k], cudaEventDisableTiming);
		}
		if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming);
		if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
		if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_launch, cudaEventDisableTiming);
		if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_switch, cudaEventDisableTiming);
		if(e != cudaSuccess) {
			set_err(0, "creating the context's streams / events failed: %s", cudaGetErrorString(e));
			cudaGetLastError();
			ccg_destroy(ctx);
			return CCG_ERR_CUDA;
		}
	}
	{
		/* stream memory operations: let the aux stream wait until the GEMM CTAs are resident */
		void *fn = 0;
		cudaDriverEntryPointQueryResult qres;
		if(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
		   qres == cudaDriverEntryPointSuccess && !getenv("CCG_NOGATE"))
			ctx->fn_wait_value = fn;
		else cudaGetLastError();
		if(cudaMalloc(&ctx->d_resident, sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); ctx->d_resident = 0; }
	}
	/* tuning knobs for experiments (scripts/one_step.py); unset in normal use */
	ctx->watchdog_cycles = getenv("CCG_WATCHDOG_S") ? (long long) (atof(getenv("CCG_WATCHDOG_S")) * 2.0e9) : 4000000000LL;
	if(getenv("CCG_NOTHIN")) ctx->dbg_nothin = atoi(getenv("CCG_NOTHIN"));
	if(getenv("CCG_KSLICES")) ctx->dbg_kslices = atoi(getenv("CCG_KSLICES"));
	if(getenv("CCG_EXPAND_SERIAL")) ctx->dbg_serial = atoi(getenv("CCG_EXPAND_SERIAL"));
	if(getenv("CCG_NOLOCK")) ctx->dbg_nolock = atoi(getenv("CCG_NOLOCK"));
	if(getenv("CCG_UMMA1")) ctx->dbg_umma1 = atoi(getenv("CCG_UMMA1"));
	if(getenv("CCG_I8")) ctx->use_i8 = atoi(getenv("CCG_I8"));
	ctx->min_slabs = getenv("CCG_MIN_SLABS") ? atoi(getenv("CCG_MIN_SLABS")) : 0;     /* 0 = automatic */
	if(getenv("CCG_FEED_SLABS")) ctx->dbg_feed_slabs = atoi(getenv("CCG_FEED_SLABS"));
	if(getenv("CCG_FEED_PLANES")) ctx->dbg_feed_planes = atoi(getenv("CCG_FEED_PLANES"));
	/* host rows are streamed K slab by K slab from this alignment length on (0 = never) */
	ctx->stream_min_chunks = getenv("CCG_STREAM_MIN_CHUNKS") ? atoi(getenv("CCG_STREAM_MIN_CHUNKS")) : 4096;
	*out = ctx;
	return CCG_OK;
}

static void free_problem(ccg_ctx *ctx) {
	cudaFree(ctx->d_planes); ctx->d_planes = 0;
	cudaFree(ctx->d_gmask); ctx->d_gmask = 0;
	cudaFree(ctx->d_inc); ctx->d_inc = 0;
	cudaFree(ctx->d_rank); ctx->d_rank = 0;
	cudaFree(ctx->d_X); ctx->d_X = 0; ctx->x_bytes = 0; ctx->tmap_thin_valid = 0;
	cudaFree(ctx->d_C); ctx->d_C = 0; ctx->c_bytes = 0;
	free(ctx->present); ctx->present = 0;
	free(ctx->need); ctx->need = 0;
	free(ctx->have); ctx->have = 0;
	free(ctx->h_rank); ctx->h_rank = 0;
	ctx->tmap_valid = 0;
	ctx->n = ctx->len = 0;
	ctx->last_Dn = 0;
	ctx->last_ntiles = 0;
}

extern "C" void ccg_destroy(ccg_ctx *ctx) {
	if(!ctx) return;
	if(ctx->multi) ccg_multi_destroy(ctx);
	cudaSetDevice(ctx->device);
	if(ctx->stream) cudaStreamSynchronize(ctx->stream);
	ccg_group_release(ctx);
	cudaFree(ctx->grp_own_win);
	cudaFree(ctx->d_row_base);
	cudaFree(ctx->d_unfed);
	free(ctx->h_row_base);
	free_problem(ctx);
	ccg_mat_free(ctx);
	ccg_trim_free(ctx);
	cudaFree(ctx->d_stage);
	cudaFree(ctx->d_motif_lens);
	cudaFree(ctx->d_motif_sets);
	cudaFree(ctx->d_acc);
	cudaFree(ctx->d_tickets);
	cudaFree(ctx->d_tiles);
	cudaFree(ctx->d_sync);
	cudaFree(ctx->d_resident);
	cudaFree(ctx->d_out_D);
	cudaFree(ctx->d_out_N);
	/* a context whose creation failed half way holds null handles */
	if(ctx->ev0) cudaEventDestroy(ctx->ev0);
	if(ctx->ev1) cudaEventDestroy(ctx->ev1);
	for(int k = 0; k < 4; ++k) if(ctx->ev_phase[k]) cudaEventDestroy(ctx->ev_phase[k]);
	if(ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
	if(ctx->ev_launch) cudaEventDestroy(ctx->ev_launch);
	if(ctx->ev_switch) cudaEventDestroy(ctx->ev_switch);
	for(int k = 0; k < 2; ++k) {
		if(ctx->ev_x[k]) cudaEventDestroy(ctx->ev_x[k]);
		if(ctx->ev_g[k]) cudaEventDestroy(ctx->ev_g[k]);
	}
	if(ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
	for(int k = 0; k < 2; ++k) {
		if(ctx->copy_stream[k]) { cudaStreamSynchronize(ctx->copy_stream[k]); cudaStreamDestroy(ctx->copy_stream[k]); }
		if(ctx->ev_up[k]) cudaEventDestroy(ctx->ev_up[k]);
		cudaFree(ctx->d_stage2[k]);
	}
	if(ctx->ev_main) cudaEventDestroy(ctx->ev_main);
	if(ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
	cudaGetLastError();
	free(ctx);
}

extern "C" int ccg_set_stream(ccg_ctx *ctx, void *cuda_stream) {
	if(!ctx) return CCG_ERR_ARG;
	if(ctx->multi) {
		if(!cuda_stream) return CCG_OK;
		set_err(ctx, "a multi-GPU context launches on its members' own streams");
		return CCG_ERR_UNSUPPORTED;
	}
	cudaStream_t next = cuda_stream ? (cudaStream_t) cuda_stream : ctx->own_stream;
	if(next != ctx->stream) {
		/* work still queued on the old stream (e.g. the clearing of a fresh sample store by ccg_set_problem) must
		 * be finished before anything launched on the new one touches the same buffers */
		CK(ctx, cudaSetDevice(ctx->device));
		CK(ctx, cudaEventRecord(ctx->ev_switch, ctx->stream));
		CK(ctx, cudaStreamWaitEvent(next, ctx->ev_switch, 0));
		ctx->stream = next;
	}
	return CCG_OK;
}

extern "C" int ccg_set_kernel(ccg_ctx *ctx, int kernel) {
	if(!ctx || kernel < CCG_KERNEL_AUTO || kernel > CCG_KERNEL_FUSED) return CCG_ERR_ARG;
	if(ctx->multi) return ccg_multi_set_kernel(ctx, kernel);
	ctx->kernel_choice = kernel;
	return CCG_OK;
}

extern "C" int ccg_set_proximity(ccg_ctx *ctx, unsigned proxi, int snp_events_only) {
	if(!ctx) return CCG_ERR_ARG;
	if(ctx->multi) {
		ccg_multi_note_special(ctx, 1, proxi != 0);
		return ccg_set_proximity(ccg_multi_member(ctx, 0), proxi, snp_events_only);
	}
	ctx->proxi = proxi;
	ctx->proxi_snp_only = snp_events_only ? 1 : 0;
	return CCG_OK;
}

extern "C" int ccg_set_scratch_limit(ccg_ctx *ctx, size_t bytes) {
	if(!ctx) return CCG_ERR_ARG;
	if(ctx->multi) {
		int rc = CCG_OK;
		for(int g = 0; ccg_multi_member(ctx, g) && !rc; ++g) rc = ccg_set_scratch_limit(ccg_multi_member(ctx, g), bytes);
		return rc;
	}
	if(ctx->x_budget != bytes && ctx->d_X) {
		/* re-size the operand panel on the next run */
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_X);
		ctx->d_X = 0;
		ctx->x_bytes = 0;
		ctx->x_chunks = 0;
		ctx->tmap_thin_valid = 0;
	}
	ctx->x_budget = bytes;
	return CCG_OK;
}

extern "C" int ccg_sync(ccg_ctx *ctx) {
	if(!ctx) return CCG_ERR_ARG;
	if(ctx->multi) return ccg_multi_sync(ctx);
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return CCG_OK;
}

extern "C" int ccg_set_tile_window(ccg_ctx *ctx, int row_lo, int row_hi, int col_lo, int col_hi) {
	if(!ctx) return CCG_ERR_ARG;
	CCG_MULTI_SOLO(ctx, "a tile window", ccg_set_tile_window(m0, row_lo, row_hi, col_lo, col_hi));
	if(row_lo < 0) {
		ctx->win_on = 0;
	} else {
		if(row_lo % CCG_UMMA_BM || col_lo % CCG_UMMA_BN || row_hi <= row_lo || col_hi <= col_lo) return CCG_ERR_ARG;
		ctx->win_on = 1;
		ctx->win[0] = row_lo / CCG_UMMA_BM;
		ctx->win[1] = (row_hi + CCG_UMMA_BM - 1) / CCG_UMMA_BM;
		ctx->win[2] = col_lo / CCG_UMMA_BN;
		ctx->win[3] = (col_hi + CCG_UMMA_BN - 1) / CCG_UMMA_BN;
	}
	update_need(ctx);
	return CCG_OK;
}

extern "C" int ccg_set_partition(ccg_ctx *ctx, int rank, int world) {
	if(!ctx || world < 1 || rank < 0 || rank >= world) return CCG_ERR_ARG;
	CCG_MULTI_SOLO(ctx, "a tile partition", ccg_set_partition(m0, rank, world));
	if(world > 1 && ctx->grp_world > 1) {
		set_err(ctx, "a context is either a member of a K-split group or a rank of a tile partition, not both");
		return CCG_ERR_ARG;
	}
	ctx->rank = rank;
	ctx->world = world;
	update_need(ctx);
	return CCG_OK;
}

/* ---- the macro-tile deal (pure host arithmetic) ---- */
extern "C" int ccg_tile_rows(void) { return CCG_UMMA_BM; }
extern "C" int ccg_tile_cols(void) { return CCG_UMMA_BN; }

/* The deal: macro tiles (tm, tn <= tm) of 256 x 256 samples are ordered along a Z-order (Morton)
 * curve and cut into `world` contiguous runs of equal estimated
 * cost.  A run is a compact 2-D region, so a rank touches only O(sqrt(tiles)) row blocks: that
 * is what keeps the per-rank encode / expansion work from being replicated on every GPU. */
static inline unsigned long long spread_bits(unsigned x) {
	unsigned long long v = x;
	v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
	v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
	v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
	v = (v | (v << 2)) & 0x3333333333333333ull;
	v = (v | (v << 1)) & 0x5555555555555555ull;
	return v;
}

struct MacroTile {
	unsigned long long key;
	int tm, tn;
};

static void macro_tiles_sorted(int n, std::vector<MacroTile> &out) {
	const int TM = (n + CCG_UMMA_BM - 1) / CCG_UMMA_BM;
	out.clear();
	for(int tm = 0; tm < TM; ++tm)
		for(int tn = 0; tn <= tm; ++tn) {
			MacroTile t;
			t.key = (spread_bits((unsigned) tm) << 1) | spread_bits((unsigned) tn);
			t.tm = tm;
			t.tn = tn;
			out.push_back(t);
		}
	std::sort(out.begin(), out.end(), [](const MacroTile &a, const MacroTile &b) { return a.key < b.key; });
}

/* Cost of giving one rank the curve segment [lo, hi): its macro tiles (GEMM time) plus the row
 * blocks they read (encode + expansion time of those 128 samples).  Measured on B200 at 5 Mbp:
 * 0.78 ms per 256 x 256 tile, 0.58 ms per row block -> 0.75 tiles per block. */
static const double kRowBlockCost = 0.75;

/* greedy walk along the curve: how many segments of cost <= cap are needed; optionally records cuts */
static int segments_for_cap(const std::vector<MacroTile> &t, int nblocks, double cap, std::vector<long long> *cuts) {
	std::vector<int> stamp((size_t) nblocks + 2, -1);
	int seg = 0;
	double cost = 0.0;
	if(cuts) { cuts->clear(); cuts->push_back(0); }
	for(size_t k = 0; k < t.size(); ++k) {
		const int blk[4] = {2 * t[k].tm, 2 * t[k].tm + 1, 2 * t[k].tn, 2 * t[k].tn + 1};
		const int nb = t[k].tm == t[k].tn ? 2 : 4;              /* a diagonal tile reads two blocks */
		double add = 1.0;
		for(int q = 0; q < nb; ++q)
			if(blk[q] < nblocks && stamp[(size_t) blk[q]] != seg) add += kRowBlockCost;
		if(cost > 0.0 && cost + add > cap) {
			++seg;
			cost = 0.0;
			if(cuts) cuts->push_back((long long) k);
			add = 1.0;
			for(int q = 0; q < nb; ++q)
				if(blk[q] < nblocks) add += kRowBlockCost;
		}
		for(int q = 0; q < nb; ++q)
			if(blk[q] < nblocks) stamp[(size_t) blk[q]] = seg;
		cost += add;
	}
	if(cuts) cuts->push_back((long long) t.size());
	return seg + 1;
}

/* cut points of the curve for `world` ranks: smallest per-rank cost cap that needs <= world segments */
static void balanced_cuts(const std::vector<MacroTile> &t, int n, int world, std::vector<long long> &cuts) {
	const int nblocks = (n + 127) / 128;
	if(world <= 1 || t.size() <= 1) {
		cuts.assign(1, 0);
		cuts.push_back((long long) t.size());
		while((int) cuts.size() < world + 1) cuts.push_back((long long) t.size());
		return;
	}
	double lo = 1.0, hi = (double) t.size() + kRowBlockCost * nblocks + 1.0;
	for(int it = 0; it < 48; ++it) {
		const double mid = 0.5 * (lo + hi);
		if(segments_for_cap(t, nblocks, mid, 0) <= world) hi = mid;
		else lo = mid;
	}
	segments_for_cap(t, nblocks, hi, &cuts);
	while((int) cuts.size() < world + 1) cuts.push_back((long long) t.size());
}

/* visits the macro tiles (tm, tn) of an n-sample triangle owned by rank, in curve order */
template <class F>
static long long for_each_macro_tile(int n, int rank, int world, F f) {
	if(n < 2 || world < 1 || rank < 0 || rank >= world) return 0;
	std::vector<MacroTile> t;
	std::vector<long long> cuts;
	macro_tiles_sorted(n, t);
	balanced_cuts(t, n, world, cuts);
	const long long lo = cuts[(size_t) rank], hi = cuts[(size_t) rank + 1];
	for(long long k = lo; k < hi; ++k) f(t[(size_t) k].tm, t[(size_t) k].tn);
	return hi - lo;
}

/* the macro tiles this context computes: its share of the curve, restricted to the tile window
 * (ccg_set_tile_window; the sample-shard ring computes one rectangular block per step) */
template <class F>
static long long for_each_ctx_tile(const ccg_ctx *ctx, F f) {
	long long k = 0;
	for_each_macro_tile(ctx->n, ctx->rank, ctx->world, [&](int tm, int tn) {
		if(ctx->win_on && (tm < ctx->win[0] || tm >= ctx->win[1] || tn < ctx->win[2] || tn >= ctx->win[3])) return;
		f(tm, tn);
		++k;
	});
	return k;
}

/* row blocks (128 slots) the owned macro tiles read: A rows 2tm, 2tm+1, B rows 2tn, 2tn+1 */
static void update_need(ccg_ctx *ctx) {
	if(!ctx->need) return;
	const int nblocks = ctx->n_pad / 128;
	memset(ctx->need, 0, (size_t) nblocks);
	for_each_ctx_tile(ctx, [&](int tm, int tn) {
		const int blk[4] = {2 * tm, 2 * tm + 1, 2 * tn, 2 * tn + 1};
		for(int q = 0; q < 4; ++q)
			if(blk[q] < nblocks) ctx->need[blk[q]] = 1;
	});
}

extern "C" long long ccg_partition_tiles(int n, int rank, int world, int *tm_out, int *tn_out, long long cap) {
	long long k = 0;
	return for_each_macro_tile(n, rank, world, [&](int tm, int tn) {
		if(k < cap && tm_out && tn_out) { tm_out[k] = tm; tn_out[k] = tn; }
		++k;
	});
}

extern "C" long long ccg_partition_cells(int n, int rank, int world) {
	long long cells = 0;
	for_each_macro_tile(n, rank, world, [&](int tm, int tn) {
		const int i0 = tm * CCG_UMMA_BM, i1 = i0 + CCG_UMMA_BM < n ? i0 + CCG_UMMA_BM : n;
		const int j0 = tn * CCG_UMMA_BN;
		for(int i = i0; i < i1; ++i) {
			int j1 = j0 + CCG_UMMA_BN;
			if(j1 > i) j1 = i;
			if(j1 > j0) cells += j1 - j0;
		}
	});
	return cells;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encoder(ccg_ctx *ctx, EncodeTiledFn *out) {
	void *fn = 0;
	cudaDriverEntryPointQueryResult qres;
	CK(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
	if(!fn || qres != cudaDriverEntryPointSuccess) {
		set_err(ctx, "cuTensorMapEncodeTiled not available from the driver");
		return CCG_ERR_CUDA;
	}
	*out = (EncodeTiledFn) fn;
	return CCG_OK;
}

static int make_planes_tmap(ccg_ctx *ctx) {
	EncodeTiledFn enc;
	int rc = get_encoder(ctx, &enc);
	if(rc) return rc;
	cuuint64_t gdim[4] = {CCG_CHUNK_WORDS, (cuuint64_t) ctx->n_pad, (cuuint64_t) ctx->nplanes, (cuuint64_t) ctx->chunks};
	cuuint64_t gstride[3] = {16, (cuuint64_t) 16 * ctx->n_pad, (cuuint64_t) 16 * ctx->n_pad * ctx->nplanes};
	cuuint32_t box[4] = {CCG_CHUNK_WORDS, CCG_TILE, (cuuint32_t) ctx->nplanes, (cuuint32_t) ccg_popc_kc()};
	cuuint32_t estr[4] = {1, 1, 1, 1};
	CUresult r = enc(&ctx->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, ctx->d_planes, gdim, gstride, box, estr,
	                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
	                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if(r != CUDA_SUCCESS) {
		set_err(ctx, "cuTensorMapEncodeTiled(planes) failed with CUresult %d", (int) r);
		return CCG_ERR_CUDA;
	}
	cuuint32_t box_pl[4] = {CCG_CHUNK_WORDS, 128, (cuuint32_t) ctx->nplanes, 1};
	r = enc(&ctx->tmap_pl, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, ctx->d_planes, gdim, gstride, box_pl, estr,
	        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
	        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if(r != CUDA_SUCCESS) {
		set_err(ctx, "cuTensorMapEncodeTiled(planes, fused box) failed with CUresult %d", (int) r);
		return CCG_ERR_CUDA;
	}
	ctx->tmap_valid = 1;
	return CCG_OK;
}

/* operand panel X (see k_pairdist_umma.cu), 128-byte swizzled boxes of 128 rows */
static int make_x_tmap(ccg_ctx *ctx) {
	EncodeTiledFn enc;
	int rc = get_encoder(ctx, &enc);
	if(rc) return rc;
	/* tile-blocked panel seen as a flat [n_pad * k-blocks][128 B] matrix: a 128-row box is one
	 * contiguous 16 KiB (row block, k-block) tile */
	cuuint64_t gdim[2] = {128, (cuuint64_t) (ctx->x_bytes / 128)};
	cuuint64_t gstride[1] = {128};
	cuuint32_t box[2] = {128, 128};
	cuuint32_t estr[2] = {1, 1};
	CUresult r = enc(&ctx->tmap_x, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ctx->d_X, gdim, gstride, box, estr,
	                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
	                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if(r != CUDA_SUCCESS) {
		set_err(ctx, "cuTensorMapEncodeTiled(X) failed with CUresult %d", (int) r);
		return CCG_ERR_CUDA;
	}
	return CCG_OK;
}

/* the panel again, with a box of `rows` rows: what one CTA stages as its half of a thin item's B operand */
cudaError_t ccg_make_thin_tmap(ccg_ctx *ctx, int rows) {
	EncodeTiledFn enc;
	if(get_encoder(ctx, &enc)) return cudaErrorUnknown;
	cuuint64_t gdim[2] = {128, (cuuint64_t) (ctx->x_bytes / 128)};
	cuuint64_t gstride[1] = {128};
	cuuint32_t box[2] = {128, (cuuint32_t) rows};
	cuuint32_t estr[2] = {1, 1};
	CUresult r = enc(&ctx->tmap_thin, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ctx->d_X, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if(r != CUDA_SUCCESS) {
		set_err(ctx, "cuTensorMapEncodeTiled(X, thin box of %d rows) failed with CUresult %d", rows, (int) r);
		return cudaErrorUnknown;
	}
	ctx->tmap_thin_rows = rows;
	ctx->tmap_thin_valid = 1;
	return cudaSuccess;
}

extern "C" int ccg_set_problem(ccg_ctx *ctx, int n, int len, int pair_mode) {
	if(!ctx || n < 0 || len < 0) return CCG_ERR_ARG;
	if(ctx->multi) return ccg_multi_set_problem(ctx, n, len, pair_mode);
	CK(ctx, cudaSetDevice(ctx->device));
	const int words = (len >> 5) + ((len & 31) ? 1 : 0);
	int chunks = (words + CCG_CHUNK_WORDS - 1) / CCG_CHUNK_WORDS;
	int n_pad = ((n + CCG_SLOT_PAD - 1) / CCG_SLOT_PAD) * CCG_SLOT_PAD;
	if(n_pad == 0) n_pad = CCG_SLOT_PAD;
	if(chunks == 0) chunks = 1;
	/* same geometry as the resident store: keep the allocations.  Stale planes of slots
	 * that are not uploaded again are harmless -- a slot that is absent or excluded gets
	 * rank -1 and none of its cells is ever written. */
	if(ctx->d_planes && ctx->tmap_valid && ctx->words == words && ctx->chunks == chunks && ctx->n_pad == n_pad &&
	   ctx->pair_mode == (pair_mode ? 1 : 0) && ctx->len == len) {
		ctx->n = n;
		memset(ctx->present, 0, (size_t) ctx->n_pad);
		memset(ctx->have, 0, (size_t) ctx->n_pad / 128);
		update_need(ctx);
		ctx->global_inc = 0;
		ctx->global_applied = 0;
		ctx->global_pending = 0;
		ctx->remask_pending = 0;
		ctx->motif_applied = 0;
		ctx->codes_upload_masked = 0;
		ctx->bor_pending = 0;
		ctx->planes_stale = 0;
		ctx->last_Dn = 0;
		ctx->last_ntiles = 0;
		return CCG_OK;
	}
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	free_problem(ctx);
	ctx->n = n;
	ctx->len = len;
	ctx->pair_mode = pair_mode ? 1 : 0;
	ctx->nplanes = pair_mode ? 3 : 2;
	ctx->words = words;
	ctx->chunks = chunks;
	ctx->n_pad = n_pad;
	ctx->planes_bytes = (size_t) ctx->chunks * ctx->nplanes * ctx->n_pad * 16;
	if(cudaMalloc(&ctx->d_planes, ctx->planes_bytes) != cudaSuccess) {
		set_err(ctx, "cudaMalloc of %zu bytes for the sample store failed: %s", ctx->planes_bytes,
		        cudaGetErrorString(cudaGetLastError()));
		ctx->d_planes = 0;
		return CCG_ERR_NOMEM;
	}
	CK(ctx, cudaMemsetAsync(ctx->d_planes, 0, ctx->planes_bytes, ctx->stream));
	CK(ctx, cudaMalloc(&ctx->d_inc, (size_t) ctx->n_pad * sizeof(unsigned)));
	CK(ctx, cudaMemsetAsync(ctx->d_inc, 0, (size_t) ctx->n_pad * sizeof(unsigned), ctx->stream));
	CK(ctx, cudaMalloc(&ctx->d_rank, (size_t) ctx->n_pad * sizeof(int)));
	if(!pair_mode) {
		CK(ctx, cudaMalloc(&ctx->d_gmask, (size_t) (ctx->words + 1) * sizeof(uint32_t)));
		CK(ctx, cudaMemsetAsync(ctx->d_gmask, 0, (size_t) (ctx->words + 1) * sizeof(uint32_t), ctx->stream));
	}
	ctx->present = (unsigned char *) calloc((size_t) ctx->n_pad, 1);
	ctx->need = (unsigned char *) calloc((size_t) ctx->n_pad / 128, 1);
	ctx->have = (unsigned char *) calloc((size_t) ctx->n_pad / 128, 1);
	ctx->h_rank = (int *) malloc((size_t) ctx->n_pad * sizeof(int));
	if(!ctx->present || !ctx->h_rank || !ctx->need || !ctx->have) return CCG_ERR_NOMEM;
	update_need(ctx);
	ctx->global_applied = 0;
	ctx->global_pending = 0;
	ctx->remask_pending = 0;
	ctx->motif_applied = 0;
	ctx->codes_upload_masked = 0;
	ctx->bor_pending = 0;
	ctx->planes_stale = 0;
	ctx->global_inc = 0;
	return make_planes_tmap(ctx);
}

static int ensure_stage(ccg_ctx *ctx, size_t bytes) {
	if(ctx->stage_bytes >= bytes) return CCG_OK;
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(ctx->d_stage);
	ctx->d_stage = 0;
	ctx->stage_bytes = 0;
	if(cudaMalloc(&ctx->d_stage, bytes) != cudaSuccess) {
		set_err(ctx, "cudaMalloc of %zu staging bytes failed", bytes);
		return CCG_ERR_NOMEM;
	}
	ctx->stage_bytes = bytes;
	return CCG_OK;
}

extern "C" int ccg_put_global_mask(ccg_ctx *ctx, const uint32_t *mask) {
	if(ctx && ctx->multi && mask) return ccg_multi_put_global_mask(ctx, mask, 0);
	if(!ctx || !ctx->d_planes || ctx->pair_mode || !mask) return CCG_ERR_ARG;
	CK(ctx, cudaSetDevice(ctx->device));
	CK(ctx, cudaMemcpyAsync(ctx->d_gmask, mask, (size_t) ctx->words * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
	unsigned inc = 0;
	for(int w = 0; w < ctx->words; ++w) inc += (unsigned) __builtin_popcount(mask[w]);
	ctx->global_inc = inc;      /* getNpos(*includes, len), fsacmpthrd.c:164 */
	return CCG_OK;
}

extern "C" int ccg_apply_global_mask(ccg_ctx *ctx, const uint32_t *mask) {
	if(ctx && ctx->multi && mask) return ccg_multi_put_global_mask(ctx, mask, 1);
	if(!ctx || !ctx->d_planes || !ctx->pair_mode || !mask) return CCG_ERR_ARG;
	CK(ctx, cudaSetDevice(ctx->device));
	NEED_PLANES(ctx);
	if(!ctx->d_gmask) CK(ctx, cudaMalloc(&ctx->d_gmask, (size_t) (ctx->words + 1) * sizeof(uint32_t)));
	CK(ctx, cudaMemcpyAsync(ctx->d_gmask, mask, (size_t) ctx->words * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
	unsigned inc = 0;
	for(int w = 0; w < ctx->words; ++w) inc += (unsigned) __builtin_popcount(mask[w]);
	ctx->global_inc = inc;
	CK(ctx, ccg_launch_apply_global_mask(ctx));
	ctx->global_applied = 1;
	return CCG_OK;
}

/* d_use <- present (and uploaded) slots with include[i] != 0; returns the first such slot in *first (-1 = none) */
static int stage_use_flags(ccg_ctx *ctx, const unsigned char *include, int lo, int hi, unsigned char **d_use_out,
                           unsigned **d_aux_out, size_t aux_words, int *first) {
	unsigned char *h_use = (unsigned char *) calloc((size_t) ctx->n_pad, 1);
	if(!h_use) return CCG_ERR_NOMEM;
	*first = -1;
	for(int i = lo; i < hi; ++i) {
		h_use[i] = ctx->present[i] && ctx->have[i >> 7] && (!include || include[i]);
		if(h_use[i] && *first < 0) *first = i;
	}
	const size_t use_bytes = ((size_t) ctx->n_pad + 15) & ~(size_t) 15;
	int rc = ensure_stage(ctx, use_bytes + aux_words * sizeof(unsigned) + 64);
	if(rc) { free(h_use); return rc; }
	unsigned char *d_use = (unsigned char *) ctx->d_stage;
	cudaError_t e = cudaMemcpyAsync(d_use, h_use, (size_t) ctx->n_pad, cudaMemcpyHostToDevice, ctx->stream);
	if(e == cudaSuccess) e = cudaMemsetAsync(d_use + use_bytes, 0, aux_words * sizeof(unsigned), ctx->stream);
	/* h_use is pageable: the copy has been staged by the time cudaMemcpyAsync returns */
	free(h_use);
	if(e != cudaSuccess) {
		set_err(ctx, "staging the slot flags failed: %s", cudaGetErrorString(e));
		return CCG_ERR_CUDA;
	}
	*d_use_out = d_use;
	*d_aux_out = (unsigned *) (d_use + use_bytes);
	return CCG_OK;
}

extern "C" int ccg_build_global_mask(ccg_ctx *ctx, const unsigned char *include, unsigned *global_inc) {
	if(ctx && ctx->multi) return ccg_multi_build_global_mask(ctx, include, global_inc);
	if(!ctx || !ctx->d_planes || !ctx->pair_mode) return CCG_ERR_ARG;
	if(ctx->world > 1) {
		set_err(ctx, "ccg_build_global_mask needs every sample on this device (partitioned contexts hold only their row blocks)");
		return CCG_ERR_UNSUPPORTED;
	}
	CK(ctx, cudaSetDevice(ctx->device));
	NEED_PLANES(ctx);
	if(!ctx->d_gmask) CK(ctx, cudaMalloc(&ctx->d_gmask, (size_t) (ctx->words + 1) * sizeof(uint32_t)));
	/* -y together with -P: the events of the proximity pass are defined on the sequences (cdist.c:109-111: maskMotifs
	 * only narrows the shared mask, getIncPosPtr then looks at the codes), so the motif sites must not be in the
	 * samples' mask planes yet, where they would read as unknown bases.  The masking is done HERE, after the
	 * proximity pass; a caller that has already run ccg_mask_motifs on this problem cannot be served. */
	const int motifs_here = ctx->motif_n > 0 && !ctx->motif_applied && ctx->words > 0;
	if(ctx->motif_n > 0 && ctx->motif_applied && ctx->proxi) {
		set_err(ctx, "shared mask with -y and -P: leave the motif masking to ccg_build_global_mask (no ccg_mask_motifs before it)");
		return CCG_ERR_UNSUPPORTED;
	}
	unsigned char *d_use = 0;
	unsigned *d_cnt = 0;
	int first = -1;
	int rc = stage_use_flags(ctx, include, 0, ctx->n, &d_use, &d_cnt, 4 + (motifs_here ? (size_t) ctx->n_pad : 0), &first);
	if(rc) return rc;
	cudaError_t e = cudaMemsetAsync(ctx->d_gmask, 0, (size_t) (ctx->words + 1) * sizeof(uint32_t), ctx->stream);
	if(e == cudaSuccess) e = ccg_launch_build_global_mask(ctx, d_use, d_cnt, 0);
	unsigned *d_final = d_cnt;
	if(e == cudaSuccess && ctx->proxi && first >= 0) {
		/* -P: every included sample also clears the runs between its close events against the first included
		 * sample (getIncPosPtr(G, seq, ref, proxi), cdist.c:111; the first sample against itself, :138) */
		e = ccg_launch_sample_proxi(ctx, 1, first, d_use, 0, 0);
		if(e == cudaSuccess) e = ccg_launch_count_mask(ctx, d_cnt + 1);
		d_final = d_cnt + 1;
	}
	if(e == cudaSuccess && motifs_here && first >= 0) {
		/* maskMotifs (cdist.c:109,137) of every slot into its mask plane, then the planes ANDed into the mask */
		e = ccg_launch_motif_mask(ctx, 0, ctx->n, d_cnt + 4);
		if(e == cudaSuccess) e = ccg_launch_build_global_mask(ctx, d_use, d_cnt + 2, 1);
		d_final = d_cnt + 2;
		ctx->motif_applied = 1;
	}
	unsigned inc = 0;
	if(e == cudaSuccess) e = cudaMemcpyAsync(&inc, d_final, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream);
	if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if(e != cudaSuccess) {
		set_err(ctx, "building the global mask failed: %s", cudaGetErrorString(e));
		return CCG_ERR_CUDA;
	}
	ctx->global_inc = inc;
	if(global_inc) *global_inc = inc;
	/* the planes are ANDed with it by the first shared-mask run (run_common): until then ccg_list_variants can
	 * still see where two samples differ outside the mask, which decides the reference's position labels */
	ctx->global_applied = 1;
	ctx->global_pending = 1;
	return CCG_OK;
}

/* -y: the motif list of getMethMotifs (methparse.c:268) */
extern "C" int ccg_set_motifs(ccg_ctx *ctx, int nmotifs, const int *lens, const unsigned char *sets) {
	if(!ctx || nmotifs < 0 || (nmotifs && (!lens || !sets))) return CCG_ERR_ARG;
	if(ctx->multi) {
		ccg_multi_note_special(ctx, 2, nmotifs != 0);
		return ccg_multi_forwarded(ctx, ccg_set_motifs(ccg_multi_member(ctx, 0), nmotifs, lens, sets));
	}
	CK(ctx, cudaSetDevice(ctx->device));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(ctx->d_motif_lens); ctx->d_motif_lens = 0;
	cudaFree(ctx->d_motif_sets); ctx->d_motif_sets = 0;
	ctx->motif_n = ctx->motif_nsets = 0;
	if(nmotifs == 0) return CCG_OK;
	int total = 0;
	for(int m = 0; m < nmotifs; ++m) {
		if(lens[m] < 1 || lens[m] > 32) {
			set_err(ctx, "motif %d has %d positions: 1 .. 32 are supported", m, lens[m]);
			return CCG_ERR_UNSUPPORTED;
		}
		total += lens[m];
	}
	if((size_t) nmotifs * sizeof(int) + (size_t) total > 40000) {
		set_err(ctx, "%d motifs with %d positions do not fit the kernel's shared memory", nmotifs, total);
		return CCG_ERR_UNSUPPORTED;
	}
	CK(ctx, cudaMalloc(&ctx->d_motif_lens, (size_t) nmotifs * sizeof(int)));
	CK(ctx, cudaMalloc(&ctx->d_motif_sets, (size_t) total));
	CK(ctx, cudaMemcpy(ctx->d_motif_lens, lens, (size_t) nmotifs * sizeof(int), cudaMemcpyHostToDevice));
	CK(ctx, cudaMemcpy(ctx->d_motif_sets, sets, (size_t) total, cudaMemcpyHostToDevice));
	ctx->motif_n = nmotifs;
	ctx->motif_nsets = total;
	return CCG_OK;
}

/* maskMotifs (meth.c:141) on the uploaded slots [first, first + count) */
extern "C" int ccg_mask_motifs(ccg_ctx *ctx, int first, int count, unsigned *inc_out) {
	CCG_MULTI_SOLO(ctx, "motif masking (-y)", ccg_mask_motifs(m0, first, count, inc_out));
	if(!ctx || !ctx->d_planes || !ctx->pair_mode || first < 0 || count < 0 || first + count > ctx->n) return CCG_ERR_ARG;
	if(count == 0) return CCG_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	NEED_PLANES(ctx);
	if(ctx->motif_n && ctx->words > 0) {
		int rc = ensure_stage(ctx, (size_t) count * sizeof(unsigned) + 64);
		if(rc) return rc;
		unsigned *d_removed = (unsigned *) ctx->d_stage;
		CK(ctx, cudaMemsetAsync(d_removed, 0, (size_t) count * sizeof(unsigned), ctx->stream));
		CK(ctx, ccg_launch_motif_mask(ctx, first, count, d_removed));
		ctx->motif_applied = 1;
	}
	if(inc_out) CK(ctx, cudaMemcpyAsync(inc_out, ctx->d_inc + first, (size_t) count * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return CCG_OK;
}

extern "C" int ccg_sample_proximity(ccg_ctx *ctx, int first, int count, int apply, unsigned *inc_out) {
	CCG_MULTI_SOLO(ctx, "proximity masking (-P)", ccg_sample_proximity(m0, first, count, apply, inc_out));
	if(!ctx || !ctx->d_planes || !ctx->pair_mode || first < 0 || count < 0 || first + count > ctx->n) return CCG_ERR_ARG;
	if(count == 0) return CCG_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	NEED_PLANES(ctx);
	unsigned *h = (unsigned *) malloc((size_t) count * 2 * sizeof(unsigned));
	if(!h) return CCG_ERR_NOMEM;
	unsigned *h_inc = h, *h_clr = h + count;
	memset(h_clr, 0, (size_t) count * sizeof(unsigned));
	cudaError_t e = cudaSuccess;
	/* against itself a sample has no SNPs: only getIncPos (events = unknown positions) clears anything */
	const int active = ctx->proxi && !ctx->proxi_snp_only && ctx->words > 0;
	if(active) {
		unsigned char *d_use = 0;
		unsigned *d_clr = 0;
		int first_used = -1;
		int rc = stage_use_flags(ctx, 0, first, first + count, &d_use, &d_clr, (size_t) ctx->n_pad, &first_used);
		if(rc) { free(h); return rc; }
		e = ccg_launch_sample_proxi(ctx, 0, 0, d_use, apply ? 2 : 0, d_clr);
		if(apply) ctx->remask_pending = 1;          /* mask plane only: the code planes follow before the first run */
		if(e == cudaSuccess)
			e = cudaMemcpyAsync(h_clr, d_clr + first, (size_t) count * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream);
	}
	if(e == cudaSuccess)
		e = cudaMemcpyAsync(h_inc, ctx->d_inc + first, (size_t) count * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream);
	if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if(e == cudaSuccess) {
		for(int i = 0; i < count; ++i) h_inc[i] -= h_clr[i];
		if(inc_out) memcpy(inc_out, h_inc, (size_t) count * sizeof(unsigned));
		if(apply && active) {
			e = cudaMemcpyAsync(ctx->d_inc + first, h_inc, (size_t) count * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream);
			if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		}
	}
	free(h);
	if(e != cudaSuccess) {
		set_err(ctx, "per-sample proximity masking failed: %s", cudaGetErrorString(e));
		return CCG_ERR_CUDA;
	}
	return CCG_OK;
}

/* The count the reference's inclusion test looks at for a shared-mask REFERENCE candidate under -y and -P (cdist.c:137-140:
 * maskMotifs + getIncPosPtr(includes, seq, seq, proxi) + getNpos), without leaving a trace in the store: the candidate's
 * planes are set aside, masked, counted and put back -- ccg_build_global_mask needs them as uploaded. */
extern "C" int ccg_sample_count_masked(ccg_ctx *ctx, int slot, unsigned *inc_out) {
	CCG_MULTI_SOLO(ctx, "motif / proximity masking (-y, -P)", ccg_sample_count_masked(m0, slot, inc_out));
	if(!ctx || !ctx->d_planes || !ctx->pair_mode || slot < 0 || slot >= ctx->n || !inc_out) return CCG_ERR_ARG;
	CK(ctx, cudaSetDevice(ctx->device));
	NEED_PLANES(ctx);
	unsigned inc0 = 0, clr = 0, inc1 = 0;
	CK(ctx, cudaMemcpyAsync(&inc0, ctx->d_inc + slot, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	*inc_out = inc0;
	if(ctx->words == 0) return CCG_OK;
	const int remask0 = ctx->remask_pending;
	void *d_raw = 0;
	CK(ctx, cudaMalloc(&d_raw, (size_t) ctx->chunks * 3 * 16));
	cudaError_t e = ccg_launch_row_planes(ctx, slot, d_raw, 0);
	int rc = CCG_OK;
	if(e == cudaSuccess && ctx->proxi && !ctx->proxi_snp_only) {
		unsigned char *d_use = 0;
		unsigned *d_clr = 0;
		int first_used = -1;
		rc = stage_use_flags(ctx, 0, slot, slot + 1, &d_use, &d_clr, (size_t) ctx->n_pad, &first_used);
		if(!rc) {
			e = ccg_launch_sample_proxi(ctx, 0, 0, d_use, 2, d_clr);
			if(e == cudaSuccess) e = cudaMemcpyAsync(&clr, d_clr + slot, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream);
			if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		}
	}
	if(!rc && e == cudaSuccess && ctx->motif_n) {
		rc = ensure_stage(ctx, sizeof(unsigned) + 64);
		if(!rc) {
			unsigned *d_removed = (unsigned *) ctx->d_stage;
			e = cudaMemsetAsync(d_removed, 0, sizeof(unsigned), ctx->stream);
			if(e == cudaSuccess) e = ccg_launch_motif_mask(ctx, slot, 1, d_removed);
		}
	}
	if(!rc && e == cudaSuccess) e = cudaMemcpyAsync(&inc1, ctx->d_inc + slot, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream);
	/* back to the sample as uploaded */
	cudaError_t er = ccg_launch_row_planes(ctx, slot, d_raw, 1);
	if(er == cudaSuccess) er = cudaMemcpyAsync(ctx->d_inc + slot, &inc0, sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream);
	if(er == cudaSuccess) er = cudaStreamSynchronize(ctx->stream);
	cudaFree(d_raw);
	ctx->remask_pending = remask0;
	if(rc) return rc;
	if(e != cudaSuccess || er != cudaSuccess) {
		set_err(ctx, "counting a masked sample failed: %s", cudaGetErrorString(e != cudaSuccess ? e : er));
		return CCG_ERR_CUDA;
	}
	*inc_out = inc1 - clr;
	return CCG_OK;
}

extern "C" int ccg_put_samples_packed(ccg_ctx *ctx, int first, int count, const uint64_t *const *seqs,
                                      const uint32_t *const *includes) {
	if(ctx && ctx->multi && seqs) return ccg_multi_put_samples_packed(ctx, first, count, seqs, includes);
	if(!ctx || !ctx->d_planes || first < 0 || count < 0 || first + count > ctx->n || !seqs) return CCG_ERR_ARG;
	if(ctx->pair_mode && !includes) return CCG_ERR_ARG;
	if(count == 0 || ctx->words == 0) return CCG_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	const size_t W = (size_t) ctx->words;
	const size_t row_bytes = W * (ctx->pair_mode ? 12 : 8);
	size_t batch = ((size_t) 256 << 20) / row_bytes;
	if(batch < 1) batch = 1;
	if(batch > (size_t) count) batch = (size_t) count;
	int rc = ensure_stage(ctx, batch * row_bytes + 64);
	if(rc) return rc;
	uint64_t *d_seq = (uint64_t *) ctx->d_stage;
	uint32_t *d_msk = ctx->pair_mode ? (uint32_t *) (d_seq + batch * W) : 0;

	int k = 0;
	while(k < count) {
		/* run of consecutive present rows, at most one staging batch; rows of row blocks that no
		 * macro tile of this rank reads are registered but not uploaded */
		if(!seqs[k] || (ctx->pair_mode && !includes[k])) { ++k; continue; }
		if(!ctx->need[(first + k) >> 7]) { ctx->present[first + k] = 1; ++k; continue; }
		int run = 0;
		/* rows at one distance from each other (one host allocation, as dist.c:143-154 makes them) cross PCIe as ONE
		 * strided copy per staging batch instead of one call per row */
		bool uniform = true;
		ptrdiff_t ds = 0, dm = 0;
		while(k + run < count && (size_t) run < batch && seqs[k + run] && (!ctx->pair_mode || includes[k + run]) &&
		      ctx->need[(first + k + run) >> 7]) {
			ctx->have[(first + k + run) >> 7] = 1;
			ctx->present[first + k + run] = 1;
			if(run == 1) {
				ds = (const char *) seqs[k + 1] - (const char *) seqs[k];
				dm = ctx->pair_mode ? (const char *) includes[k + 1] - (const char *) includes[k] : 0;
				if(ds < (ptrdiff_t) (W * 8) || (ctx->pair_mode && dm < (ptrdiff_t) (W * 4))) uniform = false;
			} else if(run > 1) {
				if((const char *) seqs[k + run] - (const char *) seqs[k + run - 1] != ds) uniform = false;
				if(ctx->pair_mode && (const char *) includes[k + run] - (const char *) includes[k + run - 1] != dm) uniform = false;
			}
			++run;
		}
		if(uniform && run > 1) {
			CK(ctx, cudaMemcpy2DAsync(d_seq, W * 8, seqs[k], (size_t) ds, W * 8, (size_t) run, cudaMemcpyHostToDevice, ctx->stream));
			if(ctx->pair_mode)
				CK(ctx, cudaMemcpy2DAsync(d_msk, W * 4, includes[k], (size_t) dm, W * 4, (size_t) run, cudaMemcpyHostToDevice, ctx->stream));
		} else {
			for(int r = 0; r < run; ++r) {
				CK(ctx, cudaMemcpyAsync(d_seq + (size_t) r * W, seqs[k + r], W * 8, cudaMemcpyHostToDevice, ctx->stream));
				if(ctx->pair_mode)
					CK(ctx, cudaMemcpyAsync(d_msk + (size_t) r * W, includes[k + r], W * 4, cudaMemcpyHostToDevice, ctx->stream));
			}
		}
		CK(ctx, ccg_launch_repack(ctx, first + k, run, d_seq, d_msk, (long) W));
		k += run;
	}
	return CCG_OK;
}

extern "C" int ccg_put_samples_packed_dev(ccg_ctx *ctx, int first, int count, const uint64_t *d_seqs,
                                          const uint32_t *d_masks, long wstride) {
	CCG_MULTI_SOLO(ctx, "an upload from device memory", ccg_put_samples_packed_dev(m0, first, count, d_seqs, d_masks, wstride));
	if(!ctx || !ctx->d_planes || first < 0 || count < 0 || first + count > ctx->n || !d_seqs || wstride < ctx->words)
		return CCG_ERR_ARG;
	if(ctx->pair_mode && !d_masks) return CCG_ERR_ARG;
	if(count == 0) return CCG_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	memset(ctx->present + first, 1, (size_t) count);
	/* runs of row blocks this rank's macro tiles read */
	int k = 0;
	while(k < count) {
		if(!ctx->need[(first + k) >> 7]) { k = (((first + k) >> 7) + 1) * 128 - first; continue; }
		int e = k;
		while(e < count && ctx->need[(first + e) >> 7]) {
			ctx->have[(first + e) >> 7] = 1;
			e = (((first + e) >> 7) + 1) * 128 - first;
		}
		if(e > count) e = count;
		CK(ctx, ccg_launch_repack(ctx, first + k, e - k, d_seqs + (size_t) k * wstride,
		                          ctx->pair_mode ? d_masks + (size_t) k * wstride : 0, wstride));
		k = e;
	}
	return CCG_OK;
}

/* ccg_fsa_cmp_thread_out on a long alignment streams the host rows straight into the operand panel: afterwards the
 * context holds results and per-sample counts, but no bit-plane store of that sample set */
static int planes_usable(ccg_ctx *ctx) {
	if(!ctx->planes_stale) return CCG_OK;
	set_err(ctx, "the samples of the last ccg_fsa_cmp_thread_out call were streamed through the operand panel and left no sample "
	        "store on the device: declare the problem again (ccg_set_problem) and upload them with ccg_put_*");
	return CCG_ERR_ARG;
}

/* builds the plane store of the lent rows when something other than the tensor path's expansion needs it */
static int materialize_borrowed(ccg_ctx *ctx) {
	if(!ctx->bor_pending) return CCG_OK;
	ctx->bor_pending = 0;
	return ccg_put_samples_packed_dev(ctx, ctx->bor_first, ctx->bor_count, ctx->bor_seqs, ctx->bor_masks, ctx->bor_wstride);
}

/* Same as ccg_put_samples_packed_dev, but the rows are LENT: they must stay valid and unchanged until the run that
 * uses them has finished (ccg_sync), and the library reads them during that run. */
extern "C" int ccg_put_samples_packed_dev_borrowed(ccg_ctx *ctx, int first, int count, const uint64_t *d_seqs,
                                                   const uint32_t *d_masks, long wstride) {
	CCG_MULTI_SOLO(ctx, "an upload from device memory", ccg_put_samples_packed_dev_borrowed(m0, first, count, d_seqs, d_masks, wstride));
	if(!ctx || !ctx->d_planes || first < 0 || count < 0 || first + count > ctx->n || !d_seqs || wstride < ctx->words)
		return CCG_ERR_ARG;
	if(ctx->pair_mode && !d_masks) return CCG_ERR_ARG;
	/* one lent range at a time: an earlier loan that the new one does not replace slot for slot gets its planes now */
	int rc = CCG_OK;
	if(ctx->bor_pending && count > 0 && first <= ctx->bor_first && first + count >= ctx->bor_first + ctx->bor_count) ctx->bor_pending = 0;
	else rc = materialize_borrowed(ctx);
	if(rc || count == 0) return rc;
	/* only the e2m1 tensor path reads lent rows; everything else goes through the planes right away */
	if(ctx->use_i8 || ctx->dbg_umma1 || ctx->proxi || ctx->motif_n || ctx->world > 1 || ctx->win_on)
		return ccg_put_samples_packed_dev(ctx, first, count, d_seqs, d_masks, wstride);
	memset(ctx->present + first, 1, (size_t) count);
	for(int b = first >> 7; b <= (first + count - 1) >> 7; ++b) ctx->have[b] = 1;
	ctx->bor_seqs = d_seqs;
	ctx->bor_masks = ctx->pair_mode ? d_masks : 0;
	ctx->bor_wstride = wstride;
	ctx->bor_first = first;
	ctx->bor_count = count;
	ctx->bor_pending = 1;
	return CCG_OK;
}

extern "C" int ccg_put_sample_codes(ccg_ctx *ctx, int idx, const unsigned char *codes) {
	if(ctx && ctx->multi && codes) return ccg_multi_put_sample_codes(ctx, idx, codes);
	if(!ctx || !ctx->d_planes || !ctx->pair_mode || idx < 0 || idx >= ctx->n || !codes) return CCG_ERR_ARG;
	CK(ctx, cudaSetDevice(ctx->device));
	size_t stride = ((size_t) ctx->len + 15) & ~(size_t) 15;
	if(stride == 0) stride = 16;
	int rc = ensure_stage(ctx, stride + 64);
	if(rc) return rc;
	CK(ctx, cudaMemcpyAsync(ctx->d_stage, codes, (size_t) ctx->len, cudaMemcpyHostToDevice, ctx->stream));
	CK(ctx, ccg_launch_encode_codes(ctx, idx, 1, (const unsigned char *) ctx->d_stage, (long) stride));
	ctx->present[idx] = 1;
	ctx->have[idx >> 7] = 1;
	return CCG_OK;
}

extern "C" int ccg_get_inc_counts(ccg_ctx *ctx, unsigned *out) {
	if(ctx && ctx->multi && out) return ccg_multi_get_inc_counts(ctx, out);
	if(!ctx || !ctx->d_inc || !out) return CCG_ERR_ARG;
	CK(ctx, cudaSetDevice(ctx->device));
	{
		int rc = materialize_borrowed(ctx);    /* the per-sample counts come out of the plane build (or of a streamed run) */
		if(rc) return rc;
	}
	CK(ctx, cudaMemcpyAsync(out, ctx->d_inc, (size_t) ctx->n * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return CCG_OK;
}

/* choose the K split: fill whole waves of resident CTAs as evenly as possible.  The persistent kernels hand out
 * (tile, K slice) items; with `slots` items in flight the last wave is only partly filled, and at 820 tiles on 74
 * CTA pairs (10 k samples, one GPU) an unsplit run idles 7.7 % of the machine in it.  The smallest split that fills
 * the waves to 99 % is taken (measured there: GEMM 304 -> 282 ms with 6 slices), else the best one found. */
static int choose_split(long long slots, long long ntiles, int iters_total, int min_iters, int max_split, int even_tail = 1) {
	int best = 1;
	double best_util = -1.0;
	for(int k = 1; k <= max_split; ++k) {
		if(k > 1 && iters_total / k < min_iters) break;
		long long items = ntiles * k;
		long long waves = (items + slots - 1) / slots;
		double util = (double) items / (double) (waves * slots);
		if(util > best_util + (even_tail ? 0.005 : 0.02)) { best_util = util; best = k; }
		if(util >= 0.99) break;
		/* plenty of waves already: a few more slices are enough to even out the tail.  Not when the run is fed from
		 * the host slab by slab (even_tail == 0): there the GEMM hides behind the PCIe upload and more slices only
		 * add accumulator traffic. */
		if(items >= 8 * slots && (!even_tail || k >= 16)) break;
	}
	return best;
}

static int ensure_tiles(ccg_ctx *ctx, const int2 *host, size_t count) {
	if(ctx->tiles_cap < count) {
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_tiles);
		ctx->d_tiles = 0;
		ctx->tiles_cap = 0;
		CK(ctx, cudaMalloc(&ctx->d_tiles, count * sizeof(int2)));
		ctx->tiles_cap = count;
	}
	CK(ctx, cudaMemcpyAsync(ctx->d_tiles, host, count * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
	/* the host list is freed by the caller right after this returns */
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return CCG_OK;
}

/* the 64x64 tiles of the macro tiles this rank owns -> ctx->d_tiles; *cnt_out = how many */
static int upload_tiles64(ccg_ctx *ctx, size_t *cnt_out) {
	const int n = ctx->n;
	size_t cap = 0;
	for_each_ctx_tile(ctx, [&](int, int) { cap += (CCG_UMMA_BM / CCG_TILE) * (CCG_UMMA_BN / CCG_TILE); });
	int2 *host = (int2 *) malloc((cap ? cap : 1) * sizeof(int2));
	if(!host) return CCG_ERR_NOMEM;
	size_t cnt = 0;
	for_each_ctx_tile(ctx, [&](int tm, int tn) {
		for(int a = 0; a < CCG_UMMA_BM / CCG_TILE; ++a)
			for(int b = 0; b < CCG_UMMA_BN / CCG_TILE; ++b) {
				int ti = tm * (CCG_UMMA_BM / CCG_TILE) + a, tj = tn * (CCG_UMMA_BN / CCG_TILE) + b;
				if(tj > ti || ti * CCG_TILE >= n || tj * CCG_TILE >= n) continue;
				host[cnt].x = ti;
				host[cnt].y = tj;
				++cnt;
			}
	});
	ctx->last_ntiles = (int) cnt;
	ctx->last_kernel_kind = CCG_KERNEL_POPC;
	*cnt_out = cnt;
	if(cnt == 0) { free(host); return CCG_OK; }
	int rc = ensure_tiles(ctx, host, cnt);
	free(host);
	return rc;
}

static int ensure_acc(ccg_ctx *ctx, size_t acc_bytes) {
	if(ctx->acc_bytes >= acc_bytes) return CCG_OK;
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(ctx->d_acc);
	ctx->d_acc = 0;
	ctx->acc_bytes = 0;
	if(cudaMalloc(&ctx->d_acc, acc_bytes) != cudaSuccess) {
		set_err(ctx, "cudaMalloc of %zu accumulator bytes failed", acc_bytes);
		return CCG_ERR_NOMEM;
	}
	ctx->acc_bytes = acc_bytes;
	return CCG_OK;
}

/* pair mode with -P: k_pairdist_proxi (the masking is sequential along the alignment: no K split, no tensor path) */
static int run_proxi(ccg_ctx *ctx, const EpilogueParams &ep) {
	size_t cnt = 0;
	int rc = upload_tiles64(ctx, &cnt);
	if(rc || cnt == 0) return rc;
	rc = ensure_acc(ctx, cnt * 2 * CCG_TILE * CCG_TILE * sizeof(uint32_t));
	if(rc) return rc;
	ProxiParams p;
	memset(&p, 0, sizeof(p));
	p.ntiles = (int) cnt;
	p.tiles = ctx->d_tiles;
	p.acc = ctx->d_acc;
	p.ep = ep;
	/* cells the kernel skips (i <= j, excluded samples) must not hold stale counts for ccg_get_raw_counts */
	CK(ctx, cudaMemsetAsync(ctx->d_acc, 0, cnt * 2 * CCG_TILE * CCG_TILE * sizeof(uint32_t), ctx->stream));
	CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	CK(ctx, ccg_launch_pair_proxi(ctx, p));
	CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	ctx->ev_valid = 1;
	snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k_pairdist_proxi tiles=%d proxi=%u", p.ntiles, ctx->proxi);
	return CCG_OK;
}

static int run_popc(ccg_ctx *ctx, const EpilogueParams &ep) {
	size_t cnt = 0;
	int rc = upload_tiles64(ctx, &cnt);
	if(rc || cnt == 0) return rc;

	PopcParams p;
	memset(&p, 0, sizeof(p));
	const int kc = ccg_popc_kc();
	const int iters_total = (ctx->chunks + kc - 1) / kc;
	p.chunks = ctx->chunks;
	p.ntiles = (int) cnt;
	p.tiles = ctx->d_tiles;
	p.ksplit = choose_split(2LL * ctx->sm_count, (long long) cnt, iters_total, 8, 64);
	p.chunks_per_split = ((iters_total + p.ksplit - 1) / p.ksplit) * kc;
	while(p.ksplit > 1 && (long long) (p.ksplit - 1) * p.chunks_per_split >= ctx->chunks) --p.ksplit;

	size_t acc_bytes = cnt * 2 * CCG_TILE * CCG_TILE * sizeof(uint32_t);
	rc = ensure_acc(ctx, acc_bytes);
	if(rc) return rc;
	if(ctx->tickets_count < cnt) {
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_tickets);
		ctx->d_tickets = 0;
		CK(ctx, cudaMalloc(&ctx->d_tickets, cnt * sizeof(unsigned)));
		ctx->tickets_count = cnt;
	}
	if(p.ksplit > 1) {
		CK(ctx, cudaMemsetAsync(ctx->d_acc, 0, acc_bytes, ctx->stream));
		CK(ctx, cudaMemsetAsync(ctx->d_tickets, 0, cnt * sizeof(unsigned), ctx->stream));
	}
	p.acc = ctx->d_acc;
	p.tickets = ctx->d_tickets;
	p.ep = ep;
	CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	CK(ctx, ccg_launch_popc(ctx, p));
	CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	ctx->ev_valid = 1;
	snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k_pairdist_popc<%d> tiles=%d ksplit=%d", ctx->nplanes,
	         p.ntiles, p.ksplit);
	return CCG_OK;
}

/* ---- K-slab streaming of host rows ----
 * Uploads the words of chunks [chunk0, chunk0 + nch) of every needed row (copy stream) and
 * repacks them into the plane store; ev_up fires when the slab's planes are complete.  Rows at
 * a uniform distance (one host allocation) go as 2-D copies, others row by row. */
/* X != NULL (e2m1 panel): the staged words are expanded straight into the slab's panel buffer X (npairs chunk pairs) by
 * the copy stream itself -- no plane store, no separate expansion pass competing with the GEMM for the SMs. */
static int feed_slab(ccg_ctx *ctx, int chunk0, int nch, int8_t *X, int npairs) {
	const int w0 = chunk0 * CCG_CHUNK_WORDS;
	int nw = nch * CCG_CHUNK_WORDS;
	if(nw > ctx->words - w0) nw = ctx->words - w0;
	if(nw > 0) {
		const int pair = ctx->feed_masks != 0;
		const size_t pw = ((size_t) nw + 3) & ~(size_t) 3;          /* staging pitch in words: vector loads in k_repack_direct */
		const size_t row_bytes = pw * (pair ? 12 : 8);
		size_t batch = ((size_t) 256 << 20) / row_bytes;
		if(batch < 32) batch = 32;
		if(batch > (size_t) ctx->n) batch = (size_t) ctx->n;
		const size_t bytes = batch * row_bytes + 64;
		/* two copy streams with a staging buffer each, used in turn: the copy engines work in parallel */
		for(int q = 0; q < 2; ++q) {
			if(ctx->stage2_bytes[q] >= bytes) continue;
			CK(ctx, cudaStreamSynchronize(ctx->copy_stream[q]));
			cudaFree(ctx->d_stage2[q]);
			ctx->d_stage2[q] = 0;
			ctx->stage2_bytes[q] = 0;
			if(cudaMalloc(&ctx->d_stage2[q], bytes) != cudaSuccess) {
				set_err(ctx, "cudaMalloc of %zu staging bytes failed", bytes);
				return CCG_ERR_NOMEM;
			}
			ctx->stage2_bytes[q] = bytes;
		}
		const uint64_t *const *seqs = ctx->feed_seqs;
		const uint32_t *const *masks = ctx->feed_masks;
		int k = 0, q = 0;
		while(k < ctx->n) {
			if(!seqs[k] || (pair && !masks[k]) || !ctx->need[k >> 7]) { ++k; continue; }
			/* run of consecutive present, needed rows */
			int run = 1;
			while(k + run < ctx->n && (size_t) run < batch && seqs[k + run] && (!pair || masks[k + run]) && ctx->need[(k + run) >> 7]) ++run;
			/* uniform row distance inside the run? */
			bool uniform = run > 1;
			const ptrdiff_t ds = run > 1 ? (const char *) seqs[k + 1] - (const char *) seqs[k] : 0;
			const ptrdiff_t dm = (run > 1 && pair) ? (const char *) masks[k + 1] - (const char *) masks[k] : 0;
			for(int r = 1; uniform && r + 1 < run; ++r) {
				if((const char *) seqs[k + r + 1] - (const char *) seqs[k + r] != ds) uniform = false;
				if(pair && (const char *) masks[k + r + 1] - (const char *) masks[k + r] != dm) uniform = false;
			}
			if(uniform && (ds < (ptrdiff_t) ((size_t) nw * 8) || (pair && dm < (ptrdiff_t) ((size_t) nw * 4)))) uniform = false;
			cudaStream_t cs = ctx->copy_stream[q];
			uint64_t *d_seq = (uint64_t *) ctx->d_stage2[q];
			uint32_t *d_msk = pair ? (uint32_t *) (d_seq + batch * pw) : 0;
			if(uniform) {
				CK(ctx, cudaMemcpy2DAsync(d_seq, pw * 8, seqs[k] + w0, (size_t) ds, (size_t) nw * 8, (size_t) run,
				                          cudaMemcpyHostToDevice, cs));
				if(pair)
					CK(ctx, cudaMemcpy2DAsync(d_msk, pw * 4, masks[k] + w0, (size_t) dm, (size_t) nw * 4, (size_t) run,
					                          cudaMemcpyHostToDevice, cs));
			} else {
				for(int r = 0; r < run; ++r) {
					CK(ctx, cudaMemcpyAsync(d_seq + (size_t) r * pw, seqs[k + r] + w0, (size_t) nw * 8, cudaMemcpyHostToDevice, cs));
					if(pair)
						CK(ctx, cudaMemcpyAsync(d_msk + (size_t) r * pw, masks[k + r] + w0, (size_t) nw * 4, cudaMemcpyHostToDevice, cs));
				}
			}
			if(X) CK(ctx, ccg_launch_expand_rows(ctx, cs, X, k, run, d_seq, d_msk, (long) pw, chunk0, npairs));
			else CK(ctx, ccg_launch_repack_direct(ctx, cs, k, run, d_seq, d_msk, (long) pw, chunk0, nch));
			k += run;
			q ^= 1;
		}
	}
	if(X && ctx->n_unfed) CK(ctx, ccg_launch_zero_panel_rows(ctx, ctx->copy_stream[0], X, ctx->d_unfed, ctx->n_unfed, npairs));
	for(int q = 0; q < 2; ++q) CK(ctx, cudaEventRecord(ctx->ev_up[q], ctx->copy_stream[q]));
	return CCG_OK;
}

/* uploads this rank's macro tiles to d_tiles[0, cnt) and, behind them, their 128-row halves
 * (2tm, tn), (2tm+1, tn) for the kernels that work on 128 x 256 tiles: d_tiles[cnt, 3 cnt) */
static int upload_macro_tiles(ccg_ctx *ctx, size_t *cnt_out) {
	size_t cap = 0;
	for_each_ctx_tile(ctx, [&](int, int) { ++cap; });
	int2 *host = (int2 *) malloc((cap ? 3 * cap : 1) * sizeof(int2));
	if(!host) return CCG_ERR_NOMEM;
	size_t cnt = 0;
	for_each_ctx_tile(ctx, [&](int tm, int tn) {
		host[cnt].x = tm;
		host[cnt].y = tn;
		host[cap + 2 * cnt].x = 2 * tm;
		host[cap + 2 * cnt].y = tn;
		host[cap + 2 * cnt + 1].x = 2 * tm + 1;
		host[cap + 2 * cnt + 1].y = tn;
		++cnt;
	});
	*cnt_out = cnt;
	int rc = cnt ? ensure_tiles(ctx, host, 3 * cnt) : CCG_OK;
	free(host);
	return rc;
}

/* The operand panel, sized once per problem geometry (cudaMemGetInfo / cudaMalloc are far too slow per run): the
 * whole K axis in one buffer when it fits, else two slab buffers so that the expansion of slab s+1 runs under the
 * GEMM of slab s. */
static int ensure_panel(ccg_ctx *ctx, bool fp4) {
	int rc;
	if(ctx->d_X) return CCG_OK;
	size_t free_b = 0, total_b = 0;
	CK(ctx, cudaMemGetInfo(&free_b, &total_b));
	/* default: up to 70 % of what is free (the planes, accumulators and outputs are already allocated) */
	size_t budget = ctx->x_budget ? ctx->x_budget : free_b / 10 * 7;
	if(budget > free_b - free_b / 8) budget = free_b - free_b / 8;
	const size_t per_chunk = (size_t) ctx->n_pad * (fp4 ? 256 : 512);
	long long fit = (long long) (budget / per_chunk);        /* chunks one buffer could hold */
	int want_slabs = 1;
	if(fit < ctx->chunks) {
		fit = (long long) (budget / (2 * per_chunk));         /* two buffers */
		if(fit < 1) {
			set_err(ctx, "not enough device memory for the operand panel (%d slots)", ctx->n_pad);
			return CCG_ERR_NOMEM;
		}
		want_slabs = (int) ((ctx->chunks + fit - 1) / fit);
	}
	/* measured (profiles/): running the expansion under the GEMM does not pay on one GPU (both
	 * are limited by HBM traffic and issue slots: 13.2 ms overlapped vs 12.3 ms back to back at
	 * 1000 x 5 Mbp), so the panel is cut only when it does not fit; then the two buffers keep
	 * the tensor pipe busy while the next slab is produced */
	fit = (ctx->chunks + want_slabs - 1) / want_slabs;              /* equal slabs */
	if(fp4) fit = (fit + 1) & ~1LL;                                  /* whole chunk pairs */
	const int nbuf = want_slabs > 1 ? 2 : 1;
	ctx->x_buf_bytes = (size_t) fit * per_chunk;
	size_t x_bytes = ctx->x_buf_bytes * nbuf;
	if(cudaMalloc(&ctx->d_X, x_bytes) != cudaSuccess) {
		ctx->d_X = 0;
		set_err(ctx, "cudaMalloc of %zu bytes for the operand panel failed", x_bytes);
		return CCG_ERR_NOMEM;
	}
	ctx->x_bytes = x_bytes;
	ctx->x_chunks = (int) fit;
	ctx->tmap_thin_valid = 0;
	rc = make_x_tmap(ctx);
	if(rc) return rc;
	return CCG_OK;
}

static int run_umma_windows(ccg_ctx *ctx, const EpilogueParams &ep, int win_rows);

static int run_umma(ccg_ctx *ctx, const EpilogueParams &ep) {
	size_t cnt = 0;
	int rc;
	const bool group = ctx->grp_world > 1;
	if(group) {
		/* accumulators larger than the peer window: windows of macro-tile rows, one after the other */
		const int win_rows = ccg_group_window_rows(ctx);
		if(win_rows > 0 && win_rows < ctx->n_pad) return run_umma_windows(ctx, ep, win_rows);
	}
	rc = upload_macro_tiles(ctx, &cnt);
	if(rc) return rc;
	ctx->last_ntiles = (int) cnt;
	ctx->last_kernel_kind = CCG_KERNEL_UMMA;
	if(cnt == 0) return CCG_OK;

	/* dense int32 accumulators S and I: the context's own, or -- member of a K-split group -- the buffer of this
	 * run inside the window the peers can read (ccg_group.cu) */
	size_t c_bytes = (size_t) 2 * ctx->n_pad * ctx->n_pad * sizeof(int);
	int *acc_S = 0, *acc_I = 0;
	void *acc_clear = 0;
	if(group) {
		rc = ccg_group_accumulators(ctx, 0, &acc_S, &acc_I, &acc_clear, &c_bytes);
		if(rc) return rc;
	} else {
		if(ctx->c_bytes < c_bytes) {
			CK(ctx, cudaStreamSynchronize(ctx->stream));
			cudaFree(ctx->d_C);
			ctx->d_C = 0;
			ctx->c_bytes = 0;
			if(cudaMalloc(&ctx->d_C, c_bytes) != cudaSuccess) {
				set_err(ctx, "cudaMalloc of %zu bytes for the int32 accumulators failed", c_bytes);
				return CCG_ERR_NOMEM;
			}
			ctx->c_bytes = c_bytes;
		}
		acc_S = ctx->d_C;
		acc_I = ctx->d_C + (size_t) ctx->n_pad * ctx->n_pad;
		acc_clear = ctx->d_C;
	}
	CK(ctx, cudaMemsetAsync(acc_clear, 0, c_bytes, ctx->stream));

	/* e2m1 operands on kind::mxf4 (2.27x the kind::i8 pipe rate, exact for these sums) unless CCG_I8=1 */
	const bool fp4 = !ctx->use_i8 && !ctx->dbg_umma1;
	/* operand panel: the K axis is cut into slabs; two slab buffers so that the expansion of
	 * slab s+1 (aux stream, HBM-write bound) runs under the GEMM of slab s (tensor bound).
	 * Sized once per problem geometry (cudaMemGetInfo / cudaMalloc are far too slow per run). */
	rc = ensure_panel(ctx, fp4);
	if(rc) return rc;
	long long slab = ctx->x_chunks;
	/* a rank of a partitioned run spends a larger share of its step expanding (it reads O(n / sqrt(world)) rows
	 * for 1 / world of the tiles): cut the K axis so that the expansion of slab s+1 runs under the GEMM of slab s.
	 * On one GPU the cut does not pay (measured 334 vs 331 ms at 10k, 7.5 vs 6.5 ms at 1k). */
	const int min_slabs = ctx->min_slabs > 0 ? ctx->min_slabs : (ctx->world > 1 ? 4 : 1);
	if(!ctx->feed_seqs && min_slabs > 1 && ctx->chunks >= 4096) {
		long long want = (ctx->chunks + min_slabs - 1) / min_slabs;
		const long long cap = ctx->x_bytes >= 2 * ctx->x_buf_bytes ? ctx->x_chunks : ctx->x_chunks / 2;
		if(want < slab && cap >= 1) slab = want < cap ? want : cap;
		if(fp4 && (slab & 1)) slab = slab > 1 ? slab - 1 : 2;
	}
	if(ctx->feed_seqs) {
		/* host rows are streamed slab by slab: shorter slabs shorten the pipeline fill (the first
		 * slab's upload is the only one not hidden behind a GEMM) */
		const int feed_slabs = ctx->dbg_feed_slabs > 0 ? ctx->dbg_feed_slabs : 16;
		long long want = (ctx->chunks + feed_slabs - 1) / feed_slabs;
		/* the GEMM of the LAST slab is the tail nothing hides: a short alignment (a member's slice of a K-split group)
		 * is cut finer -- down to 512 chunks, where a slab's GEMM is still ~1 ms of full waves */
		long long floor_chunks = ctx->stream_min_chunks / 2 < 512 ? ctx->stream_min_chunks / 2 : 512;
		if(floor_chunks < 16) floor_chunks = 16;
		if(want < floor_chunks) want = floor_chunks;
		const long long cap = ctx->x_bytes >= 2 * ctx->x_buf_bytes ? ctx->x_chunks : ctx->x_chunks / 2;
		if(want < slab && cap >= 1) slab = want < cap ? want : cap;
		if(fp4 && (slab & 1)) slab = slab > 1 ? slab - 1 : 2;
	}
	const int nslabs = (int) ((ctx->chunks + slab - 1) / slab);
	/* two slab buffers: the allocation's two halves, or -- when the whole panel fits in a single
	 * buffer and shorter slabs are streamed -- that buffer cut in two */
	const size_t per_chunk_b = (size_t) ctx->n_pad * (fp4 ? 256 : 512);
	size_t buf_off[2] = {0, ctx->x_buf_bytes};
	if(ctx->x_bytes < 2 * ctx->x_buf_bytes) {
		buf_off[1] = (((size_t) ctx->x_chunks / 2) & ~(size_t) 1) * per_chunk_b;
		if(nslabs > 1 && (size_t) slab * per_chunk_b > buf_off[1]) {
			set_err(ctx, "internal: slab of %lld chunks does not fit half of the operand panel", slab);
			return CCG_ERR_ARG;
		}
	}

	UmmaParams p;
	memset(&p, 0, sizeof(p));
	p.ntiles = (int) cnt;
	p.tiles = ctx->d_tiles;
	p.C_S = acc_S;
	p.C_I = acc_I;
	p.ldc = ctx->n_pad;
	p.fp4 = fp4 ? 1 : 0;
	p.no_mask_items = (fp4 && !ctx->pair_mode) ? 1 : 0;   /* two-plane store: k_finalize_umma uses the constant */
	long long i_const_fp4 = 0;
	CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	/* The expansion of slab s+1 runs beside the GEMM of slab s.  It must not start before every
	 * (persistent) GEMM CTA holds its SM, or expansion blocks can fill an SM and keep a GEMM CTA
	 * out for milliseconds while all others wait for it in lock-step: the GEMM CTAs count
	 * themselves in d_resident (monotonic over the run) and the aux stream waits for the count. */
	if(ctx->feed_seqs) {
		CK(ctx, cudaEventRecord(ctx->ev_main, ctx->stream));
		for(int q = 0; q < 2; ++q) CK(ctx, cudaStreamWaitEvent(ctx->copy_stream[q], ctx->ev_main, 0));
	}
	/* host rows streamed into the e2m1 panel: expanded by the copy streams, no plane store (CCG_FEED_PLANES=1: old path) */
	const bool direct_feed = ctx->feed_seqs && fp4 && !ctx->dbg_feed_planes;
	if(direct_feed) {
		std::vector<int> unfed;
		for(int b = 0; b < ctx->n_pad / 128; ++b) {
			if(!ctx->need[b]) continue;
			for(int k = b * 128; k < (b + 1) * 128; ++k)
				if(k >= ctx->n || !ctx->feed_seqs[k] || (ctx->feed_masks && !ctx->feed_masks[k])) unfed.push_back(k);
		}
		ctx->n_unfed = (int) unfed.size();
		if(ctx->unfed_cap < unfed.size()) {
			CK(ctx, cudaStreamSynchronize(ctx->stream));
			cudaFree(ctx->d_unfed);
			ctx->d_unfed = 0;
			ctx->unfed_cap = 0;
			CK(ctx, cudaMalloc(&ctx->d_unfed, unfed.size() * sizeof(int)));
			ctx->unfed_cap = unfed.size();
		}
		if(!unfed.empty()) {
			CK(ctx, cudaMemcpyAsync(ctx->d_unfed, unfed.data(), unfed.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
			CK(ctx, cudaStreamSynchronize(ctx->stream));                 /* the host list dies with this block */
		}
		ctx->planes_stale = 1;
	}
	const bool gate = ctx->fn_wait_value && ctx->d_resident && nslabs > 1;
	if(gate) CK(ctx, cudaMemsetAsync(ctx->d_resident, 0, sizeof(unsigned), ctx->stream));
	p.resident = gate ? ctx->d_resident : 0;
	unsigned resident_target = 0;
	CK(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
	CK(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
	CK(ctx, cudaEventRecord(ctx->ev_phase[0], ctx->aux_stream));
	for(int s = 0; s < nslabs; ++s) {
		const int b = s & 1;
		const int chunk0 = (int) (s * slab);
		int nch = ctx->chunks - chunk0;
		if(nch > slab) nch = (int) slab;
		/* K units of this slab: chunks, or chunk pairs for the e2m1 panel (a 128-byte channel row = 256 bases) */
		const int nu = fp4 ? (nch + 1) / 2 : nch;
		i_const_fp4 += (long long) nu * 2 * CCG_CHUNK_BASES;
		p.slab_chunks = nu;
		p.row_base = (int) (buf_off[b] / 128);
		if(ctx->dbg_umma1) {
			p.single = 1;
			p.ntiles = (int) (2 * cnt);
			p.tiles = ctx->d_tiles + cnt;
		}
		p.kslices = choose_split((long long) (p.single ? ctx->sm_count : ccg_umma_pair_slots(ctx)), (long long) p.ntiles, nu, fp4 ? 8 : 16, 512,
		                        ctx->feed_seqs ? 0 : 1);
		if(ctx->dbg_kslices > 0) p.kslices = ctx->dbg_kslices;
		/* f32 accumulators: an S item adds at most 3 * 256 per chunk pair and must stay below 2^24 (applied after
		 * the experiment override as well: exactness is not a tuning knob) */
		if(fp4 && (nu + p.kslices - 1) / p.kslices > CCG_FP4_MAX_PAIRS) p.kslices = (nu + CCG_FP4_MAX_PAIRS - 1) / CCG_FP4_MAX_PAIRS;
		p.chunks_per_slice = (nu + p.kslices - 1) / p.kslices;
		while(p.kslices > 1 && (long long) (p.kslices - 1) * p.chunks_per_slice >= nu) --p.kslices;
		/* who writes the slab's panel: the aux stream (expansion from the plane store) or -- host rows streamed into the
		 * e2m1 panel -- the two copy streams themselves, right behind their uploads */
		cudaStream_t writers[2] = {ctx->aux_stream, 0};
		int nwriters = 1;
		if(direct_feed) { writers[0] = ctx->copy_stream[0]; writers[1] = ctx->copy_stream[1]; nwriters = 2; }
		for(int wq = 0; wq < nwriters; ++wq) {
			/* buffer b is free once the GEMM of slab s-2 has read it */
			if(s >= 2) CK(ctx, cudaStreamWaitEvent(writers[wq], ctx->ev_g[b], 0));
			if(ctx->dbg_serial && s >= 1) CK(ctx, cudaStreamWaitEvent(writers[wq], ctx->ev_g[b ^ 1], 0));
			if(gate && s >= 1) {
				typedef CUresult (*WaitValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
				/* ev_launch: the GEMM of slab s-1 has been handed to the device queue */
				CK(ctx, cudaStreamWaitEvent(writers[wq], ctx->ev_launch, 0));
				CUresult r = ((WaitValueFn) ctx->fn_wait_value)((CUstream) writers[wq], (CUdeviceptr) ctx->d_resident,
				                                                resident_target, CU_STREAM_WAIT_VALUE_GEQ);
				if(r != CUDA_SUCCESS) {
					set_err(ctx, "cuStreamWaitValue32 failed with CUresult %d", (int) r);
					return CCG_ERR_CUDA;
				}
			}
		}
		if(ctx->feed_seqs) {
			rc = feed_slab(ctx, chunk0, nch, direct_feed ? ctx->d_X + buf_off[b] : 0, nu);
			if(rc) return rc;
			for(int q = 0; q < 2; ++q) CK(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_up[q], 0));
		}
		if(direct_feed) { /* the panel of this slab is complete once both copy streams are through (ev_up, waited for above) */ }
		else if(fp4) CK(ctx, ccg_launch_expand_fp4(ctx, ctx->aux_stream, ctx->d_X + buf_off[b], chunk0, nu, nslabs > 1 && !ctx->dbg_serial));
		else CK(ctx, ccg_launch_expand(ctx, ctx->aux_stream, ctx->d_X + buf_off[b], chunk0, nch, nslabs > 1 && !ctx->dbg_serial));
		CK(ctx, cudaEventRecord(ctx->ev_x[b], ctx->aux_stream));
		if(s == nslabs - 1) CK(ctx, cudaEventRecord(ctx->ev_phase[1], ctx->aux_stream));
		CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_x[b], 0));
		if(s == 0) CK(ctx, cudaEventRecord(ctx->ev_phase[2], ctx->stream));
		if(gate) CK(ctx, cudaEventRecord(ctx->ev_launch, ctx->stream));     /* everything before the GEMM of slab s */
		CK(ctx, ccg_launch_umma(ctx, p));
		resident_target += (unsigned) ctx->last_gemm_ctas;
		CK(ctx, cudaEventRecord(ctx->ev_g[b], ctx->stream));
	}
	CK(ctx, cudaEventRecord(ctx->ev_phase[3], ctx->stream));
	ctx->phase_valid = 1;
	/* shared-mask mode: every position of every chunk (pair) counts as included in the raw product */
	const int i_const = fp4 ? (int) i_const_fp4 : ctx->chunks * CCG_CHUNK_BASES;
	ctx->last_i_const = i_const;
	p.ntiles = (int) cnt;
	p.tiles = ctx->d_tiles;
	if(group) {
		/* barrier with the peers, then the reduce-scatter over NVLink fused with the epilogue of this member's rows */
		rc = ccg_group_finalize(ctx, ep, i_const, 0, ctx->n_pad);
		if(rc) return rc;
	} else CK(ctx, ccg_launch_finalize_umma(ctx, p, ep, i_const));
	CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	ctx->ev_valid = 1;
	snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k_pairdist_umma%s tiles=%d kslices=%d slabs=%d%s",
	         ctx->dbg_umma1 ? "" : (fp4 ? "2<mxf4>" : "2<i8>"), p.ntiles, p.kslices, nslabs, group ? " +k_finalize_group" : "");
	return CCG_OK;
}

/* K-split group whose n x n accumulators exceed the peer window (BASELINE configs[3]: 100,000 samples): the member's
 * whole slice of every sample is expanded once, then the lower triangle is run in windows of whole macro-tile rows;
 * each window is accumulated into one of the two window buffers, reduced over the members and finalised
 * (ccg_group_finalize) while the next window's GEMM is already queued behind it.  The sequence shards never move. */
static int run_umma_windows(ccg_ctx *ctx, const EpilogueParams &ep, int win_rows) {
	if(ctx->use_i8 || ctx->dbg_umma1 || ctx->feed_seqs) {
		set_err(ctx, "a windowed K-split run uses the e2m1 panel kernel on samples already on the device");
		return CCG_ERR_UNSUPPORTED;
	}
	const int nwin = (ctx->n_pad + win_rows - 1) / win_rows;
	const int tiles_per_win = win_rows / CCG_UMMA_BM;
	std::vector<int2> tiles;
	for_each_ctx_tile(ctx, [&](int tm, int tn) {
		int2 t;
		t.x = tm;
		t.y = tn;
		tiles.push_back(t);
	});
	ctx->last_ntiles = (int) tiles.size();
	ctx->last_kernel_kind = CCG_KERNEL_UMMA;
	if(tiles.empty()) return CCG_OK;
	/* window by window, Z-order inside a window (neighbouring CTA pairs share operand rows in L2) */
	std::stable_sort(tiles.begin(), tiles.end(), [&](const int2 &a, const int2 &b) { return a.x / tiles_per_win < b.x / tiles_per_win; });
	std::vector<size_t> start((size_t) nwin + 1, tiles.size());
	for(size_t k = tiles.size(); k-- > 0;) start[(size_t) (tiles[k].x / tiles_per_win)] = k;
	for(int w = nwin - 1; w >= 0; --w)
		if(start[(size_t) w] > start[(size_t) w + 1]) start[(size_t) w] = start[(size_t) w + 1];
	int rc = ensure_tiles(ctx, tiles.data(), tiles.size());
	if(rc) return rc;
	rc = ensure_panel(ctx, true);
	if(rc) return rc;
	if(ctx->x_chunks < ctx->chunks) {
		set_err(ctx, "windowed K-split run: the operand panel of this GPU's slice (%d slots x %d chunks) must fit the device in one "
		        "piece; use more GPUs", ctx->n_pad, ctx->chunks);
		return CCG_ERR_NOMEM;
	}
	const int nu = (ctx->chunks + 1) / 2;                       /* chunk pairs */
	const int i_const = nu * 2 * CCG_CHUNK_BASES;
	ctx->last_i_const = i_const;
	CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	CK(ctx, cudaEventRecord(ctx->ev_phase[0], ctx->stream));
	CK(ctx, ccg_launch_expand_fp4(ctx, ctx->stream, ctx->d_X, 0, nu, 0));
	CK(ctx, cudaEventRecord(ctx->ev_phase[1], ctx->stream));
	CK(ctx, cudaEventRecord(ctx->ev_phase[2], ctx->stream));
	UmmaParams p;
	memset(&p, 0, sizeof(p));
	p.ldc = ctx->n_pad;
	p.fp4 = 1;
	p.no_mask_items = ctx->pair_mode ? 0 : 1;
	p.slab_chunks = nu;
	p.row_base = 0;
	int kslices_seen = 0;
	for(int w = 0; w < nwin; ++w) {
		const int row0 = w * win_rows, row1 = row0 + win_rows < ctx->n_pad ? row0 + win_rows : ctx->n_pad;
		const size_t ntl = start[(size_t) w + 1] - start[(size_t) w];
		if(ntl == 0) continue;                                  /* rows past the last sample: the same on every member */
		void *clear = 0;
		size_t bytes = 0;
		rc = ccg_group_accumulators(ctx, row0, &p.C_S, &p.C_I, &clear, &bytes);
		if(rc) return rc;
		CK(ctx, cudaMemsetAsync(clear, 0, bytes, ctx->stream));
		p.tiles = ctx->d_tiles + start[(size_t) w];
		p.ntiles = (int) ntl;
		p.kslices = choose_split((long long) ccg_umma_pair_slots(ctx), (long long) ntl, nu, 8, 512, 1);
		if(ctx->dbg_kslices > 0) p.kslices = ctx->dbg_kslices;
		if((nu + p.kslices - 1) / p.kslices > CCG_FP4_MAX_PAIRS) p.kslices = (nu + CCG_FP4_MAX_PAIRS - 1) / CCG_FP4_MAX_PAIRS;
		p.chunks_per_slice = (nu + p.kslices - 1) / p.kslices;
		while(p.kslices > 1 && (long long) (p.kslices - 1) * p.chunks_per_slice >= nu) --p.kslices;
		if(p.kslices > kslices_seen) kslices_seen = p.kslices;
		CK(ctx, ccg_launch_umma(ctx, p));
		rc = ccg_group_finalize(ctx, ep, i_const, row0, row1);
		if(rc) return rc;
	}
	CK(ctx, cudaEventRecord(ctx->ev_phase[3], ctx->stream));
	ctx->phase_valid = 1;
	CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	ctx->ev_valid = 1;
	snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k_pairdist_umma2<mxf4> tiles=%d kslices=%d slabs=1 windows=%d +k_finalize_group",
	         (int) tiles.size(), kslices_seen, nwin);
	return CCG_OK;
}

/* tensor path with the operand expansion fused into the GEMM: no panel, no slabs */
static int run_fused(ccg_ctx *ctx, const EpilogueParams &ep) {
	size_t cnt = 0;
	int rc = upload_macro_tiles(ctx, &cnt);
	if(rc) return rc;
	ctx->last_ntiles = (int) cnt;
	ctx->last_kernel_kind = CCG_KERNEL_FUSED;
	if(cnt == 0) return CCG_OK;

	size_t c_bytes = (size_t) 2 * ctx->n_pad * ctx->n_pad * sizeof(int);
	if(ctx->c_bytes < c_bytes) {
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_C);
		ctx->d_C = 0;
		ctx->c_bytes = 0;
		if(cudaMalloc(&ctx->d_C, c_bytes) != cudaSuccess) {
			set_err(ctx, "cudaMalloc of %zu bytes for the int32 accumulators failed", c_bytes);
			return CCG_ERR_NOMEM;
		}
		ctx->c_bytes = c_bytes;
	}
	CK(ctx, cudaMemsetAsync(ctx->d_C, 0, c_bytes, ctx->stream));

	UmmaParams p;
	memset(&p, 0, sizeof(p));
	p.ntiles = (int) cnt;
	p.tiles = ctx->d_tiles;
	p.C_S = ctx->d_C;
	p.C_I = ctx->d_C + (size_t) ctx->n_pad * ctx->n_pad;
	p.ldc = ctx->n_pad;
	p.slab_chunks = ctx->chunks;
	/* the fused kernel works on the 128 x 256 halves of the macro tiles */
	p.ntiles = (int) (2 * cnt);
	p.tiles = ctx->d_tiles + cnt;
	p.kslices = choose_split((long long) ctx->sm_count, (long long) p.ntiles, ctx->chunks, 16, 1024);
	p.chunks_per_slice = (ctx->chunks + p.kslices - 1) / p.kslices;
	while(p.kslices > 1 && (long long) (p.kslices - 1) * p.chunks_per_slice >= ctx->chunks) --p.kslices;
	CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	CK(ctx, cudaEventRecord(ctx->ev_phase[2], ctx->stream));
	CK(ctx, ccg_launch_fused(ctx, p));
	CK(ctx, cudaEventRecord(ctx->ev_phase[3], ctx->stream));
	ctx->phase_valid = 1;
	p.ntiles = (int) cnt;
	p.tiles = ctx->d_tiles;
	ctx->last_i_const = ctx->chunks * CCG_CHUNK_BASES;
	CK(ctx, ccg_launch_finalize_umma(ctx, p, ep, ctx->chunks * CCG_CHUNK_BASES));
	CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	ctx->ev_valid = 1;
	snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k_pairdist_fused tiles=%d kslices=%d", p.ntiles, p.kslices);
	return CCG_OK;
}

/* compaction map: included samples in input order (fsacmpthrd.c:305-329) */
static int compute_ranks(ccg_ctx *ctx, const unsigned char *include) {
	int Dn = 0;
	for(int i = 0; i < ctx->n_pad; ++i) {
		int inc = i < ctx->n && ctx->present[i] && (!include || include[i]);
		ctx->h_rank[i] = inc ? Dn++ : -1;
	}
	return Dn;
}

static int run_common(ccg_ctx *ctx, int mode, const unsigned char *include, unsigned norm, unsigned minLength,
                      double minCov, int elem_size, double byteScale, void *d_D, void *d_N, int *Dn_out) {
	if(!ctx || !ctx->d_planes) return CCG_ERR_ARG;
	if(elem_size != 8 && elem_size != 4 && elem_size != 2 && elem_size != 1) return CCG_ERR_ARG;
	if((mode == 0) != (ctx->pair_mode == 1) && !(mode == 1 && ctx->pair_mode && ctx->global_applied)) {
		set_err(ctx, "run mode does not match ccg_set_problem(pair_mode=%d)", ctx->pair_mode);
		return CCG_ERR_ARG;
	}
	CK(ctx, cudaSetDevice(ctx->device));
	const bool group = ctx->grp_world > 1;
	if(group) {
		if(ctx->proxi || ctx->row_slot1 || ctx->world > 1 || ctx->win_on) {
			set_err(ctx, "a member of a K-split group runs plain all-vs-all comparisons only (no -P, row run, tile partition or window)");
			return CCG_ERR_UNSUPPORTED;
		}
		if(ctx->grp_total_len < ctx->len || (mode == 1 && !ctx->grp_global_inc && ctx->global_inc)) {
			set_err(ctx, "K-split group: ccg_group_set_alignment(total length%s) must precede the run", mode == 1 ? ", inclusion count of the whole global mask" : "");
			return CCG_ERR_ARG;
		}
	}
	if(ctx->remask_pending) {
		/* -y: the code planes still hold the bases of the methylation sites (k_motif.cu) */
		CK(ctx, ccg_launch_remask_all(ctx));
		ctx->remask_pending = 0;
	}
	if(mode == 1 && ctx->global_pending) {
		CK(ctx, ccg_launch_apply_global_mask(ctx));
		ctx->global_pending = 0;
	}

	if(ctx->planes_stale && !ctx->feed_seqs) {
		int rc = planes_usable(ctx);
		if(rc) return rc;
	}
	const int Dn = compute_ranks(ctx, include);
	ctx->last_Dn = Dn;
	if(Dn_out) *Dn_out = Dn;
	ctx->last_ntiles = 0;
	if(Dn < 2) return CCG_OK;
	for(int b = 0; b < ctx->n_pad / 128; ++b) {
		if(!ctx->need[b] || ctx->have[b]) continue;
		for(int i = b * 128; i < (b + 1) * 128; ++i)
			if(ctx->h_rank[i] >= 0) {
				set_err(ctx, "row block %d is needed by rank %d/%d but was not uploaded: call ccg_set_partition before "
				        "the ccg_put_* calls", b, ctx->rank, ctx->world);
				return CCG_ERR_ARG;
			}
	}
	CK(ctx, cudaMemcpyAsync(ctx->d_rank, ctx->h_rank, (size_t) ctx->n_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));

	EpilogueParams ep;
	memset(&ep, 0, sizeof(ep));
	ep.mode = mode;
	ep.elem_size = elem_size;
	ep.norm = norm;
	ep.byteScale = byteScale;
	ep.D = d_D;
	ep.N = mode == 0 ? d_N : 0;
	ep.rank = ctx->d_rank;
	ep.row_base = ctx->ep_row_base;
	if(ctx->row_slot1) ep.row_plus1 = ctx->h_rank[ctx->row_slot1 - 1] + 1;
	if(mode == 0) {
		/* fsacmpthrd.c:292: minLength = minLength < minCov * len ? minCov * len : minLength */
		const long long gate_len = group ? ctx->grp_total_len : (long long) ctx->len;       /* the whole alignment */
		if(minLength < minCov * gate_len) minLength = (unsigned) (minCov * gate_len);
		ep.minLength = minLength;
		ep.nFactor = 1.0;
	} else {
		/* fsacmpthrd.c:171-176 */
		double nFactor = 1.0;
		if(norm) { nFactor = norm; nFactor /= (int) (group ? ctx->grp_global_inc : ctx->global_inc); }
		ep.nFactor = nFactor;
	}
	/* AUTO: the tensor-core kernel wins once there is enough work to fill the machine */
	/* measured on B200 (profiles/): the tensor kernel runs the contraction ~5x faster than the
	 * XU-pipe-bound POPC kernel but pays a fixed operand-expansion pass; small problems stay on POPC */
	if(mode == 0 && ctx->proxi) {
		NEED_PLANES(ctx);
		if(ctx->feed_seqs) {
			set_err(ctx, "internal: host rows cannot be streamed into a run with proximity masking");
			return CCG_ERR_ARG;
		}
		return run_proxi(ctx, ep);
	}
	int kind = ctx->kernel_choice;
	if(kind == CCG_KERNEL_AUTO)
		kind = (Dn >= 192 && ctx->chunks >= 64) ? CCG_KERNEL_UMMA : CCG_KERNEL_POPC;
	if(ctx->feed_seqs || group) kind = CCG_KERNEL_UMMA;      /* the caller streams host rows into the tensor path's K slabs */
	/* int32 sums of the tensor path: |S| <= 3 (len + 255) must stay below 2^31; the u32 counters of the POPC path
	 * hold any alignment an int length can describe */
	if(kind != CCG_KERNEL_POPC && 3.0 * ((double) ctx->len + 256.0) >= 2147483648.0) {
		if(ctx->kernel_choice != CCG_KERNEL_AUTO || ctx->feed_seqs || group) {
			set_err(ctx, "alignment of %d bases: the tensor path's int32 sums hold 3 x length only up to 715 Mbp", ctx->len);
			return CCG_ERR_UNSUPPORTED;
		}
		kind = CCG_KERNEL_POPC;
	}
	if(ctx->bor_pending && !(kind == CCG_KERNEL_UMMA && !ctx->use_i8 && !ctx->dbg_umma1 && !ctx->feed_seqs)) {
		int rc = materialize_borrowed(ctx);
		if(rc) return rc;
	}
	if(kind == CCG_KERNEL_FUSED) return run_fused(ctx, ep);
	return kind == CCG_KERNEL_UMMA ? run_umma(ctx, ep) : run_popc(ctx, ep);
}

/* Host matrices of a K-split member: the member owns the row blocks b with b % world == rank, its device result
 * buffers hold just those rows back to back (row_base[r] = offset of compact row r), and every owned block is one
 * contiguous run of cells of the packed host matrix -- one copy per block and matrix.  The members of a group write
 * disjoint cells of the same host matrices. */
/* K-split member: row_base[r] = offset of compact row r in a buffer that holds just the rows this member owns,
 * back to back (h_row_base on the host, d_row_base for the epilogue).  Needs the rank map of this include[]. */
static int group_row_bases(ccg_ctx *ctx, const unsigned char *include, long long *owned_cells) {
	const int world = ctx->grp_world, rank = ctx->grp_rank;
	if(ctx->row_base_cap < (size_t) ctx->n_pad) {
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_row_base); ctx->d_row_base = 0;
		free(ctx->h_row_base); ctx->h_row_base = 0;
		ctx->row_base_cap = 0;
		ctx->h_row_base = (long long *) malloc((size_t) ctx->n_pad * sizeof(long long));
		if(!ctx->h_row_base) return CCG_ERR_NOMEM;
		CK(ctx, cudaMalloc(&ctx->d_row_base, (size_t) ctx->n_pad * sizeof(long long)));
		ctx->row_base_cap = (size_t) ctx->n_pad;
	}
	const int Dn = compute_ranks(ctx, include);
	long long off = 0;
	for(int i = 0; i < ctx->n; ++i) {
		const int r = ctx->h_rank[i];
		if(r < 0) continue;
		if((i / CCG_GROUP_ROW_BLOCK) % world == rank) { ctx->h_row_base[r] = off; off += r; }
		else ctx->h_row_base[r] = -1;
	}
	if(Dn > 0)
		CK(ctx, cudaMemcpyAsync(ctx->d_row_base, ctx->h_row_base, (size_t) Dn * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
	if(owned_cells) *owned_cells = off;
	return CCG_OK;
}

/* Host matrices of a K-split member: the member owns the row blocks b with b % world == rank, its device result
 * buffers hold just those rows back to back, and every owned block is one contiguous run of cells of the packed
 * host matrix -- one copy per block and matrix; the members of a group write disjoint cells of the same host
 * matrices.  With ccg_group_set_output(compact) the host buffers hold only the owned rows as well: one copy. */
static int run_to_host_group(ccg_ctx *ctx, int mode, const unsigned char *include, unsigned norm, unsigned minLength,
                             double minCov, int elem_size, double byteScale, void *D, void *N, int *Dn_out) {
	const int world = ctx->grp_world, rank = ctx->grp_rank;
	const long long own_max = ccg_group_cells(ctx->n, rank, world);
	const size_t bytes = (size_t) (own_max > 0 ? own_max : 1) * 8;
	if(ctx->out_bytes < bytes || !ctx->d_out_D) {
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_out_D); ctx->d_out_D = 0;
		cudaFree(ctx->d_out_N); ctx->d_out_N = 0;
		ctx->out_bytes = 0;
		if(cudaMalloc(&ctx->d_out_D, bytes) != cudaSuccess || cudaMalloc(&ctx->d_out_N, bytes) != cudaSuccess) {
			set_err(ctx, "cudaMalloc of 2 x %zu result bytes failed", bytes);
			return CCG_ERR_NOMEM;
		}
		ctx->out_bytes = bytes;
	}
	long long owned = 0;
	int rc = group_row_bases(ctx, include, &owned);
	if(rc) return rc;
	ctx->ep_row_base = ctx->d_row_base;
	int Dn = 0;
	rc = run_common(ctx, mode, include, norm, minLength, minCov, elem_size, byteScale, ctx->d_out_D, N ? ctx->d_out_N : 0, &Dn);
	ctx->ep_row_base = 0;
	if(rc) return rc;
	if(Dn_out) *Dn_out = Dn;
	if(Dn > 1 && ctx->grp_compact && owned > 0) {
		CK(ctx, cudaMemcpyAsync(D, ctx->d_out_D, (size_t) owned * elem_size, cudaMemcpyDeviceToHost, ctx->stream));
		if(N) CK(ctx, cudaMemcpyAsync(N, ctx->d_out_N, (size_t) owned * elem_size, cudaMemcpyDeviceToHost, ctx->stream));
	} else if(Dn > 1) {
		for(int b = rank; (long long) b * CCG_GROUP_ROW_BLOCK < ctx->n; b += world) {
			const int i0 = b * CCG_GROUP_ROW_BLOCK, i1 = i0 + CCG_GROUP_ROW_BLOCK < ctx->n ? i0 + CCG_GROUP_ROW_BLOCK : ctx->n;
			long long r0 = -1, r1 = -1;                           /* compact rows of the block: consecutive */
			for(int i = i0; i < i1; ++i)
				if(ctx->h_rank[i] >= 0) { if(r0 < 0) r0 = ctx->h_rank[i]; r1 = ctx->h_rank[i] + 1; }
			if(r0 < 0) continue;
			const size_t lo = (size_t) (r0 * (r0 - 1) / 2) * elem_size, len = (size_t) (r1 * (r1 - 1) / 2 - r0 * (r0 - 1) / 2) * elem_size;
			const size_t src = (size_t) ctx->h_row_base[r0] * elem_size;
			if(!len) continue;
			CK(ctx, cudaMemcpyAsync((char *) D + lo, (char *) ctx->d_out_D + src, len, cudaMemcpyDeviceToHost, ctx->stream));
			if(N) CK(ctx, cudaMemcpyAsync((char *) N + lo, (char *) ctx->d_out_N + src, len, cudaMemcpyDeviceToHost, ctx->stream));
		}
	}
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return CCG_OK;
}

/* device-resident run of a K-split member with compact outputs (ccg_group_set_output) */
static int run_dev_group_compact(ccg_ctx *ctx, int mode, const unsigned char *include, unsigned norm, unsigned minLength,
                                 double minCov, int elem_size, double byteScale, void *d_D, void *d_N, int *Dn) {
	if(!ctx->d_planes) return CCG_ERR_ARG;
	CK(ctx, cudaSetDevice(ctx->device));
	int rc = group_row_bases(ctx, include, 0);
	if(rc) return rc;
	ctx->ep_row_base = ctx->d_row_base;
	rc = run_common(ctx, mode, include, norm, minLength, minCov, elem_size, byteScale, d_D, d_N, Dn);
	ctx->ep_row_base = 0;
	return rc;
}

static int run_to_host(ccg_ctx *ctx, int mode, const unsigned char *include, unsigned norm, unsigned minLength,
                       double minCov, int elem_size, double byteScale, void *D, void *N, int *Dn_out) {
	if(!ctx || !D) return CCG_ERR_ARG;
	if(!ctx->d_planes || (elem_size != 8 && elem_size != 4 && elem_size != 2 && elem_size != 1)) return CCG_ERR_ARG;
	CK(ctx, cudaSetDevice(ctx->device));
	if(ctx->grp_world > 1) return run_to_host_group(ctx, mode, include, norm, minLength, minCov, elem_size, byteScale, D, N, Dn_out);
	/* worst case Dn = n */
	size_t max_cells = ctx->n > 1 ? (size_t) ctx->n * (ctx->n - 1) / 2 : 0;
	size_t bytes = max_cells * 8;
	if(ctx->out_bytes < bytes || !ctx->d_out_D) {
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_out_D); ctx->d_out_D = 0;
		cudaFree(ctx->d_out_N); ctx->d_out_N = 0;
		ctx->out_bytes = 0;
		if(bytes) {
			if(cudaMalloc(&ctx->d_out_D, bytes) != cudaSuccess || cudaMalloc(&ctx->d_out_N, bytes) != cudaSuccess) {
				set_err(ctx, "cudaMalloc of 2 x %zu result bytes failed", bytes);
				return CCG_ERR_NOMEM;
			}
		}
		ctx->out_bytes = bytes;
	}
	int Dn = 0;
	if(ctx->world > 1 && bytes) {
		/* cells owned by other ranks read back as zero */
		CK(ctx, cudaMemsetAsync(ctx->d_out_D, 0, bytes, ctx->stream));
		CK(ctx, cudaMemsetAsync(ctx->d_out_N, 0, bytes, ctx->stream));
	}
	int rc = run_common(ctx, mode, include, norm, minLength, minCov, elem_size, byteScale, ctx->d_out_D,
	                    N ? ctx->d_out_N : 0, &Dn);
	if(rc) return rc;
	if(Dn_out) *Dn_out = Dn;
	if(Dn > 1) {
		size_t cells = (size_t) Dn * (Dn - 1) / 2;
		CK(ctx, cudaMemcpyAsync(D, ctx->d_out_D, cells * elem_size, cudaMemcpyDeviceToHost, ctx->stream));
		if(N) CK(ctx, cudaMemcpyAsync(N, ctx->d_out_N, cells * elem_size, cudaMemcpyDeviceToHost, ctx->stream));
	}
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return CCG_OK;
}

extern "C" int ccg_run_pair(ccg_ctx *ctx, const unsigned char *include, unsigned norm, unsigned minLength, double minCov,
                            int elem_size, double byteScale, void *D, void *N, int *Dn) {
	if(ctx && ctx->multi) return ccg_multi_run(ctx, 1, include, norm, minLength, minCov, elem_size, byteScale, D, N, Dn, 0);
	return run_to_host(ctx, 0, include, norm, minLength, minCov, elem_size, byteScale, D, N, Dn);
}

extern "C" int ccg_run_global(ccg_ctx *ctx, const unsigned char *include, unsigned norm, int elem_size, double byteScale,
                              void *D, int *Dn, unsigned *global_inc) {
	if(ctx && ctx->multi) return ccg_multi_run(ctx, 0, include, norm, 0, 0.0, elem_size, byteScale, D, 0, Dn, global_inc);
	if(ctx && global_inc) *global_inc = ctx->grp_world > 1 ? ctx->grp_global_inc : ctx->global_inc;
	return run_to_host(ctx, 1, include, norm, 0, 0.0, elem_size, byteScale, D, 0, Dn);
}

extern "C" int ccg_run_pair_dev(ccg_ctx *ctx, const unsigned char *include, unsigned norm, unsigned minLength,
                                double minCov, int elem_size, double byteScale, void *d_D, void *d_N, int *Dn) {
	if(!d_D) return CCG_ERR_ARG;
	CCG_MULTI_SOLO(ctx, "a run into device buffers", ccg_run_pair_dev(m0, include, norm, minLength, minCov, elem_size, byteScale, d_D, d_N, Dn));
	if(ctx && ctx->grp_world > 1 && ctx->grp_compact)
		return run_dev_group_compact(ctx, 0, include, norm, minLength, minCov, elem_size, byteScale, d_D, d_N, Dn);
	return run_common(ctx, 0, include, norm, minLength, minCov, elem_size, byteScale, d_D, d_N, Dn);
}

extern "C" int ccg_run_global_dev(ccg_ctx *ctx, const unsigned char *include, unsigned norm, int elem_size,
                                  double byteScale, void *d_D, int *Dn, unsigned *global_inc) {
	if(!d_D) return CCG_ERR_ARG;
	CCG_MULTI_SOLO(ctx, "a run into device buffers", ccg_run_global_dev(m0, include, norm, elem_size, byteScale, d_D, Dn, global_inc));
	if(ctx && global_inc) *global_inc = ctx->grp_world > 1 ? ctx->grp_global_inc : ctx->global_inc;
	if(ctx && ctx->grp_world > 1 && ctx->grp_compact)
		return run_dev_group_compact(ctx, 1, include, norm, 0, 0.0, elem_size, byteScale, d_D, 0, Dn);
	return run_common(ctx, 1, include, norm, 0, 0.0, elem_size, byteScale, d_D, 0, Dn);
}

/* cmpFsaRowThrd (fsacmpthrd.c:482-580): the row of one sample against every sample uploaded into a lower slot.
 * The pair kernels run on the macro-tile row that holds the slot; the epilogue keeps that one row. */
extern "C" int ccg_run_row(ccg_ctx *ctx, int row_slot, unsigned norm, unsigned minLength, double minCov, double *D, double *N,
                           int *cols_out) {
	CCG_MULTI_SOLO(ctx, "a row run (-a)", ccg_run_row(m0, row_slot, norm, minLength, minCov, D, N, cols_out));
	if(!ctx || !ctx->d_planes || !ctx->pair_mode || !D || row_slot < 0 || row_slot >= ctx->n || !ctx->present[row_slot])
		return CCG_ERR_ARG;
	if(ctx->world > 1) {
		set_err(ctx, "ccg_run_row runs on one device");
		return CCG_ERR_UNSUPPORTED;
	}
	CK(ctx, cudaSetDevice(ctx->device));
	NEED_PLANES(ctx);
	unsigned char *use = (unsigned char *) calloc((size_t) ctx->n, 1);
	if(!use) return CCG_ERR_NOMEM;
	int cols = 0;
	for(int j = 0; j < row_slot; ++j) cols += use[j] = ctx->present[j];
	use[row_slot] = 1;
	if(cols_out) *cols_out = cols;
	if(cols == 0) { free(use); return CCG_OK; }
	const size_t bytes = (size_t) cols * sizeof(double);
	if(ctx->out_bytes < bytes || !ctx->d_out_D) {
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_out_D); ctx->d_out_D = 0;
		cudaFree(ctx->d_out_N); ctx->d_out_N = 0;
		ctx->out_bytes = 0;
		if(cudaMalloc(&ctx->d_out_D, bytes) != cudaSuccess || cudaMalloc(&ctx->d_out_N, bytes) != cudaSuccess) {
			free(use);
			set_err(ctx, "cudaMalloc of 2 x %zu result bytes failed", bytes);
			return CCG_ERR_NOMEM;
		}
		ctx->out_bytes = bytes;
	}
	if(ctx->remask_pending) {
		CK(ctx, ccg_launch_remask_all(ctx));
		ctx->remask_pending = 0;
	}
	if(ctx->proxi && ctx->words > 0) {
		/* -P: the pair's mask is the new sample's own mask after ITS builder (fsacmpthrd.c:627-628), then the
		 * per-sample builder again against the column sample (:545-546) -- k_row_proxi.  The new sample's planes are
		 * set aside as uploaded (the events are defined on the codes), masked in place for the run, and put back. */
		for(int i = 0; i < ctx->n_pad; ++i) ctx->h_rank[i] = i < ctx->n && use[i] ? 1 : -1;
		int r = 0;
		for(int i = 0; i < ctx->n; ++i)
			if(use[i]) ctx->h_rank[i] = r++;
		free(use);
		EpilogueParams ep;
		memset(&ep, 0, sizeof(ep));
		ep.elem_size = 8;
		ep.norm = norm;
		ep.byteScale = 1.0;
		ep.D = ctx->d_out_D;
		ep.N = ctx->d_out_N;
		ep.rank = ctx->d_rank;
		ep.row_plus1 = ctx->h_rank[row_slot] + 1;
		if(minLength < minCov * ctx->len) minLength = (unsigned) (minCov * ctx->len);
		ep.minLength = minLength;
		ep.nFactor = 1.0;
		void *d_raw = 0;
		unsigned char *d_use = 0;
		unsigned *d_clr = 0;
		int first_used = -1;
		int rc = stage_use_flags(ctx, 0, row_slot, row_slot + 1, &d_use, &d_clr, (size_t) ctx->n_pad, &first_used);
		if(rc) return rc;
		CK(ctx, cudaMalloc(&d_raw, (size_t) ctx->chunks * 3 * 16));
		cudaError_t e = cudaMemcpyAsync(ctx->d_rank, ctx->h_rank, (size_t) ctx->n_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
		if(e == cudaSuccess) e = ccg_launch_row_planes(ctx, row_slot, d_raw, 0);
		if(e == cudaSuccess && !ctx->proxi_snp_only) e = ccg_launch_sample_proxi(ctx, 0, 0, d_use, 1, d_clr);
		if(e == cudaSuccess) e = cudaEventRecord(ctx->ev0, ctx->stream);
		if(e == cudaSuccess) e = ccg_launch_row_proxi(ctx, row_slot, d_raw, ep);
		if(e == cudaSuccess) e = cudaEventRecord(ctx->ev1, ctx->stream);
		if(e == cudaSuccess) e = ccg_launch_row_planes(ctx, row_slot, d_raw, 1);
		if(e == cudaSuccess) e = cudaMemcpyAsync(D, ctx->d_out_D, bytes, cudaMemcpyDeviceToHost, ctx->stream);
		if(e == cudaSuccess && N) e = cudaMemcpyAsync(N, ctx->d_out_N, bytes, cudaMemcpyDeviceToHost, ctx->stream);
		if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		cudaFree(d_raw);
		if(e != cudaSuccess) {
			set_err(ctx, "row run with proximity masking failed: %s", cudaGetErrorString(e));
			return CCG_ERR_CUDA;
		}
		ctx->ev_valid = 1;
		ctx->last_Dn = 0;
		snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k_row_proxi cols=%d proxi=%u", cols, ctx->proxi);
		return CCG_OK;
	}
	const int win_on = ctx->win_on;
	int win[4];
	memcpy(win, ctx->win, sizeof(win));
	int rc = ccg_set_tile_window(ctx, row_slot / CCG_UMMA_BM * CCG_UMMA_BM, row_slot + 1, 0, row_slot > 0 ? row_slot : 1);
	if(!rc) {
		int Dn = 0;
		ctx->row_slot1 = row_slot + 1;
		rc = run_common(ctx, 0, use, norm, minLength, minCov, 8, 1.0, ctx->d_out_D, ctx->d_out_N, &Dn);
		ctx->row_slot1 = 0;
		ctx->last_Dn = 0;               /* one tile stripe only: nothing for ccg_get_raw_counts to gather */
	}
	ctx->win_on = win_on;
	memcpy(ctx->win, win, sizeof(win));
	update_need(ctx);
	free(use);
	if(rc) return rc;
	CK(ctx, cudaMemcpyAsync(D, ctx->d_out_D, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	if(N) CK(ctx, cudaMemcpyAsync(N, ctx->d_out_N, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return CCG_OK;
}

/* -V: fsacmpairint (fsacmp.c:685) / fsacmprint (:646) for every compared pair, see k_variants.cu */
static int list_variants_impl(ccg_ctx *ctx, int pair, const unsigned char *include, int last_row_only, ccg_variant_fn fn, void *user) {
	if(!ctx || !ctx->d_planes || !ctx->pair_mode || !fn) return CCG_ERR_ARG;
	if(ctx->world > 1 || ctx->grp_world > 1) {
		set_err(ctx, "variant listing (-V) is not available together with a rank partition");
		return CCG_ERR_UNSUPPORTED;
	}
	/* -P: in pair mode the reference lists under maskProxi's mask (fsacmpthrd.c:410-414), built per batch below; in
	 * shared-mask mode the proximity ranges are part of the global mask already (cdist.c:111); the row form (-a) walks
	 * the new sample's mask after the per-sample builder against each column sample (fsacmpthrd.c:545-553). */
	const int pair_proxi = pair && ctx->proxi && ctx->words > 0;
	const int row_proxi = pair_proxi && last_row_only;
	if(pair_proxi && ctx->codes_upload_masked) {
		set_err(ctx, "variant listing (-V) with proximity masking (-P): call ccg_set_proximity before the packed rows are uploaded");
		return CCG_ERR_ARG;
	}
	if(pair ? ctx->global_applied : !ctx->global_pending) {
		set_err(ctx, pair ? "variant listing in pair mode needs a store without a global mask"
		                  : "variant listing in shared-mask mode runs after ccg_build_global_mask and before ccg_run_global");
		return CCG_ERR_ARG;
	}
	CK(ctx, cudaSetDevice(ctx->device));
	NEED_PLANES(ctx);
	const int n = ctx->n;
	int *slot_of = (int *) malloc((size_t) (n ? n : 1) * sizeof(int));
	if(!slot_of) return CCG_ERR_NOMEM;
	int Dn = 0;
	for(int i = 0; i < n; ++i)
		if(ctx->present[i] && ctx->have[i >> 7] && (!include || include[i])) slot_of[Dn++] = i;
	if(Dn < 2) { free(slot_of); return CCG_OK; }
	const long long cells = (long long) Dn * (Dn - 1) / 2;
	const long long cells_lo = last_row_only ? (long long) (Dn - 1) * (Dn - 2) / 2 : 0;
	int BATCH = 1 << 18;                             /* cells per count pass */
	const size_t CAP = (size_t) 1 << 24;             /* entries per write pass (128 MiB) */
	uint32_t *d_pmask = 0;
	if(pair_proxi) {
		/* one mask column of `words` words per cell of a batch: at most 1 GiB (or half of what is free) of scratch */
		size_t free_b = 0, total_b = 0;
		CK(ctx, cudaMemGetInfo(&free_b, &total_b));
		size_t budget = (size_t) 1 << 30;
		if(budget > free_b / 2) budget = free_b / 2;
		long long fit = (long long) (budget / ((size_t) ctx->words * 4));
		fit &= ~127LL;
		if(fit < 128) fit = 128;
		if(fit < BATCH) BATCH = (int) fit;
		if(cells - cells_lo < BATCH) BATCH = (int) (((cells - cells_lo) + 127) & ~127LL);
		if(cudaMalloc(&d_pmask, (size_t) BATCH * (size_t) ctx->words * 4) != cudaSuccess) {
			cudaGetLastError();
			set_err(ctx, "cudaMalloc of %zu bytes for the per-pair proximity masks failed", (size_t) BATCH * (size_t) ctx->words * 4);
			free(slot_of);
			return CCG_ERR_NOMEM;
		}
	}
	/* row form under -P: the new sample's planes are set aside as uploaded, its own builder is applied to the store's
	 * copy (as ccg_run_row does), and everything is put back at the end */
	void *d_raw = 0;
	const int row_slot = slot_of[Dn - 1];
	if(row_proxi) {
		unsigned char *d_use = 0;
		unsigned *d_clr = 0;
		int first_used = -1;
		int rcu = stage_use_flags(ctx, 0, row_slot, row_slot + 1, &d_use, &d_clr, (size_t) ctx->n_pad, &first_used);
		cudaError_t e0 = rcu ? cudaErrorUnknown : cudaMalloc(&d_raw, (size_t) ctx->chunks * 3 * 16);
		if(e0 == cudaSuccess) e0 = ccg_launch_row_planes(ctx, row_slot, d_raw, 0);
		if(e0 == cudaSuccess && !ctx->proxi_snp_only) e0 = ccg_launch_sample_proxi(ctx, 0, 0, d_use, 1, d_clr);
		if(e0 != cudaSuccess) {
			if(!rcu) set_err(ctx, "variant listing of a row with proximity masking failed: %s", cudaGetErrorString(e0));
			cudaFree(d_raw);
			cudaFree(d_pmask);
			free(slot_of);
			return rcu ? rcu : CCG_ERR_CUDA;
		}
	}
	int *d_slot = 0;
	unsigned *d_counts = 0, *h_counts = 0;
	unsigned long long *d_off = 0, *h_off = 0, *d_ent = 0, *h_ent = 0;
	size_t ent_cap = 0;
	int rc = CCG_OK;
	cudaError_t e = cudaMalloc(&d_slot, (size_t) Dn * sizeof(int));
	if(e == cudaSuccess) e = cudaMalloc(&d_counts, (size_t) BATCH * sizeof(unsigned));
	if(e == cudaSuccess) e = cudaMalloc(&d_off, (size_t) (BATCH + 1) * sizeof(unsigned long long));
	h_counts = (unsigned *) malloc((size_t) BATCH * sizeof(unsigned));
	h_off = (unsigned long long *) malloc((size_t) (BATCH + 1) * sizeof(unsigned long long));
	if(e == cudaSuccess && (!h_counts || !h_off)) rc = CCG_ERR_NOMEM;
	if(e == cudaSuccess && !rc) e = cudaMemcpyAsync(d_slot, slot_of, (size_t) Dn * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
	for(long long c0 = cells_lo; c0 < cells && e == cudaSuccess && !rc; c0 += BATCH) {
		const int nb = (int) (cells - c0 < BATCH ? cells - c0 : BATCH);
		VariantParams p;
		memset(&p, 0, sizeof(p));
		p.ncells = nb;
		p.cell0 = c0;
		p.slot_of_rank = d_slot;
		p.counts = d_counts;
		p.pair_mask = d_pmask;
		p.pair_mask_stride = BATCH;
		p.row_raw = (const uint4 *) d_raw;
		if(d_pmask) e = row_proxi ? ccg_launch_row_proxi_mask(ctx, p, row_slot) : ccg_launch_pair_proxi_mask(ctx, p);
		if(e == cudaSuccess) e = ccg_launch_variants(ctx, p, 0, !pair);
		if(e == cudaSuccess) e = cudaMemcpyAsync(h_counts, d_counts, (size_t) nb * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream);
		if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		if(e != cudaSuccess) break;
		h_off[0] = 0;
		for(int k = 0; k < nb; ++k) h_off[k + 1] = h_off[k] + h_counts[k];
		/* the entry buffers grow to what a batch needs, up to CAP entries (most listings are small) */
		size_t want = h_off[nb] < CAP ? (size_t) h_off[nb] : CAP;
		for(int k = 0; k < nb; ++k)
			if(h_counts[k] > want) want = h_counts[k] < CAP ? h_counts[k] : CAP;
		if(want > ent_cap) {
			cudaFree(d_ent); d_ent = 0;
			cudaFreeHost(h_ent); h_ent = 0;
			ent_cap = want < 65536 ? 65536 : want;
			e = cudaMalloc(&d_ent, ent_cap * sizeof(unsigned long long));
			if(e == cudaSuccess) e = cudaMallocHost(&h_ent, ent_cap * sizeof(unsigned long long));
			if(e != cudaSuccess) break;
		}
		e = cudaMemcpyAsync(d_off, h_off, (size_t) (nb + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream);
		/* write passes over runs of cells whose lists fit the entry buffer */
		int k0 = 0;
		while(k0 < nb && e == cudaSuccess && !rc) {
			int k1 = k0;
			while(k1 < nb && h_off[k1 + 1] - h_off[k0] <= CAP) ++k1;
			if(k1 == k0) {
				set_err(ctx, "one pair has %u variants: more than the %zu the listing buffer holds", h_counts[k0], CAP);
				rc = CCG_ERR_NOMEM;
				break;
			}
			const size_t nent = (size_t) (h_off[k1] - h_off[k0]);
			if(nent) {
				p.ncells = k1 - k0;
				p.cell0 = c0 + k0;
				p.offsets = d_off + k0;
				p.entries = d_ent;
				p.pair_mask = d_pmask ? d_pmask + k0 : 0;
				e = ccg_launch_variants(ctx, p, 1, !pair);
				if(e == cudaSuccess) e = cudaMemcpyAsync(h_ent, d_ent, nent * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
				if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
				if(e != cudaSuccess) break;
				/* hand the lists out pair by pair, rows ascending, columns ascending: the reference's -t 1 order */
				long long cell = c0 + k0;
				long long r = (long long) ((1.0 + sqrt(1.0 + 8.0 * (double) cell)) * 0.5);
				while(r * (r - 1) / 2 > cell) --r;
				while((r + 1) * r / 2 <= cell) ++r;
				long long c = cell - r * (r - 1) / 2;
				for(int k = k0; k < k1 && !rc; ++k) {
					if(h_counts[k] && fn(user, slot_of[r], slot_of[c], (const uint64_t *) (h_ent + (h_off[k] - h_off[k0])), h_counts[k]))
						rc = CCG_ERR_ARG;
					if(++c == r) { ++r; c = 0; }
				}
			}
			k0 = k1;
		}
	}
	if(e != cudaSuccess) {
		set_err(ctx, "variant listing failed: %s", cudaGetErrorString(e));
		rc = CCG_ERR_CUDA;
	}
	if(d_raw) {
		/* the new sample's planes go back as uploaded */
		cudaError_t er = ccg_launch_row_planes(ctx, row_slot, d_raw, 1);
		if(er == cudaSuccess) er = cudaStreamSynchronize(ctx->stream);
		if(er != cudaSuccess && !rc) {
			set_err(ctx, "variant listing of a row: restoring the planes failed: %s", cudaGetErrorString(er));
			rc = CCG_ERR_CUDA;
		}
		cudaFree(d_raw);
	}
	cudaFree(d_pmask);
	cudaFree(d_slot);
	cudaFree(d_counts);
	cudaFree(d_off);
	cudaFree(d_ent);
	cudaFreeHost(h_ent);
	free(h_counts);
	free(h_off);
	free(slot_of);
	return rc;
}

extern "C" int ccg_list_variants(ccg_ctx *ctx, int pair, const unsigned char *include, ccg_variant_fn fn, void *user) {
	CCG_MULTI_SOLO(ctx, "variant listing (-V)", ccg_list_variants(m0, pair, include, fn, user));
	return list_variants_impl(ctx, pair, include, 0, fn, user);
}

/* -V with -a: fsacmpairint(diffile, n, j, addL, seqL, ...) of cmpFsaRowThrd (fsacmpthrd.c:552-553) */
extern "C" int ccg_list_variants_row(ccg_ctx *ctx, int row_slot, ccg_variant_fn fn, void *user) {
	CCG_MULTI_SOLO(ctx, "variant listing (-V)", ccg_list_variants_row(m0, row_slot, fn, user));
	if(!ctx || !ctx->d_planes || row_slot < 0 || row_slot >= ctx->n || !ctx->present[row_slot]) return CCG_ERR_ARG;
	unsigned char *use = (unsigned char *) calloc((size_t) ctx->n, 1);
	if(!use) return CCG_ERR_NOMEM;
	for(int j = 0; j <= row_slot; ++j) use[j] = ctx->present[j];
	int rc = list_variants_impl(ctx, 1, use, 1, fn, user);
	free(use);
	return rc;
}

extern "C" int ccg_get_raw_counts(ccg_ctx *ctx, uint32_t *mism, uint32_t *ninc) {
	CCG_MULTI_SOLO(ctx, "ccg_get_raw_counts", ccg_get_raw_counts(m0, mism, ninc));
	if(!ctx || !ctx->d_planes) return CCG_ERR_ARG;
	if(ctx->grp_world > 1) {
		set_err(ctx, "ccg_get_raw_counts: a member of a K-split group holds partial sums only");
		return CCG_ERR_UNSUPPORTED;
	}
	int Dn = ctx->last_Dn;
	if(Dn < 2) return CCG_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	size_t cells = (size_t) Dn * (Dn - 1) / 2;
	uint32_t *d_m = 0, *d_n = 0;
	CK(ctx, cudaMalloc(&d_m, cells * 4));
	if(cudaMalloc(&d_n, cells * 4) != cudaSuccess) { cudaFree(d_m); return CCG_ERR_NOMEM; }
	cudaMemsetAsync(d_m, 0, cells * 4, ctx->stream);
	cudaMemsetAsync(d_n, 0, cells * 4, ctx->stream);
	cudaError_t e;
	if(ctx->last_kernel_kind == CCG_KERNEL_UMMA || ctx->last_kernel_kind == CCG_KERNEL_FUSED) e = ccg_launch_gather_raw_dense(ctx, ctx->last_i_const, d_m, d_n);
	else e = ccg_launch_gather_raw(ctx, d_m, d_n);
	if(e == cudaSuccess && mism) e = cudaMemcpyAsync(mism, d_m, cells * 4, cudaMemcpyDeviceToHost, ctx->stream);
	if(e == cudaSuccess && ninc) e = cudaMemcpyAsync(ninc, d_n, cells * 4, cudaMemcpyDeviceToHost, ctx->stream);
	if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	cudaFree(d_m);
	cudaFree(d_n);
	if(e != cudaSuccess) {
		set_err(ctx, "raw count gather failed: %s", cudaGetErrorString(e));
		return CCG_ERR_CUDA;
	}
	return CCG_OK;
}

extern "C" int ccg_fsa_cmp_thread_out(ccg_ctx *ctx, int pair, void *D, void *N, int elem_size, double byteScale, int n,
                                      int len, const uint64_t *const *seqs, const unsigned char *include,
                                      const uint32_t *const *includes, unsigned norm, unsigned minLength, double minCov,
                                      unsigned proxi, int *Dn, unsigned *global_inc) {
	if(!seqs || !includes || n < 0) return CCG_ERR_ARG;
	if(ctx && ctx->multi)
		return ccg_multi_fsa_cmp_thread_out(ctx, pair, D, N, elem_size, byteScale, n, len, seqs, include, includes, norm, minLength,
		                                    minCov, proxi, Dn, global_inc);
	ccg_ctx *own = 0;
	int rc;
	if(!ctx) {
		rc = ccg_init(&own, -1);
		if(rc) return rc;
		ctx = own;
	}
	/* -P: the caller's per-sample masks already carry getIncPos' proximity masking (cdist.c:91); what is left
	 * for the fan-out is maskProxi per pair (fsacmpthrd.c:410).  cmpFsaThrd never looks at proxi. */
	const unsigned saved_proxi = ctx->proxi;
	ctx->proxi = pair ? proxi : 0;
	rc = ccg_set_problem(ctx, n, len, pair);
	if(!rc && !pair) rc = ccg_put_global_mask(ctx, includes[0]);
	const uint64_t **srow = 0;
	const uint32_t **mrow = 0;
	if(!rc) {
		/* excluded samples are never uploaded: NULL row = empty slot */
		srow = (const uint64_t **) malloc((size_t) (n ? n : 1) * sizeof(*srow));
		mrow = (const uint32_t **) malloc((size_t) (n ? n : 1) * sizeof(*mrow));
		if(!srow || !mrow) rc = CCG_ERR_NOMEM;
	}
	int streamed = 0;
	if(!rc) {
		int ninc = 0;
		for(int i = 0; i < n; ++i) {
			int inc = (!include || include[i]) && seqs[i] && (pair ? includes[i] != 0 : 1);
			srow[i] = inc ? seqs[i] : 0;
			mrow[i] = inc ? (pair ? includes[i] : includes[0]) : 0;
			ninc += inc;
		}
		/* tensor path on a long alignment: stream the rows K slab by K slab inside the run, so that
		 * the PCIe upload of slab s+1 hides behind the GEMM of slab s (run_umma / feed_slab) */
		int kind = ctx->kernel_choice;
		if(kind == CCG_KERNEL_AUTO) kind = (ninc >= 192 && ctx->chunks >= 64) ? CCG_KERNEL_UMMA : CCG_KERNEL_POPC;
		if(ctx->grp_world > 1) kind = CCG_KERNEL_UMMA;
		if(kind == CCG_KERNEL_UMMA && !ctx->proxi && ctx->stream_min_chunks > 0 && ctx->chunks >= ctx->stream_min_chunks &&
		   3.0 * ((double) len + 256.0) < 2147483648.0 &&
		   !(ctx->grp_world > 1 && ccg_group_window_rows(ctx) < ctx->n_pad)) {      /* windowed group runs expand the whole slice first */
			for(int i = 0; i < n; ++i) {
				if(!srow[i]) continue;
				ctx->present[i] = 1;
				if(ctx->need[i >> 7]) ctx->have[i >> 7] = 1;
			}
			if(cudaMemsetAsync(ctx->d_inc, 0, (size_t) ctx->n_pad * sizeof(unsigned), ctx->stream) != cudaSuccess) {
				set_err(ctx, "cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
				rc = CCG_ERR_CUDA;
			}
			ctx->feed_seqs = srow;
			ctx->feed_masks = pair ? mrow : 0;
			streamed = 1;
		} else {
			rc = ccg_put_samples_packed(ctx, 0, n, srow, pair ? mrow : 0);
		}
	}
	if(!rc) {
		if(pair) rc = ccg_run_pair(ctx, include, norm, minLength, minCov, elem_size, byteScale, D, N, Dn);
		else rc = ccg_run_global(ctx, include, norm, elem_size, byteScale, D, Dn, global_inc);
	}
	if(streamed) {
		/* the run has synchronised the main stream, which waited for every slab's upload */
		for(int q = 0; q < 2; ++q) cudaStreamSynchronize(ctx->copy_stream[q]);
		ctx->feed_seqs = 0;
		ctx->feed_masks = 0;
	}
	free(srow);
	free(mrow);
	ctx->proxi = saved_proxi;
	if(own) {
		if(rc) snprintf(g_init_err, sizeof(g_init_err), "%s", own->err);
		ccg_destroy(own);
	}
	return rc;
}

extern "C" void *ccg_host_alloc(size_t bytes) {
	void *p = 0;
	if(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return 0;
	return p;
}

extern "C" void ccg_host_free(void *p) {
	if(p) cudaFreeHost(p);
}

extern "C" long long ccg_launch_count(const ccg_ctx *ctx) {
	if(ctx && ctx->multi) return ccg_multi_launch_count(ctx);
	return ctx ? ctx->launches : 0;
}
extern "C" const char *ccg_last_kernel(const ccg_ctx *ctx) {
	if(ctx && ctx->multi) return ccg_multi_last_kernel(ctx);
	return ctx ? ctx->last_kernel : "";
}

extern "C" float ccg_last_phase_ms(ccg_ctx *ctx, int phase) {
	float ms = -1.0f;
	if(ctx && ctx->multi) return ccg_last_phase_ms(ccg_multi_member(ctx, 0), phase);
	if(!ctx || !ctx->phase_valid || phase < 0 || phase > 1) return ms;
	if(ctx->last_kernel_kind == CCG_KERNEL_POPC || (ctx->last_kernel_kind == CCG_KERNEL_FUSED && phase == 0)) return ms;
	if(cudaEventSynchronize(ctx->ev_phase[2 * phase + 1]) != cudaSuccess) return -1.0f;
	if(cudaEventElapsedTime(&ms, ctx->ev_phase[2 * phase], ctx->ev_phase[2 * phase + 1]) != cudaSuccess) return -1.0f;
	return ms;
}

extern "C" float ccg_last_compare_ms(ccg_ctx *ctx) {
	float ms = -1.0f;
	if(ctx && ctx->multi) return ccg_multi_last_compare_ms(ctx);
	if(!ctx || !ctx->ev_valid) return ms;
	if(cudaEventSynchronize(ctx->ev1) != cudaSuccess) return -1.0f;
	if(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return -1.0f;
	return ms;
}
