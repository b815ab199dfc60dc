/*
 * k_trim.cu -- the masks of `ccphylo trim` (fsaTrim, trim.c:77-260) on the device, behind ccg_trim_*.
 *
 * trim has no pairwise stage: per sample it (a) clears the positions that are unknown or soft-masked in the sample or in
 * the reference sample, (b) clears the methylation sites of the motif matches (maskMotifs, meth.c:141), (c) clears the
 * runs between two "events" at most proxi positions apart (getIncPos / getIncPosInsig / getIncPosInsigPrune,
 * fsacmp.c:181-353), and -- for a pseudo alignment, flag 16 -- (d) keeps only the columns where some sample differs
 * from the first one (pseudoAlnPrune fsacmp.c:504-551).  The mask is either the sample's own (pairwise flag: every
 * sample against itself) or one shared mask that every sample narrows.
 *
 * Unlike `dist`, trim translates with the 16-letter IUPAC table unless flag 4 asks for the 2-bit one (trim.c:103,
 * getIupacBitTable fsacmp.c:93-162): codes 0-3 bases, 4 unknown, 5 gap, 6-15 ambiguity letters, +16 = soft-masked
 * (lower-case) input.  The kernels therefore work on the translated code bytes themselves, one byte per position:
 *
 *   k_trim_words   one thread per 32-position word: the bits (a) clears, the event bits of (c), the column bits of (d)
 *                  HBM: 32 B (self) / 64 B (against the reference sample) read, 4-12 B written per word
 *   k_trim_proxi   the range clearing of (c) over the event words: proxi_scan_words of proxi_core.h, one thread per
 *                  segment, which finds the event its first range may start from by looking back proxi positions
 *   k_trim_and_plane / k_trim_count   (b) comes from the plane store (ccg_put_samples_packed of the reference's own
 *                  packed words + ccg_mask_motifs) and is ANDed in; getNpos
 *
 * What the reference does per position, c = the sample's code, r = the reference sample's code after ITS self pass
 * (which strips every soft flag: r < 16):
 *   getIncPos (fsacmp.c:181)             c != r || c == 4 || c & 16  -> event;  c == 4 || r == 4 || c & 16 -> cleared
 *   getIncPosInsigPrune (:240, flag 32)  c == 4 || r == 4 || c & 16 -> cleared, no event;  else c != r -> event
 *   getIncPosInsig (:297, flag 8)        c == 4 || r == 4 -> cleared;  else c != r -> event (a soft c differs from r)
 * A cleared soft position has its flag stripped in the stored sequence (c &= 15) unless r == 4 (getIncPos / Prune); the
 * column test of (d) compares the stored bytes.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ccg_internal.h"
#include "proxi_core.h"

namespace {

constexpr int TRIM_SEG_WORDS = 1024;

/* builder: 0 getIncPos, 1 getIncPosInsig, 2 getIncPosInsigPrune */
__global__ void __launch_bounds__(256)
k_trim_words(const unsigned char *__restrict__ cur, const unsigned char *__restrict__ ref, long long len, int words, int builder,
             int init, uint32_t *__restrict__ mask, uint32_t *__restrict__ events, uint32_t *__restrict__ columns) {
	const int w = blockIdx.x * blockDim.x + threadIdx.x;
	if(w >= words) return;
	const long long p0 = (long long) w * 32;
	uint32_t clr = 0, ev = 0, col = 0;
	/* the buffers are padded to whole words (zeroed behind len): two 16-byte loads per sample */
	const uint4 *c4 = reinterpret_cast<const uint4 *>(cur + p0);
	const uint4 *r4 = ref ? reinterpret_cast<const uint4 *>(ref + p0) : 0;
	uint32_t cw[8], rw[8];
	{
		const uint4 a = c4[0], b = c4[1];
		cw[0] = a.x; cw[1] = a.y; cw[2] = a.z; cw[3] = a.w; cw[4] = b.x; cw[5] = b.y; cw[6] = b.z; cw[7] = b.w;
		if(r4) {
			const uint4 x = r4[0], y = r4[1];
			rw[0] = x.x; rw[1] = x.y; rw[2] = x.z; rw[3] = x.w; rw[4] = y.x; rw[5] = y.y; rw[6] = y.z; rw[7] = y.w;
		}
	}
#pragma unroll
	for(int k = 0; k < 32; ++k) {
		const unsigned c = (cw[k >> 2] >> (8 * (k & 3))) & 0xFFu;
		const uint32_t bit = 0x80000000u >> k;
		if(!r4) {
			/* the sample against itself: getIncPos(includes, seq, seq, proxi), trim.c:183,201 */
			if(c == 4 || (c & 16)) { clr |= bit; ev |= bit; }
		} else {
			const unsigned r = (rw[k >> 2] >> (8 * (k & 3))) & 0xFFu;
			const bool unknown = c == 4 || r == 4, soft = (c & 16) != 0;
			unsigned stored = c;
			if(builder == 0) {
				if(c != r || c == 4 || soft) ev |= bit;
				if(unknown || soft) clr |= bit;
				if(!unknown && soft) stored = c & 15;
			} else if(builder == 2) {
				if(unknown || soft) clr |= bit;
				else if(c != r) ev |= bit;
				if(!unknown && soft) stored = c & 15;
			} else {
				if(unknown) clr |= bit;
				else if(c != r) ev |= bit;
			}
			if(stored != r) col |= bit;
		}
	}
	const uint32_t valid = proxi_valid_bits(len, p0);
	const uint32_t m = init ? valid : mask[w];
	mask[w] = m & ~clr;
	events[w] = ev & valid;
	if(columns && r4) columns[w] |= col & valid;
}

struct WordEvents {
	const uint32_t *events;
	__device__ __forceinline__ uint32_t operator()(long long w) const { return events[w]; }
};
struct MaskSink {
	uint32_t *mask;
	__device__ __forceinline__ void clear(long long w, uint32_t bits) {
		if(mask[w] & bits) atomicAnd(mask + w, ~bits);
	}
};

__global__ void __launch_bounds__(128)
k_trim_proxi(const uint32_t *__restrict__ events, uint32_t *mask, int words, unsigned proxi) {
	const long long seg = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	const long long w_begin = seg * TRIM_SEG_WORDS;
	if(w_begin >= words) return;
	long long w_end = w_begin + TRIM_SEG_WORDS;
	if(w_end > words) w_end = words;
	WordEvents ev = {events};
	const long long back = w_begin * 32 - (long long) proxi;
	const long long w_lo = back <= 0 ? 0 : (back >> 5);
	const long long last = proxi_last_event_before(w_lo, w_begin, ev);
	MaskSink sink = {mask};
	proxi_scan_words(last, w_begin, w_end, proxi, ev, sink);
}

/* mask &= the mask plane of slot 0 of the plane store: the methylation sites ccg_mask_motifs removed */
__global__ void __launch_bounds__(256)
k_trim_and_plane(uint32_t *__restrict__ mask, const uint32_t *__restrict__ planes, int n_pad, int words) {
	const int w = blockIdx.x * blockDim.x + threadIdx.x;
	if(w >= words) return;
	mask[w] &= planes[(((size_t) (w >> 2) * 3 + 2) * n_pad) * 4 + (size_t) (w & 3)];
}

__global__ void __launch_bounds__(256)
k_trim_count(const uint32_t *__restrict__ mask, const uint32_t *__restrict__ columns, int words, unsigned *__restrict__ count) {
	unsigned c = 0, v = 0;
	for(int w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
		c += (unsigned) __popc(mask[w]);
		if(columns) v += (unsigned) __popc(mask[w] & columns[w]);
	}
#pragma unroll
	for(int o = 16; o; o >>= 1) {
		c += __shfl_xor_sync(0xffffffffu, c, o);
		v += __shfl_xor_sync(0xffffffffu, v, o);
	}
	if((threadIdx.x & 31) == 0) {
		if(c) atomicAdd(count, c);
		if(v) atomicAdd(count + 1, v);
	}
}

/* the reference sample's stored bytes: every soft flag stripped by its own getIncPos pass (fsacmp.c:200-206) */
__global__ void __launch_bounds__(256)
k_trim_keep_ref(const unsigned char *__restrict__ cur, unsigned char *__restrict__ ref, long long padded) {
	for(long long p = (long long) blockIdx.x * blockDim.x + threadIdx.x; p < padded / 4; p += (long long) gridDim.x * blockDim.x)
		reinterpret_cast<uint32_t *>(ref)[p] = reinterpret_cast<const uint32_t *>(cur)[p] & 0x0F0F0F0Fu;
}

} // namespace

#define TCK(ctx, call)                                                                                       \
	do {                                                                                                     \
		cudaError_t e__ = (call);                                                                            \
		if(e__ != cudaSuccess) {                                                                             \
			snprintf(ctx->err, sizeof(ctx->err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
			return CCG_ERR_CUDA;                                                                             \
		}                                                                                                    \
	} while(0)

void ccg_trim_free(ccg_ctx *ctx) {
	cudaFree(ctx->trim_cur); ctx->trim_cur = 0;
	cudaFree(ctx->trim_ref); ctx->trim_ref = 0;
	cudaFree(ctx->trim_mask); ctx->trim_mask = 0;
	cudaFree(ctx->trim_events); ctx->trim_events = 0;
	cudaFree(ctx->trim_columns); ctx->trim_columns = 0;
	cudaFree(ctx->trim_count); ctx->trim_count = 0;
	free(ctx->trim_ones); ctx->trim_ones = 0;
	ctx->trim_len = -1;
	ctx->trim_has_ref = 0;
}

extern "C" int ccg_trim_begin(ccg_ctx *ctx, int len, unsigned proxi) {
	CCG_MULTI_SOLO(ctx, "trim", ccg_trim_begin(m0, len, proxi));
	if(!ctx || len < 0) return CCG_ERR_ARG;
	TCK(ctx, cudaSetDevice(ctx->device));
	TCK(ctx, cudaStreamSynchronize(ctx->stream));
	ccg_trim_free(ctx);
	ctx->trim_len = len;
	ctx->trim_words = (len + 31) / 32;
	ctx->trim_proxi = proxi;
	const size_t W = (size_t) (ctx->trim_words ? ctx->trim_words : 1), padded = W * 32;
	TCK(ctx, cudaMalloc(&ctx->trim_cur, padded));
	TCK(ctx, cudaMalloc(&ctx->trim_ref, padded));
	TCK(ctx, cudaMalloc(&ctx->trim_mask, W * 4));
	TCK(ctx, cudaMalloc(&ctx->trim_events, W * 4));
	TCK(ctx, cudaMalloc(&ctx->trim_columns, W * 4));
	TCK(ctx, cudaMalloc(&ctx->trim_count, 2 * sizeof(unsigned)));
	TCK(ctx, cudaMemsetAsync(ctx->trim_cur, 0, padded, ctx->stream));
	TCK(ctx, cudaMemsetAsync(ctx->trim_ref, 0, padded, ctx->stream));
	TCK(ctx, cudaMemsetAsync(ctx->trim_mask, 0, W * 4, ctx->stream));
	TCK(ctx, cudaMemsetAsync(ctx->trim_columns, 0, W * 4, ctx->stream));
	if(ctx->motif_n && len > 0) {
		/* maskMotifs reads the packed words: they go through slot 0 of a plane store with an all-ones mask (two slots:
		 * a store keeps only the row blocks some pair of the problem reads) */
		int rc = ccg_set_problem(ctx, 2, len, 1);
		if(rc) return rc;
		ctx->trim_ones = (uint32_t *) malloc(W * 4);
		if(!ctx->trim_ones) return CCG_ERR_NOMEM;
		for(size_t w = 0; w < W; ++w) {
			const long long left = (long long) len - (long long) w * 32;
			ctx->trim_ones[w] = left >= 32 ? 0xFFFFFFFFu : left <= 0 ? 0u : 0xFFFFFFFFu << (32 - (int) left);
		}
	}
	return CCG_OK;
}

extern "C" int ccg_trim_sample(ccg_ctx *ctx, const unsigned char *codes, const uint64_t *nibbles, int against_ref, int builder,
                               unsigned *inc_out) {
	CCG_MULTI_SOLO(ctx, "trim", ccg_trim_sample(m0, codes, nibbles, against_ref, builder, inc_out));
	if(!ctx || ctx->trim_len < 0 || !ctx->trim_cur || (ctx->trim_len && !codes) || builder < 0 || builder > 2) return CCG_ERR_ARG;
	if(against_ref && !ctx->trim_has_ref) {
		snprintf(ctx->err, sizeof(ctx->err), "ccg_trim_sample: no reference sample has been kept (ccg_trim_keep_reference)");
		return CCG_ERR_ARG;
	}
	if(ctx->motif_n && ctx->trim_len && !nibbles) return CCG_ERR_ARG;
	TCK(ctx, cudaSetDevice(ctx->device));
	const int W = ctx->trim_words;
	if(W > 0) {
		/* pageable or pinned source: the caller's buffer is free again when the call returns (synchronised below) */
		TCK(ctx, cudaMemcpyAsync(ctx->trim_cur, codes, (size_t) ctx->trim_len, cudaMemcpyHostToDevice, ctx->stream));
		k_trim_words<<<(unsigned) ((W + 255) / 256), 256, 0, ctx->stream>>>(ctx->trim_cur, against_ref ? ctx->trim_ref : 0,
		                                                                   (long long) ctx->trim_len, W, against_ref ? builder : 0,
		                                                                   against_ref ? 0 : 1, ctx->trim_mask, ctx->trim_events,
		                                                                   ctx->trim_columns);
		ctx->launches++;
		TCK(ctx, cudaGetLastError());
		if(ctx->motif_n) {
			const uint32_t *ones = ctx->trim_ones;
			int rc = ccg_put_samples_packed(ctx, 0, 1, &nibbles, &ones);
			if(!rc) rc = ccg_mask_motifs(ctx, 0, 1, 0);
			if(rc) return rc;
			k_trim_and_plane<<<(unsigned) ((W + 255) / 256), 256, 0, ctx->stream>>>(ctx->trim_mask, ctx->d_planes, ctx->n_pad, W);
			ctx->launches++;
			TCK(ctx, cudaGetLastError());
		}
		if(ctx->trim_proxi) {
			const int nseg = (W + TRIM_SEG_WORDS - 1) / TRIM_SEG_WORDS;
			k_trim_proxi<<<(unsigned) ((nseg + 127) / 128), 128, 0, ctx->stream>>>(ctx->trim_events, ctx->trim_mask, W, ctx->trim_proxi);
			ctx->launches++;
			TCK(ctx, cudaGetLastError());
		}
	}
	unsigned h[2] = {0, 0};
	if(inc_out && W > 0) {
		TCK(ctx, cudaMemsetAsync(ctx->trim_count, 0, 2 * sizeof(unsigned), ctx->stream));
		int blocks = (W + 255) / 256;
		if(blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
		k_trim_count<<<blocks, 256, 0, ctx->stream>>>(ctx->trim_mask, 0, W, ctx->trim_count);
		ctx->launches++;
		TCK(ctx, cudaMemcpyAsync(h, ctx->trim_count, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
	}
	TCK(ctx, cudaStreamSynchronize(ctx->stream));
	if(inc_out) *inc_out = h[0];
	return CCG_OK;
}

extern "C" int ccg_trim_keep_reference(ccg_ctx *ctx) {
	CCG_MULTI_SOLO(ctx, "trim", ccg_trim_keep_reference(m0));
	if(!ctx || ctx->trim_len < 0 || !ctx->trim_cur) return CCG_ERR_ARG;
	TCK(ctx, cudaSetDevice(ctx->device));
	const long long padded = (long long) (ctx->trim_words ? ctx->trim_words : 1) * 32;
	k_trim_keep_ref<<<(unsigned) (ctx->sm_count * 4), 256, 0, ctx->stream>>>(ctx->trim_cur, ctx->trim_ref, padded);
	ctx->launches++;
	TCK(ctx, cudaGetLastError());
	ctx->trim_has_ref = 1;
	return CCG_OK;
}

extern "C" int ccg_trim_get_mask(ccg_ctx *ctx, int variable_columns_only, uint32_t *mask_out, unsigned *inc_out, unsigned *var_out) {
	CCG_MULTI_SOLO(ctx, "trim", ccg_trim_get_mask(m0, variable_columns_only, mask_out, inc_out, var_out));
	if(!ctx || ctx->trim_len < 0 || !ctx->trim_mask) return CCG_ERR_ARG;
	TCK(ctx, cudaSetDevice(ctx->device));
	const int W = ctx->trim_words;
	unsigned h[2] = {0, 0};
	if(W > 0) {
		TCK(ctx, cudaMemsetAsync(ctx->trim_count, 0, 2 * sizeof(unsigned), ctx->stream));
		int blocks = (W + 255) / 256;
		if(blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
		k_trim_count<<<blocks, 256, 0, ctx->stream>>>(ctx->trim_mask, ctx->trim_columns, W, ctx->trim_count);
		ctx->launches++;
		TCK(ctx, cudaMemcpyAsync(h, ctx->trim_count, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
		if(mask_out) {
			TCK(ctx, cudaMemcpyAsync(mask_out, ctx->trim_mask, (size_t) W * 4, cudaMemcpyDeviceToHost, ctx->stream));
			if(variable_columns_only) {
				/* pseudoAlnPrune (fsacmp.c:541-547): include &= consensus */
				TCK(ctx, cudaStreamSynchronize(ctx->stream));
				uint32_t *cols = (uint32_t *) malloc((size_t) W * 4);
				if(!cols) return CCG_ERR_NOMEM;
				cudaError_t e = cudaMemcpy(cols, ctx->trim_columns, (size_t) W * 4, cudaMemcpyDeviceToHost);
				if(e == cudaSuccess)
					for(int w = 0; w < W; ++w) mask_out[w] &= cols[w];
				free(cols);
				TCK(ctx, e);
			}
		}
	}
	TCK(ctx, cudaStreamSynchronize(ctx->stream));
	if(inc_out) *inc_out = h[0];
	if(var_out) *var_out = h[1];
	return CCG_OK;
}

extern "C" int ccg_trim_end(ccg_ctx *ctx) {
	CCG_MULTI_SOLO(ctx, "trim", ccg_trim_end(m0));
	if(!ctx) return CCG_ERR_ARG;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	ccg_trim_free(ctx);
	return CCG_OK;
}
