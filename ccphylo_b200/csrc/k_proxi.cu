/*
 * k_proxi.cu -- -P proximity masking on the device (the arithmetic is in proxi_core.h).
 *
 *   k_pairdist_proxi   replaces maskProxi (proxi > 0) fsacmp.c:355-485 + fsacmpair :587-633 and the
 *                      pair loop / epilogue of cmpairFsaThrd fsacmpthrd.c:261-480.  The masking is
 *                      a sequential dependence along the alignment (whether a SNP is cleared
 *                      depends on where the previous SNP of THIS pair was), so a pair is walked by
 *                      one thread from the first to the last word with a five-register state; the
 *                      parallelism is over the n(n-1)/2 pairs.  Lanes of a warp hold 32 consecutive
 *                      column samples (512 contiguous bytes per plane row), the row samples are
 *                      warp-uniform loads.  INT-pipe bound like k_pairdist_popc; no K split.
 *   k_sample_proxi     replaces getIncPos / getIncPosInsig / getIncPosInsigPrune with proxi > 0
 *                      (fsacmp.c:181-353): the sample against itself on its own planes (pair mode,
 *                      cdist.c:91) or against the shared-mask reference sample on the global mask
 *                      (cdist.c:111).  One thread per (sample, segment of SEG_CHUNKS chunks); a range
 *                      reaches back at most proxi positions, so a thread finds the event its first
 *                      range may start from by looking back that far.  Clears are atomic ANDs
 *                      (idempotent, order independent).
 */
#include "ccg_internal.h"
#include "epilogue.cuh"
#include "proxi_core.h"

namespace {

constexpr int SEG_CHUNKS = 256;

__device__ __forceinline__ size_t plane_index(int n_pad, long long w, int plane, int slot) {
	return (((size_t) (w >> 2) * 3 + plane) * n_pad + slot) * 4 + (size_t) (w & 3);
}

struct DevEvents {
	const uint32_t *planes;
	int n_pad, slot, ref, vs_ref, snp_only;
	long long len;
	__device__ __forceinline__ uint32_t operator()(long long w) const {
		const uint32_t valid = proxi_valid_bits(len, w * 32);
		const uint32_t ms = planes[plane_index(n_pad, w, 2, slot)];
		if(!vs_ref) return proxi_events(0, snp_only, ms, 0, 0, 0, 0, 0, valid);
		const uint32_t mr = planes[plane_index(n_pad, w, 2, ref)];
		const uint32_t hs = planes[plane_index(n_pad, w, 0, slot)], ls = planes[plane_index(n_pad, w, 1, slot)];
		const uint32_t hr = planes[plane_index(n_pad, w, 0, ref)], lr = planes[plane_index(n_pad, w, 1, ref)];
		return proxi_events(1, snp_only, ms, hs, ls, mr, hr, lr, valid);
	}
};

/* the sample's own planes: count what the masking removes and, if asked, remove it */
struct SelfSink {
	uint32_t *planes;
	int n_pad, slot, apply;
	unsigned cleared;
	__device__ __forceinline__ void clear(long long w, uint32_t bits) {
		uint32_t *m = planes + plane_index(n_pad, w, 2, slot);
		uint32_t old = *m;
		if(!(old & bits)) return;
		if(apply) {
			old = atomicAnd(m, ~bits);
			/* apply == 2: the mask plane only -- the code planes keep the bases (a motif scan may still follow, and
			 * the reference's maskMotifs reads the unchanged sequence); k_remask_all clears them before the first run */
			if(apply == 1) {
				atomicAnd(planes + plane_index(n_pad, w, 0, slot), ~bits);
				atomicAnd(planes + plane_index(n_pad, w, 1, slot), ~bits);
			}
		}
		cleared += (unsigned) __popc(old & bits);
	}
};

struct GlobalSink {
	uint32_t *gmask;
	__device__ __forceinline__ void clear(long long w, uint32_t bits) {
		if(gmask[w] & bits) atomicAnd(gmask + w, ~bits);
	}
};

__global__ void __launch_bounds__(128)
k_sample_proxi(uint32_t *planes, int n_pad, int chunks, int words, long long len, unsigned proxi, int snp_only, int vs_ref,
               int ref_slot, const unsigned char *__restrict__ use, int apply, unsigned *__restrict__ cleared,
               uint32_t *gmask) {
	const int slot = blockIdx.x * 32 + threadIdx.x;
	const int seg = blockIdx.y * 4 + threadIdx.y;
	if(slot >= n_pad || !use[slot]) return;
	const long long w_begin = (long long) seg * SEG_CHUNKS * CCG_CHUNK_WORDS;
	if(w_begin >= words) return;
	long long w_end = w_begin + (long long) SEG_CHUNKS * CCG_CHUNK_WORDS;
	if(w_end > words) w_end = words;
	DevEvents ev = {planes, n_pad, slot, ref_slot, vs_ref, snp_only, len};
	/* the event a range of this segment may start from is at most proxi positions before the segment */
	long long back = w_begin * 32 - (long long) proxi;
	long long w_lo = back <= 0 ? 0 : (back >> 5);
	const long long last = proxi_last_event_before(w_lo, w_begin, ev);
	if(vs_ref) {
		GlobalSink sink = {gmask};
		proxi_scan_words(last, w_begin, w_end, proxi, ev, sink);
	} else {
		SelfSink sink = {planes, n_pad, slot, apply, 0u};
		proxi_scan_words(last, w_begin, w_end, proxi, ev, sink);
		if(sink.cleared) atomicAdd(cleared + slot, sink.cleared);
	}
}

__global__ void __launch_bounds__(256)
k_count_mask(const uint32_t *__restrict__ mask, int words, unsigned *__restrict__ count) {
	unsigned c = 0;
	for(int w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) c += (unsigned) __popc(mask[w]);
#pragma unroll
	for(int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

constexpr int PX_ROWS = 2;       /* row samples per thread */
constexpr int PX_SUB = CCG_TILE / (4 * PX_ROWS);   /* CTAs per 64 x 64 tile: 256 threads = 64 columns x 4 row groups */

__global__ void __launch_bounds__(256)
k_pairdist_proxi(const uint32_t *__restrict__ planes, int n_pad, int chunks, unsigned proxi, ProxiParams p) {
	const int lt = blockIdx.x / PX_SUB, sub = blockIdx.x % PX_SUB;
	const int ti = p.tiles[lt].x, tj = p.tiles[lt].y;
	const int tx = threadIdx.x & (CCG_TILE - 1), ty = threadIdx.x / CCG_TILE;
	const int j = tj * CCG_TILE + tx;
	int i[PX_ROWS];
	bool active[PX_ROWS];
	bool any = false;
	const int rj = p.ep.rank[j];
#pragma unroll
	for(int a = 0; a < PX_ROWS; ++a) {
		i[a] = ti * CCG_TILE + (sub * 4 + ty) * PX_ROWS + a;
		active[a] = i[a] > j && rj >= 0 && p.ep.rank[i[a]] >= 0;
		any |= active[a];
	}
	if(!any) return;

	ProxiPairState st[PX_ROWS];
#pragma unroll
	for(int a = 0; a < PX_ROWS; ++a) proxi_pair_init(st[a]);

	const uint4 *P = reinterpret_cast<const uint4 *>(planes);
	/* The walk is one dependent chain per pair, so the loads of chunk c + 1 are issued before chunk c is worked on
	 * (register double buffer): without that every iteration waits a full memory latency. */
	uint4 jn[3], in_[PX_ROWS][3];
#pragma unroll
	for(int q = 0; q < 3; ++q) {
		jn[q] = __ldg(P + (size_t) q * n_pad + j);
#pragma unroll
		for(int a = 0; a < PX_ROWS; ++a) in_[a][q] = __ldg(P + (size_t) q * n_pad + i[a]);
	}
#pragma unroll 1
	for(int c = 0; c < chunks; ++c) {
		uint4 jc[3], ic[PX_ROWS][3];
#pragma unroll
		for(int q = 0; q < 3; ++q) {
			jc[q] = jn[q];
#pragma unroll
			for(int a = 0; a < PX_ROWS; ++a) ic[a][q] = in_[a][q];
		}
		if(c + 1 < chunks) {
			const size_t row = (size_t) (c + 1) * 3;
#pragma unroll
			for(int q = 0; q < 3; ++q) {
				jn[q] = __ldg(P + (row + q) * n_pad + j);
#pragma unroll
				for(int a = 0; a < PX_ROWS; ++a) in_[a][q] = __ldg(P + (row + q) * n_pad + i[a]);
			}
		}
		const uint4 jh = jc[0], jl = jc[1], jm = jc[2];
#pragma unroll
		for(int a = 0; a < PX_ROWS; ++a) {
			if(!active[a]) continue;
			const uint4 ih = ic[a][0], il = ic[a][1], im = ic[a][2];
			const int p0 = c * CCG_CHUNK_BASES;
			const uint32_t m0 = im.x & jm.x, m1 = im.y & jm.y, m2 = im.z & jm.z, m3 = im.w & jm.w;
			proxi_pair_chunk(st[a], p0, ((il.x ^ jl.x) | (ih.x ^ jh.x)) & m0, ((il.y ^ jl.y) | (ih.y ^ jh.y)) & m1,
			                 ((il.z ^ jl.z) | (ih.z ^ jh.z)) & m2, ((il.w ^ jl.w) | (ih.w ^ jh.w)) & m3, m0, m1, m2, m3, proxi);
		}
	}

	uint32_t *accT = p.acc + (size_t) lt * 2 * CCG_TILE * CCG_TILE;
#pragma unroll
	for(int a = 0; a < PX_ROWS; ++a) {
		if(!active[a]) continue;
		unsigned mism, ninc;
		proxi_pair_finish(st[a], &mism, &ninc);
		const int e = (i[a] - ti * CCG_TILE) * CCG_TILE + tx;
		accT[e] = mism;
		accT[CCG_TILE * CCG_TILE + e] = ninc;
		ccg_write_cell(p.ep, i[a], j, mism, ninc);
	}
}

/* -a with -P: one thread per column sample walks the alignment against the new sample (proxi_row_word).  rowraw holds
 * the new sample's planes as uploaded ([chunk][h, l, m]); its mask plane in the store has been through the
 * per-sample builder by now (fsacmpthrd.c:627-628). */
__global__ void __launch_bounds__(128)
k_row_proxi(const uint32_t *__restrict__ planes, int n_pad, int chunks, long long len, unsigned proxi, int snp_only,
            const uint4 *__restrict__ rowraw, int row_slot, EpilogueParams ep) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if(j >= row_slot || ep.rank[j] < 0) return;
	const uint4 *P = reinterpret_cast<const uint4 *>(planes);
	ProxiRowState st;
	proxi_row_init(st);
#pragma unroll 1
	for(int c = 0; c < chunks; ++c) {
		const size_t row = (size_t) c * 3;
		const uint4 jh = __ldg(P + (row + 0) * n_pad + j), jl = __ldg(P + (row + 1) * n_pad + j), jm = __ldg(P + (row + 2) * n_pad + j);
		const uint4 rh = __ldg(rowraw + row), rl = __ldg(rowraw + row + 1), rm = __ldg(rowraw + row + 2);
		const uint4 own = __ldg(P + (row + 2) * n_pad + row_slot);
		const long long p0 = (long long) c * CCG_CHUNK_BASES;
#define CCG_ROW_WORD(f, q)                                                                                        \
	{                                                                                                             \
		const uint32_t valid = proxi_valid_bits(len, p0 + 32 * q);                                                \
		const uint32_t ev = proxi_events(1, snp_only, jm.f, jh.f, jl.f, rm.f, rh.f, rl.f, valid);                  \
		const uint32_t m = own.f & jm.f;                                                                          \
		const uint32_t d = ((rh.f ^ jh.f) | (rl.f ^ jl.f)) & m;                                                   \
		proxi_row_word(st, p0 + 32 * q, ev, m, d, proxi);                                                         \
	}
		CCG_ROW_WORD(x, 0) CCG_ROW_WORD(y, 1) CCG_ROW_WORD(z, 2) CCG_ROW_WORD(w, 3)
#undef CCG_ROW_WORD
	}
	unsigned mism, ninc;
	proxi_row_finish(st, &mism, &ninc);
	ccg_write_cell(ep, row_slot, j, mism, ninc);
}

/* the three plane words of every chunk of one slot <-> a compact [chunk][3] buffer */
__global__ void __launch_bounds__(256)
k_row_planes(uint32_t *planes, int n_pad, int chunks, int slot, uint4 *buf, int restore) {
	uint4 *P = reinterpret_cast<uint4 *>(planes);
	for(long long e = blockIdx.x * (long long) blockDim.x + threadIdx.x; e < (long long) chunks * 3; e += (long long) gridDim.x * blockDim.x) {
		if(restore) P[e * n_pad + slot] = buf[e];
		else buf[e] = P[e * n_pad + slot];
	}
}

} // namespace

cudaError_t ccg_launch_row_planes(ccg_ctx *ctx, int slot, void *d_buf, int restore) {
	int blocks = (ctx->chunks * 3 + 255) / 256;
	if(blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
	k_row_planes<<<blocks, 256, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->chunks, slot, (uint4 *) d_buf, restore);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_row_proxi(ccg_ctx *ctx, int row_slot, const void *d_rowraw, const EpilogueParams &ep) {
	if(row_slot <= 0) return cudaSuccess;
	k_row_proxi<<<(unsigned) ((row_slot + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->chunks, (long long) ctx->len,
	                                                                        ctx->proxi, ctx->proxi_snp_only, (const uint4 *) d_rowraw,
	                                                                        row_slot, ep);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_sample_proxi(ccg_ctx *ctx, int vs_ref, int ref_slot, const unsigned char *d_use, int apply,
                                    unsigned *d_cleared) {
	if(ctx->words == 0) return cudaSuccess;
	const int nseg = (ctx->chunks + SEG_CHUNKS - 1) / SEG_CHUNKS;
	dim3 grid((unsigned) (ctx->n_pad / 32), (unsigned) ((nseg + 3) / 4));
	k_sample_proxi<<<grid, dim3(32, 4), 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->chunks, ctx->words, (long long) ctx->len,
	                                                    ctx->proxi, ctx->proxi_snp_only, vs_ref, ref_slot, d_use, apply, d_cleared,
	                                                    ctx->d_gmask);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_count_mask(ccg_ctx *ctx, unsigned *d_count) {
	if(ctx->words == 0) return cudaSuccess;
	int blocks = (ctx->words + 255) / 256;
	if(blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
	k_count_mask<<<blocks, 256, 0, ctx->stream>>>(ctx->d_gmask, ctx->words, d_count);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_pair_proxi(ccg_ctx *ctx, const ProxiParams &p) {
	if(p.ntiles <= 0) return cudaSuccess;
	k_pairdist_proxi<<<(unsigned) p.ntiles * PX_SUB, 256, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->chunks, ctx->proxi, p);
	ctx->launches++;
	return cudaGetLastError();
}
