/*
 * k_variants.cu -- -V / --nucleotide_variations: the per-pair variant lists the reference prints while it
 * compares (fsacmpairint fsacmp.c:685-737 in pair mode, fsacmprint :646-683 in shared-mask mode, printDiff
 * :635-644), extracted on the device from the resident bit planes.
 *
 * The list of a pair is ordered along the alignment and its position labels come from a running counter (see
 * below), so one thread walks one pair from the first to the last word; the parallelism is over pairs.  A batch
 * is a run of consecutive cells of the packed lower triangle: consecutive cells share their row sample and have
 * consecutive column samples, so the lanes of a warp read neighbouring 16-byte plane words (the column) and one
 * broadcast word (the row).  Two passes over a batch: count, (host: prefix sum), write.  HBM-bound on the plane
 * reads: 2 x 48 B per pair and 128 bases.
 *
 * Labels, as the reference produces them (not alignment coordinates): a counter starts at 1; a
 * word is "walked" when its mask word is non-zero and the two packed words differ anywhere; in a walked word lane
 * k (counted from the LEAST significant end: base 31 - k of the word, plane / mask bit k) is labelled counter + k
 * and the counter then advances by (index of the highest set mask bit + 1); any other word advances it by 32.
 *
 * With -P (pair mode) the reference walks the pair's proximity mask instead (maskProxi, then fsacmpairint on its
 * output, fsacmpthrd.c:410-414), and the labels depend on every word of that mask: k_pair_proxi_mask builds it for a
 * batch of cells into a scratch buffer [word][cell] (the walk of proxi_core.h: proxi_pair_mask_word), and both passes
 * of k_variants read their mask words from there.
 */
#include "ccg_internal.h"
#include "proxi_core.h"

namespace {

__device__ __forceinline__ void cell_to_pair(long long cell, int &r, int &c) {
	/* cell = r (r - 1) / 2 + c, 0 <= c < r */
	long long rr = (long long) ((1.0 + sqrt(1.0 + 8.0 * (double) cell)) * 0.5);
	while(rr * (rr - 1) / 2 > cell) --rr;
	while((rr + 1) * rr / 2 <= cell) ++rr;
	r = (int) rr;
	c = (int) (cell - rr * (rr - 1) / 2);
}

template <bool WRITE>
__global__ void __launch_bounds__(128)
k_variants(const uint32_t *__restrict__ planes, int n_pad, int chunks, int words, const uint32_t *__restrict__ gmask,
           VariantParams p) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= p.ncells) return;
	int r, c;
	cell_to_pair(p.cell0 + t, r, c);
	const int i = p.slot_of_rank[r], j = p.slot_of_rank[c];
	const uint4 *P = reinterpret_cast<const uint4 *>(planes);
	unsigned label = 1, count = 0;
	unsigned long long *out = WRITE ? p.entries + (p.offsets[t] - p.offsets[0]) : 0;
#pragma unroll 1
	for(int ch = 0; ch < chunks; ++ch) {
		const size_t row = (size_t) ch * 3;
		/* (the row sample of -a with -P: its codes as uploaded -- the store's copy is proximity-masked for the run) */
		const uint4 ih = p.row_raw ? __ldg(p.row_raw + row) : __ldg(P + (row + 0) * n_pad + i);
		const uint4 il = p.row_raw ? __ldg(p.row_raw + row + 1) : __ldg(P + (row + 1) * n_pad + i);
		const uint4 jh = __ldg(P + (row + 0) * n_pad + j), jl = __ldg(P + (row + 1) * n_pad + j);
		uint4 m;
		if(p.pair_mask) {
			const uint32_t *pm = p.pair_mask + (size_t) ch * CCG_CHUNK_WORDS * p.pair_mask_stride + t;
			const int w0 = ch * CCG_CHUNK_WORDS;
			m.x = w0 < words ? pm[0] : 0u;
			m.y = w0 + 1 < words ? pm[p.pair_mask_stride] : 0u;
			m.z = w0 + 2 < words ? pm[2 * p.pair_mask_stride] : 0u;
			m.w = w0 + 3 < words ? pm[3 * p.pair_mask_stride] : 0u;
		} else if(gmask) {
			const int w0 = ch * CCG_CHUNK_WORDS;
			m.x = w0 < words ? __ldg(gmask + w0) : 0u;
			m.y = w0 + 1 < words ? __ldg(gmask + w0 + 1) : 0u;
			m.z = w0 + 2 < words ? __ldg(gmask + w0 + 2) : 0u;
			m.w = w0 + 3 < words ? __ldg(gmask + w0 + 3) : 0u;
		} else {
			const uint4 im = __ldg(P + (row + 2) * n_pad + i), jm = __ldg(P + (row + 2) * n_pad + j);
			m = make_uint4(im.x & jm.x, im.y & jm.y, im.z & jm.z, im.w & jm.w);
		}
#define CCG_V_WORD(f)                                                                          \
	{                                                                                          \
		const uint32_t inc = m.f;                                                              \
		const uint32_t dh = ih.f ^ jh.f, dl = il.f ^ jl.f;                                     \
		if(inc && (dh | dl)) {                                                                 \
			uint32_t snp = (dh | dl) & inc;                                                    \
			if(WRITE) {                                                                        \
				while(snp) {                                                                   \
					const int k = __ffs((int) snp) - 1;                                        \
					snp &= snp - 1u;                                                           \
					const unsigned ci = ((ih.f >> k) & 1u) << 1 | ((il.f >> k) & 1u);          \
					const unsigned cj = ((jh.f >> k) & 1u) << 1 | ((jl.f >> k) & 1u);          \
					out[count++] = ((unsigned long long) (label + (unsigned) k) << 4) | (ci << 2) | cj; \
				}                                                                              \
			} else count += (unsigned) __popc(snp);                                            \
			label += 32u - (unsigned) __clz((int) inc);                                        \
		} else label += 32u;                                                                   \
	}
		CCG_V_WORD(x) CCG_V_WORD(y) CCG_V_WORD(z) CCG_V_WORD(w)
#undef CCG_V_WORD
	}
	if(!WRITE) p.counts[t] = count;
}

/* -V with -P: maskProxi's output for every cell of the batch.  One thread per cell: first the unmasked pair mask
 * inc_i & inc_j goes into the cell's column of the scratch buffer (coalesced over the cells), then the pair is walked
 * once more and every two neighbouring SNPs at most proxi apart clear their range in it.  The walk never reaches
 * beyond the word after the current one, and all words are in place before it starts. */
struct PairMaskSink {
	uint32_t *col;                      /* pair_mask + t */
	long long stride;
	__device__ __forceinline__ void clear(long long w, uint32_t bits) {
		uint32_t *m = col + (size_t) w * stride;
		const uint32_t old = *m;
		if(old & bits) *m = old & ~bits;
	}
};

__global__ void __launch_bounds__(128)
k_pair_proxi_mask(const uint32_t *__restrict__ planes, int n_pad, int chunks, int words, unsigned proxi, VariantParams p) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= p.ncells) return;
	int r, c;
	cell_to_pair(p.cell0 + t, r, c);
	const int i = p.slot_of_rank[r], j = p.slot_of_rank[c];
	const uint4 *P = reinterpret_cast<const uint4 *>(planes);
	PairMaskSink sink = {p.pair_mask + t, p.pair_mask_stride};
#pragma unroll 1
	for(int ch = 0; ch < chunks; ++ch) {
		const size_t row = (size_t) ch * 3;
		const uint4 im = __ldg(P + (row + 2) * n_pad + i), jm = __ldg(P + (row + 2) * n_pad + j);
		const int w0 = ch * CCG_CHUNK_WORDS;
		if(w0 < words) sink.col[(size_t) w0 * sink.stride] = im.x & jm.x;
		if(w0 + 1 < words) sink.col[(size_t) (w0 + 1) * sink.stride] = im.y & jm.y;
		if(w0 + 2 < words) sink.col[(size_t) (w0 + 2) * sink.stride] = im.z & jm.z;
		if(w0 + 3 < words) sink.col[(size_t) (w0 + 3) * sink.stride] = im.w & jm.w;
	}
	long long last = -1;
#pragma unroll 1
	for(int ch = 0; ch < chunks; ++ch) {
		const size_t row = (size_t) ch * 3;
		const uint4 ih = __ldg(P + (row + 0) * n_pad + i), il = __ldg(P + (row + 1) * n_pad + i), im = __ldg(P + (row + 2) * n_pad + i);
		const uint4 jh = __ldg(P + (row + 0) * n_pad + j), jl = __ldg(P + (row + 1) * n_pad + j), jm = __ldg(P + (row + 2) * n_pad + j);
		const long long w0 = (long long) ch * CCG_CHUNK_WORDS;
		proxi_pair_mask_word(last, w0, ((ih.x ^ jh.x) | (il.x ^ jl.x)) & im.x & jm.x, words, proxi, sink);
		proxi_pair_mask_word(last, w0 + 1, ((ih.y ^ jh.y) | (il.y ^ jl.y)) & im.y & jm.y, words, proxi, sink);
		proxi_pair_mask_word(last, w0 + 2, ((ih.z ^ jh.z) | (il.z ^ jl.z)) & im.z & jm.z, words, proxi, sink);
		proxi_pair_mask_word(last, w0 + 3, ((ih.w ^ jh.w) | (il.w ^ jl.w)) & im.w & jm.w, words, proxi, sink);
	}
}

/* -V with -a and -P: the mask cmpFsaRowThrd walks for column sample j (fsacmpthrd.c:543-553) -- the new sample's own
 * mask (already through its own builder, fsacmpthrd.c:627-628) restricted to j's known positions, then the per-sample
 * builder of the new sample against j: everything between two events at most proxi apart goes, both ends included
 * (proxi_scan_words).  Events are defined on the codes as uploaded: p.row_raw for the new sample. */
struct RowMaskEvents {
	const uint32_t *planes, *raw;
	int n_pad, j, snp_only;
	long long len;
	__device__ __forceinline__ uint32_t operator()(long long w) const {
		const size_t pj = ((size_t) (w >> 2) * 3 * n_pad + j) * 4 + (size_t) (w & 3), step = (size_t) n_pad * 4;
		const size_t pr = (size_t) (w >> 2) * 12 + (size_t) (w & 3);
		return proxi_events(1, snp_only, planes[pj + 2 * step], planes[pj], planes[pj + step], raw[pr + 8], raw[pr], raw[pr + 4],
		                    proxi_valid_bits(len, w * 32));
	}
};

__global__ void __launch_bounds__(128)
k_row_proxi_mask(const uint32_t *__restrict__ planes, int n_pad, int words, long long len, unsigned proxi, int snp_only, int row_slot,
                 VariantParams p) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= p.ncells) return;
	int r, c;
	cell_to_pair(p.cell0 + t, r, c);
	const int j = p.slot_of_rank[c];
	PairMaskSink sink = {p.pair_mask + t, p.pair_mask_stride};
	const size_t step = (size_t) n_pad * 4;
#pragma unroll 1
	for(int w = 0; w < words; ++w) {
		const size_t base = ((size_t) (w >> 2) * 3 * n_pad) * 4 + (size_t) (w & 3) + 2 * step;
		sink.col[(size_t) w * sink.stride] = planes[base + (size_t) row_slot * 4] & planes[base + (size_t) j * 4];
	}
	RowMaskEvents ev = {planes, reinterpret_cast<const uint32_t *>(p.row_raw), n_pad, j, snp_only, len};
	proxi_scan_words(-1, 0, words, proxi, ev, sink);
}

} // namespace

cudaError_t ccg_launch_row_proxi_mask(ccg_ctx *ctx, const VariantParams &p, int row_slot) {
	if(p.ncells <= 0 || ctx->words == 0) return cudaSuccess;
	k_row_proxi_mask<<<(unsigned) ((p.ncells + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->words, (long long) ctx->len,
	                                                                              ctx->proxi, ctx->proxi_snp_only, row_slot, p);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_pair_proxi_mask(ccg_ctx *ctx, const VariantParams &p) {
	if(p.ncells <= 0 || ctx->words == 0) return cudaSuccess;
	k_pair_proxi_mask<<<(unsigned) ((p.ncells + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->chunks, ctx->words,
	                                                                               ctx->proxi, p);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_variants(ccg_ctx *ctx, const VariantParams &p, int write, int shared_mask) {
	if(p.ncells <= 0) return cudaSuccess;
	const unsigned blocks = (unsigned) ((p.ncells + 127) / 128);
	const uint32_t *g = shared_mask ? ctx->d_gmask : 0;
	if(write) k_variants<true><<<blocks, 128, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->chunks, ctx->words, g, p);
	else k_variants<false><<<blocks, 128, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->chunks, ctx->words, g, p);
	ctx->launches++;
	return cudaGetLastError();
}
