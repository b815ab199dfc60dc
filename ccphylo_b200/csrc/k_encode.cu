/*
 * k_encode.cu -- K1: build the device-resident bit-plane sample store.
 *
 * Replaces, per sample, the reference's host-side preparation for the compare
 * loop: qseq2nibble (qseqs.c:60-88), initIncPos (fsacmp.c:164-179),
 * getIncPos(seq, seq, 0) (fsacmp.c:181-238) and getNpos (fsacmp.c:487-503).
 *
 * Roofline: HBM streaming.  Algorithmic bytes per 32-base word and sample:
 * repack reads 8 B (u64 codes) + 4 B (u32 mask) and writes 12 B of planes
 * (coalesced both ways through a shared-memory transpose);
 * encode_codes reads 32 B of codes and writes 12 B.
 */
#include "ccg_internal.h"

/* keep the even bits of x (bit 2k -> bit k) */
__device__ __forceinline__ uint32_t compress_even(uint64_t x) {
	x &= 0x5555555555555555ull;
	x = (x | (x >> 1)) & 0x3333333333333333ull;
	x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
	x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
	x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
	x = (x | (x >> 16));
	return (uint32_t) x;
}

__device__ __forceinline__ void store_chunk(uint32_t *planes, int n_pad, int nplanes, long long chunk, int slot,
                                            const uint32_t h[4], const uint32_t l[4], const uint32_t m[4]) {
	uint4 *base = reinterpret_cast<uint4 *>(planes);
	size_t row = (size_t) chunk * nplanes;
	base[(row + 0) * n_pad + slot] = make_uint4(h[0], h[1], h[2], h[3]);
	base[(row + 1) * n_pad + slot] = make_uint4(l[0], l[1], l[2], l[3]);
	if(nplanes == 3) base[(row + 2) * n_pad + slot] = make_uint4(m[0], m[1], m[2], m[3]);
}

/* Reference packed format -> planes, plus getNpos of every mask row (fsacmp.c:487).
 * A block transposes a tile of 32 samples x 16 chunks through shared memory: the reads
 * walk each sample row contiguously (half-warp = one sample, lane = chunk: 32 B of codes +
 * 16 B of mask per lane), the writes walk each (chunk, plane) row contiguously (lane =
 * sample: 16 B per lane). */
__global__ void __launch_bounds__(256)
k_repack_packed(uint32_t *__restrict__ planes, int n_pad, int nplanes, int chunks, int words, int first, int count,
                const uint64_t *__restrict__ seqs, const uint32_t *__restrict__ masks,
                const uint32_t *__restrict__ gmask, long wstride, unsigned *__restrict__ inc) {
	__shared__ uint4 tile[3][16][33];                  /* [plane][chunk][sample], padded against bank conflicts */
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int cl = lane & 15;
	const long long c0 = (long long) blockIdx.x * 16;
	const int s0 = blockIdx.y * 32;

	for(int si = warp * 2 + (lane >> 4); si < 32; si += 16) {
		const int s = s0 + si;
		const long long c = c0 + cl;
		uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0};
		unsigned known = 0;
		if(s < count && c < chunks) {
			const uint64_t *srow = seqs + (size_t) s * wstride;
			const uint32_t *mrow = masks ? masks + (size_t) s * wstride : gmask;
#pragma unroll
			for(int q = 0; q < 4; ++q) {
				const long long w = c * 4 + q;
				if(w < words) {
					const uint64_t x = __ldg(srow + w);
					const uint32_t mk = __ldg(mrow + w);
					m[q] = mk;
					h[q] = compress_even(x >> 1) & mk;
					l[q] = compress_even(x) & mk;
					known += __popc(mk);
				}
			}
		}
		tile[0][cl][si] = make_uint4(h[0], h[1], h[2], h[3]);
		tile[1][cl][si] = make_uint4(l[0], l[1], l[2], l[3]);
		if(nplanes == 3) tile[2][cl][si] = make_uint4(m[0], m[1], m[2], m[3]);
		if(masks) {
#pragma unroll
			for(int o = 8; o; o >>= 1) known += __shfl_xor_sync(0xffffffffu, known, o);   /* within the half-warp */
			if(cl == 0 && s < count && known) atomicAdd(inc + first + s, known);
		}
	}
	__syncthreads();
	uint4 *out = reinterpret_cast<uint4 *>(planes);
	for(int r = warp; r < 16 * nplanes; r += 8) {
		const int ci = r / nplanes, pl = r % nplanes;
		const long long c = c0 + ci;
		const int s = s0 + lane;
		if(c < chunks && s < count) out[((size_t) c * nplanes + pl) * n_pad + first + s] = tile[pl][ci][lane];
	}
}

/* Translated codes (0..3 base, anything else unknown) -> planes + inc count.
 * One thread per (chunk, sample): 128 codes in, three 16-byte plane words out. */
__global__ void __launch_bounds__(256)
k_encode_codes(uint32_t *__restrict__ planes, int n_pad, int nplanes, int chunks, int len, int first, int count,
               const unsigned char *__restrict__ codes, long stride, unsigned *__restrict__ inc) {
	long long gid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	long long total = (long long) chunks * count;
	if(gid >= total) return;
	int s = (int) (gid % count);
	long long c = gid / count;
	const unsigned char *row = codes + (size_t) s * stride;
	uint32_t h[4], l[4], m[4];
	unsigned known = 0;
#pragma unroll
	for(int q = 0; q < 4; ++q) {
		long long p0 = c * CCG_CHUNK_BASES + q * 32;
		uint32_t hh = 0, ll = 0, mm = 0;
		if(p0 + 32 <= len) {
			/* stride and row base are 16-byte aligned by the staging allocator */
			const uint4 *v = reinterpret_cast<const uint4 *>(row + p0);
			uint4 a = v[0], b = v[1];
			uint32_t wv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
			for(int k = 0; k < 32; ++k) {
				uint32_t code = (wv[k >> 2] >> (8 * (k & 3))) & 0xFF;
				uint32_t kn = code < 4 ? 1u : 0u;
				mm |= kn << (31 - k);
				hh |= (kn & (code >> 1)) << (31 - k);
				ll |= (kn & code) << (31 - k);
			}
		} else {
			for(int k = 0; k < 32; ++k) {
				long long p = p0 + k;
				uint32_t code = p < len ? row[p] : 4;
				uint32_t kn = code < 4 ? 1u : 0u;
				mm |= kn << (31 - k);
				hh |= (kn & (code >> 1)) << (31 - k);
				ll |= (kn & code) << (31 - k);
			}
		}
		h[q] = hh; l[q] = ll; m[q] = mm;
		known += __popc(mm);
	}
	store_chunk(planes, n_pad, nplanes, c, first + s, h, l, m);
	if(known) atomicAdd(inc + first + s, known);
}

/* Shared-mask mode on a pair-mode store (cdist.c:101-112 followed by cmpFsaThrd): every plane
 * word of every slot is ANDed with the global mask G, so the per-pair mask m_i & m_j equals G
 * and the pair kernels count exactly the reference's shared-mask mismatches. */
__global__ void __launch_bounds__(256)
k_apply_global_mask(uint32_t *__restrict__ planes, int n_pad, int nplanes, int chunks, int words,
                    const uint32_t *__restrict__ gmask) {
	const long long total = (long long) chunks * nplanes * n_pad;
	uint4 *v = reinterpret_cast<uint4 *>(planes);
	for(long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) {
		const long long c = e / ((long long) nplanes * n_pad);
		uint32_t g[4];
#pragma unroll
		for(int q = 0; q < 4; ++q) g[q] = (c * 4 + q < words) ? gmask[c * 4 + q] : 0u;
		uint4 x = v[e];
		x.x &= g[0]; x.y &= g[1]; x.z &= g[2]; x.w &= g[3];
		v[e] = x;
	}
}

cudaError_t ccg_launch_apply_global_mask(ccg_ctx *ctx) {
	k_apply_global_mask<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->nplanes, ctx->chunks,
	                                                                   ctx->words, ctx->d_gmask);
	ctx->launches++;
	return cudaGetLastError();
}

/* tile-major raw counts -> packed lower triangle over included samples */
__global__ void __launch_bounds__(256)
k_gather_raw(const uint32_t *__restrict__ acc, int ntiles, const int2 *__restrict__ tiles, const int *__restrict__ rank,
             uint32_t *__restrict__ mism, uint32_t *__restrict__ ninc) {
	int lt = blockIdx.x;
	if(lt >= ntiles) return;
	const int ti = tiles[lt].x, tj = tiles[lt].y;
	const uint32_t *a = acc + (size_t) lt * 2 * CCG_TILE * CCG_TILE;
	for(int e = threadIdx.x; e < CCG_TILE * CCG_TILE; e += blockDim.x) {
		int i = ti * CCG_TILE + e / CCG_TILE;
		int j = tj * CCG_TILE + e % CCG_TILE;
		if(i <= j) continue;
		int r = rank[i], c = rank[j];
		if(r < 0 || c < 0) continue;
		long long cell = (long long) r * (r - 1) / 2 + c;
		if(mism) mism[cell] = a[e];
		if(ninc) ninc[cell] = a[CCG_TILE * CCG_TILE + e];
	}
}

cudaError_t ccg_launch_repack(ccg_ctx *ctx, int first, int count, const uint64_t *d_seqs, const uint32_t *d_masks,
                              long wstride) {
	if(count <= 0) return cudaSuccess;
	if(d_masks) {
		cudaError_t e = cudaMemsetAsync(ctx->d_inc + first, 0, (size_t) count * sizeof(unsigned), ctx->stream);
		if(e != cudaSuccess) return e;
	}
	dim3 grid((unsigned) ((ctx->chunks + 15) / 16), (unsigned) ((count + 31) / 32));
	k_repack_packed<<<grid, 256, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->nplanes, ctx->chunks, ctx->words, first,
	                                                count, d_seqs, d_masks, ctx->d_gmask, wstride, ctx->d_inc);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_encode_codes(ccg_ctx *ctx, int first, int count, const unsigned char *d_codes, long stride) {
	long long total = (long long) ctx->chunks * count;
	if(total == 0) return cudaSuccess;
	cudaError_t e = cudaMemsetAsync(ctx->d_inc + first, 0, (size_t) count * sizeof(unsigned), ctx->stream);
	if(e != cudaSuccess) return e;
	unsigned blocks = (unsigned) ((total + 255) / 256);
	k_encode_codes<<<blocks, 256, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->nplanes, ctx->chunks, ctx->len,
	                                                 first, count, d_codes, stride, ctx->d_inc);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_gather_raw(ccg_ctx *ctx, uint32_t *d_mism, uint32_t *d_ninc) {
	if(ctx->last_ntiles <= 0) return cudaSuccess;
	k_gather_raw<<<ctx->last_ntiles, 256, 0, ctx->stream>>>(ctx->d_acc, ctx->last_ntiles, ctx->d_tiles, ctx->d_rank,
	                                                        d_mism, d_ninc);
	ctx->launches++;
	return cudaGetLastError();
}
