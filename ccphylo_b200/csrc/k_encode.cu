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
/* The rows may hold only the words of chunks [chunk0, chunk0 + nch) (K-slab streaming of host
 * rows, see feed_slab in ccg_api.cu): seqs / masks rows then start at word 4 * chunk0. */
__global__ void __launch_bounds__(256)
k_repack_packed(uint32_t *__restrict__ planes, int n_pad, int nplanes, int chunk0, int nch, int words, int first, int count,
                const uint64_t *__restrict__ seqs, const uint32_t *__restrict__ masks,
                const uint32_t *__restrict__ gmask, long wstride, unsigned *__restrict__ inc, int keep_codes) {
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
		if(s < count && c < nch) {
			const uint64_t *srow = seqs + (size_t) s * wstride;
			const uint32_t *mrow = masks ? masks + (size_t) s * wstride : gmask + (size_t) chunk0 * 4;
#pragma unroll
			for(int q = 0; q < 4; ++q) {
				const long long w = c * 4 + q;                    /* word within the rows; absolute word = 4 chunk0 + w */
				if((long long) chunk0 * 4 + w < words) {
					const uint64_t x = __ldg(srow + w);
					const uint32_t mk = __ldg(mrow + w);
					m[q] = mk;
					/* keep_codes: the rows' masks were narrowed by the caller (-P / -y on the host) and a variant listing
					 * will compare whole words as the reference does; k_remask_all clears the codes before the first run */
					const uint32_t cm = keep_codes ? 0xFFFFFFFFu : mk;
					h[q] = compress_even(x >> 1) & cm;
					l[q] = compress_even(x) & cm;
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
		if(c < nch && s < count) out[((size_t) (chunk0 + c) * nplanes + pl) * n_pad + first + s] = tile[pl][ci][lane];
	}
}

/* Same conversion without shared memory, for the K-slab streaming path (feed_slab): it runs on
 * a copy stream BESIDE the persistent tensor kernel, which leaves ~30 KiB of shared memory per
 * SM -- the transposing kernel above would get one block per SM and stall the uploads behind
 * it.  Lane = sample: every lane streams its own row (one 32-byte sector of codes + 16 bytes of
 * mask per chunk) and the 32 lanes of a warp write 512 contiguous bytes per (chunk, plane).
 * Rows hold the words of chunks [chunk0, chunk0 + nch) at a pitch of wstride words (a multiple
 * of 4); nvalid = number of valid words per row. */
constexpr int REPACK_CPT = 8;      /* chunks per thread */
__global__ void __launch_bounds__(256)
k_repack_direct(uint32_t *__restrict__ planes, int n_pad, int nplanes, int chunk0, int nch, int nvalid, int first, int count,
                const uint64_t *__restrict__ seqs, const uint32_t *__restrict__ masks, const uint32_t *__restrict__ gmask,
                long wstride, unsigned *__restrict__ inc) {
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int s = blockIdx.y * 32 + lane;
	const int cbeg = (blockIdx.x * 8 + warp) * REPACK_CPT;
	if(s >= count || cbeg >= nch) return;
	const uint64_t *srow = seqs + (size_t) s * wstride;
	const uint32_t *mrow = masks ? masks + (size_t) s * wstride : gmask + (size_t) chunk0 * 4;
	uint4 *out = reinterpret_cast<uint4 *>(planes);
	unsigned known = 0;
	const int cend = cbeg + REPACK_CPT < nch ? cbeg + REPACK_CPT : nch;
#pragma unroll 4
	for(int c = cbeg; c < cend; ++c) {
		uint64_t x[4] = {0, 0, 0, 0};
		uint32_t mk[4] = {0, 0, 0, 0};
		if(c * 4 + 4 <= nvalid) {
			const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(srow + c * 4));
			const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2 *>(srow + c * 4 + 2));
			x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
			if(masks) {
				const uint4 m4 = __ldg(reinterpret_cast<const uint4 *>(mrow + c * 4));
				mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w;
			} else {
#pragma unroll
				for(int q = 0; q < 4; ++q) mk[q] = __ldg(mrow + c * 4 + q);
			}
		} else {
			for(int q = 0; q < 4; ++q)
				if(c * 4 + q < nvalid) { x[q] = __ldg(srow + c * 4 + q); mk[q] = __ldg(mrow + c * 4 + q); }
		}
		uint32_t h[4], l[4];
#pragma unroll
		for(int q = 0; q < 4; ++q) {
			h[q] = compress_even(x[q] >> 1) & mk[q];
			l[q] = compress_even(x[q]) & mk[q];
			known += __popc(mk[q]);
		}
		const size_t row = (size_t) (chunk0 + c) * nplanes;
		out[(row + 0) * n_pad + first + s] = make_uint4(h[0], h[1], h[2], h[3]);
		out[(row + 1) * n_pad + first + s] = make_uint4(l[0], l[1], l[2], l[3]);
		if(nplanes == 3) out[(row + 2) * n_pad + first + s] = make_uint4(mk[0], mk[1], mk[2], mk[3]);
	}
	if(masks && known) atomicAdd(inc + first + s, known);
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

/* Global mask = AND of the inclusion masks of all included samples (cdist.c:101-112: every
 * included sample clears the positions it does not know).  One warp per chunk: the mask plane
 * row of a chunk is n_pad x 16 contiguous bytes; lanes stride over the slots. */
__global__ void __launch_bounds__(256)
k_build_global_mask(const uint32_t *__restrict__ planes, int n_pad, int chunks, int words, const unsigned char *__restrict__ use,
                    uint32_t *__restrict__ gmask, unsigned *__restrict__ inc, int and_into) {
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if(warp >= chunks) return;
	const uint4 *row = reinterpret_cast<const uint4 *>(planes) + ((size_t) warp * 3 + 2) * n_pad;
	uint4 g = make_uint4(~0u, ~0u, ~0u, ~0u);
	for(int s = lane; s < n_pad; s += 32) {
		if(!use[s]) continue;
		const uint4 m = row[s];
		g.x &= m.x; g.y &= m.y; g.z &= m.z; g.w &= m.w;
	}
#pragma unroll
	for(int o = 16; o; o >>= 1) {
		g.x &= __shfl_xor_sync(0xffffffffu, g.x, o);
		g.y &= __shfl_xor_sync(0xffffffffu, g.y, o);
		g.z &= __shfl_xor_sync(0xffffffffu, g.z, o);
		g.w &= __shfl_xor_sync(0xffffffffu, g.w, o);
	}
	if(lane == 0) {
		const uint32_t w[4] = {g.x, g.y, g.z, g.w};
		unsigned known = 0;
		for(int q = 0; q < 4; ++q)
			if(warp * 4 + q < words) {
				/* and_into: the mask already holds what earlier passes left (known-ness, proximity runs) */
				const uint32_t v = and_into ? (w[q] & gmask[warp * 4 + q]) : w[q];
				gmask[warp * 4 + q] = v;
				known += __popc(v);
			}
		if(known) atomicAdd(inc, known);
	}
}

cudaError_t ccg_launch_build_global_mask(ccg_ctx *ctx, const unsigned char *d_use, unsigned *d_inc, int and_into) {
	const long long threads = (long long) ctx->chunks * 32;
	k_build_global_mask<<<(unsigned) ((threads + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->chunks, ctx->words,
	                                                                                d_use, ctx->d_gmask, d_inc, and_into);
	ctx->launches++;
	return cudaGetLastError();
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
	return ccg_launch_repack_range(ctx, ctx->stream, first, count, d_seqs, d_masks, wstride, 0, ctx->chunks);
}

/* rows hold the words of chunks [chunk0, chunk0 + nch) only; the per-sample counts accumulate
 * (the caller zeroes d_inc once before the first range) */
/* shared-memory-free variant for work that runs beside the tensor kernel; wstride must be a multiple of 4 */
cudaError_t ccg_launch_repack_direct(ccg_ctx *ctx, cudaStream_t stream, int first, int count, const uint64_t *d_seqs,
                                     const uint32_t *d_masks, long wstride, int chunk0, int nch) {
	if(count <= 0 || nch <= 0) return cudaSuccess;
	int nvalid = ctx->words - chunk0 * CCG_CHUNK_WORDS;
	if(nvalid > nch * CCG_CHUNK_WORDS) nvalid = nch * CCG_CHUNK_WORDS;
	dim3 grid((unsigned) ((nch + 8 * REPACK_CPT - 1) / (8 * REPACK_CPT)), (unsigned) ((count + 31) / 32));
	k_repack_direct<<<grid, 256, 0, stream>>>(ctx->d_planes, ctx->n_pad, ctx->nplanes, chunk0, nch, nvalid, first, count, d_seqs,
	                                         d_masks, ctx->d_gmask, wstride, ctx->d_inc);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_repack_range(ccg_ctx *ctx, cudaStream_t stream, int first, int count, const uint64_t *d_seqs,
                                    const uint32_t *d_masks, long wstride, int chunk0, int nch) {
	if(count <= 0 || nch <= 0) return cudaSuccess;
	dim3 grid((unsigned) ((nch + 15) / 16), (unsigned) ((count + 31) / 32));
	/* with -P set, a pair-mode upload keeps the code bits of the positions its mask leaves out (see the kernel) */
	const int keep = d_masks && ctx->nplanes == 3 && ctx->proxi != 0;
	if(keep) ctx->remask_pending = 1;
	else if(d_masks) ctx->codes_upload_masked = 1;
	k_repack_packed<<<grid, 256, 0, stream>>>(ctx->d_planes, ctx->n_pad, ctx->nplanes, chunk0, nch, ctx->words, first, count,
	                                         d_seqs, d_masks, ctx->d_gmask, wstride, ctx->d_inc, keep);
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
