/*
 * k_motif.cu -- -y / --methylation_motifs on the device: maskMotifs (meth.c:141-159) for freshly uploaded
 * samples.  The reference slides every motif (and its reverse complement) over the PACKED sequence of a sample
 * -- unknown bases are packed as A, qseqs.c:60 -- one start position at a time (matchMotif meth.c:76-125,
 * matchMotif32 :52-66) and clears the mask bits of the motif's upper-case positions at every match (maskMotif
 * :127-139).
 *
 * Here the match is evaluated for 32 start positions at once on the bit planes: with the base-indicator words
 * A = ~h & ~l, C = ~h & l, G = h & ~l, T = h & l of two neighbouring 32-base words side by side in 64 bits,
 *     starts = AND over motif positions k of (E_k << k),   E_k = OR of the indicators of the bases position k accepts
 * and the bits to clear are OR over the upper-case positions k of (starts >> k), which run over into the next
 * word.  One thread owns one 16-byte plane word (4 x 32 bases) of one sample: it evaluates the five 32-base words
 * whose matches can reach into it and rewrites its word of the MASK plane only; the code planes, which every
 * neighbour is reading, are re-masked later by an element-wise kernel (k_remask_all, before the first run).  Lanes of a warp are neighbouring samples
 * (coalesced 512 B rows).  HBM-bound: 3 x 32 B read + 16 B written per sample and 128 bases.
 */
#include "ccg_internal.h"

namespace {

struct Ind64 {
	unsigned long long a, c, g, t;
};

__device__ __forceinline__ Ind64 indicators(uint32_t h0, uint32_t l0, uint32_t h1, uint32_t l1) {
	const unsigned long long H = ((unsigned long long) h0 << 32) | h1, L = ((unsigned long long) l0 << 32) | l1;
	Ind64 r;
	r.a = ~H & ~L;
	r.c = ~H & L;
	r.g = H & ~L;
	r.t = H & L;
	return r;
}

/* bits to clear in words x (upper half) and x + 1 (lower half) from the matches that start in word x */
__device__ __forceinline__ unsigned long long clear_bits(const Ind64 &ind, long long x, long long len, int nmotifs,
                                                         const int *__restrict__ lens, const unsigned char *__restrict__ sets) {
	unsigned long long clr = 0;
	const long long room = len - x * 32;            /* bases from the start of word x to the end of the alignment */
	for(int m = 0; m < nmotifs; ++m) {
		const int L = lens[m];
		/* start b of word x is allowed when b + L <= room */
		const long long last = room - L;
		if(last >= 0) {
			unsigned long long starts = last >= 31 ? 0xFFFFFFFF00000000ull : (0xFFFFFFFFFFFFFFFFull << (63 - (int) last));
			unsigned long long sites = 0;
			for(int k = 0; k < L && starts; ++k) {
				const unsigned s = sets[k];
				const unsigned long long e = ((s & 1u) ? ind.a : 0ull) | ((s & 2u) ? ind.c : 0ull) | ((s & 4u) ? ind.g : 0ull) |
				                             ((s & 8u) ? ind.t : 0ull);
				starts &= e << k;
			}
			if(starts) {
				for(int k = 0; k < L; ++k)
					if(sets[k] & 16u) sites |= starts >> k;
				clr |= sites;
			}
		}
		sets += L;
	}
	return clr;
}

__global__ void __launch_bounds__(128)
k_motif_mask(uint32_t *planes, int n_pad, int chunks, long long len, int first, int count, int nmotifs,
             const int *__restrict__ lens_g, const unsigned char *__restrict__ sets_g, int nsets, unsigned *__restrict__ removed) {
	extern __shared__ unsigned char smem[];
	int *lens = reinterpret_cast<int *>(smem);
	unsigned char *sets = smem + (size_t) nmotifs * sizeof(int);
	for(int k = threadIdx.x + threadIdx.y * blockDim.x; k < nmotifs; k += blockDim.x * blockDim.y) lens[k] = lens_g[k];
	for(int k = threadIdx.x + threadIdx.y * blockDim.x; k < nsets; k += blockDim.x * blockDim.y) sets[k] = sets_g[k];
	__syncthreads();
	const int s = blockIdx.x * 32 + threadIdx.x;
	if(s >= count) return;
	const int slot = first + s;
	uint4 *P = reinterpret_cast<uint4 *>(planes);
	const uint4 zero = make_uint4(0, 0, 0, 0);
	unsigned gone_all = 0;
	for(int ch = blockIdx.y * blockDim.y + threadIdx.y; ch < chunks; ch += gridDim.y * blockDim.y) {
	const size_t row = (size_t) ch * 3;
	const uint4 h = P[(row + 0) * n_pad + slot], l = P[(row + 1) * n_pad + slot];
	const uint4 hp = ch > 0 ? P[(row - 3) * n_pad + slot] : zero, lp = ch > 0 ? P[(row - 2) * n_pad + slot] : zero;
	const uint4 hn = ch + 1 < chunks ? P[(row + 3) * n_pad + slot] : zero, ln = ch + 1 < chunks ? P[(row + 4) * n_pad + slot] : zero;
	/* words w0 - 1 .. w0 + 4 */
	const uint32_t hw[6] = {hp.w, h.x, h.y, h.z, h.w, hn.x}, lw[6] = {lp.w, l.x, l.y, l.z, l.w, ln.x};
	const long long w0 = (long long) ch * CCG_CHUNK_WORDS;
	uint32_t clr[4] = {0, 0, 0, 0};
#pragma unroll
	for(int q = 0; q < 5; ++q) {
		const long long x = w0 - 1 + q;                 /* matches starting in word x reach words x and x + 1 */
		if(x < 0) continue;
		const Ind64 ind = indicators(hw[q], lw[q], hw[q + 1], lw[q + 1]);
		const unsigned long long c = clear_bits(ind, x, len, nmotifs, lens, sets);
		if(q >= 1) clr[q - 1] |= (uint32_t) (c >> 32);
		if(q <= 3) clr[q] |= (uint32_t) c;
	}
	uint4 m = P[(row + 2) * n_pad + slot];
	const unsigned gone = (unsigned) (__popc(m.x & clr[0]) + __popc(m.y & clr[1]) + __popc(m.z & clr[2]) + __popc(m.w & clr[3]));
	if(gone) {
		m.x &= ~clr[0]; m.y &= ~clr[1]; m.z &= ~clr[2]; m.w &= ~clr[3];
		P[(row + 2) * n_pad + slot] = m;
		gone_all += gone;
	}
	}
	if(gone_all) atomicAdd(removed + s, gone_all);
}

/* included counts -= removed.  The CODE planes keep the bases of the methylation sites for now: the reference changes
 * only the inclusion mask (maskMotif meth.c:127-139), and what still looks at the sequences afterwards -- the variant
 * listing of -V (fsacmpairint / fsacmprint compare whole packed words) -- must see them unchanged.  The compare
 * kernels want code planes that are zero wherever the mask is: k_remask_all does that right before the first run. */
__global__ void __launch_bounds__(128)
k_motif_counts(int first, int count, const unsigned *__restrict__ removed, unsigned *__restrict__ inc) {
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s < count && removed[s]) inc[first + s] -= removed[s];
}

__global__ void __launch_bounds__(256)
k_remask_all(uint32_t *planes, int n_pad, int chunks) {
	uint4 *P = reinterpret_cast<uint4 *>(planes);
	const long long total = (long long) chunks * n_pad;
	for(long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) {
		const long long ch = e / n_pad, slot = e % n_pad;
		const size_t row = (size_t) ch * 3;
		const uint4 m = P[(row + 2) * n_pad + slot];
		uint4 h = P[(row + 0) * n_pad + slot], l = P[(row + 1) * n_pad + slot];
		const uint4 h2 = make_uint4(h.x & m.x, h.y & m.y, h.z & m.z, h.w & m.w), l2 = make_uint4(l.x & m.x, l.y & m.y, l.z & m.z, l.w & m.w);
		if(h2.x != h.x || h2.y != h.y || h2.z != h.z || h2.w != h.w) P[(row + 0) * n_pad + slot] = h2;
		if(l2.x != l.x || l2.y != l.y || l2.z != l.z || l2.w != l.w) P[(row + 1) * n_pad + slot] = l2;
	}
}

} // namespace

cudaError_t ccg_launch_motif_mask(ccg_ctx *ctx, int first, int count, unsigned *d_removed) {
	if(count <= 0 || ctx->words == 0 || ctx->motif_n == 0) return cudaSuccess;
	unsigned gy = (unsigned) ((ctx->chunks + 3) / 4);
	if(gy > 32768u) gy = 32768u;
	dim3 block(32, 4), grid((unsigned) ((count + 31) / 32), gy);
	const size_t smem = (size_t) ctx->motif_n * sizeof(int) + (size_t) ctx->motif_nsets;
	k_motif_mask<<<grid, block, smem, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->chunks, (long long) ctx->len, first, count,
	                                                 ctx->motif_n, ctx->d_motif_lens, ctx->d_motif_sets, ctx->motif_nsets, d_removed);
	ctx->launches++;
	cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return e;
	k_motif_counts<<<(unsigned) ((count + 127) / 128), 128, 0, ctx->stream>>>(first, count, d_removed, ctx->d_inc);
	ctx->launches++;
	ctx->remask_pending = 1;
	return cudaGetLastError();
}

/* code planes &= mask plane over the whole (three-plane) store: before the first run after ccg_mask_motifs */
cudaError_t ccg_launch_remask_all(ccg_ctx *ctx) {
	if(ctx->nplanes != 3 || ctx->words == 0) return cudaSuccess;
	k_remask_all<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->d_planes, ctx->n_pad, ctx->chunks);
	ctx->launches++;
	return cudaGetLastError();
}
