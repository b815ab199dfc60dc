/*
 * k_matdist.cu -- K5: all-vs-all distances over KMA count matrices (.mat inputs) and the C-ABI
 * around it (ccg_mat_*).
 *
 * Replaces the reference's cmpMats (matcmp.c:448-494) with the selected per-position vector
 * distance (the 18 veccmp functions of matcmp.c:63-446) and the pair loop around it
 * (ltdmatrixthrd.c:182 cmpMatThrd / :376 ltdMatrixThrd, ltdmatrix.c:32 ltdMatrix_get).  The
 * reference re-opens, inflates and re-parses sample j's file for every cell (i, j); here every
 * sample's template is parsed once on the host and kept resident in HBM.
 *
 * Device layout: counts[slot][position] = 16-byte record {A, C, G, T, -, N as u16; total as u32}
 * (the order the reference stores them in, matparse.c:254-259) plus norms[slot][position] = the
 * double sqrt(sum of squared counts) that `cos` needs, positions zero-padded to a multiple of 64.
 * Insertion rows (reference base '-') are dropped by the host parser.
 *
 * Per pair (i > j), over the positions p < len_j of the earlier sample (the one the reference
 * streams):   if(minDepth <= tot_j[p]) { ++nNucs;
 *                 if(minDepth <= tot_i[p] && 0 <= (d = veccmp(c_i[p], c_j[p], tot_i, tot_j))) { dist += d; ++rowsInc; } }
 * Gate: rowsInc < minLength || rowsInc < minCov * len_j  ->  D = -1, N = 0 ("No sufficient
 * overlap"); else D = norm ? dist / rowsInc * norm : dist, N = rowsInc.
 *
 * Work item = (16 x 16 tile of sample pairs, slice of positions).  A CTA stages 64 positions of
 * its 16 + 16 samples in shared memory (coalesced 16-byte loads along the position axis), every
 * thread owns one pair and walks the positions in order; the per-slice partial sums are written
 * to a [slice][tile][256] buffer and added in slice order by k_matdist_finalize, so the result
 * is deterministic (it differs from the reference's strictly sequential fp64 sum only by
 * rounding: the parity bar for this path is 1e-6 relative, counts are exact).
 * Roofline: CUDA-core FP64 + INT (sqrt / divide per position pair); HBM traffic is
 * 2 x 16 B x 16 samples per 256 position pairs.  -fmad=false (csrc/Makefile) keeps nvcc from
 * contracting the reference's multiply-then-add sequences.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ccg_internal.h"
#include "epilogue.cuh"

namespace {

constexpr int MT = 16;            /* tile edge in samples */
constexpr int MP = 64;            /* positions per stage (2 x 16 x 65 x 16 B = 33 KB of static shared memory) */

struct MatParams {
	unsigned minDepth;
	unsigned order;               /* l<n> / nl<n> */
	double alpha;                 /* z */
};

struct Rec {
	int c[6];                     /* A C G T - N */
	int tot;
};

__device__ __forceinline__ Rec unpack(const uint4 v) {
	Rec r;
	r.c[0] = v.x & 0xFFFF; r.c[1] = v.x >> 16;
	r.c[2] = v.y & 0xFFFF; r.c[3] = v.y >> 16;
	r.c[4] = v.z & 0xFFFF; r.c[5] = v.z >> 16;
	r.tot = (int) v.w;
	return r;
}

/* stdstat.c:132 p_chisqr with its table for q > 49 (stdstat.c:32-130) */
__device__ double fastp_dev(double q) {
	const double cut[16] = {114.5242, 109.9604, 105.3969, 100.8337, 96.27476, 91.71701, 87.16164, 82.60901, 78.05917,
	                        73.51245, 68.96954, 64.43048, 59.89615, 55.36699, 50.84417, 46.32844};
	const double val[16] = {1e-26, 1e-25, 1e-24, 1e-23, 1e-22, 1e-21, 1e-20, 1e-19, 1e-18, 1e-17, 1e-16, 1e-15, 1e-14,
	                        1e-13, 1e-12, 1e-11};
	for(int k = 0; k < 16; ++k)
		if(q > cut[k]) return val[k];
	return 1e-10;                 /* not reached: the caller only comes here for q > 49 */
}
__device__ double p_chisqr_dev(double q) {
	if(q < 0) return 1e-26;
	if(q > 49) return fastp_dev(q);
	return 1 - 1.772453850 * erf(sqrt(0.5 * q)) / 1.7724538509055160273;
}

/* the per-position distances; each follows the operation order of its reference function */
template <int M>
__device__ __forceinline__ double veccmp(const Rec &a, const Rec &b, const MatParams &mp) {
	if(M == CCG_MAT_COS) {                                   /* matcmp.c:420 */
		unsigned long long c1 = 0, c2 = 0;
		double d = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) {
			d += (double) (a.c[k] * b.c[k]);
			c1 += (unsigned long long) (long long) (a.c[k] * a.c[k]);
			c2 += (unsigned long long) (long long) (b.c[k] * b.c[k]);
		}
		if(!c1 || !c2) return -1;
		d = 1 - d / (sqrt((double) c1) * sqrt((double) c2));
		return d < 0 ? 0 : d;
	} else if(M == CCG_MAT_Z) {                              /* matcmp.c:311, bug-compatible */
		int max1 = a.c[0], max2 = b.c[0];
#pragma unroll
		for(int k = 1; k < 5; ++k) {
			if(max1 < a.c[k]) max1 = a.c[k];
			if(max2 < b.c[k]) max2 = b.c[k];
		}
		const double q1 = (double) (a.tot - (max1 << 1)) * (double) (a.tot - (max1 << 1)) / a.tot;
		const double q2 = (double) (b.tot - (max2 << 1)) * (double) (b.tot - (max2 << 1)) / b.tot;
		const int x1 = p_chisqr_dev(q1) <= mp.alpha && a.tot < (max1 << 1);
		const int x2 = p_chisqr_dev(q2) <= mp.alpha && a.tot < (max1 << 1);
		return (x1 && x2) ? 0 : -1;
	} else if(M == CCG_MAT_CHI2 || M == CCG_MAT_P) {         /* matcmp.c:383 / :346 */
		double d = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) {
			const double T = a.c[k] - b.c[k];
			if(T != 0) d += T * T / (a.c[k] + b.c[k]);
		}
		return M == CCG_MAT_CHI2 ? sqrt(d) : 1 - p_chisqr_dev(d);
	} else if(M == CCG_MAT_NCHI2 || M == CCG_MAT_NP) {       /* matcmp.c:398 / :361 */
		const int t1 = a.tot - a.c[5], t2 = b.tot - b.c[5];
		double d = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) {
			const double f1 = (double) a.c[k] / t1, f2 = (double) b.c[k] / t2;
			const double diff = f1 - f2;
			if(diff != 0) d += diff * diff / (f1 + f2);
		}
		return M == CCG_MAT_NCHI2 ? sqrt(d) : 1 - p_chisqr_dev(d);
	} else if(M == CCG_MAT_C) {                              /* matcmp.c:281 */
		double d = 0;
		int big = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) {
			if(a.c[k] < b.c[k]) { d += a.c[k]; big += b.c[k]; }
			else { d += b.c[k]; big += a.c[k]; }
		}
		if(!big) return -1;
		d = 1 - d / big;
		return d < 0 ? 0 : d;
	} else if(M == CCG_MAT_NC) {                             /* matcmp.c:246, bug-compatible: T restarts at 1 (#14) */
		const int t1 = a.tot - a.c[5], t2 = b.tot - b.c[5];
		double d = 0, T = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) {
			const double f1 = (double) a.c[k] / t1, f2 = (double) b.c[k] / t2;
			if(k) T = 1;
			if(f1 < f2) { d = k ? d + f1 : f1; T = k ? T + f2 : f2; }
			else { d = k ? d + f2 : f2; T = k ? T + f1 : f1; }
		}
		d = 1 - d / T;
		return d < 0 ? 0 : d;
	} else if(M == CCG_MAT_BC) {                             /* matcmp.c:230 */
		double d = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) d += a.c[k] < b.c[k] ? a.c[k] : b.c[k];
		d /= (a.tot - a.c[5] + b.tot - b.c[5]);
		d = 1 - 2 * d;
		return d < 0 ? 0 : d;
	} else if(M == CCG_MAT_NBC) {                            /* matcmp.c:209 */
		const int t1 = a.tot - a.c[5], t2 = b.tot - b.c[5];
		double d = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) {
			const double f1 = (double) a.c[k] / t1, f2 = (double) b.c[k] / t2;
			d = k ? d + (f1 < f2 ? f1 : f2) : (f1 < f2 ? f1 : f2);
		}
		d = 1 - d;
		return d < 0 ? 0 : d;
	} else if(M == CCG_MAT_L1) {                             /* matcmp.c:145 */
		int s = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) s += abs(a.c[k] - b.c[k]);
		return s;
	} else if(M == CCG_MAT_L2) {                             /* matcmp.c:160 */
		int s = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) s += (a.c[k] - b.c[k]) * (a.c[k] - b.c[k]);
		return sqrt((double) s);
	} else if(M == CCG_MAT_LINF) {                           /* matcmp.c:196 */
		int s = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) s = max(s, abs(a.c[k] - b.c[k]));
		return s;
	} else if(M == CCG_MAT_LN) {                             /* matcmp.c:175 */
		double d = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) {
			const double t = pow((double) abs(a.c[k] - b.c[k]), (double) mp.order);
			d = k ? d + t : t;
		}
		d = pow(d, 1.0 / mp.order);
		return d < 0 ? 0 : d;
	} else if(M == CCG_MAT_NL1 || M == CCG_MAT_NL2) {        /* matcmp.c:63 / :81 */
		const int t1 = a.tot - a.c[5], t2 = b.tot - b.c[5];
		double d = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) {
			const double t = (double) a.c[k] / t1 - (double) b.c[k] / t2;
			const double term = M == CCG_MAT_NL1 ? (t < 0 ? -t : t) : t * t;
			d = k ? d + term : term;
		}
		return M == CCG_MAT_NL1 ? d : sqrt(d);
	} else if(M == CCG_MAT_NLINF) {                          /* matcmp.c:125, bug-compatible: only the first component (#14) */
		const int t1 = a.tot - a.c[5], t2 = b.tot - b.c[5];
		const double t = (double) a.c[0] / t1 - (double) b.c[0] / t2;
		return t < 0 ? -t : t;
	} else {                                                 /* CCG_MAT_NLN, matcmp.c:99: the first term is not made absolute */
		const int t1 = a.tot - a.c[5], t2 = b.tot - b.c[5];
		double d = 0;
#pragma unroll
		for(int k = 0; k < 5; ++k) {
			double t = (double) a.c[k] / t1 - (double) b.c[k] / t2;
			if(k) t = t < 0 ? -t : t;
			const double term = pow(t, (double) mp.order);
			d = k ? d + term : term;
		}
		d = pow(d, 1.0 / mp.order);
		return d < 0 ? 0 : d;
	}
}

/* cos: the two square roots of coscmp depend on one sample each, so they are taken once per sample and
 * position on the host (ccg_mat_put_sample) instead of once per pair: sqrt(c1) * sqrt(c2) and the division are
 * the same IEEE operations on the same values, i.e. every per-position term is still bit-identical to the
 * reference's, at less than half of the FP64 work. */
template <int M>
__global__ void __launch_bounds__(MT * MT)
k_matdist(const uint4 *__restrict__ counts, const double *__restrict__ norms, long long lpad, const int2 *__restrict__ tiles,
          int ntiles, int nslices, int pos_per_slice, const int *__restrict__ lens, MatParams mp, double *__restrict__ part_dist,
          unsigned *__restrict__ part_rows) {
	__shared__ uint4 sA[MT][MP + 1], sB[MT][MP + 1];
	extern __shared__ double s_norm[];                       /* cos only: [2][MT][MP + 1] */
	double (*nA)[MP + 1] = reinterpret_cast<double (*)[MP + 1]>(s_norm);
	double (*nB)[MP + 1] = reinterpret_cast<double (*)[MP + 1]>(s_norm + MT * (MP + 1));
	const int tile = blockIdx.x % ntiles, ks = blockIdx.x / ntiles;
	const int ti = tiles[tile].x, tj = tiles[tile].y;
	const int li = threadIdx.x / MT, lj = threadIdx.x % MT;
	const int si = ti * MT + li, sj = tj * MT + lj;          /* sample slots of this thread's pair (i = row, j = column) */
	const int len_j = lens[sj];
	const long long p_begin = (long long) ks * pos_per_slice;
	long long p_end = p_begin + pos_per_slice;
	if(p_end > lpad) p_end = lpad;
	double dist = 0;
	unsigned rows = 0;
	for(long long p0 = p_begin; p0 < p_end; p0 += MP) {
		__syncthreads();
		for(int e = threadIdx.x; e < 2 * MT * MP; e += MT * MT) {
			const int which = e / (MT * MP), r = (e / MP) % MT, p = e % MP;
			const int slot = (which ? tj : ti) * MT + r;
			const uint4 v = counts[(size_t) slot * lpad + p0 + p];
			if(which) sB[r][p] = v; else sA[r][p] = v;
			if(M == CCG_MAT_COS) {
				const double nv = norms[(size_t) slot * lpad + p0 + p];
				if(which) nB[r][p] = nv; else nA[r][p] = nv;
			}
		}
		__syncthreads();
		if(sj < si) {
#pragma unroll 2
			for(int p = 0; p < MP; ++p) {
				if(p0 + p >= len_j) break;
				const Rec b = unpack(sB[lj][p]);
				if(mp.minDepth <= (unsigned) b.tot) {
					const Rec a = unpack(sA[li][p]);
					if(mp.minDepth <= (unsigned) a.tot) {
						double d;
						if(M == CCG_MAT_COS) {
							/* coscmp matcmp.c:420 with sqrt(c1), sqrt(c2) precomputed per sample */
							const double sa = nA[li][p], sb = nB[lj][p];
							double dot = 0;
#pragma unroll
							for(int k = 0; k < 5; ++k) dot += (double) (a.c[k] * b.c[k]);
							if(sa == 0 || sb == 0) d = -1;
							else {
								d = 1 - dot / (sa * sb);
								d = d < 0 ? 0 : d;
							}
						} else d = veccmp<M>(a, b, mp);
						if(0 <= d) { dist += d; ++rows; }
					}
				}
			}
		}
	}
	const size_t o = ((size_t) ks * ntiles + tile) * (MT * MT) + threadIdx.x;
	part_dist[o] = dist;
	part_rows[o] = rows;
}

/* adds the slices in order, applies the gates of cmpMats (matcmp.c:483-494) and writes the cells */
__global__ void __launch_bounds__(MT * MT)
k_matdist_finalize(const int2 *__restrict__ tiles, int ntiles, int nslices, const double *__restrict__ part_dist,
                   const unsigned *__restrict__ part_rows, const int *__restrict__ lens, const int *__restrict__ rank, int n,
                   unsigned norm, unsigned minLength, double minCov, int elem_size, double byteScale, void *D, void *N,
                   unsigned *__restrict__ rows_out, int row_slot) {
	const int tile = blockIdx.x;
	const int ti = tiles[tile].x, tj = tiles[tile].y;
	const int si = ti * MT + threadIdx.x / MT, sj = tj * MT + threadIdx.x % MT;
	if(si >= n || sj >= si) return;
	const int r = rank[si], c = rank[sj];
	if(r < 0 || c < 0) return;
	double dist = 0;
	unsigned rows = 0;
	for(int ks = 0; ks < nslices; ++ks) {
		const size_t o = ((size_t) ks * ntiles + tile) * (MT * MT) + threadIdx.x;
		dist += part_dist[o];
		rows += part_rows[o];
	}
	const int len_j = lens[sj], len_i = lens[si];
	/* the streamed sample must not be longer than the loaded one (matcmp.c:466-468), and the overlap gate */
	const bool ok = len_j <= len_i && !(rows < minLength || (double) rows < minCov * (double) len_j);
	double d, nn;
	if(!ok) { d = -1.0; nn = 0.0; }
	else { nn = (double) rows; d = norm ? dist / (double) rows * (double) norm : dist; }
	/* row_slot >= 0 (cmpMatRowThrd ltdmatrixthrd.c:111): only that sample's row, cell = column */
	if(row_slot >= 0 && si != row_slot) return;
	const long long cell = row_slot >= 0 ? (long long) c : (long long) r * (r - 1) / 2 + c;
	if(rows_out) rows_out[cell] = ok ? rows : 0u;
	if(elem_size == 8) {
		((double *) D)[cell] = d;
		if(N) ((double *) N)[cell] = nn;
	} else if(elem_size == 4) {
		((float *) D)[cell] = (float) d;
		if(N) ((float *) N)[cell] = (float) nn;
	} else {
		/* dtouc(value, 0.5), bytescale.h:22 */
		ccg_store_fixed(D, cell, elem_size, __dadd_rn(__dmul_rn(d, byteScale), 0.5));
		if(N) ccg_store_fixed(N, cell, elem_size, __dadd_rn(__dmul_rn(nn, byteScale), 0.5));
	}
}

typedef void (*MatKernel)(const uint4 *, const double *, long long, const int2 *, int, int, int, const int *, MatParams, double *,
                          unsigned *);

MatKernel pick_kernel(int method) {
	switch(method) {
		case CCG_MAT_COS: return k_matdist<CCG_MAT_COS>;
		case CCG_MAT_Z: return k_matdist<CCG_MAT_Z>;
		case CCG_MAT_CHI2: return k_matdist<CCG_MAT_CHI2>;
		case CCG_MAT_NCHI2: return k_matdist<CCG_MAT_NCHI2>;
		case CCG_MAT_C: return k_matdist<CCG_MAT_C>;
		case CCG_MAT_NC: return k_matdist<CCG_MAT_NC>;
		case CCG_MAT_P: return k_matdist<CCG_MAT_P>;
		case CCG_MAT_NP: return k_matdist<CCG_MAT_NP>;
		case CCG_MAT_BC: return k_matdist<CCG_MAT_BC>;
		case CCG_MAT_NBC: return k_matdist<CCG_MAT_NBC>;
		case CCG_MAT_L1: return k_matdist<CCG_MAT_L1>;
		case CCG_MAT_L2: return k_matdist<CCG_MAT_L2>;
		case CCG_MAT_LINF: return k_matdist<CCG_MAT_LINF>;
		case CCG_MAT_LN: return k_matdist<CCG_MAT_LN>;
		case CCG_MAT_NL1: return k_matdist<CCG_MAT_NL1>;
		case CCG_MAT_NL2: return k_matdist<CCG_MAT_NL2>;
		case CCG_MAT_NLINF: return k_matdist<CCG_MAT_NLINF>;
		case CCG_MAT_NLN: return k_matdist<CCG_MAT_NLN>;
		default: return 0;
	}
}

} // namespace

#define MCK(ctx, call)                                                                                       \
	do {                                                                                                     \
		cudaError_t e__ = (call);                                                                            \
		if(e__ != cudaSuccess) {                                                                             \
			snprintf(ctx->err, sizeof(ctx->err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
			return CCG_ERR_CUDA;                                                                             \
		}                                                                                                    \
	} while(0)

void ccg_mat_free(ccg_ctx *ctx) {
	cudaFree(ctx->mat_counts); ctx->mat_counts = 0;
	cudaFree(ctx->mat_norms); ctx->mat_norms = 0;
	cudaFree(ctx->mat_lens); ctx->mat_lens = 0;
	cudaFree(ctx->mat_part_dist); ctx->mat_part_dist = 0;
	cudaFree(ctx->mat_part_rows); ctx->mat_part_rows = 0;
	cudaFree(ctx->mat_rows); ctx->mat_rows = 0;
	cudaFree(ctx->mat_rank); ctx->mat_rank = 0;
	cudaFreeHost(ctx->mat_stage); ctx->mat_stage = 0;
	free(ctx->mat_hlens); ctx->mat_hlens = 0;
	ctx->mat_part_cap = 0;
	ctx->mat_n = 0;
	ctx->mat_lpad = 0;
}

extern "C" int ccg_mat_set_problem(ccg_ctx *ctx, int n, int max_len) {
	if(!ctx || n < 0 || max_len < 0) return CCG_ERR_ARG;
	MCK(ctx, cudaSetDevice(ctx->device));
	MCK(ctx, cudaStreamSynchronize(ctx->stream));
	ccg_mat_free(ctx);
	ctx->mat_n = n;
	ctx->mat_npad = ((n + MT - 1) / MT) * MT;
	if(ctx->mat_npad == 0) ctx->mat_npad = MT;
	ctx->mat_lpad = (((long long) max_len + MP - 1) / MP) * MP;
	if(ctx->mat_lpad == 0) ctx->mat_lpad = MP;
	const size_t bytes = (size_t) ctx->mat_npad * (size_t) ctx->mat_lpad * 16;
	if(cudaMalloc(&ctx->mat_counts, bytes) != cudaSuccess) {
		snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc of %zu bytes for %d count matrices of %d positions failed: %s", bytes, n,
		         max_len, cudaGetErrorString(cudaGetLastError()));
		ctx->mat_counts = 0;
		return CCG_ERR_NOMEM;
	}
	MCK(ctx, cudaMemsetAsync(ctx->mat_counts, 0, bytes, ctx->stream));
	/* sqrt of the squared count-vector length per sample and position (cos) */
	if(cudaMalloc(&ctx->mat_norms, bytes / 2) != cudaSuccess) {
		snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc of %zu bytes for the count-vector norms failed", bytes / 2);
		ctx->mat_norms = 0;
		return CCG_ERR_NOMEM;
	}
	MCK(ctx, cudaMemsetAsync(ctx->mat_norms, 0, bytes / 2, ctx->stream));
	MCK(ctx, cudaMalloc(&ctx->mat_lens, (size_t) ctx->mat_npad * sizeof(int)));
	MCK(ctx, cudaMalloc(&ctx->mat_rank, (size_t) ctx->mat_npad * sizeof(int)));
	ctx->mat_hlens = (int *) calloc((size_t) ctx->mat_npad, sizeof(int));
	if(!ctx->mat_hlens) return CCG_ERR_NOMEM;
	/* pinned staging for one sample */
	if(cudaHostAlloc(&ctx->mat_stage, (size_t) ctx->mat_lpad * 24, cudaHostAllocDefault) != cudaSuccess) {
		ctx->mat_stage = 0;
		snprintf(ctx->err, sizeof(ctx->err), "cudaHostAlloc of %lld staging bytes failed", ctx->mat_lpad * 16);
		return CCG_ERR_NOMEM;
	}
	return CCG_OK;
}

extern "C" int ccg_mat_put_sample(ccg_ctx *ctx, int idx, const uint16_t *counts6, const uint32_t *totals, int len) {
	if(!ctx || !ctx->mat_counts || idx < 0 || idx >= ctx->mat_n || len < 0 || len > ctx->mat_lpad || (len && !counts6)) return CCG_ERR_ARG;
	MCK(ctx, cudaSetDevice(ctx->device));
	/* the staging buffer is reused: wait for the previous upload */
	MCK(ctx, cudaStreamSynchronize(ctx->stream));
	uint16_t *st = (uint16_t *) ctx->mat_stage;
	double *sn = (double *) ((char *) ctx->mat_stage + (size_t) ctx->mat_lpad * 16);
	for(int p = 0; p < len; ++p) {
		const uint16_t *c = counts6 + (size_t) p * 6;
		uint16_t *o = st + (size_t) p * 8;
		unsigned tot = totals ? totals[p] : (unsigned) c[0] + c[1] + c[2] + c[3] + c[4] + c[5];
		o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = c[3]; o[4] = c[4]; o[5] = c[5];
		memcpy(o + 6, &tot, 4);
		/* coscmp's c1: int products summed in an unsigned long (matcmp.c:426-437) */
		unsigned long long c1 = 0;
		for(int k = 0; k < 5; ++k) c1 += (unsigned long long) (long long) ((int) c[k] * (int) c[k]);
		sn[p] = sqrt((double) c1);
	}
	uint4 *dst = (uint4 *) ctx->mat_counts + (size_t) idx * (size_t) ctx->mat_lpad;
	double *dstn = (double *) ctx->mat_norms + (size_t) idx * (size_t) ctx->mat_lpad;
	if(len) {
		MCK(ctx, cudaMemcpyAsync(dst, st, (size_t) len * 16, cudaMemcpyHostToDevice, ctx->stream));
		MCK(ctx, cudaMemcpyAsync(dstn, sn, (size_t) len * 8, cudaMemcpyHostToDevice, ctx->stream));
	}
	if(len < ctx->mat_lpad) {
		MCK(ctx, cudaMemsetAsync(dst + len, 0, (size_t) (ctx->mat_lpad - len) * 16, ctx->stream));
		MCK(ctx, cudaMemsetAsync(dstn + len, 0, (size_t) (ctx->mat_lpad - len) * 8, ctx->stream));
	}
	ctx->mat_hlens[idx] = len;
	return CCG_OK;
}

static int mat_run_impl(ccg_ctx *ctx, const unsigned char *include, int method, unsigned order, double alpha, unsigned norm,
                        unsigned minDepth, unsigned minLength, double minCov, int elem_size, double byteScale, void *D,
                        void *N, int *Dn_out, uint32_t *rows_inc, int row_slot) {
	if(!ctx || !ctx->mat_counts || !D) return CCG_ERR_ARG;
	if(elem_size != 8 && elem_size != 4 && elem_size != 2 && elem_size != 1) return CCG_ERR_ARG;
	MatKernel kern = pick_kernel(method);
	if(!kern) return CCG_ERR_ARG;
	if((method == CCG_MAT_LN || method == CCG_MAT_NLN) && order == 0) return CCG_ERR_ARG;
	MCK(ctx, cudaSetDevice(ctx->device));
	const int n = ctx->mat_n, npad = ctx->mat_npad;
	std::vector<int> rank((size_t) npad, -1);
	int Dn = 0;
	for(int i = 0; i < n; ++i)
		if(!include || include[i]) rank[(size_t) i] = Dn++;
	if(Dn_out) *Dn_out = Dn;
	if(Dn < 2) return CCG_OK;
	/* tiles of the lower triangle that hold at least one included pair */
	std::vector<int2> tiles;
	long long ordinal = 0;
	const int T = npad / MT;
	for(int ti = 0; ti < T; ++ti)
		for(int tj = 0; tj <= ti; ++tj) {
			if(row_slot >= 0 && ti != row_slot / MT) continue;
			bool any_i = false, any_j = false;
			for(int k = 0; k < MT; ++k) {
				any_i |= rank[(size_t) ti * MT + k] >= 0;
				any_j |= rank[(size_t) tj * MT + k] >= 0;
			}
			/* one process per GPU: the 16 x 16 tiles are dealt round-robin (ccg_set_partition); cells of other
			 * ranks read back as zero */
			if(any_i && any_j && (ctx->world <= 1 || (int) (ordinal++ % ctx->world) == ctx->rank)) tiles.push_back(make_int2(ti, tj));
			else if(any_i && any_j) { /* another rank's tile */ }
		}
	const int ntiles = (int) tiles.size();
	if(ntiles == 0) {
		const size_t cells0 = row_slot >= 0 ? (size_t) Dn - 1 : (size_t) Dn * (Dn - 1) / 2;
		memset(D, 0, cells0 * elem_size);
		if(N) memset(N, 0, cells0 * elem_size);
		if(rows_inc) memset(rows_inc, 0, cells0 * 4);
		return CCG_OK;
	}
	/* slices of the position axis: fill the machine a few times over, at least 1024 positions each */
	const long long stages = ctx->mat_lpad / MP;
	long long want = (8LL * ctx->sm_count * 4 + ntiles - 1) / ntiles;
	if(want < 1) want = 1;
	long long stages_per = (stages + want - 1) / want;
	if(stages_per < 8) stages_per = 8;
	if(stages_per > stages) stages_per = stages;
	const int nslices = (int) ((stages + stages_per - 1) / stages_per);
	const int pos_per_slice = (int) (stages_per * MP);

	const size_t part = (size_t) nslices * ntiles * MT * MT;
	if(ctx->mat_part_cap < part) {
		MCK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->mat_part_dist);
		cudaFree(ctx->mat_part_rows);
		ctx->mat_part_dist = 0;
		ctx->mat_part_rows = 0;
		ctx->mat_part_cap = 0;
		MCK(ctx, cudaMalloc(&ctx->mat_part_dist, part * sizeof(double)));
		MCK(ctx, cudaMalloc(&ctx->mat_part_rows, part * sizeof(unsigned)));
		ctx->mat_part_cap = part;
	}
	int rc = CCG_OK;
	const size_t cells = row_slot >= 0 ? (size_t) Dn - 1 : (size_t) Dn * (Dn - 1) / 2;
	int2 *d_tiles = 0;
	void *d_D = 0, *d_N = 0;
	unsigned *d_rows = 0;
	cudaError_t e = cudaMalloc(&d_tiles, (size_t) ntiles * sizeof(int2));
	if(e == cudaSuccess) e = cudaMalloc(&d_D, cells * 8);
	if(e == cudaSuccess && N) e = cudaMalloc(&d_N, cells * 8);
	if(e == cudaSuccess && rows_inc) e = cudaMalloc(&d_rows, cells * 4);
	if(e == cudaSuccess && ctx->world > 1) {
		e = cudaMemsetAsync(d_D, 0, cells * 8, ctx->stream);
		if(e == cudaSuccess && d_N) e = cudaMemsetAsync(d_N, 0, cells * 8, ctx->stream);
		if(e == cudaSuccess && d_rows) e = cudaMemsetAsync(d_rows, 0, cells * 4, ctx->stream);
	}
	if(e == cudaSuccess) e = cudaMemcpyAsync(d_tiles, tiles.data(), (size_t) ntiles * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream);
	if(e == cudaSuccess) e = cudaMemcpyAsync(ctx->mat_rank, rank.data(), (size_t) npad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
	if(e == cudaSuccess) e = cudaMemcpyAsync(ctx->mat_lens, ctx->mat_hlens, (size_t) npad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
	if(e == cudaSuccess) {
		MatParams mp;
		mp.minDepth = minDepth;
		mp.order = order;
		mp.alpha = alpha;
		cudaEventRecord(ctx->ev0, ctx->stream);
		const size_t dyn = method == CCG_MAT_COS ? (size_t) 2 * MT * (MP + 1) * sizeof(double) : 0;
		if(dyn) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn);
		kern<<<(unsigned) ((long long) ntiles * nslices), MT * MT, dyn, ctx->stream>>>((const uint4 *) ctx->mat_counts,
		                                                                              (const double *) ctx->mat_norms, ctx->mat_lpad, d_tiles,
		                                                                              ntiles, nslices, pos_per_slice, ctx->mat_lens, mp,
		                                                                              ctx->mat_part_dist, ctx->mat_part_rows);
		cudaEventRecord(ctx->ev1, ctx->stream);
		ctx->ev_valid = 1;
		ctx->launches++;
		k_matdist_finalize<<<ntiles, MT * MT, 0, ctx->stream>>>(d_tiles, ntiles, nslices, ctx->mat_part_dist, ctx->mat_part_rows,
		                                                       ctx->mat_lens, ctx->mat_rank, n, norm, minLength, minCov, elem_size,
		                                                       byteScale, d_D, d_N, d_rows, row_slot);
		ctx->launches++;
		e = cudaGetLastError();
	}
	if(e == cudaSuccess) e = cudaMemcpyAsync(D, d_D, cells * elem_size, cudaMemcpyDeviceToHost, ctx->stream);
	if(e == cudaSuccess && N) e = cudaMemcpyAsync(N, d_N, cells * elem_size, cudaMemcpyDeviceToHost, ctx->stream);
	if(e == cudaSuccess && rows_inc) e = cudaMemcpyAsync(rows_inc, d_rows, cells * 4, cudaMemcpyDeviceToHost, ctx->stream);
	if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if(e != cudaSuccess) {
		snprintf(ctx->err, sizeof(ctx->err), "count-matrix distance run failed: %s", cudaGetErrorString(e));
		rc = CCG_ERR_CUDA;
	}
	snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k_matdist<%d> tiles=%d slices=%d", method, ntiles, nslices);
	cudaFree(d_tiles);
	cudaFree(d_D);
	cudaFree(d_N);
	cudaFree(d_rows);
	return rc;
}

extern "C" int ccg_mat_run(ccg_ctx *ctx, const unsigned char *include, int method, unsigned order, double alpha, unsigned norm,
                           unsigned minDepth, unsigned minLength, double minCov, int elem_size, double byteScale, void *D,
                           void *N, int *Dn_out, uint32_t *rows_inc) {
	return mat_run_impl(ctx, include, method, order, alpha, norm, minDepth, minLength, minCov, elem_size, byteScale, D, N, Dn_out,
	                    rows_inc, -1);
}

/* cmpMatRowThrd (ltdmatrixthrd.c:111-181): the last uploaded sample against all the others */
extern "C" int ccg_mat_run_row(ccg_ctx *ctx, int row_slot, int method, unsigned order, double alpha, unsigned norm,
                               unsigned minDepth, unsigned minLength, double minCov, double *D, double *N, uint32_t *rows_inc) {
	if(!ctx || !ctx->mat_counts || row_slot < 0 || row_slot >= ctx->mat_n) return CCG_ERR_ARG;
	if(ctx->world > 1) return CCG_ERR_UNSUPPORTED;
	/* columns are the slots below the row: everything above it stays out */
	std::vector<unsigned char> use((size_t) ctx->mat_n, 0);
	for(int i = 0; i <= row_slot; ++i) use[(size_t) i] = 1;
	int Dn = 0;
	return mat_run_impl(ctx, use.data(), method, order, alpha, norm, minDepth, minLength, minCov, 8, 1.0, D, N, &Dn, rows_inc,
	                    row_slot);
}
