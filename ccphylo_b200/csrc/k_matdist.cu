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
 * Device layout (12 bytes per position and sample -- the payload of the reference's record, matparse.c:254-259;
 * 2,000 samples x 5 Mbp = 120 GB fit one B200):
 *     counts[slot][plane][position] u32,  plane 0 = A | C << 16,  1 = G | T << 16,  2 = '-' | N << 16
 * positions zero-padded to a multiple of 32.  The row total (the reference's 32-bit field) and, for `cos`, the
 * double sqrt(sum of squared counts) are recomputed when a CTA stages a block of positions in shared memory: once
 * per sample and position of the tile instead of once per pair.  Insertion rows (reference base '-') are dropped by
 * the host parser.
 *
 * Per pair (i > j), over the positions p < len_j of the earlier sample (the one the reference
 * streams):   if(minDepth <= tot_j[p]) { ++nNucs;
 *                 if(minDepth <= tot_i[p] && 0 <= (d = veccmp(c_i[p], c_j[p], tot_i, tot_j))) { dist += d; ++rowsInc; } }
 * Gate: rowsInc < minLength || rowsInc < minCov * len_j  ->  D = -1, N = 0 ("No sufficient
 * overlap"); else D = norm ? dist / rowsInc * norm : dist, N = rowsInc.
 *
 * Work item = (32 x 32 tile of sample pairs, slice of positions).  A CTA of 256 threads stages 32 positions of
 * its 32 + 32 samples in shared memory (coalesced loads along the position axis); every thread owns a 2 x 2 block
 * of pairs (rows li, li + 16; columns lj, lj + 16), so every record it reads from shared memory serves two pairs,
 * and walks the positions in order; the per-slice partial sums go to a [slice][tile][1024] buffer and are added in
 * slice order by k_matdist_finalize, so the result is deterministic (it differs from the reference's strictly
 * sequential fp64 sum only by rounding: the parity bar for this path is 1e-6 relative, counts are exact).
 * Roofline: CUDA-core FP64 (one divide per position pair for `cos`) + INT; HBM traffic is 2 x 12 B x 32 samples
 * per 1024 position pairs, i.e. negligible.  -fmad=false (csrc/Makefile) keeps nvcc from contracting the
 * reference's multiply-then-add sequences.
 *
 * Multi-GPU: the POSITION axis is cut (the same K split as the FASTA path): ccg_mat_run_partial returns a member's
 * raw per-pair sums over its positions, ccg_mat_finalize_host adds nothing and gates / scales on the host -- the
 * reference's own double arithmetic; a multi-GPU context (ccg_init_multi) does both behind ccg_mat_run.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ccg_internal.h"
#include "epilogue.cuh"

namespace {

constexpr int MT = 32;            /* tile edge in samples */
constexpr int TH = 16;            /* threads per tile edge: a thread owns rows li, li + TH and columns lj, lj + TH */
constexpr int MP = 32;            /* positions per stage (2 x 32 x 33 x 16 B = 33 KB of static shared memory) */

struct MatParams {
	unsigned minDepth;
	unsigned order;               /* l<n> / nl<n> */
	double alpha;                 /* z */
};

struct Rec {
	int c[6];                     /* A C G T - N */
	int tot;
};

/* shared-memory record: the three packed count words + the row total computed at staging time */
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

/* one position of one pair: the gates of cmpMats (matcmp.c:470-481) around the per-position distance */
template <int M>
__device__ __forceinline__ void pair_step(const Rec &a, const Rec &b, double sa, double sb, const MatParams &mp, double &dist, unsigned &rows) {
	if(mp.minDepth <= (unsigned) b.tot && mp.minDepth <= (unsigned) a.tot) {
		double d;
		if(M == CCG_MAT_COS) {
			/* coscmp matcmp.c:420.  sqrt(c1), sqrt(c2) depend on one sample each: taken once per sample and position
			 * at staging time; the dot product is a sum of int products converted to double one by one in the reference
			 * -- every partial sum an integer below 2^53, so the int64 sum converted once is the same double */
			long long dot = 0;
#pragma unroll
			for(int k = 0; k < 5; ++k) dot += (long long) (a.c[k] * b.c[k]);
			if(sa == 0 || sb == 0) d = -1;
			else {
				d = 1 - (double) dot / (sa * sb);
				d = d < 0 ? 0 : d;
			}
		} else d = veccmp<M>(a, b, mp);
		if(0 <= d) { dist += d; ++rows; }
	}
}

template <int M>
__global__ void __launch_bounds__(TH * TH)
k_matdist(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ tot_over, long long lpad, const int2 *__restrict__ tiles, int ntiles, int nslices,
          int pos_per_slice, const int *__restrict__ lens, MatParams mp, double *__restrict__ part_dist, unsigned *__restrict__ part_rows) {
	__shared__ uint4 sA[MT][MP + 1], sB[MT][MP + 1];
	extern __shared__ double s_norm[];                       /* cos only: [2][MT][MP + 1] */
	double (*nA)[MP + 1] = reinterpret_cast<double (*)[MP + 1]>(s_norm);
	double (*nB)[MP + 1] = reinterpret_cast<double (*)[MP + 1]>(s_norm + MT * (MP + 1));
	const int tile = blockIdx.x % ntiles, ks = blockIdx.x / ntiles;
	const int ti = tiles[tile].x, tj = tiles[tile].y;
	const int li = threadIdx.x / TH, lj = threadIdx.x % TH;
	/* the 2 x 2 pairs of this thread: q = 2 * (row half) + (column half) */
	const int si0 = ti * MT + li, si1 = si0 + TH, sj0 = tj * MT + lj, sj1 = sj0 + TH;
	const int len_j0 = lens[sj0], len_j1 = lens[sj1];
	const bool on[4] = {sj0 < si0, sj1 < si0, sj0 < si1, sj1 < si1};
	const long long p_begin = (long long) ks * pos_per_slice;
	long long p_end = p_begin + pos_per_slice;
	if(p_end > lpad) p_end = lpad;
	double dist[4] = {0, 0, 0, 0};
	unsigned rows[4] = {0, 0, 0, 0};
	for(long long p0 = p_begin; p0 < p_end; p0 += MP) {
		__syncthreads();
		for(int e = threadIdx.x; e < 2 * MT * MP; e += TH * TH) {
			const int which = e / (MT * MP), r = (e / MP) % MT, p = e % MP;
			const int slot = (which ? tj : ti) * MT + r;
			const uint32_t *src = counts + (size_t) slot * 3 * lpad + p0 + p;
			uint4 v;
			v.x = src[0];
			v.y = src[lpad];
			v.z = src[2 * lpad];
			v.w = (v.x & 0xFFFF) + (v.x >> 16) + (v.y & 0xFFFF) + (v.y >> 16) + (v.z & 0xFFFF) + (v.z >> 16);   /* the row total */
			if(tot_over) {
				/* a sample with a count above 65,535: the reference keeps 16 bits of the count but the whole depth in
				 * the row total (matparse.c:246-258); such samples carry their totals in a side plane */
				const uint32_t t = tot_over[(size_t) slot * lpad + p0 + p];
				if(t != 0xFFFFFFFFu) v.w = t;
			}
			if(which) sB[r][p] = v; else sA[r][p] = v;
			if(M == CCG_MAT_COS) {
				/* coscmp's c1: int products summed in an unsigned long (matcmp.c:426-437) */
				const Rec c = unpack(v);
				unsigned long long c1 = 0;
#pragma unroll
				for(int k = 0; k < 5; ++k) c1 += (unsigned long long) (long long) (c.c[k] * c.c[k]);
				const double nv = sqrt((double) c1);
				if(which) nB[r][p] = nv; else nA[r][p] = nv;
			}
		}
		__syncthreads();
		if(on[2]) {                                          /* the lower-left pair of the block is the last one to drop out */
#pragma unroll 2
			for(int p = 0; p < MP; ++p) {
				const long long pos = p0 + p;
				const bool in0 = pos < len_j0, in1 = pos < len_j1;
				if(!in0 && !in1) break;
				const Rec a0 = unpack(sA[li][p]), a1 = unpack(sA[li + TH][p]);
				const Rec b0 = unpack(sB[lj][p]), b1 = unpack(sB[lj + TH][p]);
				double na0 = 0, na1 = 0, nb0 = 0, nb1 = 0;
				if(M == CCG_MAT_COS) { na0 = nA[li][p]; na1 = nA[li + TH][p]; nb0 = nB[lj][p]; nb1 = nB[lj + TH][p]; }
				if(on[0] && in0) pair_step<M>(a0, b0, na0, nb0, mp, dist[0], rows[0]);
				if(on[1] && in1) pair_step<M>(a0, b1, na0, nb1, mp, dist[1], rows[1]);
				if(in0) pair_step<M>(a1, b0, na1, nb0, mp, dist[2], rows[2]);
				if(on[3] && in1) pair_step<M>(a1, b1, na1, nb1, mp, dist[3], rows[3]);
			}
		}
	}
	const size_t o = (((size_t) ks * ntiles + tile) * (TH * TH) + threadIdx.x) * 4;
#pragma unroll
	for(int q = 0; q < 4; ++q) {
		part_dist[o + q] = dist[q];
		part_rows[o + q] = rows[q];
	}
}

/* ------------------------------------------------------------------------------------------
 * `cos` (the default -d, coscmp matcmp.c:420-446) has its own kernel: everything that depends on ONE sample and position
 * is done once when a CTA stages the position -- the five counts unpacked to int, the depth gate and the "norm is zero"
 * test folded into one flag, sqrt(c1) and its reciprocal -- so a position pair costs five integer multiply-adds, one
 * int -> double conversion and a division that is done as the final Newton step of a division:
 *     den = sa * sb  (as the reference rounds it);  y = (1/sa) * (1/sb);  q0 = dot * y;  r = fma(-den, q0, dot);
 *     q = fma(r, y, q0)
 * q is the quotient dot / den rounded from an error far below half an ulp (Markstein's correction step with a
 * reciprocal good to ~2 ulp): it can differ from the correctly rounded quotient only when the exact value lies within
 * ~2^-50 ulp of a rounding boundary, and the parity bar of this path is 1e-6 relative.  The int32 dot product is exact
 * while no count of the stage exceeds 20,724 (5 * 20724^2 < 2^31); a stage with a larger count (the reference's int
 * products wrap from 46,341 on, matcmp.c:429-437) takes the literal path: wrapped products summed in 64 bits and a real
 * division.  Zero-padded positions have c1 = 0 and never count, so no length test is needed in the loop.
 * ------------------------------------------------------------------------------------------ */
constexpr int MPC = 16;           /* positions per stage of the cos kernel */
constexpr int COS_FAST_MAX = 20724;

struct CosStage {
	int4 c03[2][MT][MPC + 1];     /* A C G T */
	int c4[2][MT][MPC + 1];       /* '-' */
	double2 nr[2][MT][MPC + 1];   /* sqrt(c1), 1 / sqrt(c1); NaN, NaN when the position does not count for this sample */
	unsigned ok[2][MT];           /* bit p: position p of the stage counts (minDepth <= total && c1 != 0) */
};

__global__ void __launch_bounds__(TH * TH, 3)
k_matdist_cos(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ tot_over, long long lpad, const int2 *__restrict__ tiles,
              int ntiles, int nslices, int pos_per_slice, MatParams mp, double *__restrict__ part_dist, unsigned *__restrict__ part_rows) {
	extern __shared__ __align__(16) unsigned char cos_smem[];
	CosStage &st = *reinterpret_cast<CosStage *>(cos_smem);
	const int tile = blockIdx.x % ntiles, ks = blockIdx.x / ntiles;
	const int ti = tiles[tile].x, tj = tiles[tile].y;
	const int li = threadIdx.x / TH, lj = threadIdx.x % TH;
	const bool any_on = tj * MT + lj < ti * MT + li + TH;            /* sj0 < si1: the last pair of the block to drop out */
	const long long p_begin = (long long) ks * pos_per_slice;
	long long p_end = p_begin + pos_per_slice;
	if(p_end > lpad) p_end = lpad;
	double dist[4] = {0, 0, 0, 0};
	unsigned rows[4] = {0, 0, 0, 0};
	for(long long p0 = p_begin; p0 < p_end; p0 += MPC) {
		__syncthreads();
		int big = 0;
		for(int e = threadIdx.x; e < 2 * MT * MPC; e += TH * TH) {
			const int which = e / (MT * MPC), r = (e / MPC) % MT, p = e % MPC;
			const int slot = (which ? tj : ti) * MT + r;
			const uint32_t *src = counts + (size_t) slot * 3 * lpad + p0 + p;
			const uint32_t x = src[0], y = src[lpad], z = src[2 * lpad];
			const int c0 = x & 0xFFFF, c1 = x >> 16, c2 = y & 0xFFFF, c3 = y >> 16, c4 = z & 0xFFFF;
			uint32_t tot = (uint32_t) (c0 + c1 + c2 + c3 + c4) + (z >> 16);
			if(tot_over) {
				const uint32_t t = tot_over[(size_t) slot * lpad + p0 + p];
				if(t != 0xFFFFFFFFu) tot = t;
			}
			/* coscmp's c1: int products summed in an unsigned long (matcmp.c:426-437) */
			const unsigned long long sq = (unsigned long long) (long long) (c0 * c0) + (unsigned long long) (long long) (c1 * c1) +
			                              (unsigned long long) (long long) (c2 * c2) + (unsigned long long) (long long) (c3 * c3) +
			                              (unsigned long long) (long long) (c4 * c4);
			const double nv = sqrt((double) sq);
			big |= (c0 | c1 | c2 | c3 | c4) > COS_FAST_MAX;          /* all five are below 2^16: the OR bounds the largest */
			/* A position that does not count for this sample gets NaN norms: every pair it is part of then computes a NaN
			 * distance, which fails the `d > 0` test below and adds nothing -- no flag in the inner loop.  The counted
			 * positions (rowsInc) are the popcount of the AND of the two samples' bit masks, once per stage. */
			const bool ok = mp.minDepth <= tot && sq != 0;
			const double qnan = __longlong_as_double(0x7FF8000000000000LL);
			st.c03[which][r][p] = make_int4(c0, c1, c2, c3);
			st.c4[which][r][p] = c4;
			/* the reciprocal only seeds the correction step of the division: rsqrt's 1 ulp is plenty */
			st.nr[which][r][p] = ok ? make_double2(nv, rsqrt((double) sq)) : make_double2(qnan, qnan);
			/* MPC = 16 consecutive lanes hold the 16 positions of one sample */
			const unsigned bits = __ballot_sync(0xffffffffu, ok);
			if(p == 0) st.ok[which][r] = (bits >> (threadIdx.x & 16)) & 0xFFFFu;
		}
		/* (the OR of the counts can exceed the limit while every count is below it: that only sends a stage down the
		 * literal path more often than needed) */
		big = __syncthreads_or(big);
		if(!any_on) continue;
		{
			const unsigned ka0 = st.ok[0][li], ka1 = st.ok[0][li + TH], kb0 = st.ok[1][lj], kb1 = st.ok[1][lj + TH];
			rows[0] += (unsigned) __popc(ka0 & kb0);
			rows[1] += (unsigned) __popc(ka0 & kb1);
			rows[2] += (unsigned) __popc(ka1 & kb0);
			rows[3] += (unsigned) __popc(ka1 & kb1);
		}
		if(!big) {
#pragma unroll 4
			for(int p = 0; p < MPC; ++p) {
				const int4 a0 = st.c03[0][li][p], a1 = st.c03[0][li + TH][p], b0 = st.c03[1][lj][p], b1 = st.c03[1][lj + TH][p];
				const int a0x = st.c4[0][li][p], a1x = st.c4[0][li + TH][p], b0x = st.c4[1][lj][p], b1x = st.c4[1][lj + TH][p];
				const double2 na0 = st.nr[0][li][p], na1 = st.nr[0][li + TH][p], nb0 = st.nr[1][lj][p], nb1 = st.nr[1][lj + TH][p];
#define CCG_COS_PAIR(q, A, AX, NA, B, BX, NB)                                                            \
	{                                                                                                    \
		const int dot = A.x * B.x + A.y * B.y + A.z * B.z + A.w * B.w + AX * BX;                         \
		const double dd = (double) dot, den = NA.x * NB.x, yy = NA.y * NB.y;                             \
		const double q0 = dd * yy;                                                                       \
		const double qq = fma(fma(-den, q0, dd), yy, q0);                                                \
		const double d = 1.0 - qq;                                                                       \
		if(d > 0) dist[q] += d;                          /* the clamp; false for the NaN of a gated position */ \
	}
				CCG_COS_PAIR(0, a0, a0x, na0, b0, b0x, nb0)
				CCG_COS_PAIR(1, a0, a0x, na0, b1, b1x, nb1)
				CCG_COS_PAIR(2, a1, a1x, na1, b0, b0x, nb0)
				CCG_COS_PAIR(3, a1, a1x, na1, b1, b1x, nb1)
#undef CCG_COS_PAIR
			}
		} else {
#pragma unroll 1
			for(int p = 0; p < MPC; ++p) {
#pragma unroll
				for(int q = 0; q < 4; ++q) {
					const int ra = li + TH * (q >> 1), rb = lj + TH * (q & 1);
					if(!((st.ok[0][ra] & st.ok[1][rb]) >> p & 1u)) continue;
					const int4 a = st.c03[0][ra][p], b = st.c03[1][rb][p];
					const long long dot = (long long) (a.x * b.x) + (long long) (a.y * b.y) + (long long) (a.z * b.z) +
					                      (long long) (a.w * b.w) + (long long) (st.c4[0][ra][p] * st.c4[1][rb][p]);
					double d = 1 - (double) dot / (st.nr[0][ra][p].x * st.nr[1][rb][p].x);
					d = d < 0 ? 0 : d;
					dist[q] += d;
				}
			}
		}
	}
	const size_t o = (((size_t) ks * ntiles + tile) * (TH * TH) + threadIdx.x) * 4;
#pragma unroll
	for(int q = 0; q < 4; ++q) {
		part_dist[o + q] = dist[q];
		part_rows[o + q] = rows[q];
	}
}

/* the gates of cmpMats (matcmp.c:483-494) and the cell formats; shared by the device epilogue and the host one */
__host__ __device__ __forceinline__ void mat_cell(double dist, unsigned rows, int len_i, int len_j, unsigned norm, unsigned minLength,
                                                  double minCov, double *d_out, double *n_out, bool *ok_out) {
	/* the streamed sample must not be longer than the loaded one (matcmp.c:466-468), and the overlap gate */
	const bool ok = len_j <= len_i && !(rows < minLength || (double) rows < minCov * (double) len_j);
	if(!ok) { *d_out = -1.0; *n_out = 0.0; }
	else { *n_out = (double) rows; *d_out = norm ? dist / (double) rows * (double) norm : dist; }
	*ok_out = ok;
}

/* adds the slices in order, applies the gates of cmpMats (matcmp.c:483-494) and writes the cells; with raw != 0 the
 * sums themselves are written (ccg_mat_run_partial: a member of a position split) */
__global__ void __launch_bounds__(TH * TH)
k_matdist_finalize(const int2 *__restrict__ tiles, int ntiles, int nslices, const double *__restrict__ part_dist,
                   const unsigned *__restrict__ part_rows, const int *__restrict__ lens, const int *__restrict__ rank, int n,
                   unsigned norm, unsigned minLength, double minCov, int elem_size, double byteScale, void *D, void *N,
                   unsigned *__restrict__ rows_out, int row_slot, int raw) {
	const int tile = blockIdx.x;
	const int ti = tiles[tile].x, tj = tiles[tile].y;
#pragma unroll 1
	for(int q = 0; q < 4; ++q) {
		const int si = ti * MT + threadIdx.x / TH + TH * (q >> 1), sj = tj * MT + threadIdx.x % TH + TH * (q & 1);
		if(si >= n || sj >= si) continue;
		const int r = rank[si], c = rank[sj];
		if(r < 0 || c < 0) continue;
		double dist = 0;
		unsigned rows = 0;
		for(int ks = 0; ks < nslices; ++ks) {
			const size_t o = (((size_t) ks * ntiles + tile) * (TH * TH) + threadIdx.x) * 4 + q;
			dist += part_dist[o];
			rows += part_rows[o];
		}
		/* row_slot >= 0 (cmpMatRowThrd ltdmatrixthrd.c:111): only that sample's row, cell = column */
		if(row_slot >= 0 && si != row_slot) continue;
		const long long cell = row_slot >= 0 ? (long long) c : (long long) r * (r - 1) / 2 + c;
		if(raw) {
			((double *) D)[cell] = dist;
			rows_out[cell] = rows;
			continue;
		}
		double d, nn;
		bool ok;
		mat_cell(dist, rows, lens[si], lens[sj], norm, minLength, minCov, &d, &nn, &ok);
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
}

/* upload format -> store: a position's six u16 counts are three u32 words (A|C<<16, G|T<<16, -|N<<16 on a little-endian
 * host): de-interleave them into the three planes of the slot */
__global__ void __launch_bounds__(256)
k_mat_planes(const uint32_t *__restrict__ rows, int len, long long lpad, uint32_t *__restrict__ dst) {
	for(long long p = (long long) blockIdx.x * blockDim.x + threadIdx.x; p < lpad; p += (long long) gridDim.x * blockDim.x) {
		uint32_t w0 = 0, w1 = 0, w2 = 0;
		if(p < len) { w0 = rows[3 * p]; w1 = rows[3 * p + 1]; w2 = rows[3 * p + 2]; }
		dst[p] = w0;
		dst[lpad + p] = w1;
		dst[2 * lpad + p] = w2;
	}
}

typedef void (*MatKernel)(const uint32_t *, const uint32_t *, long long, const int2 *, int, int, int, const int *, MatParams, double *, unsigned *);

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
	cudaFree(ctx->mat_tot_over); ctx->mat_tot_over = 0;
	cudaFree(ctx->mat_lens); ctx->mat_lens = 0;
	cudaFree(ctx->mat_part_dist); ctx->mat_part_dist = 0;
	cudaFree(ctx->mat_part_rows); ctx->mat_part_rows = 0;
	cudaFree(ctx->mat_rows); ctx->mat_rows = 0;
	cudaFree(ctx->mat_out_D); ctx->mat_out_D = 0;
	cudaFree(ctx->mat_out_N); ctx->mat_out_N = 0;
	cudaFree(ctx->mat_tiles); ctx->mat_tiles = 0;
	ctx->mat_cells_cap = 0;
	ctx->mat_tiles_cap = 0;
	cudaFree(ctx->mat_rank); ctx->mat_rank = 0;
	cudaFree(ctx->mat_stage); ctx->mat_stage = 0;
	free(ctx->mat_hlens); ctx->mat_hlens = 0;
	ctx->mat_part_cap = 0;
	ctx->mat_n = 0;
	ctx->mat_lpad = 0;
}

extern "C" int ccg_mat_set_problem(ccg_ctx *ctx, int n, int max_len) {
	if(!ctx || n < 0 || max_len < 0) return CCG_ERR_ARG;
	if(ctx->multi) return ccg_multi_mat_set_problem(ctx, n, max_len);
	MCK(ctx, cudaSetDevice(ctx->device));
	MCK(ctx, cudaStreamSynchronize(ctx->stream));
	ccg_mat_free(ctx);
	ctx->mat_n = n;
	ctx->mat_npad = ((n + MT - 1) / MT) * MT;
	if(ctx->mat_npad == 0) ctx->mat_npad = MT;
	ctx->mat_lpad = (((long long) max_len + MP - 1) / MP) * MP;
	if(ctx->mat_lpad == 0) ctx->mat_lpad = MP;
	const size_t bytes = (size_t) ctx->mat_npad * (size_t) ctx->mat_lpad * 12;
	if(cudaMalloc(&ctx->mat_counts, bytes) != cudaSuccess) {
		snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc of %zu bytes for %d count matrices of %d positions failed: %s", bytes, n,
		         max_len, cudaGetErrorString(cudaGetLastError()));
		ctx->mat_counts = 0;
		return CCG_ERR_NOMEM;
	}
	MCK(ctx, cudaMemsetAsync(ctx->mat_counts, 0, bytes, ctx->stream));
	MCK(ctx, cudaMalloc(&ctx->mat_lens, (size_t) ctx->mat_npad * sizeof(int)));
	MCK(ctx, cudaMalloc(&ctx->mat_rank, (size_t) ctx->mat_npad * sizeof(int)));
	ctx->mat_hlens = (int *) calloc((size_t) ctx->mat_npad, sizeof(int));
	if(!ctx->mat_hlens) return CCG_ERR_NOMEM;
	/* device staging for one sample in the upload format */
	if(cudaMalloc(&ctx->mat_stage, (size_t) ctx->mat_lpad * 12) != cudaSuccess) {
		ctx->mat_stage = 0;
		snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc of %lld staging bytes failed", ctx->mat_lpad * 12);
		return CCG_ERR_NOMEM;
	}
	return CCG_OK;
}

extern "C" int ccg_mat_put_sample(ccg_ctx *ctx, int idx, const uint16_t *counts6, const uint32_t *totals, int len) {
	if(ctx && ctx->multi) return ccg_multi_mat_put_sample(ctx, idx, counts6, totals, len);
	if(!ctx || !ctx->mat_counts || idx < 0 || idx >= ctx->mat_n || len < 0 || len > ctx->mat_lpad || (len && !counts6)) return CCG_ERR_ARG;
	MCK(ctx, cudaSetDevice(ctx->device));
	/* the (device) staging buffer is reused in stream order; a pageable source is staged before the copy call returns,
	 * a pinned one must stay untouched until ccg_sync (include/ccphylo_gpu.h) */
	/* the reference's 32-bit row total (matparse.c:254-259) is the sum of the six counts unless a count was truncated */
	bool over = false;
	if(totals)
		for(int p = 0; p < len && !over; ++p) {
			const uint16_t *c = counts6 + (size_t) p * 6;
			over = totals[p] != (unsigned) c[0] + c[1] + c[2] + c[3] + c[4] + c[5];
		}
	const long long lp = ctx->mat_lpad;
	uint32_t *dst = (uint32_t *) ctx->mat_counts + (size_t) idx * 3 * (size_t) lp;
	if(len) MCK(ctx, cudaMemcpyAsync(ctx->mat_stage, counts6, (size_t) len * 12, cudaMemcpyHostToDevice, ctx->stream));
	{
		long long blocks = (lp + 255) / 256;
		if(blocks > 8LL * ctx->sm_count) blocks = 8LL * ctx->sm_count;
		k_mat_planes<<<(unsigned) blocks, 256, 0, ctx->stream>>>((const uint32_t *) ctx->mat_stage, len, lp, dst);
		ctx->launches++;
		MCK(ctx, cudaGetLastError());
	}
	if(over || ctx->mat_tot_over) {
		/* row totals that are not the sum of the stored (16-bit) counts: such a sample's totals go to the side plane,
		 * 0xFFFFFFFF everywhere else means "the sum" */
		if(!ctx->mat_tot_over) {
			const size_t tb = (size_t) ctx->mat_npad * (size_t) lp * 4;
			if(cudaMalloc(&ctx->mat_tot_over, tb) != cudaSuccess) {
				ctx->mat_tot_over = 0;
				snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc of %zu bytes for the row totals failed", tb);
				return CCG_ERR_NOMEM;
			}
			MCK(ctx, cudaMemsetAsync(ctx->mat_tot_over, 0xFF, tb, ctx->stream));
		}
		MCK(ctx, cudaMemsetAsync((uint32_t *) ctx->mat_tot_over + (size_t) idx * lp, 0xFF, (size_t) lp * 4, ctx->stream));
		if(over) MCK(ctx, cudaMemcpyAsync((uint32_t *) ctx->mat_tot_over + (size_t) idx * lp, totals, (size_t) len * 4, cudaMemcpyHostToDevice, ctx->stream));
	}
	ctx->mat_hlens[idx] = len;
	return CCG_OK;
}

static int mat_run_impl(ccg_ctx *ctx, const unsigned char *include, int method, unsigned order, double alpha, unsigned norm,
                        unsigned minDepth, unsigned minLength, double minCov, int elem_size, double byteScale, void *D,
                        void *N, int *Dn_out, uint32_t *rows_inc, int row_slot, int raw = 0) {
	if(!ctx || !ctx->mat_counts || !D) return CCG_ERR_ARG;
	if(raw && (!rows_inc || elem_size != 8 || N)) return CCG_ERR_ARG;
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
			/* one process per GPU: the 32 x 32 tiles are dealt round-robin (ccg_set_partition); cells of other
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
	/* tile list and device results: kept between runs and only grown (cudaMalloc / cudaFree per run cost up to
	 * hundreds of milliseconds on some hosts) */
	cudaError_t e = cudaSuccess;
	if(ctx->mat_tiles_cap < (size_t) ntiles || ctx->mat_cells_cap < cells) {
		e = cudaStreamSynchronize(ctx->stream);
		if(e == cudaSuccess && ctx->mat_tiles_cap < (size_t) ntiles) {
			cudaFree(ctx->mat_tiles);
			ctx->mat_tiles = 0;
			ctx->mat_tiles_cap = 0;
			e = cudaMalloc(&ctx->mat_tiles, (size_t) ntiles * sizeof(int2));
			if(e == cudaSuccess) ctx->mat_tiles_cap = (size_t) ntiles;
		}
		if(e == cudaSuccess && ctx->mat_cells_cap < cells) {
			cudaFree(ctx->mat_out_D);
			cudaFree(ctx->mat_out_N);
			cudaFree(ctx->mat_rows);
			ctx->mat_out_D = ctx->mat_out_N = 0;
			ctx->mat_rows = 0;
			ctx->mat_cells_cap = 0;
			e = cudaMalloc(&ctx->mat_out_D, cells * 8);
			if(e == cudaSuccess) e = cudaMalloc(&ctx->mat_out_N, cells * 8);
			if(e == cudaSuccess) e = cudaMalloc(&ctx->mat_rows, cells * 4);
			if(e == cudaSuccess) ctx->mat_cells_cap = cells;
		}
	}
	int2 *d_tiles = (int2 *) ctx->mat_tiles;
	void *d_D = ctx->mat_out_D, *d_N = N ? ctx->mat_out_N : 0;
	unsigned *d_rows = rows_inc ? ctx->mat_rows : 0;
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
		if(method == CCG_MAT_COS && !getenv("CCG_MAT_GENERIC_COS")) {
			cudaFuncSetAttribute(k_matdist_cos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(CosStage));
			k_matdist_cos<<<(unsigned) ((long long) ntiles * nslices), TH * TH, sizeof(CosStage), ctx->stream>>>(
			    (const uint32_t *) ctx->mat_counts, (const uint32_t *) ctx->mat_tot_over, ctx->mat_lpad, d_tiles, ntiles, nslices,
			    pos_per_slice, mp, ctx->mat_part_dist, ctx->mat_part_rows);
		} else {
			const size_t dyn = method == CCG_MAT_COS ? (size_t) 2 * MT * (MP + 1) * sizeof(double) : 0;
			if(dyn) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn);
			kern<<<(unsigned) ((long long) ntiles * nslices), TH * TH, dyn, ctx->stream>>>((const uint32_t *) ctx->mat_counts,
			                                                                              (const uint32_t *) ctx->mat_tot_over, ctx->mat_lpad, d_tiles,
			                                                                              ntiles, nslices, pos_per_slice, ctx->mat_lens, mp,
			                                                                              ctx->mat_part_dist, ctx->mat_part_rows);
		}
		cudaEventRecord(ctx->ev1, ctx->stream);
		ctx->ev_valid = 1;
		ctx->launches++;
		k_matdist_finalize<<<ntiles, TH * TH, 0, ctx->stream>>>(d_tiles, ntiles, nslices, ctx->mat_part_dist, ctx->mat_part_rows,
		                                                       ctx->mat_lens, ctx->mat_rank, n, norm, minLength, minCov, elem_size,
		                                                       byteScale, d_D, d_N, d_rows, row_slot, raw);
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
	snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "%s<%d> tiles=%d slices=%d",
	         method == CCG_MAT_COS && !getenv("CCG_MAT_GENERIC_COS") ? "k_matdist_cos" : "k_matdist", method, ntiles, nslices);
	return rc;
}

extern "C" int ccg_mat_run(ccg_ctx *ctx, const unsigned char *include, int method, unsigned order, double alpha, unsigned norm,
                           unsigned minDepth, unsigned minLength, double minCov, int elem_size, double byteScale, void *D,
                           void *N, int *Dn_out, uint32_t *rows_inc) {
	if(ctx && ctx->multi)
		return ccg_multi_mat_run(ctx, include, method, order, alpha, norm, minDepth, minLength, minCov, elem_size, byteScale, D, N, Dn_out,
		                         rows_inc);
	return mat_run_impl(ctx, include, method, order, alpha, norm, minDepth, minLength, minCov, elem_size, byteScale, D, N, Dn_out,
	                    rows_inc, -1);
}

/* cmpMatRowThrd (ltdmatrixthrd.c:111-181): the last uploaded sample against all the others */
/* A member of a position split: the raw per-pair sums over THIS context's positions -- dist[cell] the fp64 sum of
 * the per-position distances, rows[cell] the positions that counted (rowsInc of cmpMats, matcmp.c:470-481) -- packed
 * over the included samples, no gate, no scaling.  The caller adds the members' sums and hands them to
 * ccg_mat_finalize_host. */
extern "C" int ccg_mat_run_partial(ccg_ctx *ctx, const unsigned char *include, int method, unsigned order, double alpha,
                                   unsigned minDepth, double *dist, uint32_t *rows, int *Dn_out) {
	if(ctx && ctx->multi) {
		ccg_set_err(ctx, "ccg_mat_run_partial is the per-member call: use it on a single-device context");
		return CCG_ERR_UNSUPPORTED;
	}
	return mat_run_impl(ctx, include, method, order, alpha, 0, minDepth, 0, 0.0, 8, 1.0, dist, 0, Dn_out, rows, -1, 1);
}

static inline int host_cvttsd2si(double x) {
	if(!(x > -2147483649.0 && x < 2147483648.0)) return (int) 0x80000000;
	return (int) x;
}

/* The tail of cmpMats (matcmp.c:483-494) + the cell formats on the HOST, in the reference's own double arithmetic:
 * dist / rows: the summed raw sums of all members; lens: the whole length of every sample slot. */
extern "C" int ccg_mat_finalize_host(int n, const unsigned char *include, const int *lens, const double *dist, const uint32_t *rows,
                                     unsigned norm, unsigned minLength, double minCov, int elem_size, double byteScale, void *D, void *N,
                                     uint32_t *rows_inc, int *Dn_out) {
	if(n < 0 || !lens || !dist || !rows || !D) return CCG_ERR_ARG;
	if(elem_size != 8 && elem_size != 4 && elem_size != 2 && elem_size != 1) return CCG_ERR_ARG;
	std::vector<int> slot;
	for(int i = 0; i < n; ++i)
		if(!include || include[i]) slot.push_back(i);
	const int Dn = (int) slot.size();
	if(Dn_out) *Dn_out = Dn;
	long long cell = 0;
	for(int r = 1; r < Dn; ++r)
		for(int c = 0; c < r; ++c, ++cell) {
			double d, nn;
			bool ok;
			mat_cell(dist[cell], rows[cell], lens[slot[(size_t) r]], lens[slot[(size_t) c]], norm, minLength, minCov, &d, &nn, &ok);
			if(rows_inc) rows_inc[cell] = ok ? rows[cell] : 0u;
			if(elem_size == 8) {
				((double *) D)[cell] = d;
				if(N) ((double *) N)[cell] = nn;
			} else if(elem_size == 4) {
				((float *) D)[cell] = (float) d;
				if(N) ((float *) N)[cell] = (float) nn;
			} else {
				const int td = host_cvttsd2si(d * byteScale + 0.5), tn = host_cvttsd2si(nn * byteScale + 0.5);
				if(elem_size == 2) {
					((unsigned short *) D)[cell] = (unsigned short) td;
					if(N) ((unsigned short *) N)[cell] = (unsigned short) tn;
				} else {
					((unsigned char *) D)[cell] = (unsigned char) td;
					if(N) ((unsigned char *) N)[cell] = (unsigned char) tn;
				}
			}
		}
	return CCG_OK;
}

extern "C" int ccg_mat_run_row(ccg_ctx *ctx, int row_slot, int method, unsigned order, double alpha, unsigned norm,
                               unsigned minDepth, unsigned minLength, double minCov, double *D, double *N, uint32_t *rows_inc) {
	if(ctx && ctx->multi)
		return ccg_multi_forwarded(ctx, ccg_mat_run_row(ccg_multi_mat_solo(ctx), row_slot, method, order, alpha, norm, minDepth, minLength,
		                                                minCov, D, N, rows_inc));
	if(!ctx || !ctx->mat_counts || row_slot < 0 || row_slot >= ctx->mat_n) return CCG_ERR_ARG;
	if(ctx->world > 1) return CCG_ERR_UNSUPPORTED;
	/* columns are the slots below the row: everything above it stays out */
	std::vector<unsigned char> use((size_t) ctx->mat_n, 0);
	for(int i = 0; i <= row_slot; ++i) use[(size_t) i] = 1;
	int Dn = 0;
	return mat_run_impl(ctx, use.data(), method, order, alpha, norm, minDepth, minLength, minCov, 8, 1.0, D, N, &Dn, rows_inc,
	                    row_slot);
}
