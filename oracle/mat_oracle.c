/*
 * mat_oracle.c -- CPU restatement of ccphylo's `dist` count-matrix (.mat) path.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as fsa_oracle.h): nothing under ccphylo_b200/ or
 * include/ may link, import or call this.
 *
 * Parity status: PINNED BY EXECUTION -- tests/golden/mat_dist.json holds the .phy / .num /
 * stderr text the unmodified reference binary (oracle/_ref/ccphylo) printed for the inputs
 * scripts/make_golden_mat.py generated; tests/test_mat_oracle_golden.py requires this file to
 * reproduce that text byte for byte (every -d method, -W, -E, gz input, union input).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 * Counts are 6 x u16 per position in the reference's storage order A, C, G, T, -, N
 * (matparse.c:254-259); tot = row total.  Reference bugs that change results are kept and
 * marked.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

enum {
	M_COS = 0, M_Z, M_CHI2, M_NCHI2, M_C, M_NC, M_P, M_NP, M_BC, M_NBC, M_L1, M_L2, M_LINF, M_LN, M_NL1, M_NL2, M_NLINF, M_NLN
};

/* stdstat.c:32-130 fastp, only the branch p_chisqr can reach (q > 49) */
static double fastp(long double q) {
	static const double cut[16] = {114.5242, 109.9604, 105.3969, 100.8337, 96.27476, 91.71701, 87.16164, 82.60901, 78.05917,
	                               73.51245, 68.96954, 64.43048, 59.89615, 55.36699, 50.84417, 46.32844};
	static const double val[16] = {1e-26, 1e-25, 1e-24, 1e-23, 1e-22, 1e-21, 1e-20, 1e-19, 1e-18, 1e-17, 1e-16, 1e-15,
	                               1e-14, 1e-13, 1e-12, 1e-11};
	for(int k = 0; k < 16; ++k)
		if(q > cut[k]) return val[k];
	return 1e-10;
}

/* stdstat.c:132-144 p_chisqr */
static double p_chisqr(long double q) {
	if(q < 0) return 1e-26;
	if(q > 49) return fastp(q);
	return 1 - 1.772453850 * erf(sqrt(0.5 * q)) / tgamma(0.5);
}

/* matcmp.c:63-446: one per-position distance; < 0 (or NaN) means "position not comparable" */
double orc_mat_veccmp(int method, unsigned order, double alpha, const uint16_t *c1, const uint16_t *c2, int tot1, int tot2) {
	int i, t1, t2;
	double d, tmp, f1, f2, T;
	switch(method) {
		case M_COS: {                                       /* coscmp matcmp.c:420-446: int products */
			unsigned long a = 0, b = 0;
			d = 0;
			for(i = 0; i < 5; ++i) {
				int x = c1[i], y = c2[i];
				d += x * y;
				a += x * x;
				b += y * y;
			}
			if(!a || !b) return -1;
			d = 1 - d / (sqrt(a) * sqrt(b));
			return d < 0 ? 0 : d;
		}
		case M_Z: {                                         /* zcmp matcmp.c:311-344; #13: the arg-max indices are overwritten */
			int max1 = c1[0], max2 = c2[0], x1, x2;
			for(i = 1; i < 5; ++i) {
				if(max1 < c1[i]) max1 = c1[i];
				if(max2 < c2[i]) max2 = c2[i];
			}
			x1 = p_chisqr(pow(tot1 - (max1 << 1), 2) / tot1) <= alpha && tot1 < (max1 << 1);
			x2 = p_chisqr(pow(tot2 - (max2 << 1), 2) / tot2) <= alpha && tot1 < (max1 << 1);
			if(x1 && x2) return x1 == x2 ? 0 : 1;
			return -1;
		}
		case M_CHI2:                                        /* chi2cmp matcmp.c:383-396 */
		case M_P:                                           /* pcmp matcmp.c:346-359 */
			d = 0;
			for(i = 0; i < 5; ++i)
				if((T = c1[i] - c2[i])) d += T * T / (c1[i] + c2[i]);
			return method == M_CHI2 ? sqrt(d) : 1 - p_chisqr(d);
		case M_NCHI2:                                       /* nchi2cmp matcmp.c:398-418 */
		case M_NP:                                          /* npcmp matcmp.c:361-381 */
			t1 = tot1 - c1[5];
			t2 = tot2 - c2[5];
			d = 0;
			for(i = 0; i < 5; ++i) {
				f1 = (double) c1[i] / t1;
				f2 = (double) c2[i] / t2;
				if((tmp = f1 - f2) != 0) d += tmp * tmp / (f1 + f2);
			}
			return method == M_NCHI2 ? sqrt(d) : 1 - p_chisqr(d);
		case M_C: {                                         /* ccmp matcmp.c:281-309 */
			int big = 0;
			d = 0;
			for(i = 0; i < 5; ++i) {
				if(c1[i] < c2[i]) { d += c1[i]; big += c2[i]; }
				else { d += c2[i]; big += c1[i]; }
			}
			if(!big) return -1;
			d = 1 - d / big;
			return d < 0 ? 0 : d;
		}
		case M_NC:                                          /* nccmp matcmp.c:246-279; #14: T restarts at 1 in every step */
			t1 = tot1 - c1[5];
			t2 = tot2 - c2[5];
			f1 = (double) c1[0] / t1;
			f2 = (double) c2[0] / t2;
			if(f1 < f2) { d = f1; T = f2; } else { d = f2; T = f1; }
			for(i = 1; i < 5; ++i) {
				f1 = (double) c1[i] / t1;
				f2 = (double) c2[i] / t2;
				T = 1;
				if(f1 < f2) { d += f1; T += f2; } else { d += f2; T += f1; }
			}
			d = 1 - d / T;
			return d < 0 ? 0 : d;
		case M_BC:                                          /* bccmp matcmp.c:230-244 */
			d = 0;
			for(i = 0; i < 5; ++i) d += c1[i] < c2[i] ? c1[i] : c2[i];
			d /= (tot1 - c1[5] + tot2 - c2[5]);
			d = 1 - 2 * d;
			return d < 0 ? 0 : d;
		case M_NBC:                                         /* nbccmp matcmp.c:209-228 */
			t1 = tot1 - c1[5];
			t2 = tot2 - c2[5];
			d = 0;
			for(i = 0; i < 5; ++i) {
				f1 = (double) c1[i] / t1;
				f2 = (double) c2[i] / t2;
				d += f1 < f2 ? f1 : f2;
			}
			d = 1 - d;
			return d < 0 ? 0 : d;
		case M_L1: {                                        /* l1cmp matcmp.c:145-158 */
			int s = 0;
			for(i = 0; i < 5; ++i) s += abs(c1[i] - c2[i]);
			return s;
		}
		case M_L2: {                                        /* l2cmp matcmp.c:160-173 */
			int s = 0;
			for(i = 0; i < 5; ++i) s += (c1[i] - c2[i]) * (c1[i] - c2[i]);
			return sqrt(s);
		}
		case M_LINF: {                                      /* linfcmp matcmp.c:196-207 */
			int s = 0;
			for(i = 0; i < 5; ++i)
				if(s < abs(c1[i] - c2[i])) s = abs(c1[i] - c2[i]);
			return s;
		}
		case M_LN:                                          /* lncmp matcmp.c:175-194 */
			d = 0;
			for(i = 0; i < 5; ++i) d += pow(abs(c1[i] - c2[i]), order);
			d = pow(d, 1.0 / order);
			return d < 0 ? 0 : d;
		case M_NL1:                                         /* nl1cmp matcmp.c:63-79 */
		case M_NL2:                                         /* nl2cmp matcmp.c:81-97 */
			t1 = tot1 - c1[5];
			t2 = tot2 - c2[5];
			d = 0;
			for(i = 0; i < 5; ++i) {
				tmp = (double) c1[i] / t1 - (double) c2[i] / t2;
				d += method == M_NL1 ? (tmp < 0 ? -tmp : tmp) : tmp * tmp;
			}
			return method == M_NL1 ? d : sqrt(d);
		case M_NLINF:                                       /* nlinfcmp matcmp.c:125-143; #14: the pointers never advance */
			t1 = tot1 - c1[5];
			t2 = tot2 - c2[5];
			tmp = (double) c1[0] / t1 - (double) c2[0] / t2;
			return tmp < 0 ? -tmp : tmp;
		default:                                            /* nlncmp matcmp.c:99-123: the first term keeps its sign */
			t1 = tot1 - c1[5];
			t2 = tot2 - c2[5];
			d = pow((double) c1[0] / t1 - (double) c2[0] / t2, order);
			for(i = 1; i < 5; ++i) {
				tmp = (double) c1[i] / t1 - (double) c2[i] / t2;
				tmp = tmp < 0 ? -tmp : tmp;
				d += pow(tmp, order);
			}
			d = pow(d, 1.0 / order);
			return d < 0 ? 0 : d;
	}
}

/* cmpMats matcmp.c:448-494 for one pair: sample i was loaded (ci, ti, len_i), sample j (the earlier
 * one) is streamed.  Returns the value cmpMats returns (-1: insufficient overlap, -2: sample j fails
 * its own gate) and the rowsInc it leaves in mat2->total. */
double orc_mat_pair(int method, unsigned order, double alpha, const uint16_t *ci, const uint32_t *ti, int len_i,
                    const uint16_t *cj, const uint32_t *tj, int len_j, unsigned norm, unsigned minDepth, unsigned minLength,
                    double minCov, unsigned *rows_inc) {
	unsigned rowNum = 0, rowsInc = 0, nNucs = 0;
	double dist = 0, d;
	*rows_inc = 0;
	for(int p = 0; p < len_j; ++p) {
		if((unsigned) len_i < ++rowNum) return -1;
		if(minDepth <= tj[p]) {
			++nNucs;
			if(minDepth <= ti[p] && 0 <= (d = orc_mat_veccmp(method, order, alpha, ci + 6 * (long) p, cj + 6 * (long) p, (int) ti[p], (int) tj[p]))) {
				dist += d;
				++rowsInc;
			}
		}
	}
	if(nNucs < minLength || nNucs < minCov * rowNum) return -2.0;
	if(rowsInc < minLength || rowsInc < minCov * rowNum) return -1.0;
	*rows_inc = rowsInc;
	return norm ? dist / rowsInc * norm : dist;
}

/* All pairs over the included samples (ltdmatrixthrd.c:182-375 cell order): counts / totals are
 * [n][lmax] arrays, lens[n] the rows of each sample.  D, N: packed doubles over the included
 * samples; a -1 pair gets D = -1, N = 0.  Returns Dn, or -1 when a pair returns -2 (the reference
 * exits there). */
int orc_mat_matrix(int method, unsigned order, double alpha, int n, long lmax, const uint16_t *counts, const uint32_t *totals,
                   const int *lens, const unsigned char *include, unsigned norm, unsigned minDepth, unsigned minLength,
                   double minCov, double *D, double *N) {
	int Dn = 0;
	long cell = 0;
	for(int i = 0; i < n; ++i) {
		if(include && !include[i]) continue;
		for(int j = 0; j < i; ++j) {
			if(include && !include[j]) continue;
			unsigned rows = 0;
			double v = orc_mat_pair(method, order, alpha, counts + 6 * lmax * i, totals + lmax * i, lens[i], counts + 6 * lmax * j,
			                        totals + lmax * j, lens[j], norm, minDepth, minLength, minCov, &rows);
			if(v == -2.0) return -1;
			D[cell] = v;
			N[cell] = rows;
			++cell;
		}
		++Dn;
	}
	return Dn;
}

/* Test / bench infrastructure: writes one KMA count matrix (.mat, plain text) for the reference binary to read --
 * "#name", then per position "ref\tA\tC\tG\tT\tN\t-" (the file order; counts6 holds A,C,G,T,-,N), then a blank line. */
#include <stdio.h>
int orc_write_mat(const char *path, const char *name, const unsigned char *refbases, const uint16_t *counts6, long len) {
	FILE *f = fopen(path, "w");
	long p;
	if(!f) return -1;
	fprintf(f, "#%s\n", name);
	for(p = 0; p < len; ++p) {
		const uint16_t *c = counts6 + 6 * p;
		fprintf(f, "%c\t%u\t%u\t%u\t%u\t%u\t%u\n", "ACGT"[refbases[p] & 3], c[0], c[1], c[2], c[3], c[5], c[4]);
	}
	fputc('\n', f);
	return fclose(f);
}
