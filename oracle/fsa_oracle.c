/*
 * fsa_oracle.c -- CPU restatement of ccphylo's `dist` FASTA hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see fsa_oracle.h).  Parity pinned by execution
 * against the unmodified reference built into oracle/_ref/.
 *
 * Written from the behavioural description of the reference (SURVEY.md
 * App. A), popcount based, with none of the reference's bit-serial loops.
 */
#include "fsa_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* byte -> code table                      (reference fsacmp.c:32-91)  */
/* ------------------------------------------------------------------ */
void orc_code_table(unsigned flag, unsigned char table[256]) {
	static const char unknown_upper[] = "N-RYSWKMBDHVX";
	static const char unknown_lower[] = "ryswkmbdhvx";
	const char *p;
	int low = (flag & 8) ? 1 : 0;

	memset(table, 32, 256);
	table['A'] = 0; table['C'] = 1; table['G'] = 2; table['T'] = 3; table['U'] = 3;
	for(p = unknown_upper; *p; ++p) table[(unsigned char) *p] = 4;
	for(p = unknown_lower; *p; ++p) table[(unsigned char) *p] = 4;
	/* lower case = "insignificant" calls: unknown unless flag bit 8 */
	table['a'] = low ? 0 : 4;
	table['c'] = low ? 1 : 4;
	table['g'] = low ? 2 : 4;
	table['t'] = low ? 3 : 4;
	table['u'] = low ? 3 : 4;
	table['n'] = 4;
}

/* (reference seqparse.c:195-248) */
long orc_translate(const unsigned char *bytes, long nbytes, unsigned flag,
                   unsigned char *codes) {
	unsigned char table[256];
	long i, len = 0;

	orc_code_table(flag, table);
	for(i = 0; i < nbytes; ++i) {
		unsigned char c = table[bytes[i]];
		if(c < 32) codes[len++] = c;
	}
	return len;
}

/* (reference qseqs.c:60-88) */
int orc_pack(const unsigned char *codes, int len, uint64_t *words) {
	int p, unknown = 0, W = orc_words(len);

	memset(words, 0, (size_t) W * sizeof(uint64_t));
	for(p = 0; p < len; ++p) {
		unsigned c = codes[p];
		if(c == 4) {
			++unknown;
		} else {
			words[p >> 5] |= (uint64_t) (c & 3) << (62 - 2 * (p & 31));
		}
	}
	return unknown;
}

/* (reference fsacmp.c:164-179, :181-238 with proxi == 0, :487-503) */
int orc_known_mask(const unsigned char *codes, int len, uint32_t *mask) {
	int p, inc = 0, W = orc_words(len);

	memset(mask, 0, (size_t) W * sizeof(uint32_t));
	for(p = 0; p < len; ++p) {
		if(codes[p] != 4) {
			mask[p >> 5] |= 1u << (31 - (p & 31));
			++inc;
		}
	}
	return inc;
}

/* (reference fsacmp.c:181-238 with proxi == 0, seq != ref) */
void orc_and_known(uint32_t *mask, const unsigned char *seq_codes,
                   const unsigned char *ref_codes, int len) {
	int p;

	for(p = 0; p < len; ++p) {
		if(seq_codes[p] == 4 || ref_codes[p] == 4) {
			mask[p >> 5] &= ~(1u << (31 - (p & 31)));
		}
	}
}

/* (reference fsacmp.c:487-503) */
int orc_mask_count(const uint32_t *mask, int len) {
	int w, inc = 0, W = orc_words(len);

	for(w = 0; w < W; ++w) inc += __builtin_popcount(mask[w]);
	return inc;
}

/* Spread the 32 mask bits to the even bit of each 2-bit lane. */
static inline uint64_t spread32(uint32_t m) {
	uint64_t x = m;
	x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
	x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
	x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
	x = (x | (x << 2)) & 0x3333333333333333ull;
	x = (x | (x << 1)) & 0x5555555555555555ull;
	return x;
}

/* number of 2-bit lanes that differ between a and b, restricted to the lanes
 * whose mask bit is set (mask bit k <-> lane bits 2k+1:2k). */
static inline uint32_t lane_mism(uint64_t a, uint64_t b, uint32_t m) {
	uint64_t x = a ^ b;
	uint64_t d = (x | (x >> 1)) & 0x5555555555555555ull;
	return (uint32_t) __builtin_popcountll(d & spread32(m));
}

/* (reference fsacmp.c:355-389 maskProxi with proxi == 0, :587-633 fsacmpair) */
void orc_pair_counts(const uint64_t *seq_i, const uint64_t *seq_j,
                     const uint32_t *inc_i, const uint32_t *inc_j, int len,
                     uint32_t *mism, uint32_t *ninc) {
	int w, W = orc_words(len);
	uint32_t d = 0, n = 0;

	for(w = 0; w < W; ++w) {
		uint32_t m = inc_i[w] & inc_j[w];
		n += (uint32_t) __builtin_popcount(m);
		d += lane_mism(seq_i[w], seq_j[w], m);
	}
	*mism = d;
	*ninc = n;
}

/* (reference fsacmp.c:552-585) */
uint32_t orc_masked_mism(const uint64_t *seq_i, const uint64_t *seq_j,
                         const uint32_t *mask, int len) {
	int w, W = orc_words(len);
	uint32_t d = 0;

	for(w = 0; w < W; ++w) d += lane_mism(seq_i[w], seq_j[w], mask[w]);
	return d;
}

/* ------------------------------------------------------------------ */
/* -y methylation motif masking                                        */
/* ------------------------------------------------------------------ */
/* (reference meth.c:52-159: matchMotif32 compares a window with up to four alternative words per motif, i.e.
 * position k matches when the sequence base is one of the bases the motif's IUPAC letter stands for; matchMotif
 * tries every start 0 .. len - motif length; maskMotif clears the upper-case positions of the motif.  Motifs of
 * more than 32 positions shift by a negative count there (meth.c:99) and are not restated.) */
long orc_mask_motifs(const uint64_t *seq, uint32_t *mask, int len, int nmotifs, const int *lens,
                     const unsigned char *sets) {
	long found = 0;
	int m;

	for(m = 0; m < nmotifs; sets += lens[m], ++m) {
		const int L = lens[m];
		long p;
		for(p = 0; p + L <= len; ++p) {
			int k, ok = 1;
			for(k = 0; k < L && ok; ++k) {
				const long q = p + k;
				const unsigned code = (unsigned) (seq[q >> 5] >> (62 - 2 * (q & 31))) & 3;
				ok = (sets[k] >> code) & 1;
			}
			if(!ok) continue;
			++found;
			for(k = 0; k < L; ++k) {
				const long q = p + k;
				if(sets[k] & 16) mask[q >> 5] &= ~(1u << (31 - (q & 31)));
			}
		}
	}
	return found;
}

/* ------------------------------------------------------------------ */
/* -V variant listing                                                  */
/* ------------------------------------------------------------------ */
/* (reference fsacmp.c:646-683 fsacmprint, :685-737 fsacmpairint, printDiff :635-644)
 * The reference walks a word only when the mask word is non-zero and the two packed words differ (anywhere,
 * masked or not); it then reads lanes from the LEAST significant end -- lane k is base 31 - k of the word --
 * while labelling them pos, pos + 1, ... from a counter that starts at 1, and it stops at the highest set mask
 * bit, so the counter advances by (index of that bit + 1) for a walked word and by 32 for every other word.
 * The printed positions are therefore not alignment coordinates; they are reproduced as
 * they are. */
long orc_list_variants(const uint64_t *seq_i, const uint64_t *seq_j, const uint32_t *mask, int len,
                       uint64_t *out, long cap) {
	int w, W = orc_words(len);
	unsigned label = 1;
	long count = 0;

	for(w = 0; w < W; ++w) {
		const uint32_t inc = mask[w];
		if(inc && seq_i[w] != seq_j[w]) {
			int k;
			for(k = 0; k < 32 && (inc >> k) != 0; ++k) {
				const unsigned ci = (unsigned) (seq_i[w] >> (2 * k)) & 3, cj = (unsigned) (seq_j[w] >> (2 * k)) & 3;
				if(((inc >> k) & 1) && ci != cj) {
					if(out && count < cap) out[count] = ((uint64_t) (label + (unsigned) k) << 4) | (ci << 2) | cj;
					++count;
				}
			}
			label += (unsigned) k;
		} else label += 32;
	}
	return count;
}

/* ------------------------------------------------------------------ */
/* -P proximity masking                                                */
/* ------------------------------------------------------------------ */
static inline void clear_range(uint32_t *mask, long lo, long hi, int len) {
	long p;
	if(lo < 0) lo = 0;
	if(hi >= len) hi = (long) len - 1;   /* positions >= len do not exist (the reference's stores there are out of bounds) */
	for(p = lo; p <= hi; ++p) mask[p >> 5] &= ~(1u << (31 - (p & 31)));
}

/* (reference fsacmp.c:181-238 getIncPos, :240-295 getIncPosInsigPrune, :297-353 getIncPosInsig,
 * selected by -f 32 / -f 8 at dist.c:802-806)
 * Positions where seq or ref is unknown are cleared.  An "event" is, for getIncPos (variant 0), every
 * position with seq != ref or seq unknown; for the other two (variant 1) every position where both are
 * known and differ.  When an event lies at most proxi after the previous event, everything from the
 * previous event to this one (both inclusive) is cleared; the first event never clears anything (the
 * reference's range loop compares int -1 as unsigned). */
void orc_inc_pos(uint32_t *mask, const unsigned char *seq_codes, const unsigned char *ref_codes, int len,
                 unsigned proxi, int variant) {
	long p, last = -1;

	for(p = 0; p < len; ++p) {
		const unsigned c = seq_codes[p], r = ref_codes[p];
		const int unknown = c == 4 || r == 4;
		const int event = variant == 0 ? (c != r || c == 4) : (!unknown && c != r);
		if(unknown) mask[p >> 5] &= ~(1u << (31 - (p & 31)));
		if(event) {
			if(last >= 0 && (unsigned long) (p - last) <= proxi) clear_range(mask, last, p, len);
			last = p;
		}
	}
}

/* (reference fsacmp.c:181-353 on the code bytes of getIupacBitTable, as fsaTrim trim.c:77-260 calls them)
 * Per position, c = the sample's byte, r = the reference sample's stored byte (its own pass stripped every soft flag):
 *   getIncPos            event: c != r, c unknown or c soft;  cleared: c or r unknown, or c soft (then c &= 15 unless unknown)
 *   getIncPosInsigPrune  cleared: c or r unknown, or c soft (c &= 15 as above) -- no event there;  otherwise event: c != r
 *   getIncPosInsig       cleared: c or r unknown;  otherwise event: c != r (a soft c differs from the stripped r)
 * The sample against itself is getIncPos with r = c: unknown and soft positions are cleared and are events, every
 * soft flag goes.  Ranges as in orc_inc_pos. */
void orc_trim_pass(uint32_t *mask, unsigned char *seq, const unsigned char *ref, int len, unsigned proxi, int builder,
                   uint32_t *columns) {
	long p, last = -1;

	for(p = 0; p < len; ++p) {
		const unsigned c = seq[p], r = ref ? ref[p] : (unsigned) (seq[p] & 15u);
		const int unknown = c == 4 || r == 4, soft = (c & 16u) != 0;
		int event, clear;
		if(!ref || builder == 0) {
			event = c != r || c == 4 || soft;
			clear = unknown || soft;
			if(!unknown && soft) seq[p] = (unsigned char) (c & 15u);
		} else if(builder == 2) {
			clear = unknown || soft;
			event = !clear && c != r;
			if(!unknown && soft) seq[p] = (unsigned char) (c & 15u);
		} else {
			clear = unknown;
			event = !clear && c != r;
		}
		if(clear) mask[p >> 5] &= ~(1u << (31 - (p & 31)));
		if(columns && ref && seq[p] != r) columns[p >> 5] |= 1u << (31 - (p & 31));
		if(event) {
			if(last >= 0 && (unsigned long) (p - last) <= proxi) clear_range(mask, last, p, len);
			last = p;
		}
	}
}

/* (reference fsacmp.c:355-485 maskProxi with proxi > 0, then :587-633 fsacmpair)
 * inc = inc_i & inc_j; the SNPs of the pair are the included positions whose 2-bit codes differ.  The
 * reference walks them from the last to the first with a position counter that is one too high, so for
 * two neighbouring SNPs p < q with q - p <= proxi it clears p+1 .. q+1: q and everything between go, p
 * stays.  Counts are then taken under the cleared mask (orc_pair_counts_proxi); -V walks the
 * same mask (fsacmpthrd.c:410-414), so it is also handed out as it is (orc_mask_proxi, m = ceil(len/32) words). */
void orc_mask_proxi(const uint64_t *seq_i, const uint64_t *seq_j, const uint32_t *inc_i, const uint32_t *inc_j, int len,
                    unsigned proxi, uint32_t *m) {
	int w, W = orc_words(len);
	long *snp, ns = 0, k;

	for(w = 0; w < W; ++w) m[w] = inc_i[w] & inc_j[w];
	if(!proxi) return;
	snp = malloc((size_t) (len ? len : 1) * sizeof(long));
	for(w = 0; w < W; ++w) {
		int b;
		for(b = 0; b < 32; ++b) {
			if((m[w] >> (31 - b)) & 1) {
				const unsigned ci = (unsigned) (seq_i[w] >> (62 - 2 * b)) & 3, cj = (unsigned) (seq_j[w] >> (62 - 2 * b)) & 3;
				if(ci != cj) snp[ns++] = (long) w * 32 + b;
			}
		}
	}
	for(k = 0; k + 1 < ns; ++k) {
		if((unsigned long) (snp[k + 1] - snp[k]) <= proxi) clear_range(m, snp[k] + 1, snp[k + 1] + 1, len);
	}
	free(snp);
}

void orc_pair_counts_proxi(const uint64_t *seq_i, const uint64_t *seq_j, const uint32_t *inc_i,
                           const uint32_t *inc_j, int len, unsigned proxi, uint32_t *mism, uint32_t *ninc) {
	int w, W = orc_words(len);
	uint32_t *m, d = 0, n = 0;

	if(!proxi) {
		orc_pair_counts(seq_i, seq_j, inc_i, inc_j, len, mism, ninc);
		return;
	}
	m = malloc((size_t) (W ? W : 1) * sizeof(uint32_t));
	orc_mask_proxi(seq_i, seq_j, inc_i, inc_j, len, proxi, m);
	for(w = 0; w < W; ++w) {
		n += (uint32_t) __builtin_popcount(m[w]);
		d += lane_mism(seq_i[w], seq_j[w], m[w]);
	}
	free(m);
	*mism = d;
	*ninc = n;
}

/* ------------------------------------------------------------------ */
/* raw all-vs-all integer matrices (test helper, threaded over rows)   */
/* ------------------------------------------------------------------ */
typedef struct {
	int n, len, tid, nthreads;
	long wstride;
	const uint64_t *seqs;
	const uint32_t *masks;
	uint32_t *mism, *ninc;
} RawJob;

static void *raw_worker(void *arg) {
	RawJob *job = arg;
	int i, j;

	/* interleave rows so each thread gets a similar number of cells */
	for(i = 1 + job->tid; i < job->n; i += job->nthreads) {
		size_t row = (size_t) i * (i - 1) / 2;
		for(j = 0; j < i; ++j) {
			orc_pair_counts(job->seqs + i * job->wstride, job->seqs + j * job->wstride,
			                job->masks + i * job->wstride, job->masks + j * job->wstride,
			                job->len, job->mism + row + j, job->ninc + row + j);
		}
	}
	return 0;
}

void orc_raw_pair_matrix(int n, int len, const uint64_t *seqs,
                         const uint32_t *masks, long wstride,
                         uint32_t *mism, uint32_t *ninc, int nthreads) {
	pthread_t ids[256];
	RawJob jobs[256];
	int t;

	if(nthreads < 1) nthreads = 1;
	if(nthreads > 256) nthreads = 256;
	for(t = 0; t < nthreads; ++t) {
		jobs[t] = (RawJob){n, len, t, nthreads, wstride, seqs, masks, mism, ninc};
		if(t && pthread_create(ids + t, 0, raw_worker, jobs + t)) {
			raw_worker(jobs + t);
			ids[t] = 0;
		}
	}
	raw_worker(jobs);
	for(t = 1; t < nthreads; ++t) {
		if(ids[t]) pthread_join(ids[t], 0);
	}
}

/* ------------------------------------------------------------------ */
/* epilogues                                                           */
/* ------------------------------------------------------------------ */

/* double -> integer cell as gcc/x86-64 does it for the reference's
 * `unsigned short = double` / `unsigned char = double` stores: a truncating
 * 32-bit cvttsd2si whose low bits are kept. */
static inline int32_t trunc_i32(double x) {
	if(!(x > -2147483649.0 && x < 2147483648.0)) return INT32_MIN;
	return (int32_t) x;
}

/* (reference fsacmpthrd.c:419-475; bytescale.h:22 dtouc) */
void orc_pair_cell(uint32_t mism, uint32_t inc, unsigned norm,
                   unsigned minLength, int elem_size, double byteScale,
                   void *Dcell, void *Ncell) {
	uint64_t scaled = (uint64_t) mism * norm;
	int ok = minLength <= inc;

	if(elem_size == 8) {
		double d;
		if(!ok) d = -1.0;
		else if(norm) { d = (double) scaled; d /= inc; }
		else d = (double) mism;
		*(double *) Dcell = d;
		if(Ncell) *(double *) Ncell = (double) inc;
	} else if(elem_size == 4) {
		float f;
		if(!ok) f = -1.0f;
		else if(norm) { f = (float) scaled; f /= inc; }
		else f = (float) mism;
		*(float *) Dcell = f;
		if(Ncell) *(float *) Ncell = (float) inc;
	} else {
		double d, nn;
		if(!ok) d = -1.0 * byteScale + 0;
		else if(norm) d = ((double) scaled * byteScale + 0.5) / inc;
		else d = (double) (uint64_t) mism * byteScale + 0.5;
		nn = inc * byteScale + 0.5;
		if(elem_size == 2) {
			*(uint16_t *) Dcell = (uint16_t) trunc_i32(d);
			if(Ncell) *(uint16_t *) Ncell = (uint16_t) trunc_i32(nn);
		} else {
			*(uint8_t *) Dcell = (uint8_t) trunc_i32(d);
			if(Ncell) *(uint8_t *) Ncell = (uint8_t) trunc_i32(nn);
		}
	}
}

static int compact_included(int n, const unsigned char *include, int *idx) {
	int i, Dn = 0;
	for(i = 0; i < n; ++i) {
		if(include[i]) idx[Dn++] = i;
	}
	return Dn;
}

/* (reference fsacmpthrd.c:261-480) */
int orc_fsa_cmp_pair(int n, int len, const uint64_t *seqs, long wstride,
                     const unsigned char *include, const uint32_t *masks,
                     unsigned norm, unsigned minLength, double minCov,
                     int elem_size, double byteScale, void *D, void *N) {
	return orc_fsa_cmp_pair_proxi(n, len, seqs, wstride, include, masks, norm, minLength, minCov, 0, elem_size,
	                              byteScale, D, N);
}

/* the same with -P proxi (fsacmpthrd.c:409-410) */
int orc_fsa_cmp_pair_proxi(int n, int len, const uint64_t *seqs, long wstride,
                           const unsigned char *include, const uint32_t *masks,
                           unsigned norm, unsigned minLength, double minCov, unsigned proxi,
                           int elem_size, double byteScale, void *D, void *N) {
	int *idx = malloc((size_t) (n > 0 ? n : 1) * sizeof(int));
	int Dn = compact_included(n, include, idx), r, c;
	size_t cell = 0;

	/* fsacmpthrd.c:292: threshold re-derived from minCov, truncating */
	if(minLength < minCov * len) minLength = (unsigned) (minCov * len);

	for(r = 1; r < Dn; ++r) {
		for(c = 0; c < r; ++c, ++cell) {
			uint32_t mism, inc;
			int i = idx[r], j = idx[c];
			orc_pair_counts_proxi(seqs + i * wstride, seqs + j * wstride,
			                      masks + i * wstride, masks + j * wstride, len, proxi, &mism, &inc);
			orc_pair_cell(mism, inc, norm, minLength, elem_size, byteScale,
			              (char *) D + cell * elem_size,
			              N ? (char *) N + cell * elem_size : 0);
		}
	}
	free(idx);
	return Dn;
}

/* (reference fsacmpthrd.c:108-259; intended pair selection) */
int orc_fsa_cmp_global(int n, int len, const uint64_t *seqs, long wstride,
                       const unsigned char *include, const uint32_t *mask,
                       unsigned norm, int elem_size, double byteScale,
                       void *D, unsigned *global_inc) {
	int *idx = malloc((size_t) (n > 0 ? n : 1) * sizeof(int));
	int Dn = compact_included(n, include, idx), r, c;
	int inc = orc_mask_count(mask, len);
	double nFactor;
	size_t cell = 0;

	if(global_inc) *global_inc = (unsigned) inc;
	/* fsacmpthrd.c:171-176 */
	if(norm) { nFactor = norm; nFactor /= inc; }
	else nFactor = 1.0;

	for(r = 1; r < Dn; ++r) {
		for(c = 0; c < r; ++c, ++cell) {
			uint64_t dist = orc_masked_mism(seqs + idx[r] * wstride, seqs + idx[c] * wstride, mask, len);
			double v = nFactor * dist;          /* fsacmpthrd.c:247-255 */
			char *cellp = (char *) D + cell * elem_size;
			if(elem_size == 8) *(double *) cellp = v;
			else if(elem_size == 4) *(float *) cellp = (float) v;
			else if(elem_size == 2) *(uint16_t *) cellp = (uint16_t) trunc_i32(v * byteScale + 0.5);
			else *(uint8_t *) cellp = (uint8_t) trunc_i32(v * byteScale + 0.5);
		}
	}
	free(idx);
	return Dn;
}
