/*
 * proxi_core.h -- the arithmetic of -P proximity masking, shared by the CUDA kernels of
 * k_proxi.cu and by a host-compiled unit test (tests/csrc/proxi_core_test.cpp) that runs the very
 * same functions word by word against the oracle.  No memory traffic in here except through
 * the `Sink` of proxi_scan_words.
 *
 * Reference behaviour restated:
 *
 *  per pair  -- maskProxi fsacmp.c:355-485 followed by fsacmpair :587-633.  inc = inc_i & inc_j;
 *     the SNPs of the pair are the included positions whose 2-bit codes differ.  The reference
 *     walks them downwards with a position counter that is one too high: for neighbouring SNPs
 *     p < q with q - p <= proxi it clears p+1 .. q+1.  Over a maximal chain a = s_1 < .. < s_k = b
 *     of SNPs whose gaps are all <= proxi (a "cluster", k >= 2) that clears a+1 .. b+1: a is the
 *     only SNP left.  Hence
 *         mismatches = number of clusters (chains of length 1 included)
 *         included   = popcount(inc) - sum over clusters with k >= 2 of popcount(inc[a+1 .. b+1])
 *     and both follow from one ascending pass that remembers the last SNP, the included
 *     positions seen since, and the inclusion bit right after it.
 *
 *  per sample -- getIncPos fsacmp.c:181-238 (events: seq != ref or seq unknown), getIncPosInsig /
 *     getIncPosInsigPrune :240-353 (events: both known and different; -f 8 / -f 32, dist.c:802):
 *     an event at most proxi after the previous event clears everything from the previous
 *     event to this one, both inclusive.
 */
#ifndef CCG_PROXI_CORE_H
#define CCG_PROXI_CORE_H

#include <stdint.h>

#ifdef __CUDACC__
#define CCG_HD __host__ __device__ __forceinline__
#else
#define CCG_HD static inline
#endif

CCG_HD int ccg_popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
	return __popc(x);
#else
	return __builtin_popcount(x);
#endif
}

/* index (0 = most significant = first base of the word) of the first set bit; x != 0 */
CCG_HD int ccg_first_bit(uint32_t x) {
#ifdef __CUDA_ARCH__
	return __clz((int) x);
#else
	return __builtin_clz(x);
#endif
}

/* ------------------------------------------------------------------------------------------
 * per pair
 * ------------------------------------------------------------------------------------------ */
enum { PROXI_IN_CLUSTER = 1, PROXI_AFTER_INC = 2, PROXI_NEED_AFTER = 4 };

struct ProxiPairState {
	int last;            /* position of the last SNP so far, -1 = none */
	unsigned gap;        /* included positions after `last` seen so far */
	unsigned clusters;   /* = surviving mismatches */
	unsigned cleared;    /* included positions removed by proximity masking */
	unsigned total;      /* popcount(inc) */
	unsigned flags;      /* PROXI_IN_CLUSTER: `last` closes a chain of >= 2 SNPs; PROXI_AFTER_INC: the position right
	                        after `last` is included; PROXI_NEED_AFTER: that position is the first of the next word */
};

CCG_HD void proxi_pair_init(ProxiPairState &st) {
	st.last = -1;
	st.gap = st.clusters = st.cleared = st.total = st.flags = 0;
}

/* One 32-base word, words in ascending order.  p0 = position of the word's first base; d = SNP bits
 * (already restricted to m); m = inc_i & inc_j; bit 31 - k <-> base p0 + k. */
CCG_HD void proxi_pair_word(ProxiPairState &st, int p0, uint32_t d, uint32_t m, unsigned proxi) {
	st.total += (unsigned) ccg_popc32(m);
	if(st.flags & PROXI_NEED_AFTER) {
		st.flags &= ~(unsigned) PROXI_NEED_AFTER;
		if(m >> 31) st.flags |= PROXI_AFTER_INC;
	}
	if(d == 0) {
		st.gap += (unsigned) ccg_popc32(m);
		return;
	}
	uint32_t rest = m;                           /* included bits of this word not yet added to gap */
	while(d) {
		const int b = ccg_first_bit(d);
		const uint32_t bit = 0x80000000u >> b;
		d &= ~bit;
		st.gap += (unsigned) ccg_popc32(rest & ~(bit | (bit - 1u)));   /* bases before b */
		rest &= bit - 1u;                                               /* bases after b */
		const int p = p0 + b;
		if(st.last >= 0 && (unsigned) (p - st.last) <= proxi) {
			st.cleared += st.gap + 1u;           /* last+1 .. p, p itself is included */
			st.flags |= PROXI_IN_CLUSTER;
		} else {
			st.clusters += 1u;
			if((st.flags & (PROXI_IN_CLUSTER | PROXI_AFTER_INC)) == (PROXI_IN_CLUSTER | PROXI_AFTER_INC)) st.cleared += 1u;
			st.flags &= ~(unsigned) PROXI_IN_CLUSTER;
		}
		st.last = p;
		st.gap = 0;
		st.flags &= ~(unsigned) (PROXI_AFTER_INC | PROXI_NEED_AFTER);
		if(b < 31) {
			if(m & (bit >> 1)) st.flags |= PROXI_AFTER_INC;
		} else st.flags |= PROXI_NEED_AFTER;
	}
	st.gap += (unsigned) ccg_popc32(rest);
}

/* One 128-base chunk (four words, ascending) in one go: the same state transitions as four proxi_pair_word calls,
 * but with ONE loop over the chunk's SNPs.  On the device the lanes of a warp walk different pairs, so a loop runs
 * as often as its busiest lane needs: per word that is "at least once" for 86 % of the words at typical SNP
 * densities, per chunk it is about two rounds for four words. */
CCG_HD void proxi_pair_chunk(ProxiPairState &st, int p0, uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3, uint32_t m0,
                             uint32_t m1, uint32_t m2, uint32_t m3, unsigned proxi) {
	/* included positions before word w of the chunk */
	const unsigned c1 = (unsigned) ccg_popc32(m0), c2 = c1 + (unsigned) ccg_popc32(m1), c3 = c2 + (unsigned) ccg_popc32(m2),
	               c4 = c3 + (unsigned) ccg_popc32(m3);
	st.total += c4;
	if(st.flags & PROXI_NEED_AFTER) {
		st.flags &= ~(unsigned) PROXI_NEED_AFTER;
		if(m0 >> 31) st.flags |= PROXI_AFTER_INC;
	}
	unsigned base = 0;                           /* included positions of the chunk already accounted for in gap / cleared */
	while(d0 | d1 | d2 | d3) {
		const int w = d0 ? 0 : d1 ? 1 : d2 ? 2 : 3;
		const uint32_t dw = w == 0 ? d0 : w == 1 ? d1 : w == 2 ? d2 : d3;
		const uint32_t mw = w == 0 ? m0 : w == 1 ? m1 : w == 2 ? m2 : m3;
		const uint32_t mnext = w == 0 ? m1 : w == 1 ? m2 : m3;                   /* unused for w == 3 */
		const unsigned cw = w == 0 ? 0u : w == 1 ? c1 : w == 2 ? c2 : c3;
		const int b = ccg_first_bit(dw);
		const uint32_t bit = 0x80000000u >> b;
		if(w == 0) d0 &= ~bit; else if(w == 1) d1 &= ~bit; else if(w == 2) d2 &= ~bit; else d3 &= ~bit;
		const unsigned upto = cw + (unsigned) ccg_popc32(mw & ~(bit | (bit - 1u)));   /* included before the SNP */
		st.gap += upto - base;
		base = upto + 1u;                                                        /* the SNP itself is included */
		const int p = p0 + 32 * w + b;
		if(st.last >= 0 && (unsigned) (p - st.last) <= proxi) {
			st.cleared += st.gap + 1u;
			st.flags |= PROXI_IN_CLUSTER;
		} else {
			st.clusters += 1u;
			if((st.flags & (PROXI_IN_CLUSTER | PROXI_AFTER_INC)) == (PROXI_IN_CLUSTER | PROXI_AFTER_INC)) st.cleared += 1u;
			st.flags &= ~(unsigned) PROXI_IN_CLUSTER;
		}
		st.last = p;
		st.gap = 0;
		st.flags &= ~(unsigned) (PROXI_AFTER_INC | PROXI_NEED_AFTER);
		if(b < 31) {
			if(mw & (bit >> 1)) st.flags |= PROXI_AFTER_INC;
		} else if(w < 3) {
			if(mnext >> 31) st.flags |= PROXI_AFTER_INC;
		} else st.flags |= PROXI_NEED_AFTER;
	}
	st.gap += c4 - base;
}

/* after the last word: mismatches and included positions under the proximity mask */
CCG_HD void proxi_pair_finish(const ProxiPairState &st, unsigned *mism, unsigned *ninc) {
	unsigned cleared = st.cleared;
	if((st.flags & (PROXI_IN_CLUSTER | PROXI_AFTER_INC)) == (PROXI_IN_CLUSTER | PROXI_AFTER_INC)) cleared += 1u;
	*mism = st.clusters;
	*ninc = st.total - cleared;
}

/* -V with -P: the pair's mask itself (maskProxi's `include`, which fsacmpairint then walks, fsacmpthrd.c:410-414).
 * One word of the pair per call, words ascending; d = SNP bits of word w under the unmasked pair mask, `last` = position
 * of the last SNP so far (-1 = none), W = words of the alignment.  For two neighbouring SNPs last < p at most proxi apart
 * the positions last + 1 .. p + 1 go: sink.clear(word, bits) is asked for words <= w + 1 (p + 1 may be the first base of
 * the next word), so the caller has words 0 .. w + 1 of the mask in place before the call. */
template <class Sink>
CCG_HD void proxi_pair_mask_word(long long &last, long long w, uint32_t d, long long W, unsigned proxi, Sink &sink) {
	while(d) {
		const int b = ccg_first_bit(d);
		d &= ~(0x80000000u >> b);
		const long long p = w * 32 + b;
		if(last >= 0 && (unsigned long long) (p - last) <= proxi) {
			const long long lo = last + 1, hi = p + 1;
			for(long long x = lo >> 5; x <= (hi >> 5) && x < W; ++x) {
				uint32_t bits = 0xFFFFFFFFu;
				if(x == (lo >> 5)) bits &= 0xFFFFFFFFu >> (int) (lo & 31);           /* bases lo % 32 .. 31 */
				if(x == (hi >> 5)) bits &= 0xFFFFFFFFu << (31 - (int) (hi & 31));    /* bases 0 .. hi % 32 */
				sink.clear(x, bits);
			}
		}
		last = p;
	}
}

/* ------------------------------------------------------------------------------------------
 * one row against an existing matrix (-a with -P)
 * ------------------------------------------------------------------------------------------ */
/* cmpFsaRowThrd (fsacmpthrd.c:543-556): the pair's mask starts as the new sample's own mask, then goes through
 * the PER-SAMPLE builder against the column sample -- getIncPosPtr(includeseq, seq, ref, proxi), not maskProxi --
 * and fsacmpair counts under it.  So an event at most proxi after the previous event removes everything from
 * the previous event to this one, both inclusive; what is wanted here are the counts of what is removed. */
struct ProxiRowState {
	long long last;          /* position of the last event, -1 = none */
	unsigned pend_inc, pend_snp;   /* included positions / SNPs strictly after `last` seen so far */
	unsigned last_inc, last_snp;   /* `last` itself is included / a SNP and not yet removed */
	unsigned cleared_inc, cleared_snp, total_inc, total_snp;
};

CCG_HD void proxi_row_init(ProxiRowState &st) {
	st.last = -1;
	st.pend_inc = st.pend_snp = st.last_inc = st.last_snp = 0;
	st.cleared_inc = st.cleared_snp = st.total_inc = st.total_snp = 0;
}

/* ev = event bits of the word (proxi_events), m = included positions before the pair's proximity masking,
 * d = SNPs among them; words ascending, p0 = position of the word's first base */
CCG_HD void proxi_row_word(ProxiRowState &st, long long p0, uint32_t ev, uint32_t m, uint32_t d, unsigned proxi) {
	st.total_inc += (unsigned) ccg_popc32(m);
	st.total_snp += (unsigned) ccg_popc32(d);
	uint32_t rest = 0xFFFFFFFFu;                 /* bases of this word not yet looked at */
	while(ev) {
		const int b = ccg_first_bit(ev);
		const uint32_t bit = 0x80000000u >> b;
		ev &= ~bit;
		const uint32_t before = rest & ~(bit | (bit - 1u));
		st.pend_inc += (unsigned) ccg_popc32(m & before);
		st.pend_snp += (unsigned) ccg_popc32(d & before);
		rest = bit - 1u;
		const long long p = p0 + b;
		const unsigned inc_p = (m & bit) ? 1u : 0u, snp_p = (d & bit) ? 1u : 0u;
		if(st.last >= 0 && (unsigned long long) (p - st.last) <= proxi) {
			st.cleared_inc += st.pend_inc + st.last_inc + inc_p;
			st.cleared_snp += st.pend_snp + st.last_snp + snp_p;
			st.last_inc = st.last_snp = 0;
		} else {
			st.last_inc = inc_p;
			st.last_snp = snp_p;
		}
		st.pend_inc = st.pend_snp = 0;
		st.last = p;
	}
	st.pend_inc += (unsigned) ccg_popc32(m & rest);
	st.pend_snp += (unsigned) ccg_popc32(d & rest);
}

CCG_HD void proxi_row_finish(const ProxiRowState &st, unsigned *mism, unsigned *ninc) {
	*mism = st.total_snp - st.cleared_snp;
	*ninc = st.total_inc - st.cleared_inc;
}

/* ------------------------------------------------------------------------------------------
 * per sample
 * ------------------------------------------------------------------------------------------ */
/* Event bits of one word.  vs_ref == 0: the sample against itself (pair mode, cdist.c:91): its
 * unknown positions (getIncPos) or nothing (the Insig builders).  vs_ref != 0: against the
 * shared-mask reference sample (cdist.c:111).  ms / mr = known masks, (hs, ls) / (hr, lr) = code
 * bit planes with unknown positions cleared, valid = bits of positions < len. */
CCG_HD uint32_t proxi_events(int vs_ref, int snp_only, uint32_t ms, uint32_t hs, uint32_t ls, uint32_t mr, uint32_t hr,
                             uint32_t lr, uint32_t valid) {
	if(!vs_ref) return snp_only ? 0u : (~ms & valid);
	const uint32_t both = ms & mr, diff = (hs ^ hr) | (ls ^ lr);
	return snp_only ? (both & diff) : ((~both | diff) & valid);
}

CCG_HD uint32_t proxi_valid_bits(long long len, long long p0) {
	const long long left = len - p0;
	if(left >= 32) return 0xFFFFFFFFu;
	if(left <= 0) return 0u;
	return 0xFFFFFFFFu << (32 - (int) left);
}

/* Words [w_begin, w_end) of one sample; `last` = position of the last event before word w_begin
 * (-1 = none within reach); events(w) returns the event bits of word w; sink.clear(w, bits)
 * removes the given positions of word w (it may be asked for words before w_begin: a range
 * reaches back to the previous event). */
template <class Events, class Sink>
CCG_HD void proxi_scan_words(long long last, long long w_begin, long long w_end, unsigned proxi, Events &events, Sink &sink) {
	for(long long w = w_begin; w < w_end; ++w) {
		uint32_t ev = events(w);
		if(!ev) continue;
		uint32_t pend = 0;
		while(ev) {
			const int b = ccg_first_bit(ev);
			const uint32_t bit = 0x80000000u >> b;
			ev &= ~bit;
			const long long p = w * 32 + b;
			if(last >= 0 && (unsigned long long) (p - last) <= proxi) {
				const long long wl = last >> 5;
				const uint32_t upto_b = ~(bit - 1u);                 /* bases 0 .. b */
				const uint32_t from_last = 0xFFFFFFFFu >> (int) (last & 31);   /* bases last%32 .. 31 */
				if(wl == w) pend |= upto_b & from_last;
				else {
					sink.clear(wl, from_last);
					for(long long x = wl + 1; x < w; ++x) sink.clear(x, 0xFFFFFFFFu);
					pend |= upto_b;
				}
			}
			last = p;
		}
		if(pend) sink.clear(w, pend);
	}
}

/* position of the last event in words [w_lo, w_hi), scanning downwards; -1 if none */
template <class Events>
CCG_HD long long proxi_last_event_before(long long w_lo, long long w_hi, Events &events) {
	for(long long w = w_hi - 1; w >= w_lo; --w) {
		const uint32_t ev = events(w);
		if(ev) {
#ifdef __CUDA_ARCH__
			const int low = __ffs((int) ev) - 1;
#else
			const int low = __builtin_ctz(ev);
#endif
			return w * 32 + (31 - low);
		}
	}
	return -1;
}

#endif
