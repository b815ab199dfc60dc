/*
 * proxi_core_test.cpp -- the -P arithmetic the CUDA kernels run (ccphylo_b200/csrc/proxi_core.h), compiled for
 * the host and driven word by word exactly as k_pairdist_proxi / k_sample_proxi drive it, against the oracle
 * (liboracle.so: orc_pair_counts_proxi, orc_inc_pos, themselves pinned to the reference's maskProxi / getIncPos*
 * in tests/test_oracle_vs_reference.py).  Prints "OK <cases>" or the first difference.
 *
 *   proxi_core_test <seed>
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fsa_oracle.h"
#include "proxi_core.h"

static uint64_t rng_state;
static uint32_t rnd() {
	rng_state ^= rng_state << 13;
	rng_state ^= rng_state >> 7;
	rng_state ^= rng_state << 17;
	return (uint32_t) (rng_state >> 16);
}
static double urnd() { return (rnd() & 0xFFFFFF) / (double) 0x1000000; }

/* a sample derived from `base`: SNPs and unknowns in clusters, so that every proxi meets close and distant events */
static void make_sample(const std::vector<unsigned char> &base, std::vector<unsigned char> &out, double snp, double unk) {
	const int len = (int) base.size();
	out = base;
	for(int p = 0; p < len; ++p) {
		if(urnd() < snp) {
			int burst = 1 + (int) (rnd() % 3);
			for(int k = 0; k < burst && p < len; ++k, p += 1 + (int) (rnd() % 6))
				out[p] = (unsigned char) ((base[p] + 1 + rnd() % 3) & 3);
		} else if(urnd() < unk) {
			int run = 1 + (int) (rnd() % 40);
			for(int k = 0; k < run && p < len; ++k, ++p) out[p] = 4;
		}
	}
}

/* the bit planes of the device store (k_encode.cu): code bits and known mask per 32-base word, base k <-> bit 31 - k,
 * code planes cleared where the mask is */
struct Planes {
	std::vector<uint32_t> h, l, m;
};
static void to_planes(const uint64_t *seq, const uint32_t *mask, int W, Planes &pl) {
	pl.h.assign((size_t) W, 0);
	pl.l.assign((size_t) W, 0);
	pl.m.assign(mask, mask + W);
	for(int w = 0; w < W; ++w) {
		uint32_t h = 0, l = 0;
		for(int b = 0; b < 32; ++b) {
			const unsigned c = (unsigned) (seq[w] >> (62 - 2 * b)) & 3;
			h |= (uint32_t) (c >> 1) << (31 - b);
			l |= (uint32_t) (c & 1) << (31 - b);
		}
		pl.h[w] = h & mask[w];
		pl.l[w] = l & mask[w];
	}
}

struct HostEvents {
	const Planes *s, *r;
	int vs_ref, snp_only;
	long long len;
	uint32_t operator()(long long w) const {
		const uint32_t valid = proxi_valid_bits(len, w * 32);
		if(!vs_ref) return proxi_events(0, snp_only, s->m[w], 0, 0, 0, 0, 0, valid);
		return proxi_events(1, snp_only, s->m[w], s->h[w], s->l[w], r->m[w], r->h[w], r->l[w], valid);
	}
};
struct HostSink {
	uint32_t *mask;
	void clear(long long w, uint32_t bits) { mask[w] &= ~bits; }
};

static long cases = 0;

static int check_pairs(int n, int len, unsigned proxi, unsigned variant_for_masks) {
	const int W = orc_words(len);
	std::vector<unsigned char> base((size_t) len);
	for(int p = 0; p < len; ++p) base[p] = (unsigned char) (rnd() & 3);
	std::vector<std::vector<unsigned char>> codes((size_t) n);
	std::vector<std::vector<uint64_t>> seq((size_t) n, std::vector<uint64_t>((size_t) W + 1));
	std::vector<std::vector<uint32_t>> mask((size_t) n, std::vector<uint32_t>((size_t) W + 1));
	std::vector<Planes> pl((size_t) n);
	for(int i = 0; i < n; ++i) {
		make_sample(base, codes[i], 0.03, 0.01);
		orc_pack(codes[i].data(), len, seq[i].data());
		orc_known_mask(codes[i].data(), len, mask[i].data());
		/* the per-sample builder of cdist.c:91 first, as the caller of the pair pass has done */
		orc_inc_pos(mask[i].data(), codes[i].data(), codes[i].data(), len, proxi, (int) variant_for_masks);
		to_planes(seq[i].data(), mask[i].data(), W, pl[i]);
	}
	for(int i = 1; i < n; ++i) {
		for(int j = 0; j < i; ++j) {
			uint32_t want_m = 0, want_n = 0;
			orc_pair_counts_proxi(seq[i].data(), seq[j].data(), mask[i].data(), mask[j].data(), len, proxi, &want_m, &want_n);
			ProxiPairState st;
			proxi_pair_init(st);
			for(int w = 0; w < W; ++w) {
				const uint32_t m = pl[i].m[w] & pl[j].m[w];
				const uint32_t d = ((pl[i].l[w] ^ pl[j].l[w]) | (pl[i].h[w] ^ pl[j].h[w])) & m;
				proxi_pair_word(st, w * 32, d, m, proxi);
			}
			unsigned got_m, got_n;
			proxi_pair_finish(st, &got_m, &got_n);
			/* the chunked form the kernel uses: four words per call, the tail padded with empty words */
			ProxiPairState sc;
			proxi_pair_init(sc);
			for(int w = 0; w < W; w += 4) {
				uint32_t mm[4] = {0, 0, 0, 0}, dd[4] = {0, 0, 0, 0};
				for(int q = 0; q < 4 && w + q < W; ++q) {
					mm[q] = pl[i].m[w + q] & pl[j].m[w + q];
					dd[q] = ((pl[i].l[w + q] ^ pl[j].l[w + q]) | (pl[i].h[w + q] ^ pl[j].h[w + q])) & mm[q];
				}
				proxi_pair_chunk(sc, w * 32, dd[0], dd[1], dd[2], dd[3], mm[0], mm[1], mm[2], mm[3], proxi);
			}
			unsigned chunk_m, chunk_n;
			proxi_pair_finish(sc, &chunk_m, &chunk_n);
			if(chunk_m != got_m || chunk_n != got_n) {
				printf("pair len=%d proxi=%u (%d,%d): chunked %u/%u, word by word %u/%u\n", len, proxi, i, j, chunk_m, chunk_n, got_m, got_n);
				return 1;
			}
			++cases;
			if(got_m != want_m || got_n != want_n) {
				printf("pair len=%d proxi=%u (%d,%d): got %u/%u want %u/%u\n", len, proxi, i, j, got_m, got_n, want_m, want_n);
				return 1;
			}
			/* the mask itself, as k_pair_proxi_mask builds it for -V: word w + 1 is put in place before word w is scanned */
			std::vector<uint32_t> want_mask((size_t) W + 1, 0), got_mask((size_t) W + 1, 0);
			orc_mask_proxi(seq[i].data(), seq[j].data(), mask[i].data(), mask[j].data(), len, proxi, want_mask.data());
			HostSink msink = {got_mask.data()};
			long long last = -1;
			if(W) got_mask[0] = pl[i].m[0] & pl[j].m[0];
			for(int w = 0; w < W; ++w) {
				if(w + 1 < W) got_mask[w + 1] = pl[i].m[w + 1] & pl[j].m[w + 1];
				const uint32_t m = pl[i].m[w] & pl[j].m[w];
				const uint32_t d = ((pl[i].l[w] ^ pl[j].l[w]) | (pl[i].h[w] ^ pl[j].h[w])) & m;
				proxi_pair_mask_word(last, w, d, W, proxi, msink);
			}
			if(memcmp(want_mask.data(), got_mask.data(), (size_t) W * sizeof(uint32_t)) != 0) {
				int w = 0;
				while(want_mask[w] == got_mask[w]) ++w;
				printf("pair mask len=%d proxi=%u (%d,%d): word %d got %08x want %08x\n", len, proxi, i, j, w, got_mask[w], want_mask[w]);
				return 1;
			}
		}
	}
	return 0;
}

/* the segmented scan of k_sample_proxi: every segment finds the event its first range may start from by looking
 * back proxi positions, then scans its own words; segments run in an arbitrary order */
static int check_samples(int len, unsigned proxi, int variant, int seg_words) {
	const int W = orc_words(len);
	std::vector<unsigned char> base((size_t) len), s0, s1;
	for(int p = 0; p < len; ++p) base[p] = (unsigned char) (rnd() & 3);
	make_sample(base, s0, 0.02, 0.02);
	make_sample(base, s1, 0.02, 0.02);
	std::vector<uint64_t> q0((size_t) W + 1), q1((size_t) W + 1);
	std::vector<uint32_t> k0((size_t) W + 1), k1((size_t) W + 1);
	orc_pack(s0.data(), len, q0.data());
	orc_pack(s1.data(), len, q1.data());
	orc_known_mask(s0.data(), len, k0.data());
	orc_known_mask(s1.data(), len, k1.data());
	Planes p0, p1;
	to_planes(q0.data(), k0.data(), W, p0);
	to_planes(q1.data(), k1.data(), W, p1);
	for(int vs_ref = 0; vs_ref < 2; ++vs_ref) {
		/* oracle: initIncPos, then the builder on (s1, s1) or (s1, s0) */
		std::vector<uint32_t> want((size_t) W + 1, 0), got;
		for(int w = 0; w < W; ++w) want[w] = proxi_valid_bits(len, (long long) w * 32);
		got = want;
		orc_inc_pos(want.data(), s1.data(), vs_ref ? s0.data() : s1.data(), len, proxi, variant);
		/* device scheme: known-ness first (the mask plane / k_build_global_mask), then the proximity ranges */
		for(int w = 0; w < W; ++w) got[w] &= vs_ref ? (k0[w] & k1[w]) : k1[w];
		HostEvents ev = {&p1, &p0, vs_ref, variant, len};
		HostSink sink = {got.data()};
		const int nseg = (W + seg_words - 1) / seg_words;
		for(int k = 0; k < nseg; ++k) {
			const int seg = (k * 7 + 3) % nseg;               /* any order; 7 is coprime to most nseg, repeats are harmless */
			for(int pass = 0; pass < 2; ++pass) {
				const int sg = pass ? k : seg;
				const long long wb = (long long) sg * seg_words;
				long long we = wb + seg_words;
				if(we > W) we = W;
				const long long back = wb * 32 - (long long) proxi;
				const long long w_lo = back <= 0 ? 0 : (back >> 5);
				const long long last = proxi_last_event_before(w_lo, wb, ev);
				proxi_scan_words(last, wb, we, proxi, ev, sink);
			}
		}
		++cases;
		if(memcmp(want.data(), got.data(), (size_t) W * sizeof(uint32_t)) != 0) {
			int w = 0;
			while(want[w] == got[w]) ++w;
			printf("sample len=%d proxi=%u variant=%d vs_ref=%d seg=%d: word %d got %08x want %08x\n", len, proxi, variant,
			       vs_ref, seg_words, w, got[w], want[w]);
			return 1;
		}
	}
	return 0;
}

/* -a with -P: the row state machine against getIncPos on the pair's mask + plain counting */
static int check_rows(int len, unsigned proxi, int variant) {
	const int W = orc_words(len);
	std::vector<unsigned char> base((size_t) len), add, col;
	for(int p = 0; p < len; ++p) base[p] = (unsigned char) (rnd() & 3);
	make_sample(base, add, 0.03, 0.02);
	make_sample(base, col, 0.03, 0.02);
	std::vector<uint64_t> qa((size_t) W + 1), qc((size_t) W + 1);
	std::vector<uint32_t> ka((size_t) W + 1), kc((size_t) W + 1), own((size_t) W + 1);
	orc_pack(add.data(), len, qa.data());
	orc_pack(col.data(), len, qc.data());
	orc_known_mask(add.data(), len, ka.data());
	orc_known_mask(col.data(), len, kc.data());
	/* includeadd: the new sample's own builder (fsacmpthrd.c:627-628) */
	own = ka;
	orc_inc_pos(own.data(), add.data(), add.data(), len, proxi, variant);
	/* oracle: copy, getIncPosPtr(copy, col, add, proxi), fsacmpair */
	std::vector<uint32_t> pm = own;
	orc_inc_pos(pm.data(), col.data(), add.data(), len, proxi, variant);
	std::vector<uint32_t> ones((size_t) W + 1, 0xFFFFFFFFu);
	uint32_t want_m = 0, want_n = 0;
	orc_pair_counts(qa.data(), qc.data(), pm.data(), ones.data(), len, &want_m, &want_n);
	/* device scheme: raw planes of both for the events, the new sample's masked mask for the counts */
	Planes pa, pc;
	to_planes(qa.data(), ka.data(), W, pa);
	to_planes(qc.data(), kc.data(), W, pc);
	ProxiRowState st;
	proxi_row_init(st);
	for(int w = 0; w < W; ++w) {
		const uint32_t valid = proxi_valid_bits(len, (long long) w * 32);
		const uint32_t ev = proxi_events(1, variant, pc.m[w], pc.h[w], pc.l[w], pa.m[w], pa.h[w], pa.l[w], valid);
		const uint32_t m = own[w] & pc.m[w];
		const uint32_t d = ((pa.h[w] ^ pc.h[w]) | (pa.l[w] ^ pc.l[w])) & m;
		proxi_row_word(st, (long long) w * 32, ev, m, d, proxi);
	}
	unsigned got_m, got_n;
	proxi_row_finish(st, &got_m, &got_n);
	++cases;
	if(got_m != want_m || got_n != want_n) {
		printf("row len=%d proxi=%u variant=%d: got %u/%u want %u/%u\n", len, proxi, variant, got_m, got_n, want_m, want_n);
		return 1;
	}
	return 0;
}

int main(int argc, char **argv) {
	rng_state = 0x9E3779B97F4A7C15ull ^ (uint64_t) (argc > 1 ? atoll(argv[1]) : 1) * 0x100000001B3ull;
	static const unsigned proxis[] = {1, 2, 3, 5, 8, 31, 32, 33, 63, 64, 65, 100, 1000, 100000};
	static const int lens[] = {1, 2, 31, 32, 33, 64, 95, 128, 129, 700, 4099};
	for(unsigned pi = 0; pi < sizeof(proxis) / sizeof(*proxis); ++pi) {
		for(unsigned li = 0; li < sizeof(lens) / sizeof(*lens); ++li) {
			for(unsigned variant = 0; variant < 2; ++variant) {
				if(check_pairs(5, lens[li], proxis[pi], variant)) return 1;
				for(int rep = 0; rep < 6; ++rep)
					if(check_rows(lens[li], proxis[pi], (int) variant)) return 1;
				for(int seg = 1; seg <= 64; seg *= 4)
					if(check_samples(lens[li], proxis[pi], (int) variant, seg)) return 1;
			}
		}
	}
	printf("OK %ld\n", cases);
	return 0;
}
