/* Host-side check of the Phylip writer: the multi-threaded writer must print byte for byte what a
 * plain per-cell fprintf loop in the reference's format (phy.c:59-123) prints, for every cell type. */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "phy_writer.h"

static char *slurp(const char *path, size_t *len) {
	FILE *f = fopen(path, "rb");
	fseek(f, 0, SEEK_END);
	*len = (size_t) ftell(f);
	fseek(f, 0, SEEK_SET);
	char *d = malloc(*len + 1);
	if(fread(d, 1, *len, f) != *len) exit(2);
	fclose(f);
	return d;
}

/* phy_format_fixed against snprintf("%.*f"): ratios as the epilogue produces them, dyadic ties (exactly half a unit of
 * the last printed digit), values next to a tie, tiny / huge / negative values, raw bit patterns; every precision 0 .. 18 */
static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd(void) {
	rng_state ^= rng_state << 13;
	rng_state ^= rng_state >> 7;
	rng_state ^= rng_state << 17;
	return rng_state;
}

static int check_fixed(double d, int precision, long *fallbacks) {
	char a[512], b[64];
	const int la = snprintf(a, sizeof(a), "%.*f", precision, d);
	const size_t lb = phy_format_fixed(b, d, precision);
	if(lb == 0) { ++*fallbacks; return 0; }
	if((size_t) la != lb || memcmp(a, b, lb)) {
		b[lb] = 0;
		fprintf(stderr, "MISMATCH %.17g precision %d: snprintf '%s' phy_format_fixed '%s'\n", d, precision, a, b);
		return 1;
	}
	return 0;
}

static int fixed_sweep(long rounds) {
	long bad = 0, fallbacks = 0, total = 0;
	for(long r = 0; r < rounds && bad < 10; ++r) {
		const int p = (int) (rnd() % 19);
		double v[12];
		const uint64_t x = rnd(), y = rnd();
		v[0] = (double) (x % 5000000) * 1000000.0 / (double) (1 + y % 5000000);       /* mismatches x norm / included */
		v[1] = (double) (float) v[0];                                                  /* a float cell */
		v[2] = (double) (x % 100000) / (double) (1 + y % 1000);
		v[3] = ldexp((double) (2 * (x % 1000000) + 1), -1 - (int) (y % 40));           /* odd / 2^k: ties at some precision */
		v[4] = nextafter(v[3], 0.0);
		v[5] = nextafter(v[3], 1e300);
		v[6] = -v[0];
		v[7] = ldexp((double) (x >> 11), -(int) (y % 1100));                           /* down to subnormals */
		v[8] = (double) (x % 1000) + 0.5;                                              /* p = 0 ties */
		v[9] = (double) (x >> 20) / 1024.0;                                            /* up to 2^34, exact binary fractions */
		memcpy(&v[10], &x, 8);                                                         /* any bit pattern */
		v[11] = ((double) (x % 2000001) - 1000000.0) / 1e9;                            /* around zero: "-0.000000000" */
		for(int k = 0; k < 12; ++k, ++total) bad += check_fixed(v[k], p, &fallbacks);
	}
	/* a few fixed points */
	const double fixed[] = {0.5, 1.5, 2.5, 0.125, 0.375, 1e-10, -1e-10, 0.9999999995, 0.99999999949999, 999999.9999999995, 4503599627370495.5,
	                        4503599627370496.0, 1e18, 9.223372036854e18, 5e-324, 1.7976931348623157e308, -0.0, 0.0};
	for(size_t k = 0; k < sizeof(fixed) / sizeof(fixed[0]); ++k)
		for(int p = 0; p <= 18; ++p, ++total) bad += check_fixed(fixed[k], p, &fallbacks);
	/* out of range: left to snprintf */
	char b[64];
	if(phy_format_fixed(b, 1.5, 19) || phy_format_fixed(b, 1.5, -1) || phy_format_fixed(b, NAN, 9) || phy_format_fixed(b, INFINITY, 9)) ++bad;
	printf("%s %ld values, %ld left to snprintf\n", bad ? "FAIL" : "OK", total, fallbacks);
	return bad != 0 || fallbacks * 4 > total;
}

int main(int argc, char **argv) {
	if(argc > 2 && strcmp(argv[1], "fixed") == 0) return fixed_sweep(atol(argv[2]));
	const int n = argc > 1 ? atoi(argv[1]) : 700, dn = n - 3;
	const char *dir = argc > 2 ? argv[2] : "/tmp";
	char **names = malloc((size_t) n * sizeof(char *));
	unsigned char *include = malloc((size_t) n);
	for(int i = 0; i < n; ++i) {
		names[i] = malloc(64);
		snprintf(names[i], 64, i % 5 == 0 ? "\"dir/sub/sample_%d.fsa\"" : "some/dir/sample_number_%d.fsa", i);
		include[i] = !(i == 1 || i == 40 || i == n - 2);
	}
	size_t cells = (size_t) dn * (dn - 1) / 2;
	int rc = 0;
	for(int elem = 8; elem >= 1; elem >>= 1) {
		void *buf = malloc(cells * 8);
		uint32_t x = 12345u + (unsigned) elem;
		for(size_t k = 0; k < cells; ++k) {
			x = x * 1664525u + 1013904223u;
			double v = (x >> 8) % 7 == 0 ? -1.0 : (x >> 8) % 3 == 0 ? (double) ((x >> 10) % 100000) : (double) (x >> 6) / 977.0;
			if(elem == 8) ((double *) buf)[k] = v;
			else if(elem == 4) ((float *) buf)[k] = (float) v;
			else if(elem == 2) ((uint16_t *) buf)[k] = (uint16_t) (x >> 9);
			else ((uint8_t *) buf)[k] = (uint8_t) (x >> 13);
		}
		const double scale = elem == 2 ? 100.0 : elem == 1 ? 0.25 : 1.0;
		for(unsigned flags = 0; flags <= 5; flags += 5) {
			char pa[512], pb[512];
			snprintf(pa, sizeof(pa), "%s/phy_ref_%d_%u.txt", dir, elem, flags);
			snprintf(pb, sizeof(pb), "%s/phy_mt_%d_%u.txt", dir, elem, flags);
			/* reference-format loop */
			FILE *f = fopen(pa, "wb");
			if(flags & 4) fprintf(f, "#%s\n", "tmpl");
			fprintf(f, "%10d\n", dn);
			size_t k = 0;
			for(int i = 0, r = 0; r < dn; ++i) {
				if(!include[i]) continue;
				char tmp[64];
				strcpy(tmp, names[i]);
				char *nm = tmp;
				size_t l = strlen(nm);
				if(nm[0] == '"' && nm[l - 1] == '"') { nm[l - 1] = 0; ++nm; }
				char *sl = strrchr(nm, '/');
				if(sl) nm = sl + 1;
				if(flags & 1) fprintf(f, "%s", nm); else fprintf(f, "%-10.10s", nm);
				for(int c = 0; c < r; ++c, ++k) {
					double d = elem == 8 ? ((double *) buf)[k] : elem == 4 ? ((float *) buf)[k] :
					           elem == 2 ? ((uint16_t *) buf)[k] / scale : ((uint8_t *) buf)[k] / scale;
					if(d == (int) d) fprintf(f, "\t%d", (int) d); else fprintf(f, "\t%.*f", 9, d);
				}
				fprintf(f, "\n");
				++r;
			}
			fclose(f);
			for(int threads = 1; threads <= 7; threads += 6) {
				f = fopen(pb, "wb");
				phy_write_mt(f, buf, elem, scale, dn, names, include, "tmpl", flags, 9, threads);
				fclose(f);
				size_t la, lb;
				char *a = slurp(pa, &la), *b = slurp(pb, &lb);
				if(la != lb || memcmp(a, b, la)) {
					fprintf(stderr, "MISMATCH elem=%d flags=%u threads=%d (%zu vs %zu bytes)\n", elem, flags, threads, la, lb);
					rc = 1;
				}
				free(a);
				free(b);
			}
		}
		free(buf);
	}
	printf(rc ? "FAIL\n" : "OK\n");
	return rc;
}
