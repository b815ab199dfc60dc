/* Host-side check of the Phylip writer: the multi-threaded writer must print byte for byte what a
 * plain per-cell fprintf loop in the reference's format (phy.c:59-123) prints, for every cell type. */
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

int main(int argc, char **argv) {
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
