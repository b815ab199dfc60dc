/* motifs_test.c -- prints what ccphylo_b200/host/motifs.c makes of a motif file: one line per motif (the file's
 * motifs, each followed by its reverse complement), the position codes as decimal numbers.  tests/test_host_c.py
 * compares that with oracle.parse_motifs, which is pinned to the reference's getMethMotifs + maskMotifs.
 *
 *   motifs_test <file> */
#include <stdio.h>

#include "motifs.h"

int main(int argc, char **argv) {
	MotifList m;
	if(argc < 2) return 2;
	if(motifs_load(argv[1], &m)) return 1;
	const unsigned char *s = m.sets;
	for(int k = 0; k < m.n; s += m.lens[k], ++k) {
		for(int q = 0; q < m.lens[k]; ++q) printf(q ? " %d" : "%d", s[q]);
		printf("\n");
	}
	motifs_free(&m);
	return 0;
}
