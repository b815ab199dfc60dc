/* readers_test.c -- what the driver's input layer (ccphylo_b200/host/fsa_reader.c, mat_reader.c) makes of a file,
 * as text, for tests/test_host_c.py to compare with the oracle's translate (pinned to the reference's table) and
 * with a plain Python parse of the .mat text.
 *
 *   readers_test fsa <flag> <file>            every record: ">header" line, then the codes as digits
 *   readers_test mat <minDepth> <file> <template>   "len nNucs" line, then one "A C G T - N total" line per row */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fsa_reader.h"
#include "mat_reader.h"

int main(int argc, char **argv) {
	if(argc >= 4 && strcmp(argv[1], "fsa") == 0) {
		unsigned char table[256];
		fsa_code_table((unsigned) atoi(argv[2]), table);
		FsaReader *r = fsa_open(argv[3]);
		if(!r) return 1;
		printf("first=%d plain=%d\n", fsa_peek(r), fsa_is_plain(r));
		ByteBuf header, codes;
		bytebuf_init(&header, 16);
		bytebuf_init(&codes, 16);
		long long off = 0;
		while(fsa_next_header_off(r, &header, &off)) {
			printf(">%s @%lld\n", (const char *) header.data, off);
			if(!fsa_read_codes(r, table, &codes)) {
				printf("(no sequence)\n");
				break;
			}
			for(size_t k = 0; k < codes.len; ++k) putchar('0' + codes.data[k]);
			putchar('\n');
		}
		fsa_close(r);
		return 0;
	}
	if(argc >= 5 && strcmp(argv[1], "mat") == 0) {
		MatSample m;
		mat_sample_init(&m);
		int st = mat_load_template(argv[3], argv[4], (unsigned) atoi(argv[2]), &m);
		printf("status=%d\n", st);
		if(st == 1) {
			printf("%zu %u\n", m.len, m.nNucs);
			for(size_t p = 0; p < m.len; ++p)
				printf("%u %u %u %u %u %u %u\n", m.counts[6 * p], m.counts[6 * p + 1], m.counts[6 * p + 2], m.counts[6 * p + 3],
				       m.counts[6 * p + 4], m.counts[6 * p + 5], m.totals[p]);
		}
		mat_sample_free(&m);
		return 0;
	}
	return 2;
}
