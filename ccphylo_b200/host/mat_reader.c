/* mat_reader.c -- see mat_reader.h */
#define _POSIX_C_SOURCE 200809L
#include "mat_reader.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fsa_reader.h"

void mat_sample_init(MatSample *m) {
	memset(m, 0, sizeof(*m));
}

void mat_sample_free(MatSample *m) {
	free(m->counts);
	free(m->totals);
	memset(m, 0, sizeof(*m));
}

static void grow(MatSample *m) {
	m->cap = m->cap ? m->cap << 1 : (size_t) 1 << 16;
	m->counts = realloc(m->counts, m->cap * 6 * sizeof(uint16_t));
	m->totals = realloc(m->totals, m->cap * sizeof(uint32_t));
	if(!m->counts || !m->totals) {
		fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
		exit(errno ? errno : 1);
	}
}

int mat_peek(const char *path) {
	FsaReader *r = fsa_open(path);
	if(!r) return -1;
	int c = fsa_peek(r);
	fsa_close(r);
	return c;
}

int mat_load_template(const char *path, const char *target, unsigned minDepth, MatSample *out) {
	FsaReader *r = fsa_open(path);
	ByteBuf line;
	int found = 0;
	out->len = 0;
	out->nNucs = 0;
	if(!r) return -1;
	if(fsa_peek(r) < 0) {
		fsa_close(r);
		errno = 0;
		return -1;
	}
	bytebuf_init(&line, 256);
	while(fsa_read_line(r, &line)) {
		if(!found) {
			if(line.data[0] == '#' && strcmp((const char *) line.data + 1, target) == 0) found = 1;
			continue;
		}
		if(line.len == 0 || line.data[0] == '#') break;          /* end of the template */
		/* row: ref, then tab-separated numbers in the file order A C G T N - */
		const unsigned char *p = line.data + 1;
		unsigned v[6] = {0, 0, 0, 0, 0, 0};
		int f = -1;
		for(; *p; ++p) {
			if(*p == '\t') { if(++f > 5) break; }
			else if(f >= 0) v[f] = 10 * v[f] + (unsigned) (*p - '0');
		}
		if(line.data[0] == '-') continue;                        /* insertion relative to the template */
		if(out->len == out->cap) grow(out);
		uint16_t *c = out->counts + out->len * 6;
		c[0] = (uint16_t) v[0]; c[1] = (uint16_t) v[1]; c[2] = (uint16_t) v[2]; c[3] = (uint16_t) v[3];
		c[4] = (uint16_t) v[5];                                  /* '-' is stored before N */
		c[5] = (uint16_t) v[4];
		const uint32_t tot = v[0] + v[1] + v[2] + v[3] + v[4] + v[5];
		out->totals[out->len] = tot;
		out->nNucs += minDepth <= tot;
		++out->len;
	}
	bytebuf_free(&line);
	fsa_close(r);
	return found;
}
