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

/* row: the reference base, then tab-separated numbers in the file order A C G T N - (a NUL byte ends the row as it
 * ended the reference's C string) */
static inline void store_row(MatSample *out, unsigned char ref, const unsigned v[6], unsigned minDepth);

static inline void add_row(MatSample *out, const unsigned char *row, const unsigned char *end, unsigned minDepth) {
	unsigned v[6] = {0, 0, 0, 0, 0, 0};
	int f = -1;
	for(const unsigned char *p = row + 1; p < end && *p; ++p) {
		if(*p == '\t') { if(++f > 5) break; }
		else if(f >= 0) v[f] = 10 * v[f] + (unsigned) (*p - '0');
	}
	store_row(out, row[0], v, minDepth);
}

/* the row as KMA writes it -- one byte, then six times a tab and 1 .. 9 digits, then the newline -- parsed in one pass
 * without looking for the newline first; returns the start of the next row, or NULL for anything else (the caller then
 * takes add_row, which accepts what the reference's parser accepts).  Needs FAST_ROW_MAX readable bytes at row. */
#define FAST_ROW_MAX 72
static inline const unsigned char *fast_row(const unsigned char *row, unsigned v[6]) {
	const unsigned char *q = row + 1;
	for(int f = 0; f < 6; ++f) {
		if(*q != '\t') return 0;
		unsigned d = (unsigned) q[1] - '0';
		if(d > 9) return 0;
		unsigned x = d;
		int nd = 1;
		for(q += 2; (d = (unsigned) *q - '0') <= 9; ++q) {
			if(++nd > 9) return 0;
			x = 10 * x + d;
		}
		v[f] = x;
	}
	return *q == '\n' ? q + 1 : 0;
}

static inline void store_row(MatSample *out, unsigned char ref, const unsigned v[6], unsigned minDepth) {
	if(ref == '-') return;                                    /* insertion relative to the template */
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
	while(!found && fsa_read_line(r, &line))
		if(line.data[0] == '#' && strcmp((const char *) line.data + 1, target) == 0) found = 1;
	/* the template's rows, scanned in place in the reader's buffer (a row is ~15 bytes: copying each into a line buffer
	 * first cost more than parsing it); only the row the end of the window cuts goes through fsa_read_line */
	for(int more = found; more;) {
		size_t have;
		const unsigned char *base = fsa_window(r, &have);
		if(!have) break;
		const unsigned char *p = base, *end = base + have;
		while(p < end) {
			if(end - p >= FAST_ROW_MAX && *p != '\n' && *p != '#') {
				unsigned v[6];
				const unsigned char *next = fast_row(p, v);
				if(next) {
					store_row(out, *p, v, minDepth);
					p = next;
					continue;
				}
			}
			const unsigned char *nl = memchr(p, '\n', (size_t) (end - p));
			if(!nl) break;
			if(nl == p || *p == '#') { more = 0; break; }        /* end of the template */
			add_row(out, p, nl, minDepth);
			p = nl + 1;
		}
		fsa_consume(r, (size_t) (p - base));
		if(more && p < end) {
			if(!fsa_read_line(r, &line) || line.len == 0 || line.data[0] == '#') break;
			add_row(out, line.data, line.data + line.len, minDepth);
		}
	}
	bytebuf_free(&line);
	fsa_close(r);
	return found;
}
