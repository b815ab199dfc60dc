/*
 * fsa_reader.c -- see fsa_reader.h.  Written against the observable behaviour of the
 * reference parsers (seqparse.c / filebuff.c), not their code: zlib's gz* layer gives the
 * same transparent handling of plain, gzip and concatenated-gzip inputs.
 */
#define _POSIX_C_SOURCE 200809L
#include "fsa_reader.h"

#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define CHUNK (1 << 20)

struct FsaReader {
	gzFile gz;
	unsigned char *buf;
	size_t pos, avail;
	long long base;          /* stream offset of buf[0] */
	int eof;
};

void bytebuf_init(ByteBuf *b, size_t cap) {
	b->cap = cap ? cap : 64;
	b->len = 0;
	b->data = malloc(b->cap);
	if(!b->data) {
		fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
		exit(errno ? errno : 1);
	}
}

void bytebuf_free(ByteBuf *b) {
	free(b->data);
	b->data = 0;
	b->len = b->cap = 0;
}

static inline void bytebuf_push(ByteBuf *b, unsigned char c) {
	if(b->len == b->cap) {
		b->cap <<= 1;
		b->data = realloc(b->data, b->cap);
		if(!b->data) {
			fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
			exit(errno ? errno : 1);
		}
	}
	b->data[b->len++] = c;
}

void fsa_code_table(unsigned flag, unsigned char table[256]) {
	const char *unknown = "N-RYSWKMBDHVXryswkmbdhvxn";
	const char *p;
	int lower = (flag & 8) != 0;

	memset(table, 32, 256);
	table['A'] = 0; table['C'] = 1; table['G'] = 2; table['T'] = 3; table['U'] = 3;
	for(p = unknown; *p; ++p) table[(unsigned char) *p] = 4;
	table['a'] = lower ? 0 : 4;
	table['c'] = lower ? 1 : 4;
	table['g'] = lower ? 2 : 4;
	table['t'] = lower ? 3 : 4;
	table['u'] = lower ? 3 : 4;
}

FsaReader *fsa_open(const char *path) {
	FsaReader *r = calloc(1, sizeof(*r));
	if(!r) return 0;
	r->gz = (path[0] == '-' && path[1] == 0) ? gzdopen(0, "rb") : gzopen(path, "rb");
	if(!r->gz) {
		free(r);
		return 0;
	}
	gzbuffer(r->gz, CHUNK);
	r->buf = malloc(CHUNK);
	if(!r->buf) {
		gzclose(r->gz);
		free(r);
		return 0;
	}
	return r;
}

void fsa_close(FsaReader *r) {
	if(!r) return;
	gzclose(r->gz);
	free(r->buf);
	free(r);
}

static int refill(FsaReader *r) {
	int got;
	if(r->eof) return 0;
	r->base += (long long) r->avail;
	got = gzread(r->gz, r->buf, CHUNK);
	if(got <= 0) {
		r->eof = 1;
		r->avail = r->pos = 0;
		return 0;
	}
	r->pos = 0;
	r->avail = (size_t) got;
	return 1;
}

long long fsa_tell(const FsaReader *r) {
	return r->base + (long long) r->pos;
}

int fsa_is_plain(FsaReader *r) {
	/* gzdirect is only meaningful once the header has been looked at */
	if(r->pos == r->avail && !r->eof) refill(r);
	return gzdirect(r->gz) ? 1 : 0;
}

int fsa_seek(FsaReader *r, long long offset) {
	if(gzseek(r->gz, (z_off_t) offset, SEEK_SET) < 0) return -1;
	r->base = offset;
	r->pos = r->avail = 0;
	r->eof = 0;
	return 0;
}

int fsa_peek(FsaReader *r) {
	if(r->pos == r->avail && !refill(r)) return -1;
	return r->buf[r->pos];
}

int fsa_next_header(FsaReader *r, ByteBuf *header) {
	return fsa_next_header_off(r, header, 0);
}

int fsa_next_header_off(FsaReader *r, ByteBuf *header, long long *offset) {
	header->len = 0;
	/* find the next '>' */
	for(;;) {
		unsigned char *hit;
		if(r->pos == r->avail && !refill(r)) return 0;
		hit = memchr(r->buf + r->pos, '>', r->avail - r->pos);
		if(hit) {
			if(offset) *offset = r->base + (long long) (hit - r->buf);
			r->pos = (size_t) (hit - r->buf) + 1;
			break;
		}
		r->pos = r->avail;
	}
	/* the header line */
	for(;;) {
		unsigned char c;
		if(r->pos == r->avail && !refill(r)) return 0;     /* no newline before end of file: no entry */
		c = r->buf[r->pos++];
		if(c == '\n') break;
		bytebuf_push(header, c);
	}
	while(header->len && isspace(header->data[header->len - 1])) --header->len;
	bytebuf_push(header, 0);
	--header->len;
	return 1;
}

int fsa_read_codes(FsaReader *r, const unsigned char table[256], ByteBuf *codes) {
	codes->len = 0;
	if(r->pos == r->avail && !refill(r)) return 0;
	for(;;) {
		const unsigned char *p = r->buf + r->pos, *end = r->buf + r->avail;
		/* make room for the whole chunk once, then translate without bounds checks */
		if(codes->cap - codes->len < (size_t) (end - p)) {
			while(codes->cap - codes->len < (size_t) (end - p)) codes->cap <<= 1;
			codes->data = realloc(codes->data, codes->cap);
			if(!codes->data) {
				fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
				exit(errno ? errno : 1);
			}
		}
		unsigned char *out = codes->data + codes->len;
		while(p < end && *p != '>') {
			unsigned char c = table[*p++];
			*out = c;
			out += c < 32;
		}
		codes->len = (size_t) (out - codes->data);
		r->pos = (size_t) (p - r->buf);
		if(p < end) return 1;                /* stopped at the next record */
		if(!refill(r)) return 1;             /* end of file ends the record */
	}
}

const unsigned char *fsa_window(FsaReader *r, size_t *have) {
	if(r->pos == r->avail && !refill(r)) {
		*have = 0;
		return 0;
	}
	*have = r->avail - r->pos;
	return r->buf + r->pos;
}

void fsa_consume(FsaReader *r, size_t n) {
	r->pos += n;
}

int fsa_read_line(FsaReader *r, ByteBuf *line) {
	line->len = 0;
	if(r->pos == r->avail && !refill(r)) return 0;
	for(;;) {
		const unsigned char *p = r->buf + r->pos;
		const size_t have = r->avail - r->pos;
		const unsigned char *nl = memchr(p, '\n', have);
		const size_t take = nl ? (size_t) (nl - p) : have;
		if(line->cap - line->len < take + 1) {
			while(line->cap - line->len < take + 1) line->cap <<= 1;
			line->data = realloc(line->data, line->cap);
			if(!line->data) {
				fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
				exit(errno ? errno : 1);
			}
		}
		memcpy(line->data + line->len, p, take);
		line->len += take;
		r->pos += take + (nl ? 1 : 0);
		if(nl || !refill(r)) break;
	}
	line->data[line->len] = 0;
	return 1;
}
