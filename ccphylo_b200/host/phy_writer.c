/* phy_writer.c -- see phy_writer.h
 *
 * The reference walks the cells with one fprintf each (phy.c:100-118): minutes of a single core at
 * n >= 10^4.  Here the rows are formatted in parallel into per-thread buffers, block of rows by block
 * of rows, and written in order; integral cells (counts, -1, unnormalised distances -- the common
 * case) take a hand-rolled integer path, every other cell goes through snprintf("%.*f") so the
 * digits are the C library's, exactly as the reference prints them.
 */
#include "phy_writer.h"

#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const char *base_name(char *name) {
	size_t len = strlen(name);
	/* enclosing quotes are dropped (the closing one is cut off in place, as the reference does) */
	if(len && ((name[0] == '"' && name[len - 1] == '"') || (name[0] == '\'' && name[len - 1] == '\''))) {
		name[len - 1] = 0;
		++name;
	}
	const char *slash = strrchr(name, '/');
	return slash ? slash + 1 : name;
}

static inline double cell_value(const void *cells, int elem_size, double byteScale, size_t k) {
	switch(elem_size) {
		case 8: return ((const double *) cells)[k];
		case 4: return ((const float *) cells)[k];
		case 2: return ((const uint16_t *) cells)[k] / byteScale;
		default: return ((const uint8_t *) cells)[k] / byteScale;
	}
}

typedef struct {
	char *data;
	size_t len, cap;
} Buf;

static inline void buf_room(Buf *b, size_t need) {
	if(b->cap - b->len >= need) return;
	while(b->cap - b->len < need) b->cap = b->cap ? b->cap << 1 : (size_t) 1 << 16;
	b->data = realloc(b->data, b->cap);
	if(!b->data) {
		fprintf(stderr, "Error: out of memory while formatting the matrix\n");
		exit(1);
	}
}

/* "\t%d" */
static inline void put_int(Buf *b, int v) {
	char tmp[12];
	int n = 0;
	unsigned u = v < 0 ? 0u - (unsigned) v : (unsigned) v;
	do { tmp[n++] = (char) ('0' + u % 10); u /= 10; } while(u);
	buf_room(b, 14);
	b->data[b->len++] = '\t';
	if(v < 0) b->data[b->len++] = '-';
	while(n) b->data[b->len++] = tmp[--n];
}

typedef struct {
	const void *cells;
	int elem_size, precision;
	double byteScale;
	const char **names;          /* per matrix row: printable name */
	unsigned flags;
	int row_lo, row_hi;          /* this job's rows */
	Buf out;
} RowJob;

static void format_rows(RowJob *j) {
	j->out.len = 0;
	for(int r = j->row_lo; r < j->row_hi; ++r) {
		const char *name = j->names[r];
		const size_t nl = strlen(name);
		buf_room(&j->out, nl + 16);
		if(j->flags & 1) {
			memcpy(j->out.data + j->out.len, name, nl);
			j->out.len += nl;
		} else {
			j->out.len += (size_t) snprintf(j->out.data + j->out.len, 16, "%-10.10s", name);
		}
		size_t k = (size_t) r * (size_t) (r > 0 ? r - 1 : 0) / 2;
		for(int c = 0; c < r; ++c, ++k) {
			const double d = cell_value(j->cells, j->elem_size, j->byteScale, k);
			/* integral test in the reference's own terms: d == (int) d */
			if(d >= -2147483648.0 && d <= 2147483647.0 && d == (double) (int) d) put_int(&j->out, (int) d);
			else {
				buf_room(&j->out, 400);
				j->out.len += (size_t) snprintf(j->out.data + j->out.len, 400, "\t%.*f", j->precision, d);
			}
		}
		buf_room(&j->out, 1);
		j->out.data[j->out.len++] = '\n';
	}
}

static void *row_worker(void *arg) {
	format_rows((RowJob *) arg);
	return 0;
}

void phy_write_mt(FILE *out, const void *cells, int elem_size, double byteScale, int dn, char **names,
                  const unsigned char *include, const char *comment, unsigned flags, int precision, int threads) {
	if(flags & 4) fprintf(out, "#%s\n", comment ? comment : "(null)");
	fprintf(out, "%10d\n", dn);
	if(dn <= 0) return;
	/* names of the matrix rows, in input order */
	const char **rown = malloc((size_t) dn * sizeof(*rown));
	if(!rown) {
		fprintf(stderr, "Error: out of memory while formatting the matrix\n");
		exit(1);
	}
	for(int i = 0, r = 0; r < dn; ++i) {
		if(include && !include[i]) continue;
		rown[r++] = base_name(names[i]);
	}
	if(threads < 1) threads = 1;
	if(threads > 64) threads = 64;
	if((long long) dn * dn < 200000) threads = 1;          /* small matrices: not worth a thread */
	RowJob *jobs = calloc((size_t) threads, sizeof(RowJob));
	pthread_t *th = calloc((size_t) threads, sizeof(pthread_t));
	if(!jobs || !th) {
		fprintf(stderr, "Error: out of memory while formatting the matrix\n");
		exit(1);
	}
	/* blocks of rows with about equal numbers of cells, `threads` blocks in flight, written in order;
	 * a block holds at most ~4M cells so the buffers stay small */
	int row = 0;
	while(row < dn) {
		int started = 0;
		for(int t = 0; t < threads && row < dn; ++t) {
			long long budget = 4LL << 20, got = 0;
			int hi = row;
			while(hi < dn && (got == 0 || got + hi <= budget)) { got += hi; ++hi; }
			RowJob *j = &jobs[t];
			j->cells = cells; j->elem_size = elem_size; j->precision = precision; j->byteScale = byteScale;
			j->names = rown; j->flags = flags; j->row_lo = row; j->row_hi = hi;
			row = hi;
			if(threads == 1 || pthread_create(&th[t], 0, row_worker, j)) {
				format_rows(j);
				th[t] = 0;
			}
			++started;
		}
		for(int t = 0; t < started; ++t) {
			if(threads > 1 && th[t]) pthread_join(th[t], 0);
			fwrite(jobs[t].out.data, 1, jobs[t].out.len, out);
		}
	}
	for(int t = 0; t < threads; ++t) free(jobs[t].out.data);
	free(jobs);
	free(th);
	free(rown);
}

void phy_write(FILE *out, const void *cells, int elem_size, double byteScale, int dn, char **names,
               const unsigned char *include, const char *comment, unsigned flags, int precision) {
	phy_write_mt(out, cells, elem_size, byteScale, dn, names, include, comment, flags, precision, 1);
}
