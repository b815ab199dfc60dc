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

/* "%.*f" without the C library: a double is M x 2^E with a 53-bit integer M, so M x 10^precision fits 128 bits for
 * precision <= 18 and the decimal digits of the cell are the integer  round(M x 10^p / 2^-E)  -- rounded to nearest,
 * ties to even, on the EXACT value, which is what glibc's printf does (round-to-nearest mode).  50 million cells of a
 * normalised 10,000-sample matrix cost minutes of snprintf on one core; this is ~15x faster and byte-identical
 * (tests/csrc/phy_writer_test.c compares it with snprintf on tens of millions of values, every precision).  Returns the
 * number of bytes written to dst (room for 48), or 0 for what it leaves to snprintf: non-finite values, precision > 18,
 * magnitudes from 2^53 on, results of more than 19 digits. */
size_t phy_format_fixed(char *dst, double d, int precision) {
	static const uint64_t pow10[19] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
	                                   1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
	                                   100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull,
	                                   1000000000000000000ull};
	if(precision < 0 || precision > 18) return 0;
	uint64_t bits;
	memcpy(&bits, &d, 8);
	const int neg = (int) (bits >> 63);
	const int bexp = (int) ((bits >> 52) & 0x7FF);
	uint64_t M = bits & 0xFFFFFFFFFFFFFull;
	if(bexp == 0x7FF) return 0;                               /* inf / nan */
	int E;                                                   /* |d| = M x 2^E */
	if(bexp == 0) E = -1074;                                 /* zero and subnormals */
	else { M |= 1ull << 52; E = bexp - 1075; }
	if(E >= 0) return 0;                                     /* integers from 2^52 on */
	const int sh = -E;                                       /* 1 .. 1074 */
	uint64_t q;
	if(sh >= 114) q = 0;                                     /* M x 10^18 < 2^113 <= half a unit */
	else {
		const unsigned __int128 P = (unsigned __int128) M * pow10[precision];
		const unsigned __int128 whole = P >> sh;
		const unsigned __int128 rem = P - (whole << sh), half = (unsigned __int128) 1 << (sh - 1);
		if(whole >> 63) return 0;                              /* more digits than a uint64 holds with room to round up */
		q = (uint64_t) whole;
		if(rem > half || (rem == half && (q & 1))) ++q;
	}
	const uint64_t ip = q / pow10[precision], fp = q % pow10[precision];
	char tmp[24];
	int n = 0;
	size_t len = 0;
	if(neg) dst[len++] = '-';
	uint64_t u = ip;
	do { tmp[n++] = (char) ('0' + u % 10); u /= 10; } while(u);
	while(n) dst[len++] = tmp[--n];
	if(precision) {
		dst[len++] = '.';
		u = fp;
		for(int k = precision - 1; k >= 0; --k) { dst[len + (size_t) k] = (char) ('0' + u % 10); u /= 10; }
		len += (size_t) precision;
	}
	return len;
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
				j->out.data[j->out.len] = '\t';
				const size_t got = phy_format_fixed(j->out.data + j->out.len + 1, d, j->precision);
				if(got) j->out.len += got + 1;
				else j->out.len += (size_t) snprintf(j->out.data + j->out.len, 400, "\t%.*f", j->precision, d);
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
	/* two sets of jobs: while the rows of one batch are written, the next batch is being formatted */
	RowJob *jobs = calloc((size_t) threads * 2, sizeof(RowJob));
	pthread_t *th = calloc((size_t) threads * 2, sizeof(pthread_t));
	if(!jobs || !th) {
		fprintf(stderr, "Error: out of memory while formatting the matrix\n");
		exit(1);
	}
	/* blocks of rows with about equal numbers of cells, `threads` blocks in flight, written in order;
	 * a block holds at most ~4M cells so the buffers stay small */
	int row = 0, started[2] = {0, 0}, set = 0, pending = -1;
	while(row < dn || pending >= 0) {
		/* start the next batch in the free set ... */
		int now = -1;
		if(row < dn) {
			now = set;
			set ^= 1;
			started[now] = 0;
			for(int t = 0; t < threads && row < dn; ++t) {
				long long budget = 4LL << 20, got = 0;
				int hi = row;
				while(hi < dn && (got == 0 || got + hi <= budget)) { got += hi; ++hi; }
				RowJob *j = &jobs[now * threads + t];
				j->cells = cells; j->elem_size = elem_size; j->precision = precision; j->byteScale = byteScale;
				j->names = rown; j->flags = flags; j->row_lo = row; j->row_hi = hi;
				row = hi;
				if(threads == 1 || pthread_create(&th[now * threads + t], 0, row_worker, j)) {
					format_rows(j);
					th[now * threads + t] = 0;
				}
				++started[now];
			}
		}
		/* ... and meanwhile write the batch before it (its threads were joined when it became `pending`) */
		if(pending >= 0)
			for(int t = 0; t < started[pending]; ++t) fwrite(jobs[pending * threads + t].out.data, 1, jobs[pending * threads + t].out.len, out);
		if(now >= 0)
			for(int t = 0; t < started[now]; ++t)
				if(threads > 1 && th[now * threads + t]) pthread_join(th[now * threads + t], 0);
		pending = now;
	}
	threads *= 2;
	for(int t = 0; t < threads; ++t) free(jobs[t].out.data);
	free(jobs);
	free(th);
	free(rown);
}

void phy_write(FILE *out, const void *cells, int elem_size, double byteScale, int dn, char **names,
               const unsigned char *include, const char *comment, unsigned flags, int precision) {
	phy_write_mt(out, cells, elem_size, byteScale, dn, names, include, comment, flags, precision, 1);
}
