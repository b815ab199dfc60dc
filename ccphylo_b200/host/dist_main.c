/*
 * dist_main.c -- `ccphylo-b200 dist`: the host side of the reference's `ccphylo dist`
 * (dist.c:473 main_dist, dist.c:42 makeMatrix, cdist.c:36 ltdFsaMatrix_get, cdist.c:196
 * ltdMsaMatrix_get) in front of the CUDA library.  Same options, same Phylip / .num text, same
 * stderr lines, so `ccphylo union | ccphylo-b200 dist | ccphylo tree` keeps working.
 *
 * What is different by design (B200-first):
 *   - input files are parsed by -t host threads in parallel; the translated codes of each
 *     included sample go straight to the device (ccg_put_sample_codes), which packs them,
 *     builds the inclusion mask and counts it -- the host keeps O(threads) sequences, not n;
 *   - the O(n^2 L) comparison and the epilogue run on the GPU (ccg_run_pair / ccg_run_global);
 *     there is no CPU fallback: without a usable device the program stops with an error;
 *   - -P (proximity) runs on the device too: the per-sample builder right after the upload
 *     (ccg_sample_proximity), the per-pair maskProxi inside the compare (ccg_run_pair), the
 *     shared-mask variant in ccg_build_global_mask;
 *   - -V (variant listing) comes from the device as well (ccg_list_variants, same labels as the reference);
 *     -a appends one row to an existing matrix (ccg_run_row / ccg_mat_run_row);
 *   - -y masks methylation motifs on the device right after each upload (ccg_mask_motifs);
 *   - -y with -P in shared-mask mode: the motif sites go into the shared mask after the proximity pass
 *     (ccg_build_global_mask); -a ignores -y, as the reference does.
 */
#define _POSIX_C_SOURCE 200809L
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include "ccphylo_gpu.h"
#include "cmdline.h"
#include "dist_opts.h"
#include "fsa_reader.h"
#include "motifs.h"
#include "ordered_pool.h"
#include "phy_update.h"
#include "phy_writer.h"

#define VERSION "0.1.0 (B200 path of ccphylo dist 0.8.5)"

/* CCPHYLO_GPU_STATS=1: phase timings on stderr (off by default: scripts parse the reference's stderr lines) */
#include <time.h>
static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}
static void stat_line(const char *what, double t0) {
	if(getenv("CCPHYLO_GPU_STATS")) fprintf(stderr, "# gpu-stats\t%s\t%.3f s\n", what, now_s() - t0);
}

static void die_errno(void) {
	fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
	exit(errno ? errno : 1);
}

static void die_gpu(ccg_ctx *ctx, int rc) {
	fprintf(stderr, "GPU error: %s (%s)\n", ccg_strerror(rc), ccg_last_error(ctx));
	exit(rc ? rc : 1);
}

static FILE *open_out(const char *name) {
	if(name[0] == '-' && name[1] == 0) return stdout;
	FILE *f = fopen(name, "wb");
	if(!f) {
		fprintf(stderr, "Filename:\t%s\n", name);
		die_errno();
	}
	/* large rows: fewer write calls */
	setvbuf(f, 0, _IOFBF, 1 << 22);
	return f;
}

/* ------------------------------------------------------------------------------------------
 * result matrices: packed lower triangle in pinned host memory
 * ------------------------------------------------------------------------------------------ */
/* -H / --mmap (matrix.c:116-231 ltdMatrixMinit): the matrix lives in an unlinked temporary file that is mapped into
 * memory, for sample sets whose matrices exceed the host's RAM.  The file goes where the reference's tmpF (tmp.c:27-76)
 * puts it: -T dir/ -> dir/.kma-XXXXXX, -T prefix -> prefix.tmp<k>, no -T -> tmpfile(). */
static FILE *tmp_file(const char *location) {
	static int counter = 0;
	if(!location || !*location) return tmpfile();
	const size_t len = strlen(location);
	char *name = malloc(len + 32);
	if(!name) return 0;
	FILE *f = 0;
	if(location[len - 1] == '/') {
		sprintf(name, "%s.kma-XXXXXX", location);
		const int fd = mkstemp(name);
		if(fd >= 0 && (f = fdopen(fd, "wb+"))) unlink(name);
	} else {
		sprintf(name, "%s.tmp%d", location, counter++);
		if((f = fopen(name, "wb+"))) unlink(name);
	}
	free(name);
	return f;
}

void *dist_alloc_cells(const DistOpts *o, size_t n, int elem) {
	const size_t cells = n > 1 ? n * (n - 1) / 2 : 1, bytes = cells * (size_t) elem;
	void *p;
	if(o->mmap_matrix) {
		FILE *f = tmp_file(o->tmpdir);
		if(!f || ftruncate(fileno(f), (off_t) bytes)) die_errno();
		p = mmap(0, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fileno(f), 0);
		if(p == MAP_FAILED) {
			fprintf(stderr, "MMAP failed:\n");
			die_errno();
		}
		fclose(f);                               /* the mapping keeps the (unlinked) file alive */
		return p;
	}
	p = ccg_host_alloc(bytes);
	if(!p) {
		fprintf(stderr, "Error: cannot allocate %zu bytes of pinned host memory\n", bytes);
		exit(1);
	}
	return p;
}

void dist_free_cells(const DistOpts *o, void *p, size_t n, int elem) {
	if(!p) return;
	if(o->mmap_matrix) munmap(p, (n > 1 ? n * (n - 1) / 2 : 1) * (size_t) elem);
	else ccg_host_free(p);
}

/* ------------------------------------------------------------------------------------------
 * multi-file FASTA input: parallel parse, ordered consumption
 * ------------------------------------------------------------------------------------------ */
enum { PARSE_OK = 0, PARSE_NOT_FASTA, PARSE_NO_TEMPLATE, PARSE_NO_SEQ, PARSE_OPEN_FAILED };

typedef struct {
	int status, err;
	ByteBuf codes;
	unsigned known;          /* codes < 4 */
} Parsed;

typedef struct {
	const DistOpts *o;
	unsigned char table[256];
} FsaJob;

static void parse_one(int job, void *state, void *user) {
	const FsaJob *fj = (const FsaJob *) user;
	Parsed *r = (Parsed *) state;
	const char *path = fj->o->filenames[job];
	ByteBuf header;
	r->status = PARSE_NO_TEMPLATE;
	r->known = 0;
	r->codes.len = 0;
	FsaReader *fr = fsa_open(path);
	if(!fr) {
		r->status = PARSE_OPEN_FAILED;
		r->err = errno;
		return;
	}
	if(fsa_peek(fr) != '>') {
		r->status = fsa_peek(fr) < 0 ? PARSE_OPEN_FAILED : PARSE_NOT_FASTA;
		r->err = 0;
		fsa_close(fr);
		return;
	}
	bytebuf_init(&header, 256);
	while(fsa_next_header(fr, &header)) {
		if(strcmp((const char *) header.data, fj->o->targetTemplate) != 0) continue;
		if(!fsa_read_codes(fr, fj->table, &r->codes)) r->status = PARSE_NO_SEQ;
		else {
			unsigned known = 0;
			for(size_t k = 0; k < r->codes.len; ++k) known += r->codes.data[k] < 4;
			r->known = known;
			r->status = PARSE_OK;
		}
		break;
	}
	bytebuf_free(&header);
	fsa_close(fr);
}

/* -y: motifs loaded by make_matrix (dist.c:127-131); n == 0 without -y */
static MotifList g_motifs;
/* shared-mask mode with -y and -P: the motif sites stay out of the samples' own masks (they would read as unknown
 * bases in the proximity pass); ccg_build_global_mask applies them after that pass */
static int g_defer_motifs = 0;

/* one sample into its slot: the device packs it, builds its mask and -- with -y -- takes the methylation sites
 * of every motif match out of it (maskMotifs, cdist.c:90,109,137) */
static void upload_sample(ccg_ctx *ctx, int slot, const ByteBuf *codes, unsigned *inc, int proxi_apply) {
	int rc = ccg_put_sample_codes(ctx, slot, codes->data);
	/* the staging copy is asynchronous and the parser's buffer is about to be reused */
	if(!rc) rc = ccg_sync(ctx);
	/* -P in pair mode: getIncPosPtr(includes[i], seq, seq, proxi) (cdist.c:91) only CLEARS mask bits, and which ones
	 * depends on the sequence alone -- so it commutes with maskMotifs (cdist.c:90), and it runs first here because the
	 * device finds the unknown positions in the still pristine mask */
	if(!rc && proxi_apply) rc = ccg_sample_proximity(ctx, slot, 1, 1, inc);
	if(!rc && g_defer_motifs) {
		if(inc) rc = ccg_sample_count_masked(ctx, slot, inc);
	} else if(!rc && (g_motifs.n || (inc && !proxi_apply))) rc = ccg_mask_motifs(ctx, slot, 1, inc);
	if(rc) die_gpu(ctx, rc);
}

/* The count the inclusion test of cdist.c:91-100 / :138-147 looks at, for a candidate whose codes are in r.
 * Plain: the number of known bases, which the parser thread has counted.  With -y the methylation sites of the
 * motif matches are gone from it, and with -P and the event definition of getIncPos (fsacmp.c:181: not -f 8 /
 * -f 32) so are the bases between two unknown positions at most proxi apart; the device applies (pair mode) or
 * just counts (the shared-mask reference candidate, whose ranges ccg_build_global_mask clears later) that on the
 * freshly uploaded sample.  Later shared-mask samples are gated on their known bases alone (cdist.c:102).
 * *uploaded tells the caller that slot already holds the sample. */
static unsigned candidate_count(const DistOpts *o, ccg_ctx *ctx, int slot, const ByteBuf *codes, unsigned known, int len,
                                int pair, int is_ref_candidate, int *uploaded) {
	*uploaded = 0;
	const int proxi_counts = o->proxi && !(o->flag & (8 | 32));
	if((!proxi_counts && !g_motifs.n) || len <= 0) return known;
	if(!pair && !is_ref_candidate) return known;
	unsigned inc = 0;
	upload_sample(ctx, slot, codes, &inc, proxi_counts && pair);
	*uploaded = 1;
	if(g_defer_motifs) return inc;               /* ccg_sample_count_masked has counted both maskings */
	if(proxi_counts && !pair) {
		/* the shared-mask reference candidate: its ranges are only counted here, ccg_build_global_mask clears them */
		int rc = ccg_sample_proximity(ctx, slot, 1, 0, &inc);
		if(rc) die_gpu(ctx, rc);
	}
	return inc;
}

/* -V: where fsaCmpThreadOut's `diffile` lines go (dist.c:85-95); NULL without -V */
static FILE *g_diffile = 0;

/* The device context of a FASTA run.  The reference spreads the pair loop over -t threads inside fsaCmpThreadOut
 * (fsacmpthrd.c:76-106); here the same calls spread over every visible GPU: ccg_init_multi returns one handle, the
 * library cuts the alignment between the GPUs when the job is worth it and works on one device otherwise.
 * -P, -y and -V need a whole sample on one device, so those runs open a single-device context.
 * CCPHYLO_GPUS=N limits the number of GPUs (1 = single device). */
static ccg_ctx *open_fasta_context(const DistOpts *o) {
	ccg_ctx *ctx = 0;
	const char *env = getenv("CCPHYLO_GPUS");
	const int want = env ? atoi(env) : 0;
	const int plain = !o->proxi && !g_motifs.n && !g_diffile;
	int rc = (plain && want != 1) ? ccg_init_multi(&ctx, want) : ccg_init(&ctx, -1);
	if(rc) die_gpu(0, rc);
	return ctx;
}

/* printDiff (fsacmp.c:635-644) for one pair's list from ccg_list_variants: "(%d, %d)\t%c%d%c\n" per variant.  A listing
 * is millions of short lines, so they are put together by hand in a buffer (the pair's prefix once, the label's digits
 * backwards) instead of going through fprintf one by one. */
static int print_variants(void *user, int sample_i, int sample_j, const uint64_t *variants, size_t count) {
	FILE *f = (FILE *) user;
	static char buf[1 << 16];
	char prefix[48];
	const int plen = snprintf(prefix, sizeof(prefix), "(%d, %d)\t", sample_i, sample_j);
	size_t used = 0;
	for(size_t k = 0; k < count; ++k) {
		if(used + (size_t) plen + 32 > sizeof(buf)) {
			if(fwrite(buf, 1, used, f) != used) return 1;
			used = 0;
		}
		memcpy(buf + used, prefix, (size_t) plen);
		used += (size_t) plen;
		buf[used++] = "ACGT"[(variants[k] >> 2) & 3];
		/* the reference prints the unsigned counter with %d */
		int label = (int) (variants[k] >> 4);
		char digits[12];
		int nd = 0;
		unsigned mag = label < 0 ? 0u - (unsigned) label : (unsigned) label;
		do { digits[nd++] = (char) ('0' + mag % 10); mag /= 10; } while(mag);
		if(label < 0) buf[used++] = '-';
		while(nd) buf[used++] = digits[--nd];
		buf[used++] = "ACGT"[variants[k] & 3];
		buf[used++] = '\n';
	}
	if(used && fwrite(buf, 1, used, f) != used) return 1;
	return 0;
}

/* shared tail of the two FASTA modes: compare on the device and print (cdist.c:170-192, dist.c:174-180) */
static int compare_and_print(const DistOpts *o, ccg_ctx *ctx, int n, int len, unsigned minLength, unsigned char *include,
                             int included, char **names, const char *comment, FILE *outfile, FILE *noutfile, int n_into_out) {
	const int pair = (o->flag & 2) != 0;
	const long long cells_all = (long long) n * (n - 1) / 2;
	if(cells_all < o->threads)
		fprintf(stderr, "Adjustning number of nodes to %d, to conform with the matrix size.\n", (int) cells_all);
	if(!included) {
		fprintf(stderr, "All sequences were trimmed away.\n");
		return 0;
	}
	void *D = dist_alloc_cells(o, (size_t) included, o->elem_size);
	void *N = (pair && noutfile) ? dist_alloc_cells(o, (size_t) included, o->elem_size) : 0;
	int Dn = 0, rc;
	const double t_cmp = now_s();
	if(pair) {
		if(g_diffile) {
			rc = ccg_list_variants(ctx, 1, include, print_variants, g_diffile);
			if(rc) die_gpu(ctx, rc);
		}
		/* minLength was already maxed with minCov * len; the library repeats that (fsacmpthrd.c:292) */
		rc = ccg_run_pair(ctx, include, o->norm, minLength, o->minCov, o->elem_size, o->byteScale, D, N, &Dn);
		if(rc) die_gpu(ctx, rc);
	} else {
		unsigned ginc = 0;
		rc = ccg_build_global_mask(ctx, include, &ginc);
		if(rc) die_gpu(ctx, rc);
		fprintf(stderr, "# %d / %d bases included in distance matrix.\n", (int) ginc, len);
		if(g_diffile) {
			rc = ccg_list_variants(ctx, 0, include, print_variants, g_diffile);
			if(rc) die_gpu(ctx, rc);
		}
		rc = ccg_run_global(ctx, include, o->norm, o->elem_size, o->byteScale, D, &Dn, &ginc);
		if(rc) die_gpu(ctx, rc);
	}
	stat_line("  compare (device, incl. result copy)", t_cmp);
	if(getenv("CCPHYLO_GPU_STATS")) {
		int active = 1;
		const int gpus = ccg_multi_gpus(ctx, &active);
		fprintf(stderr, "# gpu-stats\t%d of %d GPU(s)\t%s\n", active, gpus, ccg_last_kernel(ctx));
	}
	if(Dn > 1) {
		phy_write_mt(outfile, D, o->elem_size, o->byteScale, Dn, names, include, comment, o->flag, o->precision, o->threads);
		if(N) phy_write_mt(n_into_out ? outfile : noutfile, N, o->elem_size, o->byteScale, Dn, names, include, comment, o->flag, o->precision, o->threads);
	}
	dist_free_cells(o, D, (size_t) included, o->elem_size);
	dist_free_cells(o, N, (size_t) included, o->elem_size);
	return Dn;
}

/* ltdFsaMatrix_get (cdist.c:36-194) */
static void dist_fasta_files(const DistOpts *o, FILE *outfile, FILE *noutfile) {
	const int n = (int) o->numFile;
	FsaJob fj;
	fj.o = o;
	fsa_code_table(o->flag, fj.table);
	int nthreads = o->threads < 1 ? 1 : o->threads;
	if(nthreads > n) nthreads = n;
	const int window = nthreads + 2;
	Parsed *slots = calloc((size_t) window, sizeof(Parsed));
	if(!slots) die_errno();
	for(int k = 0; k < window; ++k) bytebuf_init(&slots[k].codes, 1 << 20);
	OrderedPool *pool = pool_start(n, nthreads, window, slots, sizeof(Parsed), parse_one, &fj);
	if(!pool) die_errno();

	double t0 = now_s();
	ccg_ctx *ctx = open_fasta_context(o);
	int rc;
	stat_line("device context", t0);
	t0 = now_s();
	unsigned char *include = malloc((size_t) n);
	if(!include) die_errno();
	memset(include, 1, (size_t) n);
	unsigned minLength = o->minLength;
	int len = 0, have_ref = 0, included = n;
	const int pair = (o->flag & 2) != 0;
	rc = ccg_set_proximity(ctx, o->proxi, (o->flag & (8 | 32)) != 0);
	if(!rc && g_motifs.n) rc = ccg_set_motifs(ctx, g_motifs.n, g_motifs.lens, g_motifs.sets);
	if(rc) die_gpu(ctx, rc);

	for(int i = 0; i < n; ++i) {
		Parsed *r = (Parsed *) pool_take(pool, i);
		const char *path = o->filenames[i];
		switch(r->status) {
			case PARSE_OPEN_FAILED:
				if(r->err) {
					errno = r->err;
					fprintf(stderr, "Filename:\t%s\n", path);
					die_errno();
				}
				fprintf(stderr, "Cannot determine format of file:\t%s\n", path);
				exit(1);
			case PARSE_NOT_FASTA:
				fprintf(stderr, "\"%s\" is not fasta.\n", path);
				exit(1);
			case PARSE_NO_TEMPLATE:
				fprintf(stderr, "Missing template entry (\"%s\") in file:\t%s\n", o->targetTemplate, path);
				include[i] = 0;
				--included;
				break;
			case PARSE_NO_SEQ:
				fprintf(stderr, "Missing template sequence (\"%s\") in file:\t%s\n", o->targetTemplate, path);
				include[i] = 0;
				--included;
				break;
			default: {
				if(have_ref) {
					if((int) r->codes.len != len) {
						fprintf(stderr, "Sequences does not match: %s\n", path);
						exit(1);
					}
				} else {
					/* until a sample passes, every candidate redefines the alignment length (cdist.c:113-147) */
					len = (int) r->codes.len;
					if(minLength < o->minCov * len) minLength = (unsigned) (o->minCov * len);
					/* a pair-mode store serves both modes: the shared mask is ANDed in afterwards */
					rc = ccg_set_problem(ctx, n, len, 1);
					if(rc) die_gpu(ctx, rc);
				}
				int uploaded = 0;
				const unsigned inc = candidate_count(o, ctx, i, &r->codes, r->known, len, pair, !have_ref, &uploaded);
				if(inc < minLength) {
					fprintf(stderr, "# Excluded:\t%s\t( %d / %d )\n", path, (int) inc, len);
					include[i] = 0;
					--included;
				} else {
					fprintf(stderr, "# Included:\t%s\t( %d / %d )\n", path, (int) inc, len);
					have_ref = 1;
					if(len > 0 && !uploaded) upload_sample(ctx, i, &r->codes, 0, 0);
				}
			}
		}
		pool_release(pool, i);
	}
	pool_finish(pool);
	stat_line("parse + upload", t0);
	t0 = now_s();
	if(!have_ref) included = 0;
	compare_and_print(o, ctx, n, len, minLength, include, included, o->filenames, o->targetTemplate, outfile, noutfile, 0);
	stat_line("compare + print", t0);
	ccg_destroy(ctx);
	for(int k = 0; k < window; ++k) bytebuf_free(&slots[k].codes);
	free(slots);
	free(include);
}

/* ltdMsaMatrix_get (cdist.c:196-390): one multi-FASTA alignment, records are the samples */
typedef struct {
	int status;              /* PARSE_OK | PARSE_NO_SEQ | PARSE_OPEN_FAILED */
	ByteBuf header, codes;
	unsigned known;
} MsaParsed;

typedef struct {
	const char *path;
	const long long *offsets;     /* stream offset of every record's '>' */
	unsigned char table[256];
} MsaJob;

static void msa_parse_one(int job, void *state, void *user) {
	const MsaJob *mj = (const MsaJob *) user;
	MsaParsed *r = (MsaParsed *) state;
	r->known = 0;
	r->codes.len = 0;
	r->status = PARSE_OPEN_FAILED;
	FsaReader *fr = fsa_open(mj->path);
	if(!fr) return;
	if(fsa_seek(fr, mj->offsets[job]) == 0 && fsa_next_header(fr, &r->header)) {
		if(!fsa_read_codes(fr, mj->table, &r->codes)) r->status = PARSE_NO_SEQ;
		else {
			unsigned known = 0;
			for(size_t k = 0; k < r->codes.len; ++k) known += r->codes.data[k] < 4;
			r->known = known;
			r->status = PARSE_OK;
		}
	}
	fsa_close(fr);
}

static void dist_fasta_msa(const DistOpts *o, FILE *outfile, FILE *noutfile) {
	const char *path = o->numFile ? o->filenames[0] : "-";
	MsaJob mj;
	fsa_code_table(o->flag, mj.table);
	const int pair = (o->flag & 2) != 0;

	/* The device store is sized up front, so the records are counted first (a second pass is far cheaper
	 * than keeping the alignment in host memory); the same pass notes where every record starts, so that a
	 * plain (uncompressed) alignment can then be parsed by -t threads in parallel, each seeking to its
	 * record.  Compressed input is read twice, sequentially. */
	if(strcmp(path, "-") == 0) {
		fprintf(stderr, "MSA input from stdin is not supported on the GPU path (the alignment is read twice).\n");
		exit(1);
	}
	FsaReader *fr = fsa_open(path);
	if(!fr) {
		fprintf(stderr, "Filename:\t%s\n", path);
		die_errno();
	}
	if(fsa_peek(fr) < 0) {
		fprintf(stderr, "Cannot determine format of file:\t%s\n", path);
		exit(1);
	}
	const int plain = fsa_is_plain(fr);
	ByteBuf header;
	bytebuf_init(&header, 256);
	int nrec = 0, cap = 1024;
	long long *offsets = malloc((size_t) cap * sizeof(long long)), off = 0;
	if(!offsets) die_errno();
	while(fsa_next_header_off(fr, &header, &off)) {
		if(nrec == cap) {
			cap <<= 1;
			offsets = realloc(offsets, (size_t) cap * sizeof(long long));
			if(!offsets) die_errno();
		}
		offsets[nrec++] = off;
	}
	fsa_close(fr);

	int nthreads = o->threads < 1 ? 1 : o->threads;
	if(nthreads > nrec) nthreads = nrec > 0 ? nrec : 1;
	const int parallel = plain && nthreads > 1;
	const int window = parallel ? nthreads + 2 : 1;
	MsaParsed *slots = calloc((size_t) window, sizeof(MsaParsed));
	if(!slots) die_errno();
	for(int k = 0; k < window; ++k) {
		bytebuf_init(&slots[k].header, 256);
		bytebuf_init(&slots[k].codes, 1 << 20);
	}
	mj.path = path;
	mj.offsets = offsets;
	OrderedPool *pool = 0;
	if(parallel) {
		pool = pool_start(nrec, nthreads, window, slots, sizeof(MsaParsed), msa_parse_one, &mj);
		if(!pool) die_errno();
	} else {
		fr = fsa_open(path);
		if(!fr) die_errno();
	}

	ccg_ctx *ctx = open_fasta_context(o);
	int rc = ccg_set_proximity(ctx, o->proxi, (o->flag & (8 | 32)) != 0);
	if(!rc && g_motifs.n) rc = ccg_set_motifs(ctx, g_motifs.n, g_motifs.lens, g_motifs.sets);
	if(rc) die_gpu(ctx, rc);
	char **names = calloc((size_t) (nrec ? nrec : 1), sizeof(char *));
	if(!names) die_errno();
	unsigned minLength = o->minLength;
	int len = 0, have_ref = 0, n = 0;
	for(int job = 0; job < nrec; ++job) {
		MsaParsed *r;
		if(parallel) r = (MsaParsed *) pool_take(pool, job);
		else {
			r = &slots[0];
			r->status = PARSE_OPEN_FAILED;
			if(fsa_next_header(fr, &r->header)) {
				if(!fsa_read_codes(fr, mj.table, &r->codes)) r->status = PARSE_NO_SEQ;
				else {
					unsigned known = 0;
					for(size_t k = 0; k < r->codes.len; ++k) known += r->codes.data[k] < 4;
					r->known = known;
					r->status = PARSE_OK;
				}
			}
		}
		if(r->status != PARSE_OK) {
			/* a header at the very end of the file is no record (seqparse.c:28); anything else is an I/O error */
			if(r->status == PARSE_OPEN_FAILED) {
				fprintf(stderr, "Filename:\t%s\n", path);
				die_errno();
			}
			if(parallel) pool_release(pool, job);
			continue;
		}
		const char *name = (const char *) r->header.data;
		if(have_ref) {
			if((int) r->codes.len != len) {
				fprintf(stderr, "Sequences does not match: >%s\n", name);
				exit(1);
			}
		} else {
			len = (int) r->codes.len;
			if(minLength < o->minCov * len) minLength = (unsigned) (o->minCov * len);
			rc = ccg_set_problem(ctx, nrec, len, 1);
			if(rc) die_gpu(ctx, rc);
		}
		int uploaded = 0;
		const unsigned known = candidate_count(o, ctx, n, &r->codes, r->known, len, pair, !have_ref, &uploaded);
		/* shared-mask mode keeps a later record only if it EXCEEDS the threshold (cdist.c:270) */
		const int keep = (have_ref && !pair) ? (minLength < known) : !(known < minLength);
		if(!keep) fprintf(stderr, "# Excluded:\t%s\t( %d / %d )\n", name, (int) known, len);
		else {
			fprintf(stderr, "# Included:\t%s\t( %d / %d )\n", name, (int) known, len);
			have_ref = 1;
			names[n] = strdup(name);
			if(!names[n]) die_errno();
			if(len > 0 && !uploaded) upload_sample(ctx, n, &r->codes, 0, 0);
			++n;
		}
		if(parallel) pool_release(pool, job);
	}
	if(parallel) pool_finish(pool);
	else fsa_close(fr);
	if(n == 0) {
		/* no record was kept: the reference shrinks its arrays to zero bytes, takes realloc's NULL for a failure and
		 * leaves through ERROR() with errno still 0 (cdist.c:323-329) -- this line, exit code 0, empty outputs */
		fprintf(stderr, "Error: %d (%s)\n", 0, strerror(0));
		fflush(outfile);
		if(noutfile) fflush(noutfile);
		exit(0);
	}
	if(n == 1 && o->mmap_matrix) {
		/* one record kept and -H: the reference sizes its file-backed matrix for 0 cells, seeks to byte -1 of the
		 * temporary file and leaves through ERROR() (matrix.c:150-155, cdist.c:333) */
		errno = EINVAL;
		fflush(outfile);
		if(noutfile) fflush(noutfile);
		die_errno();
	}
	/* excluded records were dropped: the n kept samples occupy slots 0..n-1 */
	unsigned char *include = malloc((size_t) (nrec ? nrec : 1));
	if(!include) die_errno();
	memset(include, 0, (size_t) (nrec ? nrec : 1));
	memset(include, 1, (size_t) n);
	/* the reference prints the N block into the .phy stream, right behind the D block (cdist.c:364-369) */
	compare_and_print(o, ctx, n, len, minLength, include, n, names, 0, outfile, noutfile, 1);
	ccg_destroy(ctx);
	for(int k = 0; k < n; ++k) free(names[k]);
	free(names);
	free(include);
	free(offsets);
	for(int k = 0; k < window; ++k) {
		bytebuf_free(&slots[k].header);
		bytebuf_free(&slots[k].codes);
	}
	free(slots);
	bytebuf_free(&header);
}

/* ------------------------------------------------------------------------------------------
 * -a: one more sample against an existing matrix (add2Matrix dist.c:331-411, ltdFsaRowThrd
 * fsacmpthrd.c:582-667, printphyUpdate phy.c:201-250)
 * ------------------------------------------------------------------------------------------ */
static int add_fasta_row(const DistOpts *o, const PhyNames *phy, double *D, double *N) {
	const int n = phy->n;
	FsaJob fj;
	DistOpts one = *o;
	char *addname = o->addfilename;
	one.filenames = &addname;
	fj.o = &one;
	fsa_code_table(o->flag, fj.table);
	Parsed added;
	memset(&added, 0, sizeof(added));
	bytebuf_init(&added.codes, 1 << 20);
	parse_one(0, &added, &fj);
	if(added.status == PARSE_OPEN_FAILED) {
		errno = added.err;
		fprintf(stderr, "Filename:\t%s\n", addname);
		die_errno();
	}
	if(added.status == PARSE_NO_TEMPLATE || added.status == PARSE_NOT_FASTA) {
		fprintf(stderr, "Missing template entry (\"%s\") in file:\t%s\n", o->targetTemplate, addname);
		exit(1);
	}
	if(added.status == PARSE_NO_SEQ) {
		fprintf(stderr, "\"%s\" is not fasta.\n", addname);
		exit(1);
	}
	const int len = (int) added.codes.len;
	unsigned minLength = o->minLength;
	if(minLength < o->minCov * len) minLength = (unsigned) (o->minCov * len);
	ccg_ctx *ctx = 0;
	int rc = ccg_init(&ctx, -1);
	if(rc) die_gpu(0, rc);
	rc = ccg_set_proximity(ctx, o->proxi, (o->flag & (8 | 32)) != 0);
	if(!rc) rc = ccg_set_problem(ctx, n + 1, len, 1);
	if(rc) die_gpu(ctx, rc);
	/* the new sample's own count (fsacmpthrd.c:627-629): with -P and getIncPos' events the bases between close
	 * unknown positions do not count; ccg_run_row applies that masking itself, here it is only counted */
	unsigned inc = added.known;
	if(len > 0) {
		rc = ccg_put_sample_codes(ctx, n, added.codes.data);
		if(!rc) rc = ccg_sync(ctx);
		if(!rc && o->proxi && !(o->flag & (8 | 32))) rc = ccg_sample_proximity(ctx, n, 1, 0, &inc);
		if(rc) die_gpu(ctx, rc);
	}
	if(inc < minLength) {
		fprintf(stderr, "Template (\"%s\") did not exceed threshold for inclusion:\t%s\n", o->targetTemplate, addname);
		ccg_destroy(ctx);
		return 1;
	}
	int threads = o->threads < 1 ? 1 : o->threads;
	if(n < threads) fprintf(stderr, "Adjustning number of nodes to %d, to conform with the matrix size.\n", (threads = n));
	if(n < 1) {
		ccg_destroy(ctx);
		return 0;
	}
	/* the samples of the existing matrix, parsed by the host threads, into the slots below */
	DistOpts old = *o;
	old.filenames = phy->paths;
	fj.o = &old;
	const int window = threads + 2;
	Parsed *slots = calloc((size_t) window, sizeof(Parsed));
	if(!slots) die_errno();
	for(int k = 0; k < window; ++k) bytebuf_init(&slots[k].codes, 1 << 20);
	OrderedPool *pool = pool_start(n, threads, window, slots, sizeof(Parsed), parse_one, &fj);
	if(!pool) die_errno();
	for(int j = 0; j < n; ++j) {
		Parsed *r = (Parsed *) pool_take(pool, j);
		if(r->status == PARSE_OPEN_FAILED && r->err) {
			errno = r->err;
			fprintf(stderr, "Filename:\t%s\n", phy->paths[j]);
			die_errno();
		}
		if(r->status != PARSE_OK) {
			/* the reference goes on with whatever its buffers hold (fsacmpthrd.c:538-540) */
			fprintf(stderr, "Missing template entry (\"%s\") in file:\t%s\n", o->targetTemplate, phy->paths[j]);
			exit(1);
		}
		if((int) r->codes.len != len) {
			fprintf(stderr, "New sequence does not match the existing sequences.\n");
			exit(1);
		}
		if(len > 0) {
			rc = ccg_put_sample_codes(ctx, j, r->codes.data);
			if(!rc) rc = ccg_sync(ctx);
			if(rc) die_gpu(ctx, rc);
		}
		pool_release(pool, j);
	}
	pool_finish(pool);
	for(int k = 0; k < window; ++k) bytebuf_free(&slots[k].codes);
	free(slots);
	if(len > 0 && o->diffilename) {
		/* ltdFsaRowThrd appends to the variant file (fsacmpthrd.c:632-637) */
		FILE *diffile = strcmp(o->diffilename, "-") == 0 ? stdout : fopen(o->diffilename, "ab");
		if(!diffile) {
			fprintf(stderr, "Filename:\t%s\n", o->diffilename);
			die_errno();
		}
		rc = ccg_list_variants_row(ctx, n, print_variants, diffile);
		if(rc) die_gpu(ctx, rc);
		if(diffile != stdout) fclose(diffile);
		else fflush(stdout);
	}
	if(len > 0) {
		int cols = 0;
		rc = ccg_run_row(ctx, n, o->norm, minLength, o->minCov, D, N, &cols);
		if(rc) die_gpu(ctx, rc);
	} else {
		for(int j = 0; j < n; ++j) {
			D[j] = minLength ? -1.0 : 0.0;
			N[j] = 0.0;
		}
	}
	for(int j = 0; j < n; ++j)
		if(D[j] == -1.0) fprintf(stderr, "No sufficient overlap with sample:\t%s\n", phy->paths[j]);
	ccg_destroy(ctx);
	bytebuf_free(&added.codes);
	return 0;
}

static int add_to_matrix(const DistOpts *o) {
	/* directory of the first input file: the names of the matrix are looked up there (dist.c:344-356) */
	char *dir = strdup(o->filenames[0]);
	if(!dir) die_errno();
	char *slash = strrchr(dir, '/');
	if(slash) slash[1] = 0;
	PhyNames phy;
	const int ok = read_phy_names(o->outputfilename, dir, o->sep, &phy);
	free(dir);
	if(ok == 0) {
		errno |= 1;
		die_errno();
	}
	if(ok < 0) return 1;
	FsaReader *fr = fsa_open(o->addfilename);
	if(!fr) {
		fprintf(stderr, "Filename:\t%s\n", o->addfilename);
		die_errno();
	}
	const int informat = fsa_peek(fr);
	fsa_close(fr);
	const int n = phy.n;
	double *D = malloc((size_t) (n ? n : 1) * sizeof(double)), *N = malloc((size_t) (n ? n : 1) * sizeof(double));
	if(!D || !N) die_errno();
	int failed;
	if(informat == '>') failed = add_fasta_row(o, &phy, D, N);
	else failed = dist_mat_add_row(o, n, phy.paths, D, N);
	if(failed) {
		fprintf(stderr, "Distance measures failed and thus the matrix was not updated.\n");
		return 1;
	}
	phy_append_row(o->outputfilename, n + 1, o->addfilename, D, o->flag, o->precision);
	if(o->noutputfilename) phy_append_row(o->noutputfilename, n + 1, o->addfilename, N, o->flag, o->precision);
	for(int i = 0; i < n; ++i) free(phy.paths[i]);
	free(phy.paths);
	free(D);
	free(N);
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * makeMatrix (dist.c:42-329)
 * ------------------------------------------------------------------------------------------ */
static void make_matrix(DistOpts *o) {
	FILE *outfile = open_out(o->outputfilename);
	FILE *noutfile = 0;
	if(o->noutputfilename) {
		if(strcmp(o->noutputfilename, o->outputfilename) == 0) noutfile = outfile;
		else noutfile = open_out(o->noutputfilename);
	}
	if(o->methfilename && motifs_load(o->methfilename, &g_motifs)) exit(1);
	if(o->diffilename) {
		if(strcmp(o->diffilename, o->outputfilename) == 0) g_diffile = outfile;
		else g_diffile = open_out(o->diffilename);
	}
	int informat;
	if(o->flag & 16) informat = '>';
	else if(o->numFile) {
		FsaReader *fr = fsa_open(o->filenames[0]);
		if(!fr) {
			fprintf(stderr, "Filename:\t%s\n", o->filenames[0]);
			die_errno();
		}
		informat = fsa_peek(fr);
		fsa_close(fr);
		if(informat < 0) {
			fprintf(stderr, "Cannot determine format of file:\t%s\n", o->filenames[0]);
			exit(1);
		}
	} else informat = '#';                       /* stdin is a union file (dist.c:104) */
	if(informat != '>') informat = '#';

	if(o->targetTemplate && o->numFile > 1) {
		if(informat == '>') dist_fasta_files(o, outfile, noutfile);
		else dist_mat_files(o, outfile, noutfile);
	} else if(o->numFile < 2 && informat == '#') {
		dist_mat_union(o, outfile, noutfile);
	} else if(o->numFile < 2) {
		dist_fasta_msa(o, outfile, noutfile);
	} else {
		fprintf(stderr, "Invalid argument combination.\n");
		exit(1);
	}
	if(g_diffile && g_diffile != outfile && g_diffile != stdout) fclose(g_diffile);
	if(outfile != stdout) fclose(outfile);
	else fflush(stdout);
	if(noutfile && noutfile != outfile && noutfile != stdout) fclose(noutfile);
}

/* ------------------------------------------------------------------------------------------
 * main_dist (dist.c:473-840): options
 * ------------------------------------------------------------------------------------------ */
static int help_message(FILE *out) {
	static const struct { char c; const char *name, *desc, *def; } rows[] = {
		{'i', "input", "Input file(s)", "stdin"},
		{'o', "output", "Output file", "stdout"},
		{'n', "nucleotide_numbers", "Output number of nucleotides included", "False/None"},
		{'S', "separator", "Separator", "\\t"},
		{'x', "print_precision", "Floating point print precision", "9"},
		{'y', "methylation_motifs", "Mask methylation motifs from <file>", "False/None"},
		{'V', "nucleotide_variations", "Output nucleotide variations", "False/None"},
		{'r', "reference", "Target reference", "None"},
		{'a', "add", "Add file to existing matrix", ""},
		{'E', "min_depth", "Minimum depth", "15"},
		{'C', "min_cov", "Minimum coverage", "50.0%"},
		{'L', "min_len", "Minimum overlapping length", "1"},
		{'W', "normalization_weight", "Normalization weight", "0 / None"},
		{'P', "proximity", "Minimum proximity between SNPs", "0"},
		{'f', "flag", "Output flags", "1"},
		{'F', "flag_help", "Help on option \"-f\"", ""},
		{'d', "distance", "Distance method", "cos"},
		{'D', "distance_help", "Help on option \"-d\"", ""},
		{'l', "significance_lvl", "Minimum lvl. of significance", "0.05"},
		{'p', "float_precision", "Float precision on distance matrix", "double"},
		{'s', "short_precision", "Short precision on distance matrix", "double / 1e0"},
		{'b', "byte_precision", "Byte precision on distance matrix", "double / 1e0"},
		{'H', "mmap", "Accepted for compatibility (matrices live in pinned host memory)", "False"},
		{'T', "tmp", "Accepted for compatibility", ""},
		{'t', "threads", "Number of host threads (file parsing)", "1"},
		{'h', "help", "Shows this helpmessage", ""},
	};
	fprintf(out, "#ccphylo-b200 dist: distances between samples from KMA consensus alignments or count matrices, computed on a B200 GPU.\n");
	fprintf(out, "#   %-24s\t%-32s\t%s\n", "Options are:", "Desc:", "Default:");
	for(size_t k = 0; k < sizeof(rows) / sizeof(rows[0]); ++k)
		fprintf(out, "#    -%c, --%-16s\t%-32s\t%s\n", rows[k].c, rows[k].name, rows[k].desc, rows[k].def);
	return out == stderr;
}

static char short_of(const char *longname) {
	static const struct { const char *name; char c; } map[] = {
		{"input", 'i'}, {"output", 'o'}, {"nucleotide_numbers", 'n'}, {"separator", 'S'}, {"print_precision", 'x'},
		{"methylation_motifs", 'y'}, {"nucleotide_variations", 'V'}, {"reference", 'r'}, {"add", 'a'}, {"min_depth", 'E'},
		{"min_cov", 'C'}, {"min_len", 'L'}, {"normalization_weight", 'W'}, {"proximity", 'P'}, {"flag", 'f'},
		{"flag_help", 'F'}, {"distance", 'd'}, {"distance_help", 'D'}, {"significance_lvl", 'l'}, {"float_precision", 'p'},
		{"short_precision", 's'}, {"byte_precision", 'b'}, {"mmap", 'H'}, {"tmp", 'T'}, {"threads", 't'}, {"help", 'h'},
	};
	for(size_t k = 0; k < sizeof(map) / sizeof(map[0]); ++k)
		if(strcmp(map[k].name, longname) == 0) return map[k].c;
	return 0;
}

int main_dist(int argc, char **argv) {
	DistOpts o;
	memset(&o, 0, sizeof(o));
	o.precision = 9;
	o.elem_size = 8;
	o.byteScale = 1.0;
	o.flag = 1;
	o.minDepth = 15;
	o.minLength = 1;
	o.threads = 1;
	o.outputfilename = "-";
	o.minCov = 0.5;
	o.alpha = 0.05;
	o.method = "cos";
	o.sep = '\t';
	int flag_help = 0;

	OptScan sc;
	optscan_init(&sc, argc - 1, argv + 1);
	char c, longname[64];
	while(optscan_next(&sc, &c, longname, sizeof(longname))) {
		char word[80];
		if(!c) {
			c = short_of(longname);
			if(!c) {
				snprintf(word, sizeof(word), "--%s", longname);
				die_unknown(word);
			}
		}
		switch(c) {
			case 'i': o.filenames = optscan_list(&sc, (int *) &o.numFile); break;
			case 'o': o.outputfilename = optscan_arg(&sc); break;
			case 'n': o.noutputfilename = optscan_arg(&sc); break;
			case 'S': o.sep = (char) optscan_char(&sc); break;
			case 'x': o.precision = (int) optscan_long(&sc); break;
			case 'y': o.methfilename = optscan_arg(&sc); break;
			case 'V': o.diffilename = optscan_arg(&sc); break;
			case 'r': o.targetTemplate = optscan_arg(&sc); break;
			case 'a': o.addfilename = optscan_arg(&sc); break;
			case 'E': o.minDepth = (unsigned) optscan_double(&sc); break;
			case 'C': o.minCov = optscan_double(&sc) / 100; break;
			case 'L': o.minLength = (unsigned) optscan_long(&sc); break;
			case 'W': o.norm = (unsigned) optscan_long(&sc); break;
			case 'P': o.proxi = (unsigned) optscan_long(&sc); break;
			case 'f': o.flag = (unsigned) optscan_long(&sc); break;
			case 'F': flag_help = 1; break;
			case 'd': o.method = optscan_arg(&sc); break;
			case 'D': o.method = 0; break;
			case 'l': o.alpha = optscan_double(&sc); break;
			case 'p': o.elem_size = 4; break;
			case 's':
				o.elem_size = 2;
				if(!sc.name[1]) sc.name[0] = 'p';   /* the reference reports a bad value of the short form "at p" (dist.c:651-653) */
				o.byteScale = optscan_optional_double(&sc, o.byteScale);
				break;
			case 'b': o.elem_size = 1; o.byteScale = optscan_optional_double(&sc, o.byteScale); break;
			case 'H': o.mmap_matrix = 1; break;
			case 'T': o.tmpdir = optscan_arg(&sc); break;
			case 't': o.threads = (int) optscan_long(&sc); break;
			case 'h': return help_message(stdout);
			default:
				/* a short option is named by its letter alone (the reference cuts the word behind it and prints from the
				 * letter on, dist.c:671-672 / trim.c) */
				snprintf(word, sizeof(word), "%c", c);
				die_unknown(word);
		}
	}
	/* trailing words are input files (dist.c:685-689) */
	if(sc.pos < sc.argc) {
		if(strcmp(sc.argv[sc.pos], "--") == 0) ++sc.pos;
		if(sc.pos < sc.argc) {
			o.filenames = sc.argv + sc.pos;
			o.numFile = (unsigned) (sc.argc - sc.pos);
		}
	}
	if(o.minCov < 0 || 1 < o.minCov) die_invalid("\"--min_cov\"");
	if(o.byteScale == 0) die_invalid(o.elem_size == 2 ? "\"--short_precision\"" : "\"--byte_precision\"");
	if(o.alpha < 0) die_invalid("\"--significance_lvl\"");
	if(flag_help) {
		fprintf(stdout, "# Format flags output, add them to combine them.\n#\n"
		                "#   1:\tRelaxed Phylip\n"
		                "#   2:\tDistances are pairwise, always true on *.mat files\n"
		                "#   4:\tInclude template name in phylip file\n"
		                "#   8:\tInclude insignificant bases in distance calculation, only affects fasta input\n"
		                "#  16:\tDistances based on fasta input\n"
		                "#  32:\tDo not include insignificant bases in pruning\n#\n");
		return 0;
	}
	if(!o.method) {
		dist_mat_method_help(stdout);
		return 0;
	}
	if(dist_mat_parse_method(&o)) die_invalid(o.method_err);
	if(!o.numFile && o.targetTemplate) o.numFile = 1;

	/* -y with -P: in pair mode (-f bit 2) both maskings run per sample (they commute); in shared-mask mode the motif
	 * sites are applied to the shared mask after the proximity pass (ccg_build_global_mask) */
	g_defer_motifs = o.methfilename && !o.addfilename && o.proxi && !(o.flag & 2);
	if(o.addfilename && o.filenames) return add_to_matrix(&o);
	make_matrix(&o);
	return 0;
}

int main_trim(int argc, char **argv);      /* trim_main.c */

int main(int argc, char **argv) {
	if(argc < 2 || strcmp(argv[1], "-h") == 0 || strcmp(argv[1], "--help") == 0) {
		fprintf(argc < 2 ? stderr : stdout,
		        "# ccphylo-b200 %s\n# usage: ccphylo-b200 dist|trim [options]   (`ccphylo-b200 dist -h`, `ccphylo-b200 trim -h` list them)\n", VERSION);
		return argc < 2;
	}
	if(strcmp(argv[1], "-v") == 0 || strcmp(argv[1], "--version") == 0) {
		fprintf(stdout, "ccphylo-b200-%s\n", VERSION);
		return 0;
	}
	if(strcmp(argv[1], "dist") == 0) return main_dist(argc - 1, argv + 1);
	if(strcmp(argv[1], "trim") == 0) return main_trim(argc - 1, argv + 1);
	fprintf(stderr, "Invalid tool specified: %s (this build provides `dist` and `trim`)\n", argv[1]);
	return 1;
}
