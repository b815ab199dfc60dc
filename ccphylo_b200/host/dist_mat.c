/*
 * dist_mat.c -- `dist` on KMA count matrices: the multi-file mode (-r template, *.mat[.gz]
 * files; reference ltdMatrixThrd ltdmatrixthrd.c:376) and the union mode (`ccphylo union |
 * ccphylo-b200 dist`; reference ltdMatrix_get ltdmatrix.c:32 driven by dist.c:181-266).
 *
 * Each sample's template is parsed once by the host threads and uploaded to the device; the
 * reference instead re-opens and re-parses sample j for every cell (i, j).  Sample gate (both
 * modes, every sample): nNucs < minLength || nNucs < minCov * rows excludes it, where nNucs
 * counts the rows with minDepth <= total (ltdmatrixthrd.c:455-458,523-526).  The pair gate and
 * the -d arithmetic live in csrc/k_matdist.cu.
 */
#define _POSIX_C_SOURCE 200809L
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ccphylo_gpu.h"
#include "dist_opts.h"
#include "fsa_reader.h"
#include "mat_reader.h"
#include "ordered_pool.h"
#include "phy_writer.h"

static void die_errno(void) {
	fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
	exit(errno ? errno : 1);
}

static void die_gpu(ccg_ctx *ctx, int rc) {
	fprintf(stderr, "GPU error: %s (%s)\n", ccg_strerror(rc), ccg_last_error(ctx));
	exit(rc ? rc : 1);
}

/* ---- -d (dist.c:720-786) ---- */
void dist_mat_method_help(FILE *out) {
	fprintf(out, "# Distance calculation methods:\n#\n"
	             "# cos:\tCalculate distance between positions as the angle between the count vectors.\n"
	             "# z:\tMake consensus comparison if vectors passes a McNemar test\n"
	             "# chi2:\tCalculate the chi square distance\n"
	             "# nchi2:\tCalculate the normalized chi square distance\n"
	             "# c:\tCalculate the Clausen distance between the count vectors. d(A,B) = (||A-B||_1 / sum(max{Ai, Bi}))\n"
	             "# nc:\tCalculate the normalized Clausen distance between the count vectors.\n"
	             "# bc:\tCalculate the Bray-Curtis dissimilarity between the count vectors.\n"
	             "# nbc:\tCalculate the normalized Bray-Curtis dissimilarity between the count vectors.\n"
	             "# ln:\tCalculate distance between positions as the n-norm distance between the count vectors. Replace \"n\" with the waned norm\n"
	             "# linf:\tCalculate distance between positions as the l_infinity distance between the count vectors.\n"
	             "# nln:\tCalculate distance between positions as the normalized n-norm distance between the count vectors. Replace last \"n\" with the waned norm\n"
	             "# nlinf:\tCalculate distance between positions as the normalized l_infinity distance between the count vectors.\n#\n");
}

int dist_mat_parse_method(DistOpts *o) {
	static const struct { const char *name; int id; } fixed[] = {
		{"cos", CCG_MAT_COS}, {"z", CCG_MAT_Z}, {"chi2", CCG_MAT_CHI2}, {"nchi2", CCG_MAT_NCHI2}, {"nc", CCG_MAT_NC},
		{"c", CCG_MAT_C}, {"np", CCG_MAT_NP}, {"p", CCG_MAT_P}, {"nbc", CCG_MAT_NBC}, {"bc", CCG_MAT_BC},
		{"nl1", CCG_MAT_NL1}, {"nl2", CCG_MAT_NL2}, {"nlinf", CCG_MAT_NLINF}, {"l1", CCG_MAT_L1}, {"l2", CCG_MAT_L2},
		{"linf", CCG_MAT_LINF},
	};
	const char *m = o->method;
	char *end;
	for(size_t k = 0; k < sizeof(fixed) / sizeof(fixed[0]); ++k)
		if(strcmp(m, fixed[k].name) == 0) {
			o->method_id = fixed[k].id;
			return 0;
		}
	if(m[0] == 'l') {
		o->method_id = CCG_MAT_LN;
		o->method_order = (unsigned) strtoul(m + 1, &end, 10);
		if(*end) { snprintf(o->method_err, sizeof(o->method_err), "\"-d ln\""); return 1; }
		return 0;
	}
	if(strncmp(m, "nl", 2) == 0) {
		o->method_id = CCG_MAT_NLN;
		o->method_order = (unsigned) strtoul(m + 2, &end, 10);
		if(*end) { snprintf(o->method_err, sizeof(o->method_err), "\"-d nln\""); return 1; }
		return 0;
	}
	snprintf(o->method_err, sizeof(o->method_err), "\"-d\"");
	return 1;
}

/* ---- loading ---- */
typedef struct {
	int status;        /* mat_load_template's return */
	int err;
	MatSample m;
} MatParsed;

typedef struct {
	char **filenames;
	const unsigned char *want;      /* NULL = every file */
	const char *target;
	unsigned minDepth;
} MatJob;

static void load_one(int job, void *state, void *user) {
	const MatJob *mj = (const MatJob *) user;
	MatParsed *r = (MatParsed *) state;
	r->m.len = 0;
	r->m.nNucs = 0;
	if(mj->want && !mj->want[job]) {
		r->status = 2;              /* not a member of this union entry */
		return;
	}
	errno = 0;
	r->status = mat_load_template(mj->filenames[job], mj->target, mj->minDepth, &r->m);
	r->err = errno;
}

/* One template over `n` files: load, gate, upload, compare, print.  `want` preselects the files
 * (union entries), include receives the final flags.  threaded_msgs selects the wording of the
 * "No sufficient overlap" line (ltdmatrixthrd.c:320 vs ltdmatrix.c:153). */
static void one_template(const DistOpts *o, ccg_ctx *ctx, int n, char **filenames, const unsigned char *want, const char *target,
                         unsigned char *include, FILE *outfile, FILE *noutfile, int threaded_msgs) {
	MatJob mj = {filenames, want, target, o->minDepth};
	int nthreads = o->threads < 1 ? 1 : o->threads;
	if(nthreads > n) nthreads = n;
	const int window = nthreads + 2;
	MatParsed *slots = calloc((size_t) window, sizeof(MatParsed));
	if(!slots) die_errno();
	for(int k = 0; k < window; ++k) mat_sample_init(&slots[k].m);
	OrderedPool *pool = pool_start(n, nthreads, window, slots, sizeof(MatParsed), load_one, &mj);
	if(!pool) die_errno();

	/* the device store is sized by the first usable sample (all samples of a template have its rows) */
	int have_problem = 0, rc, included = 0;
	long max_len = 0;
	/* The union twin (ltdMatrix_get, ltdmatrix.c:32) walks the rows from file 1 on and never loads file 0 as a row: that
	 * sample meets its gate only when the first later sample that passes streams it as a column (cmpMats returns -2,
	 * :157-158, include[0] = 0).  So its "did not exceed threshold" line comes at that moment -- and not at all when no
	 * later sample passes. */
	const int first_member = 0;
	int first_line_due = 0;
	for(int i = 0; i < n; ++i) {
		MatParsed *r = (MatParsed *) pool_take(pool, i);
		include[i] = 0;
		if(r->status == 2) {
			/* not in this entry */
		} else if(r->status < 0) {
			if(r->err) {
				errno = r->err;
				fprintf(stderr, "Filename:\t%s\n", filenames[i]);
				die_errno();
			}
			fprintf(stderr, "Cannot determine format of file:\t%s\n", filenames[i]);
			exit(1);
		} else if(r->status == 0) {
			fprintf(stderr, "Template (\"%s\") is not included in:\t%s\n", target, filenames[i]);
		} else if(r->m.nNucs < o->minLength || r->m.nNucs < o->minCov * (double) r->m.len) {
			if(!threaded_msgs && i == first_member) first_line_due = 1;
			else fprintf(stderr, "Template (\"%s\") did not exceed threshold for inclusion:\t%s\n", target, filenames[i]);
		} else {
			if(first_line_due) {
				fprintf(stderr, "Template (\"%s\") did not exceed threshold for inclusion:\t%s\n", target, filenames[first_member]);
				first_line_due = 0;
			}
			if(!have_problem) {
				max_len = (long) r->m.len;
				rc = ccg_mat_set_problem(ctx, n, (int) max_len);
				if(rc) die_gpu(ctx, rc);
				have_problem = 1;
			}
			if((long) r->m.len > max_len) {
				/* a later sample with more rows than the first: the reference cannot compare it either
				 * (cmpMats returns -1, matcmp.c:466) */
				fprintf(stderr, "Template (\"%s\") has %zu rows in %s but %ld in the first sample.\n", target, r->m.len, filenames[i], max_len);
				exit(1);
			}
			rc = ccg_mat_put_sample(ctx, i, r->m.counts, r->m.totals, (int) r->m.len);
			if(rc) die_gpu(ctx, rc);
			rc = ccg_sync(ctx);
			if(rc) die_gpu(ctx, rc);
			include[i] = 1;
			++included;
		}
		pool_release(pool, i);
	}
	pool_finish(pool);
	for(int k = 0; k < window; ++k) mat_sample_free(&slots[k].m);
	free(slots);
	if(included < 2) return;                        /* nothing to print (dist.c:175: 1 < D->n) */

	const size_t cells = (size_t) included * (included - 1) / 2;
	void *D = dist_alloc_cells(o, (size_t) included, o->elem_size);
	void *N = dist_alloc_cells(o, (size_t) included, o->elem_size);
	uint32_t *rows = malloc(cells * sizeof(uint32_t));
	if(!D || !N || !rows) die_errno();
	int Dn = 0;
	rc = ccg_mat_run(ctx, include, o->method_id, o->method_order, o->alpha, o->norm, o->minDepth, o->minLength, o->minCov,
	                 o->elem_size, o->byteScale, D, N, &Dn, rows);
	if(rc) die_gpu(ctx, rc);
	/* pairs without sufficient overlap, in the reference's cell order */
	int *slot_of = malloc((size_t) Dn * sizeof(int));
	if(!slot_of) die_errno();
	for(int i = 0, r = 0; i < n; ++i)
		if(include[i]) slot_of[r++] = i;
	size_t k = 0;
	for(int r = 1; r < Dn; ++r)
		for(int c = 0; c < r; ++c, ++k)
			/* (with -L 0 -C 0 a pair without a single comparable position passes the gate: no line, the cell is 0/0) */
			if(rows[k] == 0 && (o->minLength > 0 || o->minCov > 0)) {
				/* the threaded loop names the row sample by its COMPACT row number (filenames[pi], ltdmatrixthrd.c:320, pi counts
				 * the included samples): with an excluded file in front that is another file's name; kept, the line is compared */
				if(threaded_msgs) fprintf(stderr, "No sufficient overlap between samples:\t%s\t%s\n", filenames[r], filenames[slot_of[c]]);
				else fprintf(stderr, "No sufficient overlap between samples:\t%s, %s\n", filenames[slot_of[r]], filenames[slot_of[c]]);
			}
	phy_write_mt(outfile, D, o->elem_size, o->byteScale, Dn, filenames, include, target, o->flag, o->precision, o->threads);
	if(noutfile) phy_write_mt(noutfile, N, o->elem_size, o->byteScale, Dn, filenames, include, target, o->flag, o->precision, o->threads);
	free(slot_of);
	free(rows);
	dist_free_cells(o, D, (size_t) included, o->elem_size);
	dist_free_cells(o, N, (size_t) included, o->elem_size);
}

/* ltdRowThrd (ltdmatrixthrd.c:564-611) + cmpMatRowThrd (:111-181) */
int dist_mat_add_row(const DistOpts *o, int n, char **paths, double *D, double *N) {
	const char *target = o->targetTemplate;
	MatSample added;
	mat_sample_init(&added);
	errno = 0;
	int st = mat_load_template(o->addfilename, target, o->minDepth, &added);
	if(st < 0) {
		fprintf(stderr, "Filename:\t%s\n", o->addfilename);
		die_errno();
	}
	if(st == 0) {
		fprintf(stderr, "Malformed matrix in:\t%s\n", o->addfilename);
		exit(1);
	}
	if(added.nNucs < o->minLength || added.nNucs < o->minCov * (double) added.len) {
		fprintf(stderr, "Template (\"%s\") did not exceed threshold for inclusion:\t%s\n", target, o->addfilename);
		return 1;
	}
	int nthreads = o->threads < 1 ? 1 : o->threads;
	if(n < nthreads) fprintf(stderr, "Adjustning number of nodes to %d, to conform with the matrix size.\n", (nthreads = n));
	if(n < 1) return 0;
	ccg_ctx *ctx = 0;
	int rc = ccg_init(&ctx, -1);
	if(rc) die_gpu(0, rc);
	rc = ccg_mat_set_problem(ctx, n + 1, (int) added.len);
	if(!rc) rc = ccg_mat_put_sample(ctx, n, added.counts, added.totals, (int) added.len);
	if(!rc) rc = ccg_sync(ctx);
	if(rc) die_gpu(ctx, rc);

	MatJob mj = {paths, 0, target, o->minDepth};
	const int window = nthreads + 2;
	MatParsed *slots = calloc((size_t) window, sizeof(MatParsed));
	if(!slots) die_errno();
	for(int k = 0; k < window; ++k) mat_sample_init(&slots[k].m);
	OrderedPool *pool = pool_start(n, nthreads, window, slots, sizeof(MatParsed), load_one, &mj);
	if(!pool) die_errno();
	/* a column sample with more rows than the new one cannot be compared (cmpMats returns -1, matcmp.c:466) */
	unsigned char *too_long = calloc((size_t) n, 1);
	if(!too_long) die_errno();
	for(int j = 0; j < n; ++j) {
		MatParsed *r = (MatParsed *) pool_take(pool, j);
		if(r->status < 0) {
			if(r->err) errno = r->err;
			fprintf(stderr, "Filename:\t%s\n", paths[j]);
			die_errno();
		}
		if(r->status == 0 || r->m.nNucs < o->minLength || r->m.nNucs < o->minCov * (double) r->m.len) {
			/* cmpMats -2 (matcmp.c:455,481): fatal in row mode (ltdmatrixthrd.c:160-162) */
			fprintf(stderr, "Template (\"%s\") did not exceed threshold for inclusion:\t%s\n", target, paths[j]);
			exit(1);
		}
		size_t len = r->m.len;
		if(len > added.len) {
			too_long[j] = 1;
			len = added.len;
		}
		rc = ccg_mat_put_sample(ctx, j, r->m.counts, r->m.totals, (int) len);
		if(!rc) rc = ccg_sync(ctx);
		if(rc) die_gpu(ctx, rc);
		pool_release(pool, j);
	}
	pool_finish(pool);
	for(int k = 0; k < window; ++k) mat_sample_free(&slots[k].m);
	free(slots);
	uint32_t *rows = malloc((size_t) n * sizeof(uint32_t));
	if(!rows) die_errno();
	rc = ccg_mat_run_row(ctx, n, o->method_id, o->method_order, o->alpha, o->norm, o->minDepth, o->minLength, o->minCov, D, N, rows);
	if(rc) die_gpu(ctx, rc);
	for(int j = 0; j < n; ++j) {
		if(too_long[j]) {
			D[j] = -1.0;
			N[j] = 0.0;
			rows[j] = 0;
		}
		if(rows[j] == 0 && (too_long[j] || o->minLength > 0 || o->minCov > 0)) fprintf(stderr, "No sufficient overlap with sample:\t%s\n", paths[j]);
	}
	free(rows);
	free(too_long);
	mat_sample_free(&added);
	ccg_destroy(ctx);
	return 0;
}

/* every visible GPU (the library cuts the position axis between them when the matrices are long enough);
 * CCPHYLO_GPUS=N limits the number, 1 = a single-device context */
static ccg_ctx *open_mat_context(void) {
	ccg_ctx *ctx = 0;
	const char *env = getenv("CCPHYLO_GPUS");
	const int want = env ? atoi(env) : 0;
	int rc = want == 1 ? ccg_init(&ctx, -1) : ccg_init_multi(&ctx, want);
	if(rc) die_gpu(0, rc);
	return ctx;
}

void dist_mat_files(const DistOpts *o, FILE *outfile, FILE *noutfile) {
	const int n = (int) o->numFile;
	ccg_ctx *ctx = open_mat_context();
	unsigned char *include = calloc((size_t) n, 1);
	if(!include) die_errno();
	one_template(o, ctx, n, o->filenames, 0, o->targetTemplate, include, outfile, noutfile, 1);
	free(include);
	ccg_destroy(ctx);
}

/* ---- union input (unionparse.c:46 header "N\tfile...\n", :134 rows "template\tnum\tidx...\n") ---- */
static char *dup_with_room(const char *s, size_t len, size_t room) {
	char *d = malloc(len + room + 1);
	if(!d) die_errno();
	memcpy(d, s, len);
	d[len] = 0;
	return d;
}

void dist_mat_union(const DistOpts *o, FILE *outfile, FILE *noutfile) {
	const char *path = o->numFile ? o->filenames[0] : "-";
	FsaReader *fr = fsa_open(path);
	if(!fr) {
		fprintf(stderr, "Filename:\t%s\n", path);
		die_errno();
	}
	ByteBuf line;
	bytebuf_init(&line, 1 << 16);
	if(!fsa_read_line(fr, &line)) {
		fprintf(stderr, "Malformed union input.\n");
		exit(1);
	}
	/* header: count, then the sample files */
	char *p = (char *) line.data, *end;
	const long n = strtol(p, &end, 10);
	if(end == p || *end != '\t' || n < 1) {
		fprintf(stderr, "Malformed union input.\n");
		exit(1);
	}
	char **filenames = calloc((size_t) n, sizeof(char *));
	if(!filenames) die_errno();
	p = end + 1;
	for(long k = 0; k < n; ++k) {
		char *tab = strchr(p, '\t');
		size_t len = tab ? (size_t) (tab - p) : strlen(p);
		/* the sample's matrix file: name cut at its last '.', plus ".mat.gz" -- or ".mat" when that
		 * does not exist (dist.c:223-250) */
		char *name = dup_with_room(p, len, 8);
		char *dot = strrchr(name, '.');
		if(dot) { *dot = 0; len = (size_t) (dot - name); }
		strcpy(name + len, (o->flag & 16) ? ".fsa.gz" : ".mat.gz");
		FILE *probe = fopen(name, "rb");
		if(probe) fclose(probe);
		else name[len + 4] = 0;
		filenames[k] = name;
		if(!tab && k + 1 < n) {
			fprintf(stderr, "Malformed union input.\n");
			exit(1);
		}
		p = tab ? tab + 1 : p + len;
	}
	if(o->flag & 16) {
		fprintf(stderr, "Union input with FASTA consensus files (-f 16) is not available on the GPU path.\n");
		exit(1);
	}
	ccg_ctx *ctx = open_mat_context();
	unsigned char *want = malloc((size_t) n), *include = malloc((size_t) n);
	if(!want || !include) die_errno();
	while(fsa_read_line(fr, &line)) {
		if(line.len == 0) continue;
		char *target = (char *) line.data;
		char *tab = strchr(target, '\t');
		if(!tab) continue;
		*tab = 0;
		long num = strtol(tab + 1, &end, 10);
		memset(want, 0, (size_t) n);
		p = end;
		for(long k = 0; k < num && *p == '\t'; ++k) {
			long idx = strtol(p + 1, &end, 10);
			if(idx >= 0 && idx < n) want[idx] = 1;
			p = end;
		}
		one_template(o, ctx, (int) n, filenames, want, target, include, outfile, noutfile, 0);
	}
	free(want);
	free(include);
	ccg_destroy(ctx);
	for(long k = 0; k < n; ++k) free(filenames[k]);
	free(filenames);
	bytebuf_free(&line);
	fsa_close(fr);
}
