/* dist_opts.h -- parsed options of `dist` (dist.c:484-506 defaults) shared by the FASTA and .mat drivers */
#ifndef CCB_DIST_OPTS_H
#define CCB_DIST_OPTS_H

#include <stdio.h>

typedef struct {
	unsigned numFile;
	char **filenames;
	char *outputfilename, *noutputfilename, *methfilename, *diffilename, *addfilename;
	char *targetTemplate;
	char *method;              /* -d */
	int method_id;             /* CCG_MAT_* */
	unsigned method_order;     /* n of l<n> / nl<n> */
	char method_err[32];
	double minCov, alpha, byteScale;
	unsigned flag, norm, minDepth, minLength, proxi;
	int elem_size, precision, threads;
	char sep;
	int mmap_matrix;           /* -H: result matrices on the disk (matrix.c:116 ltdMatrixMinit) */
	char *tmpdir;              /* -T: where their files go (tmp.c:27 tmpF) */
} DistOpts;

/* result matrices: packed lower triangle of n samples, elem bytes per cell -- pinned host memory, or with -H a
 * mapping of an unlinked temporary file (dist_main.c) */
void *dist_alloc_cells(const DistOpts *o, size_t n, int elem);
void dist_free_cells(const DistOpts *o, void *p, size_t n, int elem);

/* dist_mat.c: KMA .mat count-matrix inputs (ltdmatrixthrd.c:376, ltdmatrix.c:32) */
void dist_mat_files(const DistOpts *o, FILE *outfile, FILE *noutfile);
void dist_mat_union(const DistOpts *o, FILE *outfile, FILE *noutfile);
/* -a on .mat input (ltdRowThrd ltdmatrixthrd.c:564): the row of o->addfilename against the n samples in paths;
 * 0 on success, 1 when the new sample fails its gate */
int dist_mat_add_row(const DistOpts *o, int n, char **paths, double *D, double *N);
void dist_mat_method_help(FILE *out);
/* 0 on success; on failure o->method_err names the option for "Invalid value parsed at ..." */
int dist_mat_parse_method(DistOpts *o);

#endif
