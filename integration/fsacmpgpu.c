/* fsacmpgpu.c -- the reference-side binding: a drop-in for fsaCmpThreadOut (fsacmpthrd.c:76) that sends the pair loop
 * to the B200 library.  Built into the UNMODIFIED reference by oracle/Makefile (target ref_gpu: the reference's own
 * objects + this file, the four call sites cdist.c:181,184,351,354 redirected with -DfsaCmpThreadOut=fsaCmpGpuOut for
 * cdist.c only), see INTEGRATION.md. */
#include <stdio.h>
#include <stdlib.h>
#include "ccphylo_gpu.h"     /* this repo: include/ccphylo_gpu.h */
#include "bytescale.h"       /* ByteScale */
#include "fsacmp.h"
#include "fsacmpthrd.h"
#include "matrix.h"

static void *cells(Matrix *M, int *elem) {
	if(M->mat)  { *elem = 8; return *M->mat; }
	if(M->fmat) { *elem = 4; return *M->fmat; }
	if(M->smat) { *elem = 2; return *M->smat; }
	*elem = 1; return *M->bmat;
}

static void die(ccg_ctx *ctx, int rc) {            /* no CPU fallback: report and stop, like ERROR() */
	fprintf(stderr, "ccphylo_gpu: %s (%s)\n", ccg_strerror(rc), ccg_last_error(ctx));
	exit(rc);
}

/* printDiff (fsacmp.c:635-644) for one pair's list */
static int print_variants(void *user, int i, int j, const uint64_t *v, size_t count) {
	size_t k;
	for(k = 0; k < count; ++k)
		fprintf((FILE *) user, "(%d, %d)\t%c%d%c\n", i, j, "ACGT"[(v[k] >> 2) & 3], (int) (v[k] >> 4), "ACGT"[v[k] & 3]);
	return 0;
}

/* same argument list as fsaCmpThreadOut; func selects the mode exactly as before */
void fsaCmpGpuOut(int tnum, void *(*func)(void *), Matrix *D, Matrix *N, int n, int len,
                  long unsigned **seqs, unsigned char *include, unsigned **includes,
                  unsigned norm, unsigned minLength, double minCov, FILE *diffile,
                  char *targetTemplate, Qseqs *ref, Qseqs **filenames, unsigned proxi) {
	int elem, Dn = 0, pair = (func == &cmpairFsaThrd), rc;
	unsigned inc = 0;
	void *Dcells = cells(D, &elem);
	void *Ncells = (pair && N) ? cells(N, &elem) : 0;
	ccg_ctx *ctx = 0;

	if(diffile && !pair) {
		/* -V in shared-mask mode walks the UNMASKED sequences beside the global mask (fsacmprint fsacmp.c:646): that
		 * stays on the reference's own code.  (ccphylo-b200 dist builds the shared mask on the device and lists those
		 * variants there: host/dist_main.c.) */
		fsaCmpThreadOut(tnum, func, D, N, n, len, seqs, include, includes, norm, minLength,
		                minCov, diffile, targetTemplate, ref, filenames, proxi);
		return;
	}
	if(diffile) {
		/* -V, pair mode: fsacmpairint's lines (fsacmp.c:685-737) come from ccg_list_variants, in the order of a -t 1 run;
		 * with -P under maskProxi's per-pair mask (fsacmpthrd.c:410-414) */
		if((rc = ccg_init(&ctx, -1))) die(0, rc);
		if((rc = ccg_set_proximity(ctx, proxi, 0))) die(ctx, rc);
		if((rc = ccg_set_problem(ctx, n, len, 1))) die(ctx, rc);
		{
			const uint64_t **s = malloc((size_t) (n ? n : 1) * sizeof(*s));
			const uint32_t **m = malloc((size_t) (n ? n : 1) * sizeof(*m));
			int i;
			if(!s || !m) exit(1);
			for(i = 0; i < n; ++i) {
				s[i] = include[i] ? (const uint64_t *) seqs[i] : 0;
				m[i] = include[i] ? (const uint32_t *) includes[i] : 0;
			}
			if((rc = ccg_put_samples_packed(ctx, 0, n, s, m))) die(ctx, rc);
			if((rc = ccg_sync(ctx))) die(ctx, rc);
			free(s);
			free(m);
		}
		if((rc = ccg_list_variants(ctx, 1, include, print_variants, diffile))) die(ctx, rc);
		if((rc = ccg_run_pair(ctx, include, norm, minLength, minCov, elem, ByteScale, Dcells, Ncells, &Dn))) die(ctx, rc);
		ccg_destroy(ctx);
	} else {
		/* every visible GPU: the library cuts the alignment between them when the job is worth it */
		if((rc = ccg_init_multi(&ctx, 0))) die(0, rc);
		rc = ccg_fsa_cmp_thread_out(ctx, pair, Dcells, Ncells, elem, ByteScale, n, len,
		                            (const uint64_t *const *) seqs, include,
		                            (const uint32_t *const *) includes, norm, minLength, minCov,
		                            proxi, &Dn, &inc);
		if(rc) die(ctx, rc);
		ccg_destroy(ctx);
	}
	D->n = Dn;
	if(pair && N) N->n = Dn;        /* cmpFsaThrd never touches N: the .num file stays empty (dist.c:177) */
	if(!pair) fprintf(stderr, "# %u / %d bases included in distance matrix.\n", inc, len); /* fsacmpthrd.c:165 */
}
