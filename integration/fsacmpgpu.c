#include <stdio.h>
#include <stdlib.h>
#include "ccphylo_gpu.h"     /* this repo: include/ccphylo_gpu.h */
#include "bytescale.h"       /* ByteScale */
#include "fsacmpthrd.h"
#include "matrix.h"

static void *cells(Matrix *M, int *elem) {
	if(M->mat)  { *elem = 8; return *M->mat; }
	if(M->fmat) { *elem = 4; return *M->fmat; }
	if(M->smat) { *elem = 2; return *M->smat; }
	*elem = 1; return *M->bmat;
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

	if(diffile) {                   /* -V (per-pair variant listing) stays on the reference's own CPU code */
		fsaCmpThreadOut(tnum, func, D, N, n, len, seqs, include, includes, norm, minLength,
		                minCov, diffile, targetTemplate, ref, filenames, proxi);
		return;
	}
	rc = ccg_fsa_cmp_thread_out(0, pair, Dcells, Ncells, elem, ByteScale, n, len,
	                            (const uint64_t *const *) seqs, include,
	                            (const uint32_t *const *) includes, norm, minLength, minCov,
	                            proxi, &Dn, &inc);
	if(rc) {                        /* no CPU fallback: report and stop, like ERROR() */
		fprintf(stderr, "ccphylo_gpu: %s (%s)\n", ccg_strerror(rc), ccg_last_error(0));
		exit(rc);
	}
	D->n = Dn;
	if(pair && N) N->n = Dn;        /* cmpFsaThrd never touches N: the .num file stays empty (dist.c:177) */
	if(!pair) fprintf(stderr, "# %u / %d bases included in distance matrix.\n", inc, len); /* fsacmpthrd.c:165 */
}
