/*
 * ref_shim.c -- flat C entry points onto the UNMODIFIED reference.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is ours; it is compiled together with
 * the reference's own sources, taken where they lie under /root/reference
 * (never copied into this repo), into oracle/_ref/libccphylo_ref.so by
 * oracle/Makefile.  It exists so tests and bench.py can drive the reference's
 * real hot path -- fsaCmpThreadOut (fsacmpthrd.c:76) with cmpairFsaThrd
 * (:261) / cmpFsaThrd (:108) -- on in-memory inputs through ctypes, and the
 * reference's encoders (fsacmp.c:32 get2BitTable, qseqs.c:60 qseq2nibble,
 * fsacmp.c:164 initIncPos, :181 getIncPos, :487 getNpos) to pin the oracle.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "bytescale.h"
#include "fsacmp.h"
#include "fsacmpthrd.h"
#include "matrix.h"
#include "qseqs.h"

/* Translate bytes exactly like FileBuffgetFsaSeq (seqparse.c:217-231) does
 * with the reference's own table, pack with qseq2nibble, build the pair-mode
 * mask with initIncPos + getIncPosPtr-selected builder.  `codes` (size
 * nbytes) receives the translated codes; returns len.  *inc_out = getNpos. */
long refshim_encode(const unsigned char *bytes, long nbytes, unsigned flag,
                    unsigned proxi, unsigned char *codes, uint64_t *words,
                    uint32_t *mask, int *unknown_out, int *inc_out) {
	unsigned char *table = get2BitTable(flag);
	Qseqs seq;
	long i, len = 0;

	for(i = 0; i < nbytes; ++i) {
		unsigned char c = table[bytes[i]];
		if(c < 32) codes[len++] = c;
	}
	seq.size = (unsigned) len;
	seq.len = (unsigned) len;
	seq.seq = codes;
	if(flag & 32) getIncPosPtr = &getIncPosInsigPrune;
	else if(flag & 8) getIncPosPtr = &getIncPosInsig;
	else getIncPosPtr = &getIncPos;
	initIncPos(mask, (int) len);
	*unknown_out = qseq2nibble(&seq, (long unsigned *) words);
	getIncPosPtr(mask, &seq, &seq, proxi);
	*inc_out = getNpos(mask, (int) len);
	free(table - 128);
	return len;
}

/* Global-mode mask accumulation, cdist.c:110-111: getIncPosPtr(G, seq, ref). */
void refshim_and_known(uint32_t *mask, unsigned char *seq_codes,
                       unsigned char *ref_codes, int len, unsigned proxi) {
	Qseqs seq, ref;
	seq.size = seq.len = (unsigned) len; seq.seq = seq_codes;
	ref.size = ref.len = (unsigned) len; ref.seq = ref_codes;
	getIncPos(mask, &seq, &ref, proxi);
}

/* The builder -f selects (dist.c:802-806): flag & 32 getIncPosInsigPrune, flag & 8 getIncPosInsig, else getIncPos. */
void refshim_inc_pos(uint32_t *mask, unsigned char *seq_codes, unsigned char *ref_codes, int len, unsigned proxi,
                     unsigned flag) {
	Qseqs seq, ref;
	seq.size = seq.len = (unsigned) len; seq.seq = seq_codes;
	ref.size = ref.len = (unsigned) len; ref.seq = ref_codes;
	if(flag & 32) getIncPosInsigPrune(mask, &seq, &ref, proxi);
	else if(flag & 8) getIncPosInsig(mask, &seq, &ref, proxi);
	else getIncPos(mask, &seq, &ref, proxi);
}

void refshim_init_mask(uint32_t *mask, int len) { initIncPos(mask, len); }
int refshim_mask_count(uint32_t *mask, int len) { return getNpos(mask, len); }

/* One pair through the reference kernels: maskProxi (fsacmp.c:355) +
 * fsacmpair (:587). */
uint64_t refshim_pair(uint64_t *seq_i, uint64_t *seq_j, uint32_t *inc_i,
                      uint32_t *inc_j, int len, unsigned proxi) {
	uint32_t *scratch = malloc(((size_t) len / 32 + 1) * sizeof(uint32_t));
	uint64_t r;
	maskProxi(scratch, inc_i, inc_j, (long unsigned *) seq_i, (long unsigned *) seq_j, (unsigned) len, proxi);
	r = fsacmpair((long unsigned *) seq_i, (long unsigned *) seq_j, scratch, len);
	free(scratch);
	return r;
}

/* maskProxi alone: the pair's mask into out (len / 32 + 1 words are written) */
void refshim_mask_proxi(uint64_t *seq_i, uint64_t *seq_j, uint32_t *inc_i, uint32_t *inc_j, int len, unsigned proxi, uint32_t *out) {
	maskProxi(out, inc_i, inc_j, (long unsigned *) seq_i, (long unsigned *) seq_j, (unsigned) len, proxi);
}

/* -a: getSizePhy + getFilenamesPhy (phy.c:509-650) on an existing Phylip file: the names (prefixed with dir) joined
 * by '\n' into buf; returns n, -1 if the names cannot be read, -2 if bytes are left behind the n rows (dist.c:366) */
#include "phy.h"
int refshim_phy_names(char *phyname, char *dir, char sep, char *buf, long cap) {
	FileBuff *infile = setFileBuff(1048576);
	Qseqs **names;
	int n, i, left;
	long used = 0;
	openAndDetermine(infile, phyname);
	n = getSizePhy(infile);
	names = getFilenamesPhy(dir, n, infile, sep);
	if(!names) return -1;
	left = infile->bytes;
	buf[0] = 0;
	for(i = 0; i < n; ++i) {
		long l = (long) strlen((char *) names[i]->seq);
		if(used + l + 2 > cap) break;
		memcpy(buf + used, names[i]->seq, (size_t) l);
		used += l;
		buf[used++] = '\n';
		buf[used] = 0;
	}
	closeFileBuff(infile);
	return left ? -2 : n;
}

/* -a: printphyUpdate (phy.c:201-250) after setPrecisionPhy */
void refshim_phy_update(char *phyname, int n, char *name, double *row, unsigned flag, int precision) {
	FILE *f = fopen(phyname, "rb+");
	setPrecisionPhy(precision);
	printphyUpdate(f, n, name, row, flag);
	fclose(f);
}

/* -y: getMethMotifs (methparse.c:268) on the motif file, then maskMotifs (meth.c:141) on one packed sequence */
#include "filebuff.h"
#include "meth.h"
#include "methparse.h"
int refshim_mask_motifs(char *motif_path, uint64_t *seq, uint32_t *mask, int len) {
	FileBuff *infile = setFileBuff(1048576);
	Qseqs *qseq = setQseqs(1024);
	MethMotif *motif;
	int n;
	openAndDetermine(infile, motif_path);
	motif = getMethMotifs(infile, qseq);
	closeFileBuff(infile);
	n = maskMotifs((long unsigned *) seq, mask, len, motif);
	destroyMethMotifs(motif);
	destroyQseqs(qseq);
	destroyFileBuff(infile);
	return n;
}

/* -V: the lines fsacmpairint (pair != 0, fsacmp.c:685) / fsacmprint (fsacmp.c:646) print for one pair under `mask`,
 * as text into buf (capacity cap, 0-terminated); returns the function's own return value. */
#define _GNU_SOURCE
#include <stdio.h>
FILE *open_memstream(char **ptr, size_t *sizeloc);
uint64_t refshim_variants(int pair, int si, int sj, uint64_t *seq_i, uint64_t *seq_j, uint32_t *mask, int len, char *buf,
                          long cap) {
	char *text = 0;
	size_t size = 0;
	FILE *f = open_memstream(&text, &size);
	uint64_t r;
	if(pair) r = fsacmpairint(f, si, sj, (long unsigned *) seq_i, (long unsigned *) seq_j, mask, len);
	else r = fsacmprint(f, si, sj, (long unsigned *) seq_i, (long unsigned *) seq_j, mask, len);
	fclose(f);
	if((long) size >= cap) size = (size_t) cap - 1;
	memcpy(buf, text, size);
	buf[size] = 0;
	free(text);
	return r;
}

/* The drop-in boundary itself: fsaCmpThreadOut (fsacmpthrd.c:76), called the
 * way cdist.c:181 / :184 call it.  seqs is n x wstride u64, masks n x wstride
 * u32 (pair) or 1 x wstride (global).  D / N receive Dn(Dn-1)/2 packed cells
 * of elem_size bytes.  Returns Dn. */
int refshim_fsa_cmp(int tnum, int pair, int n, int len, uint64_t *seqs,
                    long wstride, unsigned char *include, uint32_t *masks,
                    unsigned norm, unsigned minLength, double minCov,
                    unsigned proxi, int elem_size, double byteScale,
                    void *D_out, void *N_out) {
	long unsigned **seqp = malloc((size_t) n * sizeof(*seqp));
	unsigned **incp = malloc((size_t) n * sizeof(*incp));
	Matrix *D, *N;
	size_t cells;
	int i, Dn;

	for(i = 0; i < n; ++i) {
		seqp[i] = (long unsigned *) (seqs + (size_t) i * wstride);
		incp[i] = masks + (pair ? (size_t) i * wstride : 0);
	}
	ByteScale = byteScale;
	ltdMatrixInit(-elem_size);
	D = ltdMatrixInit(n);
	N = (pair && N_out) ? ltdMatrixInit(n) : 0;
	fsaCmpThreadOut(tnum, pair ? &cmpairFsaThrd : &cmpFsaThrd, D, N, n, len,
	                seqp, include, incp, norm, minLength, minCov, 0, 0, 0, 0, proxi);
	Dn = D->n;
	cells = Dn > 1 ? (size_t) Dn * (Dn - 1) / 2 : 0;
	if(cells) {
		void *src = D->mat ? (void *) *D->mat : D->fmat ? (void *) *D->fmat
		          : D->smat ? (void *) *D->smat : (void *) *D->bmat;
		memcpy(D_out, src, cells * elem_size);
		if(N) {
			src = N->mat ? (void *) *N->mat : N->fmat ? (void *) *N->fmat
			    : N->smat ? (void *) *N->smat : (void *) *N->bmat;
			memcpy(N_out, src, cells * elem_size);
		}
	}
	Matrix_destroy(D);
	if(N) Matrix_destroy(N);
	ltdMatrixInit(-(int) sizeof(double));
	ByteScale = 1.0;
	free(seqp);
	free(incp);
	return Dn;
}
