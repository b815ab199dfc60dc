/*
 * mock_ccg.c -- TEST INFRASTRUCTURE: a CPU stand-in for libccphylo_gpu.so built on the oracle (oracle/fsa_oracle.c,
 * oracle/mat_oracle.c), so that the HOST driver (ccphylo_b200/host/ *.c: option scanner, FASTA / .mat readers, parser
 * pool, gates, messages, Phylip writer, file-backed matrices) can run where there is no GPU -- under
 * -fsanitize=address,undefined, against the reference binary, on random command lines (tests/test_host_driver_cpu.py).
 * It is never built into the product and never travels as a library: the driver binary linked with it is called
 * ccphylo-b200-mock and lives in the test's temporary directory.
 *
 * Covered: `dist` on FASTA input -- pair mode with -P (per-sample builder, per-pair maskProxi), -y, -V; shared-mask mode
 * with -P or -y (not both); `dist` on .mat input (every -d method); -a; `trim` (orc_trim_pass).  Everything else
 * (shared-mask -P with -y, device pointers, groups) answers CCG_ERR_UNSUPPORTED.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ccphylo_gpu.h"
#include "fsa_oracle.h"

double orc_mat_pair(int method, unsigned order, double alpha, const uint16_t *ci, const uint32_t *ti, int len_i, const uint16_t *cj,
                    const uint32_t *tj, int len_j, unsigned norm, unsigned minDepth, unsigned minLength, double minCov, unsigned *rows_inc);
int orc_mat_matrix(int method, unsigned order, double alpha, int n, long lmax, const uint16_t *counts, const uint32_t *totals,
                   const int *lens, const unsigned char *include, unsigned norm, unsigned minDepth, unsigned minLength,
                   double minCov, double *D, double *N);

struct ccg_ctx {
	int n, len, words, pair;
	unsigned proxi;
	int snp_only;
	uint64_t *seqs;              /* n x words */
	uint32_t *masks;             /* n x words */
	unsigned char **codes;       /* per slot, kept for the proximity builder */
	unsigned char *present;
	uint32_t *gmask;
	int have_gmask;
	int nmotifs;
	int *mlens;
	unsigned char *msets;
	/* count matrices */
	int mat_n, mat_max;
	uint16_t *mat_counts;
	uint32_t *mat_totals;
	int *mat_lens;
	unsigned char *mat_present;
	/* trim */
	int t_len, t_words, t_has_ref;
	unsigned t_proxi;
	uint32_t *t_mask, *t_cols;
	unsigned char *t_cur, *t_ref;
	char err[256];
};

static char g_err[256] = "";

const char *ccg_strerror(int code) {
	switch(code) {
		case CCG_OK: return "ok";
		case CCG_ERR_ARG: return "invalid argument";
		case CCG_ERR_NOMEM: return "out of memory";
		case CCG_ERR_UNSUPPORTED: return "unsupported by the CPU mock";
		default: return "error";
	}
}
const char *ccg_last_error(const ccg_ctx *ctx) { return ctx ? ctx->err : g_err; }

static int unsupported(ccg_ctx *ctx, const char *what) {
	snprintf(ctx ? ctx->err : g_err, 256, "mock: %s", what);
	return CCG_ERR_UNSUPPORTED;
}

int ccg_init(ccg_ctx **out, int device) {
	*out = calloc(1, sizeof(ccg_ctx));
	return *out ? CCG_OK : CCG_ERR_NOMEM;
}
int ccg_init_multi(ccg_ctx **out, int ngpus) { return ccg_init(out, 0); }
int ccg_multi_gpus(const ccg_ctx *ctx, int *active) { if(active) *active = 1; return 1; }
const char *ccg_last_kernel(const ccg_ctx *ctx) { return "mock (oracle on the CPU)"; }
int ccg_sync(ccg_ctx *ctx) { return CCG_OK; }
void *ccg_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void ccg_host_free(void *p) { free(p); }

static void free_problem(ccg_ctx *c) {
	if(c->codes) for(int i = 0; i < c->n; ++i) free(c->codes[i]);
	free(c->codes); free(c->seqs); free(c->masks); free(c->present); free(c->gmask);
	c->codes = 0; c->seqs = 0; c->masks = 0; c->present = 0; c->gmask = 0;
	c->n = 0;
	c->have_gmask = 0;
}
static void free_mat(ccg_ctx *c) {
	free(c->mat_counts); free(c->mat_totals); free(c->mat_lens); free(c->mat_present);
	c->mat_counts = 0; c->mat_totals = 0; c->mat_lens = 0; c->mat_present = 0;
	c->mat_n = 0;
}
void ccg_destroy(ccg_ctx *c) {
	if(!c) return;
	free_problem(c);
	free_mat(c);
	free(c->t_mask); free(c->t_cols); free(c->t_cur); free(c->t_ref);
	free(c->mlens);
	free(c->msets);
	free(c);
}

int ccg_set_proximity(ccg_ctx *c, unsigned proxi, int snp_events_only) {
	c->proxi = proxi;
	c->snp_only = snp_events_only != 0;
	return CCG_OK;
}

int ccg_set_motifs(ccg_ctx *c, int nmotifs, const int *lens, const unsigned char *sets) {
	free(c->mlens); free(c->msets);
	c->mlens = 0; c->msets = 0;
	c->nmotifs = nmotifs;
	if(!nmotifs) return CCG_OK;
	int total = 0;
	for(int m = 0; m < nmotifs; ++m) {
		if(lens[m] < 1 || lens[m] > 32) return CCG_ERR_UNSUPPORTED;
		total += lens[m];
	}
	c->mlens = malloc((size_t) nmotifs * sizeof(int));
	c->msets = malloc((size_t) total);
	if(!c->mlens || !c->msets) return CCG_ERR_NOMEM;
	memcpy(c->mlens, lens, (size_t) nmotifs * sizeof(int));
	memcpy(c->msets, sets, (size_t) total);
	return CCG_OK;
}

int ccg_set_problem(ccg_ctx *c, int n, int len, int pair_mode) {
	if(n < 0 || len < 0) return CCG_ERR_ARG;
	free_problem(c);
	c->n = n;
	c->len = len;
	c->words = orc_words(len) > 0 ? orc_words(len) : 1;
	c->pair = pair_mode != 0;
	c->seqs = calloc((size_t) (n ? n : 1) * c->words, 8);
	c->masks = calloc((size_t) (n ? n : 1) * c->words, 4);
	c->codes = calloc((size_t) (n ? n : 1), sizeof(*c->codes));
	c->present = calloc((size_t) (n ? n : 1), 1);
	c->gmask = calloc((size_t) c->words, 4);
	return (c->seqs && c->masks && c->codes && c->present && c->gmask) ? CCG_OK : CCG_ERR_NOMEM;
}

int ccg_put_sample_codes(ccg_ctx *c, int idx, const unsigned char *codes) {
	if(idx < 0 || idx >= c->n || !c->pair) return CCG_ERR_ARG;
	free(c->codes[idx]);
	c->codes[idx] = malloc((size_t) c->len + 1);
	if(!c->codes[idx]) return CCG_ERR_NOMEM;
	memcpy(c->codes[idx], codes, (size_t) c->len);
	memset(c->seqs + (size_t) idx * c->words, 0, (size_t) c->words * 8);
	memset(c->masks + (size_t) idx * c->words, 0, (size_t) c->words * 4);
	orc_pack(codes, c->len, c->seqs + (size_t) idx * c->words);
	orc_known_mask(codes, c->len, c->masks + (size_t) idx * c->words);
	c->present[idx] = 1;
	return CCG_OK;
}

/* packed rows as the reference holds them (the integration stub's way in); no code bytes: what needs them
 * (ccg_sample_proximity, shared-mask -P) is not available on such a store */
int ccg_put_samples_packed(ccg_ctx *c, int first, int count, const uint64_t *const *seqs, const uint32_t *const *includes) {
	for(int k = 0; k < count; ++k) {
		const int s = first + k;
		if(s < 0 || s >= c->n) return CCG_ERR_ARG;
		if(!seqs[k] || (c->pair && (!includes || !includes[k]))) { c->present[s] = 0; continue; }
		memcpy(c->seqs + (size_t) s * c->words, seqs[k], (size_t) orc_words(c->len) * 8);
		if(c->pair) memcpy(c->masks + (size_t) s * c->words, includes[k], (size_t) orc_words(c->len) * 4);
		c->present[s] = 1;
	}
	return CCG_OK;
}

/* the one-call drop-in with fsaCmpThreadOut's argument list */
int ccg_fsa_cmp_thread_out(ccg_ctx *c, int pair, void *D, void *N, int elem_size, double byteScale, int n, int len,
                           const uint64_t *const *seqs, const unsigned char *include, const uint32_t *const *includes, unsigned norm,
                           unsigned minLength, double minCov, unsigned proxi, int *Dn, unsigned *global_inc) {
	int rc = ccg_set_proximity(c, proxi, 0);
	if(!rc) rc = ccg_set_problem(c, n, len, 1);
	if(rc) return rc;
	const uint64_t **s = malloc((size_t) (n ? n : 1) * sizeof(*s));
	const uint32_t **m = malloc((size_t) (n ? n : 1) * sizeof(*m));
	if(!s || !m) { free(s); free(m); return CCG_ERR_NOMEM; }
	for(int i = 0; i < n; ++i) {
		const int in = !include || include[i];
		s[i] = in ? seqs[i] : 0;
		m[i] = in ? (pair ? includes[i] : includes[0]) : 0;
	}
	rc = ccg_put_samples_packed(c, 0, n, s, m);
	free(s);
	free(m);
	if(rc) return rc;
	if(pair) return ccg_run_pair(c, include, norm, minLength, minCov, elem_size, byteScale, D, N, Dn);
	memcpy(c->gmask, includes[0], (size_t) orc_words(len) * 4);
	c->have_gmask = 1;
	return ccg_run_global(c, include, norm, elem_size, byteScale, D, Dn, global_inc);
}

int ccg_sample_proximity(ccg_ctx *c, int first, int count, int apply, unsigned *inc_out) {
	for(int k = 0; k < count; ++k) {
		const int s = first + k;
		if(s < 0 || s >= c->n || !c->present[s] || !c->codes[s]) return CCG_ERR_ARG;
		uint32_t *m = c->masks + (size_t) s * c->words, *tmp = 0;
		if(!apply) {
			tmp = malloc((size_t) c->words * 4);
			if(!tmp) return CCG_ERR_NOMEM;
			memcpy(tmp, m, (size_t) c->words * 4);
			m = tmp;
		}
		orc_inc_pos(m, c->codes[s], c->codes[s], c->len, c->proxi, c->snp_only);
		if(inc_out) inc_out[k] = (unsigned) orc_mask_count(m, c->len);
		free(tmp);
	}
	return CCG_OK;
}

int ccg_sample_count_masked(ccg_ctx *c, int slot, unsigned *inc_out) { return unsupported(c, "shared-mask mode with -y and -P"); }

int ccg_mask_motifs(ccg_ctx *c, int first, int count, unsigned *inc_out) {
	for(int k = 0; k < count; ++k) {
		const int s = first + k;
		if(s < 0 || s >= c->n || !c->present[s]) return CCG_ERR_ARG;
		uint32_t *m = c->masks + (size_t) s * c->words;
		if(c->nmotifs) orc_mask_motifs(c->seqs + (size_t) s * c->words, m, c->len, c->nmotifs, c->mlens, c->msets);
		if(inc_out) inc_out[k] = (unsigned) orc_mask_count(m, c->len);
	}
	return CCG_OK;
}

int ccg_build_global_mask(ccg_ctx *c, const unsigned char *include, unsigned *global_inc) {
	if(c->proxi && c->nmotifs) return unsupported(c, "shared-mask mode with -P and -y");
	int any = 0, ref = -1;
	if(c->proxi) {
		/* cdist.c:86-112 with -P: initIncPos, then getIncPosPtr(includes[0], seq, ref, proxi) for every included sample,
		 * ref = the first included one (the proximity events are defined on the sequences, not on the samples' masks) */
		for(int p = 0; p < c->words * 32; ++p) {
			if(p % 32 == 0) c->gmask[p / 32] = 0;
			if(p < c->len) c->gmask[p / 32] |= 1u << (31 - p % 32);
		}
		for(int i = 0; i < c->n; ++i) {
			if(!c->present[i] || (include && !include[i])) continue;
			if(ref < 0) ref = i;
			any = 1;
			orc_inc_pos(c->gmask, c->codes[i], c->codes[ref], c->len, c->proxi, c->snp_only);
		}
	} else {
		for(int w = 0; w < c->words; ++w) c->gmask[w] = 0xFFFFFFFFu;
		for(int i = 0; i < c->n; ++i) {
			if(!c->present[i] || (include && !include[i])) continue;
			any = 1;
			for(int w = 0; w < c->words; ++w) c->gmask[w] &= c->masks[(size_t) i * c->words + w];
		}
	}
	if(!any) memset(c->gmask, 0, (size_t) c->words * 4);
	c->have_gmask = 1;
	if(global_inc) *global_inc = (unsigned) orc_mask_count(c->gmask, c->len);
	return CCG_OK;
}

/* included = uploaded and not excluded by the caller */
static unsigned char *effective_include(const ccg_ctx *c, const unsigned char *include) {
	unsigned char *e = malloc((size_t) (c->n ? c->n : 1));
	if(!e) return 0;
	for(int i = 0; i < c->n; ++i) e[i] = (unsigned char) (c->present[i] && (!include || include[i]));
	return e;
}

int ccg_run_pair(ccg_ctx *c, const unsigned char *include, unsigned norm, unsigned minLength, double minCov, int elem_size,
                 double byteScale, void *D, void *N, int *Dn) {
	if(!c->pair) return CCG_ERR_ARG;
	unsigned char *e = effective_include(c, include);
	if(!e) return CCG_ERR_NOMEM;
	int dn = 0;
	for(int i = 0; i < c->n; ++i) dn += e[i];
	void *Ntmp = N ? 0 : malloc(((size_t) dn * (dn > 0 ? dn - 1 : 0) / 2 + 1) * (size_t) elem_size);
	if(c->proxi)
		dn = orc_fsa_cmp_pair_proxi(c->n, c->len, c->seqs, c->words, e, c->masks, norm, minLength, minCov, c->proxi, elem_size, byteScale, D,
		                            N ? N : Ntmp);
	else dn = orc_fsa_cmp_pair(c->n, c->len, c->seqs, c->words, e, c->masks, norm, minLength, minCov, elem_size, byteScale, D, N ? N : Ntmp);
	free(Ntmp);
	free(e);
	if(Dn) *Dn = dn;
	return CCG_OK;
}

int ccg_run_global(ccg_ctx *c, const unsigned char *include, unsigned norm, int elem_size, double byteScale, void *D, int *Dn,
                   unsigned *global_inc) {
	if(!c->have_gmask) return CCG_ERR_ARG;
	unsigned char *e = effective_include(c, include);
	if(!e) return CCG_ERR_NOMEM;
	unsigned ginc = 0;
	const int dn = orc_fsa_cmp_global(c->n, c->len, c->seqs, c->words, e, c->gmask, norm, elem_size, byteScale, D, &ginc);
	free(e);
	if(Dn) *Dn = dn;
	if(global_inc) *global_inc = ginc;
	return CCG_OK;
}

int ccg_list_variants(ccg_ctx *c, int pair, const unsigned char *include, ccg_variant_fn fn, void *user) {
	uint32_t *mask = malloc((size_t) c->words * 4);
	uint64_t *out = malloc(((size_t) c->len + 1) * 8);
	if(!mask || !out) { free(mask); free(out); return CCG_ERR_NOMEM; }
	int stop = 0;
	for(int i = 0; i < c->n && !stop; ++i) {
		if(!c->present[i] || (include && !include[i])) continue;
		for(int j = 0; j < i && !stop; ++j) {
			if(!c->present[j] || (include && !include[j])) continue;
			const uint64_t *si = c->seqs + (size_t) i * c->words, *sj = c->seqs + (size_t) j * c->words;
			const uint32_t *mi = c->masks + (size_t) i * c->words, *mj = c->masks + (size_t) j * c->words;
			if(!pair) memcpy(mask, c->gmask, (size_t) c->words * 4);
			else if(c->proxi) orc_mask_proxi(si, sj, mi, mj, c->len, c->proxi, mask);
			else for(int w = 0; w < c->words; ++w) mask[w] = mi[w] & mj[w];
			const long cnt = orc_list_variants(si, sj, mask, c->len, out, (long) c->len + 1);
			if(cnt > 0 && fn(user, i, j, out, (size_t) cnt)) stop = 1;
		}
	}
	free(mask);
	free(out);
	return CCG_OK;
}

/* ---- count matrices ---- */
int ccg_mat_set_problem(ccg_ctx *c, int n, int max_len) {
	free_mat(c);
	c->mat_n = n;
	c->mat_max = max_len > 0 ? max_len : 1;
	c->mat_counts = calloc((size_t) (n ? n : 1) * c->mat_max * 6, 2);
	c->mat_totals = calloc((size_t) (n ? n : 1) * c->mat_max, 4);
	c->mat_lens = calloc((size_t) (n ? n : 1), sizeof(int));
	c->mat_present = calloc((size_t) (n ? n : 1), 1);
	return (c->mat_counts && c->mat_totals && c->mat_lens && c->mat_present) ? CCG_OK : CCG_ERR_NOMEM;
}

int ccg_mat_put_sample(ccg_ctx *c, int idx, const uint16_t *counts6, const uint32_t *totals, int len) {
	if(idx < 0 || idx >= c->mat_n || len < 0 || len > c->mat_max) return CCG_ERR_ARG;
	memcpy(c->mat_counts + (size_t) idx * c->mat_max * 6, counts6, (size_t) len * 12);
	for(int p = 0; p < len; ++p) {
		uint32_t t = 0;
		if(totals) t = totals[p];
		else for(int k = 0; k < 6; ++k) t += counts6[(size_t) p * 6 + k];
		c->mat_totals[(size_t) idx * c->mat_max + p] = t;
	}
	c->mat_lens[idx] = len;
	c->mat_present[idx] = 1;
	return CCG_OK;
}

static void store_cell(void *buf, size_t k, int elem_size, double byteScale, double v) {
	if(elem_size == 8) ((double *) buf)[k] = v;
	else if(elem_size == 4) ((float *) buf)[k] = (float) v;
	else if(elem_size == 2) ((uint16_t *) buf)[k] = (uint16_t) (v * byteScale + 0.5);      /* dtouc, bytescale.h:22 */
	else ((uint8_t *) buf)[k] = (uint8_t) (v * byteScale + 0.5);
}

int ccg_mat_run(ccg_ctx *c, const unsigned char *include, int method, unsigned order, double alpha, unsigned norm, unsigned minDepth,
                unsigned minLength, double minCov, int elem_size, double byteScale, void *D, void *N, int *Dn, uint32_t *rows_inc) {
	const int n = c->mat_n;
	unsigned char *e = malloc((size_t) (n ? n : 1));
	if(!e) return CCG_ERR_NOMEM;
	int dn = 0;
	for(int i = 0; i < n; ++i) dn += (e[i] = (unsigned char) (c->mat_present[i] && (!include || include[i])));
	const size_t cells = (size_t) dn * (dn > 0 ? dn - 1 : 0) / 2;
	double *d = malloc((cells + 1) * 8), *nn = malloc((cells + 1) * 8);
	if(!d || !nn) { free(e); free(d); free(nn); return CCG_ERR_NOMEM; }
	const int got = orc_mat_matrix(method, order, alpha, n, c->mat_max, c->mat_counts, c->mat_totals, c->mat_lens, e, norm, minDepth, minLength,
	                               minCov, d, nn);
	free(e);
	if(got < 0) { free(d); free(nn); return unsupported(c, "a pair whose sample fails its own gate (the reference exits there)"); }
	for(size_t k = 0; k < cells; ++k) {
		store_cell(D, k, elem_size, byteScale, d[k]);
		if(N) store_cell(N, k, elem_size, byteScale, nn[k]);
		if(rows_inc) rows_inc[k] = (uint32_t) nn[k];
	}
	free(d);
	free(nn);
	if(Dn) *Dn = got;
	return CCG_OK;
}

/* ---- -a: one row against the slots below it (cmpFsaRowThrd fsacmpthrd.c:482-580, as oracle.fsa_cmp_row restates it): the pair's
 * mask is the new sample's own mask after its builder, put through the per-sample builder against the column sample ---- */
static uint32_t *row_pair_mask(const ccg_ctx *c, int row, int j, const uint32_t *own) {
	uint32_t *m = malloc((size_t) c->words * 4);
	if(!m) return 0;
	if(c->proxi) {
		memcpy(m, own, (size_t) c->words * 4);
		orc_inc_pos(m, c->codes[j], c->codes[row], c->len, c->proxi, c->snp_only);
	} else
		for(int w = 0; w < c->words; ++w) m[w] = own[w] & c->masks[(size_t) j * c->words + w];
	return m;
}

static uint32_t *row_own_mask(const ccg_ctx *c, int row) {
	uint32_t *own = malloc((size_t) c->words * 4);
	if(!own) return 0;
	memcpy(own, c->masks + (size_t) row * c->words, (size_t) c->words * 4);
	if(c->proxi) orc_inc_pos(own, c->codes[row], c->codes[row], c->len, c->proxi, c->snp_only);
	return own;
}

int ccg_run_row(ccg_ctx *c, int row_slot, unsigned norm, unsigned minLength, double minCov, double *D, double *N, int *cols) {
	if(row_slot < 0 || row_slot >= c->n || !c->present[row_slot]) return CCG_ERR_ARG;
	if(minLength < minCov * c->len) minLength = (unsigned) (minCov * c->len);
	uint32_t *own = row_own_mask(c, row_slot), *ones = malloc((size_t) c->words * 4);
	if(!own || !ones) { free(own); free(ones); return CCG_ERR_NOMEM; }
	memset(ones, 0xFF, (size_t) c->words * 4);
	int k = 0;
	for(int j = 0; j < row_slot; ++j) {
		if(!c->present[j]) continue;
		uint32_t *m = row_pair_mask(c, row_slot, j, own), mism = 0, inc = 0;
		if(!m) { free(own); free(ones); return CCG_ERR_NOMEM; }
		orc_pair_counts(c->seqs + (size_t) row_slot * c->words, c->seqs + (size_t) j * c->words, m, ones, c->len, &mism, &inc);
		free(m);
		if(minLength <= inc) {
			D[k] = norm ? (double) mism * (double) norm / (double) inc : (double) mism;
			if(N) N[k] = inc;
		} else {
			D[k] = -1.0;
			if(N) N[k] = 0.0;
		}
		++k;
	}
	free(own);
	free(ones);
	if(cols) *cols = k;
	return CCG_OK;
}

int ccg_list_variants_row(ccg_ctx *c, int row_slot, ccg_variant_fn fn, void *user) {
	if(row_slot < 0 || row_slot >= c->n || !c->present[row_slot]) return CCG_ERR_ARG;
	uint32_t *own = row_own_mask(c, row_slot);
	uint64_t *out = malloc(((size_t) c->len + 1) * 8);
	if(!own || !out) { free(own); free(out); return CCG_ERR_NOMEM; }
	for(int j = 0; j < row_slot; ++j) {
		if(!c->present[j]) continue;
		uint32_t *m = row_pair_mask(c, row_slot, j, own);
		if(!m) break;
		const long cnt = orc_list_variants(c->seqs + (size_t) row_slot * c->words, c->seqs + (size_t) j * c->words, m, c->len, out,
		                                   (long) c->len + 1);
		free(m);
		if(cnt > 0 && fn(user, row_slot, j, out, (size_t) cnt)) break;
	}
	free(own);
	free(out);
	return CCG_OK;
}

/* ---- what the mock does not stand in for ---- */
/* -a on .mat input (cmpMatRowThrd ltdmatrixthrd.c:111-181): the new sample is the loaded one, every column sample is streamed */
int ccg_mat_run_row(ccg_ctx *c, int row_slot, int method, unsigned order, double alpha, unsigned norm, unsigned minDepth,
                    unsigned minLength, double minCov, double *D, double *N, uint32_t *rows_inc) {
	if(row_slot < 0 || row_slot >= c->mat_n || !c->mat_present[row_slot]) return CCG_ERR_ARG;
	const size_t L = (size_t) c->mat_max;
	for(int j = 0; j < row_slot; ++j) {
		unsigned rows = 0;
		double v = -1.0;
		if(c->mat_present[j])
			v = orc_mat_pair(method, order, alpha, c->mat_counts + 6 * L * row_slot, c->mat_totals + L * row_slot, c->mat_lens[row_slot],
			                 c->mat_counts + 6 * L * j, c->mat_totals + L * j, c->mat_lens[j], norm, minDepth, minLength, minCov, &rows);
		if(v == -2.0) return unsupported(c, "a column sample that fails its own gate (the caller checks that first)");
		if(v == -1.0) rows = 0;
		D[j] = v;
		if(N) N[j] = rows;
		if(rows_inc) rows_inc[j] = rows;
	}
	return CCG_OK;
}
/* ---- trim: orc_trim_pass behind the ccg_trim_* calls (see include/ccphylo_gpu.h for the order of the calls) ---- */
static void trim_free(ccg_ctx *c) {
	free(c->t_mask); free(c->t_cols); free(c->t_cur); free(c->t_ref);
	c->t_mask = c->t_cols = 0;
	c->t_cur = c->t_ref = 0;
	c->t_has_ref = 0;
}

int ccg_trim_begin(ccg_ctx *c, int len, unsigned proxi) {
	trim_free(c);
	c->t_len = len;
	c->t_words = orc_words(len) > 0 ? orc_words(len) : 1;
	c->t_proxi = proxi;
	c->t_mask = calloc((size_t) c->t_words, 4);
	c->t_cols = calloc((size_t) c->t_words, 4);
	c->t_cur = calloc((size_t) len + 1, 1);
	c->t_ref = calloc((size_t) len + 1, 1);
	return (c->t_mask && c->t_cols && c->t_cur && c->t_ref) ? CCG_OK : CCG_ERR_NOMEM;
}

int ccg_trim_sample(ccg_ctx *c, const unsigned char *codes, const uint64_t *nibbles, int against_ref, int builder, unsigned *inc_out) {
	if(!c->t_mask || (against_ref && !c->t_has_ref) || builder < 0 || builder > 2) return CCG_ERR_ARG;
	const int len = c->t_len, W = c->t_words;
	if(len) memcpy(c->t_cur, codes, (size_t) len);
	if(!against_ref)
		for(int p = 0; p < W * 32; ++p) {                       /* initIncPos (fsacmp.c:164) */
			if(p % 32 == 0) c->t_mask[p / 32] = 0;
			if(p < len) c->t_mask[p / 32] |= 1u << (31 - p % 32);
		}
	orc_trim_pass(c->t_mask, c->t_cur, against_ref ? c->t_ref : 0, len, c->t_proxi, builder, against_ref ? c->t_cols : 0);
	if(c->nmotifs && len) {
		if(!nibbles) return CCG_ERR_ARG;
		orc_mask_motifs(nibbles, c->t_mask, len, c->nmotifs, c->mlens, c->msets);
	}
	if(inc_out) *inc_out = (unsigned) orc_mask_count(c->t_mask, len);
	return CCG_OK;
}

int ccg_trim_keep_reference(ccg_ctx *c) {
	if(!c->t_mask) return CCG_ERR_ARG;
	for(int p = 0; p < c->t_len; ++p) c->t_ref[p] = (unsigned char) (c->t_cur[p] & 15u);
	c->t_has_ref = 1;
	return CCG_OK;
}

int ccg_trim_get_mask(ccg_ctx *c, int variable_columns_only, uint32_t *mask_out, unsigned *inc_out, unsigned *var_out) {
	if(!c->t_mask) return CCG_ERR_ARG;
	unsigned var = 0;
	for(int w = 0; w < orc_words(c->t_len); ++w) {
		var += (unsigned) __builtin_popcount(c->t_mask[w] & c->t_cols[w]);
		if(mask_out) mask_out[w] = variable_columns_only ? (c->t_mask[w] & c->t_cols[w]) : c->t_mask[w];
	}
	if(inc_out) *inc_out = (unsigned) orc_mask_count(c->t_mask, c->t_len);
	if(var_out) *var_out = var;
	return CCG_OK;
}

int ccg_trim_end(ccg_ctx *c) {
	trim_free(c);
	return CCG_OK;
}
