/*
 * fsa_oracle.h -- CPU restatement of ccphylo's `dist` FASTA hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product path
 * (ccphylo_b200/, include/) may link, import or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker.
 *
 * Parity status: PINNED BY EXECUTION.  The reference has no tests of its own
 * (SURVEY.md section 4); this restatement is pinned against the unmodified
 * reference compiled from /root/reference into oracle/_ref/ (see
 * oracle/Makefile and tests/test_oracle_vs_reference.py) and against the
 * golden vectors in tests/golden/ that were produced by that binary.
 *
 * Every function cites the reference file:line whose behaviour it restates
 * (paths relative to /root/reference).
 */
#ifndef FSA_ORACLE_H
#define FSA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* fsacmp.c:32-91 get2BitTable -- byte -> code (0..3 base, 4 unknown, 32 drop). */
void orc_code_table(unsigned flag, unsigned char table[256]);

/* seqparse.c:195-248 FileBuffgetFsaSeq -- translate sequence bytes, keeping
 * only codes < 32.  Returns the number of codes written. */
long orc_translate(const unsigned char *bytes, long nbytes, unsigned flag,
                   unsigned char *codes);

/* qseqs.c:60-88 qseq2nibble -- 32 codes per u64, first base in the top bits,
 * code 4 stored as 00, tail word left-aligned.  Returns #unknown. */
int orc_pack(const unsigned char *codes, int len, uint64_t *words);

/* fsacmp.c:164-179 initIncPos + fsacmp.c:181-238 getIncPos(seq, seq, 0):
 * bit (31 - p%32) of word p/32 is set iff code[p] != 4.  Returns popcount
 * (fsacmp.c:487-503 getNpos). */
int orc_known_mask(const unsigned char *codes, int len, uint32_t *mask);

/* fsacmp.c:181-238 getIncPos(include, seq, ref, 0) on an existing mask:
 * clears bits where seq or ref is unknown (global-mode accumulation,
 * cdist.c:110-111). */
void orc_and_known(uint32_t *mask, const unsigned char *seq_codes,
                   const unsigned char *ref_codes, int len);

/* fsacmp.c:487-503 getNpos. */
int orc_mask_count(const uint32_t *mask, int len);

/* fsacmp.c:355-389 maskProxi (proxi == 0) + fsacmp.c:587-633 fsacmpair:
 * n = popcount(inc_i & inc_j); mism = #2-bit lanes that differ under it. */
void orc_pair_counts(const uint64_t *seq_i, const uint64_t *seq_j,
                     const uint32_t *inc_i, const uint32_t *inc_j, int len,
                     uint32_t *mism, uint32_t *ninc);

/* fsacmp.c:181-353 getIncPos (variant 0) / getIncPosInsigPrune, getIncPosInsig (variant 1) with -P
 * proxi on an existing mask: clears unknown positions and everything between two events at most
 * proxi apart (see fsa_oracle.c for what counts as an event). */
void orc_inc_pos(uint32_t *mask, const unsigned char *seq_codes,
                 const unsigned char *ref_codes, int len, unsigned proxi, int variant);

/* fsacmp.c:355-485 maskProxi with -P proxi: the pair's mask (ceil(len/32) words) */
void orc_mask_proxi(const uint64_t *seq_i, const uint64_t *seq_j, const uint32_t *inc_i, const uint32_t *inc_j, int len,
                    unsigned proxi, uint32_t *mask_out);
/* fsacmp.c:355-485 maskProxi with -P proxi + fsacmp.c:587-633 fsacmpair. */
void orc_pair_counts_proxi(const uint64_t *seq_i, const uint64_t *seq_j,
                           const uint32_t *inc_i, const uint32_t *inc_j, int len,
                           unsigned proxi, uint32_t *mism, uint32_t *ninc);

/* `ccphylo trim` (fsaTrim trim.c:77-260): one sample's pass over an inclusion mask on trim's alphabet (getIupacBitTable
 * fsacmp.c:93-162: 0-3 bases, 4 unknown, 5 gap, 6-15 ambiguity letters, +16 = soft-masked input).  ref == NULL: the sample
 * against itself, getIncPos(includes, seq, seq, proxi) (trim.c:183,201).  Otherwise builder 0 = getIncPos (fsacmp.c:181),
 * 1 = getIncPosInsig (:297, flag 8), 2 = getIncPosInsigPrune (:240, flag 32) against the reference sample's stored bytes.
 * Clears unknown / soft-masked positions as the builder does, then everything from an event to the next one at most proxi
 * later (both inclusive); strips the soft flag of `seq` where the reference strips it (the bytes trim prints and
 * pseudoAlnPrune fsacmp.c:504 compares); ORs into `columns` (may be NULL) the positions where the stored byte differs
 * from the reference sample's. */
void orc_trim_pass(uint32_t *mask, unsigned char *seq, const unsigned char *ref, int len, unsigned proxi, int builder,
                   uint32_t *columns);

/* meth.c:141-159 maskMotifs (-y) with the motif list of methparse.c:268 getMethMotifs: nmotifs motifs (the file's
 * motifs and their reverse complements), motif m has lens[m] <= 32 positions whose codes follow each other in
 * `sets`: bits 0..3 = the bases A, C, G, T the position accepts, bit 4 = methylation site.  Wherever a motif
 * matches the PACKED sequence (unknown bases read as A, qseqs.c:60) the mask bits of its methylation sites are
 * cleared.  Returns the number of matches (meth.c:153). */
long orc_mask_motifs(const uint64_t *seq, uint32_t *mask, int len, int nmotifs, const int *lens,
                     const unsigned char *sets);

/* fsacmp.c:646-683 fsacmprint / :685-737 fsacmpairint (-V): the variants of one pair under `mask` (pair mode: the
 * pair's mask; shared-mask mode: the global mask) in the order and with the position labels the reference
 * prints.  out[k] = label << 4 | code_i << 2 | code_j; returns how many there are (out may be NULL / cap 0). */
long orc_list_variants(const uint64_t *seq_i, const uint64_t *seq_j, const uint32_t *mask, int len,
                       uint64_t *out, long cap);

/* fsacmp.c:552-585 fsacmp: mismatches under one shared mask. */
uint32_t orc_masked_mism(const uint64_t *seq_i, const uint64_t *seq_j,
                         const uint32_t *mask, int len);

/* Words per sample: ceil(len / 32). */
static inline int orc_words(int len) { return (len >> 5) + ((len & 31) ? 1 : 0); }

/* Raw integer matrices over ALL n samples (no exclusion, no epilogue):
 * strict lower triangle, row-major, row i at i(i-1)/2.  nthreads > 1 uses
 * pthreads over rows (only used to make big test cases finish quickly). */
void orc_raw_pair_matrix(int n, int len, const uint64_t *seqs,
                         const uint32_t *masks, long wstride,
                         uint32_t *mism, uint32_t *ninc, int nthreads);

/* fsacmpthrd.c:261-480 cmpairFsaThrd -- pair mode incl. compaction of
 * excluded samples and the epilogue (:419-475) for the four cell types
 * elem_size 8 (double) / 4 (float) / 2 (u16) / 1 (u8).  seqs is n x wstride
 * u64, masks n x wstride u32.  D and N hold Dn(Dn-1)/2 cells; N may be NULL.
 * Returns Dn (number of included samples). */
int orc_fsa_cmp_pair(int n, int len, const uint64_t *seqs, long wstride,
                     const unsigned char *include, const uint32_t *masks,
                     unsigned norm, unsigned minLength, double minCov,
                     int elem_size, double byteScale, void *D, void *N);

int orc_fsa_cmp_pair_proxi(int n, int len, const uint64_t *seqs, long wstride,
                           const unsigned char *include, const uint32_t *masks,
                           unsigned norm, unsigned minLength, double minCov, unsigned proxi,
                           int elem_size, double byteScale, void *D, void *N);

/* fsacmpthrd.c:108-259 cmpFsaThrd -- global-mask mode.  mask is the single
 * shared mask (includes[0]).  Restates the INTENDED pair selection (included
 * samples only); the reference's own selection is wrong when a sample is
 * excluded, so parity with the reference is asserted
 * only on inputs without exclusions.  *global_inc gets popcount(mask). */
int orc_fsa_cmp_global(int n, int len, const uint64_t *seqs, long wstride,
                       const unsigned char *include, const uint32_t *mask,
                       unsigned norm, int elem_size, double byteScale,
                       void *D, unsigned *global_inc);

/* One cell of the pair-mode epilogue (fsacmpthrd.c:419-475). */
void orc_pair_cell(uint32_t mism, uint32_t inc, unsigned norm,
                   unsigned minLength, int elem_size, double byteScale,
                   void *Dcell, void *Ncell);

#ifdef __cplusplus
}
#endif
#endif
