/*
 * mat_reader.h -- KMA count matrices (*.mat, *.mat.gz): "#template" line, then one row per
 * alignment column "ref\tA\tC\tG\tT\tN\t-", a blank line ends the template.  Counterpart of the
 * reference's matparse.c (FileBuffSkipTemplate :142, FileBuffLoadMat :213, FileBuffGetRow :45).
 * The whole template of a sample is read ONCE; rows whose reference base is '-' (insertions
 * relative to the template) are dropped, which is what stripMat (matcmp.c:27) intends.
 */
#ifndef CCB_MAT_READER_H
#define CCB_MAT_READER_H

#include <stddef.h>
#include <stdint.h>

typedef struct {
	uint16_t *counts;      /* len x 6: A, C, G, T, -, N (the reference's storage order, matparse.c:254-259) */
	uint32_t *totals;      /* len: sum of the six numbers as parsed */
	size_t len, cap;
	unsigned nNucs;        /* rows with minDepth <= total */
} MatSample;

void mat_sample_init(MatSample *m);
void mat_sample_free(MatSample *m);
/* 1: template found and loaded; 0: not in this file; -1: cannot open / empty (errno set or 0) */
int mat_load_template(const char *path, const char *target, unsigned minDepth, MatSample *out);
/* first byte of the (decompressed) file, -1 if unreadable */
int mat_peek(const char *path);

#endif
