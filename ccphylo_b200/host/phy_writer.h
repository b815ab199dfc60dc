/*
 * phy_writer.h -- Phylip text output of a packed lower-triangular matrix, byte-compatible with
 * the reference's printphy (phy.c:59-123): "%10d\n" sample count, then per included sample its
 * name (directory part and enclosing quotes stripped; "relaxed" = whole name, strict =
 * "%-10.10s") followed by tab-separated cells; a cell whose value is integral prints as "%d",
 * any other as "%.<precision>f".  Cells of 2 and 1 bytes are fixed-point (value = cell /
 * byteScale, bytescale.h:22).
 */
#ifndef CCB_PHY_WRITER_H
#define CCB_PHY_WRITER_H

#include <stdio.h>

/* flags: bit 1 relaxed names, bit 4 print "#<comment>" first (dist.c:706-718).
 * names has one entry per input sample; include (may be NULL) selects the dn samples of the
 * matrix in input order.  cells: dn(dn-1)/2 values of elem_size bytes. */
void phy_write(FILE *out, const void *cells, int elem_size, double byteScale, int dn, char **names,
               const unsigned char *include, const char *comment, unsigned flags, int precision);
/* the same text, rows formatted by `threads` host threads (blocks of rows, written in order) */
void phy_write_mt(FILE *out, const void *cells, int elem_size, double byteScale, int dn, char **names,
                  const unsigned char *include, const char *comment, unsigned flags, int precision, int threads);

/* "%.*f" of one cell into dst (room for 48 bytes) without the C library, byte-identical to snprintf; returns the length,
 * or 0 when the value is left to snprintf (non-finite, precision > 18, |d| >= 2^52, more than 19 digits) */
size_t phy_format_fixed(char *dst, double d, int precision);

#endif
