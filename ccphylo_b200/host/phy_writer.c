/* phy_writer.c -- see phy_writer.h */
#include "phy_writer.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const char *base_name(char *name) {
	size_t len = strlen(name);
	/* enclosing quotes are dropped (the closing one is cut off in place, as the reference does) */
	if(len && ((name[0] == '"' && name[len - 1] == '"') || (name[0] == '\'' && name[len - 1] == '\''))) {
		name[len - 1] = 0;
		++name;
	}
	const char *slash = strrchr(name, '/');
	return slash ? slash + 1 : name;
}

static inline double cell_value(const void *cells, int elem_size, double byteScale, size_t k) {
	switch(elem_size) {
		case 8: return ((const double *) cells)[k];
		case 4: return ((const float *) cells)[k];
		case 2: return ((const uint16_t *) cells)[k] / byteScale;
		default: return ((const uint8_t *) cells)[k] / byteScale;
	}
}

void phy_write(FILE *out, const void *cells, int elem_size, double byteScale, int dn, char **names,
               const unsigned char *include, const char *comment, unsigned flags, int precision) {
	size_t k = 0;
	int row = 0;
	if(flags & 4) fprintf(out, "#%s\n", comment ? comment : "(null)");
	fprintf(out, "%10d\n", dn);
	for(int i = 0; row != dn; ++i) {
		if(include && !include[i]) continue;
		const char *name = base_name(names[i]);
		if(flags & 1) fputs(name, out);
		else fprintf(out, "%-10.10s", name);
		for(int j = 0; j < row; ++j, ++k) {
			const double d = cell_value(cells, elem_size, byteScale, k);
			/* integral test in the reference's own terms: d == (int) d */
			if(d >= -2147483648.0 && d <= 2147483647.0 && d == (double) (int) d) fprintf(out, "\t%d", (int) d);
			else fprintf(out, "\t%.*f", precision, d);
		}
		fputc('\n', out);
		++row;
	}
}
