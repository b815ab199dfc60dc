/* phy_update.c -- see phy_update.h */
#define _POSIX_C_SOURCE 200809L
#include "phy_update.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fsa_reader.h"

static void die_errno(void) {
	fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
	exit(errno ? errno : 1);
}

/* getSizePhy + getFilenamesPhy (phy.c:509-650): the sample count and the row names of a single-matrix Phylip file */
int read_phy_names(const char *phyname, const char *dir, char sep, PhyNames *out) {
	FILE *f = fopen(phyname, "rb");
	if(!f) {
		fprintf(stderr, "Filename:\t%s\n", phyname);
		die_errno();
	}
	ByteBuf text;
	bytebuf_init(&text, 1 << 16);
	for(;;) {
		if(text.cap - text.len < 65536) {
			text.cap <<= 1;
			text.data = realloc(text.data, text.cap);
			if(!text.data) die_errno();
		}
		size_t got = fread(text.data + text.len, 1, text.cap - text.len, f);
		if(!got) break;
		text.len += got;
	}
	fclose(f);
	const unsigned char *p = text.data, *end = text.data + text.len;
	if(p < end && *p == '#') {
		while(p < end && *p != '\n') ++p;
		if(p < end) ++p;
	}
	int n = 0;
	while(p < end && *p != '\n') {
		if('0' <= *p && *p <= '9') n = 10 * n + (*p - '0');
		++p;
	}
	if(p < end) ++p;
	out->n = n;
	out->paths = calloc((size_t) (n ? n : 1), sizeof(char *));
	if(!out->paths) die_errno();
	const size_t dlen = strlen(dir);
	for(int i = 0; i < n; ++i) {
		if(p >= end) {
			fprintf(stderr, "Malformatted phylip file, name on row: %d\n", i + 1);
			return 0;
		}
		const unsigned char *q = p;
		while(q < end && *q != (unsigned char) sep && *q != '\n') ++q;
		if(q + 1 >= end) {
			/* the reference wants at least one more byte behind the character that ends a name (phy.c:617-624): a
			 * file whose last row is a bare name -- only a one-sample matrix -- is "malformatted" to it */
			fprintf(stderr, "Malformatted phylip file, name on row: %d\n", i + 1);
			return 0;
		}
		size_t nl = (size_t) (q - p);
		while(nl && (p[nl - 1] == ' ' || (p[nl - 1] >= '\t' && p[nl - 1] <= '\r'))) --nl;     /* isspace */
		char *path = malloc(dlen + nl + 1);
		if(!path) die_errno();
		memcpy(path, dir, dlen);
		memcpy(path + dlen, p, nl);
		path[dlen + nl] = 0;
		out->paths[i] = path;
		while(q < end && *q != '\n') ++q;
		if(q >= end && i != n - 1) {
			fprintf(stderr, "Malformatted phylip file, missing newline at row:\t%d\n", i + 1);
			return 0;
		}
		p = q < end ? q + 1 : q;
	}
	const int more = p < end;
	bytebuf_free(&text);
	if(more) {
		fprintf(stderr, "Cannot update a multi distance phylip file.\n");
		return -1;
	}
	return 1;
}

/* printphyUpdate (phy.c:201-250): new count over the first ten bytes, the new row at the end */
void phy_append_row(const char *phyname, int n, char *name, const double *row, unsigned flag, int precision) {
	FILE *f = fopen(phyname, "rb+");
	if(!f) {
		fprintf(stderr, "Filename:\t%s\n", phyname);
		die_errno();
	}
	fprintf(f, "%10d", n);
	fflush(f);
	fseek(f, 0, SEEK_END);
	size_t len = strlen(name);
	if(len && ((name[0] == '"' && name[len - 1] == '"') || (name[0] == '\'' && name[len - 1] == '\''))) {
		name[len - 1] = 0;
		++name;
	}
	const char *slash = strrchr(name, '/');
	if(slash) name = (char *) slash + 1;
	if(flag & 1) fprintf(f, "%s", name);
	else fprintf(f, "%-10.10s", name);
	for(int j = 0; j + 1 < n; ++j) {
		const double d = row[j];
		if(d == (int) d) fprintf(f, "\t%d", (int) d);
		else fprintf(f, "\t%.*f", precision, d);
	}
	fprintf(f, "\n");
	fclose(f);
}

