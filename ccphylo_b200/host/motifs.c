/* motifs.c -- see motifs.h */
#include "motifs.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fsa_reader.h"

static int popcount4(unsigned s) { return (int) ((s & 1u) + ((s >> 1) & 1u) + ((s >> 2) & 1u) + ((s >> 3) & 1u)); }

/* letter -> accepted bases | 16 for upper case; 0 = not a motif letter (methparse.c:27-81) */
static unsigned letter_set(int ch) {
	static const char letters[] = "acgturyswkmbdhvxn";
	static const unsigned char set[] = {1, 2, 4, 8, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15, 15};
	const int upper = ch >= 'A' && ch <= 'Z';
	const int lc = upper ? ch - 'A' + 'a' : ch;
	const char *p = lc ? strchr(letters, lc) : 0;
	if(!p) return 0;
	return set[p - letters] | (upper ? 16u : 0u);
}

static void push_motif(MotifList *out, const unsigned char *pos, int len, size_t *cap_m, size_t *cap_s, size_t *nsets) {
	if((size_t) out->n == *cap_m) {
		*cap_m = *cap_m ? *cap_m * 2 : 16;
		out->lens = realloc(out->lens, *cap_m * sizeof(int));
	}
	if(*nsets + (size_t) len > *cap_s) {
		while(*nsets + (size_t) len > *cap_s) *cap_s = *cap_s ? *cap_s * 2 : 256;
		out->sets = realloc(out->sets, *cap_s);
	}
	if(!out->lens || !out->sets) {
		fprintf(stderr, "Error: %d (%s)\n", ENOMEM, strerror(ENOMEM));
		exit(ENOMEM);
	}
	/* qseq2methMotif (methparse.c:177-266) stores, per position, as many alternative words as the motif's most
	 * ambiguous letter has bases, and pads the alternatives of a less ambiguous position with
	 * `bases[*seq & 31]`.  For an upper-case letter that index is >= 16, past the 16-entry table: what is read
	 * there is not defined by the language.  Built with the reference's own Makefile (gcc -O3) the 32-entry
	 * `nums` table lies right behind `bases`, and the padding base is the NUMBER of bases the letter stands
	 * for: such a position also accepts base code 1 (C), 2 (G) or 3 (T).  That is what users of the reference
	 * binary get, so it is what is reproduced here by default; CCPHYLO_MOTIF_STRICT=1 keeps the letters' own sets. */
	int most = 0;
	for(int k = 0; k < len; ++k)
		if(popcount4(pos[k] & 15u) > most) most = popcount4(pos[k] & 15u);
	const int strict = getenv("CCPHYLO_MOTIF_STRICT") != 0;
	for(int k = 0; k < len; ++k) {
		unsigned v = pos[k];
		const int nb = popcount4(v & 15u);
		if(!strict && (v & 16u) && nb < most) v |= 1u << nb;
		out->sets[*nsets + (size_t) k] = (unsigned char) v;
	}
	out->lens[out->n++] = len;
	*nsets += (size_t) len;
}

static void finish_record(MotifList *out, ByteBuf *cur, size_t *cap_m, size_t *cap_s, size_t *nsets) {
	const int len = (int) cur->len;
	if(len == 0) return;             /* (an empty record sends the reference into a 2^32-step loop, methparse.c:191) */
	push_motif(out, cur->data, len, cap_m, cap_s, nsets);
	/* strrcMeth (methparse.c:83-103): swap and complement from both ends; for an odd length it then complements
	 * "the middle" through a pointer that still stands on the last left element, so that element ends up as the
	 * uncomplemented letter it was swapped with and the middle letter is never complemented */
	unsigned char *rc = malloc((size_t) len);
	if(!rc) exit(ENOMEM);
	for(int k = 0; k < len; ++k) {
		const unsigned v = cur->data[len - 1 - k], s = v & 15u;
		rc[k] = (unsigned char) (((s & 1u) << 3) | ((s & 2u) << 1) | ((s & 4u) >> 1) | ((s & 8u) >> 3) | (v & 16u));
	}
	if(len & 1) {
		const int mid = len >> 1;
		rc[mid] = cur->data[mid];
		if(mid >= 1) rc[mid - 1] = cur->data[len - mid];
	}
	push_motif(out, rc, len, cap_m, cap_s, nsets);
	free(rc);
	cur->len = 0;
}

int motifs_load(const char *path, MotifList *out) {
	memset(out, 0, sizeof(*out));
	FsaReader *fr = fsa_open(path);
	if(!fr) {
		fprintf(stderr, "Filename:\t%s\n", path);
		fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
		return 1;
	}
	ByteBuf line, cur;
	bytebuf_init(&line, 256);
	bytebuf_init(&cur, 64);
	size_t cap_m = 0, cap_s = 0, nsets = 0;
	int too_long = 0;
	while(fsa_read_line(fr, &line)) {
		size_t k = 0;
		if(line.len && line.data[0] == '>') {
			finish_record(out, &cur, &cap_m, &cap_s, &nsets);
			continue;
		}
		for(; k < line.len; ++k) {
			const unsigned v = letter_set(line.data[k]);
			if(!v) continue;
			if(cur.len == 32) { too_long = 1; continue; }
			if(cur.len == cur.cap) {
				cur.cap <<= 1;
				cur.data = realloc(cur.data, cur.cap);
				if(!cur.data) exit(ENOMEM);
			}
			cur.data[cur.len++] = (unsigned char) v;
		}
	}
	finish_record(out, &cur, &cap_m, &cap_s, &nsets);
	bytebuf_free(&line);
	bytebuf_free(&cur);
	fsa_close(fr);
	if(too_long) {
		/* the reference shifts by a negative count for those (meth.c:99) */
		fprintf(stderr, "Motifs of more than 32 positions are not supported (%s).\n", path);
		return 1;
	}
	return 0;
}

void motifs_free(MotifList *m) {
	free(m->lens);
	free(m->sets);
	memset(m, 0, sizeof(*m));
}
