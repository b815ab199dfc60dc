/*
 * motifs.h -- -y / --methylation_motifs: the motif file of the reference's getMethMotifs (methparse.c:268-296).
 * FASTA-like: an optional ">name" line, then the motif in IUPAC letters over one or more lines; lower case = plain
 * position, UPPER case = methylation site (masked at every match); '-', '.' and anything that is not an IUPAC
 * letter is dropped (getMethBitTable methparse.c:27-81).  Every motif is followed by its reverse complement.
 */
#ifndef CCB_MOTIFS_H
#define CCB_MOTIFS_H

typedef struct {
	int n;                    /* motifs including the reverse complements */
	int *lens;
	unsigned char *sets;      /* concatenated positions: bits 0..3 accepted bases A C G T, bit 4 methylation site */
} MotifList;

/* 0 on success; prints the reason and returns non-zero otherwise */
int motifs_load(const char *path, MotifList *out);
void motifs_free(MotifList *m);

#endif
