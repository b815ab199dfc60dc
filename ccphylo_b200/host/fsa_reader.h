/*
 * fsa_reader.h -- buffered (gzip-transparent) file reader and KMA consensus FASTA parsing
 * for the `dist` driver.  Behavioural counterpart of the reference's filebuff.c:52
 * (openAndDetermine), seqparse.c:128 (FileBuffgetFsaHeader), :195 (FileBuffgetFsaSeq) and :28
 * (FileBuffgetFsa), and of the byte -> code table of fsacmp.c:32 (get2BitTable).
 */
#ifndef FSA_READER_H
#define FSA_READER_H

#include <stddef.h>

typedef struct FsaReader FsaReader;

/* growable byte buffer */
typedef struct {
	unsigned char *data;
	size_t len, cap;
} ByteBuf;

void bytebuf_init(ByteBuf *b, size_t cap);
void bytebuf_free(ByteBuf *b);

/* byte -> code table: 0..3 bases, 4 unknown, 32 = not a sequence byte (dropped).
 * flag bit 8 makes lower-case acgtu count as bases (fsacmp.c:32-91). */
void fsa_code_table(unsigned flag, unsigned char table[256]);

/* opens a plain or gzip file ("-" = stdin); NULL on failure (errno set) */
FsaReader *fsa_open(const char *path);
void fsa_close(FsaReader *r);
/* first byte of the (decompressed) stream without consuming it, or -1 (filebuff.c:26 fileExist) */
int fsa_peek(FsaReader *r);

/* advance to the next '>' and read the header line (without '>', trailing white space
 * stripped) into header; 0 at end of file (seqparse.c:128) */
int fsa_next_header(FsaReader *r, ByteBuf *header);
/* same; *offset receives the stream offset of the record's '>' */
int fsa_next_header_off(FsaReader *r, ByteBuf *header, long long *offset);
/* translate the sequence up to the next '>' or end of file, keeping codes < 32; returns 0
 * only if the stream was already exhausted (seqparse.c:195) */
int fsa_read_codes(FsaReader *r, const unsigned char table[256], ByteBuf *codes);

/* offset (in the decompressed stream) of the next byte to be read */
long long fsa_tell(const FsaReader *r);
/* 1 when the file is not compressed, i.e. fsa_seek is cheap */
int fsa_is_plain(FsaReader *r);
/* continue reading at `offset` of the (decompressed) stream; 0 on success */
int fsa_seek(FsaReader *r, long long offset);

/* the unread bytes of the reader's buffer (refilled when empty; *have = 0 at end of file), for parsers that scan whole
 * lines in place; fsa_consume marks the first n of them read.  A line cut by the end of the window is fetched whole
 * with fsa_read_line. */
const unsigned char *fsa_window(FsaReader *r, size_t *have);
void fsa_consume(FsaReader *r, size_t n);

/* next text line without its '\n' (0-terminated, len excludes the terminator); 0 at end of file */
int fsa_read_line(FsaReader *r, ByteBuf *line);

#endif
