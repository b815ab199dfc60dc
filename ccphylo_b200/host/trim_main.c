/*
 * trim_main.c -- `ccphylo-b200 trim`: the host side of the reference's `ccphylo trim` (trim.c:307 main_trim, :77
 * fsaTrim, :38 printTrimFsa) in front of the CUDA library.  Same options, same FASTA text on the output, same stderr
 * lines.  The masks -- unknown / soft-masked positions, methylation sites (-y), proximity runs (-P), variable columns
 * (flag 16) -- are built on the device (ccg_trim_*, csrc/k_trim.cu); the host parses, translates with trim's own table,
 * applies the inclusion gates and prints.  No CPU fallback: without a device the program stops.
 *
 * Control flow kept from fsaTrim, quirks included (they decide what the output looks like):
 *   - the pairwise flag (2) sends EVERY sample through the "no reference yet" branch (trim.c:189-223 never sets `ref`
 *     in that mode): own mask, getIncPos whatever -f 8 / -f 32 say, no "# Included" line, printed at once;
 *   - otherwise the first sample that passes becomes the reference, later samples narrow the shared mask, and the
 *     stored sequences are printed LAST TO FIRST at the end (trim.c:249-255);
 *   - without -r (every record of the file is a sample) the name table follows the reference's indexing: a record that
 *     is read but not kept hands ITS name to the previously kept one, and leaves a hole in the slot array that the
 *     backwards print loop then runs into (trim.c:224-232, :249-255; the array is re-based whenever it doubles, :140).
 * Where the reference reads or writes out of bounds this driver does not follow: a pairwise sample below the threshold
 * is not printed (the reference dereferences the mask it has just freed, trim.c:206-209,228), soft-masked letters that
 * the reference leaves flagged (beside an unknown reference base; all of them under -f 8 without -f 1 / -f 4) print as
 * their own letter instead of a byte from beyond `bases[16]` (trim.c:40,50,60-64), and a run in which no sample
 * passes ends with "All sequences were trimmed away." instead of getNpos(NULL).
 */
#define _POSIX_C_SOURCE 200809L
#include <ctype.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ccphylo_gpu.h"
#include "cmdline.h"
#include "fsa_reader.h"
#include "motifs.h"

typedef struct {
	unsigned numFile;
	char **filenames;
	char *outputfilename, *methfilename, *targetTemplate;
	unsigned minLength, flag, proxi;
	double minCov;
} TrimOpts;

static void die_errno(void) {
	fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
	exit(errno ? errno : 1);
}

static void die_gpu(ccg_ctx *ctx, int rc) {
	fprintf(stderr, "GPU error: %s (%s)\n", ccg_strerror(rc), ccg_last_error(ctx));
	exit(rc ? rc : 1);
}

/* getIupacBitTable (fsacmp.c:93-162): flag bit 1 turns lower-case letters into unknowns, otherwise they keep their
 * letter and carry the soft flag 16 */
static void trim_iupac_table(unsigned flag, unsigned char table[256]) {
	static const char letters[] = "ACGTN-RYSWKMBDHV";
	memset(table, 32, 256);
	for(int k = 0; k < 16; ++k) table[(unsigned char) letters[k]] = (unsigned char) k;
	table['U'] = 3;
	table['X'] = 4;
	for(int k = 0; k < 16; ++k) {
		const int c = letters[k];
		if(c == '-' || c == 'N') continue;
		table[tolower(c)] = (flag & 1) ? 4 : (unsigned char) (k | 16);
	}
	table['u'] = (flag & 1) ? 4 : (3 | 16);
	table['n'] = 4;
	table['x'] = 4;
}

/* qseq2nibble (qseqs.c:60-87): 2 bits per position, the first position of a word in its top bits; an unknown adds
 * nothing, any other code is ORed in as it is -- codes above 3 reach into the bits of the positions before them */
static void pack_nibbles(const unsigned char *codes, int len, uint64_t *dest) {
	for(int i = 0; i < len; i += 32) {
		const int end = i + 32 < len ? i + 32 : len;
		uint64_t nuc = 0;
		for(int j = i; j < end; ++j) nuc = codes[j] == 4 ? nuc << 2 : (nuc << 2) | codes[j];
		dest[i >> 5] = nuc;
	}
	if(len & 31) dest[(len - 1) >> 5] <<= (64 - ((len & 31) << 1));
}

static const char *strip_dir(const char *s) {
	const char *slash = strrchr(s, '/');
	return slash ? slash + 1 : s;
}

/* printTrimFsa (trim.c:38-75) */
static void print_trim_fsa(FILE *out, const char *name, const unsigned char *codes, int len, const uint32_t *mask, unsigned flag,
                           unsigned char *line) {
	static const char bases[16] = {'A', 'C', 'G', 'T', 'N', '-', 'R', 'Y', 'S', 'W', 'K', 'M', 'B', 'D', 'H', 'V'};
	fprintf(out, ">%s\n", strip_dir(name));
	size_t k = 0;
	if((flag & 18) == 16) {
		for(int i = 0; i < len; ++i)
			if((mask[i >> 5] >> (31 - (i & 31))) & 1) line[k++] = (unsigned char) bases[codes[i] & 15];
	} else {
		for(int i = 0; i < len; ++i) {
			const char b = bases[codes[i] & 15];
			if((mask[i >> 5] >> (31 - (i & 31))) & 1) line[k++] = (unsigned char) b;
			else line[k++] = (flag & 1) ? 'N' : (unsigned char) tolower(b);
		}
	}
	line[k++] = '\n';
	fwrite(line, 1, k, out);
}

static void fsa_trim(const TrimOpts *o) {
	const int pair = (o->flag & 2) != 0;
	const int builder = (o->flag & 32) ? 2 : (o->flag & 8) ? 1 : 0;
	const char *target = o->targetTemplate;
	unsigned char table[256];
	if(o->flag & 4) fsa_code_table(o->flag, table);
	else trim_iupac_table(o->flag, table);

	MotifList motifs;
	memset(&motifs, 0, sizeof(motifs));
	if(o->methfilename && motifs_load(o->methfilename, &motifs)) exit(1);
	FILE *out = stdout;
	if(!(o->outputfilename[0] == '-' && o->outputfilename[1] == 0)) {
		out = fopen(o->outputfilename, "wb");
		if(!out) {
			fprintf(stderr, "Filename:\t%s\n", o->outputfilename);
			die_errno();
		}
		setvbuf(out, 0, _IOFBF, 1 << 22);
	}

	ccg_ctx *ctx = 0;
	int rc = ccg_init(&ctx, -1);
	if(rc) die_gpu(0, rc);
	if(motifs.n && (rc = ccg_set_motifs(ctx, motifs.n, motifs.lens, motifs.sets))) die_gpu(ctx, rc);

	/* the stored sequences of the shared-mask mode: a slot per file (-r) or per kept record, with the reference's
	 * running slot position (see the header comment) */
	size_t maxSeqs = o->numFile ? o->numFile : 1, cap = maxSeqs + 1, pos = 0;
	int numSeqs = 0, includeN = 0, have_ref = 0, begun_len = -1;
	unsigned char **slots = calloc(cap, sizeof(*slots));
	char **seqnames = (!pair && !target) ? calloc(maxSeqs, sizeof(*seqnames)) : 0;
	if(!slots || (!pair && !target && !seqnames)) die_errno();
	ByteBuf header, codes;
	bytebuf_init(&header, 256);
	bytebuf_init(&codes, 1 << 20);
	uint64_t *nibbles = 0;
	uint32_t *mask = 0;
	unsigned char *line = 0;
	int len = 0;
	unsigned minLength = o->minLength;

	for(unsigned f = 0; f < o->numFile; ++f) {
		const char *path = o->filenames[f];
		FsaReader *fr = fsa_open(path);
		if(!fr) {
			fprintf(stderr, "Filename:\t%s\n", path);
			die_errno();
		}
		if(fsa_peek(fr) != '>') {
			fprintf(stderr, "\"%s\" is not fasta.\n", path);
			exit(1);
		}
		int header_ok;
		do {
			if((size_t) numSeqs == maxSeqs) {
				/* trim.c:131-144: the arrays double, the slot pointer is re-based on the number of kept samples and the
				 * new half is zeroed -- whatever an earlier hole had pushed up there is gone from the output */
				maxSeqs <<= 1;
				for(size_t k = (size_t) numSeqs; k < 2 * (size_t) numSeqs && k < cap; ++k) {
					free(slots[k]);
					slots[k] = 0;
				}
				if(seqnames) {
					seqnames = realloc(seqnames, maxSeqs * sizeof(*seqnames));
					if(!seqnames) die_errno();
					memset(seqnames + numSeqs, 0, (maxSeqs - (size_t) numSeqs) * sizeof(*seqnames));
				}
				pos = (size_t) numSeqs;
			}
			if(pos + 1 >= cap) {
				const size_t ncap = 2 * cap + 16;
				slots = realloc(slots, ncap * sizeof(*slots));
				if(!slots) die_errno();
				memset(slots + cap, 0, (ncap - cap) * sizeof(*slots));
				cap = ncap;
			}
			/* the entry: the next record, or with -r the first one of that name */
			header_ok = 0;
			while(fsa_next_header(fr, &header)) {
				if(!target || strcmp((const char *) header.data, target) == 0) { header_ok = 1; break; }
			}
			const int got = header_ok && fsa_read_codes(fr, table, &codes);
			if(got) {
				const char *shown = target ? path : (const char *) header.data;
				if(have_ref) {
					if((int) codes.len != len) {
						fprintf(stderr, "Sequences does not match: %s %s\n", (const char *) header.data, path);
						exit(1);
					}
					/* shared mask, a later sample (trim.c:168-184): gated on its known positions alone */
					int Ns = 0;
					for(int k = 0; k < len; ++k) Ns += codes.data[k] == 4;
					const int inc = len - Ns;
					if((unsigned) inc < minLength) {
						fprintf(stderr, "# Excluded:\t%s\t( %d / %d )\n", shown, inc, len);
					} else {
						fprintf(stderr, "# Included:\t%s\t( %d / %d )\n", shown, inc, len);
						if(motifs.n) pack_nibbles(codes.data, len, nibbles);
						rc = ccg_trim_sample(ctx, codes.data, nibbles, 1, builder, 0);
						if(rc) die_gpu(ctx, rc);
						slots[pos] = malloc((size_t) len + 1);
						if(!slots[pos]) die_errno();
						memcpy(slots[pos], codes.data, (size_t) len);
						++numSeqs;
						++includeN;
					}
				} else {
					/* no reference yet (every sample under the pairwise flag): its own mask (trim.c:189-223) */
					len = (int) codes.len;
					if(minLength < o->minCov * len) minLength = (unsigned) (o->minCov * len);
					const size_t W = (size_t) len / 32 + 1;
					nibbles = realloc(nibbles, W * sizeof(*nibbles));
					mask = realloc(mask, W * sizeof(*mask));
					line = realloc(line, (size_t) len + 2);
					if(!nibbles || !mask || !line) die_errno();
					if(len != begun_len) {
						rc = ccg_trim_begin(ctx, len, o->proxi);
						if(rc) die_gpu(ctx, rc);
						begun_len = len;
					}
					if(motifs.n) pack_nibbles(codes.data, len, nibbles);
					unsigned inc = 0;
					rc = ccg_trim_sample(ctx, codes.data, nibbles, 0, 0, &inc);
					if(rc) die_gpu(ctx, rc);
					int kept = 1;
					if(inc < minLength) {
						fprintf(stderr, "# Excluded:\t%s\t( %d / %d )\n", shown, (int) inc, len);
						kept = 0;
					} else if(!pair) {
						fprintf(stderr, "# Included:\t%s\t( %d / %d )\n", shown, (int) inc, len);
						rc = ccg_trim_keep_reference(ctx);
						if(rc) die_gpu(ctx, rc);
						slots[pos] = malloc((size_t) len + 1);
						if(!slots[pos]) die_errno();
						memcpy(slots[pos], codes.data, (size_t) len);
						have_ref = 1;
						++numSeqs;
					}
					++includeN;
					if(pair && kept) {
						rc = ccg_trim_get_mask(ctx, 0, mask, 0, 0);
						if(rc) die_gpu(ctx, rc);
						print_trim_fsa(out, shown, codes.data, len, mask, o->flag, line);
					}
				}
				if(seqnames && numSeqs > 0) {
					free(seqnames[numSeqs - 1]);
					seqnames[numSeqs - 1] = strdup((const char *) header.data);
					if(!seqnames[numSeqs - 1]) die_errno();
				}
				if(!pair) ++pos;
			} else if(target && !pair) ++pos;
		} while(!target && header_ok);
		if(target && (!header_ok || !codes.len))
			fprintf(stderr, "Missing template entry (\"%s\") in file:\t%s\n", target, path);
		fsa_close(fr);
	}

	if(!includeN || (!pair && !have_ref)) {
		fprintf(stderr, "All sequences were trimmed away.\n");
	} else if(!pair) {
		unsigned inc = 0, var = 0;
		rc = ccg_trim_get_mask(ctx, (o->flag & 16) != 0, mask, &inc, &var);
		if(rc) die_gpu(ctx, rc);
		fprintf(stderr, "# %d / %d bases included in distance matrix.\n", (int) inc, len);
		if(o->flag & 16) fprintf(stderr, "# %d / %d positions with variance\n", (int) var, (int) inc);
		/* last to first (trim.c:249-255) */
		const size_t count = target ? o->numFile : (size_t) numSeqs;
		for(size_t i = count; i > 0; --i) {
			if(pos == 0) break;
			--pos;
			const char *name = target ? o->filenames[i - 1] : seqnames[i - 1];
			if(pos < cap && slots[pos] && name) print_trim_fsa(out, name, slots[pos], len, mask, o->flag, line);
		}
	}

	if(begun_len >= 0) ccg_trim_end(ctx);
	ccg_destroy(ctx);
	for(size_t k = 0; k < cap; ++k) free(slots[k]);
	free(slots);
	if(seqnames) {
		for(size_t k = 0; k < maxSeqs; ++k) free(seqnames[k]);
		free(seqnames);
	}
	free(nibbles);
	free(mask);
	free(line);
	bytebuf_free(&header);
	bytebuf_free(&codes);
	motifs_free(&motifs);
	if(out != stdout) fclose(out);
	else fflush(stdout);
}

static int help_message(FILE *out) {
	static const struct { char c; const char *name, *desc, *def; } rows[] = {
		{'i', "input", "Input file(s)", "stdin"},
		{'o', "output", "Output file", "stdout"},
		{'y', "methylation_motifs", "Mask methylation motifs from <file>", "False/None"},
		{'r', "reference", "Target reference identifier", "None"},
		{'C', "min_cov", "Minimum coverage", "50.0%"},
		{'L', "min_len", "Minimum overlapping length", "1"},
		{'P', "proximity", "Minimum proximity between SNPs", "0"},
		{'f', "flag", "Output flags", "0"},
		{'F', "flag_help", "Help on option \"-f\"", ""},
		{'h', "help", "Shows this helpmessage", ""},
	};
	fprintf(out, "#ccphylo-b200 trim: trims multiple alignments from different files, and merge them into one (masks built on a B200 GPU)\n");
	fprintf(out, "#   %-24s\t%-32s\t%s\n", "Options are:", "Desc:", "Default:");
	for(size_t k = 0; k < sizeof(rows) / sizeof(rows[0]); ++k)
		fprintf(out, "#    -%c, --%-16s\t%-32s\t%s\n", rows[k].c, rows[k].name, rows[k].desc, rows[k].def);
	return out == stderr;
}

static char short_of(const char *longname) {
	static const struct { const char *name; char c; } map[] = {
		{"input", 'i'}, {"output", 'o'}, {"methylation_motifs", 'y'}, {"reference", 'r'}, {"min_cov", 'C'}, {"min_len", 'L'},
		{"proximity", 'P'}, {"flag", 'f'}, {"flag_help", 'F'}, {"help", 'h'},
	};
	for(size_t k = 0; k < sizeof(map) / sizeof(map[0]); ++k)
		if(strcmp(map[k].name, longname) == 0) return map[k].c;
	return 0;
}

/* main_trim (trim.c:307-473) */
int main_trim(int argc, char **argv) {
	TrimOpts o;
	memset(&o, 0, sizeof(o));
	o.minLength = 1;
	o.minCov = 0.5;
	o.outputfilename = "-";
	int flag_help = 0;

	OptScan sc;
	optscan_init(&sc, argc - 1, argv + 1);
	char c, longname[64];
	while(optscan_next(&sc, &c, longname, sizeof(longname))) {
		char word[80];
		if(!c) {
			c = short_of(longname);
			if(!c) {
				snprintf(word, sizeof(word), "--%s", longname);
				die_unknown(word);
			}
		}
		switch(c) {
			case 'i': o.filenames = optscan_list(&sc, (int *) &o.numFile); break;
			case 'o': o.outputfilename = optscan_arg(&sc); break;
			case 'y': o.methfilename = optscan_arg(&sc); break;
			case 'r': o.targetTemplate = optscan_arg(&sc); break;
			case 'C': o.minCov = optscan_double(&sc) / 100; break;
			case 'L': o.minLength = (unsigned) optscan_long(&sc); break;
			case 'P': o.proxi = (unsigned) optscan_long(&sc); break;
			case 'f': o.flag = (unsigned) optscan_long(&sc); break;
			case 'F': flag_help = 1; break;
			case 'h': return help_message(stdout);
			default:
				/* a short option is named by its letter alone (the reference cuts the word behind it and prints from the
				 * letter on, dist.c:671-672 / trim.c) */
				snprintf(word, sizeof(word), "%c", c);
				die_unknown(word);
		}
	}
	if(sc.pos < sc.argc) {
		if(strcmp(sc.argv[sc.pos], "--") == 0) ++sc.pos;
		if(sc.pos < sc.argc) {
			o.filenames = sc.argv + sc.pos;
			o.numFile = (unsigned) (sc.argc - sc.pos);
		}
	}
	if(flag_help) {
		fprintf(stdout, "# Format flags output, add them to combine them.\n#\n"
		                "#   1:\tHard mask\n"
		                "#   2:\tPairwise comparison\n"
		                "#   4:\tMask gaps and ambiguous bases\n"
		                "#   8:\tUnmask soft masked bases in input\n"
		                "#  16:\tCreate pseudo alignment, not compatible with pairwise comparison\n"
		                "#  32:\tDo not include insignificant bases in pruning\n#\n");
		return 0;
	}
	static char *stdin_name[] = {"-"};
	if(!o.numFile) {
		/* the reference reads nothing without -i (and follows a null pointer with -r alone, trim.c:466-468): stdin here */
		o.filenames = stdin_name;
		o.numFile = 1;
	}
	fsa_trim(&o);
	return 0;
}
