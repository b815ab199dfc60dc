/*
 * cmdline.h -- option scanner for the `dist` driver.
 *
 * Accepts the command-line dialect of the reference's hand-rolled parser (dist.c:508-689 over
 * cmdline.c): `--name value`, `--name=value`, `-x value`, `-xvalue`, bundled short flags
 * (`-pf3`: flags without an argument may be followed by one that takes the rest of the word or
 * the next word), list options that run until the next word starting with '-' (a lone "-" is a
 * file name: stdin), optional arguments that are taken only when the next word does not start
 * with '-', and trailing non-option words (input files).  The error texts are the reference's
 * (cmdline.h:23-26) because scripts match on them.
 */
#ifndef CCB_CMDLINE_H
#define CCB_CMDLINE_H

typedef struct {
	int argc;
	char **argv;
	int pos;          /* next word */
	char *word;       /* current word when inside a bundle of short options, else NULL */
	int off;          /* next character of the bundle */
	char name[64];    /* last option as the user wrote it, for messages */
} OptScan;

void optscan_init(OptScan *s, int argc, char **argv);
/* Next option: returns 1 and sets *shortopt (a letter, or 0 for a long option whose name is copied
 * to longopt without the leading dashes and without "=value").  Returns 0 when the options are
 * exhausted: the words from s->pos on are positional. */
int optscan_next(OptScan *s, char *shortopt, char *longopt, int longcap);
/* mandatory argument of the current option ("Missing argument at <opt>." + exit 1 otherwise) */
char *optscan_arg(OptScan *s);
/* optional argument: NULL when the next word starts with '-' or there is none */
char *optscan_optional_arg(OptScan *s);
/* list argument: first element pointer and length (>= 1) */
char **optscan_list(OptScan *s, int *count);
long optscan_long(OptScan *s);
double optscan_double(OptScan *s);
double optscan_optional_double(OptScan *s, double def);
int optscan_char(OptScan *s);

void die_missing(const char *opt);
void die_invalid(const char *opt);
void die_unknown(const char *word);

#endif
