/* cmdline.c -- see cmdline.h */
#include "cmdline.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void die_missing(const char *opt) {
	fprintf(stderr, "Missing argument at %s.\n", opt);
	exit(1);
}

void die_invalid(const char *opt) {
	fprintf(stderr, "Invalid value parsed at %s.\n", opt);
	exit(1);
}

void die_unknown(const char *word) {
	fprintf(stderr, "Unknown argument or option: \"%s\"\n", word);
	exit(1);
}

void optscan_init(OptScan *s, int argc, char **argv) {
	memset(s, 0, sizeof(*s));
	s->argc = argc;
	s->argv = argv;
	s->pos = 0;
}

int optscan_next(OptScan *s, char *shortopt, char *longopt, int longcap) {
	*shortopt = 0;
	if(longopt && longcap) *longopt = 0;
	/* inside a bundle of short options */
	if(s->word && s->off > 0 && s->word[s->off]) {
		*shortopt = s->word[s->off++];
		s->name[0] = *shortopt;
		s->name[1] = 0;
		return 1;
	}
	s->word = 0;
	if(s->pos >= s->argc) return 0;
	char *w = s->argv[s->pos];
	if(w[0] != '-' || w[1] == 0) return 0;        /* first positional word (a lone "-" is a file) */
	++s->pos;
	if(w[1] == '-') {
		if(w[2] == 0) return 0;                   /* "--" ends the options */
		const char *name = w + 2;
		size_t len = strcspn(name, "=");
		if(longopt && longcap) {
			size_t c = len < (size_t) longcap - 1 ? len : (size_t) longcap - 1;
			memcpy(longopt, name, c);
			longopt[c] = 0;
		}
		snprintf(s->name, sizeof(s->name), "%.*s", (int) (len < 60 ? len : 60), name);
		/* "--name=value": the value is the rest of this word */
		s->word = w;
		s->off = (int) (2 + len + (name[len] == '=' ? 1 : 0));
		if(!name[len]) s->word = 0;
		else if(!w[s->off]) s->word = 0;
		/* mark that what follows in word is an attached value, not more short options */
		if(s->word) s->off = -s->off;
		return 1;
	}
	s->word = w;
	s->off = 1;
	*shortopt = s->word[s->off++];
	s->name[0] = *shortopt;
	s->name[1] = 0;
	return 1;
}

/* value attached to the current word, if any; consumes it */
static char *attached(OptScan *s) {
	if(!s->word) return 0;
	int off = s->off < 0 ? -s->off : s->off;
	char *v = s->word + off;
	s->word = 0;
	return *v ? v : 0;
}

char *optscan_arg(OptScan *s) {
	char *v = attached(s);
	if(v) return v;
	if(s->pos >= s->argc) die_missing(s->name);
	return s->argv[s->pos++];
}

char *optscan_optional_arg(OptScan *s) {
	char *v = attached(s);
	if(v) return v;
	if(s->pos >= s->argc || s->argv[s->pos][0] == '-') return 0;
	return s->argv[s->pos++];
}

char **optscan_list(OptScan *s, int *count) {
	char **first;
	char *v = attached(s);
	int n = 0;
	if(v) {
		/* "-ifile more files": the attached value is the first element, in place */
		s->argv[s->pos - 1] = v;
		first = &s->argv[s->pos - 1];
		n = 1;
	} else {
		if(s->pos >= s->argc) die_missing(s->name);
		first = &s->argv[s->pos];
	}
	while(s->pos < s->argc && (s->argv[s->pos][0] != '-' || s->argv[s->pos][1] == 0)) {
		++s->pos;
		++n;
	}
	*count = n;
	return first;
}

long optscan_long(OptScan *s) {
	char name[64], *end;
	snprintf(name, sizeof(name), "%s", s->name);
	char *v = optscan_arg(s);
	long x = strtol(v, &end, 10);
	if(*end) die_invalid(name);
	return x;
}

double optscan_double(OptScan *s) {
	char name[64], *end;
	snprintf(name, sizeof(name), "%s", s->name);
	char *v = optscan_arg(s);
	double x = strtod(v, &end);
	if(*end) die_invalid(name);
	return x;
}

double optscan_optional_double(OptScan *s, double def) {
	char name[64], *end;
	snprintf(name, sizeof(name), "%s", s->name);
	char *v = optscan_optional_arg(s);
	if(!v) return def;
	double x = strtod(v, &end);
	if(*end) die_invalid(name);
	return x;
}

int optscan_char(OptScan *s) {
	char name[64];
	snprintf(name, sizeof(name), "%s", s->name);
	char *v = optscan_arg(s);
	if(!v[0] || v[1]) die_invalid(name);
	return (unsigned char) v[0];
}
