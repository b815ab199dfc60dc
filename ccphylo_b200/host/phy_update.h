/*
 * phy_update.h -- the Phylip side of `dist -a`: the sample count and row names of an existing single-matrix file
 * (getSizePhy + getFilenamesPhy, phy.c:509-650) and the append of one row (printphyUpdate, phy.c:201-250).
 */
#ifndef CCB_PHY_UPDATE_H
#define CCB_PHY_UPDATE_H

typedef struct {
	int n;                   /* rows of the existing matrix */
	char **paths;            /* directory of the first -i argument + the row's name */
} PhyNames;

/* 1: ok; 0: malformed file (message printed); -1: the file holds more than one matrix (message printed) */
int read_phy_names(const char *phyname, const char *dir, char sep, PhyNames *out);
/* new count over the first ten bytes, the new row (n - 1 cells) at the end; name is stripped of its directory and
 * of enclosing quotes in place */
void phy_append_row(const char *phyname, int n, char *name, const double *row, unsigned flag, int precision);

#endif
