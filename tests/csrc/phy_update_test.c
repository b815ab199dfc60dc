/* phy_update_test.c -- ccphylo_b200/host/phy_update.c from the command line, for tests/test_host_c.py:
 *   phy_update_test names <phy> <dir> <sep>                       prints the status, then one path per line
 *   phy_update_test append <phy> <n> <name> <flag> <precision> <cell>...   appends the row */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "phy_update.h"

int main(int argc, char **argv) {
	if(argc >= 5 && strcmp(argv[1], "names") == 0) {
		PhyNames phy;
		int ok = read_phy_names(argv[2], argv[3], argv[4][0], &phy);
		printf("%d\n", ok);
		if(ok == 1)
			for(int i = 0; i < phy.n; ++i) printf("%s\n", phy.paths[i]);
		return 0;
	}
	if(argc >= 7 && strcmp(argv[1], "append") == 0) {
		const int n = atoi(argv[3]);
		double *row = calloc((size_t) (n > 0 ? n : 1), sizeof(double));
		for(int k = 0; k < n - 1 && 7 + k < argc; ++k) row[k] = strtod(argv[7 + k], 0);
		phy_append_row(argv[2], n, argv[4], row, (unsigned) atoi(argv[5]), atoi(argv[6]));
		free(row);
		return 0;
	}
	return 2;
}
