/* mbwt_new -- drop-in for the reference's `mbwt_new` (/root/reference/mk_bwts_sa_new.c:36-86):
 * as mk_bwts, except that without an outfile the output goes to a fresh
 * BWTS_out_XXXXXX.bwts in the current directory and its name is printed
 * (mk_bwts_sa_new.c:60-72).  The seam :50-55 is the one library call. */
#define _GNU_SOURCE
#include "map_file.h"
#include "tool_common.h"

static unsigned char *T;
static long len;

int main(int argc, char **argv)
{
	if (argc < 2) {
		fprintf(stderr, "Usage: mk_bwts_sa <infile> [<outfile.bwts>]\n");
		fprintf(stderr, "If unspecified, output is written to a temp file\n");
		exit(1);
	}
	char *outname = argc < 3 ? NULL : argv[2];
	map_in(T, len, argv[1]);

	unsigned char *bwts = run_transform(0, T, len);

	FILE *fp;
	char generated[] = "BWTS_out_XXXXXX.bwts";
	if (outname) {
		fp = fopen(outname, "wb");
	} else {
		int fd = mkstemps(generated, 5);
		printf("Writing to %s\n", generated);
		fp = fdopen(fd, "w");
		outname = generated;
	}
	if (!fp) {
		fprintf(stderr, "Couldn't open BWTS file for writing\n");
		perror(outname);
		exit(1);
	}
	write_output(bwts, len, fp, "Write BWTS");
	finish(fp);
	return 0;
}
