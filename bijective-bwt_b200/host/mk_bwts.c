/* mk_bwts -- drop-in for the reference's `mk_bwts` (/root/reference/mk_bwts_sa.c:33-65):
 * same argv, same usage text, same mmap input, same raw output (stdout when no outfile).
 * The seam mk_bwts_sa.c:47-52 (malloc sa / divsufsort / make_bwts_sa) is the one library call. */
#include "map_file.h"
#include "tool_common.h"

static unsigned char *T;
static long len;

int main(int argc, char **argv)
{
	if (argc < 2) {
		fprintf(stderr, "Usage: mk_bwts_sa <infile> [<outfile.bwts>]\n");
		fprintf(stderr, "If unspecified, output is written to standard output\n");
		exit(1);
	}
	char *outname = argc < 3 ? NULL : argv[2];
	map_in(T, len, argv[1]);

	unsigned char *bwts = run_transform(0, T, len);

	FILE *fp = outname ? fopen(outname, "w") : stdout;
	if (!fp) {
		fprintf(stderr, "Couldn't open BWTS file for writing\n");
		perror(outname);
		exit(1);
	}
	write_output(bwts, len, fp, "Write BWTS");
	finish(fp);
	return 0;
}
