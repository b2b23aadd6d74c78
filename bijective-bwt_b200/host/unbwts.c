/* unbwts -- drop-in for the reference's `unbwts` (/root/reference/unbwts.c:19-92,136-176):
 * same argv and usage text; without an outfile the output goes to <infile>_XXXXXX and the
 * name is printed.  The seam unbwts.c:31-86 is the one library call. */
#define _GNU_SOURCE
#include "map_file.h"
#include "tool_common.h"

static unsigned char *BWTS;
static long len;

static void write_out(const unsigned char *data, long n, char *out_arg, char *in_arg)
{
	char *name;
	FILE *fp = NULL;
	if (out_arg != NULL) {
		name = out_arg;
		fp = fopen(name, "wb");
	} else {
		if (asprintf(&name, "%s_XXXXXX", in_arg) <= 0) {
			fprintf(stderr, "Allocating outfile name failed. Abort\n");
			exit(1);
		}
		int fd = mkstemps(name, 0);
		printf("Writing to %s\n", name);
		fp = fdopen(fd, "w");
	}
	if (!fp) {
		fprintf(stderr, "Couldn't open output file for writing\n");
		perror(name);
		exit(1);
	}
	write_output(data, n, fp, "Write text");
	fclose(fp);
}

int main(int argc, char **argv)
{
	if (argc < 2) {
		fprintf(stderr, "Usage: unbwts <infile.bwts> [<outfile>]\n");
		fprintf(stderr, "If output file name is unspecified, a name is generated\n");
		exit(1);
	}
	char *inname = argv[1];
	char *outname = argc < 3 ? NULL : argv[2];
	map_in(BWTS, len, inname);

	unsigned char *text = run_transform(1, BWTS, len);

	write_out(text, len, outname, inname);
	finish(NULL);
	return 0;
}
