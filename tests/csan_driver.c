/* Tiny C driver for compute-sanitizer runs (no Python in the way):
 *   csan_driver <kind> <seed> <bytes> [chunk] [shift]
 * forward + inverse round trip through the C ABI; exit 0 iff lossless.
 * CSAN_TUNE="key:value,key:value" applies further bwts_b200_tune settings (e.g. "7:2,9:2"
 * forces the binned rank scatter and the binned emit on small inputs).
 *   gcc -O2 -o tests/csan_driver tests/csan_driver.c -Lbijective-bwt_b200 -lbwts_b200 -lbwts_gen \
 *       -Wl,-rpath,$PWD/bijective-bwt_b200
 *   compute-sanitizer --tool memcheck|racecheck --error-exitcode 9 tests/csan_driver 3 5 300000 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/bwts_b200.h"
int bwts_gen(int kind, unsigned long long seed, unsigned char *out, long n);
int main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: csan_driver kind seed bytes [chunk] [shift]\n"); return 2; }
	int kind = atoi(argv[1]);
	unsigned long long seed = strtoull(argv[2], 0, 10);
	long n = atol(argv[3]);
	if (argc > 4) bwts_b200_tune(0, atol(argv[4]));
	if (argc > 5) bwts_b200_tune(1, atol(argv[5]));
	const char *extra = getenv("CSAN_TUNE");
	while (extra && *extra) {
		int key = atoi(extra);
		const char *colon = strchr(extra, ':');
		if (!colon) break;
		bwts_b200_tune(key, atol(colon + 1));
		extra = strchr(colon, ',');
		if (extra) extra++;
	}
	unsigned char *x = malloc(n), *y = malloc(n), *z = malloc(n);
	bwts_gen(kind, seed, x, n);
	int rc = bwts_b200_forward(x, n, y, 0);
	printf("forward rc=%d (%s)\n", rc, bwts_b200_strerror(rc));
	if (rc) return 1;
	rc = bwts_b200_inverse(y, n, z, 0);
	printf("inverse rc=%d (%s)\n", rc, bwts_b200_strerror(rc));
	if (rc) return 1;
	int ok = memcmp(x, z, n) == 0;
	printf("round trip %s\n", ok ? "ok" : "BROKEN");
	return ok ? 0 : 1;
}
