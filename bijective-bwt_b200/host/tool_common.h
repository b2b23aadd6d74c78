/* tool_common.h -- shared by mk_bwts.c, mbwt_new.c, unbwts.c: run the library on the mapped
 * input, honouring two ADDITIVE environment knobs that default to the reference's behaviour:
 *   BWTS_B200_BLOCK=<bytes>     independent blocks of that size (default: whole file, one block)
 *   BWTS_B200_DEVICES=<n>       deal blocks over GPUs 0..n-1          (default: 1)
 *   BWTS_B200_TIMINGS=1         per-phase device times to stderr, the counterpart of the
 *                               reference's -DSHOW_TIMINGS (/root/reference/mk_bwts_sa.c:13-22)
 */
#ifndef BWTS_B200_TOOL_COMMON_H
#define BWTS_B200_TOOL_COMMON_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bwts_b200.h"

static long env_long(const char *name, long dflt)
{
	const char *v = getenv(name);
	return (v && *v) ? atol(v) : dflt;
}

static void print_timings(bwts_b200_ctx *ctx)
{
	bwts_b200_stats s;
	if (bwts_b200_get_stats(ctx, &s) != 0) return;
	for (int c = 0; c < BWTS_B200_NCLASS; c++)
		if (s.class_launches[c])
			fprintf(stderr, "%s time %0.3f (%ld launches)\n", bwts_b200_class_name(c),
			        s.class_ms[c] / 1000.0, s.class_launches[c]);
	fprintf(stderr, "Transform time %0.3f\n", s.total_ms / 1000.0);
}

/* direction 0 = forward, 1 = inverse.  Exits with the reference's convention on failure. */
static unsigned char *run_transform(int direction, const unsigned char *in, long len)
{
	unsigned char *out = (unsigned char *)malloc((size_t)len);
	if (!out) {
		fprintf(stderr, "Out of memory\n");
		exit(1);
	}
	long block = env_long("BWTS_B200_BLOCK", 0);
	int ndev = (int)env_long("BWTS_B200_DEVICES", 1);
	int rc;
	if (block > 0 || ndev > 1) {
		rc = direction ? bwts_b200_inverse_blocks(in, len, block, out, NULL, ndev)
		               : bwts_b200_forward_blocks(in, len, block, out, NULL, ndev);
	} else if (env_long("BWTS_B200_TIMINGS", 0)) {
		bwts_b200_ctx *ctx = bwts_b200_create(0);
		if (!ctx) {
			rc = BWTS_B200_ENODEV;
		} else {
			rc = direction ? bwts_b200_inverse_host(ctx, in, len, out) : bwts_b200_forward_host(ctx, in, len, out);
			if (rc == 0) print_timings(ctx);
			bwts_b200_destroy(ctx);
		}
	} else {
		rc = direction ? bwts_b200_inverse(in, len, out, 0) : bwts_b200_forward(in, len, out, 0);
	}
	if (rc != 0) {
		fprintf(stderr, "bwts_b200: %s\n", bwts_b200_strerror(rc));
		exit(1);
	}
	return out;
}

#endif
