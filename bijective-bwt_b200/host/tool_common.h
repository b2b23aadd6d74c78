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
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include "../../include/bwts_b200.h"

static long env_long(const char *name, long dflt)
{
	const char *v = getenv(name);
	return (v && *v) ? atol(v) : dflt;
}

/* BWTS_B200_TIMINGS=1: the reference's `<label> time <seconds>` lines (-DSHOW_TIMINGS,
 * /root/reference/mk_bwts_sa.c:13-22) in the reference's order -- Suffix sort (:50), Compute ISA
 * (:124), Fix sort order (:168), Generate BWTS (:190); the tool adds Write BWTS (:62) after its
 * fwrite -- booked from CUDA-event times of the kernels of each phase, then one line of per-transform
 * diagnostics, the counterpart of /root/reference/mk_bwts_new_algo.c:127, then the kernel classes. */
static void print_timings(bwts_b200_ctx *ctx)
{
	bwts_b200_stats s;
	if (bwts_b200_get_stats(ctx, &s) != 0) return;
	for (int p = 0; p < BWTS_B200_NPHASE; p++) {
		const char *name = bwts_b200_phase_name(s.direction, p);
		if (name) fprintf(stderr, "%s time %0.3f\n", name, s.phase_ms[p] / 1000.0);
	}
	if (s.direction == 0)
		fprintf(stderr, "Factors: %10ld; longest: %10ld; alphabet bits: %d; initial depth: %d; live after initial sort: %10ld; "
		        "doubling rounds: %d (warp-local %d, CTA-local %d, tuple set %d); radix passes: %d; live sum: %ld; workspace bytes/byte: %.1f\n",
		        s.factors, s.longest_factor, s.alphabet_bits, s.initial_depth, s.first_live, s.rounds, s.local_rounds,
		        s.cta_rounds, s.tuple_rounds, s.radix_passes, s.live_sum, s.len ? (double)s.arena_bytes / (double)s.len : 0.0);
	else
		fprintf(stderr, "Cycles: %10ld; sublists: %10ld; unreached: %10ld; workspace bytes/byte: %.1f\n", s.factors,
		        s.splitters, s.unreached, s.len ? (double)s.arena_bytes / (double)s.len : 0.0);
	for (int c = 0; c < BWTS_B200_NCLASS; c++)
		if (s.class_launches[c])
			fprintf(stderr, "  class %s time %0.6f (%ld launches)\n", bwts_b200_class_name(c),
			        s.class_ms[c] / 1000.0, s.class_launches[c]);
	fprintf(stderr, "Transform time %0.3f (H2D %0.3f, D2H %0.3f)\n", s.total_ms / 1000.0, s.h2d_ms / 1000.0, s.d2h_ms / 1000.0);
}

/* fwrite with the reference's last mark: "Write BWTS time" (/root/reference/mk_bwts_sa.c:60-62) */
/* Leave without tearing the CUDA context down piece by piece (0.3-0.5 s with a 70 GB workspace): the
 * output is flushed, the kernel reclaims everything else. */
static void finish(FILE *fp)
{
	if (fp && fp != stdout) fclose(fp);
	fflush(NULL);
	_exit(0);
}

static void write_output(const unsigned char *data, long len, FILE *fp, const char *label)
{
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	fwrite(data, 1, (size_t)len, fp);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (env_long("BWTS_B200_TIMINGS", 0))
		fprintf(stderr, "%s time %0.3f\n", label, (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
}

/* direction 0 = forward, 1 = inverse.  Exits with the reference's convention on failure. */
static unsigned char *run_transform(int direction, const unsigned char *in, long len)
{
	/* the output buffer: anonymous, pre-faulted mapping (a fresh malloc would take its page faults inside the
	 * device-to-host copy, one per 4 KiB) */
	unsigned char *out = (unsigned char *)mmap(NULL, (size_t)len, PROT_READ | PROT_WRITE,
	                                           MAP_PRIVATE | MAP_ANONYMOUS | MAP_POPULATE, -1, 0);
	if (out == (unsigned char *)MAP_FAILED) {
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
		/* wall-clock marks of the host side as well (lines start with a blank: not phase lines) */
		struct timespec t0, t1, t2, t3;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		bwts_b200_ctx *ctx = bwts_b200_create(0);
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (!ctx) {
			rc = BWTS_B200_ENODEV;
		} else {
			rc = bwts_b200_reserve(ctx, len);
			clock_gettime(CLOCK_MONOTONIC, &t2);
			if (rc == 0)
				rc = direction ? bwts_b200_inverse_host(ctx, in, len, out) : bwts_b200_forward_host(ctx, in, len, out);
			clock_gettime(CLOCK_MONOTONIC, &t3);
			if (rc == 0) {
				print_timings(ctx);
				fprintf(stderr, " host: context %0.3f s, workspace %0.3f s, copy in + transform + copy out %0.3f s\n",
				        (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec),
				        (double)(t2.tv_sec - t1.tv_sec) + 1e-9 * (double)(t2.tv_nsec - t1.tv_nsec),
				        (double)(t3.tv_sec - t2.tv_sec) + 1e-9 * (double)(t3.tv_nsec - t2.tv_nsec));
			}
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
