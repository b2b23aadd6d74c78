/*
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Stand-in for the two symbols of libdivsufsort (Yuta Mori; version unpinned
 * by the reference: only `-ldivsufsort`, /root/reference/Makefile:4) that the
 * reference's forward tools use (/root/reference/mk_bwts_sa.c:6,26,47-48;
 * /root/reference/mk_bwts_sa_new.c:6,26,50-51).  libdivsufsort is absent from
 * this image.  A suffix array is mathematically unique for a given byte string,
 * so any correct sorter reproduces the reference's bytes exactly; timings of
 * the "Suffix sort" phase are those of this substitute (oracle/sais.c), not of
 * libdivsufsort.
 */
#ifndef ORACLE_SHIM_DIVSUFSORT_H
#define ORACLE_SHIM_DIVSUFSORT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t saidx_t;
typedef uint8_t sauchar_t;

/* Suffix array of T[0..n) into SA[0..n); returns 0 on success, <0 on error. */
int divsufsort(const unsigned char *T, saidx_t *SA, saidx_t n);

#ifdef __cplusplus
}
#endif

#endif
