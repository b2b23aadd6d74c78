/*
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * CPU suffix sorter (induced sorting, Nong/Zhang/Chan "SA-IS") exposing the
 * libdivsufsort entry point the reference calls:
 *     divsufsort(T, sa, len)      /root/reference/mk_bwts_sa.c:47-48
 *                                 /root/reference/mk_bwts_sa_new.c:50-51
 * libdivsufsort itself (third party, un-vendored, version unpinned) is absent
 * from this image.  The suffix array of a string is unique, so this sorter
 * yields the same SA -- hence the same BWTS bytes -- as the real library.
 *
 * Convention: no sentinel is stored; the suffix starting at n-1 is treated as
 * if it were followed by a symbol smaller than every byte.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "shim/divsufsort.h"

typedef int32_t sidx;

#define SYM(i) (wide ? ((const sidx *)T)[i] : (sidx)((const unsigned char *)T)[i])
#define IS_S(i) (ty[(i) >> 3] & (1u << ((i) & 7)))
#define IS_LMS(i) ((i) > 0 && IS_S(i) && !IS_S((i) - 1))

static void symbol_counts(const void *T, int wide, sidx n, sidx sigma, sidx *cnt)
{
    for (sidx c = 0; c < sigma; c++) cnt[c] = 0;
    for (sidx i = 0; i < n; i++) cnt[SYM(i)]++;
}

/* heads != 0: bucket start offsets, otherwise one-past-the-end offsets */
static void bucket_offsets(const sidx *cnt, sidx sigma, sidx *bkt, int heads)
{
    sidx run = 0;
    for (sidx c = 0; c < sigma; c++) {
        run += cnt[c];
        bkt[c] = heads ? run - cnt[c] : run;
    }
}

/* Given LMS suffixes seeded at bucket tails (other slots -1), induce L then S. */
static void induce(const void *T, int wide, const unsigned char *ty, sidx *SA,
                   sidx n, sidx sigma, const sidx *cnt, sidx *bkt)
{
    bucket_offsets(cnt, sigma, bkt, 1);
    /* the (virtual) empty suffix ranks first and induces n-1, which is L-type */
    SA[bkt[SYM(n - 1)]++] = n - 1;
    for (sidx i = 0; i < n; i++) {
        sidx j = SA[i];
        if (j > 0 && !IS_S(j - 1)) SA[bkt[SYM(j - 1)]++] = j - 1;
    }
    bucket_offsets(cnt, sigma, bkt, 0);
    for (sidx i = n - 1; i >= 0; i--) {
        sidx j = SA[i];
        if (j > 0 && IS_S(j - 1)) SA[--bkt[SYM(j - 1)]] = j - 1;
    }
}

static int lms_substrings_equal(const void *T, int wide, const unsigned char *ty,
                                sidx n, sidx a, sidx b)
{
    for (sidx d = 0;; d++) {
        sidx ia = a + d, ib = b + d;
        if (ia >= n || ib >= n) return 0; /* one of them runs into the end marker */
        if (SYM(ia) != SYM(ib)) return 0;
        if ((IS_S(ia) != 0) != (IS_S(ib) != 0)) return 0;
        if (d > 0 && (IS_LMS(ia) || IS_LMS(ib))) return 1;
    }
}

static int sais_level(const void *T, int wide, sidx *SA, sidx n, sidx sigma)
{
    if (n == 0) return 0;
    if (n == 1) { SA[0] = 0; return 0; }

    unsigned char *ty = (unsigned char *)calloc(((size_t)n >> 3) + 1, 1);
    sidx *cnt = (sidx *)malloc(sizeof(sidx) * (size_t)sigma);
    sidx *bkt = (sidx *)malloc(sizeof(sidx) * (size_t)sigma);
    if (!ty || !cnt || !bkt) { free(ty); free(cnt); free(bkt); return -2; }

    /* suffix types; n-1 is L-type (bit clear) */
    for (sidx i = n - 2; i >= 0; i--) {
        sidx a = SYM(i), b = SYM(i + 1);
        if (a < b || (a == b && IS_S(i + 1))) ty[i >> 3] |= (unsigned char)(1u << (i & 7));
    }

    symbol_counts(T, wide, n, sigma, cnt);

    /* stage 1: sort the LMS substrings */
    bucket_offsets(cnt, sigma, bkt, 0);
    for (sidx i = 0; i < n; i++) SA[i] = -1;
    sidx m = 0;
    for (sidx i = 1; i < n; i++)
        if (IS_LMS(i)) { SA[--bkt[SYM(i)]] = i; m++; }
    induce(T, wide, ty, SA, n, sigma, cnt, bkt);

    /* compact the sorted LMS positions to SA[0..m) */
    sidx w = 0;
    for (sidx i = 0; i < n; i++) {
        sidx j = SA[i];
        if (IS_LMS(j)) SA[w++] = j;
    }
    /* name them; name of LMS position j is parked at SA[m + j/2] */
    for (sidx i = m; i < n; i++) SA[i] = -1;
    sidx names = 0, last = -1;
    for (sidx i = 0; i < m; i++) {
        sidx j = SA[i];
        if (last < 0 || !lms_substrings_equal(T, wide, ty, n, last, j)) names++;
        last = j;
        SA[m + (j >> 1)] = names - 1;
    }
    /* reduced string, in text order, packed at the tail of SA */
    sidx *S1 = SA + n - m;
    {
        sidx r = m;
        for (sidx i = n - 1; i >= m; i--)
            if (SA[i] >= 0) S1[--r] = SA[i];
    }

    /* stage 2: order the LMS suffixes */
    if (names < m) {
        int rc = sais_level(S1, 1, SA, m, names);
        if (rc) { free(ty); free(cnt); free(bkt); return rc; }
    } else {
        for (sidx i = 0; i < m; i++) SA[S1[i]] = i;
    }

    /* stage 3: translate to text positions, seed buckets in sorted order, induce */
    {
        sidx r = 0;
        for (sidx i = 1; i < n; i++)
            if (IS_LMS(i)) S1[r++] = i;
    }
    for (sidx i = 0; i < m; i++) SA[i] = S1[SA[i]];
    for (sidx i = m; i < n; i++) SA[i] = -1;
    bucket_offsets(cnt, sigma, bkt, 0);
    for (sidx i = m - 1; i >= 0; i--) {
        sidx j = SA[i];
        SA[i] = -1;
        SA[--bkt[SYM(j)]] = j;
    }
    induce(T, wide, ty, SA, n, sigma, cnt, bkt);

    free(ty); free(cnt); free(bkt);
    return 0;
}

int divsufsort(const unsigned char *T, saidx_t *SA, saidx_t n)
{
    if (!T || !SA || n < 0) return -1;
    return sais_level(T, 0, SA, n, 256);
}
