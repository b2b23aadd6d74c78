/*
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this file's
 * library; the product (libbwts_b200.so, the host tools) never does.
 *
 * CPU restatement of the reference's two hot paths, in plain C:
 *   forward  = suffix array + sequential SA/ISA fix-up + scatter
 *              (/root/reference/mk_bwts_sa.c:47-52,74-195; same algorithm in
 *               /root/reference/mk_bwts_sa_new.c:50-55,95-241)
 *   inverse  = byte counts, exclusive scan, stable LF map, cycle walk
 *              (/root/reference/unbwts.c:31-86)
 * plus Duval's factorisation, used by the tests as an independent statement of
 * the Lyndon boundaries the reference finds as prefix minima of the ISA
 * (/root/reference/mk_bwts_sa.c:126-129).
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (its Makefile:30-38 targets need an absent testdata/ and exercise only
 * mk_bwts_new_algo).  This restatement is therefore pinned against the
 * UNMODIFIED reference sources compiled into oracle/_ref/ (oracle/Makefile),
 * whose outputs are committed as tests/golden/ fixtures by
 * tests/golden/make_golden.py, and against a brute-force statement of the
 * BWTS definition (tests/bwts_definition.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "shim/divsufsort.h"

/* ---- forward ----------------------------------------------------------- */

typedef struct {
    const unsigned char *text;
    int32_t n;
    int32_t *sa;   /* rank -> text position */
    int32_t *isa;  /* text position -> rank  */
} fwd_state;

/* Move rank slot r+1 down to r (one step of the reference's "shift SA/ISA one
 * position to left", mk_bwts_sa.c:104-106 and :150-152). */
static inline void pull_down(fwd_state *st, int32_t r)
{
    int32_t p = st->sa[r + 1];
    st->sa[r] = p;
    st->isa[p] = r;
}

static inline void seat(fwd_state *st, int32_t pos, int32_t r)
{
    st->sa[r] = pos;
    st->isa[pos] = r;
}

/* Head of the factor [s, s+flen) currently at rank r: advance it past every
 * following suffix that (a) starts beyond the end of this factor and (b) is
 * not larger than the factor read as an infinite repetition.  Follows
 * move_lyndonword_head, mk_bwts_sa.c:74-112. */
static int32_t slide_head(fwd_state *st, int32_t s, int32_t flen, int32_t r)
{
    const int32_t n = st->n;
    for (; r + 1 < n; r++) {
        const int32_t q = st->sa[r + 1];
        if (!(q > s + flen)) break;                       /* :80 loop guard */
        int32_t span = n - q < flen ? n - q : flen;       /* :83 */
        int c = memcmp(st->text + s, st->text + q, (size_t)span); /* :86 */
        if (c < 0) break;                                 /* :88-90 */
        if (c == 0 && q + flen < n && r < st->isa[q + flen]) break; /* :91-101 */
        pull_down(st, r);
    }
    seat(st, s, r);
    return r;
}

/* Interior positions of the factor [s, e), visited from e-1 down to s+1: each
 * is bubbled right inside its first-byte bucket while the neighbour starts
 * later in the text, shares the first byte, and the rank just fixed for the
 * cyclic successor is not below the neighbour's successor rank.  The first
 * position that stays put ends the factor.  Follows mk_bwts_sa.c:133-160. */
static void reseat_interior(fwd_state *st, int32_t s, int32_t e, int32_t head_rank)
{
    const int32_t n = st->n;
    int32_t succ_rank = head_rank;
    for (int32_t j = e - 1; j > s; j--) {
        const int32_t from = st->isa[j];
        int32_t r = from;
        while (r < n - 1) {
            const int32_t q = st->sa[r + 1];
            if (j > q) break;
            if (st->text[j] != st->text[q]) break;
            if (succ_rank < st->isa[q + 1]) break;
            pull_down(st, r);
            r++;
        }
        seat(st, j, r);
        succ_rank = r;
        if (r == from) break;
    }
}

int oracle_bwts_forward(const unsigned char *in, long len, unsigned char *out)
{
    if (len <= 0 || len > 0x7fffffffL || !in || !out) return -1;
    fwd_state st;
    st.text = in;
    st.n = (int32_t)len;
    st.sa = (int32_t *)malloc(sizeof(int32_t) * (size_t)len);
    st.isa = (int32_t *)malloc(sizeof(int32_t) * (size_t)len);
    if (!st.sa || !st.isa) { free(st.sa); free(st.isa); return -2; }
    if (divsufsort(in, st.sa, st.n) != 0) { free(st.sa); free(st.isa); return -3; }
    const int32_t n = st.n;

    for (int32_t r = 0; r < n; r++) st.isa[st.sa[r]] = r;          /* :119-122 */

    /* factor starts = strict prefix minima of isa; the final factor (the one
     * whose head has rank 0, or that reaches the end of the text) is left
     * untouched, as in the reference (:126-165). */
    int32_t fstart = 0, frank = st.isa[0];
    for (int32_t i = 1; i < n && frank > 0; i++) {
        if (st.isa[i] < frank) {
            int32_t hr = slide_head(&st, fstart, i - fstart, frank);
            reseat_interior(&st, fstart, i, hr);
            fstart = i;
            frank = st.isa[i];
        }
    }
    free(st.sa);
    st.sa = NULL;

    /* emit: rank(i) receives the byte cyclically preceding i inside its own
     * factor (:170-188).  Note isa[] of already-fixed factor heads only grew,
     * and the prefix-minimum test is re-run on the fixed isa exactly as the
     * reference does. */
    int32_t cur_rank = n, have = 0;
    for (int32_t i = 0; i < n; i++) {
        if (st.isa[i] < cur_rank) {
            if (have) out[cur_rank] = in[i - 1];
            cur_rank = st.isa[i];
            have = 1;
        } else {
            out[st.isa[i]] = in[i - 1];
        }
    }
    out[0] = in[n - 1];
    free(st.isa);
    return 0;
}

/* ---- inverse ----------------------------------------------------------- */

int oracle_bwts_inverse(const unsigned char *in, long len, unsigned char *out)
{
    if (len <= 0 || len > 0x7fffffffL || !in || !out) return -1;
    const int32_t n = (int32_t)len;
    int32_t first[256];
    memset(first, 0, sizeof first);
    for (int32_t i = 0; i < n; i++) first[in[i]]++;               /* unbwts.c:31-36 */
    for (int32_t c = 0, run = 0; c < 256; c++) {                  /* :38-43 */
        int32_t k = first[c];
        first[c] = run;
        run += k;
    }
    int32_t *lf = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    if (!lf) return -2;
    for (int32_t i = 0; i < n; i++) lf[i] = first[in[i]]++;       /* :50-52 */

    /* cycles in order of ascending smallest index; bytes land back to front
     * (:62-86).  A visited slot is marked with -1. */
    int32_t wpos = n - 1;
    for (int32_t start = 0; start < n; start++) {
        if (lf[start] < 0) continue;
        int32_t p = start;
        do {
            int32_t nx = lf[p];
            out[wpos--] = in[p];
            lf[p] = -1;
            p = nx;
        } while (p != start);
    }
    free(lf);
    return 0;
}

/* ---- helpers for tests -------------------------------------------------- */

/* Duval's algorithm: writes the start of every Lyndon factor, returns count. */
long oracle_lyndon_starts(const unsigned char *in, long len, int32_t *starts)
{
    long cnt = 0, f = 0;
    while (f < len) {
        long i = f, k = f + 1;
        while (k < len && in[i] <= in[k]) {
            if (in[i] < in[k]) i = f; else i++;
            k++;
        }
        long p = k - i;
        while (f <= i) {
            if (starts) starts[cnt] = (int32_t)f;
            cnt++;
            f += p;
        }
    }
    return cnt;
}

int oracle_suffix_array(const unsigned char *in, long len, int32_t *sa)
{
    if (len < 0 || len > 0x7fffffffL) return -1;
    return divsufsort(in, sa, (int32_t)len);
}

/* stable LF map only (the front half of the inverse), for kernel-level tests */
int oracle_lf_map(const unsigned char *in, long len, int32_t *lf)
{
    if (len <= 0 || len > 0x7fffffffL) return -1;
    int32_t first[256];
    memset(first, 0, sizeof first);
    for (long i = 0; i < len; i++) first[in[i]]++;
    for (int32_t c = 0, run = 0; c < 256; c++) {
        int32_t k = first[c];
        first[c] = run;
        run += k;
    }
    for (long i = 0; i < len; i++) lf[i] = first[in[i]]++;
    return 0;
}
