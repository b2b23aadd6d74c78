/*
 * Deterministic synthetic inputs for the configurations BASELINE.json names
 * (SURVEY.md section 8d).  PRNG = splitmix64.  Built twice: as libbwts_gen.so
 * (ctypes, used by tests and bench.py) and as the `gen_input` tool.
 *
 *   kind 1  uniform random bytes                                  (C1)
 *   kind 2  English-like order-2 Markov text, 64 symbols, ~2 bits/char  (C2, C5 blocks)
 *   kind 3  64 KiB kind-2 block tiled, one byte substitution per MiB    (C3)
 *   kind 4  DNA over ACGT: 30 % copied segments (20-2000 B) of earlier
 *           text, else 20-500 i.i.d. bases                              (C4)
 *   kind 6  Fibonacci word over {a,b}                                   (C3 stress)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } rng_t;

static inline uint64_t rng_next(rng_t *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline uint64_t rng_below(rng_t *r, uint64_t bound) { return rng_next(r) % bound; }

static const char SYMS[65] =
    " abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ.,;:!?'\"-(\n";

static void gen_random(rng_t *r, unsigned char *out, long n)
{
    for (long i = 0; i < n; i++) out[i] = (unsigned char)(rng_next(r) >> 56);
}

static void gen_text(rng_t *r, unsigned char *out, long n)
{
    /* one permutation of the 64 symbols per two-symbol context */
    unsigned char *perm = (unsigned char *)malloc(64 * 64 * 64);
    for (int ctx = 0; ctx < 64 * 64; ctx++) {
        unsigned char *p = perm + ctx * 64;
        for (int k = 0; k < 64; k++) p[k] = (unsigned char)k;
        for (int k = 63; k > 0; k--) {
            int j = (int)rng_below(r, (uint64_t)k + 1);
            unsigned char t = p[k]; p[k] = p[j]; p[j] = t;
        }
    }
    int a = 0, b = 0;
    for (long i = 0; i < n; i++) {
        uint64_t x = rng_next(r);
        int g = x ? __builtin_ctzll(x) : 63;
        if (g > 63) g = 63;
        int c = perm[(a * 64 + b) * 64 + g];
        out[i] = (unsigned char)SYMS[c];
        a = b; b = c;
    }
    free(perm);
}

static void gen_tiled(rng_t *r, unsigned char *out, long n)
{
    const long tile = 64 * 1024;
    long first = n < tile ? n : tile;
    gen_text(r, out, first);
    for (long i = first; i < n; i++) out[i] = out[i - tile];
    /* one substitution per MiB (at least one if the input is shorter) */
    long subs = n >> 20;
    if (subs == 0) subs = 1;
    for (long k = 0; k < subs; k++) {
        long lo = k << 20, hi = lo + (1 << 20);
        if (hi > n) hi = n;
        if (lo >= hi) break;
        long p = lo + (long)rng_below(r, (uint64_t)(hi - lo));
        out[p] = (unsigned char)(rng_next(r) >> 56);
    }
}

static void gen_dna(rng_t *r, unsigned char *out, long n)
{
    static const char B[4] = { 'A', 'C', 'G', 'T' };
    long i = 0;
    while (i < n) {
        int copy = i > 0 && rng_below(r, 10) < 3;
        if (copy) {
            long L = 20 + (long)rng_below(r, 1981);
            long from = (long)rng_below(r, (uint64_t)i);
            for (long k = 0; k < L && i < n; k++, i++) out[i] = out[from + k];
        } else {
            long L = 20 + (long)rng_below(r, 481);
            for (long k = 0; k < L && i < n; k++, i++) out[i] = (unsigned char)B[rng_next(r) >> 62];
        }
    }
}

static void gen_fibonacci(unsigned char *out, long n)
{
    /* S(1)=a, S(2)=ab, S(k)=S(k-1)S(k-2); built by self-copy */
    if (n <= 0) return;
    out[0] = 'a';
    if (n == 1) return;
    out[1] = 'b';
    long prev = 1, cur = 2;
    while (cur < n) {
        long add = prev;
        if (cur + add > n) add = n - cur;
        memcpy(out + cur, out, (size_t)add);
        long t = cur; cur += prev; prev = t;
    }
}

int bwts_gen(int kind, uint64_t seed, unsigned char *out, long n)
{
    rng_t r = { seed };
    if (!out || n < 0) return -1;
    switch (kind) {
    case 1: gen_random(&r, out, n); return 0;
    case 2: gen_text(&r, out, n); return 0;
    case 3: gen_tiled(&r, out, n); return 0;
    case 4: gen_dna(&r, out, n); return 0;
    case 6: gen_fibonacci(out, n); return 0;
    default: return -1;
    }
}

#ifdef GEN_INPUT_MAIN
int main(int argc, char **argv)
{
    if (argc < 5) {
        fprintf(stderr, "Usage: gen_input <kind 1|2|3|4|6> <seed> <bytes> <outfile>\n");
        return 1;
    }
    int kind = atoi(argv[1]);
    uint64_t seed = strtoull(argv[2], NULL, 10);
    long n = atol(argv[3]);
    unsigned char *buf = (unsigned char *)malloc(n > 0 ? (size_t)n : 1);
    if (!buf || bwts_gen(kind, seed, buf, n) != 0) { fprintf(stderr, "gen_input: bad arguments\n"); return 1; }
    FILE *f = fopen(argv[4], "wb");
    if (!f) { perror(argv[4]); return 1; }
    fwrite(buf, 1, (size_t)n, f);
    fclose(f);
    return 0;
}
#endif
