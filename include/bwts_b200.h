/*
 * bwts_b200.h -- C ABI of libbwts_b200.so: the bijective Burrows-Wheeler transform
 * (BWTS, Gil & Scott) and its inverse as hand-written CUDA for NVIDIA B200 (sm_100a).
 *
 * The reference (NealB/Bijective-BWT) has no library API: its hot paths are the bodies
 * of three `main`s working on file-scope globals.  Each entry point below replaces one
 * such seam; a reference-side caller keeps `map_in(...)` before it and `fwrite(...)`
 * after it.
 *
 *   bwts_b200_forward   replaces  mk_bwts_sa.c:47-52      (malloc sa; divsufsort; make_bwts_sa)
 *                       and       mk_bwts_sa_new.c:50-55  (same, refactored)
 *   bwts_b200_inverse   replaces  unbwts.c:31-86          (counts, scan, LF map, cycle walk)
 *   bwts_b200_*_blocks  additive: independent fixed-size blocks dealt over several GPUs
 *                       (the reference transforms the whole file as one block).
 *
 * Byte layout is the reference's: `len` raw input bytes in, exactly `len` raw bytes
 * out, no header, no primary index, no terminator (mk_bwts_sa.c:60, unbwts.c:173).
 * Bytes order as unsigned char (mk_bwts_sa.c:24,86).
 *
 * Ownership: the caller owns `in` and `out` (never aliased); the library owns all
 * device memory.  Every function returns 0 on success or a negative BWTS_B200_E*
 * code; nothing prints, nothing calls exit() -- the host tools turn a failure into
 * the reference's "message to stderr, exit(1)" convention.  There is no CPU
 * fallback: without a CUDA device every transform fails with BWTS_B200_ENODEV.
 *
 * Limits: 1 <= len < 2^31 per block, the reference's own range (`int` / `saidx_t`,
 * mk_bwts_sa.c:26-27, unbwts.c:12-13).  The device workspace is 71 bytes per input byte (+ 64 MiB,
 * + 2 per byte for the host-buffer calls): 2 GiB blocks need ~157 GB of the B200's 192 GB.
 */
#ifndef BWTS_B200_H
#define BWTS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define BWTS_B200_OK         0
#define BWTS_B200_EINVAL    -1   /* NULL pointer, len <= 0, bad device / block size      */
#define BWTS_B200_ETOOBIG   -2   /* len > BWTS_B200_MAX_LEN                               */
#define BWTS_B200_ENODEV    -3   /* no usable CUDA device (no CPU fallback exists)        */
#define BWTS_B200_ENOMEM    -4   /* device or pinned-host allocation failed               */
#define BWTS_B200_ECUDA     -5   /* a CUDA call or kernel failed; see bwts_b200_last_cuda_error */
#define BWTS_B200_EINTERNAL -6   /* invariant violated (a bug)                            */

#define BWTS_B200_MAX_LEN ((1L << 31) - 1)

typedef struct bwts_b200_ctx bwts_b200_ctx;

/* ---- one-call entry points (host buffers) -------------------------------------- */

/* Forward BWTS of in[0..len) into out[0..len) on CUDA device `device`.
 * Replaces: sa = malloc(...); divsufsort(T, sa, len); bwts = make_bwts_sa();
 * (/root/reference/mk_bwts_sa.c:47-52, mk_bwts_sa_new.c:50-55).                     */
int bwts_b200_forward(const unsigned char *in, long len, unsigned char *out, int device);

/* Inverse BWTS.  Replaces /root/reference/unbwts.c:31-86.                            */
int bwts_b200_inverse(const unsigned char *in, long len, unsigned char *out, int device);

/* Independent blocks of `block_len` bytes (last one shorter), block j = bytes
 * [j*block_len, min((j+1)*block_len, len)); output block j = transform of input block
 * j at the same offsets.  Blocks are dealt round-robin over devices[0..ndev); per device
 * a three-stage pipeline (pinned staging + H2D of block b+1 | transform of block b | D2H
 * + un-staging of block b-1, each on its own stream and host thread) overlaps the copies
 * with the kernels; no inter-GPU traffic.  block_len <= 0 or
 * >= len means "whole input is one block" (= the reference's behaviour).
 * devices == NULL means devices 0..ndev-1.                                           */
int bwts_b200_forward_blocks(const unsigned char *in, long len, long block_len,
                             unsigned char *out, const int *devices, int ndev);
int bwts_b200_inverse_blocks(const unsigned char *in, long len, long block_len,
                             unsigned char *out, const int *devices, int ndev);

/* ---- context API (reusable workspace; one context per device per host thread) ---- */

int  bwts_b200_device_count(void);                 /* >= 0, or 0 when no driver/device */
bwts_b200_ctx *bwts_b200_create(int device);       /* NULL on failure                   */
void bwts_b200_destroy(bwts_b200_ctx *ctx);
int  bwts_b200_reserve(bwts_b200_ctx *ctx, long max_len);   /* pre-size the workspace   */

/* host buffers: pinned staging + H2D, transform, D2H; blocks until `out` is complete  */
int bwts_b200_forward_host(bwts_b200_ctx *ctx, const unsigned char *in, long len, unsigned char *out);
int bwts_b200_inverse_host(bwts_b200_ctx *ctx, const unsigned char *in, long len, unsigned char *out);

/* device-resident buffers (d_in, d_out: device pointers on ctx's device, not aliased; any
 * alignment -- an input that is not 16-byte aligned costs one device-to-device copy).
 * Work is issued on `stream` (a cudaStream_t passed as void*; NULL = the context's own
 * stream).  The call returns after the result is complete in d_out (the driver loop
 * reads a few counters back per doubling round).                                      */
int bwts_b200_forward_device(bwts_b200_ctx *ctx, const void *d_in, long len, void *d_out, void *stream);
int bwts_b200_inverse_device(bwts_b200_ctx *ctx, const void *d_in, long len, void *d_out, void *stream);

/* ---- introspection --------------------------------------------------------------- */

#define BWTS_B200_NCLASS 16
#define BWTS_B200_NPHASE 8
typedef struct {
    long   len;                 /* length of the last transform                         */
    int    direction;           /* 0 forward, 1 inverse                                 */
    long   factors;             /* forward: Lyndon factors; inverse: cycles             */
    long   longest_factor;      /* forward only                                         */
    int    alphabet_bits;       /* forward: bits per packed symbol                      */
    int    initial_depth;       /* forward: symbols in the initial key (k0)             */
    int    rounds;              /* forward: doubling rounds after the initial sort      */
    int    radix_passes;        /* forward: onesweep passes launched                    */
    int    local_rounds;        /* forward: rounds served by the warp-local sort        */
    int    cta_rounds;          /* forward: rounds whose large-group set was sorted CTA-locally */
    long   live_sum;            /* forward: sum over rounds of live elements            */
    long   splitters;           /* inverse: sublists                                    */
    long   unreached;           /* inverse: elements ranked by the self-walk fallback   */
    long   launches;            /* kernels launched by the last transform               */
    double total_ms;            /* device time of the last transform (CUDA events)      */
    /* per kernel class: launches, summed CUDA-event milliseconds, summed algorithmic
     * bytes (DESIGN.md section 4 states the per-element figures)                       */
    long   class_launches[BWTS_B200_NCLASS];
    double class_ms[BWTS_B200_NCLASS];
    double class_bytes[BWTS_B200_NCLASS];
    /* host-buffer entry points only: CUDA-event times of the two copies around the transform */
    double h2d_ms;
    double d2h_ms;
    int    lyndon_fallback;     /* forward: 1 if the Lyndon boundaries came from the suffix-sort fallback */
    /* device time per phase, the counterpart of the reference's MARK_TIME marks
     * (/root/reference/mk_bwts_sa.c:13-22,50,124,168,190); names: bwts_b200_phase_name       */
    double phase_ms[BWTS_B200_NPHASE];
    long   arena_bytes;         /* device workspace held by the context after this transform */
    long   first_live;          /* forward: rotations still tied after the initial sort      */
    int    tuple_rounds;        /* forward: rounds in which the text-ordered tuple set ran   */
    long   tuple_live_sum;      /* forward: sum over those rounds of its members             */
    int    inverse_attempts;    /* inverse: 1 + restarts with another splitter hash (fallback over budget) */
    int    binned_rounds;       /* forward: re-ranks whose ranks went out through the binned scatter */
} bwts_b200_stats;

int bwts_b200_get_stats(const bwts_b200_ctx *ctx, bwts_b200_stats *out);
const char *bwts_b200_class_name(int cls);          /* NULL when cls is out of range   */
/* forward (direction 0): "Suffix sort", "Compute ISA", "Fix sort order", "Generate BWTS" -- the
 * reference's own labels, in its order; inverse (direction 1): "Count bytes", "LF map",
 * "Walk sublists", "Rank sublists", "Place bytes".  NULL past the last phase.           */
const char *bwts_b200_phase_name(int direction, int phase);
int bwts_b200_set_profile(bwts_b200_ctx *ctx, int on);   /* per-launch events (default on) */

const char *bwts_b200_strerror(int code);
int bwts_b200_last_cuda_error(const bwts_b200_ctx *ctx);  /* cudaError_t of the last failure */
const char *bwts_b200_version(void);

/* Test hooks: override tuning constants so that small inputs exercise the multi-tile /
 * multi-chunk paths.  key: 0 = Lyndon chunk bytes (>= 1), 1 = inverse splitter shift
 * (density 2^-(32-shift), 20..31), 2 = onesweep tile shape (0..4), 3 = disable the
 * warp-local sort path (1), 4 = force the suffix-sort fallback for the Lyndon boundaries (1),
 * 5 = no copy/compute overlap between the blocks of one device (1), 6 = cap on the bits of
 * the initial packed key (8..64), 7 = binned rank scatter of the re-ranks (1 = never, 2 = the first
 * re-rank of inputs of any size, 3 = the first re-rank only, 4 = every re-rank of both rank-ordered sets; default:
 * the first re-rank of inputs of 4 Mi bytes and more, later ones from 128 Mi bytes on when a third of the set's
 * ranks moved in the round before), 8 = never sort the large-group set
 * CTA-locally (1), 9 = emit through rank windows (1), binned by rank region as one packed word per
 * element (2; default from 512 Mi bytes) or as (rank, byte) pairs in two streams (3), 10 = cudaLimitMaxL2FetchGranularity (32 / 64 / 128; measured: no
 * effect), 12 = inverse through two read-only walks (1) instead of the staged single walk, 13 =
 * sublists per warp of the staged walk, 14 = largest group the text-ordered tuple set takes (1 =
 * set switched off, 2..32; default 8), 15 = inverse marks reached elements with one bit each (1) or one
 * count per 128 (2; default: counts above 256 MiB), 16 = step budget of the inverse's fallback walks per attempt
 * (default 32 n), 17 = Lyndon chunk-minimum scan (1 = Hillis-Steele levels, 2 = CTA-wide levels; default:
 * chosen by the match lengths of the first level), 18 = CTA-local sort as the bitonic network of
 * round 1 (1) instead of the radix sort in shared memory, 20 = tuple set with one thread per group and
 * groups of up to 32 (1; measured slower than one thread per member), 21 = digit histograms of the initial sort by a
 * sweep over the keys (1) instead of the window histogram taken while the keys are built, 22 = initial keys of whole
 * symbols only (1; default: spare key bits hold the top bits of the next symbol), 23 = MiB of output per emit window
 * (16..1024; default 64: the scatter target of one sweep stays in L2).  value 0 = default. */
int bwts_b200_tune(int key, long value);

/* ---- "next" row: suffix array behind libdivsufsort's own seam ----- */
/* Same contract as divsufsort(T, SA, n) (/root/reference/mk_bwts_sa.c:48): SA[0..n) =
 * start offsets of the suffixes of T in ascending order; returns 0 or a negative code. */
int bwts_b200_divsufsort(const unsigned char *T, int *SA, int n, int device);

#ifdef __cplusplus
}
#endif

#endif /* BWTS_B200_H */
