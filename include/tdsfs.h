/*
 * tdsfs.h -- C ABI of libtdsfs.so: the B200 (sm_100a) implementation of the 2DSFS-scan hot path.
 *
 * The reference (uricchio/2DSFS-scan) has no FFI; its boundary is the Python surface of
 * scripts/src/twoDSFS_class.py and scripts/sims_scan.py (paths below are relative to the reference
 * repository).  Every entry point here names the reference code whose body it replaces.  The Python
 * drop-in layer (2dsfs-scan_b200/twoDSFS_class.py, sims_scan.py) binds these with ctypes; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - return 0 on success, non-zero (TDSFS_ERR_*) otherwise; tdsfs_last_error() gives a thread-local message.
 *   - input pointers may be HOST or DEVICE memory (detected with cudaPointerGetAttributes); device inputs are
 *     adopted without a copy and must outlive the scan.  Output pointers are caller-owned HOST buffers.
 *   - the library owns every device allocation behind the handle; it never returns memory it allocated.
 *   - calls are synchronous unless stated (they synchronise the handle's stream before returning);
 *     one handle = one GPU, not re-entrant.
 *   - there is NO CPU fallback: without a CUDA device tdsfs_create fails.
 */
#ifndef TDSFS_H
#define TDSFS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tdsfs_ctx tdsfs_t;

enum {
  TDSFS_OK = 0,
  TDSFS_ERR_ARG = 1,    /* bad argument / buffer too small */
  TDSFS_ERR_CUDA = 2,   /* CUDA runtime error (message has the cudaError string) */
  TDSFS_ERR_STATE = 3,  /* call order violated (e.g. scan before background) */
  TDSFS_ERR_RANGE = 4,  /* an allele count exceeds 2n of the declared panel: the reference raises KeyError
                           (twoDSFS_class.py:433) for the same input */
  TDSFS_ERR_RETRY = 5   /* asynchronous pass only: a SNP did not fit the 4-byte per-SNP record; the handle has switched
                           to 8-byte records and the pass must be run again (synchronous calls and tdsfs_run_bp retry
                           by themselves) */
};

/* result flags (per candidate window) */
enum {
  TDSFS_F_T2D_NONE = 1,    /* calculate_likelihood_2D returned None (twoDSFS_class.py:645-647, :668-670) */
  TDSFS_F_T1D_P1_NONE = 2, /* calculate_likelihood_1D returned None for pop1 (:497-499, :520-522) */
  TDSFS_F_T1D_P2_NONE = 4,
  TDSFS_F_EMPTY = 8,       /* window holds no SNP: the reference never emits it (:902 `if window_data`) */
  TDSFS_F_SKIPPED = 16     /* fixed-SNP window whose 2D spectrum sums to 0 (:1496) */
};

/* background modes for tdsfs_background() */
enum {
  TDSFS_BG_NONE = 0,      /* keys only; background supplied later with tdsfs_set_background (scan_precomputed_BG :1161) */
  TDSFS_BG_PER_CHROM = 1, /* every chromosome its own background (combined_scan :809-825, scan_perChr_bySNPs :1450-1460) */
  TDSFS_BG_GENOME = 2,    /* one background over all loaded SNPs (whole-genome usage :1970-1981; the north-star mode) */
  TDSFS_BG_CHROM = 3      /* background = one chromosome (scan_chooseChr :1021-1034) */
};

/* Sparse correction for calls the 2-bit code cannot hold (half calls './1', haploid '1', ...):
 * the sample is stored as MISSING and (dref, dalt) is added to its population's counts, reproducing the
 * per-character counting of make_data_dict_vcf (twoDSFS_class.py:128-129).  Sorted by snp. */
typedef struct {
  int64_t snp;
  int32_t pop; /* 0 = pop1, 1 = pop2 */
  int32_t dref;
  int32_t dalt;
} tdsfs_fixup_t;

/* Struct-of-arrays result, caller-owned host buffers of capacity >= the candidate count
 * (tdsfs_candidates_bp / _snp).  Replaces the dict of dicts built at twoDSFS_class.py:937-945.
 * Any member may be NULL (skipped). */
typedef struct {
  int32_t* chrom;     /* chromosome index (order of chrom_off) */
  int64_t* start;     /* label start: 1 + k*W (fixed-bp) / first SNP pos or previous end + 1 (fixed-SNP, :1527/:1535) */
  int64_t* end;       /* label end: start + W - 1 / last SNP pos */
  int32_t* snp_count; /* count_snps (:291-302) */
  int32_t* n2d;       /* N of the 2D likelihood = SNPs in interior bins (:642) */
  int32_t* n1d_p1;
  int32_t* n1d_p2;
  double* T2D;        /* 2*(ll_fg - ll_bg), +inf when a populated bin has zero background */
  double* T1D_p1;
  double* T1D_p2;
  uint8_t* flags;     /* TDSFS_F_* */
} tdsfs_result_t;

/* ---- lifetime ------------------------------------------------------------------------------------------ */
int tdsfs_create(int device, tdsfs_t** out);
void tdsfs_destroy(tdsfs_t* ctx);
const char* tdsfs_last_error(void);
/* Run the handle's work on an existing CUDA stream (cudaStream_t as void*; NULL = the handle's own stream). */
int tdsfs_set_stream(tdsfs_t* ctx, void* cuda_stream);
/* sync = 0: tdsfs_background / tdsfs_finalize_background / tdsfs_scan_*(out = NULL) only enqueue work on the stream
 * and return; the caller synchronises (used to time the whole pass on the device and to chain the all-reduce). */
int tdsfs_set_sync(tdsfs_t* ctx, int sync);

/* Declared diploid panel sizes -> spectrum shape (2*n1+1) x (2*n2+1); fold = joint minor-allele fold.
 * Replaces the constructor state pop1_size/pop2_size/fold (twoDSFS_class.py:21-33). */
int tdsfs_set_panel(tdsfs_t* ctx, int32_t n1, int32_t n2, int32_t fold);

/* ---- data (the data_dict, twoDSFS_class.py:132-134, in array form) ---------------------------------------
 * SNPs sorted by (chromosome, position); chrom_off[C+1] row offsets of the chromosomes (host memory).
 * snp_flags (optional, S bytes): bit0 = SNP passes the spectrum filters (start/end position :179-182 and
 * variant_type :185-187), bit1 = SNP counts in snp_count (count_snps :298-301).  NULL = all pass. */

/* Counts-level entry: cnt[S][4] = (ref1, alt1, ref2, alt2), the 'calls' tuples of the data_dict. */
int tdsfs_load_counts(tdsfs_t* ctx, const uint16_t* cnt, int64_t S, const int32_t* pos, const int64_t* chrom_off,
                      int32_t C, const uint8_t* snp_flags);

/* Genotype-level entry: 2-bit-per-call matrix, stored as bit planes.  Per SNP: RW = words1 + words2 uint32 words; a
 * population is words/2 pairs (lo word, hi word) of 32 samples each: bit b of the lo / hi word of pair g is the low /
 * high bit of the code of sample 32g + b.  pop1 pairs then pop2 pairs, each population zero padded to a whole pair
 * (words1, words2 even: words = 2 * ceil(samples / 32)).
 * Memory layout "B32" (block transposed): SNPs in blocks of 32; word w of SNP s at uint32 index
 * ((s / 32) * RW + w) * 32 + s % 32; the buffer holds ceil(S / 32) whole blocks (rows beyond S zero).
 * Codes: 0 = 0/0, 1 = 0/1, 3 = 1/1, 2 = missing, so #missing = popcount(hi & ~lo), alt = popcount(all words) - #missing,
 * ref = 2 * (samples - #missing) - alt.  ns1/ns2 = number of sample columns in each block.
 * Replaces the per-sample counting loop of make_data_dict_vcf (:118-130).  Host G is uploaded in chunks
 * asynchronously; the upload overlaps the count kernel of tdsfs_background. */
int tdsfs_load_genotypes(tdsfs_t* ctx, const void* G, int64_t S, int32_t words1, int32_t words2, int32_t ns1,
                         int32_t ns2, const int32_t* pos, const int64_t* chrom_off, int32_t C,
                         const tdsfs_fixup_t* fixups, int64_t n_fixups, const uint8_t* snp_flags);

/* ---- spectra + background --------------------------------------------------------------------------------
 * Derives every SNP's folded 2D bin and raw 1D alt counts (calculate_2d_sfs :190-217, calculate_1d_sfs :427-433)
 * and accumulates the integer background spectra of the requested mode.  bg_chrom: chromosome index for
 * TDSFS_BG_CHROM.  [bg_pos_lo, bg_pos_hi] restricts the SNPs that enter the background (sims_scan.py:664-666
 * uses 0..500000); pass -1, -1 for no restriction. */
int tdsfs_background(tdsfs_t* ctx, int32_t mode, int32_t bg_chrom, int64_t bg_pos_lo, int64_t bg_pos_hi);

/* Number of background groups (C for PER_CHROM, else 1) and the packed device histogram
 * [group][ (2n1+1)(2n2+1) | 2n1+1 | 2n2+1 ] of uint32 counts, for an in-place all-reduce (sum) across GPUs
 * before tdsfs_finalize_background.  n_words = total uint32 words. */
int tdsfs_background_device(tdsfs_t* ctx, void** dev_ptr, int64_t* n_words, int32_t* n_groups);

/* ---- peer-memory exchange of the background (multi-GPU, one process per GPU on one NVLink/NVSwitch node) -----
 * The genome-wide background of a chromosome-sharded scan is the SUM of the ranks' histograms; the reference
 * computes it in one process (calculate_2d_sfs on the whole data_dict, :809-825), so this step has no reference
 * counterpart.  Instead of a library all-reduce, every rank maps every other rank's histogram (CUDA IPC) and one
 * kernel per rank pulls its slice of all histograms over NVLink, sums it and pushes the sum back into all of them,
 * between two flag barriers in peer memory.  Protocol, collectively on all ranks:
 *   tdsfs_background(...)                       once, so that the histogram exists (single group: GENOME / CHROM)
 *   tdsfs_peer_export(ctx, rank, world, blob)   fills blob[TDSFS_PEER_BLOB_BYTES]; exchange blobs out of band
 *   tdsfs_peer_import(ctx, blobs)               blobs = world x TDSFS_PEER_BLOB_BYTES in rank order
 *   per scan: tdsfs_background -> tdsfs_peer_allreduce_background -> tdsfs_finalize_background -> tdsfs_scan_*
 *   tdsfs_peer_close(ctx)                       on every rank BEFORE any rank changes its panel or is destroyed
 * tdsfs_peer_allreduce_background is asynchronous on the handle's stream; all ranks must call it the same number
 * of times.  A rank that never arrives makes the others flag TDSFS_ERR_CUDA (tdsfs_check) after ~60 s (TDSFS_PEER_TIMEOUT_S), not hang. */
#define TDSFS_PEER_BLOB_BYTES 192
#define TDSFS_PEER_MAX_RANKS 16
int tdsfs_peer_export(tdsfs_t* ctx, int32_t rank, int32_t world, void* blob);
int tdsfs_peer_import(tdsfs_t* ctx, const void* blobs);
int tdsfs_peer_allreduce_background(tdsfs_t* ctx);
/* The same exchange and tdsfs_finalize_background in ONE launch per rank: barrier, pull / sum / push of this rank's slice,
 * barrier, then the ln tables of the complete local histogram.  A tdsfs_finalize_background that follows is a no-op. */
int tdsfs_peer_reduce_finalize(tdsfs_t* ctx);
int tdsfs_peer_close(tdsfs_t* ctx);

/* Copy one group's integer spectra to the host: sfs2d[(2n1+1)*(2n2+1)] row-major (i,j) = calculate_2d_sfs,
 * sfs1d_p1[2n1+1] / sfs1d_p2[2n2+1] = calculate_1d_sfs (unfolded).  Any pointer may be NULL. */
int tdsfs_get_background(tdsfs_t* ctx, int32_t group, uint64_t* sfs2d, uint64_t* sfs1d_p1, uint64_t* sfs1d_p2);

/* Precomputed background (scan_precomputed_BG :1161, sims_scan.process_window :451): b2d[(2n1+1)(2n2+1)] and,
 * per population, the values bg[k] seen by the folded foreground keys k = 0..n (n+1 doubles).  Counts or
 * normalised floats.  All chromosomes are scored against it. */
int tdsfs_set_background(tdsfs_t* ctx, const double* b2d, const double* b1d_p1, const double* b1d_p2);

/* Build the log tables the scorer reads (ln b_k, interior totals B).  Call after tdsfs_background (+ all-reduce)
 * or tdsfs_set_background. */
int tdsfs_finalize_background(tdsfs_t* ctx);

/* ---- scans -----------------------------------------------------------------------------------------------
 * Fixed-bp windows (window walk :843-949: window index (pos-1)//W per chromosome, empty windows flagged
 * TDSFS_F_EMPTY) and fixed-SNP windows (:1515-1535: full chunks of N SNPs, partial tail dropped).
 * `out` may be NULL: results then stay on the device (tdsfs_fetch_results copies them later); cap = capacity of
 * the out arrays; *n_windows = number of candidate windows written. */
/* Optional: launch the window-boundary kernel NOW on a side stream (it depends on positions only), so that it overlaps
 * tdsfs_background and the multi-GPU all-reduce; the next tdsfs_scan_* of the same size waits for it instead of launching it. */
int tdsfs_plan_bp(tdsfs_t* ctx, int64_t W);
int tdsfs_plan_snp(tdsfs_t* ctx, int64_t N);
int tdsfs_candidates_bp(tdsfs_t* ctx, int64_t W, int64_t* n_candidates);
int tdsfs_candidates_snp(tdsfs_t* ctx, int64_t N, int64_t* n_candidates);
int tdsfs_scan_bp(tdsfs_t* ctx, int64_t W, tdsfs_result_t* out, int64_t cap, int64_t* n_windows);
int tdsfs_scan_snp(tdsfs_t* ctx, int64_t N, tdsfs_result_t* out, int64_t cap, int64_t* n_windows);
int tdsfs_fetch_results(tdsfs_t* ctx, tdsfs_result_t* out, int64_t cap, int64_t* n_windows);

/* Synchronise the stream and report deferred device-side errors (TDSFS_ERR_RANGE) after asynchronous calls. */
int tdsfs_check(tdsfs_t* ctx);

/* One-call convenience used by the end-to-end benchmark: background(mode) -> finalize -> scan_bp. */
int tdsfs_run_bp(tdsfs_t* ctx, int32_t bg_mode, int64_t W, tdsfs_result_t* out, int64_t cap, int64_t* n_windows);

/* Legacy Poisson composite score of every fixed-bp window, the first-generation script's calculate_p_window
 * (scripts/twoDSFS.py:385-463; the class copy :304-393 cannot run): per window the UNFOLDED 2D spectrum keyed by raw alt
 * counts (twoDSFS.py:211-303: SNPs with both alt counts 0 skipped, a pseudo-count 1/total added to every bin), S_w = the sum
 * of all its bins, and the sum over the bins with S_w q != 0 of poisson.logpmf(int(count), S_w q) (:336-374).
 * q2d[(2n1+1)(2n2+1)] = the normalised background (normalize_2d_sfs :324-334).  Needs tdsfs_set_panel(..., fold = 0) and
 * tdsfs_background (any mode: it writes the per-SNP bins).  Results: T2D = the score, n2d = SNPs counted in the spectrum,
 * snp_count as in the other scans; the 1D fields are zero. */
int tdsfs_set_poisson_background(tdsfs_t* ctx, const double* q2d);
int tdsfs_scan_poisson_bp(tdsfs_t* ctx, int64_t W, tdsfs_result_t* out, int64_t cap, int64_t* n_windows);

/* One whole pass, asynchronous on the handle's stream: plan -> count kernel -> (peer exchange +) ln tables -> finish
 * kernel, results left on the device (tdsfs_fetch_results), deferred errors in tdsfs_check.  Called repeatedly with the
 * same arguments on device-resident data it captures the pass as a CUDA graph on the second call and replays it afterwards.
 * With the peer exchange mapped (tdsfs_peer_import) the exchange is part of the pass: every rank calls it the same number
 * of times. */
int tdsfs_step_bp(tdsfs_t* ctx, int32_t bg_mode, int64_t W);

/* Spectra of one scanned window (calculate_2d_sfs / calculate_1d_sfs on window_data): dense outputs, any NULL. */
int tdsfs_window_spectra(tdsfs_t* ctx, int64_t window, uint64_t* sfs2d, uint64_t* sfs1d_p1, uint64_t* sfs1d_p2);

/* ---- likelihood of explicit spectra ------------------------------------------------------------------------
 * calculate_likelihood_1D / _2D (:478-537, :625-684) on already-interior vectors: x[n] integer counts,
 * b[n] background values, B = sum of b as the caller's language sums it.  *flag = 1 when the reference returns
 * None (sum x == 0 or B == 0). */
int tdsfs_likelihood(tdsfs_t* ctx, const int64_t* x, const double* b, int64_t n, double B, double* T, int32_t* flag);

/* Legacy Poisson composite score, calculate_p (:249-289, twoDSFS.py:336-374): sum over bins with mu != 0 of
 * poisson.logpmf(x, mu), mu = S_w * p_bg. */
int tdsfs_poisson_score(tdsfs_t* ctx, const int64_t* x, const double* mu, int64_t n, double* score);

/* ---- synthetic input (benchmark configs of BASELINE.json) + instrumentation ---------------------------------- */
/* Fill a DEVICE genotype matrix with the synthetic panel of SURVEY.md 8(d): per-SNP ancestral frequency
 * log-uniform, per-population drift, Binomial(2,p) calls, iid missing.  snp0 = global index of row 0. */
int tdsfs_synth_genotypes(tdsfs_t* ctx, void* G_dev, int64_t S, int64_t snp0, int32_t words1, int32_t words2,
                          int32_t ns1, int32_t ns2, uint64_t seed, double missing_rate, double fst);
/* CUDA-event times (ms) of the last background / finalize / scan calls (synchronises the stream):
 * [0]=count kernel (K1), [1]=finalize, [2]=boundaries (K2), [3]=score small windows (K3/K4), [4]=score large windows,
 * [5]=background call, [6]=scan call, [7]=K1 start -> last score kernel end, [8]=peer exchange kernel (n >= 9). */
int tdsfs_timings(tdsfs_t* ctx, float* ms, int32_t n);
int64_t tdsfs_launch_count(tdsfs_t* ctx); /* kernels launched by this handle so far */
/* Which path scored the last scan: *fused = 1 when the fused pair ran (tdsfs_plan_* before tdsfs_background on the
 * genotype-level entry: the count kernel leaves every window's background-independent sums and one finish kernel gathers
 * ln b over the per-SNP records), 0 for the table scorer; *record_bytes = 4 (narrow per-SNP record) or 8. */
int tdsfs_scan_info(tdsfs_t* ctx, int32_t* fused, int32_t* record_bytes);
/* Diagnostics (handles created with TDSFS_TAIL_STAMPS=1 in the environment, else TDSFS_ERR_STATE): SM clock (clock64) of
 * CTA 0 of the last count kernel that ran a tail: [0] kernel entry, [1] its histograms flushed, [2] past the grid barrier,
 * [3] past the first peer barrier, [4] its slice pushed and fenced, [5] past the second peer barrier, [6] ln tables written,
 * [7] totals written.  No reference counterpart. */
int tdsfs_tail_stamps(tdsfs_t* ctx, uint64_t* out8);
int tdsfs_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TDSFS_H */
