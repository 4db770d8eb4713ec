/*
 * include/ldpgta.h -- C-ABI of libldpgta.so, the B200 (sm_100a) implementation of
 * LD-PGTA's windowed haplotype-likelihood engine.
 *
 * The reference (mccoy-lab/LD-PGTA) is pure Python and has no FFI of its own; the
 * entry points below are the operations its hot path performs, cut where the
 * reference's model classes and driver call into big-int arithmetic.  Each one
 * cites the reference code it replaces (paths into the reference repository).
 * The Python classes in ldpgta_b200/ (homogeneous, recent_admixture,
 * distant_admixture, bootstrap, aneuploidy_test) bind these with ctypes; the
 * stub a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative LDPGTA_E_* code;
 *     ldpgta_last_error() returns a thread-local description of the last failure
 *   - plain pointers and sizes only; `const T *` inputs are caller-owned HOST buffers
 *     unless the name ends in _dev, outputs are caller-allocated
 *   - a context owns one CUDA device, one stream and the packed panel rows of one
 *     (sample, chromosome); it is thread-compatible, not thread-safe
 *   - haplotype rows are little-endian arrays of 64-bit words, bit i of word j =
 *     haplotype 64*j+i (any bijection haplotype->bit gives the same popcounts);
 *     bits at or above the group's haplotype count MUST be zero on input
 *   - a "draw" is the ordered tuple of reads sampled from one genomic window by
 *     ANEUPLOIDY_TEST.pick_reads (:226-233); read i of a draw is bit i of the
 *     subset masks that index joint-frequency tables
 */
#ifndef LDPGTA_H
#define LDPGTA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldpgta_ctx ldpgta_ctx;

enum {
    LDPGTA_OK = 0,
    LDPGTA_E_INVALID = -1,     /* bad argument (the message says which)            */
    LDPGTA_E_CUDA = -2,        /* CUDA runtime failure                              */
    LDPGTA_E_UNSUPPORTED = -3, /* outside the implemented envelope  */
    LDPGTA_E_STATE = -4        /* call order: rows/masks not uploaded yet           */
};

/* Model classes: HOMOGENOUES_MODELS.homogeneous, RECENT_ADMIXTURE_MODELS.recent_admixture,
 * DISTANT_ADMIXTURE_MODELS.distant_admixture (selection: ANEUPLOIDY_TEST.py:263-273). */
enum { LDPGTA_HOMOGENEOUS = 0, LDPGTA_RECENT_ADMIXTURE = 1, LDPGTA_DISTANT_ADMIXTURE = 2 };

#define LDPGTA_MAX_GROUPS 8
#define LDPGTA_MAX_READS_PER_DRAW 18   /* reads per draw handled on device = the largest model table the reference ships (MODELS18.p) */
#define LDPGTA_MAX_PEERS 16             /* GPUs of one node that can share a chromosome (ldpgta_comm_*) */
#define LDPGTA_COMM_HANDLE_BYTES 128    /* opaque handle of a rank's exchange buffer */
#define LDPGTA_NUM_LLR_PAIRS 5         /* (BPH,SPH) (disomy,monosomy) (BPH,disomy) (disomy,SPH) (SPH,monosomy) */

const char *ldpgta_last_error(void);
int ldpgta_version(void);
int ldpgta_device_count(void);

/* One context per (sample, chromosome) and per model object: replaces the state built by
 * homogeneous.__init__ (HOMOGENOUES_MODELS.py:52-73) and twins.  n_hap[g] = number of
 * haplotypes in reference-panel group g (number_of_haplotypes_per_group). */
int ldpgta_create(int device, int n_groups, const uint32_t *n_hap, ldpgta_ctx **out);
int ldpgta_destroy(ldpgta_ctx *ctx);
/* words of a caller-side row of group g = ceil(n_hap[g] / 64) */
int ldpgta_row_words(const ldpgta_ctx *ctx, int group);

/* hap_dict (build_hap_dict, HOMOGENOUES_MODELS.py:76-100): one row per observed
 * (pos, base) allele, rows[n_rows][ldpgta_row_words].  complement[i] != 0 means the
 * allele is the REF allele: the device stores row ^ (2^N - 1) (:94); pass NULL for none.
 * Every group must receive the same number of rows, in the same allele order. */
int ldpgta_upload_allele_rows(ldpgta_ctx *ctx, int group, const uint64_t *rows, int64_t n_rows,
                              const uint8_t *complement);

/* A reference panel resident on the device, shared by any number of contexts (samples) on that device:
 * the HAP table of MAKE_REF_PANEL (one row per legend SNP and group, MAKE_REF_PANEL.py:352-368) packed once.
 * ldpgta_gather_allele_rows then performs build_hap_dict's lookup (HOMOGENOUES_MODELS.py:84-94) as a device
 * gather: allele row a of every group = panel row legend_idx[a], complemented inside N bits when
 * complement[a] != 0 (the observed base is the REF allele).  Equivalent to ldpgta_upload_allele_rows without
 * re-packing and re-uploading panel rows for every sample. */
typedef struct ldpgta_panel ldpgta_panel;
int ldpgta_panel_create(int device, int n_groups, const uint32_t *n_hap, int64_t n_snps, ldpgta_panel **out);
int ldpgta_panel_upload(ldpgta_panel *panel, int group, int64_t first_snp, const uint64_t *rows, int64_t n_rows);
int ldpgta_panel_destroy(ldpgta_panel *panel);
int ldpgta_gather_allele_rows(ldpgta_ctx *ctx, const ldpgta_panel *panel, const int64_t *legend_idx,
                              const uint8_t *complement, int64_t n_alleles);

/* build_intrenal_hap_dict (HOMOGENOUES_MODELS.py:102-127) hoisted out of the draw loop:
 * mask of read r = AND of the allele rows allele_idx[read_offsets[r] .. read_offsets[r+1]),
 * computed on device for every group.  Replaces any earlier read masks. */
int ldpgta_build_read_masks(ldpgta_ctx *ctx, const int32_t *allele_idx, const int64_t *read_offsets,
                            int64_t n_reads);

/* Read scores of supporting_dictionaries.build_score_dict (ANEUPLOIDY_TEST.py:151-189), used by
 * build_gw_dict to filter reads (min_score, :215).  Reads are given like in ldpgta_build_read_masks;
 * alpha[g] = ancestral share of group g / its haplotype count (:162), score_weight = the weight the
 * reference applies to every group's haplotype count in :181-184 (the last group's alpha),
 * min_HF as in the CLI.  out_scores[n_reads]; needs the allele rows of every group. */
int ldpgta_read_scores(ldpgta_ctx *ctx, const int32_t *allele_idx, const int64_t *read_offsets, int64_t n_reads,
                       const double *alpha, double score_weight, double min_HF, int32_t *out_scores);

/* Alternative to the two calls above: upload ready-made read masks. */
int ldpgta_upload_read_masks(ldpgta_ctx *ctx, int group, const uint64_t *rows, int64_t n_rows);
/* Read masks straight from device memory laid out [n_rows][padded_words] with
 * padded_words = ldpgta_padded_row_words(); the context keeps the pointer, no copy. */
int ldpgta_padded_row_words(const ldpgta_ctx *ctx, int group);
int ldpgta_attach_read_masks_dev(ldpgta_ctx *ctx, int group, const uint64_t *rows_dev, int64_t n_rows);
int64_t ldpgta_num_reads(const ldpgta_ctx *ctx);

/* joint_frequencies_combo(normalize=False) (HOMOGENOUES_MODELS.py:129-163) for ONE draw of
 * n reads (2..18): out_counts[g][S] for S = 0 .. 2^n-1 (S = subset bitmask, entry 0 unused = 0).
 * Parity hook: counts are bit-exact. */
int ldpgta_joint_counts(ldpgta_ctx *ctx, const int32_t *read_idx, int n, uint32_t *out_counts);

/* One call of model.get_likelihoods(alleles) / model.likelihoods(alleles) on an arbitrary tuple of
 * "haplotypes" (HOMOGENOUES_MODELS.py:272-286 / :185-225): haplotype i is the list of allele rows
 * allele_idx[read_offsets[i] .. read_offsets[i+1]) (build_intrenal_hap_dict :102-127 runs on device
 * into scratch rows).  general_formula != 0 forces the partition polynomials of likelihoods() even
 * for 2..4 haplotypes.  out[4]; counts_out (optional) [G][2^n] as in ldpgta_joint_counts. */
int ldpgta_get_likelihoods(ldpgta_ctx *ctx, int model_kind, const int32_t *allele_idx,
                           const int64_t *read_offsets, int n, const double *proportions,
                           int general_formula, double *out, uint32_t *counts_out);

/* get_likelihoods (HOMOGENOUES_MODELS.py:272-286, RECENT_...:314-328, DISTANT_...:300-316)
 * for a batch of draws = the body of the bootstrap loop ANEUPLOIDY_TEST.py:277-286.
 *   read_idx      concatenated read indices (rows of the read-mask table)
 *   draw_offsets  n_draws+1 offsets into read_idx; draw d has 2..18 reads
 *   proportions   LDPGTA_DISTANT_ADMIXTURE only: ancestral proportion per group, in the
 *                 order of the ancestral_makeup dict (sum order of DISTANT_...:138-139)
 *   out           [n_draws][4] = monosomy, disomy, SPH, BPH (likelihoods_tuple order)
 * Dispatch follows the reference: 2,3,4 reads use the closed forms on normalised
 * frequencies, other sizes the general partition polynomials on integer counts. */
int ldpgta_likelihoods_batch(ldpgta_ctx *ctx, int model_kind, const int32_t *read_idx,
                             const int64_t *draw_offsets, int64_t n_draws, const double *proportions,
                             double *out);
/* The same for draws that all hold n reads (n = min(R-1, max_reads), ANEUPLOIDY_TEST.py:229, is
 * one value for all windows with R > max_reads): read_idx[n_draws][n], no offsets.  With
 * page-locked buffers (ldpgta_host_alloc) the call streams chunks through upload -> index
 * validation + kernel -> download on three streams, and a call that repeats the previous call's
 * shape and buffers replays the whole pipeline as one CUDA graph; this is the call bench.py times
 * for `e2e`. */
int ldpgta_likelihoods_uniform(ldpgta_ctx *ctx, int model_kind, int n, const int32_t *read_idx,
                               int64_t n_draws, const double *proportions, double *out);

/* The bootstrap loop with the draws made on the device (pick_reads, ANEUPLOIDY_TEST.py:226-233, for all
 * windows at once): window w holds the read rows window_reads[window_offsets[w] .. window_offsets[w+1])
 * and gets draw_offsets[w+1] - draw_offsets[w] draws of n distinct reads each (uniform over n-subsets;
 * Philox4x32-10 counter-based generator, draw d is a pure function of (seed, d), so results do not
 * depend on the launch shape).  Only the window lists travel up, not the draws: at 1000 subsamples
 * per window that is 30x less than the indices.  out[n_draws][4]; idx_out (optional) [n_draws][n]
 * returns the draws.  Not the reference's MT19937 stream -- use ldpgta_likelihoods_batch to replay it. */
int ldpgta_bootstrap_uniform(ldpgta_ctx *ctx, int model_kind, int n, const int64_t *window_offsets,
                             const int32_t *window_reads, int64_t n_windows, const int64_t *draw_offsets,
                             uint64_t seed, const double *proportions, double *out, int32_t *idx_out);

/* Same computation for draws of one size n with everything already resident in HBM:
 * read_idx_dev[n_draws][n], out_dev[n_draws][4].  Asynchronous on the context's stream;
 * ldpgta_synchronize() waits.  This is the call bench.py times for `value`. */
int ldpgta_likelihoods_uniform_dev(ldpgta_ctx *ctx, int model_kind, int n, const int32_t *read_idx_dev,
                                   int64_t n_draws, const double *proportions, double *out_dev);
int ldpgta_synchronize(ldpgta_ctx *ctx);
/* cudaStream_t of the context (as an integer) so callers can record events on it. */
uint64_t ldpgta_stream(const ldpgta_ctx *ctx);
/* number of kernels this context has launched since creation */
int64_t ldpgta_launch_count(const ldpgta_ctx *ctx);
/* number of host-pipeline calls that ran as one CUDA graph launch (see ldpgta_likelihoods_uniform) */
int64_t ldpgta_graph_replays(const ldpgta_ctx *ctx);

/* ---- the bootstrap loop and statistics() as one device-resident pipeline -------------------------------------
 * The reference's bootstrap() returns the likelihoods of every draw and statistics() consumes them in place
 * (ANEUPLOIDY_TEST.py:277-287 -> :289-325).  The calls below keep that hand-off inside HBM: the likelihoods go from the
 * draw kernels to the LLR / window / chromosome reductions without visiting the host, and what comes back is what the
 * caller asks for -- 80 bytes per window and a few hundred per chromosome, or additionally the per-draw likelihoods.
 *
 * ldpgta_set_windows makes the result of build_gw_dict (:191-224) and the draw plan (:226-248) resident:
 *   window_offsets, window_reads  CSR of read rows per window (window w = window_reads[window_offsets[w] ..
 *                                 window_offsets[w+1])); window_reads NULL = the rows window_offsets[w] ..
 *                                 window_offsets[w+1] themselves (read masks stored window by window; window_offsets[0]
 *                                 may then be any row, e.g. the first window of a shard);
 *                                 window_offsets NULL = a replay-only plan (ldpgta_replay_stats)
 *   draw_offsets    [n_windows+1] draws per window = effective_number_of_subsamples (:235-248), at least one each
 *                   (windows the reference skips are not listed)
 *   reads_per_draw  [n_windows]   min(R - 1, max_reads) of :231, 2..18
 *   segment_offsets [n_segments+1] windows of each (embryo, chromosome) unit statistics() reduces together;
 *                   NULL = one segment.  Several chromosomes or embryos can share one context this way.
 *   segment_cols    [n_segments] columns of zip(*A) every window of the WHOLE chromosome has (:73); NULL = the minimum
 *                   over the listed windows.  segment_first [n_segments] = len(A[0]) (:71); NULL = the draws of the
 *                   segment's first listed window.  Ranks that hold a shard of a chromosome pass the global values.
 * Replaces any earlier plan of the context; the read masks must be resident. */
int ldpgta_set_windows(ldpgta_ctx *ctx, const int64_t *window_offsets, const int32_t *window_reads, int64_t n_windows,
                       const int64_t *draw_offsets, const int32_t *reads_per_draw, const int64_t *segment_offsets,
                       int n_segments, const int32_t *segment_cols, const int32_t *segment_first);
/* bootstrap() + statistics() for the resident plan, draws made on the device (Philox4x32-10 + Floyd as in
 * ldpgta_bootstrap_uniform; the draws of size class n use seed + n and are numbered in window order).  Nothing but the
 * seed goes up.  Every output is optional (NULL = not wanted, not downloaded):
 *   lik_out          [n_draws][4]          per-draw likelihoods (streamed down in chunks under the kernels)
 *   idx_out          concatenated draws in window order, reads_per_draw[w] indices per draw
 *   llr_out          [5][n_draws]          per-draw LLRs; also left resident for ldpgta_bin_stats(x = NULL)
 *   window_mean_var  [5][n_windows][2]     mean and sample variance (NaN for single-draw windows)
 *   segment_partials [n_segments][5][ncp+3] as chrom_partials of ldpgta_window_stats, ncp = max segment_cols, columns at
 *                                          or beyond a segment's own count are zero
 *   segment_stats    [n_segments][5][3]    mean_of_mean, std_of_mean, fraction_of_negative_LLRs (:311-313)
 * Memory: a one-group plan of 4-read draws with 256 or more draws per window on average (e.g. 1000 subsamples) whose
 * windows are read lists keeps, from its first call on, a window-ordered copy of the listed read masks (644 bytes per
 * listed read + 16 per draw) for the window-resident draw kernel; the copy is skipped when it would not leave 1 GiB free,
 * refreshed when the read masks change, and released with the plan.  Results do not depend on it. */
int ldpgta_bootstrap_stats(ldpgta_ctx *ctx, int model_kind, uint64_t seed, const double *proportions, double *lik_out,
                           int32_t *idx_out, double *llr_out, double *window_mean_var, double *segment_partials,
                           double *segment_stats);
/* The same for draws chosen by the caller (the reference's random.sample stream, :231): read_idx = the draws of the
 * plan in window order, reads_per_draw[w] read rows per draw. */
int ldpgta_replay_stats(ldpgta_ctx *ctx, int model_kind, const int32_t *read_idx, const double *proportions,
                        double *lik_out, double *llr_out, double *window_mean_var, double *segment_partials,
                        double *segment_stats);
/* Device-resident variants, asynchronous on the context's stream (bench.py `value`, multi-GPU shards): the first leaves
 * likelihoods and window statistics in the plan's buffers and writes the segment partial sums to
 * segment_partials_dev[n_segments][5][ncp+3] (NULL = the plan's own buffer, then the statistics are finished too); shards
 * of one chromosome all-reduce (sum) that buffer and call the second, which turns partials into segment_stats_dev
 * [n_segments][5][3] (NULL = the plan's buffers). */
int ldpgta_bootstrap_stats_dev(ldpgta_ctx *ctx, int model_kind, uint64_t seed, const double *proportions,
                               double *segment_partials_dev);
int ldpgta_finish_segments_dev(ldpgta_ctx *ctx, const double *segment_partials_dev, double *segment_stats_dev);
/* ---- the cross-GPU exchange: the sum of the ranks' segment (chromosome) partial sums ----------------------------------
 * One process per GPU.  When the windows of one chromosome are dealt to several GPUs, the column sums of
 * mean_and_std_of_mean_of_rnd_var (ANEUPLOIDY_TEST.py:65-76) have to be added across them; nothing else of the path
 * crosses GPUs.  The sum is ONE kernel over NVLink / NVSwitch peer memory (store the partials into every peer's buffer,
 * flag the peers, poll locally, add the world's slots from own memory in rank order: every rank gets bit-identical sums);
 * no library collective is involved.
 *   ldpgta_comm_create   allocates this rank's exchange buffer for up to max_doubles values and writes its handle
 *                        (LDPGTA_COMM_HANDLE_BYTES bytes, to be passed to every peer by any byte transport)
 *   ldpgta_comm_connect  handles[world][LDPGTA_COMM_HANDLE_BYTES] of all ranks, in rank order (CUDA IPC between
 *                        processes, direct peer access between contexts of one process)
 *   ldpgta_allreduce_chrom_dev  partials_dev[n_doubles] (device memory, e.g. the segment_partials_dev of
 *                        ldpgta_bootstrap_stats_dev) becomes the sum over ranks, in place, asynchronously on the context's
 *                        stream; every rank must call it the same number of times with the same n_doubles.  A peer that
 *                        does not arrive within ~2 s makes the NEXT call fail (the kernel never spins forever). */
int ldpgta_comm_create(ldpgta_ctx *ctx, int rank, int world, int64_t max_doubles, unsigned char *handle_out);
int ldpgta_comm_connect(ldpgta_ctx *ctx, const unsigned char *handles);
int ldpgta_allreduce_chrom_dev(ldpgta_ctx *ctx, double *partials_dev, int64_t n_doubles);
int ldpgta_comm_destroy(ldpgta_ctx *ctx);

/* Measurement aid: with enable != 0 the calls above bracket their draw kernels (k_small / k_mid / k_lanes / k_general, not
 * the sampler or the reductions) with CUDA events on the context's stream; ldpgta_kernel_times waits for the stream,
 * writes the elapsed milliseconds of every bracket since the last call (at most max_entries) and returns their number.
 * enable == 2 brackets the reductions of ldpgta_bootstrap_stats_dev (window statistics + segment sums) instead, 3 its
 * window-statistics kernel alone, 4 its sampler.  Bits 8..23 of enable = stride k: only every k-th bracket is recorded (a pair
 * of events costs the stream about 5 us, 1 % of a configs[1] step). */
int ldpgta_kernel_timing(ldpgta_ctx *ctx, int enable);
int ldpgta_kernel_times(ldpgta_ctx *ctx, double *ms_out, int max_entries);
/* out[10] = n_windows, n_draws, n_segments, ncp, number of draw-size classes, total read indices of the plan, and the
 * device addresses of the plan's likelihoods [n_draws][4], window_mean_var, segment_partials, segment_stats buffers. */
int ldpgta_plan_info(const ldpgta_ctx *ctx, int64_t *out);

/* LLR + reductions of statistics() (ANEUPLOIDY_TEST.py:78-90, 51-56, 65-76, 289-325) for one
 * chromosome shard.  lik[n_draws][4] as produced above, windows = consecutive runs of draws:
 * window w owns draws [window_offsets[w], window_offsets[w+1]).
 *   llr            [5][n_draws]        per-draw log-likelihood ratios (sentinels +-1.23456789, 0)
 *   window_mean_var[5][n_windows][2]   mean and SAMPLE variance per window (variance = NaN when
 *                                      a window has a single draw; the reference raises there)
 *   chrom_partials [5][n_cols + 3]     per pair: [0] sum over windows of (sum_i x_wi),
 *                                      [1] number of windows, [2] number of windows with mean < 0,
 *                                      [3 + i] column sums  sum_w x_wi  for i < n_cols
 * n_cols = the smallest draw count over the windows of the WHOLE chromosome (zip truncation of
 * :73); shards of one chromosome add their chrom_partials (that sum is the only cross-GPU
 * exchange of the path) and ldpgta_finish_chromosome() turns the sum into mean/std.  Likelihoods that are
 * already on the device never need this call: see ldpgta_bootstrap_stats / ldpgta_replay_stats. */
int ldpgta_window_stats(ldpgta_ctx *ctx, const double *lik, const int64_t *window_offsets,
                        int64_t n_windows, int n_cols, double *llr, double *window_mean_var,
                        double *chrom_partials);
/* mean_and_std_of_mean_of_rnd_var (ANEUPLOIDY_TEST.py:65-76) from summed partials; pure host
 * arithmetic.  first_window_draws = len(A[0]) of the reference.  out[5][3] = mean_of_mean,
 * std_of_mean, fraction_of_negative_LLRs. */
int ldpgta_finish_chromosome(const double *chrom_partials, int n_cols, int first_window_draws, double *out);

/* binning() of the LLR consumers (PLOT_PANEL.py:121-144, BALANCED_ROC_CURVE.py:131-154): per bin
 * = run of windows [bin_ranges[b][0], bin_ranges[b][1]) the mean LLR and the standard error of
 * mean_and_std_of_mean_of_rnd_var (PLOT_PANEL.py:37-48; N = draws of the bin's first window,
 * columns truncated to the bin's shortest window).
 *   x          [n_series][n_draws] host values, one LLR series per row, windows as in
 *              ldpgta_window_stats; or NULL = the 5 LLR series the last ldpgta_window_stats call
 *              of this context left on the device (n_series must be 5, same window_offsets),
 *              so the per-draw LLRs never travel back to the host;
 *   bin_ranges [n_bins][2]  window index ranges, every bin non-empty;
 *   out        [n_series][n_bins][2] = (mean, std). */
int ldpgta_bin_stats(ldpgta_ctx *ctx, const double *x, int n_series, const int64_t *window_offsets,
                     int64_t n_windows, const int64_t *bin_ranges, int64_t n_bins, double *out);

/* Page-locked host buffers so that the batch calls above copy without staging. */
int ldpgta_host_alloc(void **ptr, int64_t bytes);
int ldpgta_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* LDPGTA_H */
