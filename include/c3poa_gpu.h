/*
 * c3poa_gpu.h -- C ABI of the B200-native C3POa per-read consensus hot path.
 *
 * The reference (rvolden/C3POa, 100 % Python) has no FFI of its own: the path
 * sits behind three Python call sites.  Each entry point below replaces one of
 * them, batch-wise, and is what a ctypes/Cython binding in the reference would
 * bind (INTEGRATION.md shows the stubs):
 *
 *   c3_conk_batch       <- conk.conk(splint, seq, penalty)            C3POa.py:123
 *   c3_peaks_batch      <- call_peaks(scores, min_dist, iters, w, o)  C3POa.py:124, bin/call_peaks.py:8-16
 *   c3_poa_batch        <- poa.msa_aligner(match=5).msa(seqs, ...)    bin/determine_consensus.py:30-47
 *   c3_consensus_batch  <- body of `for read in reads:` of analyze_reads up to the
 *                          pre-racon consensus                         C3POa.py:112-165
 *   c3_stage/c3_run/c3_fetch  = c3_consensus_batch split in three so that a
 *                          caller can overlap copies with compute (and so the
 *                          bench can time the device-resident region alone).
 *
 * Conventions: plain pointers and sizes, caller-owned HOST buffers, int return
 * codes (0 = ok, <0 = error; c3_last_error() gives the text), no exceptions and
 * no callbacks across the boundary.  One handle per device; a handle is not
 * thread-safe; distinct handles may be driven from distinct host threads.
 * There is NO CPU fallback: every entry point fails if the device is unusable.
 */
#ifndef C3POA_GPU_H
#define C3POA_GPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct c3_handle c3_handle;

/* pyabpoa.msa_aligner keyword arguments (reference overrides only match=5) */
typedef struct {
    int32_t match, mismatch;        /* 5, 4 */
    int32_t gap_open1, gap_ext1;    /* 4, 2 */
    int32_t gap_open2, gap_ext2;    /* 24, 1 */
    int32_t wb;                     /* extra_b = 10 */
    int32_t simd_bits;              /* 256: band granule of the AVX2 build (16 x int16 / 8 x int32) */
    double  wf;                     /* extra_f = 0.01 */
} c3_poa_params;

/* per-read result of the fused path */
typedef struct {
    int32_t status;     /* 0 consensus produced (n_sub>=3 POA, or n_sub==1 copy);
                           1 read skipped: no peaks                (C3POa.py:125-126,131-132);
                           2 n_sub==2: pairwise path, out_cons holds the two abPOA MSA rows [row0|row1],
                             cons_len columns each (bin/determine_consensus.py:33-41);
                             n_sub==0: zero-repeat path, bounds only;
                           <0 device-side error (workspace overflow ...), never silent */
    int32_t n_peaks;    /* after shift/filter (C3POa.py:127-130) */
    int32_t n_sub;      /* repeats */
    int32_t n_dang;
    int32_t cons_len;
    int32_t poa_nodes;  /* final graph size */
    int64_t poa_cells;  /* banded DP cells computed for this read (for GCUPS) */
} c3_read_result;

/* stage timings of the last c3_run / *_batch call, CUDA events on the launch stream */
typedef struct {
    float encode_ms, conk_ms, peaks_ms, split_ms, poa_ms, total_ms;
    int32_t kernel_launches;   /* kernels launched by that call */
    int32_t poa_items;         /* reads that went through the POA kernel */
    /* POA stage by kernel (sums of per-launch event pairs on the launch stream; they add up to poa_ms less the host's
     * work-order preparation): group DP kernel, group graph kernels (init + per-alignment), warp-per-read kernel,
     * lane kernel */
    float poa_dp_ms, poa_graph_ms, poa_warp_ms, poa_lane_ms;
    int32_t poa_dp_launches, poa_graph_launches;
} c3_timings;

const char *c3_version(void);
int  c3_device_count(void);
int  c3_init(int device_ordinal, c3_handle **out);
void c3_destroy(c3_handle *h);
const char *c3_last_error(const c3_handle *h);
void c3_default_poa_params(c3_poa_params *p);
int  c3_get_timings(const c3_handle *h, c3_timings *out);
/* Which POA kernel serves B3/B4.  0 = auto: the group path (c3_poa_grp_dp_kernel + c3_poa_graph_kernel: 8 lanes per
 * read for the DP, one thread per read for the graph phases) takes the reads with a mean subread length <= 2 600,
 * <= 32 subreads and int16 score mode when a batch holds at least 12 000 of them; everything else, and whatever the
 * group path declines, runs through the warp-per-read kernel in the same call.  1 = warp kernel only.  2 = the
 * thread-per-read "lane" kernel of round 1 whenever a read is eligible.  3 = the group path whenever a read is eligible
 * (no minimum batch, no length limit).  Results are identical in every mode.                                     */
int  c3_set_poa_mode(c3_handle *h, int32_t mode);
/* Named switches of the abPOA restatement, both 0 by default: two upstream branches that are
 * recalled from abPOA 1.0.5 but could not be checked offline.  int8_lanes: the band granule becomes simd_bits/8 when the
 * score bound fits int8 (only sequences under ~25 nt); end_clamp: a row's band end is clamped to one SIMD vector past
 * the largest predecessor band end.  With either set every read runs through the warp-per-read kernel.  The oracle has
 * the same two switches (c3o_poa_para_t).  Replaces nothing in the reference: pins pyabpoa 1.0.5's behaviour
 * (/root/reference/setup.sh:8, bin/determine_consensus.py:30) once it can be compared.                          */
int  c3_set_abpoa_switches(c3_handle *h, int32_t int8_lanes, int32_t end_clamp);
/* Reads of the last B3/B4 call handed to the fast kernel of the mode (group path / lane kernel), and how many of them
 * it finished (the rest went to the warp kernel).                                                              */
int  c3_lane_counts(c3_handle *h, int32_t *out_given, int32_t *out_done);

/* B1.  reads: concatenated ASCII, read_off[n_reads+1]; splints likewise
 * (already strand-resolved, splint_idx[r] selects one).  out_profile has the
 * same CSR layout as reads: one int32 per read position (diagonal).         */
int c3_conk_batch(c3_handle *h, int32_t n_reads, const char *reads, const int64_t *read_off,
                  int32_t n_splints, const char *splints, const int32_t *splint_off,
                  const int32_t *splint_idx, int32_t penalty, int32_t *out_profile);

/* B2.  profile: int32 CSR (off[n+1]).  coef: `window` SG coefficients computed
 * by the caller with the reference's own pinv line.  out_smoothed (optional,
 * may be NULL) is float64 CSR like profile; out_median optional [n];
 * out_peaks is [n][max_peaks]; out_n_peaks [n] (a count > max_peaks or <0
 * flags an error for that read).                                            */
int c3_peaks_batch(c3_handle *h, int32_t n, const int32_t *profile, const int64_t *off,
                   const double *coef, int32_t window, int32_t iters, int32_t min_dist,
                   double height_mult, double gate_mult, double *out_smoothed,
                   double *out_median, int32_t *out_peaks, int32_t max_peaks,
                   int32_t *out_n_peaks);

/* B3.  n_groups independent msa() calls.  seqs: concatenated ASCII,
 * seq_off[n_seqs+1]; group_off[n_groups+1] indexes seqs.  out_cons is
 * [n_groups][cons_cap] ASCII; out_msa (optional) is [n_seqs][msa_cap] with
 * out_msa_len[n_groups]; only groups of exactly 2 sequences get MSA rows.   */
int c3_poa_batch(c3_handle *h, int32_t n_groups, const char *seqs, const int64_t *seq_off,
                 const int32_t *group_off, const c3_poa_params *params, char *out_cons,
                 int32_t cons_cap, int32_t *out_cons_len, int64_t *out_cells,
                 int32_t *out_nodes, int32_t *out_status, char *out_msa, int32_t msa_cap,
                 int32_t *out_msa_len);

/* B4, split in three.  c3_stage copies inputs host->device; c3_run launches the
 * kernels (conk -> SG/peaks -> split -> POA) and waits; c3_fetch copies the
 * results device->host.  Output layout: out_peaks [n][max_peaks] (shifted,
 * filtered), out_sub_bounds [n][max_peaks][2], out_dang_bounds [n][2][2],
 * out_cons [n][cons_cap].                                                   */
int c3_stage(c3_handle *h, int32_t n_reads, const char *reads, const int64_t *read_off,
             int32_t n_splints, const char *splints, const int32_t *splint_off,
             const int32_t *splint_idx);
int c3_run(c3_handle *h, int32_t penalty, const double *coef, int32_t window, int32_t iters,
           int32_t min_dist, const c3_poa_params *params, int32_t max_peaks, int32_t cons_cap);
int c3_fetch(c3_handle *h, int32_t *out_peaks, int32_t *out_sub_bounds, int32_t *out_dang_bounds,
             char *out_cons, c3_read_result *out_results);
int c3_consensus_batch(c3_handle *h, int32_t n_reads, const char *reads, const int64_t *read_off,
                       int32_t n_splints, const char *splints, const int32_t *splint_off,
                       const int32_t *splint_idx, int32_t penalty, const double *coef,
                       int32_t window, int32_t iters, int32_t min_dist,
                       const c3_poa_params *params, int32_t max_peaks, int32_t cons_cap,
                       int32_t *out_peaks, int32_t *out_sub_bounds, int32_t *out_dang_bounds,
                       char *out_cons, c3_read_result *out_results);

/* Output records of a batch, host code: for the reads with group[i] == which_group (group may be NULL:
 * all reads) and status 0, appends ">{name}_{avg_qual}_{len}_{repeats}_{cons_len}\n{cons}\n" to out_cons
 * (/root/reference/C3POa.py:167-173; avg_qual formatted like Python's str(round(x, 2))) and the subread / dangling
 * FASTQ records "@{name}_{k}\n{seq}\n+\n{qual}\n" to out_sub (/root/reference/bin/determine_consensus.py:57-77).
 * names: NUL-terminated, name_off[i] = start of read i's name (as c3_fastq_next returns them); the other arrays are
 * the inputs / outputs of c3_consensus_batch.  stats[4] (optional): records written, reads without peaks, reads left
 * to the caller (status 2: pairwise / zero-repeat paths), reads with errors.  Returns 0, -2 if a buffer is too small
 * (size them with 2 x bases + (repeats + 2) x (name + 32) per read), -3 on inconsistent bounds.                  */
int c3_format_batch(int32_t n_reads, const char *names, const int64_t *name_off, const char *seq, const char *qual,
                    const int64_t *off, const int64_t *qual_sum, const c3_read_result *res,
                    const int32_t *sub_bounds, const int32_t *dang_bounds, int32_t max_peaks, const char *cons,
                    int32_t cons_cap, const int32_t *group, int32_t which_group, char *out_cons, int64_t out_cons_cap,
                    int64_t *out_cons_len, char *out_sub, int64_t out_sub_cap, int64_t *out_sub_len, int64_t *stats);

/* Pinned (page-locked) host buffers for callers that want full-speed H2D/D2H
 * (any host pointer is accepted by the batch calls; pageable memory is staged
 * by the driver).  Returns NULL on failure.                                  */
void *c3_host_alloc(size_t bytes);
void  c3_host_free(void *p);

/* Splint assignment on the GPU: the conk profile of every candidate (splint x strand,
 * concatenated ASCII with cand_off[n_cands+1]) against every read; a candidate's score is the maximum of its
 * raw profile, out_best[r] the arg-max (lowest index on ties), out_scores (optional) is [n_cands][n_reads].
 * Stands in for the BLAT pre-step (bin/preprocess.py:12-45,74-76); it is NOT BLAT-equivalent: the acceptance
 * criterion is agreement with the known splint/strand of synthetic reads (tests), thresholding is the caller's. */
int c3_assign_splints(c3_handle *h, int32_t n_reads, const char *reads, const int64_t *read_off,
                      int32_t n_cands, const char *cands, const int32_t *cand_off, int32_t penalty,
                      int32_t *out_best, int32_t *out_scores);

/* Ingest: FASTQ/FASTA (plain or gzip) -> length filter -> packed batch in the layout c3_stage()
 * takes; replaces the reference's mappy.fastx_read passes (C3POa.py:201-206,239-254).  c3_fastq_next fills up to
 * max_reads reads / max_bases bases: seq (and qual if non-NULL) concatenated with off[n+1], NUL-terminated names
 * with name_off[n+1], per-read Phred sums (header quality, C3POa.py:168); reads shorter than min_len are skipped
 * and counted in *n_short.  Returns the number of reads (0 at end of file) or <0.                              */
typedef struct c3_fastq c3_fastq;
int  c3_fastq_open(const char *path, c3_fastq **out);
int  c3_fastq_next(c3_fastq *fq, int32_t max_reads, int64_t max_bases, int32_t min_len, char *seq, char *qual,
                   int64_t *off, char *names, int64_t names_cap, int64_t *name_off, int64_t *qual_sum,
                   int64_t *n_short);
void c3_fastq_close(c3_fastq *fq);

/* Micro-benchmark used for the integer-pipe roofline denominator: independent
 * VIADDMNMX chains on every SM, CUDA-event timed.  Returns int-ops/s.        */
int c3_measure_int_peak(c3_handle *h, double *out_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif
