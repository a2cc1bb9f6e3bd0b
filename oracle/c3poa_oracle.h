/*
 * c3poa_oracle.h -- CPU restatement of C3POa's per-read consensus hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (c3poa_b200/, the C-ABI
 * library) may include, link or call this.  Allowed users: tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status:
 *   stage 2  (savitzky_golay, call_peaks)  -- PINNED against the reference's own
 *            Python (bin/call_peaks.py, bin/savitzky_golay.py) via tests/golden/.
 *   stage 3a (peak shift / subread split)  -- PINNED against C3POa.py:127-155.
 *   stage 1  (conk)   -- PARITY UNPINNED: conk (github rvolden/conk, unpinned HEAD,
 *            setup.sh:12) is not vendored in /root/reference; restated from its
 *            call contract (C3POa.py:123-130) and published description.
 *   stage 3b (abPOA 1.0.5) -- PARITY UNPINNED: pyabpoa==1.0.5 (setup.sh:8) is not
 *            vendored; restated from the published abPOA algorithm (adaptive
 *            banded convex-gap POA + heaviest bundling).
 */
#ifndef C3POA_ORACLE_H
#define C3POA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- stage 1: conk(splint, seq, penalty)  (call site C3POa.py:123) ---- */
#define C3O_CONK_MATCH     5   /* EDNAFULL lineage (water -> gonk -> conk)  */
#define C3O_CONK_MISMATCH (-4)
/* out[d] for d in [0, lr): sum over the diagonal j-i=d of the local-alignment
 * matrix H[i][j] = max(0, H[i-1][j-1]+s, H[i-1][j]-penalty, H[i][j-1]-penalty). */
int c3o_conk(const char *splint, int ls, const char *seq, int lr, int penalty,
             int32_t *out);

/* ---- stage 2: savitzky_golay x iters + call_peaks (bin/call_peaks.py:8-16) ---- */
/* y: n doubles; coef: window doubles (host-computed with the reference's pinv
 * line, bin/savitzky_golay.py:30-31); out: n doubles.  Summation order is fixed:
 * acc=0; for k=0..window-1: acc = acc + coef[k]*ypad[t+k]  (no FMA).          */
int c3o_savgol(const double *y, int n, const double *coef, int window, double *out);
/* returns number of peaks (>=0) or <0 on error; peaks ascending.
 * smoothed_out may be NULL.  median_out may be NULL.                          */
int c3o_call_peaks(const int32_t *scores, int n, int min_dist, int iters,
                   const double *coef, int window, double *smoothed_out,
                   double *median_out, int32_t *peaks_out, int max_peaks);

/* ---- stage 3a: shift/filter peaks and split (C3POa.py:127-155) ---- */
/* peaks_io: in = call_peaks output, out = shifted+filtered peaks (count returned
 * through *n_peaks_io).  sub_bounds: [2*i]=start,[2*i+1]=end of kept subreads.
 * dang_bounds: up to 2 (start,end) pairs.  Returns 0, or 1 when the read is
 * skipped by the reference (`continue` at C3POa.py:125-126,131-132).          */
int c3o_split(int32_t *peaks_io, int *n_peaks_io, int ls, int lr,
              int32_t *sub_bounds, int *n_sub, int32_t *dang_bounds, int *n_dang);

/* ---- stage 3b: abPOA 1.0.5 msa (bin/determine_consensus.py:30-47) ---- */
typedef struct {
    int match, mismatch;          /* 5, 4 (reference: msa_aligner(match=5)) */
    int gap_open1, gap_ext1;      /* 4, 2 */
    int gap_open2, gap_ext2;      /* 24, 1 */
    int wb;                       /* extra_b = 10 */
    double wf;                    /* extra_f = 0.01 */
    int simd_bits;                /* 256 (AVX2 build): band rounded to 16 int16 / 8 int32 lanes */
    /* Named switches for two upstream branches that are recalled, not restated; both default 0:
     * int8_lanes: abPOA picks int8 SIMD lanes (granule simd_bits/8) when the score bound fits int8;
     * end_clamp:  a row's band end is clamped to (largest predecessor band end vector + 1).               */
    int int8_lanes, end_clamp;
} c3o_poa_para_t;
void c3o_poa_default_para(c3o_poa_para_t *p);

typedef struct {
    int64_t cells;        /* banded DP cells computed (rows 1..n-2, cols beg..min(end,qlen)) */
    int32_t node_n;       /* final graph nodes incl. src/sink */
    int32_t n_aln;        /* alignments performed */
    int32_t last_score;   /* best score of the last alignment */
    int32_t status;       /* 0 ok; <0 internal error */
} c3o_poa_stats_t;

/* seqs: n_seq pointers to ASCII sequences.  cons_out: buffer of cons_cap bytes
 * (ASCII, not NUL terminated), *cons_len receives the length.  msa_out may be
 * NULL; otherwise n_seq rows of msa_cap bytes, *msa_len receives the column
 * count.  dbg may be NULL; otherwise 4 int32 per alignment: score, n_ops,
 * node_n after merge, cells.                                                  */
int c3o_poa_msa(const c3o_poa_para_t *para, int n_seq, const char *const *seqs,
                const int32_t *seq_lens, char *cons_out, int cons_cap, int *cons_len,
                char *msa_out, int msa_cap, int *msa_len, c3o_poa_stats_t *stats,
                int32_t *dbg);

/* ---- whole per-read path (analyze_reads body, C3POa.py:112-165, pre-racon) ---- */
typedef struct {
    int32_t status;       /* 0 = consensus produced; 1 = skipped (no peaks); 2 = repeats<3 path not taken here */
    int32_t n_peaks;
    int32_t n_sub;
    int32_t n_dang;
    int32_t cons_len;
    int32_t poa_nodes;   /* final graph size (0 when no POA ran) */
    int64_t poa_cells;
    int64_t conk_cells;
} c3o_read_result_t;

/* Batch driver used as the CPU baseline: OpenMP over reads.  reads are ASCII,
 * CSR offsets (n_reads+1).  splint already strand-resolved per read
 * (splint_idx[n_reads] indexes splints[]).  Outputs: peaks (CSR, capacity
 * max_peaks per read), sub bounds, cons (capacity cons_cap per read).       */
int c3o_consensus_batch(int n_reads, const char *reads, const int64_t *read_off,
                        int n_splints, const char *const *splints, const int32_t *splint_lens,
                        const int32_t *splint_idx, int penalty, int min_dist, int iters,
                        const double *coef, int window, const c3o_poa_para_t *para,
                        int max_peaks, int32_t *peaks_out, int32_t *sub_bounds_out,
                        int32_t *dang_bounds_out, int cons_cap, char *cons_out,
                        c3o_read_result_t *results, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
