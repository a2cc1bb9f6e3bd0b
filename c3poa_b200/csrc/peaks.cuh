// peaks.cuh -- stage 2 (Savitzky-Golay x iters + call_peaks) and stage 3a (split).
//
// Replaces call_peaks(scores, min_dist, iters, window, order)
//   (/root/reference/bin/call_peaks.py:8-16, bin/savitzky_golay.py:33-36, and the
//   scipy.signal.find_peaks semantics pinned in SURVEY.md appendix B.2), and the
//   peak shift / subread split of analyze_reads (/root/reference/C3POa.py:127-155).
//
// One CTA per read (persistent, atomic work counter).  fp64 arithmetic is kept as
// separate IEEE multiplies and adds in a fixed order (k = 0..window-1), so the
// smoothed profile is bit-identical to the CPU oracle and within ~1e-13 of
// numpy's BLAS dot product.
#pragma once
#include "common.cuh"

#ifndef C3_PK_THREADS
#define C3_PK_THREADS 256
#endif
#define C3_PK_TILE (4 * C3_PK_THREADS)   // four consecutive outputs per thread
#define C3_PK_MAXWIN 127
#define C3_PK_MAXC 1024      // candidate local maxima above the height threshold, per read

struct c3_peaks_args {
    const int32_t *prof;      // int32 CSR
    const int64_t *off;       // [n+1]
    int n;
    double coefv[C3_PK_MAXWIN];   // [window] by value: the taps are read from the kernel-parameter constant bank
    int window, iters, min_dist;
    double height_mult, gate_mult;
    double *scratch;          // per-CTA 2 * scratch_stride doubles
    int64_t scratch_stride;
    double *out_smoothed;     // optional CSR
    double *out_median;       // optional [n]
    int32_t *out_peaks;       // [n][max_peaks]
    int32_t *out_n_peaks;     // [n]
    int max_peaks;
    unsigned *counter;
};

__device__ __forceinline__ unsigned long long c3_dkey(double x)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double c3_dunkey(unsigned long long k)
{
    unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// k-th smallest (0-based) of buf[0..n): MSB-first radix select, 8 bits per pass; as soon as at most C3_PK_SMALL
// candidates share the prefix (after 2-3 passes for a smoothed profile) they are gathered and the answer is the one
// whose rank among them is k.  cand: C3_PK_SMALL keys of shared memory.
#define C3_PK_SMALL 512
__device__ double c3_radix_select(const double *buf, int n, int k, unsigned *hist, unsigned long long *sh,
                                  unsigned long long *cand)
{
    unsigned long long prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        // smoothed profiles share their leading bytes: without aggregation the first passes are n atomics on one
        // address.  Lanes with the same digit elect one of them to add their count.
        for (int i0 = 0; i0 < n; i0 += blockDim.x) {
            const int i = i0 + threadIdx.x;
            unsigned digit = 0xffffffffu;
            if (i < n) {
                const unsigned long long kx = c3_dkey(buf[i]);
                if ((kx & mask) == prefix) digit = (unsigned)(kx >> shift) & 255u;
            }
            const unsigned peers = __match_any_sync(C3_FULL, digit);
            if (digit != 0xffffffffu && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[digit], (unsigned)__popc(peers));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned cum = 0; int b = 0;
            for (; b < 256; ++b) { if (cum + hist[b] > (unsigned)k) break; cum += hist[b]; }
            sh[0] = prefix | ((unsigned long long)b << shift);
            sh[1] = (unsigned long long)(k - (int)cum);
            sh[2] = (unsigned long long)hist[b];           // candidates left
            sh[3] = 0;
        }
        __syncthreads();
        prefix = sh[0]; k = (int)sh[1];
        const int left = (int)sh[2];
        mask |= 0xffull << shift;
        __syncthreads();
        if (left <= C3_PK_SMALL && shift > 0) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const unsigned long long kx = c3_dkey(buf[i]);
                if ((kx & mask) == prefix) cand[atomicAdd((unsigned *)&sh[3], 1u)] = kx;
            }
            __syncthreads();
            for (int t = threadIdx.x; t < left; t += blockDim.x) {
                const unsigned long long kt = cand[t];
                int rank = 0;
                for (int u = 0; u < left; ++u) { const unsigned long long ku = cand[u]; rank += (ku < kt) || (ku == kt && u < t); }
                if (rank == k) sh[0] = kt;
            }
            __syncthreads();
            prefix = sh[0];
            __syncthreads();
            break;
        }
    }
    return c3_dunkey(prefix);
}


// 4 consecutive outputs of the FIR from a staged tile: out[j] = sum_k coef[k] * tile[u + j + k], k ascending, separate
// IEEE multiply and add (the oracle's order).  The window is read once as 16-byte words; the taps come from the constant
// bank (kernel parameters).  WT > 0: compile-time window (41, the reference's), fully unrolled; WT = 0: any odd window.
template <int WT>
__device__ __forceinline__ void c3_pk_fir4(const c3_peaks_args &A, const double *tile_u, const int W, double (&a)[4])
{
    a[0] = a[1] = a[2] = a[3] = 0.0;
    const double2 *tp = reinterpret_cast<const double2 *>(tile_u);
    double2 p0 = tp[0], p1 = tp[1];
    if (WT > 0) {
#pragma unroll
        for (int k = 0; k < WT; k += 2) {
            const double2 p2 = tp[(k >> 1) + 2];
            const double c0 = A.coefv[k];
            a[0] = __dadd_rn(a[0], __dmul_rn(c0, p0.x)); a[1] = __dadd_rn(a[1], __dmul_rn(c0, p0.y));
            a[2] = __dadd_rn(a[2], __dmul_rn(c0, p1.x)); a[3] = __dadd_rn(a[3], __dmul_rn(c0, p1.y));
            if (k + 1 < WT) {
                const double c1 = A.coefv[k + 1];
                a[0] = __dadd_rn(a[0], __dmul_rn(c1, p0.y)); a[1] = __dadd_rn(a[1], __dmul_rn(c1, p1.x));
                a[2] = __dadd_rn(a[2], __dmul_rn(c1, p1.y)); a[3] = __dadd_rn(a[3], __dmul_rn(c1, p2.x));
            }
            p0 = p1; p1 = p2;
        }
    } else {
        for (int k = 0; k < W; k += 2) {
            const double2 p2 = tp[(k >> 1) + 2];
            const double c0 = A.coefv[k];
            a[0] = __dadd_rn(a[0], __dmul_rn(c0, p0.x)); a[1] = __dadd_rn(a[1], __dmul_rn(c0, p0.y));
            a[2] = __dadd_rn(a[2], __dmul_rn(c0, p1.x)); a[3] = __dadd_rn(a[3], __dmul_rn(c0, p1.y));
            if (k + 1 < W) {
                const double c1 = A.coefv[k + 1];
                a[0] = __dadd_rn(a[0], __dmul_rn(c1, p0.y)); a[1] = __dadd_rn(a[1], __dmul_rn(c1, p1.x));
                a[2] = __dadd_rn(a[2], __dmul_rn(c1, p1.y)); a[3] = __dadd_rn(a[3], __dmul_rn(c1, p2.x));
            }
            p0 = p1; p1 = p2;
        }
    }
}

// padded signal (bin/savitzky_golay.py:33-35) at padded index t: left mirror about y0, the signal, right mirror about yl
template <typename F>
__device__ __forceinline__ double c3_pk_padded(const int t, const int n, const int half, const double y0, const double yl, F at)
{
    if (t < half) return y0 - fabs(at(half - t) - y0);
    if (t < half + n) return at(t - half);
    return yl + fabs(at(n - 2 - (t - half - n)) - yl);
}

#ifndef C3_PK_MINB
#define C3_PK_MINB 6
#endif
__global__ void __launch_bounds__(C3_PK_THREADS, C3_PK_MINB) c3_peaks_kernel(c3_peaks_args A)
{
    __shared__ __align__(16) double s_tile[C3_PK_TILE + C3_PK_MAXWIN + 9];
    __shared__ double s_pr[C3_PK_MAXC];
    __shared__ int s_pos[C3_PK_MAXC];
    __shared__ int s_order[C3_PK_MAXC];
    __shared__ unsigned char s_keep[C3_PK_MAXC];
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_sh[4];
    __shared__ double s_red[C3_PK_THREADS / 32];
    __shared__ int s_cnt[C3_PK_THREADS + 1];
    __shared__ int s_r;

    const int tid = threadIdx.x;
    const int W = A.window, half = (A.window - 1) / 2;
    double *bufA = A.scratch + (int64_t)blockIdx.x * 2 * A.scratch_stride;
    double *bufB = bufA + A.scratch_stride;

    for (;;) {
        if (tid == 0) s_r = (int)atomicAdd(A.counter, 1u);
        __syncthreads();
        const int r = s_r;
        __syncthreads();
        if (r >= A.n) break;
        const int64_t off = A.off[r];
        const int n = (int)(A.off[r + 1] - off);
        const int32_t *prof = A.prof + off;
        if (n < half + 2) {                      // reference would fail / reads are >= lencutoff
            if (tid == 0) { A.out_n_peaks[r] = -1; if (A.out_median) A.out_median[r] = 0.0; }
            continue;
        }
        // ---------------- Savitzky-Golay passes ----------------
        double *cur = nullptr;
        for (int it = 0; it < A.iters; ++it) {
            const bool from_int = (cur == nullptr);
            const double *src = cur;
            double *dst = (cur == bufA) ? bufB : bufA;
            // edge anchors (bin/savitzky_golay.py:33-34)
            const double y0 = from_int ? (double)prof[0] : src[0];
            const double yl = from_int ? (double)prof[n - 1] : src[n - 1];
            auto at = [&](const int idx) { return from_int ? (double)prof[idx] : src[idx]; };
            for (int t0 = 0; t0 < n; t0 += C3_PK_TILE) {
                const int tn = min(C3_PK_TILE, n - t0);
                for (int u = tid; u < tn + W - 1 + 8; u += blockDim.x) {
                    const int t = t0 + u;
                    s_tile[u] = t < n + W - 1 ? c3_pk_padded(t, n, half, y0, yl, at) : 0.0;
                }
                __syncthreads();
                const int u4 = 4 * tid;
                if (u4 < tn) {
                    double a[4];
                    if (W == 41) c3_pk_fir4<41>(A, s_tile + u4, W, a); else c3_pk_fir4<0>(A, s_tile + u4, W, a);
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (u4 + k < tn) dst[t0 + u4 + k] = a[k];
                }
                __syncthreads();
            }
            cur = dst;
        }
        if (cur == nullptr) {                     // iters == 0: operate on the raw scores
            for (int i = tid; i < n; i += blockDim.x) bufA[i] = (double)prof[i];
            cur = bufA;
        }
        __syncthreads();
        const double *x = cur;
        if (A.out_smoothed)
            for (int i = tid; i < n; i += blockDim.x) A.out_smoothed[off + i] = x[i];
        // ---------------- median (np.median) and max ----------------
        double med;
        if (n & 1) med = c3_radix_select(x, n, n / 2, s_hist, s_sh, reinterpret_cast<unsigned long long *>(s_pr));
        else {
            // hi = element n/2 of the sorted profile; lo = element n/2 - 1 = hi itself when at most n/2 - 1 samples are
            // smaller than hi, else the largest sample below hi: one pass instead of a second select
            const double hi = c3_radix_select(x, n, n / 2, s_hist, s_sh, reinterpret_cast<unsigned long long *>(s_pr));
            double below = -1.0e308; int nb = 0;
            for (int i = tid; i < n; i += blockDim.x) { const double v = x[i]; if (v < hi) { ++nb; below = fmax(below, v); } }
            for (int o = 16; o > 0; o >>= 1) { below = fmax(below, __shfl_xor_sync(C3_FULL, below, o)); nb += __shfl_xor_sync(C3_FULL, nb, o); }
            if ((tid & 31) == 0) { s_red[tid >> 5] = below; s_cnt[tid >> 5] = nb; }
            __syncthreads();
            below = s_red[0]; nb = s_cnt[0];
            for (int i = 1; i < C3_PK_THREADS / 32; ++i) { below = fmax(below, s_red[i]); nb += s_cnt[i]; }
            __syncthreads();
            const double lo = (nb > n / 2 - 1) ? below : hi;
            med = __ddiv_rn(__dadd_rn(lo, hi), 2.0);
        }
        double mx = -1.0e308;
        for (int i = tid; i < n; i += blockDim.x) mx = fmax(mx, x[i]);
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(C3_FULL, mx, o));
        if ((tid & 31) == 0) s_red[tid >> 5] = mx;
        __syncthreads();
        mx = s_red[0];
        for (int i = 1; i < C3_PK_THREADS / 32; ++i) mx = fmax(mx, s_red[i]);
        __syncthreads();
        if (A.out_median && tid == 0) A.out_median[r] = med;
        if (mx < __dmul_rn(A.gate_mult, med)) {      // bin/call_peaks.py:13-14
            if (tid == 0) A.out_n_peaks[r] = 0;
            continue;
        }
        const double hmin = __dmul_rn(med, A.height_mult);          // bin/call_peaks.py:15
        // ---------------- local maxima (scipy _local_maxima_1d) + height ----------------
        const int span = n - 2;                                     // candidates i in [1, n-2]
        const int chunk = (span + blockDim.x - 1) / blockDim.x;
        const int c0 = 1 + tid * chunk, c1 = min(n - 1, c0 + chunk);
        int mycount = 0;
        for (int pass = 0; pass < 2; ++pass) {
            int w = (pass == 1) ? s_cnt[tid] : 0;
            for (int i = c0; i < c1; ++i) {
                const double xi = x[i];
                if (x[i - 1] < xi) {
                    int ia = i + 1;
                    while (ia < n - 1 && x[ia] == xi) ++ia;
                    if (x[ia] < xi) {
                        const int mid = (i + ia - 1) >> 1;
                        const double xm = x[mid];
                        if (hmin <= xm) {
                            if (pass == 0) ++mycount;
                            else { if (w < C3_PK_MAXC) { s_pos[w] = mid; s_pr[w] = xm; } ++w; }
                        }
                    }
                }
            }
            if (pass == 0) {
                s_cnt[tid + 1] = mycount;
                if (tid == 0) s_cnt[0] = 0;
                __syncthreads();
                if (tid == 0) for (int i = 1; i <= (int)blockDim.x; ++i) s_cnt[i] += s_cnt[i - 1];
                __syncthreads();
            }
        }
        const int nc = s_cnt[blockDim.x];
        __syncthreads();
        if (nc > C3_PK_MAXC) {
            if (tid == 0) A.out_n_peaks[r] = -2;
            continue;
        }
        // ---------------- distance rule (scipy _select_by_peak_distance) ----------------
        for (int t = tid; t < nc; t += blockDim.x) {
            const double pt = s_pr[t];
            int rank = 0;
            for (int u = 0; u < nc; ++u) {
                const double pu = s_pr[u];
                rank += (pu < pt) || (pu == pt && u < t);
            }
            s_order[rank] = t;
            s_keep[t] = 1;
        }
        __syncthreads();
        if (tid == 0) {
            for (int t = nc - 1; t >= 0; --t) {
                const int j = s_order[t];
                if (!s_keep[j]) continue;
                int k = j - 1;
                while (k >= 0 && s_pos[j] - s_pos[k] < A.min_dist) { s_keep[k] = 0; --k; }
                k = j + 1;
                while (k < nc && s_pos[k] - s_pos[j] < A.min_dist) { s_keep[k] = 0; ++k; }
            }
            int np = 0;
            int32_t *op = A.out_peaks + (int64_t)r * A.max_peaks;
            for (int t = 0; t < nc; ++t)
                if (s_keep[t]) { if (np < A.max_peaks) op[np] = s_pos[t]; ++np; }
            A.out_n_peaks[r] = np <= A.max_peaks ? np : -3;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// stage 3a: peak shift / filter / subread split  (C3POa.py:127-155), one thread per read
// ---------------------------------------------------------------------------
struct c3_read_result_dev {     // mirrors c3_read_result in include/c3poa_gpu.h
    int32_t status, n_peaks, n_sub, n_dang, cons_len, poa_nodes;
    long long poa_cells;
};

#define C3_SPLIT_MAXP 256

__global__ void c3_split_kernel(int n_reads, const int64_t *__restrict__ read_off,
                                const int32_t *__restrict__ sp_off, const int32_t *__restrict__ sp_idx,
                                int32_t *peaks, const int32_t *__restrict__ n_peaks_raw, int max_peaks,
                                int32_t *sub_bounds, int32_t *dang_bounds, c3_read_result_dev *res,
                                int32_t *stats /* [0]=max qlen, [1]=max n_sub, [2]=max total sub len, [3]=#poa items */)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    c3_read_result_dev out;
    out.status = 0; out.n_peaks = 0; out.n_sub = 0; out.n_dang = 0; out.cons_len = 0; out.poa_nodes = 0; out.poa_cells = 0;
    const int lr = (int)(read_off[r + 1] - read_off[r]);
    const int si = sp_idx[r];
    const int ls = sp_off[si + 1] - sp_off[si];
    int32_t *pk = peaks + (int64_t)r * max_peaks;
    int32_t *sb = sub_bounds + (int64_t)r * max_peaks * 2;
    int32_t *db = dang_bounds + (int64_t)r * 4;
    db[0] = db[1] = db[2] = db[3] = 0;
    int np = n_peaks_raw[r];
    if (np < 0 || np > max_peaks || np > C3_SPLIT_MAXP) { out.status = np < 0 ? -100 + np : -110; res[r] = out; return; }
    int k = 0;
    for (int i = 0; i < np; ++i) {                         // C3POa.py:127-130
        const int p = pk[i] + ls / 2;
        if (p < lr) pk[k++] = p;
    }
    np = k; out.n_peaks = np;
    if (np == 0) { out.status = 1; res[r] = out; return; }  // C3POa.py:125-126,131-132
    if (np > 1) {
        int rl[C3_SPLIT_MAXP], srt[C3_SPLIT_MAXP];
        const int nl = np - 1;
        for (int i = 0; i < nl; ++i) {
            const double xx = (double)(pk[i + 1] - pk[i]);
            rl[i] = (int)(50.0 * rint(__ddiv_rn(xx, 50.0)));   // rounding(): Python round = half-to-even
            int q = i;                                     // insertion sort
            while (q > 0 && srt[q - 1] > rl[i]) { srt[q] = srt[q - 1]; --q; }
            srt[q] = rl[i];
        }
        const double med = (nl & 1) ? (double)srt[nl / 2]
                                    : __ddiv_rn(__dadd_rn((double)srt[nl / 2 - 1], (double)srt[nl / 2]), 2.0);
        const double lo = __dmul_rn(med, 0.8), hi = __dmul_rn(med, 1.2);
        int ns = 0, tot = 0, mq = 0;
        for (int i = 0; i < nl; ++i) {
            const double v = (double)rl[i];
            if (lo <= v && v <= hi) {
                sb[2 * ns] = pk[i]; sb[2 * ns + 1] = pk[i + 1];
                const int L = pk[i + 1] - pk[i];
                tot += L; mq = max(mq, L);
                ++ns;
            }
        }
        out.n_sub = ns;
        int nd = 0;
        if (pk[0] > 100) { db[0] = 0; db[1] = pk[0]; nd = 1; }
        if (lr - pk[np - 1] > 100) { db[2 * nd] = pk[np - 1]; db[2 * nd + 1] = lr; ++nd; }
        out.n_dang = nd;
        if (ns >= 2) {                                     // POA work: consensus (>=3) or the two MSA rows (==2)
            out.poa_cells = tot;                           // subread bases, for the host's work ordering; the POA
                                                           // kernel overwrites it with the DP cell count
            atomicMax(&stats[0], mq); atomicMax(&stats[1], ns); atomicMax(&stats[2], tot); atomicAdd(&stats[3], 1);
        }
        if (ns == 2 || ns == 0) out.status = 2;            // pairwise / zero-repeat paths
    } else {
        db[0] = 0; db[1] = pk[0]; db[2] = pk[0]; db[3] = lr;
        out.n_dang = 2; out.status = 2;
    }
    res[r] = out;
}
