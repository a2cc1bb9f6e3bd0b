// conk.cuh -- stage 1: splint-vs-read local-alignment diagonal score profile.
//
// Replaces conk.conk(splint, seq, penalty) (call site /root/reference/C3POa.py:123).
// H[i][j] = max(0, H[i-1][j-1]+s, H[i-1][j]-pen, H[i][j-1]-pen), s = +5/-4;
// profile[d] = sum of H over the diagonal j-i = d, d in [0, Lr).
//
// Mapping: one warp per read.  The splint's rows are split over the 32 lanes,
// R consecutive rows per lane, held in registers; the warp sweeps the read as
// a skewed wavefront (lane t works on column step-t), passing the bottom cell
// of each lane's row block and the running diagonal sum to lane t+1 with
// warp shuffles.  Each diagonal's sum therefore travels down the lanes with
// the wavefront and leaves the last lane complete -- no atomics, no matrix in
// memory.  Rows are padded at the TOP (sentinel bases never match, so padded
// rows stay 0); columns past the read end are run as "virtual" columns with
// the scores forced to 0 so that unfinished diagonals drain to the last lane.
// Splints longer than 32*R rows take several passes with the boundary row
// staged in a per-warp global scratch row.
#pragma once
#include "common.cuh"

#define C3_CONK_MATCH 5
#define C3_CONK_MISMATCH (-4)
#define C3_CONK_THREADS 128
#define C3_CONK_MAXR 16

template <int R>
__global__ void __launch_bounds__(C3_CONK_THREADS)
c3_conk_kernel(const uint8_t *__restrict__ codes, const int64_t *__restrict__ read_off, int n_reads,
               const uint8_t *__restrict__ sp_codes, const int32_t *__restrict__ sp_off,
               const int32_t *__restrict__ sp_idx, int penalty, int32_t *__restrict__ prof,
               int32_t *brow, int64_t brow_stride, unsigned *counter)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int32_t *my_brow = brow ? brow + (int64_t)gwarp * 2 * brow_stride : nullptr;
    const int BIG = 1 << 20;
    constexpr int RPP = 32 * R;

    for (;;) {
        int r = 0;
        if (lane == 0) r = (int)atomicAdd(counter, 1u);
        r = __shfl_sync(C3_FULL, r, 0);
        if (r >= n_reads) break;
        const int64_t off = read_off[r];
        const int Lr = (int)(read_off[r + 1] - off);
        const uint8_t *seq = codes + off;
        int32_t *out = prof + off;
        const int si = sp_idx[r];
        const uint8_t *sp = sp_codes + sp_off[si];
        const int Ls = sp_off[si + 1] - sp_off[si];
        const int npass = (Ls + RPP - 1) / RPP;
        const int pad = npass * RPP - Ls;

        for (int pass = 0; pass < npass; ++pass) {
            int a[R], h[R], p[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int i = pass * RPP + lane * R + q - pad;
                int c = i >= 0 ? (int)sp[i] : 5;
                a[q] = c >= 4 ? 5 : c;          // N / padding never match (read N is 4)
                h[q] = 0; p[q] = 0;
            }
            const int i_last = pass * RPP + RPP - 1 - pad;
            const int32_t *bin = my_brow ? my_brow + ((pass + 1) & 1) * brow_stride : nullptr;
            int32_t *bout = my_brow ? my_brow + (pass & 1) * brow_stride : nullptr;
            const bool has_top = pass > 0;
            const bool has_bot = pass < npass - 1;
            int top_prev = 0, inc_hold = 0, bot_send = 0, p_send = 0, keep = 0;
            const int nsteps = ((Lr + RPP + 31 + 31) >> 5) << 5;
            int b_next = ((unsigned)(-lane) < (unsigned)Lr) ? (int)seq[-lane] : 4;

            for (int step = 0; step < nsteps; ++step) {
                const int j = step - lane;
                const bool valid = (unsigned)j < (unsigned)Lr;
                const int b = b_next;
                b_next = ((unsigned)(j + 1) < (unsigned)Lr) ? (int)seq[j + 1] : 4;
                const int mat = valid ? C3_CONK_MATCH : -BIG;
                const int mis = valid ? C3_CONK_MISMATCH : -BIG;
                const int npen = valid ? -penalty : -BIG;
                int top_cur = __shfl_up_sync(C3_FULL, bot_send, 1);
                int recv = __shfl_up_sync(C3_FULL, p_send, 1);
                if (lane == 0) {
                    recv = 0;
                    top_cur = (has_top && valid) ? bin[j] : 0;
                }
                const int inc = inc_hold;
                inc_hold = recv;
                int diag = top_prev, up = top_cur;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int left = h[q];
                    const int s = (b == a[q]) ? mat : mis;
                    const int t1 = __viaddmax_s32_relu(left, npen, diag + s);
                    const int hn = __viaddmax_s32(up, npen, t1);
                    diag = left; h[q] = hn; up = hn;
                }
                top_prev = top_cur;
#pragma unroll
                for (int q = R - 1; q >= 1; --q) p[q] = p[q - 1] + h[q];
                p[0] = inc + h[0];
                p_send = p[R - 1];
                bot_send = h[R - 1];
                // the last lane's outgoing sum is a finished diagonal; gather 32 of them, store coalesced
                const int o = __shfl_sync(C3_FULL, p_send, 31);
                if (lane == (step & 31)) keep = o;
                if ((step & 31) == 31) {
                    const int d = (step - 31 + lane) - 31 - i_last;
                    if (d >= 0 && d < Lr) {
                        if (pass == 0) out[d] = keep; else out[d] += keep;
                    }
                }
                if (has_bot && lane == 31 && valid) bout[j] = bot_send;
            }
            __syncwarp();
        }
    }
}


// ---------------------------------------------------------------------------
// Packed variant: TWO reads per warp, one in each 16-bit half of every register (VIADDMNMX.S16x2[.RELU], VIADD.16x2,
// VIMNMX.U16x2): 8 instructions per pair of cells, one of them on the FMA pipe, where the kernel above needs 6 per cell
// on the integer pipe it is bound by.  Same cells, same sums, same output.
//   * both reads of a pair have the same splint (the host pairs reads by splint index and 2 kb length class), so the
//     splint rows are shared; the score of a pair of cells is a select by the per-half mismatch mask of (read base XOR
//     splint base);
//   * H <= 5 * Ls fits a half as long as the splint fits one pass (Ls <= 32 * R <= 480: the caller checks);
//   * the diagonal sums do not fit 16 bits.  Inside a lane the sum of its R cells of a diagonal is carried packed
//     (<= 15 * 2400); the 32-bit sum arriving from the lane above bypasses the rows through a per-lane delay line in
//     shared memory (written at step t, read at step t + R) and is added when the diagonal leaves the lane.
// pairs: [2 * n_pairs] read indices, the second -1 for a read without partner; *n_pairs_dev is read on the device.
// ---------------------------------------------------------------------------
#define C3_CONK2_NEG 0xC180u           // -16000 as a half: "no cell here" for the adds, far from wrapping

struct c3_false_t { static constexpr bool value = false; };
struct c3_true_t { static constexpr bool value = true; };

template <int R>
__global__ void __launch_bounds__(C3_CONK_THREADS)
c3_conk2_kernel(const uint8_t *__restrict__ codes, const int64_t *__restrict__ read_off, const int32_t *__restrict__ pairs,
                const int *__restrict__ n_pairs_dev, const uint8_t *__restrict__ sp_codes, const int32_t *__restrict__ sp_off,
                const int32_t *__restrict__ sp_idx, int penalty, int32_t *__restrict__ prof, unsigned *counter)
{
    __shared__ int s_ring[C3_CONK_THREADS / 32][2][16][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int RPP = 32 * R;
    const int n_pairs = *n_pairs_dev;
    const unsigned pen16 = (unsigned)(-penalty) & 0xffffu;

    for (;;) {
        int pr = 0;
        if (lane == 0) pr = (int)atomicAdd(counter, 1u);
        pr = __shfl_sync(C3_FULL, pr, 0);
        if (pr >= n_pairs) break;
        const int rA = pairs[2 * pr], rB = pairs[2 * pr + 1];
        if (rA < 0) continue;
        const int64_t offA = read_off[rA];
        const int LrA = (int)(read_off[rA + 1] - offA);
        const int64_t offB = rB >= 0 ? read_off[rB] : offA;
        const int LrB = rB >= 0 ? (int)(read_off[rB + 1] - offB) : 0;
        const uint8_t *seqA = codes + offA, *seqB = codes + offB;
        int32_t *outA = prof + offA, *outB = prof + offB;
        const int si = sp_idx[rA];
        const uint8_t *sp = sp_codes + sp_off[si];
        const int Ls = sp_off[si + 1] - sp_off[si];
        const int pad = RPP - Ls;                          // one pass: Ls <= RPP
        unsigned a2[R], h2[R], p2[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int i = lane * R + q - pad;
            int c = i >= 0 ? (int)sp[i] : 5;
            c = c >= 4 ? 5 : c;                            // N / padding never match (read N is 4)
            a2[q] = (unsigned)c * 0x00010001u; h2[q] = 0u; p2[q] = 0u;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) { s_ring[wib][0][t][lane] = 0; s_ring[wib][1][t][lane] = 0; }
        const int i_last = RPP - 1 - pad;
        const int Lmax = max(LrA, LrB);
        const int nsteps = ((Lmax + RPP + 31 + 31) >> 5) << 5;
        unsigned top_prev = 0u, bot_send = 0u;
        int p_sendA = 0, p_sendB = 0, keepA = 0, keepB = 0;
        int bA_next = ((unsigned)(-lane) < (unsigned)LrA) ? (int)seqA[-lane] : 4;
        int bB_next = ((unsigned)(-lane) < (unsigned)LrB) ? (int)seqB[-lane] : 4;
        const unsigned mat_ok = (unsigned)C3_CONK_MATCH * 0x00010001u, mis_ok = ((unsigned)C3_CONK_MISMATCH & 0xffffu) * 0x00010001u,
                       pen_ok = pen16 * 0x00010001u;

        // One step of the wavefront.  TAIL = false: every lane's column lies before the end of both reads (all steps up to
        // the shorter read's length), so the per-half constants are fixed; columns before the start need nothing special
        // (their base code 4 never matches and all their inputs are 0, so their cells stay 0).  TAIL = true: columns past
        // a read's end are forced to 0 through the "no cell here" constants, per half.
        auto body = [&](const int step, auto tail_c) {
            constexpr bool TAIL = decltype(tail_c)::value;
            const int j = step - lane;
            const unsigned b2 = (unsigned)bA_next | ((unsigned)bB_next << 16);
            unsigned matv = mat_ok, misv = mis_ok, npen = pen_ok;
            if (TAIL) {
                const bool vA = (unsigned)j < (unsigned)LrA, vB = (unsigned)j < (unsigned)LrB;
                bA_next = ((unsigned)(j + 1) < (unsigned)LrA) ? (int)seqA[j + 1] : 4;
                bB_next = ((unsigned)(j + 1) < (unsigned)LrB) ? (int)seqB[j + 1] : 4;
                matv = (vA ? (unsigned)C3_CONK_MATCH : C3_CONK2_NEG) | ((vB ? (unsigned)C3_CONK_MATCH : C3_CONK2_NEG) << 16);
                misv = (vA ? ((unsigned)C3_CONK_MISMATCH & 0xffffu) : C3_CONK2_NEG) | ((vB ? ((unsigned)C3_CONK_MISMATCH & 0xffffu) : C3_CONK2_NEG) << 16);
                npen = (vA ? pen16 : C3_CONK2_NEG) | ((vB ? pen16 : C3_CONK2_NEG) << 16);
            } else {
                // (j + 1 <= step + 1 <= the shorter length: at worst one byte past a read, inside the padded buffer, and
                // that value is only used by a TAIL step, which masks it)
                bA_next = j + 1 >= 0 ? (int)seqA[j + 1] : 4;
                bB_next = j + 1 >= 0 ? (int)seqB[j + 1] : 4;
            }
            unsigned top_cur = __shfl_up_sync(C3_FULL, bot_send, 1);
            int recvA = __shfl_up_sync(C3_FULL, p_sendA, 1);
            int recvB = __shfl_up_sync(C3_FULL, p_sendB, 1);
            if (lane == 0) { recvA = 0; recvB = 0; top_cur = 0u; }
            // the sums arriving now leave this lane R steps from now
            const int incA = s_ring[wib][0][(step - R) & 15][lane], incB = s_ring[wib][1][(step - R) & 15][lane];
            s_ring[wib][0][step & 15][lane] = recvA; s_ring[wib][1][step & 15][lane] = recvB;
            unsigned diag = top_prev, up = top_cur;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const unsigned left = h2[q];
                const unsigned mask = __vminu2(b2 ^ a2[q], 0x00010001u) * 0xffffu;     // per half: all ones = mismatch
                const unsigned sc = (mask & misv) | (~mask & matv);
                const unsigned t1 = __viaddmax_s16x2_relu(left, npen, __vadd2(diag, sc));
                const unsigned hn = __viaddmax_s16x2(up, npen, t1);
                diag = left; h2[q] = hn; up = hn;
            }
            top_prev = top_cur;
#pragma unroll
            for (int q = R - 1; q >= 1; --q) p2[q] = __vadd2(p2[q - 1], h2[q]);
            p2[0] = h2[0];
            p_sendA = (int)(p2[R - 1] & 0xffffu) + incA;
            p_sendB = (int)(p2[R - 1] >> 16) + incB;
            bot_send = h2[R - 1];
            // the last lane's outgoing sums are finished diagonals; gather 32 of them, store coalesced
            const int oA = __shfl_sync(C3_FULL, p_sendA, 31), oB = __shfl_sync(C3_FULL, p_sendB, 31);
            if (lane == (step & 31)) { keepA = oA; keepB = oB; }
            if ((step & 31) == 31) {
                const int d = (step - 31 + lane) - 31 - i_last;
                if (d >= 0 && d < LrA) outA[d] = keepA;
                if (d >= 0 && d < LrB) outB[d] = keepB;
            }
        };
        const int n_fast = rB >= 0 ? min(LrA, LrB) : 0;
        int step = 0;
        for (; step < n_fast; ++step) body(step, c3_false_t());
        for (; step < nsteps; ++step) body(step, c3_true_t());
        __syncwarp();
    }
}

// pairing of the reads [r0, r1) by (splint index, 2 kb length class): histogram, scan with every class padded to an even
// number of slots, scatter.  Empty slots hold -1 (the list is preset to 0xff bytes).
#define C3_CONK2_MAXKEYS 4096
__global__ void c3_conk2_count_kernel(int r0, int r1, const int64_t *read_off, const int32_t *sp_idx, unsigned *hist)
{
    const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= r1) return;
    const int L = (int)(read_off[r + 1] - read_off[r]);
    atomicAdd(&hist[sp_idx[r] * 32 + min(31, L >> 11)], 1u);
}
__global__ void __launch_bounds__(1024) c3_conk2_scan_kernel(int nkeys, const unsigned *hist, unsigned *start, int *n_pairs)
{
    __shared__ unsigned s_sum[1024];
    const int t = threadIdx.x;
    const int per = (nkeys + 1023) / 1024;
    unsigned loc = 0;
    for (int k = t * per; k < min(nkeys, (t + 1) * per); ++k) loc += (hist[k] + 1u) & ~1u;
    s_sum[t] = loc;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const unsigned v = t >= d ? s_sum[t - d] : 0u;
        __syncthreads();
        s_sum[t] += v;
        __syncthreads();
    }
    unsigned acc = s_sum[t] - loc;
    for (int k = t * per; k < min(nkeys, (t + 1) * per); ++k) { start[k] = acc; acc += (hist[k] + 1u) & ~1u; }
    if (t == 1023) *n_pairs = (int)(s_sum[1023] / 2u);
}
__global__ void c3_conk2_scatter_kernel(int r0, int r1, const int64_t *read_off, const int32_t *sp_idx, const unsigned *start,
                                        unsigned *fill, int32_t *pairs)
{
    const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= r1) return;
    const int L = (int)(read_off[r + 1] - read_off[r]);
    const int key = sp_idx[r] * 32 + min(31, L >> 11);
    pairs[start[key] + atomicAdd(&fill[key], 1u)] = r;
}
