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
