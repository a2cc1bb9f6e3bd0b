// poa_lane.cuh -- stage 3b, fast kernel: one THREAD per read, 32 reads per warp in lockstep.
//
// Same algorithm and the same results as poa.cuh (abPOA 1.0.5 semantics; replaces
// poa.msa_aligner(match=5).msa(...), /root/reference/bin/determine_consensus.py:30-47), other mapping:
//
//   poa.cuh       one warp per read, lanes over band columns.  A typical row is ~65 columns wide, so half the
//                 lanes idle and every row pays ~300 warp-uniform bookkeeping instructions 32 times over.
//   poa_lane.cuh  one thread per read.  Each thread walks its own graph; the 32 threads of a warp run their
//                 s-th DP row at the same time.  Row bookkeeping is now useful work in every lane, the
//                 horizontal (F) gap is a plain in-thread recurrence (no warp scan), the row arg-max is one
//                 packed max per cell (no REDUX), and backtrack / merge / consensus run 32 at a time.
//
// DP rows live in a per-warp arena, interleaved by lane so that the warp's loads and stores are coalesced:
// the s-th row step of a warp owns `mv` vectors (mv = widest band among the 32 threads, in 16-column
// vectors = abPOA's int16 SIMD granule); vector vi holds H, E1, E2 as 3 x 4 int4 per lane at
//     int4 index  base4 + vi*384 + (array*4 + quarter)*32 + lane.
// Columns past a row's band end inside its last vector hold NEG_INF, so successors load whole vectors.
// Graph, row records, cigar and the int8 query profile are thread-private (c3_poa_ws without cells).
//
// Scope: int16-lane score mode with the 256-bit granule (pn = 16), banded (wb >= 0), consensus output
// (2-sequence MSA rows stay with poa.cuh).  Anything else -- and any capacity overflow -- leaves the item
// not-done; the caller then runs it through c3_poa_kernel, which also owns all error reporting.
//
// The per-thread phases are host+device functions: tests/lane_emul.cu runs them on the CPU, 32 states in
// lockstep, against the oracle (no GPU needed).
#pragma once
#include "poa.cuh"

#define C3L_E_RETRY (-299)
#define C3L_VSTRIDE 400            // int4 per 16-column vector of one warp step: 3 arrays x 4 quarters x 32 lanes + 32 x (F1,F2) carry-in
#ifndef C3L_NC
#define C3L_NC 8                   // columns per software-pipeline step of the row loop (8 or 16)
#endif
#ifndef C3L_THREADS
#define C3L_THREADS 64
#endif
#ifndef C3L_MINB
#define C3L_MINB 4
#endif

// The serial per-thread phases (graph walks, backtrack, merge, consensus) loop under warp-uniform control:
// `while (C3L_ANY(cond)) { if (cond) { one step } }` -- every thread of the warp makes the same number of
// trips and the warp reconverges at each one, so 32 reads advance one step per pass instead of running one
// after the other.  All 32 threads must reach such a loop (no early return ahead of it).
#if defined(__CUDA_ARCH__)
#define C3L_ADDMAX(a, b, c) __viaddmax_s32((a), (b), (c))
#define C3L_MAX3(a, b, c) __vimax3_s32((a), (b), (c))
#define C3L_ANY(x) __any_sync(C3_FULL, (x))
#define C3L_LDCS4(p) __ldcs(p)           // DP rows are read once by the successor row: stream them (evict-first)
#define C3L_LDCS1(p) __ldcs(p)
#else
#define C3L_ANY(x) (x)
#define C3L_LDCS4(p) (*(p))
#define C3L_LDCS1(p) (*(p))
static inline int c3l_hmax(int a, int b) { return a > b ? a : b; }
#define C3L_ADDMAX(a, b, c) c3l_hmax((a) + (b), (c))
#define C3L_MAX3(a, b, c) c3l_hmax(c3l_hmax((a), (b)), (c))
#endif

// -DC3L_PROF: per-phase cycle counters (lane 0 of every warp adds clock64() deltas); read with c3_debug_lane_prof
#if defined(C3L_PROF) && defined(__CUDACC__)
__device__ unsigned long long c3l_prof[16];
#endif
#if defined(C3L_PROF) && defined(__CUDA_ARCH__)
#define C3L_TICK(id) do { const long long t1_ = clock64(); if (lane == 0) atomicAdd(&c3l_prof[id], (unsigned long long)(t1_ - tprof_)); tprof_ = clock64(); } while (0)
#define C3L_TICK_INIT long long tprof_ = clock64()
#define C3L_COUNT(id, v) do { if (lane == 0) atomicAdd(&c3l_prof[id], (unsigned long long)(v)); } while (0)
#else
#define C3L_TICK(id) do { } while (0)
#define C3L_TICK_INIT do { } while (0)
#define C3L_COUNT(id, v) do { } while (0)
#endif

struct c3l_state {
    int item, on, err, nseq;
    const uint8_t *ibase; const int32_t *bnd;
    int node_n, pool_n;
    long long cells_total;
    const uint8_t *q; int qlen, n, w, aligning;
    int v, rcount;                                    // row walk: current node, rows done
    c3_nrec nd, nd1, nd2; uint32_t hrv, hr1, hr2;     // look-ahead: records of v and of the next two nodes of the list
    uint2 pe, pe1;                                    // first overflow in-edge of v / next(v) (in_n > 1)
    c3_prow ra, rb;                                   // row records of v's first two predecessors (unless == vlast)
    c3_prow last; int vlast;                          // the row just computed
    int beg, end, nvec, beg_sn, end_sn; c3_prow r0, r1;   // the row between setup and compute
};

// int32 index of column c (relative to the row's band start) of array a in the row stored at off4
C3_HD __forceinline__ int c3l_ci(int off4, int a, int c, int lane)
{
    return ((off4 + (c >> 4) * C3L_VSTRIDE + ((a << 2) + ((c >> 2) & 3)) * 32 + lane) << 2) + (c & 3);
}

// int32 index of the (F1, F2) pair entering vector vi of the row stored at off4
C3_HD __forceinline__ int c3l_fi(int off4, int vi, int lane) { return ((off4 + vi * C3L_VSTRIDE + 384) << 2) + 2 * lane; }

// ---------------------------------------------------------------------------
// item start: first sequence -> linear graph
// ---------------------------------------------------------------------------
C3_HD __forceinline__ void c3l_item_begin(c3l_state &S, const c3_poa_args &A, const c3_poa_ws &W, const int item, const bool have)
{
    S.item = item; S.on = 0; S.err = 0; S.nseq = 0; S.node_n = 0; S.pool_n = 0; S.cells_total = 0; S.aligning = 0;
    S.v = C3_SINK; S.nvec = 0;
    const uint8_t *q = nullptr;
    int L = 0, cnt = 0;
    if (have) {
        const int nseq = A.n_seqs[(int64_t)item * A.n_seqs_stride];
        if (nseq >= A.min_seqs && nseq <= A.max_seqs && nseq >= 1 && !(A.msa2 && nseq == 2)) {
            S.ibase = A.codes + A.item_base[item];
            S.bnd = A.bounds + (int64_t)item * A.max_seqs * 2;
            S.on = 1; S.nseq = nseq;
            q = S.ibase + S.bnd[0];
            L = S.bnd[1] - S.bnd[0];
            if (L <= 0 || L > 65000 || L + 2 > A.node_cap) S.err = C3L_E_RETRY;
            else cnt = L + 2;
        }
    }
    for (int i = 0; C3L_ANY(i < cnt); ++i) {
        if (i >= cnt) continue;
        c3_pnode n;
        n.in_more = n.out_more = C3_NONE; n.rmask = 1; n.spare = 0;
        n.aln0 = n.aln1 = n.aln2 = n.aln3 = C3_NONE; n.max_out = C3_NONE; n.aln_n = 0;
        if (i == C3_SRC) {
            n.base = 4; n.in_n = 0; n.out_n = 1; n.in0 = C3_NONE; n.out0 = 2; n.w0 = 1;
            n.prev = C3_NONE; n.next = 2;
        } else if (i == C3_SINK) {
            n.base = 4; n.in_n = 1; n.out_n = 0; n.in0 = (uint16_t)(L + 1); n.out0 = C3_NONE; n.w0 = 0;
            n.prev = (uint16_t)(L + 1); n.next = C3_NONE;
        } else {
            n.base = q[i - 2]; n.in_n = 1; n.out_n = 1; n.w0 = 1;
            n.in0 = (uint16_t)(i == 2 ? C3_SRC : i - 1);
            n.out0 = (uint16_t)(i == L + 1 ? C3_SINK : i + 1);
            n.prev = n.in0; n.next = n.out0;
        }
        W.nodes[i] = n;
    }
    S.node_n = cnt;
}

// ---------------------------------------------------------------------------
// alignment start: score mode, band half-width, remaining-path lengths, query profile, source-row band.
// Returns the number of vectors of the source row (0: this thread does not align now).
// ---------------------------------------------------------------------------
C3_HD __forceinline__ int c3l_align_begin(c3l_state &S, const c3_poa_args &A, const c3_poa_para_dev &P, const c3_poa_ws &W, const int sq)
{
    S.aligning = 0; S.nvec = 0; S.v = C3_SINK;
    bool act = S.on && !S.err && sq < S.nseq;
    const uint8_t *q = nullptr;
    int qlen = 0;
    if (act) {
        q = S.ibase + S.bnd[2 * sq];
        qlen = S.bnd[2 * sq + 1] - S.bnd[2 * sq];
        const int n = S.node_n;
        const int len = qlen > n ? qlen : n;
        const int max_score = max(qlen * 5, len * P.e1 + P.o1);
        const int pn = (max_score <= 32767 - P.mismatch - P.o1 - P.e1) ? P.simd_bits / 16 : P.simd_bits / 32;
        if (qlen <= 0 || qlen > 65000 || qlen + 32 > A.qp_stride || pn != 16 || P.wb < 0) { S.err = C3L_E_RETRY; act = false; qlen = 0; }
        else { S.q = q; S.qlen = qlen; S.n = n; S.w = P.wb + (int)(P.wf * (double)qlen); }
    }
    // remaining path length along the heaviest out-edges: hops(v -> sink), reverse list walk
    int v = C3_NONE;
    if (act) { W.hr[C3_SINK] = C3_SINK; v = W.nodes[C3_SINK].prev; }
    while (C3L_ANY(v != C3_NONE)) {
        if (v == C3_NONE) continue;
        const c3_pnode *nd = &W.nodes[v];
        int best_w = nd->w0, best = nd->out0;
        uint16_t e = nd->out_more;
        while (e != C3_NONE) {
            const c3_pedge pe = W.pool[e];
            if ((int)pe.w > best_w) { best_w = pe.w; best = pe.id; }
            e = pe.next;
        }
        W.hr[v] = ((W.hr[best] >> 16) + 1u) << 16;
        v = nd->prev;
    }
    // query profile: int8 scores per node base, 16 columns per store; column j scores q[j-1], j = 0 scores 0
    {
        const int qs = A.qp_stride, jmax = act ? ((qlen + 16) & ~15) : 0;   // the DP reads vectors up to column qlen | 15
        for (int j0 = 0; C3L_ANY(j0 < jmax); j0 += 16) {
            if (j0 >= jmax) continue;
            uint32_t wv[4][4];
#pragma unroll
            for (int b4 = 0; b4 < 4; ++b4)
#pragma unroll
                for (int t = 0; t < 4; ++t) wv[b4][t] = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int j = j0 + k;
                if (j >= 1 && j <= qlen) {
                    const int qc = q[j - 1];
#pragma unroll
                    for (int b4 = 0; b4 < 4; ++b4)
                        wv[b4][k >> 2] |= (uint32_t)(uint8_t)(int8_t)c3_score(P, b4, qc) << (8 * (k & 3));
                }
            }
#pragma unroll
            for (int b4 = 0; b4 < 4; ++b4)
                *reinterpret_cast<uint4 *>(W.qp + b4 * qs + j0) = make_uint4(wv[b4][0], wv[b4][1], wv[b4][2], wv[b4][3]);
        }
    }
    if (!act) return 0;
    // source row band
    const int rem = (int)(W.hr[C3_SRC] >> 16) - 1;
    const int rr = qlen - rem;
    const int beg = max(0, min(0, rr) - S.w);
    const int end = min(qlen, max(0, rr) + S.w);
    const int b0 = (beg >> 4) << 4, e0 = min(qlen, (((end >> 4) + 1) << 4) - 1);
    S.beg = b0; S.end = e0; S.nvec = (e0 - b0 + 16) >> 4;
    S.aligning = 1;
    return S.nvec;
}

// node / edge / row records by index, with the list terminator (C3_NONE) mapped to the sink so that the
// look-ahead past the end of the walk stays inside the workspace
C3_HD __forceinline__ int c3l_cl(const int v) { return v == (int)C3_NONE ? C3_SINK : v; }
C3_HD __forceinline__ uint2 c3l_ld_edge(const c3_pedge *p) { return *reinterpret_cast<const uint2 *>(p); }
#define C3L_E_ID(e) ((int)((e).x & 0xffffu))
#define C3L_E_NEXT(e) ((int)((e).y & 0xffffu))

// source row cells + records; positions the row walk on the first node after the source and fills the
// look-ahead (records of the next two nodes, first overflow in-edges, predecessor row records)
C3_HD __forceinline__ void c3l_source_row(c3l_state &S, const c3_poa_para_dev &P, const c3_poa_ws &W, int32_t *ar, const int lane)
{
    if (!S.aligning) return;
    const int oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const int b0 = S.beg, wd = S.end - S.beg + 1;
    c3_prow ri; ri.off = 0; ri.beg = (uint16_t)b0; ri.end = (uint16_t)S.end; ri.mp = 1;   // successors of the source start at column 1
    ri.in0 = C3_NONE; ri.base = 4; ri.npre = 0; ri.link = 0;
    W.rows[C3_SRC] = ri;
    S.last = ri; S.vlast = C3_SRC;
    ri.mp = C3_NONE;
    W.ord[0] = ri;
    for (int vi = 0; vi < S.nvec; ++vi)
        *reinterpret_cast<int2 *>(ar + c3l_fi(0, vi, lane)) = make_int2(C3_NEG_INF, C3_NEG_INF);
    for (int c = 0; c < 16 * S.nvec; ++c) {
        int h = C3_NEG_INF, x1 = C3_NEG_INF, x2 = C3_NEG_INF;
        if (b0 == 0 && c < wd) {
            if (c == 0) { h = 0; x1 = -oe1; x2 = -oe2; }
            else h = max(-(P.o1 + P.e1 * c), -(P.o2 + P.e2 * c));
        }
        ar[c3l_ci(0, 0, c, lane)] = h; ar[c3l_ci(0, 1, c, lane)] = x1; ar[c3l_ci(0, 2, c, lane)] = x2;
    }
    S.v = W.nodes[C3_SRC].next; S.rcount = 1;
    S.nd = c3_ld_node(&W.nodes[S.v]); S.hrv = W.hr[S.v];
    const int v1 = c3l_cl(C3_N_NEXT(S.nd));
    S.nd1 = c3_ld_node(&W.nodes[v1]); S.hr1 = W.hr[v1];
    const int v2 = c3l_cl(C3_N_NEXT(S.nd1));
    S.nd2 = c3_ld_node(&W.nodes[v2]); S.hr2 = W.hr[v2];
    S.pe = make_uint2(0u, 0u); S.pe1 = make_uint2(0u, 0u);
    if (C3_N_INN(S.nd) > 1) S.pe = c3l_ld_edge(&W.pool[C3_N_INMORE(S.nd)]);
    if (C3_N_INN(S.nd1) > 1) S.pe1 = c3l_ld_edge(&W.pool[C3_N_INMORE(S.nd1)]);
    S.ra = ri; S.rb = ri;
    if (C3_N_IN0(S.nd) != C3_SRC) S.ra = W.rows[c3l_cl(C3_N_IN0(S.nd))];
    if (C3_N_INN(S.nd) > 1 && C3L_E_ID(S.pe) != C3_SRC) S.rb = W.rows[C3L_E_ID(S.pe)];
    S.nvec = 0;
}

// ---------------------------------------------------------------------------
// row setup: adaptive band of the current node's row.  Returns its number of vectors (0: no row).
// The records of the first two predecessors were requested while the previous row was computed (S.ra, S.rb);
// a predecessor that IS the previous row comes from S.last.
// ---------------------------------------------------------------------------
C3_HD __forceinline__ int c3l_row_setup(c3l_state &S, const c3_poa_args &A, const c3_poa_ws &W)
{
    S.nvec = 0;
    if (!S.aligning || S.err || S.v == C3_SINK) return 0;
    const int qlen = S.qlen, w = S.w;
    const int rem = (int)(S.hrv >> 16) - 1;
    const int rr = qlen - rem;
    const int npre = C3_N_INN(S.nd);
    if (npre > C3_MAXPRE) { S.err = C3L_E_RETRY; return 0; }       // c3_poa_kernel's limit: let it report
    const c3_prow r0 = (C3_N_IN0(S.nd) == S.vlast) ? S.last : S.ra;
    int mpl = min(S.n, (int)r0.mp), mpr = r0.mp, min_pre_beg = r0.beg;
    c3_prow r1; r1.off = 0; r1.beg = 16; r1.end = 0; r1.mp = 0; r1.in0 = C3_NONE; r1.link = 0; r1.base = 4; r1.npre = 0;   // empty band
    if (npre > 1) {
        r1 = (C3L_E_ID(S.pe) == S.vlast) ? S.last : S.rb;
        mpl = min(mpl, (int)r1.mp); mpr = max(mpr, (int)r1.mp); min_pre_beg = min(min_pre_beg, (int)r1.beg);
        int e = C3L_E_NEXT(S.pe);
        for (int k = 2; k < npre; ++k) {
            const c3_pedge pe = W.pool[e]; e = pe.next;
            const c3_prow ri = W.rows[pe.id];
            mpl = min(mpl, (int)ri.mp); mpr = max(mpr, (int)ri.mp); min_pre_beg = min(min_pre_beg, (int)ri.beg);
        }
    }
    int beg = max(0, min(mpl, rr) - w);
    int end = min(qlen, max(mpr, rr) + w);
    const int beg_sn = max(beg >> 4, min_pre_beg >> 4);
    const int end_sn = max(end >> 4, beg_sn);
    beg = beg_sn << 4; end = min(qlen, ((end_sn + 1) << 4) - 1);
    if (end - beg + 1 <= 0) { S.err = C3L_E_RETRY; return 0; }
    S.beg = beg; S.end = end; S.beg_sn = beg_sn; S.end_sn = end_sn; S.r0 = r0; S.r1 = r1;
    S.nvec = end_sn - beg_sn + 1;
    return S.nvec;
}

// C3L_NC columns (from j0, a multiple of C3L_NC) of one predecessor's H, E1, E2, or NEG_INF outside its band
C3_HD __forceinline__ void c3l_load_part(const int4 *ar4, const c3_prow &rp, const int j0, const int lane,
                                         int (&hv)[C3L_NC], int (&v1)[C3L_NC], int (&v2)[C3L_NC])
{
    const int pb = rp.beg, pe = rp.end;
    if (j0 >= pb && j0 <= pe) {
        const int c = j0 - pb;
        const int4 *src = ar4 + rp.off + (c >> 4) * C3L_VSTRIDE + ((c >> 2) & 3) * 32 + lane;
#pragma unroll
        for (int qd = 0; qd < C3L_NC / 4; ++qd) {
            const int4 a = C3L_LDCS4(src + qd * 32), b = C3L_LDCS4(src + (4 + qd) * 32), c = C3L_LDCS4(src + (8 + qd) * 32);
            hv[4 * qd] = a.x; hv[4 * qd + 1] = a.y; hv[4 * qd + 2] = a.z; hv[4 * qd + 3] = a.w;
            v1[4 * qd] = b.x; v1[4 * qd + 1] = b.y; v1[4 * qd + 2] = b.z; v1[4 * qd + 3] = b.w;
            v2[4 * qd] = c.x; v2[4 * qd + 1] = c.y; v2[4 * qd + 2] = c.z; v2[4 * qd + 3] = c.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < C3L_NC; ++k) { hv[k] = C3_NEG_INF; v1[k] = C3_NEG_INF; v2[k] = C3_NEG_INF; }
    }
}

// ---------------------------------------------------------------------------
// row compute: all columns of the row set up by c3l_row_setup in steps of C3L_NC, row record, advance to the
// next node.  While the cells are computed, everything the NEXT row's setup reads is requested: the row
// records of its first two predecessors, the node record two rows ahead and that node's first overflow edge.
// ---------------------------------------------------------------------------
C3_HD __forceinline__ void c3l_row_compute(c3l_state &S, const c3_poa_args &A, const c3_poa_para_dev &P, const c3_poa_ws &W,
                                           int32_t *ar, const int base4, const int lane)
{
    if (S.nvec <= 0) return;
    const int e1 = P.e1, e2 = P.e2, oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const int v = S.v;
    // ---- look-ahead for the next row (v1) and the one after (v2) ----
    const int v1 = c3l_cl(C3_N_NEXT(S.nd)), v2 = c3l_cl(C3_N_NEXT(S.nd1)), v3 = c3l_cl(C3_N_NEXT(S.nd2));
    const c3_nrec nd3 = c3_ld_node(&W.nodes[v3]);
    const uint32_t hr3 = W.hr[v3];
    uint2 pe2 = make_uint2(0u, 0u);
    if (C3_N_INN(S.nd2) > 1) pe2 = c3l_ld_edge(&W.pool[C3_N_INMORE(S.nd2)]);
    c3_prow ra_n = S.last, rb_n = S.last;
    {
        const int in0n = c3l_cl(C3_N_IN0(S.nd1));
        if (in0n != v) ra_n = W.rows[in0n];
        if (C3_N_INN(S.nd1) > 1 && C3L_E_ID(S.pe1) != v) rb_n = W.rows[C3L_E_ID(S.pe1)];
    }
    (void)v2;
    const int npre = C3_N_INN(S.nd), nbase = C3_N_BASE(S.nd);
    const int beg = S.beg, end = S.end, nvec = S.nvec;
    const c3_prow r0 = S.r0, r1 = S.r1;
    const int4 *ar4 = reinterpret_cast<const int4 *>(ar);
    int4 *out4 = reinterpret_cast<int4 *>(ar) + base4 + lane;
    const int8_t *qprow = W.qp + (nbase < 4 ? nbase : 0) * A.qp_stride;
    int f1 = C3_NEG_INF, f2 = C3_NEG_INF, carry0 = C3_NEG_INF, carry1 = C3_NEG_INF;
    int bestkey = -0x7fffffff - 1;
    // software pipeline: both predecessors' cells and the profile words of step h+1 are requested before
    // step h is computed
    int nh[C3L_NC], nx1[C3L_NC], nx2[C3L_NC], ph[C3L_NC], px1[C3L_NC], px2[C3L_NC];
    uint32_t nsw[C3L_NC / 4];
#pragma unroll
    for (int t = 0; t < C3L_NC / 4; ++t) nsw[t] = 0u;
    c3l_load_part(ar4, r0, beg, lane, nh, nx1, nx2);
    c3l_load_part(ar4, r1, beg, lane, ph, px1, px2);
    if (nbase < 4) {
#pragma unroll
        for (int t = 0; t < C3L_NC / 4; ++t) nsw[t] = *reinterpret_cast<const uint32_t *>(qprow + beg + 4 * t);
    }
    const int nstep = nvec * (16 / C3L_NC);
    for (int h = 0; h < nstep; ++h) {
        const int j0 = beg + C3L_NC * h;
        int m[C3L_NC], x1[C3L_NC], x2[C3L_NC];
        uint32_t swv[C3L_NC / 4];
#pragma unroll
        for (int t = 0; t < C3L_NC / 4; ++t) swv[t] = nsw[t];
        m[0] = max(carry0, carry1);
#pragma unroll
        for (int k = 1; k < C3L_NC; ++k) m[k] = max(nh[k - 1], ph[k - 1]);
        carry0 = nh[C3L_NC - 1]; carry1 = ph[C3L_NC - 1];
#pragma unroll
        for (int k = 0; k < C3L_NC; ++k) { x1[k] = max(nx1[k], px1[k]); x2[k] = max(nx2[k], px2[k]); }
        if (h + 1 < nstep) {
            c3l_load_part(ar4, r0, j0 + C3L_NC, lane, nh, nx1, nx2);
            c3l_load_part(ar4, r1, j0 + C3L_NC, lane, ph, px1, px2);
            if (nbase < 4) {
#pragma unroll
                for (int t = 0; t < C3L_NC / 4; ++t) nsw[t] = *reinterpret_cast<const uint32_t *>(qprow + j0 + C3L_NC + 4 * t);
            }
        }
        if (npre > 2) {                                  // third and further predecessors: rare, not pipelined
            int e = C3L_E_NEXT(S.pe);
            for (int k = 2; k < npre; ++k) {
                const c3_pedge pe = W.pool[e]; e = pe.next;
                const c3_prow rp = W.rows[pe.id];
                int hv[C3L_NC], v1[C3L_NC], v2[C3L_NC];
                c3l_load_part(ar4, rp, j0, lane, hv, v1, v2);
                const int jc = j0 - 1;
                int prev = C3_NEG_INF;
                if (h > 0 && jc >= (int)rp.beg && jc <= (int)rp.end) prev = C3L_LDCS1(ar + c3l_ci(rp.off, 0, jc - rp.beg, lane));
                m[0] = max(m[0], prev);
#pragma unroll
                for (int t = 1; t < C3L_NC; ++t) m[t] = max(m[t], hv[t - 1]);
#pragma unroll
                for (int t = 0; t < C3L_NC; ++t) { x1[t] = max(x1[t], v1[t]); x2[t] = max(x2[t], v2[t]); }
            }
        }
        const int vi = (C3L_NC * h) >> 4;
        if (((C3L_NC * h) & 15) == 0)                   // F entering this 16-column vector: the backtrack restarts from it
            *reinterpret_cast<int2 *>(ar + c3l_fi(base4, vi, lane)) = make_int2(f1, f2);
        int hme[C3L_NC];
#pragma unroll
        for (int k = 0; k < C3L_NC; ++k) {
            const int sc = (int)(int8_t)(swv[k >> 2] >> (8 * (k & 3)));
            hme[k] = C3L_MAX3(m[k] + sc, x1[k], x2[k]);
        }
        const int lim = end - j0;                       // last active column of this step (>= C3L_NC - 1: all)
        if (lim < C3L_NC - 1) {
#pragma unroll
            for (int k = 0; k < C3L_NC; ++k) if (k > lim) hme[k] = C3_NEG_INF;
        }
        int hh[C3L_NC], n1[C3L_NC], n2[C3L_NC];
#pragma unroll
        for (int k = 0; k < C3L_NC; ++k) {
            hh[k] = C3L_MAX3(hme[k], f1, f2);
            f1 = C3L_ADDMAX(f1, -e1, hme[k] - oe1);
            f2 = C3L_ADDMAX(f2, -e2, hme[k] - oe2);
            n1[k] = C3L_ADDMAX(hh[k], -oe1, x1[k] - e1);
            n2[k] = C3L_ADDMAX(hh[k], -oe2, x2[k] - e2);
        }
        if (lim < C3L_NC - 1) {
#pragma unroll
            for (int k = 0; k < C3L_NC; ++k) if (k > lim) { hh[k] = C3_NEG_INF; n1[k] = C3_NEG_INF; n2[k] = C3_NEG_INF; }
        }
        int4 *dst = out4 + vi * C3L_VSTRIDE + (((C3L_NC * h) >> 2) & 3) * 32;
#pragma unroll
        for (int qd = 0; qd < C3L_NC / 4; ++qd) {
            dst[qd * 32] = make_int4(hh[4 * qd], hh[4 * qd + 1], hh[4 * qd + 2], hh[4 * qd + 3]);
            dst[(4 + qd) * 32] = make_int4(n1[4 * qd], n1[4 * qd + 1], n1[4 * qd + 2], n1[4 * qd + 3]);
            dst[(8 + qd) * 32] = make_int4(n2[4 * qd], n2[4 * qd + 1], n2[4 * qd + 2], n2[4 * qd + 3]);
        }
        // simd_abpoa_ada_max_i as one packed max: value in the high half, tie-break priority in the low half
        // (lowest SIMD lane, then the last vector, then the earliest vector)
        const int vp = (vi == nvec - 1) ? 0xfff : (0xffe - vi);
        const int k16 = (C3L_NC * h) & 15;
#pragma unroll
        for (int k = 0; k < C3L_NC; ++k) {
            const int hc = max(hh[k], -32768);
            bestkey = C3L_ADDMAX((int)((unsigned)hc << 16) + vp, (15 - k16 - k) << 12, bestkey);
        }
    }
    int best_i = -1;
    if ((bestkey >> 16) > -32768) {
        const int sl = 15 - ((bestkey >> 12) & 15);
        const int vp = bestkey & 0xfff;
        const int sn = (vp == 0xfff) ? S.end_sn : S.beg_sn + (0xffe - vp);
        best_i = (sn << 4) + sl;
    }
    c3_prow ri; ri.off = base4; ri.beg = (uint16_t)beg; ri.end = (uint16_t)end; ri.mp = (uint16_t)(best_i + 1);
    ri.in0 = (uint16_t)C3_N_IN0(S.nd); ri.base = (uint8_t)nbase; ri.npre = (uint8_t)npre;
    ri.link = (uint16_t)S.rcount;
    W.rows[v] = ri;
    S.last = ri; S.vlast = v;
    ri.link = (uint16_t)v; ri.mp = r0.link;
    W.ord[S.rcount] = ri;
    S.cells_total += end - beg + 1;
    ++S.rcount;
    S.v = v1; S.nd = S.nd1; S.hrv = S.hr1; S.nd1 = S.nd2; S.hr1 = S.hr2; S.nd2 = nd3; S.hr2 = hr3;
    S.pe = S.pe1; S.pe1 = pe2; S.ra = ra_n; S.rb = rb_n;
    S.nvec = 0;
}

// ---------------------------------------------------------------------------
// alignment end: backtrack (abPOA's M -> E1 -> E2 -> F1 -> F2 order and op-mask state machine) and graph
// merge (abpoa_add_graph_alignment), one thread per read, one step per warp-uniform trip
// ---------------------------------------------------------------------------
C3_HD __forceinline__ void c3l_align_end(c3l_state &S, const c3_poa_args &A, const c3_poa_para_dev &P, const c3_poa_ws &W,
                                         const int32_t *ar, const int lane, const int sq)
{
    bool run = S.aligning && !S.err;
    S.aligning = 0;
    C3L_TICK_INIT;
    const int e1 = P.e1, e2 = P.e2, oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const uint8_t *q = S.q; const int qlen = S.qlen;
    unsigned long long *cg = W.cigar;
    int nc = 0, j = 0, hij = 0;
    c3_prow rt; rt.off = 0; rt.beg = rt.end = 0; rt.mp = C3_NONE; rt.in0 = C3_NONE; rt.link = C3_SRC; rt.base = 4; rt.npre = 0;
    if (run) {
        const c3_nrec sk = c3_ld_node(&W.nodes[C3_SINK]);
        int best_score = -0x7fffffff - 1, bj = -1, bk = -1;
        int e = C3_N_INMORE(sk);
        const int skn = C3_N_INN(sk);
        for (int k = 0; k < skn; ++k) {
            int p;
            if (k == 0) p = C3_N_IN0(sk); else { const c3_pedge pe = W.pool[e]; p = pe.id; e = pe.next; }
            const c3_prow rp = W.rows[p];
            const int en = min(qlen, (int)rp.end);
            const int val = ar[c3l_ci(rp.off, 0, en - rp.beg, lane)];
            if (val > best_score) { best_score = val; bj = en; bk = rp.link; }
        }
        if (bk < 0 || qlen - bj + 8 > A.cigar_cap) { S.err = C3L_E_RETRY; run = false; }
        else {
            j = bj; rt = W.ord[bk];
            for (int t = qlen; t > bj; --t)
                cg[qlen - t] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(t - 1) << 32);
            nc = qlen - bj;
            hij = ar[c3l_ci(rt.off, 0, j - rt.beg, lane)];
        }
    }
    int cur_op = C3_OP_ALL;
    while (C3L_ANY(run && rt.link != C3_SRC && j > 0)) {
        C3L_COUNT(11, 1);
        if (!(run && rt.link != C3_SRC && j > 0)) continue;
        const int i = rt.link;
        const int b = rt.beg;
        int hit = 0;
        unsigned long long opw = 0;
        if (j >= b && j <= (int)rt.end) {
            const int s = c3_score(P, rt.base, q[j - 1]);
            const int npre = rt.npre;
            const int in_more = npre > 1 ? (int)W.nodes[i].in_more : (int)C3_NONE;
            if (cur_op & C3_OP_M) {
                int e = in_more;
                for (int k = 0; k < npre; ++k) {
                    int pk;
                    if (k == 0) pk = rt.mp; else { const c3_pedge pe = W.pool[e]; pk = W.rows[pe.id].link; e = pe.next; }
                    const c3_prow pr = W.ord[pk];
                    if (j - 1 < max((int)pr.beg, b) || j - 1 > (int)pr.end) continue;
                    const int ph = ar[c3l_ci(pr.off, 0, j - 1 - pr.beg, lane)];
                    if (ph + s == hij) {
                        opw = C3_CG_MATCH | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                        rt = pr; --j; hit = 1; cur_op = C3_OP_ALL; hij = ph;
                        break;
                    }
                }
            }
            if (!hit && (cur_op & C3_OP_E)) {
                int e = in_more;
                for (int k = 0; k < npre; ++k) {
                    int pk;
                    if (k == 0) pk = rt.mp; else { const c3_pedge pe = W.pool[e]; pk = W.rows[pe.id].link; e = pe.next; }
                    const c3_prow pr = W.ord[pk];
                    if (j < (int)pr.beg || j > (int)pr.end) continue;
                    const int pc = j - pr.beg;
                    const int ph = ar[c3l_ci(pr.off, 0, pc, lane)], pe1 = ar[c3l_ci(pr.off, 1, pc, lane)], pe2 = ar[c3l_ci(pr.off, 2, pc, lane)];
                    if (cur_op & C3_OP_E1) {
                        if (cur_op & C3_OP_M) {
                            if (hij == pe1) { cur_op = (ph - oe1 == pe1) ? (C3_OP_M | C3_OP_F) : C3_OP_E1; hit = 1; }
                        } else if (ar[c3l_ci(rt.off, 1, j - b, lane)] == pe1 - e1) {
                            cur_op = (ph - oe1 == pe1) ? (C3_OP_M | C3_OP_F) : C3_OP_E1; hit = 1;
                        }
                    }
                    if (!hit && (cur_op & C3_OP_E2)) {
                        if (cur_op & C3_OP_M) {
                            if (hij == pe2) { cur_op = (ph - oe2 == pe2) ? (C3_OP_M | C3_OP_F) : C3_OP_E2; hit = 1; }
                        } else if (ar[c3l_ci(rt.off, 2, j - b, lane)] == pe2 - e2) {
                            cur_op = (ph - oe2 == pe2) ? (C3_OP_M | C3_OP_F) : C3_OP_E2; hit = 1;
                        }
                    }
                    if (hit) {
                        opw = C3_CG_DEL | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                        rt = pr; hij = ph;
                        break;
                    }
                }
            }
            if (!hit && (cur_op & C3_OP_F) && j - 1 >= b) {
                // F is stored only where it enters a 16-column vector: rebuild F[j-1] and F[j] from there
                const int cm = j - 1 - b, vs = cm >> 4, tm = cm & 15;
                const int2 fin = *reinterpret_cast<const int2 *>(ar + c3l_fi(rt.off, vs, lane));
                const int4 *hp = reinterpret_cast<const int4 *>(ar) + rt.off + vs * C3L_VSTRIDE + lane;
                int hq[16];
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    const int4 a4 = hp[qd * 32];
                    hq[4 * qd] = a4.x; hq[4 * qd + 1] = a4.y; hq[4 * qd + 2] = a4.z; hq[4 * qd + 3] = a4.w;
                }
                int f1 = fin.x, f2 = fin.y, f1l = C3_NEG_INF, f2l = C3_NEG_INF, hl = C3_NEG_INF;
                C3L_COUNT(12, 1);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (c <= tm) {
                        hl = hq[c];
                        f1l = f1; f2l = f2;
                        f1 = max(f1 - e1, hl - oe1); f2 = max(f2 - e2, hl - oe2);
                    }
                }
                if (cur_op & C3_OP_F1) {
                    if (!(cur_op & C3_OP_M) || hij == f1) {
                        if (hl - oe1 == f1) { cur_op = C3_OP_M | C3_OP_E; hit = 1; }
                        else if (f1l - e1 == f1) { cur_op = C3_OP_F1; hit = 1; }
                    }
                }
                if (!hit && (cur_op & C3_OP_F2)) {
                    if (!(cur_op & C3_OP_M) || hij == f2) {
                        if (hl - oe2 == f2) { cur_op = C3_OP_M | C3_OP_E; hit = 1; }
                        else if (f2l - e2 == f2) { cur_op = C3_OP_F2; hit = 1; }
                    }
                }
                if (hit) {
                    opw = C3_CG_INS | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                    --j; hij = hl;
                }
            }
        }
        if (!hit) { S.err = C3L_E_RETRY; run = false; continue; }
        cg[nc] = opw;
        ++nc;
        if (nc + j + 8 > A.cigar_cap) { S.err = C3L_E_RETRY; run = false; }
    }
    if (run) {
        for (int t = j; t > 0; --t)
            cg[nc + j - t] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(t - 1) << 32);
        nc += j;
    } else nc = 0;

    C3L_TICK(6);
    // ---- merge: the cigar is walked from its tail = forward order ----
    c3_graph g; g.nodes = W.nodes; g.pool = W.pool; g.node_n = S.node_n; g.pool_n = S.pool_n;
    g.node_cap = A.node_cap; g.pool_cap = A.pool_cap; g.err = 0;
    int last_id = C3_SRC, last_new = 0;
    for (int t = nc - 1; C3L_ANY(t >= 0 && !g.err); --t) {
        if (!(t >= 0 && !g.err)) continue;
        const unsigned long long opc = cg[t];
        const int kc = (int)(opc & 0xff), nid = (int)((opc >> 8) & 0xffff), qp = (int)(opc >> 32);
        if (kc == (int)C3_CG_DEL) continue;
        if (kc == (int)C3_CG_MATCH) {
            const uint8_t bq = q[qp];
            const c3_pnode nm = g.nodes[nid];
            if (nm.base != bq) {
                int al = -1;
                for (int k = 0; k < nm.aln_n; ++k) {
                    const int a = c3_aln_get(nm, k);
                    if (g.nodes[a].base == bq) { al = a; break; }
                }
                if (al != -1) {
                    c3_g_add_edge(g, last_id, al, 1 - last_new);
                    last_id = al; last_new = 0;
                    if (sq < 16) g.nodes[al].rmask |= (uint16_t)(1u << sq);
                } else {
                    const int id = c3_g_add_node(g, bq);
                    if (!g.err) {
                        c3_list_insert_before(g, id, nid);
                        c3_g_add_edge(g, last_id, id, 0);
                        last_id = id; last_new = 1;
                        if (sq < 16) g.nodes[id].rmask = (uint16_t)(1u << sq);
                        for (int k = 0; k < nm.aln_n; ++k) {
                            const int a = c3_aln_get(nm, k);
                            c3_aln_push(&g.nodes[a], (uint16_t)id);
                            c3_aln_push(&g.nodes[id], (uint16_t)a);
                        }
                        c3_aln_push(&g.nodes[nid], (uint16_t)id);
                        c3_aln_push(&g.nodes[id], (uint16_t)nid);
                    }
                }
            } else {
                c3_g_add_edge(g, last_id, nid, 1 - last_new);
                last_id = nid; last_new = 0;
                if (sq < 16) g.nodes[nid].rmask |= (uint16_t)(1u << sq);
            }
        } else {
            const int id = c3_g_add_node(g, q[qp]);
            if (!g.err) {
                c3_list_insert_after(g, id, c3_group_tail(g, last_id));
                c3_g_add_edge(g, last_id, id, 0);
                last_id = id; last_new = 1;
                if (sq < 16) g.nodes[id].rmask = (uint16_t)(1u << sq);
            }
        }
    }
    C3L_TICK(7);
    if (run) {
        if (!g.err) c3_g_add_edge(g, last_id, C3_SINK, 1 - last_new);
        if (g.err) S.err = C3L_E_RETRY;
        else { S.node_n = g.node_n; S.pool_n = g.pool_n; }
    }
}

// ---------------------------------------------------------------------------
// item end: heaviest bundling (abpoa_heaviest_bundling) + consensus walk and outputs.  A failed item writes
// no result (done stays 0).
// ---------------------------------------------------------------------------
C3_HD __forceinline__ void c3l_item_end(c3l_state &S, const c3_poa_args &A, const c3_poa_ws &W, int32_t *done)
{
    bool act = S.on && !S.err;
    char *co = A.cons + (int64_t)S.item * A.cons_cap;
    int32_t *score = (int32_t *)W.hr;
    int v = act ? C3_SINK : C3_NONE;
    while (C3L_ANY(v != C3_NONE)) {
        if (v == C3_NONE) continue;
        c3_pnode *nd = &W.nodes[v];
        if (v == C3_SINK) { nd->max_out = C3_NONE; score[v] = 0; }
        else if (v == C3_SRC) {
            int max_id = -1, path_score = -1, path_w = -1;
            uint16_t e = nd->out_more;
            for (int k = 0; k < nd->out_n; ++k) {
                int o, wv;
                if (k == 0) { o = nd->out0; wv = nd->w0; } else { const c3_pedge pe = W.pool[e]; o = pe.id; wv = pe.w; e = pe.next; }
                if (wv > path_w || (wv == path_w && score[o] > path_score)) { max_id = o; path_score = score[o]; path_w = wv; }
            }
            nd->max_out = (uint16_t)max_id;
        } else {
            int max_w = -0x7fffffff - 1, max_id = -1;
            uint16_t e = nd->out_more;
            for (int k = 0; k < nd->out_n; ++k) {
                int o, wv;
                if (k == 0) { o = nd->out0; wv = nd->w0; } else { const c3_pedge pe = W.pool[e]; o = pe.id; wv = pe.w; e = pe.next; }
                if (max_w < wv) { max_w = wv; max_id = o; }
                else if (max_w == wv && score[max_id] <= score[o]) max_id = o;
            }
            score[v] = max_w + score[max_id];
            nd->max_out = (uint16_t)max_id;
        }
        v = nd->prev;
    }
    int cons_len = 0;
    int id = act ? (int)W.nodes[C3_SRC].max_out : C3_SINK;
    while (C3L_ANY(id != C3_SINK)) {
        if (id == C3_SINK) continue;
        if (id == C3_NONE || cons_len >= A.cons_cap) { act = false; id = C3_SINK; continue; }
        const c3_pnode nd = W.nodes[id];
        co[cons_len++] = "ACGTN"[nd.base];
        id = nd.max_out;
    }
    if (!act) return;
    const int64_t o = (int64_t)S.item * A.out_stride;
    A.status[o] = 0;
    A.cons_len[o] = cons_len;
    A.nodes_out[o] = S.node_n;
    *(long long *)((int32_t *)A.cells_out + (int64_t)S.item * A.cells_stride) = S.cells_total;
    done[S.item] = 1;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// the kernel: persistent warps, 32 items per fetch
// ---------------------------------------------------------------------------
struct c3_lane_args {
    c3_poa_args A;                 // ws / ws_stride: per-THREAD workspace (cell_cap = 0); order / n_work: eligible items
    int4 *arena; long long arena_stride4;   // per-warp DP arena, in int4
    int arena_cap4;                // int4 per warp
    int32_t *done;                 // [n_items] 1 = finished here
};

__global__ void __launch_bounds__(C3L_THREADS, C3L_MINB) c3_poa_lane_kernel(c3_lane_args L)
{
    const c3_poa_args &A = L.A;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const c3_poa_ws W = c3_poa_ws_carve(A.ws + ((int64_t)gwarp * 32 + lane) * A.ws_stride, A.node_cap, A.pool_cap, 0, A.cigar_cap);
    int32_t *ar = reinterpret_cast<int32_t *>(L.arena + (int64_t)gwarp * L.arena_stride4);
    const c3_poa_para_dev P = A.P;
    c3l_state S;
    C3L_TICK_INIT;
    for (;;) {
        int first = 0;
        if (lane == 0) first = (int)atomicAdd(A.counter, 32u);
        first = __shfl_sync(C3_FULL, first, 0);
        if (first >= A.n_work) break;
        const bool have = first + lane < A.n_work;
        const int item = have ? (A.order ? A.order[first + lane] : first + lane) : 0;
        c3l_item_begin(S, A, W, item, have);
        __syncwarp();
        C3L_TICK(0);
        const int max_nseq = __reduce_max_sync(C3_FULL, (S.on && !S.err) ? S.nseq : 0);
        for (int sq = 1; sq < max_nseq; ++sq) {
            int nv = c3l_align_begin(S, A, P, W, sq);
            __syncwarp();
            int mv = __reduce_max_sync(C3_FULL, nv);
            C3L_TICK(1);
            if (mv == 0) continue;
            int used4 = mv * C3L_VSTRIDE;
            if (used4 > L.arena_cap4) { if (S.aligning) { S.err = C3L_E_RETRY; S.aligning = 0; } continue; }
            c3l_source_row(S, P, W, ar, lane);
            __syncwarp();
            C3L_TICK(2);
            for (;;) {
                nv = c3l_row_setup(S, A, W);
                __syncwarp();
                mv = __reduce_max_sync(C3_FULL, nv);
                C3L_TICK(3);
                if (mv == 0) break;
                if (used4 + mv * C3L_VSTRIDE > L.arena_cap4) { if (S.aligning) S.err = C3L_E_RETRY; break; }
                c3l_row_compute(S, A, P, W, ar, used4, lane);
                used4 += mv * C3L_VSTRIDE;
                __syncwarp();
                C3L_TICK(4);
                C3L_COUNT(9, 1); C3L_COUNT(10, mv);
            }
            c3l_align_end(S, A, P, W, ar, lane, sq);
            __syncwarp();
            C3L_TICK(5);
        }
        c3l_item_end(S, A, W, L.done);
        __syncwarp();
        C3L_TICK(8);
    }
}
#endif
