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
// DP rows live in a per-warp arena, interleaved by lane so that the warp's loads and stores are coalesced.
// Cells are STORED as int16 (abPOA's own lane width in this score mode; arithmetic stays int32 in registers):
// the s-th row step of a warp owns `mv` vectors (mv = widest band among the 32 threads, in 16-column vectors
// = abPOA's int16 SIMD granule); vector vi holds, per lane, H / E1 / E2 as 3 x 2 uint4 (8 columns each) and the
// (F1, F2) pair entering the vector:
//     uint4 index  base4 + vi*208 + (array*2 + half)*32 + lane,       F pair: int2 at uint4 index base4 + vi*208 + 192
// "Unreachable" is stored as C3L_FLOOR (-30720): every packed operation clamps there, and a loaded H at the floor
// reads back as NEG_INF in the backtrack.  That is exact as long as every reachable cell of a row stays well above
// the floor, which the row loop checks from the row's first cell (else the read goes to the warp kernel).
// Columns past a row's band end inside its last vector hold the floor, so successors load whole vectors.
// The row just computed is also kept in shared memory (same packed form, ring of `smR` vector slots per lane,
// slot = absolute vector index mod smR): the usual first predecessor -- the previous row -- never comes from L2.
// Graph, row records, cigar and the column codes of the query are thread-private (c3_poa_ws without cells).
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
#define C3L_VSTRIDE 208            // uint4 per 16-column vector of one warp step: 3 arrays x 2 halves x 32 lanes + 32 x (F1,F2) carry-in
#ifndef C3L_SMEM_KB
#define C3L_SMEM_KB 200            // shared memory per SM the rings may take (the rest of the 228 KB stays L1)
#endif
#define C3L_RSLOT 6                // uint4 per lane and ring slot: H, E1, E2 x 2 halves
#define C3L_FLOOR (-30720)         // stored "unreachable"; every packed operation clamps here, 2048 above the int16 wrap
#define C3L_FLOOR2 0x88008800u
#define C3L_LOW_GUARD 64           // reachable cells must stay above C3L_FLOOR + C3L_LOW_GUARD
#ifndef C3L_THREADS
#define C3L_THREADS 64
#endif
#ifndef C3L_MINB
#define C3L_MINB 6                 // CTAs per SM the registers are bounded for (168): 12 warps per SM.  Measured on B200 per 100k
#endif                             // cfg2 reads: 4 -> 393 ms, 6 -> 374 ms, 8 (128 registers, spills, ring too small) -> 610 ms

// The serial per-thread phases (graph walks, backtrack, merge, consensus) loop under warp-uniform control:
// `while (C3L_ANY(cond)) { if (cond) { one step } }` -- every thread of the warp makes the same number of
// trips and the warp reconverges at each one, so 32 reads advance one step per pass instead of running one
// after the other.  All 32 threads must reach such a loop (no early return ahead of it).
#if defined(__CUDA_ARCH__)
#define C3L_ADDMAX(a, b, c) __viaddmax_s32((a), (b), (c))
#define C3L_MAX3(a, b, c) __vimax3_s32((a), (b), (c))
#define C3L_ANY(x) __any_sync(C3_FULL, (x))
#define C3L_LDCS4(p) __ldcs(p)           // DP rows are read once by the successor row: stream them (evict-first)
#define C3L_UMIN(a, b) min((unsigned)(a), (unsigned)(b))
#define C3L_PACK2(lo, hi) __byte_perm((unsigned)(lo), (unsigned)(hi), 0x5410)
#define C3L_VMAX2(a, b) __vmaxs2((a), (b))          // per-halfword signed max: VIMNMX.S16x2
#define C3L_VADDMAX2(a, b, c) __viaddmax_s16x2((a), (b), (c))   // per-halfword max(a + b, c): VIADDMNMX.S16x2
#define C3L_VMAX3_2(a, b, c) __vimax3_s16x2((a), (b), (c))      // VIMNMX3.S16x2
__device__ __forceinline__ unsigned c3l_prmt(unsigned a, unsigned b, unsigned sel)
{
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));   // generic mode: selector bit 3 = replicate the byte's sign
    return d;
}
#define C3L_PRMT(a, b, sel) c3l_prmt((a), (b), (sel))
#define C3L_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#ifndef C3L_NO_PF2
#define C3L_PREFETCH2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#else
#define C3L_PREFETCH2(p) do { (void)(p); } while (0)
#endif
#else
#define C3L_ANY(x) (x)
#define C3L_LDCS4(p) (*(p))
static inline unsigned c3l_humin(unsigned a, unsigned b) { return a < b ? a : b; }
#define C3L_UMIN(a, b) c3l_humin((unsigned)(a), (unsigned)(b))
#define C3L_PACK2(lo, hi) (((unsigned)(lo) & 0xffffu) | ((unsigned)(hi) << 16))
static inline unsigned c3l_hvmax2(unsigned a, unsigned b)
{
    const int al = (int16_t)(a & 0xffffu), ah = (int16_t)(a >> 16), bl = (int16_t)(b & 0xffffu), bh = (int16_t)(b >> 16);
    return ((unsigned)(al > bl ? al : bl) & 0xffffu) | ((unsigned)(ah > bh ? ah : bh) << 16);
}
#define C3L_VMAX2(a, b) c3l_hvmax2((a), (b))
static inline unsigned c3l_hvaddmax2(unsigned a, unsigned b, unsigned c)
{
    const int16_t sl = (int16_t)(uint16_t)((a & 0xffffu) + (b & 0xffffu)), sh = (int16_t)(uint16_t)((a >> 16) + (b >> 16));
    const int16_t cl = (int16_t)(c & 0xffffu), ch = (int16_t)(c >> 16);
    return ((unsigned)(uint16_t)(sl > cl ? sl : cl)) | ((unsigned)(uint16_t)(sh > ch ? sh : ch) << 16);
}
static inline unsigned c3l_hprmt(unsigned a, unsigned b, unsigned sel)
{
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned d = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned n = (sel >> (4 * i)) & 0xfu;
        unsigned byte = (unsigned)(v >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        d |= byte << (8 * i);
    }
    return d;
}
#define C3L_VADDMAX2(a, b, c) c3l_hvaddmax2((a), (b), (c))
#define C3L_VMAX3_2(a, b, c) c3l_hvmax2(c3l_hvmax2((a), (b)), (c))
#define C3L_PRMT(a, b, sel) c3l_hprmt((a), (b), (sel))
#define C3L_PREFETCH(p) do { (void)(p); } while (0)
#define C3L_PREFETCH2(p) do { (void)(p); } while (0)
static inline int c3l_hmax(int a, int b) { return a > b ? a : b; }
#define C3L_ADDMAX(a, b, c) c3l_hmax((a) + (b), (c))
#define C3L_MAX3(a, b, c) c3l_hmax(c3l_hmax((a), (b)), (c))
#endif

// -DC3L_PROF: per-phase cycle counters (lane 0 of every warp adds clock64() deltas); read with c3_debug_lane_prof
#if defined(C3L_PROF) && defined(__CUDACC__)
__device__ unsigned long long c3l_prof[24];
#endif
#if defined(C3L_PROF) && defined(__CUDA_ARCH__)
#define C3L_TICK(id) do { const long long t1_ = clock64(); if (lane == 0) atomicAdd(&c3l_prof[id], (unsigned long long)(t1_ - tprof_)); tprof_ = clock64(); } while (0)
#define C3L_TICK_INIT long long tprof_ = clock64()
#define C3L_COUNT(id, v) do { if (lane == 0) atomicAdd(&c3l_prof[id], (unsigned long long)(v)); } while (0)
// inside the row loop: `dep` makes the clock read wait for the values it names
#define C3L_TICK2_INIT long long tp2_ = clock64()
#define C3L_TICK2(id, dep) do { long long t1_; asm volatile("{ .reg .b32 t; mov.b32 t, %1; mov.u64 %0, %%clock64; }" : "=l"(t1_) : "r"((unsigned)(dep)) : "memory"); \
    if (lane == 0) atomicAdd(&c3l_prof[id], (unsigned long long)(t1_ - tp2_)); tp2_ = clock64(); } while (0)
#else
#define C3L_TICK2_INIT do { } while (0)
#define C3L_TICK2(id, dep) do { } while (0)
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
    c3_nrec nd, nd1, nd2, nd3; uint32_t hrv, hr1, hr2, hr3;   // look-ahead: records of v and of the next three nodes of the list
    uint2 pe, pe1, pe2;                               // first overflow in-edge of v, v+1, v+2 (in_n > 1)
    c3_prow ra, rb, ra1, rb1;                         // row records of the first two predecessors of v and of v+1 (unless among
                                                      // the two rows computed last when they were requested)
    c3_prow last2; int vlast2;                        // the row before the one just computed
    c3_prow last; int vlast, last_sm;                 // the row just computed; last_sm: it is in the shared-memory ring
    int beg, end, nvec, beg_sn, end_sn; c3_prow r0, r1;   // the row between setup and compute
};

// int16 index of column c (relative to the row's band start) of array a in the row stored at off4
C3_HD __forceinline__ int c3l_ci(int off4, int a, int c, int lane)
{
    return ((off4 + (c >> 4) * C3L_VSTRIDE + ((a << 1) + ((c >> 3) & 1)) * 32 + lane) << 3) + (c & 7);
}
// stored cells: H at the floor reads back as NEG_INF; E1 / E2 stay as stored (they only feed max() and inequalities)
C3_HD __forceinline__ int c3l_map(const int v) { return v <= C3L_FLOOR ? C3_NEG_INF : v; }
C3_HD __forceinline__ int c3l_ld_h(const int32_t *ar, const int idx16) { return c3l_map((int)reinterpret_cast<const int16_t *>(ar)[idx16]); }
C3_HD __forceinline__ int c3l_ld_e(const int32_t *ar, const int idx16) { return (int)reinterpret_cast<const int16_t *>(ar)[idx16]; }

// node / edge / row records by index, with the list terminator (C3_NONE) mapped to the sink so that the
// look-ahead past the end of the walk stays inside the workspace
C3_HD __forceinline__ int c3l_cl(const int v) { return v == (int)C3_NONE ? C3_SINK : v; }
C3_HD __forceinline__ uint2 c3l_ld_edge(const c3_pedge *p) { return *reinterpret_cast<const uint2 *>(p); }
#define C3L_E_ID(e) ((int)((e).x & 0xffffu))
#define C3L_E_NEXT(e) ((int)((e).y & 0xffffu))

// int32 index of the (F1, F2) pair entering vector vi of the row stored at off4
C3_HD __forceinline__ int c3l_fi(int off4, int vi, int lane) { return ((off4 + vi * C3L_VSTRIDE + 192) << 2) + 2 * lane; }

C3_HD __forceinline__ void c3l_unpack8(const uint4 w, int (&o)[8])
{
    o[0] = (int)(int16_t)(w.x & 0xffffu); o[1] = (int)w.x >> 16;
    o[2] = (int)(int16_t)(w.y & 0xffffu); o[3] = (int)w.y >> 16;
    o[4] = (int)(int16_t)(w.z & 0xffffu); o[5] = (int)w.z >> 16;
    o[6] = (int)(int16_t)(w.w & 0xffffu); o[7] = (int)w.w >> 16;
}
C3_HD __forceinline__ uint4 c3l_vmax8(const uint4 a, const uint4 b)
{
    return make_uint4(C3L_VMAX2(a.x, b.x), C3L_VMAX2(a.y, b.y), C3L_VMAX2(a.z, b.z), C3L_VMAX2(a.w, b.w));
}

// ---------------------------------------------------------------------------
// item start: first sequence -> linear graph
// ---------------------------------------------------------------------------
C3_HD __forceinline__ void c3l_item_begin(c3l_state &S, const c3_poa_args &A, const c3_poa_ws &W, const int item, const bool have)
{
    const int P_ms = A.P.match + A.P.mismatch;
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
            if (L <= 0 || L > 65000 || L + 2 > A.node_cap || P_ms > 255) S.err = C3L_E_RETRY;
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
            if (n.base >= 4) S.err = C3L_E_RETRY;           // N: scored 0 against everything -- the warp kernel's business
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
    // (the next node's record is requested before this node's hops are resolved: one load latency per node)
    int v = C3_NONE;
    if (act) { W.hr[C3_SINK] = C3_SINK; v = W.nodes[C3_SINK].prev; }
    c3_nrec cur = c3_ld_node(&W.nodes[c3l_cl(v)]);
    int cur_om = W.nodes[c3l_cl(v)].out_more;
    int vdone = C3_SINK; uint32_t hdone = C3_SINK;
    while (C3L_ANY(v != C3_NONE)) {
        if (v == C3_NONE) continue;
        const int pv = C3_N_PREV(cur);
        const c3_nrec nxt = c3_ld_node(&W.nodes[c3l_cl(pv)]);
        const int nxt_om = W.nodes[c3l_cl(pv)].out_more;
        int best_w = C3_N_W0(cur), best = C3_N_OUT0(cur);
        int e = cur_om;
        while (e != (int)C3_NONE) {
            const c3_pedge pe = W.pool[e];
            if ((int)pe.w > best_w) { best_w = pe.w; best = pe.id; }
            e = pe.next;
        }
        // (the heaviest successor is usually the node handled one trip earlier: its word is still in a register)
        const uint32_t hb = (best == vdone) ? hdone : W.hr[best];
        hdone = ((hb >> 16) + 1u) << 16; vdone = v;
        W.hr[v] = hdone;
        v = pv; cur = nxt; cur_om = nxt_om;
    }
    // column codes: qa[j] = base code of column j (= q[j-1]), 16 columns per store; j = 0 and the padding hold 4.
    // The row loop scores 4 columns per word against the node base on the fly (no per-base profile in memory: with
    // 256 threads per SM a 4-row profile does not stay in L1).  A sequence with an N is not handled here.
    {
        const int jmax = act ? ((qlen + 16) & ~15) : 0;   // the DP reads vectors up to column qlen | 15
        bool has_n = false;
        for (int j0 = 0; C3L_ANY(j0 < jmax); j0 += 16) {
            if (j0 >= jmax) continue;
            uint32_t wv[4] = {0x04040404u, 0x04040404u, 0x04040404u, 0x04040404u};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int j = j0 + k;
                if (j >= 1 && j <= qlen) {
                    const uint32_t qc = q[j - 1];
                    has_n |= qc >= 4;
                    wv[k >> 2] = (wv[k >> 2] & ~(0xffu << (8 * (k & 3)))) | (qc << (8 * (k & 3)));
                }
            }
            *reinterpret_cast<uint4 *>(W.qp + j0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
        if (has_n) { S.err = C3L_E_RETRY; act = false; }
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
    int16_t *ar16 = reinterpret_cast<int16_t *>(ar);
    for (int vi = 0; vi < S.nvec; ++vi)
        *reinterpret_cast<int2 *>(ar + c3l_fi(0, vi, lane)) = make_int2(C3_NEG_INF, C3_NEG_INF);
    for (int c = 0; c < 16 * S.nvec; ++c) {
        int h = C3L_FLOOR, x1 = C3L_FLOOR, x2 = C3L_FLOOR;
        if (b0 == 0 && c < wd) {
            if (c == 0) { h = 0; x1 = -oe1; x2 = -oe2; }
            else h = max(-(P.o1 + P.e1 * c), -(P.o2 + P.e2 * c));
            if (h < C3L_FLOOR + 1024) S.err = C3L_E_RETRY;
        }
        ar16[c3l_ci(0, 0, c, lane)] = (int16_t)h; ar16[c3l_ci(0, 1, c, lane)] = (int16_t)x1; ar16[c3l_ci(0, 2, c, lane)] = (int16_t)x2;
    }
    S.last_sm = 0;
    S.v = W.nodes[C3_SRC].next; S.rcount = 1;
    S.nd = c3_ld_node(&W.nodes[S.v]); S.hrv = W.hr[S.v];
    const int v1 = c3l_cl(C3_N_NEXT(S.nd));
    S.nd1 = c3_ld_node(&W.nodes[v1]); S.hr1 = W.hr[v1];
    const int v2 = c3l_cl(C3_N_NEXT(S.nd1));
    S.nd2 = c3_ld_node(&W.nodes[v2]); S.hr2 = W.hr[v2];
    const int v3 = c3l_cl(C3_N_NEXT(S.nd2));
    S.nd3 = c3_ld_node(&W.nodes[v3]); S.hr3 = W.hr[v3];
    S.pe = S.pe1 = S.pe2 = make_uint2(0u, 0u);
    if (C3_N_INN(S.nd) > 1) S.pe = c3l_ld_edge(&W.pool[C3_N_INMORE(S.nd)]);
    if (C3_N_INN(S.nd1) > 1) S.pe1 = c3l_ld_edge(&W.pool[C3_N_INMORE(S.nd1)]);
    if (C3_N_INN(S.nd2) > 1) S.pe2 = c3l_ld_edge(&W.pool[C3_N_INMORE(S.nd2)]);
    // the first two nodes after the source can only have the source / the first node as predecessors: S.last, S.last2
    S.ra = S.rb = S.ra1 = S.rb1 = ri; S.last2 = ri; S.vlast2 = -1;
    S.nvec = 0;
}

// ---------------------------------------------------------------------------
// row setup: adaptive band of the current node's row.  Returns its number of vectors (0: no row).
// The records of the first two predecessors were requested two rows ago (S.ra, S.rb); a predecessor that is one of
// the two rows computed since comes from S.last / S.last2.
// ---------------------------------------------------------------------------
C3_HD __forceinline__ int c3l_row_setup(c3l_state &S, const c3_poa_args &A, const c3_poa_ws &W)
{
    S.nvec = 0;
    if (!S.aligning || S.err || S.v == C3_SINK) return 0;
    const int qlen = S.qlen, w = S.w;
    const int rem = (int)(S.hrv >> 16) - 1;
    const int rr = qlen - rem;
    const int npre = C3_N_INN(S.nd);
    if (npre > C3_MAXPRE || C3_N_BASE(S.nd) >= 4) { S.err = C3L_E_RETRY; return 0; }   // c3_poa_kernel's limit / an N node: let it handle
    const c3_prow r0 = (C3_N_IN0(S.nd) == S.vlast) ? S.last : (C3_N_IN0(S.nd) == S.vlast2) ? S.last2 : S.ra;
    int mpl = min(S.n, (int)r0.mp), mpr = r0.mp, min_pre_beg = r0.beg;
    c3_prow r1; r1.off = 0; r1.beg = 16; r1.end = 0; r1.mp = 0; r1.in0 = C3_NONE; r1.link = 0; r1.base = 4; r1.npre = 0;   // empty band
    if (npre > 1) {
        r1 = (C3L_E_ID(S.pe) == S.vlast) ? S.last : (C3L_E_ID(S.pe) == S.vlast2) ? S.last2 : S.rb;
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

// 8 columns (from j0, a multiple of 8) of one predecessor's packed H, E1, E2 -- from the arena, or from the
// shared-memory ring when the predecessor is the row just computed; outside its band: all C3L_FLOOR
C3_HD __forceinline__ void c3l_load8(const uint4 *ar4, const uint4 *sm, const bool from_sm, const int slot,
                                     const c3_prow &rp, const int j0, const int lane, uint4 &a, uint4 &b, uint4 &c)
{
    if (j0 >= (int)rp.beg && j0 <= (int)rp.end) {
        if (from_sm) {
            const uint4 *src = sm + (slot * C3L_RSLOT + ((j0 >> 3) & 1)) * 32 + lane;
            a = src[0]; b = src[64]; c = src[128];
        } else {
            const int cc = j0 - rp.beg;
            const uint4 *src = ar4 + rp.off + (cc >> 4) * C3L_VSTRIDE + ((cc >> 3) & 1) * 32 + lane;
            a = C3L_LDCS4(src); b = C3L_LDCS4(src + 64); c = C3L_LDCS4(src + 128);
        }
    } else {
        a = b = c = make_uint4(C3L_FLOOR2, C3L_FLOOR2, C3L_FLOOR2, C3L_FLOOR2);
    }
}

// Folds one predecessor row into the ring over the band [beg, beg + 8 nstep): mode 1 = the ring becomes that row
// (floor outside its band), mode 2 = the row is in the ring already, only the steps outside its band are set to the
// floor, mode 0 = per-halfword max with what the ring holds.  4 steps of arena loads are in flight at a time.
C3_HD __forceinline__ void c3l_fold_pred(const uint4 *ar4, uint4 *sm, const int smR, const int slot0, const c3_prow &rp,
                                         const int beg, const int nstep, const int lane, const int mode)
{
    for (int h0 = 0; h0 < nstep; h0 += 4) {
        uint4 a[4], b[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j0 = beg + 8 * (h0 + u);
            const bool inb = h0 + u < nstep && j0 >= (int)rp.beg && j0 <= (int)rp.end;
            a[u] = b[u] = c[u] = make_uint4(C3L_FLOOR2, C3L_FLOOR2, C3L_FLOOR2, C3L_FLOOR2);
            if (inb && mode != 2) {
                const int cc = j0 - rp.beg;
                const uint4 *src = ar4 + rp.off + (cc >> 4) * C3L_VSTRIDE + ((cc >> 3) & 1) * 32 + lane;
                a[u] = C3L_LDCS4(src); b[u] = C3L_LDCS4(src + 64); c[u] = C3L_LDCS4(src + 128);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int h = h0 + u, j0 = beg + 8 * h;
            if (h >= nstep) continue;
            const bool inb = j0 >= (int)rp.beg && j0 <= (int)rp.end;
            int sl = slot0 + (h >> 1); if (sl >= smR) sl -= smR;
            uint4 *d2 = sm + (sl * C3L_RSLOT + (h & 1)) * 32 + lane;
            if (mode == 0) {
                if (inb) { d2[0] = c3l_vmax8(d2[0], a[u]); d2[64] = c3l_vmax8(d2[64], b[u]); d2[128] = c3l_vmax8(d2[128], c[u]); }
            } else if (mode == 1 || !inb) {
                d2[0] = a[u]; d2[64] = b[u]; d2[128] = c[u];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// row compute: all columns of the row set up by c3l_row_setup in steps of 8, row record, advance to the
// next node.  While the cells are computed, everything the NEXT row's setup reads is requested: the row
// records of its first two predecessors, the node record two rows ahead and that node's first overflow edge.
// sm: this warp's shared-memory ring (smR vector slots per lane; smR = 0: none).
// ---------------------------------------------------------------------------
C3_HD __forceinline__ void c3l_row_compute(c3l_state &S, const c3_poa_args &A, const c3_poa_para_dev &P, const c3_poa_ws &W,
                                           int32_t *ar, const int base4, const int lane, uint4 *sm, const int smR)
{
    if (S.nvec <= 0) return;
    const int e1 = P.e1, e2 = P.e2, oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const int v = S.v;
    // ---- look-ahead for the next row (v1) and the one after ----
    const int v1 = c3l_cl(C3_N_NEXT(S.nd)), v4 = c3l_cl(C3_N_NEXT(S.nd3));
    const c3_nrec nd4 = c3_ld_node(&W.nodes[v4]);
    const uint32_t hr4 = W.hr[v4];
    uint2 pe3 = make_uint2(0u, 0u);
    if (C3_N_INN(S.nd3) > 1) pe3 = c3l_ld_edge(&W.pool[C3_N_INMORE(S.nd3)]);
    c3_prow ra2 = S.last, rb2 = S.last;                 // predecessors of v+2 (rows v and v+1 do not exist yet: last / last2 then)
    {
        const int p0 = c3l_cl(C3_N_IN0(S.nd2)), p1 = C3L_E_ID(S.pe2);
        if (p0 != v && p0 != v1) ra2 = W.rows[p0];
        if (C3_N_INN(S.nd2) > 1 && p1 != v && p1 != v1) rb2 = W.rows[p1];
    }
    // cells of the NEXT row's predecessors that will come from the arena: start their lines towards L2 a whole row
    // ahead, over the columns of this row's band (the next band is about the same)
    {
        const int p0 = c3l_cl(C3_N_IN0(S.nd1)), p1 = C3L_E_ID(S.pe1);
        const int lo = S.beg, hi = S.end + 16;
        if (p0 != v) {
            const c3_prow rq = (p0 == S.vlast) ? S.last : S.ra1;
            const uint4 *a4 = reinterpret_cast<const uint4 *>(ar);
            for (int j0 = max(lo, (int)rq.beg); j0 <= min(hi, (int)rq.end); j0 += 8) {
                const int cc = j0 - rq.beg;
                const uint4 *src = a4 + rq.off + (cc >> 4) * C3L_VSTRIDE + ((cc >> 3) & 1) * 32 + lane;
                C3L_PREFETCH2(src); C3L_PREFETCH2(src + 64); C3L_PREFETCH2(src + 128);
            }
        }
        if (C3_N_INN(S.nd1) > 1 && p1 != v) {
            const c3_prow rq = (p1 == S.vlast) ? S.last : S.rb1;
            const uint4 *a4 = reinterpret_cast<const uint4 *>(ar);
            for (int j0 = max(lo, (int)rq.beg); j0 <= min(hi, (int)rq.end); j0 += 8) {
                const int cc = j0 - rq.beg;
                const uint4 *src = a4 + rq.off + (cc >> 4) * C3L_VSTRIDE + ((cc >> 3) & 1) * 32 + lane;
                C3L_PREFETCH2(src); C3L_PREFETCH2(src + 64); C3L_PREFETCH2(src + 128);
            }
        }
    }
    const int npre = C3_N_INN(S.nd), nbase = C3_N_BASE(S.nd);
    const int beg = S.beg, end = S.end, nvec = S.nvec;
    const c3_prow r0 = S.r0, r1 = S.r1;
    C3L_TICK2_INIT;
    C3L_TICK2(16, r0.off + r1.off + beg);
    // the first predecessor comes from the ring when it is the row just computed and this band does not start
    // left of it (then no slot is overwritten before it is read); this row goes into the ring if it fits
    const bool p0_sm = S.last_sm && C3_N_IN0(S.nd) == S.vlast && beg >= (int)r0.beg;
    const bool keep_sm = nvec <= smR;
    int slot = smR > 0 ? S.beg_sn % smR : 0;
    const uint4 *ar4 = reinterpret_cast<const uint4 *>(ar);
    const int nstep = nvec * 2;
    // Anything but "one predecessor, and it is in the ring" is first FOLDED into the ring: a per-halfword max over all
    // predecessor rows on the packed cells, with one walk of the edge list and the loads of 4 steps in flight
    // at a time.  The row loop below then reads nothing but the ring (no arena load sits on its critical path).
    // A row wider than the ring takes the slow path: cells of the first two predecessors straight from the arena,
    // further ones along the edge chain per step.
    bool p0_ring = p0_sm;
    c3_prow r0e = r0, r1e = r1;
    // (Measured and dropped: folding the second predecessor as well -- 392 vs 374 ms per 100k reads, the fold's loads are
    // exposed once per row while the in-loop register prefetch overlaps them with the arithmetic; the column codes in
    // the ring -- 398 ms, the larger ring costs L1.)
    const bool fold = keep_sm && npre > 2;              // the second predecessor stays in the row loop (register prefetch)
    if (fold) {
        const bool have0 = p0_sm && beg >= (int)r0.beg;
        if (!have0) c3l_fold_pred(ar4, sm, smR, slot, r0, beg, nstep, lane, 1);
        else if (end > (int)r0.end) c3l_fold_pred(ar4, sm, smR, slot, r0, beg, nstep, lane, 2);
        if (npre > 1) {
            int e = C3L_E_NEXT(S.pe);
            for (int k = 2; k < npre; ++k) {
                const c3_pedge pe = W.pool[e]; e = pe.next;
                const c3_prow rp = W.rows[pe.id];
                c3l_fold_pred(ar4, sm, smR, slot, rp, beg, nstep, lane, 0);
            }
        }
        p0_ring = true;
        r0e.beg = (uint16_t)beg; r0e.end = (uint16_t)(beg + 16 * nvec - 1);
    }
    C3L_TICK2(17, sm[lane].x);
    uint4 *out4 = reinterpret_cast<uint4 *>(ar) + base4 + lane;
    const int8_t *qprow = W.qp;                         // column codes
    // The row loop works on the packed cells, two columns per 32-bit word (VIADDMNMX.S16x2 / VIMNMX3.S16x2); only the
    // horizontal gap F, a strictly sequential recurrence, and the row arg-max run per column in int32.
    const uint32_t nb4 = (uint32_t)nbase * 0x01010101u, m4 = (uint32_t)P.match * 0x01010101u;
    const uint32_t dms = (uint32_t)(256 - (P.match + P.mismatch));      // byte: match  ->  match - (match + mismatch)
    const uint32_t ne1 = C3L_PACK2(-e1, -e1), ne2 = C3L_PACK2(-e2, -e2), noe1 = C3L_PACK2(-oe1, -oe1), noe2 = C3L_PACK2(-oe2, -oe2);
    int g1 = C3_NEG_INF + oe1, g2 = C3_NEG_INF + oe2;   // F1 + (o1 + e1), F2 + (o2 + e2): one add-max per column and gap type
    uint32_t carry = C3L_FLOOR2;                        // high half: merged predecessor H at column j0 - 1
    int bestkey = -0x7fffffff - 1, h_first = 0;
    // software pipeline: both predecessors' cells and the column codes of step h+1 are requested before step h
    uint4 na, nb, nc, pa, pb, pc;
    uint2 nsw = make_uint2(0u, 0u);
    c3l_load8(ar4, sm, p0_ring, slot, r0e, beg, lane, na, nb, nc);
    c3l_load8(ar4, sm, false, 0, r1e, beg, lane, pa, pb, pc);
    nsw = *reinterpret_cast<const uint2 *>(qprow + beg);
    C3L_TICK2(18, nsw.x + na.x + pa.x);
    // (#pragma unroll 2 here: measured 422 vs 371 ms -- spills at 168 registers)
    for (int h = 0; h < nstep; ++h) {
        const int j0 = beg + 8 * h;
        uint32_t hv[4] = {C3L_VMAX2(na.x, pa.x), C3L_VMAX2(na.y, pa.y), C3L_VMAX2(na.z, pa.z), C3L_VMAX2(na.w, pa.w)};
        uint32_t x1[4] = {C3L_VMAX2(nb.x, pb.x), C3L_VMAX2(nb.y, pb.y), C3L_VMAX2(nb.z, pb.z), C3L_VMAX2(nb.w, pb.w)};
        uint32_t x2[4] = {C3L_VMAX2(nc.x, pc.x), C3L_VMAX2(nc.y, pc.y), C3L_VMAX2(nc.z, pc.z), C3L_VMAX2(nc.w, pc.w)};
        const uint2 sw = nsw;
        C3L_TICK2(13, hv[0] + x1[1] + x2[2] + sw.x);     // waited for the loads of this step
        const int wslot = slot;
        if (h & 1) { ++slot; if (slot >= smR) slot = 0; }
        if (h + 1 < nstep) {
            c3l_load8(ar4, sm, p0_ring, slot, r0e, j0 + 8, lane, na, nb, nc);
            c3l_load8(ar4, sm, false, 0, r1e, j0 + 8, lane, pa, pb, pc);
            nsw = *reinterpret_cast<const uint2 *>(qprow + j0 + 8);
        }
        if (npre > 2 && !fold) {                         // wider than the ring: follow the edge chain per step
            int e = C3L_E_NEXT(S.pe);
            for (int k = 2; k < npre; ++k) {
                const c3_pedge pe = W.pool[e]; e = pe.next;
                const c3_prow rp = W.rows[pe.id];
                uint4 qa, qb, qc;
                c3l_load8(ar4, sm, false, 0, rp, j0, lane, qa, qb, qc);
                const int jc = j0 - 1;
                if (h > 0 && jc >= (int)rp.beg && jc <= (int)rp.end) {
                    const int prev = c3l_ld_e(ar, c3l_ci(rp.off, 0, jc - rp.beg, lane));
                    carry = C3L_VMAX2(carry, C3L_PACK2(C3L_FLOOR, prev));
                }
                hv[0] = C3L_VMAX2(hv[0], qa.x); hv[1] = C3L_VMAX2(hv[1], qa.y); hv[2] = C3L_VMAX2(hv[2], qa.z); hv[3] = C3L_VMAX2(hv[3], qa.w);
                x1[0] = C3L_VMAX2(x1[0], qb.x); x1[1] = C3L_VMAX2(x1[1], qb.y); x1[2] = C3L_VMAX2(x1[2], qb.z); x1[3] = C3L_VMAX2(x1[3], qb.w);
                x2[0] = C3L_VMAX2(x2[0], qc.x); x2[1] = C3L_VMAX2(x2[1], qc.y); x2[2] = C3L_VMAX2(x2[2], qc.z); x2[3] = C3L_VMAX2(x2[3], qc.w);
            }
        }
        const int vi = h >> 1;
        if (!(h & 1))                                   // F entering this 16-column vector: the backtrack restarts from it
            *reinterpret_cast<int2 *>(ar + c3l_fi(base4, vi, lane)) = make_int2(g1 - oe1, g2 - oe2);
        // M: merged predecessor H one column to the left
        uint32_t mw[4];
        mw[0] = C3L_PRMT(carry, hv[0], 0x5432u);
#pragma unroll
        for (int t = 1; t < 4; ++t) mw[t] = C3L_PRMT(hv[t - 1], hv[t], 0x5432u);
        carry = hv[3];
        // scores: 4 columns per word, byte = match where the column's base equals the node's, else -mismatch;
        // sign-extended to halfwords
        uint32_t sb[2] = {sw.x ^ nb4, sw.y ^ nb4};
#pragma unroll
        for (int t = 0; t < 2; ++t) sb[t] = m4 + (((sb[t] + 0x7f7f7f7fu) >> 7) & 0x01010101u) * dms;
        const uint32_t sc[4] = {C3L_PRMT(sb[0], 0u, 0x9180u), C3L_PRMT(sb[0], 0u, 0xb3a2u), C3L_PRMT(sb[1], 0u, 0x9180u), C3L_PRMT(sb[1], 0u, 0xb3a2u)};
        uint32_t hme[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) hme[t] = C3L_VMAX3_2(C3L_VADDMAX2(mw[t], sc[t], C3L_FLOOR2), x1[t], x2[t]);
        const int lim = end - j0;                       // last active column of this step (>= 7: all)
        if (lim < 7) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (2 * t > lim) hme[t] = C3L_FLOOR2;
                else if (2 * t + 1 > lim) hme[t] = (hme[t] & 0xffffu) | (C3L_FLOOR2 & 0xffff0000u);
            }
        }
        int hk[8], fm[8], hhk[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) { hk[2 * t] = (int)(int16_t)(hme[t] & 0xffffu); hk[2 * t + 1] = (int)hme[t] >> 16; }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            fm[k] = C3L_MAX3(g1 - oe1, g2 - oe2, C3L_FLOOR);
            hhk[k] = max(hk[k], fm[k]);
            g1 = C3L_ADDMAX(g1, -e1, hk[k]);
            g2 = C3L_ADDMAX(g2, -e2, hk[k]);
        }
        if (h == 0) h_first = hhk[0];
        uint32_t hh2[4], n1[4], n2[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            hh2[t] = C3L_VMAX2(hme[t], C3L_PACK2(fm[2 * t], fm[2 * t + 1]));
            n1[t] = C3L_VADDMAX2(x1[t], ne1, C3L_VADDMAX2(hh2[t], noe1, C3L_FLOOR2));
            n2[t] = C3L_VADDMAX2(x2[t], ne2, C3L_VADDMAX2(hh2[t], noe2, C3L_FLOOR2));
        }
        if (lim < 7) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (2 * t > lim) { hh2[t] = n1[t] = n2[t] = C3L_FLOOR2; hhk[2 * t] = hhk[2 * t + 1] = C3L_FLOOR; }
                else if (2 * t + 1 > lim) {
                    hh2[t] = (hh2[t] & 0xffffu) | (C3L_FLOOR2 & 0xffff0000u);
                    n1[t] = (n1[t] & 0xffffu) | (C3L_FLOOR2 & 0xffff0000u);
                    n2[t] = (n2[t] & 0xffffu) | (C3L_FLOOR2 & 0xffff0000u);
                    hhk[2 * t + 1] = C3L_FLOOR;
                }
            }
        }
        const uint4 oa = make_uint4(hh2[0], hh2[1], hh2[2], hh2[3]), ob = make_uint4(n1[0], n1[1], n1[2], n1[3]),
                    oc = make_uint4(n2[0], n2[1], n2[2], n2[3]);
        C3L_TICK2(14, oa.x + ob.y + oc.z);               // arithmetic
        uint4 *dst = out4 + vi * C3L_VSTRIDE + (h & 1) * 32;
        dst[0] = oa; dst[64] = ob; dst[128] = oc;
        if (keep_sm) {
            uint4 *d2 = sm + (wslot * C3L_RSLOT + (h & 1)) * 32 + lane;
            d2[0] = oa; d2[64] = ob; d2[128] = oc;
        }
        // simd_abpoa_ada_max_i as one packed max: value in the high half, tie-break priority in the low half
        // (lowest SIMD lane, then the last vector, then the earliest vector)
        const int vp = (vi == nvec - 1) ? 0xfff : (0xffe - vi);
        const int k16 = (h & 1) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            bestkey = C3L_ADDMAX(hhk[k] * 65536 + vp, (15 - k16 - k) << 12, bestkey);
        C3L_TICK2(15, bestkey);                          // stores issued + arg-max
    }
    // Exactness of the int16 form: a cell can lose at most D per column to its left neighbour (gap open / extend),
    // so a first band cell comfortably above the floor means that every cell of the row is reachable and was never
    // clamped.  Anything else leaves the read to the warp kernel.
    {
        const int D = max(min(oe1, oe2), max(e1, e2));
        if (h_first < C3L_FLOOR + C3L_LOW_GUARD + oe2 + D * (end - beg + 1)) S.err = C3L_E_RETRY;
    }
    int best_i = -1;
    if ((bestkey >> 16) > C3L_FLOOR) {
        const int sl = 15 - ((bestkey >> 12) & 15);
        const int vp = bestkey & 0xfff;
        const int sn = (vp == 0xfff) ? S.end_sn : S.beg_sn + (0xffe - vp);
        best_i = (sn << 4) + sl;
    }
    c3_prow ri; ri.off = base4; ri.beg = (uint16_t)beg; ri.end = (uint16_t)end; ri.mp = (uint16_t)(best_i + 1);
    ri.in0 = (uint16_t)C3_N_IN0(S.nd); ri.base = (uint8_t)nbase; ri.npre = (uint8_t)npre;
    ri.link = (uint16_t)S.rcount;
    W.rows[v] = ri;
    S.last2 = S.last; S.vlast2 = S.vlast;
    S.last = ri; S.vlast = v; S.last_sm = keep_sm ? 1 : 0;
    ri.link = (uint16_t)v; ri.mp = r0.link;
    ri.in0 = npre > 1 ? r1.link : (uint16_t)C3_NONE;   // by position: in0 = position of the SECOND predecessor's row
    W.ord[S.rcount] = ri;
    S.cells_total += end - beg + 1;
    ++S.rcount;
    S.v = v1; S.nd = S.nd1; S.hrv = S.hr1; S.nd1 = S.nd2; S.hr1 = S.hr2; S.nd2 = S.nd3; S.hr2 = S.hr3; S.nd3 = nd4; S.hr3 = hr4;
    S.pe = S.pe1; S.pe1 = S.pe2; S.pe2 = pe3; S.ra = S.ra1; S.rb = S.rb1; S.ra1 = ra2; S.rb1 = rb2;
    S.nvec = 0;
    C3L_TICK2(19, S.nd3.a.x + S.hr3 + S.pe2.x + S.ra1.off + S.rb1.off);
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
            const int val = c3l_ld_h(ar, c3l_ci(rp.off, 0, en - rp.beg, lane));
            if (val > best_score) { best_score = val; bj = en; bk = rp.link; }
        }
        if (bk < 0 || qlen - bj + 8 > A.cigar_cap) { S.err = C3L_E_RETRY; run = false; }
        else {
            j = bj; rt = W.ord[bk];
            for (int t = qlen; t > bj; --t)
                cg[qlen - t] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(t - 1) << 32);
            nc = qlen - bj;
            hij = c3l_ld_h(ar, c3l_ci(rt.off, 0, j - rt.beg, lane));
        }
    }
    int cur_op = C3_OP_ALL;
    // the record of the current row's first predecessor is requested as soon as the row is known (one step ahead)
    c3_prow pr0 = W.ord[rt.mp == C3_NONE ? 0 : rt.mp];
    c3_prow pr1 = W.ord[rt.in0 == C3_NONE ? 0 : rt.in0];        // second predecessor likewise (rows with two or more)
    while (C3L_ANY(run && rt.link != C3_SRC && j > 0)) {
        C3L_COUNT(11, 1);
        if (!(run && rt.link != C3_SRC && j > 0)) continue;
        const int row_before = rt.link;
        const int i = rt.link;
        const int b = rt.beg;
        int hit = 0;
        unsigned long long opw = 0;
        if (j >= b && j <= (int)rt.end) {
            const int s = c3_score(P, rt.base, q[j - 1]);
            const int npre = rt.npre;
            // third and further predecessors follow the edge list (rare); the first two come from pr0 / pr1
            if (cur_op & C3_OP_M) {
                const bool ok0 = !(j - 1 < max((int)pr0.beg, b) || j - 1 > (int)pr0.end);
                const bool ok1 = npre > 1 && !(j - 1 < max((int)pr1.beg, b) || j - 1 > (int)pr1.end);
                int ph0 = 0, ph1 = 0;                  // both cells requested before either is compared
                if (ok0) ph0 = c3l_ld_h(ar, c3l_ci(pr0.off, 0, j - 1 - pr0.beg, lane));
                if (ok1) ph1 = c3l_ld_h(ar, c3l_ci(pr1.off, 0, j - 1 - pr1.beg, lane));
                if (ok0 && ph0 + s == hij) {
                    opw = C3_CG_MATCH | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                    rt = pr0; --j; hit = 1; cur_op = C3_OP_ALL; hij = ph0;
                } else if (ok1 && ph1 + s == hij) {
                    opw = C3_CG_MATCH | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                    rt = pr1; --j; hit = 1; cur_op = C3_OP_ALL; hij = ph1;
                } else if (npre > 2) {
                    int e = W.pool[W.nodes[i].in_more].next;
                    for (int k = 2; k < npre; ++k) {
                        const c3_pedge pe = W.pool[e]; e = pe.next;
                        const c3_prow pr = W.ord[W.rows[pe.id].link];
                        if (j - 1 < max((int)pr.beg, b) || j - 1 > (int)pr.end) continue;
                        const int ph = c3l_ld_h(ar, c3l_ci(pr.off, 0, j - 1 - pr.beg, lane));
                        if (ph + s == hij) {
                            opw = C3_CG_MATCH | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                            rt = pr; --j; hit = 1; cur_op = C3_OP_ALL; hij = ph;
                            break;
                        }
                    }
                }
            }
            if (!hit && (cur_op & (C3_OP_E | C3_OP_F))) {
                // No match / mismatch move from here.  Everything the gap tests of this cell can need -- the first two
                // predecessors' H, E1, E2 at column j, this row's E1, E2, and the vector of this row that F is rebuilt
                // from -- is requested at once, so the tests cost one memory round trip, not one per question.
                const bool want_e = (cur_op & C3_OP_E) != 0, want_f = (cur_op & C3_OP_F) && j - 1 >= b;
                const bool oke0 = want_e && j >= (int)pr0.beg && j <= (int)pr0.end;
                const bool oke1 = want_e && npre > 1 && j >= (int)pr1.beg && j <= (int)pr1.end;
                int eh0 = 0, e10 = 0, e20 = 0, eh1 = 0, e11 = 0, e21 = 0, ce1 = 0, ce2 = 0;
                int2 fin = make_int2(0, 0);
                uint4 hq0 = make_uint4(0u, 0u, 0u, 0u), hq1 = hq0;
                const int cm = j - 1 - b, vs = cm >> 4, tm = cm & 15;
                if (oke0) {
                    const int pc = j - pr0.beg;
                    eh0 = c3l_ld_h(ar, c3l_ci(pr0.off, 0, pc, lane)); e10 = c3l_ld_e(ar, c3l_ci(pr0.off, 1, pc, lane)); e20 = c3l_ld_e(ar, c3l_ci(pr0.off, 2, pc, lane));
                }
                if (oke1) {
                    const int pc = j - pr1.beg;
                    eh1 = c3l_ld_h(ar, c3l_ci(pr1.off, 0, pc, lane)); e11 = c3l_ld_e(ar, c3l_ci(pr1.off, 1, pc, lane)); e21 = c3l_ld_e(ar, c3l_ci(pr1.off, 2, pc, lane));
                }
                if (want_e && !(cur_op & C3_OP_M)) { ce1 = c3l_ld_e(ar, c3l_ci(rt.off, 1, j - b, lane)); ce2 = c3l_ld_e(ar, c3l_ci(rt.off, 2, j - b, lane)); }
                if (want_f) {
                    fin = *reinterpret_cast<const int2 *>(ar + c3l_fi(rt.off, vs, lane));
                    const uint4 *hp = reinterpret_cast<const uint4 *>(ar) + rt.off + vs * C3L_VSTRIDE + lane;
                    hq0 = hp[0]; hq1 = hp[32];
                }
                if (want_e) {
                    int e = C3_NONE;
                    for (int k = 0; k < npre; ++k) {
                        c3_prow pr = k == 0 ? pr0 : pr1;
                        int ph = k == 0 ? eh0 : eh1, pe1 = k == 0 ? e10 : e11, pe2 = k == 0 ? e20 : e21;
                        bool inb = k == 0 ? oke0 : oke1;
                        if (k >= 2) {
                            if (k == 2) e = W.pool[W.nodes[i].in_more].next;
                            const c3_pedge pe = W.pool[e]; pr = W.ord[W.rows[pe.id].link]; e = pe.next;
                            inb = j >= (int)pr.beg && j <= (int)pr.end;
                            if (inb) {
                                const int pc = j - pr.beg;
                                ph = c3l_ld_h(ar, c3l_ci(pr.off, 0, pc, lane)); pe1 = c3l_ld_e(ar, c3l_ci(pr.off, 1, pc, lane)); pe2 = c3l_ld_e(ar, c3l_ci(pr.off, 2, pc, lane));
                            }
                        }
                        if (!inb) continue;
                        if (cur_op & C3_OP_E1) {
                            if (cur_op & C3_OP_M) {
                                if (hij == pe1) { cur_op = (ph - oe1 == pe1) ? (C3_OP_M | C3_OP_F) : C3_OP_E1; hit = 1; }
                            } else if (ce1 == pe1 - e1) {
                                cur_op = (ph - oe1 == pe1) ? (C3_OP_M | C3_OP_F) : C3_OP_E1; hit = 1;
                            }
                        }
                        if (!hit && (cur_op & C3_OP_E2)) {
                            if (cur_op & C3_OP_M) {
                                if (hij == pe2) { cur_op = (ph - oe2 == pe2) ? (C3_OP_M | C3_OP_F) : C3_OP_E2; hit = 1; }
                            } else if (ce2 == pe2 - e2) {
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
                // (cur_op may just have been rewritten by an E hit; the F test only runs without one)
                if (!hit && want_f && (cur_op & C3_OP_F)) {
                    // F is stored only where it enters a 16-column vector: rebuild F[j-1] and F[j] from there
                    int hq[16];
                    {
                        int t0[8];
                        c3l_unpack8(hq0, t0);
#pragma unroll
                        for (int k = 0; k < 8; ++k) hq[k] = c3l_map(t0[k]);
                        c3l_unpack8(hq1, t0);
#pragma unroll
                        for (int k = 0; k < 8; ++k) hq[8 + k] = c3l_map(t0[k]);
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
        }
        if (!hit) { S.err = C3L_E_RETRY; run = false; continue; }
        if ((int)rt.link != row_before) {
            pr0 = W.ord[rt.mp == C3_NONE ? 0 : rt.mp];
            pr1 = W.ord[rt.in0 == C3_NONE ? 0 : rt.in0];
        }
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
    unsigned long long opn = nc > 0 ? cg[nc - 1] : 0ull;
    for (int t = nc - 1; C3L_ANY(t >= 0 && !g.err); --t) {
        if (!(t >= 0 && !g.err)) continue;
        const unsigned long long opc = opn;
        if (t > 0) {                                       // next op now, and its node record on the way into L1
            opn = cg[t - 1];
            const int nn = (int)((opn >> 8) & 0xffff);
            if (nn != (int)C3_NONE) C3L_PREFETCH(&g.nodes[nn]);
        }
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
    int sm_vec;                    // shared-memory ring: vector slots per lane (dynamic shared memory = warps x sm_vec x 3.5 KB)
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
    extern __shared__ uint4 c3l_smem[];
    uint4 *sm = c3l_smem + (size_t)(threadIdx.x >> 5) * L.sm_vec * (C3L_RSLOT * 32);
    const int smR = L.sm_vec;
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
                c3l_row_compute(S, A, P, W, ar, used4, lane, sm, smR);
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
