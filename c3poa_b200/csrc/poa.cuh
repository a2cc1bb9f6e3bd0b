// poa.cuh -- stage 3b: adaptive-banded partial-order alignment + heaviest-bundling
// consensus with abPOA 1.0.5 semantics (convex gap, global mode, extra_b/extra_f band).
//
// Replaces poa.msa_aligner(match=5).msa(subreads, out_cons=True, out_msa=True)
//   (/root/reference/bin/determine_consensus.py:30-47).
//
// One warp per read (persistent grid, atomic work counter, largest reads first).  Per added subread:
//   prepare   (lanes over nodes)  heaviest successor per node, then abPOA's "remaining path length"
//                                 by in-place pointer jumping on packed (hops, next) words
//   profile   (lanes over cols)   int8 substitution scores of the subread against A/C/G/T
//   DP        (lanes over cols)   rows = graph nodes in a maintained topological list order; each lane owns
//                                 4 consecutive band columns (128 per pass); predecessor rows come as 128-bit
//                                 loads from a shared-memory ring of recent rows (else HBM); the horizontal
//                                 (F) gap is an in-lane recurrence + one warp prefix-max of lane aggregates;
//                                 row arg-max by two REDUX; band anchors are pulled from predecessor records
//   backtrack (all lanes)         c3_bt_merge: the first-predecessor chain inside a 32-row window is resolved
//                                 by pointer doubling and up to 31 match/mismatch moves are verified at once;
//                                 other moves take the generic step (abPOA's M -> E1 -> E2 -> F1 -> F2 order and
//                                 op-mask state machine); F is rebuilt along one row when an insertion is traced
//   merge     (all lanes)         32 cigar ops at a time: weight bumps of existing edges in parallel, new nodes
//                                 and edges through lane 0 in order; new nodes are spliced into the list next to
//                                 their aligned group (a valid topological order without re-sorting; every
//                                 quantity the DP derives is order-independent)
// then either the two MSA rows (2-sequence groups, c3_emit_msa) or heaviest bundling and the
// consensus walk (c3_consensus, single thread).
//
// The graph, DP rows, row records and cigar live in a per-warp HBM workspace (L1/L2 cached).
#pragma once
#include "common.cuh"

#ifndef C3_POA_THREADS
#define C3_POA_THREADS 128
#endif
#ifndef C3_POA_MINB
#define C3_POA_MINB 5       // resident CTAs per SM the register allocation is bounded for (<= 102 regs; measured best)
#endif
#define C3_HD __host__ __device__
#define C3_NONE 0xffffu
#define C3_SRC 0
#define C3_SINK 1
#define C3_NEG_INF (-(1 << 29))
#define C3_NEG_HALF (-(1 << 28))
#define C3_MAXPRE 48
#ifndef C3_RING
#define C3_RING 2            // recent DP rows kept in shared memory per warp (rows of <= 128 columns); 1, 2 or 4.
                             // Measured: 1, 2 and 4 slots perform alike; 2 leaves more of the SM's 228 KB to L1.
#endif

#define C3_OP_M 0x1
#define C3_OP_E1 0x2
#define C3_OP_E2 0x4
#define C3_OP_E 0x6
#define C3_OP_F1 0x8
#define C3_OP_F2 0x10
#define C3_OP_F 0x18
#define C3_OP_ALL 0x1f

#define C3_CG_MATCH 0ull
#define C3_CG_INS 1ull
#define C3_CG_DEL 2ull

// error codes written to c3_read_result.status
#define C3_E_NODES (-201)     // node capacity
#define C3_E_POOL (-202)      // edge pool capacity
#define C3_E_CELLS (-203)     // DP cell pool capacity
#define C3_E_PRE (-204)       // in-degree above C3_MAXPRE
#define C3_E_BAND (-205)      // empty band
#define C3_E_BT (-206)        // backtrack found no move
#define C3_E_CIGAR (-207)     // cigar capacity
#define C3_E_QLEN (-208)      // sequence too long / empty
#define C3_E_CONS (-209)      // consensus capacity
#define C3_E_BEST (-210)

struct __align__(16) c3_pnode {
    // first 16 bytes: everything a DP row needs (one 128-bit load)
    uint16_t next, prev;          // maintained topological list order
    uint16_t in0, in_more;        // first in neighbour (C3_NONE when absent); pool index of the 2nd in edge
    uint8_t base, in_n, out_n, aln_n;
    uint16_t out0, w0;            // first out neighbour and its weight
    // second 16 bytes: graph mutation / consensus
    uint16_t out_more, rmask;     // pool index of the 2nd out edge; rmask: bit r set = read r passes here (r < 16)
    uint16_t spare, aln0;         // aligned node ids (insertion order)
    uint16_t aln1, aln2;
    uint16_t aln3, max_out;       // heaviest-bundling successor
};
static_assert(sizeof(c3_pnode) == 32, "node record must be 32 bytes");

struct c3_pedge { uint16_t id, w, next, pad; };              // overflow edge (in or out list)
// banded row: columns beg..end stored as ng = ceil(width/4) groups of 4 int32 per array; arrays
// H,E1,E2 back to back at cells[off + a*4*ng]; pad cells (> end) hold NEG_INF.  F is not stored:
// the backtrack recomputes it along one row when it needs it.  mp = (arg-max column of the row)+1,
// pulled by the successors for their adaptive band; in0/base/npre spare the node-record load.
// rows[v] (by node id): link = position in processing order, mp = arg-max column + 1.
// ord[k]  (by position): link = node id, mp = position of in0 (the first predecessor's row).
struct __align__(16) c3_prow { int32_t off; uint16_t beg, end; uint16_t mp, in0; uint16_t link; uint8_t base, npre; };
static_assert(sizeof(c3_prow) == 16, "row record must be 16 bytes");
C3_HD __forceinline__ int c3_row_ng(const c3_prow &r) { return ((int)r.end - (int)r.beg + 4) >> 2; }

// first half of a node record as one 128-bit load + field decode (avoids a local-memory struct copy)
struct c3_nrec { uint4 a; };
C3_HD __forceinline__ c3_nrec c3_ld_node(const c3_pnode *p)
{
    c3_nrec r; r.a = *reinterpret_cast<const uint4 *>(p); return r;
}
#define C3_N_NEXT(r) ((int)((r).a.x & 0xffffu))
#define C3_N_PREV(r) ((int)((r).a.x >> 16))
#define C3_N_IN0(r) ((int)((r).a.y & 0xffffu))
#define C3_N_INMORE(r) ((int)((r).a.y >> 16))
#define C3_N_BASE(r) ((int)((r).a.z & 0xffu))
#define C3_N_INN(r) ((int)(((r).a.z >> 8) & 0xffu))
#define C3_N_OUTN(r) ((int)(((r).a.z >> 16) & 0xffu))
#define C3_N_OUT0(r) ((int)((r).a.w & 0xffffu))
#define C3_N_W0(r) ((int)((r).a.w >> 16))

struct c3_poa_para_dev {
    int match, mismatch, o1, e1, o2, e2, wb, simd_bits;
    double wf;
    int int8_lanes, end_clamp;     // named switches of the abPOA restatement, c3_set_abpoa_switches; default 0
};

#ifdef C3_POA_STATS
__device__ unsigned long long c3_poa_stats[16];
#define C3_STAT(i, v) do { if (lane == 0) atomicAdd(&c3_poa_stats[i], (unsigned long long)(v)); } while (0)
#else
#define C3_STAT(i, v) do { } while (0)
#endif

struct c3_poa_args {
    const uint8_t *codes;          // base codes of all sequences
    const int64_t *item_base;      // [n_items] offset of the item's sequence block in codes
    const int32_t *bounds;         // [n_items][max_seqs][2] (start,end) relative to item_base
    const int32_t *n_seqs;         // [n_items] (via stride, see n_seqs_stride)
    int n_seqs_stride;             // in int32 units (lets n_seqs alias c3_read_result.n_sub)
    int n_items, max_seqs, min_seqs;
    int msa2;                      // 1: groups of exactly 2 sequences return their two MSA rows instead of a consensus
    int ok_status;                 // status written for a successful msa2 item (fused path keeps 2, c3_poa_batch uses 0)
    c3_poa_para_dev P;
    // per-warp workspace
    uint8_t *ws; int64_t ws_stride;
    int node_cap, pool_cap, cell_cap, cigar_cap, qp_stride;   // cell_cap in int32, multiple of 4
    // outputs
    char *cons; int cons_cap;
    int32_t *status; int32_t *cons_len; int32_t *nodes_out; long long *cells_out;
    int out_stride, cells_stride;  // strides of the int32 outputs / of cells_out, in int32 units
    unsigned *counter;
    const int32_t *order;          // optional work order (largest estimated cost first); n_work entries
    int n_work;                    // number of work items handed out (== n_items when order is null)
    const int32_t *done;           // optional [n_items]: 1 = already finished by c3_poa_lane_kernel, skip
};

struct c3_poa_ws {
    c3_pnode *nodes; c3_pedge *pool; c3_prow *rows; c3_prow *ord; uint32_t *hr; int32_t *cells; unsigned long long *cigar;
    int8_t *qp;       // query profile: 4 rows (A,C,G,T node base) x qp_stride scores, index j = column
};

__host__ __device__ inline int64_t c3_poa_ws_bytes(int node_cap, int pool_cap, int cell_cap, int cigar_cap, int qp_stride)
{
    int64_t b = 0;
    b += (int64_t)node_cap * 32; b += (int64_t)pool_cap * 8; b += (int64_t)node_cap * 32;
    b += (int64_t)node_cap * 4; b += (int64_t)cell_cap * 4; b += (int64_t)cigar_cap * 8;
    b += (int64_t)qp_stride * 4;
    return (b + 255) & ~(int64_t)255;
}

C3_HD __forceinline__ c3_poa_ws c3_poa_ws_carve(uint8_t *base, int node_cap, int pool_cap, int cell_cap, int cigar_cap)
{
    c3_poa_ws w;
    w.nodes = (c3_pnode *)base; base += (int64_t)node_cap * 32;
    w.pool = (c3_pedge *)base; base += (int64_t)pool_cap * 8;
    w.rows = (c3_prow *)base; base += (int64_t)node_cap * 16;
    w.ord = (c3_prow *)base; base += (int64_t)node_cap * 16;
    w.hr = (uint32_t *)base; base += (int64_t)node_cap * 4;
    w.cells = (int32_t *)base; base += (int64_t)cell_cap * 4;
    w.cigar = (unsigned long long *)base; base += (int64_t)cigar_cap * 8;
    w.qp = (int8_t *)base;
    return w;
}

__device__ __forceinline__ long long c3_mkkey(int v, unsigned prio)
{
    return (long long)(((unsigned long long)(unsigned)v << 32) | (unsigned long long)prio);
}

C3_HD __forceinline__ int c3_score(const c3_poa_para_dev &P, int a, int b)
{
    return (a >= 4 || b >= 4) ? 0 : (a == b ? P.match : -P.mismatch);
}

C3_HD __forceinline__ uint16_t c3_aln_get(const c3_pnode &n, int k)
{
    return k == 0 ? n.aln0 : k == 1 ? n.aln1 : k == 2 ? n.aln2 : n.aln3;
}
C3_HD __forceinline__ void c3_aln_push(c3_pnode *n, uint16_t id)
{
    const int k = n->aln_n;
    if (k == 0) n->aln0 = id; else if (k == 1) n->aln1 = id; else if (k == 2) n->aln2 = id; else if (k == 3) n->aln3 = id; else return;
    n->aln_n = (uint8_t)(k + 1);
}

// ---------------- graph mutation (lane 0 only) ----------------
struct c3_graph { c3_pnode *nodes; c3_pedge *pool; int node_n, pool_n, node_cap, pool_cap, err; };

C3_HD __forceinline__ int c3_g_add_node(c3_graph &g, uint8_t base)
{
    if (g.node_n >= g.node_cap) { g.err = C3_E_NODES; return 0; }
    c3_pnode n;
    n.next = n.prev = n.in0 = n.out0 = n.in_more = n.out_more = C3_NONE;
    n.w0 = 0; n.rmask = 0; n.spare = 0; n.aln0 = n.aln1 = n.aln2 = n.aln3 = C3_NONE; n.max_out = C3_NONE;
    n.base = base; n.in_n = n.out_n = n.aln_n = 0;
    g.nodes[g.node_n] = n;
    return g.node_n++;
}

// abpoa_add_graph_edge: find (optional) else append at the END of both lists
C3_HD inline void c3_g_add_edge(c3_graph &g, int from, int to, int check)
{
    c3_pnode *f = &g.nodes[from], *t = &g.nodes[to];
    const int fo = f->out_n;
    if (check && fo > 0) {
        if (f->out0 == to) { f->w0 = (uint16_t)(f->w0 + 1); return; }
        uint16_t e = f->out_more;
        while (e != C3_NONE) {
            if (g.pool[e].id == to) { g.pool[e].w = (uint16_t)(g.pool[e].w + 1); return; }
            e = g.pool[e].next;
        }
    }
    const int ti = t->in_n;
    if (g.pool_n + 2 > g.pool_cap || fo >= 250 || ti >= 250) { g.err = C3_E_POOL; return; }
    if (ti == 0) t->in0 = (uint16_t)from;
    else {
        const uint16_t ne = (uint16_t)g.pool_n++;
        g.pool[ne].id = (uint16_t)from; g.pool[ne].w = 0; g.pool[ne].next = C3_NONE; g.pool[ne].pad = 0;
        if (t->in_more == C3_NONE) t->in_more = ne;
        else { uint16_t e = t->in_more; while (g.pool[e].next != C3_NONE) e = g.pool[e].next; g.pool[e].next = ne; }
    }
    t->in_n = (uint8_t)(ti + 1);
    if (fo == 0) { f->out0 = (uint16_t)to; f->w0 = 1; }
    else {
        const uint16_t ne = (uint16_t)g.pool_n++;
        g.pool[ne].id = (uint16_t)to; g.pool[ne].w = 1; g.pool[ne].next = C3_NONE; g.pool[ne].pad = 0;
        if (f->out_more == C3_NONE) f->out_more = ne;
        else { uint16_t e = f->out_more; while (g.pool[e].next != C3_NONE) e = g.pool[e].next; g.pool[e].next = ne; }
    }
    f->out_n = (uint8_t)(fo + 1);
}

C3_HD __forceinline__ void c3_list_insert_before(c3_graph &g, int x, int y)
{
    const uint16_t p = g.nodes[y].prev;
    g.nodes[x].prev = p; g.nodes[x].next = (uint16_t)y;
    g.nodes[p].next = (uint16_t)x; g.nodes[y].prev = (uint16_t)x;
}
C3_HD __forceinline__ void c3_list_insert_after(c3_graph &g, int x, int a)
{
    const uint16_t nx = g.nodes[a].next;
    g.nodes[x].prev = (uint16_t)a; g.nodes[x].next = nx;
    g.nodes[a].next = (uint16_t)x; g.nodes[nx].prev = (uint16_t)x;
}
// last list element of the contiguous aligned block that contains `a`, looking forward
C3_HD inline int c3_group_tail(const c3_graph &g, int a)
{
    const c3_pnode na = g.nodes[a];
    int e = a;
    for (;;) {
        const int nx = g.nodes[e].next;
        bool in_group = false;
        for (int k = 0; k < na.aln_n; ++k) in_group |= (c3_aln_get(na, k) == nx);
        if (!in_group) break;
        e = nx;
    }
    return e;
}

// ---------------------------------------------------------------------------
// Backtrack + graph merge of one aligned sequence.  All 32 lanes call it with identical
// (warp-uniform) arguments; returns 0 or a C3_E_* code, updates node_n / pool_n.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int c3_bt_merge(const c3_poa_args &A, const c3_poa_para_dev &P, const c3_poa_ws &W,
                                        const uint8_t *q, const int qlen, const int sq, const int lane,
                                        int &node_n, int &pool_n)
{
    const int o1 = P.o1, e1 = P.e1, o2 = P.o2, e2 = P.e2, oe1 = o1 + e1, oe2 = o2 + e2;
    (void)o1; (void)o2;
    int err = 0;
    // ---- best end cell over the sink's predecessors (uniform across the warp) ----
    unsigned long long *cg = W.cigar;
    int nc = 0;
    int j, kpos;                                  // current column, position (processing order) of the current row
    c3_prow rt;                                   // ord[]-style record of the current row
    {
        const c3_nrec sk = c3_ld_node(&W.nodes[C3_SINK]);
        int best_score = -0x7fffffff - 1, bj = -1, bk = -1;
        int e = C3_N_INMORE(sk);
        const int skn = C3_N_INN(sk);
        for (int k = 0; k < skn; ++k) {
            int p;
            if (k == 0) p = C3_N_IN0(sk); else { const c3_pedge pe = W.pool[e]; p = pe.id; e = pe.next; }
            const c3_prow rp = W.rows[p];
            const int en = min(qlen, (int)rp.end);
            const int val = W.cells[rp.off + en - rp.beg];
            if (val > best_score) { best_score = val; bj = en; bk = rp.link; }
        }
        if (bk < 0) return C3_E_BEST;
        kpos = bk; j = bj; rt = W.ord[kpos];
        if (qlen - bj + 8 > A.cigar_cap) return C3_E_CIGAR;
        for (int t = qlen - lane; t > bj; t -= 32)          // trailing query bases: insertions
            cg[qlen - t] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(t - 1) << 32);
        nc = qlen - bj;
    }
    // ---- backtrack: warp-cooperative.  The chain of first predecessors (in0) inside a window of
    // 32 processed rows is resolved by pointer doubling over warp shuffles; up to 31 consecutive
    // match/mismatch moves along it are then verified at once (one gather of row records, one of
    // cells).  Everything else takes the generic one-step path (abPOA's M -> E1 -> E2 -> F1 -> F2). ----
    int cur_op = C3_OP_ALL;
    while (!err && rt.link != C3_SRC && j > 0) {
        if (cur_op == C3_OP_ALL) {
            // window slot w = lane  <->  position kpos - w
            const int kw = kpos - lane;
            c3_prow rw = rt;
            if (kw >= 0 && lane > 0) rw = W.ord[kw];
            // F0[w] = window slot of in0(w); 32 = outside the window / none
            int f = 32;
            if (kw >= 0 && rw.mp != C3_NONE) { const int d = kpos - (int)rw.mp; if (d < 32) f = d; }
            int tbl[5];
            tbl[0] = f;
#pragma unroll
            for (int b2 = 1; b2 < 5; ++b2) {
                const int prev = tbl[b2 - 1];
                const int nx = __shfl_sync(C3_FULL, prev, prev & 31);
                tbl[b2] = prev < 32 ? nx : 32;
            }
            // slot of the t-th row of the chain (t = lane): compose the set bits of t
            int sl = 0;
#pragma unroll
            for (int b2 = 0; b2 < 5; ++b2) {
                const int nx = __shfl_sync(C3_FULL, tbl[b2], sl & 31);
                if ((lane >> b2) & 1) sl = sl < 32 ? nx : 32;
            }
            const bool have = sl < 32;
            const int4 rv = *reinterpret_cast<const int4 *>(&rw);
            int4 cv;
            cv.x = __shfl_sync(C3_FULL, rv.x, sl & 31); cv.y = __shfl_sync(C3_FULL, rv.y, sl & 31);
            cv.z = __shfl_sync(C3_FULL, rv.z, sl & 31); cv.w = __shfl_sync(C3_FULL, rv.w, sl & 31);
            const c3_prow rc = *reinterpret_cast<const c3_prow *>(&cv);      // record of chain row t
            const int jt = j - lane;
            const bool inb = have && jt >= 1 && jt >= (int)rc.beg && jt <= (int)rc.end;
            int ht = C3_NEG_INF;
            if (inb) ht = W.cells[rc.off + jt - rc.beg];
            const int beg_next = __shfl_down_sync(C3_FULL, (int)rc.beg, 1);
            const int end_next = __shfl_down_sync(C3_FULL, (int)rc.end, 1);
            const int h_next = __shfl_down_sync(C3_FULL, ht, 1);
            const int have_next = __shfl_down_sync(C3_FULL, (int)have, 1);
            const int st = inb ? c3_score(P, rc.base, q[jt - 1]) : 0;
            const bool ok = lane < 31 && inb && have_next && rc.link != C3_SRC &&
                            jt - 1 >= max(beg_next, (int)rc.beg) && jt - 1 <= end_next && ht == h_next + st;
            const unsigned okm = __ballot_sync(C3_FULL, ok);
            int L = __ffs(~okm) - 1;                                    // leading run of verified moves
            L = min(L, A.cigar_cap - 8 - j - nc);
            C3_STAT(9, 1); C3_STAT(10, L > 0 ? L : 0);
            if (L > 0) {
                if (lane < L) cg[nc + lane] = C3_CG_MATCH | ((unsigned long long)rc.link << 8) | ((unsigned long long)(jt - 1) << 32);
                nc += L; j -= L;
                kpos -= __shfl_sync(C3_FULL, sl, L);
                int4 nv;
                nv.x = __shfl_sync(C3_FULL, cv.x, L); nv.y = __shfl_sync(C3_FULL, cv.y, L);
                nv.z = __shfl_sync(C3_FULL, cv.z, L); nv.w = __shfl_sync(C3_FULL, cv.w, L);
                rt = *reinterpret_cast<const c3_prow *>(&nv);
                continue;
            }
        }
        // generic single step
        C3_STAT(11, 1);
        const int i = rt.link;
        const int b = rt.beg, st4 = 4 * c3_row_ng(rt);
        const int32_t *H = W.cells + rt.off, *E1 = H + st4, *E2 = E1 + st4;
        if (j < b || j > (int)rt.end) { err = C3_E_BT; break; }
        const int s = c3_score(P, rt.base, q[j - 1]);
        const int hij = H[j - b];
        const int npre = rt.npre;
        const int in_more = npre > 1 ? (int)W.nodes[i].in_more : (int)C3_NONE;
        int hit = 0;
        unsigned long long opw = 0;
        if (cur_op & C3_OP_M) {
            int e = in_more;
            for (int k = 0; k < npre; ++k) {
                int pk;                                              // position of predecessor k
                if (k == 0) pk = rt.mp; else { const c3_pedge pe = W.pool[e]; pk = W.rows[pe.id].link; e = pe.next; }
                const c3_prow pr = W.ord[pk];
                if (j - 1 < max((int)pr.beg, b) || j - 1 > (int)pr.end) continue;
                if (W.cells[pr.off + j - 1 - pr.beg] + s == hij) {
                    opw = C3_CG_MATCH | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                    kpos = pk; rt = pr; --j; hit = 1; cur_op = C3_OP_ALL;
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
                const int pw = 4 * c3_row_ng(pr), pc = j - pr.beg;
                const int32_t *pH = W.cells + pr.off;
                const int ph = pH[pc], pe1 = pH[pw + pc], pe2 = pH[2 * pw + pc];
                if (cur_op & C3_OP_E1) {
                    if (cur_op & C3_OP_M) {
                        if (hij == pe1) { cur_op = (ph - oe1 == pe1) ? (C3_OP_M | C3_OP_F) : C3_OP_E1; hit = 1; }
                    } else if (E1[j - b] == pe1 - e1) {
                        cur_op = (ph - oe1 == pe1) ? (C3_OP_M | C3_OP_F) : C3_OP_E1; hit = 1;
                    }
                }
                if (!hit && (cur_op & C3_OP_E2)) {
                    if (cur_op & C3_OP_M) {
                        if (hij == pe2) { cur_op = (ph - oe2 == pe2) ? (C3_OP_M | C3_OP_F) : C3_OP_E2; hit = 1; }
                    } else if (E2[j - b] == pe2 - e2) {
                        cur_op = (ph - oe2 == pe2) ? (C3_OP_M | C3_OP_F) : C3_OP_E2; hit = 1;
                    }
                }
                if (hit) {
                    opw = C3_CG_DEL | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                    kpos = pk; rt = pr;
                    break;
                }
            }
        }
        if (!hit && (cur_op & C3_OP_F)) {
            if (j - 1 >= b) {
                // F is not stored: rebuild F[j] and F[j-1] of this row from its H
                // (F[c+1] = max(F[c]-e, H[c]-o-e); equal to the DP's F wherever F decides)
                int f1 = C3_NEG_INF, f2 = C3_NEG_INF, f1l = C3_NEG_INF, f2l = C3_NEG_INF, hl = C3_NEG_INF;
                for (int c = 0; c < j - b; ++c) {
                    hl = H[c];
                    f1l = f1; f2l = f2;
                    f1 = max(f1 - e1, hl - oe1); f2 = max(f2 - e2, hl - oe2);
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
            }
            if (hit) { opw = C3_CG_INS | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32); --j; }
        }
        if (!hit) { err = C3_E_BT; break; }
        if (lane == 0) cg[nc] = opw;
        ++nc;
        if (nc + j + 8 > A.cigar_cap) { err = C3_E_CIGAR; break; }
    }
    if (err) return err;
    for (int t = j - lane; t > 0; t -= 32)                      // leading query bases: insertions
        cg[nc + j - t] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(t - 1) << 32);
    nc += j;
    __syncwarp();

    // ---- merge (abpoa_add_graph_alignment); the cigar is walked from its tail = forward order.
    // 32 ops at a time: ops that only bump the weight of an existing edge between two matched
    // nodes are applied by all lanes at once; the rest (new nodes / edges) goes through lane 0
    // in order. ----
    {
        c3_graph g; g.nodes = W.nodes; g.pool = W.pool; g.node_n = node_n; g.pool_n = pool_n;
        g.node_cap = A.node_cap; g.pool_cap = A.pool_cap; g.err = 0;
        int last_id = C3_SRC, last_new = 0;                     // uniform
        for (int tb = nc - 1; tb >= 0; tb -= 32) {
            const int t = tb - lane;
            const bool have = t >= 0;
            const unsigned long long op = have ? cg[t] : C3_CG_DEL;
            const int kind = (int)(op & 0xff), node_id = (int)((op >> 8) & 0xffff), qpos = (int)(op >> 32);
            const bool is_match = have && kind == (int)C3_CG_MATCH;
            bool eq = false;
            if (is_match) eq = W.nodes[node_id].base == q[qpos];
            if (eq && sq < 16) W.nodes[node_id].rmask |= (uint16_t)(1u << sq);
            const unsigned m_nondel = __ballot_sync(C3_FULL, have && kind != (int)C3_CG_DEL);
            const unsigned m_eq = __ballot_sync(C3_FULL, eq);
            const unsigned lower = m_nondel & ((1u << lane) - 1u);
            const int pl = lower ? 31 - __clz(lower) : -1;      // lane of the previous non-deletion op
            const int pred_node = __shfl_sync(C3_FULL, node_id, pl < 0 ? 0 : pl);
            const int from = pl >= 0 ? pred_node : last_id;
            const bool from_ok = pl >= 0 ? ((m_eq >> pl) & 1u) != 0 : last_new == 0;
            bool done = false;
            if (eq && from_ok) {                                 // bump the existing edge from -> node_id
                c3_pnode *f = &W.nodes[from];
                if (f->out_n > 0) {
                    if ((int)f->out0 == node_id) { f->w0 = (uint16_t)(f->w0 + 1); done = true; }
                    else {
                        uint16_t e = f->out_more;
                        while (e != C3_NONE) {
                            if ((int)W.pool[e].id == node_id) { W.pool[e].w = (uint16_t)(W.pool[e].w + 1); done = true; break; }
                            e = W.pool[e].next;
                        }
                    }
                }
            }
            const unsigned m_cx = m_nondel & ~__ballot_sync(C3_FULL, done);
            C3_STAT(12, __popc(m_nondel)); C3_STAT(13, __popc(m_cx));
            __syncwarp();
            if (m_cx) {
                if (lane == 0) {
                    unsigned mc = m_cx;
                    while (mc && !g.err) {
                        const int c = __ffs(mc) - 1; mc &= mc - 1;
                        const unsigned lowc = m_nondel & ((1u << c) - 1u);
                        const int pc = lowc ? 31 - __clz(lowc) : -1;
                        if (pc >= 0 && ((m_eq >> pc) & 1u)) { last_id = (int)((cg[tb - pc] >> 8) & 0xffff); last_new = 0; }
                        const unsigned long long opc = cg[tb - c];
                        const int kc = (int)(opc & 0xff), nid = (int)((opc >> 8) & 0xffff), qp = (int)(opc >> 32);
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
                                    if (g.err) break;
                                    c3_list_insert_before(g, id, nid);
                                    c3_g_add_edge(g, last_id, id, 0);
                                    last_id = id; last_new = 1;
                                    if (sq < 16) g.nodes[id].rmask = (uint16_t)(1u << sq);
                                    for (int k = 0; k < nm.aln_n; ++k) {     // abpoa_add_graph_aligned_node
                                        const int a = c3_aln_get(nm, k);
                                        c3_aln_push(&g.nodes[a], (uint16_t)id);
                                        c3_aln_push(&g.nodes[id], (uint16_t)a);
                                    }
                                    c3_aln_push(&g.nodes[nid], (uint16_t)id);
                                    c3_aln_push(&g.nodes[id], (uint16_t)nid);
                                }
                            } else {
                                c3_g_add_edge(g, last_id, nid, 1 - last_new);
                                last_id = nid; last_new = 0;
                            }
                        } else {                                     // insertion
                            const int id = c3_g_add_node(g, q[qp]);
                            if (g.err) break;
                            c3_list_insert_after(g, id, c3_group_tail(g, last_id));
                            c3_g_add_edge(g, last_id, id, 0);
                            last_id = id; last_new = 1;
                            if (sq < 16) g.nodes[id].rmask = (uint16_t)(1u << sq);
                        }
                    }
                }
                g.err = __shfl_sync(C3_FULL, g.err, 0);
                g.node_n = __shfl_sync(C3_FULL, g.node_n, 0);
                g.pool_n = __shfl_sync(C3_FULL, g.pool_n, 0);
                last_id = __shfl_sync(C3_FULL, last_id, 0);
                last_new = __shfl_sync(C3_FULL, last_new, 0);
            }
            if (m_nondel) {                                          // state after the chunk
                const int ln = 31 - __clz(m_nondel);
                if ((m_eq >> ln) & 1u) { last_id = __shfl_sync(C3_FULL, node_id, ln); last_new = 0; }
            }
            __syncwarp();
            if (g.err) break;
        }
        if (!g.err && lane == 0) c3_g_add_edge(g, last_id, C3_SINK, 1 - last_new);
        g.err = __shfl_sync(C3_FULL, g.err, 0);
        g.pool_n = __shfl_sync(C3_FULL, g.pool_n, 0);
        err = g.err;
        node_n = g.node_n; pool_n = g.pool_n;
    }
    return err;
}

// Two MSA rows of a 2-sequence graph (the reference's pairwise path, bin/determine_consensus.py:33-41
// -> abpoa_generate_rc_msa with LIFO rank order), written as [row0 | row1].  Warp-uniform call;
// returns the number of columns or a C3_E_* code.
__device__ __forceinline__ int c3_emit_msa(const c3_poa_args &A, const c3_poa_ws &W, const int node_n, const int lane, char *co)
{
    int32_t *rank = (int32_t *)W.hr, *indeg = (int32_t *)W.rows, *stk = (int32_t *)W.ord;
    for (int v = lane; v < node_n; v += 32) { rank[v] = 0; indeg[v] = W.nodes[v].in_n; }
    __syncwarp();
    int msa_len = 0;
    if (lane == 0) {
        int top = 0, msa_rank = 0, ok = 0;
        stk[top++] = C3_SRC; rank[C3_SRC] = -1;
        while (top > 0) {
            const int cur = stk[--top];
            const c3_pnode nd = W.nodes[cur];
            if (rank[cur] < 0) {
                rank[cur] = msa_rank;
                for (int k = 0; k < nd.aln_n; ++k) rank[c3_aln_get(nd, k)] = msa_rank;
                ++msa_rank;
            }
            if (cur == C3_SINK) { ok = 1; break; }
            uint16_t e = nd.out_more;
            for (int k = 0; k < nd.out_n; ++k) {
                int o;
                if (k == 0) o = nd.out0; else { const c3_pedge pe = W.pool[e]; o = pe.id; e = pe.next; }
                if (--indeg[o] == 0) {
                    const c3_pnode on = W.nodes[o];
                    bool ready = true;
                    for (int a = 0; a < on.aln_n; ++a) if (indeg[c3_aln_get(on, a)] != 0) { ready = false; break; }
                    if (!ready) continue;
                    stk[top++] = o; rank[o] = -1;
                    for (int a = 0; a < on.aln_n; ++a) { const int al = c3_aln_get(on, a); stk[top++] = al; rank[al] = -1; }
                }
            }
        }
        msa_len = ok ? rank[C3_SINK] - 1 : -1;
    }
    msa_len = __shfl_sync(C3_FULL, msa_len, 0);
    __syncwarp();
    if (msa_len < 0 || 2 * msa_len > A.cons_cap) return C3_E_CONS;
    {
        for (int c = lane; c < 2 * msa_len; c += 32) co[c] = '-';
        __syncwarp();
        for (int v = 2 + lane; v < node_n; v += 32) {
            const c3_pnode nd = W.nodes[v];
            int rk = rank[v];
            for (int k = 0; k < nd.aln_n; ++k) rk = max(rk, rank[c3_aln_get(nd, k)]);
            const char ch = "ACGTN"[nd.base];
            if (nd.rmask & 1) co[rk - 1] = ch;
            if (nd.rmask & 2) co[msa_len + rk - 1] = ch;
        }
    }
    return msa_len;
}

// Heaviest bundling (abpoa_heaviest_bundling) + consensus walk.  Single-thread routine: called by
// one lane per graph; returns the consensus length or a C3_E_* code.
C3_HD __forceinline__ int c3_consensus(const c3_poa_args &A, const c3_poa_ws &W, char *co)
{
    int cons_len = 0;
    int32_t *score = (int32_t *)W.hr;
    int v = C3_SINK;
    while (v != C3_NONE) {
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
    int id = W.nodes[C3_SRC].max_out;
    while (id != C3_SINK) {
        if (id == C3_NONE || cons_len >= A.cons_cap) return C3_E_CONS;
        const c3_pnode nd = W.nodes[id];
        co[cons_len++] = "ACGTN"[nd.base];
        id = nd.max_out;
    }
    return cons_len;
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(C3_POA_THREADS, C3_POA_MINB) c3_poa_kernel(c3_poa_args A)
{
    __shared__ int4 s_ring[C3_POA_THREADS / 32][C3_RING][3 * 32];   // H,E1,E2 of recent rows: 32 groups each
    __shared__ c3_prow s_rrec[C3_POA_THREADS / 32][C3_RING];
    __shared__ int s_rid[C3_POA_THREADS / 32][4];                    // node id held by each ring slot (-1: none)
    __shared__ const int4 *s_pptr[C3_POA_THREADS / 32][C3_MAXPRE];   // predecessors 1..: row base (ring or HBM)
    __shared__ int s_pstr[C3_POA_THREADS / 32][C3_MAXPRE];           // array stride in groups
    __shared__ int s_pbe[C3_POA_THREADS / 32][C3_MAXPRE];            // beg | end << 16

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + wib;
    const c3_poa_ws W = c3_poa_ws_carve(A.ws + (int64_t)gwarp * A.ws_stride, A.node_cap, A.pool_cap, A.cell_cap, A.cigar_cap);
    const c3_poa_para_dev P = A.P;
    const int o1 = P.o1, e1 = P.e1, o2 = P.o2, e2 = P.e2, oe1 = o1 + e1, oe2 = o2 + e2;
    const int4 **pptr = s_pptr[wib];
    int *pstr = s_pstr[wib], *pbe = s_pbe[wib], *rid = s_rid[wib];
    c3_prow *rrec = s_rrec[wib];
    int4 (*ring)[3 * 32] = s_ring[wib];

    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(A.counter, 1u);
        item = __shfl_sync(C3_FULL, item, 0);
        if (item >= A.n_work) break;
        if (A.order) item = A.order[item];
        if (A.done && A.done[item]) continue;
        const int nseq = A.n_seqs[(int64_t)item * A.n_seqs_stride];
        if (nseq < A.min_seqs || nseq > A.max_seqs) continue;
        const uint8_t *ibase = A.codes + A.item_base[item];
        const int32_t *bnd = A.bounds + (int64_t)item * A.max_seqs * 2;
        int err = 0;
        long long cells_total = 0;
        int node_n = 0, pool_n = 0;

        // ---------------- first sequence -> linear graph ----------------
        {
            const uint8_t *q = ibase + bnd[0];
            const int L = bnd[1] - bnd[0];
            if (L <= 0 || L > 65000 || L + 2 > A.node_cap) err = C3_E_QLEN;
            else {
                for (int i = lane; i < L + 2; i += 32) {
                    c3_pnode n;
                    n.in_more = n.out_more = C3_NONE; n.rmask = 1; n.spare = 0;     // the first sequence (read 0) passes through every initial node
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
                node_n = L + 2;
            }
            __syncwarp();
        }

        // ---------------- align + merge the remaining sequences ----------------
        for (int sq = 1; sq < nseq && !err; ++sq) {
            const uint8_t *q = ibase + bnd[2 * sq];
            const int qlen = bnd[2 * sq + 1] - bnd[2 * sq];
            if (qlen <= 0 || qlen > 65000) { err = C3_E_QLEN; break; }
            const int n = node_n;
            // score width -> SIMD lanes of the reference build -> band granule
            const int len = qlen > n ? qlen : n;
            const int max_score = max(qlen * 5, len * e1 + o1);
            int pn = (max_score <= 32767 - P.mismatch - o1 - e1) ? P.simd_bits / 16 : P.simd_bits / 32;
            if (P.int8_lanes && max_score <= 127 - P.mismatch - o1 - e1) pn = P.simd_bits / 8;
            int w = P.wb < 0 ? qlen : P.wb + (int)(P.wf * (double)qlen);
            asm volatile("" : "+r"(w));      // keep the band half-width in a register (ptxas re-derived the fp64 product per row)

            // ---- prepare: band bookkeeping reset, heaviest successor, remain by pointer jumping ----
            for (int v = lane; v < n; v += 32) {
                const c3_pnode *nd = &W.nodes[v];
                uint32_t hv;
                if (v == C3_SINK) hv = C3_SINK;
                else {
                    int best_w = nd->w0, best = nd->out0;
                    uint16_t e = nd->out_more;
                    while (e != C3_NONE) {
                        const c3_pedge pe = W.pool[e];
                        if ((int)pe.w > best_w) { best_w = pe.w; best = pe.id; }
                        e = pe.next;
                    }
                    hv = (1u << 16) | (uint32_t)best;
                }
                W.hr[v] = hv;
            }
            __syncwarp();
            for (int round = 0; round < 20; ++round) {
                bool changed = false;
                for (int v = lane; v < n; v += 32) {
                    const uint32_t a = W.hr[v];
                    const uint32_t hnode = a & 0xffffu;
                    if (hnode != C3_SINK) {
                        const uint32_t b = W.hr[hnode];
                        W.hr[v] = ((a >> 16) + (b >> 16)) << 16 | (b & 0xffffu);
                        changed = true;
                    }
                }
                __syncwarp();
                if (!__any_sync(C3_FULL, changed)) break;
            }
            // remain(v) = hops(v -> sink) - 1  (sink: -1)

            // ---- query profile (int8 scores per node base, indexed by column j; j = 0 scores 0) ----
            {
                const int qs = A.qp_stride;
                for (int j = lane; j < qs; j += 32) {
                    const int qc = (j >= 1 && j <= qlen) ? (int)q[j - 1] : 4;
#pragma unroll
                    for (int b4 = 0; b4 < 4; ++b4)
                        W.qp[b4 * qs + j] = (int8_t)((j >= 1 && j <= qlen) ? c3_score(P, b4, qc) : 0);
                }
            }
            const int pn_shift = 31 - __clz(pn);
            // ---- DP ----
            int cell_used = 0;
            // source row
            {
                const int rem = (int)(W.hr[C3_SRC] >> 16) - 1;
                const int rr = qlen - rem;
                const int beg = max(0, min(0, rr) - w);
                const int end = min(qlen, max(0, rr) + w);
                const int b0 = (beg >> pn_shift) << pn_shift, e0 = min(qlen, (((end >> pn_shift) + 1) << pn_shift) - 1);
                const int wd = e0 - b0 + 1, ng = (wd + 3) >> 2;
                if (12 * ng > A.cell_cap) { err = C3_E_CELLS; break; }
                const bool in_ring = ng <= 32;
                if (lane == 0) {
                    c3_prow ri; ri.off = 0; ri.beg = (uint16_t)b0; ri.end = (uint16_t)e0; ri.mp = 1;   // successors of the source start at column 1
                    ri.in0 = C3_NONE; ri.base = 4; ri.npre = 0; ri.link = 0;
                    W.rows[C3_SRC] = ri;
                    rrec[0] = ri; rid[0] = in_ring ? C3_SRC : -1;
                    ri.mp = C3_NONE;
                    W.ord[0] = ri;                                   // link = node id of the source = 0; no predecessor
                    for (int t = 1; t < 4; ++t) rid[t] = -1;
                }
                int32_t *H = W.cells, *E1 = H + 4 * ng, *E2 = E1 + 4 * ng;
                int32_t *rg = reinterpret_cast<int32_t *>(&ring[0][0]);
                for (int c = lane; c < 4 * ng; c += 32) {
                    int h = C3_NEG_INF, x1 = C3_NEG_INF, x2 = C3_NEG_INF;
                    if (b0 == 0 && c < wd) {
                        if (c == 0) { h = 0; x1 = -oe1; x2 = -oe2; }
                        else h = max(-(o1 + e1 * c), -(o2 + e2 * c));
                    }
                    H[c] = h; E1[c] = x1; E2[c] = x2;
                    if (in_ring) { rg[c] = h; rg[128 + c] = x1; rg[256 + c] = x2; }
                }
                cell_used = 12 * ng;
                __syncwarp();
            }
            int v = W.nodes[C3_SRC].next;
            int rcount = 1;                                        // rows written so far (ring slot = rcount % C3_RING)
            c3_nrec nd = c3_ld_node(&W.nodes[v]);
            uint32_t hrv = W.hr[v];
            while (v != C3_SINK) {
                // prefetch the next row's node record while this row computes
                const int vnext = C3_N_NEXT(nd);
                const c3_nrec nd_next = c3_ld_node(&W.nodes[vnext]);
                const uint32_t hr_next = W.hr[vnext];
                const int rem = (int)(hrv >> 16) - 1;
                const int rr = qlen - rem;
                const int npre = C3_N_INN(nd), nbase = C3_N_BASE(nd);
                if (npre > C3_MAXPRE) { err = C3_E_PRE; break; }
                // predecessors (in-edge order): recent rows come from the shared-memory ring, others from HBM
                // the slot this row is about to overwrite is not a valid source (a predecessor C3_RING rows back
                // would be read while other lanes already store the new row there): such rows come from HBM
                const int slot = rcount & (C3_RING - 1);
                const int rid0 = slot == 0 ? -1 : rid[0], rid1 = slot == 1 ? -1 : rid[1];
                const int rid2 = slot == 2 ? -1 : rid[2], rid3 = slot == 3 ? -1 : rid[3];
                c3_prow r0; const int4 *p0ptr; int p0str;
                {
                    const int p = C3_N_IN0(nd);
                    const int sl = p == rid0 ? 0 : p == rid1 ? 1 : p == rid2 ? 2 : p == rid3 ? 3 : -1;
                    C3_STAT(0, 1); C3_STAT(1, sl < 0); C3_STAT(2, npre > 1); C3_STAT(3, npre);
                    if (sl >= 0) { r0 = rrec[sl]; p0ptr = &ring[sl][0]; p0str = 32; }
                    else { r0 = W.rows[p]; p0ptr = reinterpret_cast<const int4 *>(W.cells + r0.off); p0str = c3_row_ng(r0); }
                }
                int mpl = min(n, (int)r0.mp), mpr = r0.mp, min_pre_beg = r0.beg, max_pre_end = r0.end;
                if (npre > 1) {
                    int e = C3_N_INMORE(nd);
                    for (int k = 1; k < npre; ++k) {
                        const c3_pedge pe = W.pool[e]; e = pe.next;
                        const int p = pe.id;
                        const int sl = p == rid0 ? 0 : p == rid1 ? 1 : p == rid2 ? 2 : p == rid3 ? 3 : -1;
                        c3_prow ri; const int4 *pp; int ps;
                        if (sl >= 0) { ri = rrec[sl]; pp = &ring[sl][0]; ps = 32; }
                        else { ri = W.rows[p]; pp = reinterpret_cast<const int4 *>(W.cells + ri.off); ps = c3_row_ng(ri); }
                        if (lane == 0) { pptr[k] = pp; pstr[k] = ps; pbe[k] = (int)ri.beg | ((int)ri.end << 16); }
                        mpl = min(mpl, (int)ri.mp); mpr = max(mpr, (int)ri.mp); min_pre_beg = min(min_pre_beg, (int)ri.beg);
                        max_pre_end = max(max_pre_end, (int)ri.end);
                    }
                }
                int beg = max(0, min(mpl, rr) - w);
                int end = min(qlen, max(mpr, rr) + w);
                const int beg_sn = max(beg >> pn_shift, min_pre_beg >> pn_shift);
                int end_sn = end >> pn_shift;
                if (P.end_clamp) end_sn = min(end_sn, (max_pre_end >> pn_shift) + 1);
                end_sn = max(end_sn, beg_sn);
                beg = beg_sn << pn_shift; end = min(qlen, ((end_sn + 1) << pn_shift) - 1);
                const int wd = end - beg + 1;
                if (wd <= 0) { err = C3_E_BAND; break; }
                const int ng = (wd + 3) >> 2;
                if (cell_used + 12 * ng > A.cell_cap) { err = C3_E_CELLS; break; }
                const int off = cell_used; cell_used += 12 * ng; cells_total += wd;
                if (npre > 1) __syncwarp();
                int4 *rowv = reinterpret_cast<int4 *>(W.cells + off);
                C3_STAT(4, ng); C3_STAT(5, ng > 32); C3_STAT(6, wd);
                const bool to_ring = ng <= 32;
                int4 *ringv = &ring[slot][0];
                const int8_t *qprow = W.qp + (nbase < 4 ? nbase : 0) * A.qp_stride;
                const int p0b = r0.beg, p0e = r0.end, p0ng = (p0e - p0b + 4) >> 2;
                int carry1 = C3_NEG_INF, carry2 = C3_NEG_INF;      // F entering lane 0 of the pass
                const bool partial_row = (wd & 3) != 0;
                int bestv = C3_NEG_INF; unsigned bestp = 0;
                for (int g00 = 0; g00 < ng; g00 += 32) {
                    const int gl = g00 + lane;                     // group index within the row
                    const int g0 = beg + 4 * gl;                   // first column of this lane's group
                    const bool gact = gl < ng;
                    int m[4], x1[4], x2[4];
                    {   // predecessor 0 (registers)
                        const int gi = (g0 - p0b) >> 2;
                        int4 hv = make_int4(C3_NEG_INF, C3_NEG_INF, C3_NEG_INF, C3_NEG_INF), ev1 = hv, ev2 = hv;
                        if (gact && g0 >= p0b && gi < p0ng) {
                            hv = p0ptr[gi]; ev1 = p0ptr[p0str + gi]; ev2 = p0ptr[2 * p0str + gi];
                        }
                        int prev = __shfl_up_sync(C3_FULL, hv.w, 1);
                        if (lane == 0) {   // column g0-1: never from outside this row's band on the first pass
                            const int jc = g0 - 1;
                            prev = (g00 > 0 && jc >= p0b && jc <= p0e) ? reinterpret_cast<const int *>(p0ptr)[jc - p0b] : C3_NEG_INF;
                        }
                        m[0] = prev; m[1] = hv.x; m[2] = hv.y; m[3] = hv.z;
                        x1[0] = ev1.x; x1[1] = ev1.y; x1[2] = ev1.z; x1[3] = ev1.w;
                        x2[0] = ev2.x; x2[1] = ev2.y; x2[2] = ev2.z; x2[3] = ev2.w;
                    }
                    for (int k = 1; k < npre; ++k) {
                        const int4 *pr = pptr[k];
                        const int ps = pstr[k], pb = pbe[k] & 0xffff, pe = (pbe[k] >> 16) & 0xffff;
                        const int png = (pe - pb + 4) >> 2;
                        const int gi = (g0 - pb) >> 2;
                        int4 hv = make_int4(C3_NEG_INF, C3_NEG_INF, C3_NEG_INF, C3_NEG_INF), ev1 = hv, ev2 = hv;
                        if (gact && g0 >= pb && gi < png) { hv = pr[gi]; ev1 = pr[ps + gi]; ev2 = pr[2 * ps + gi]; }
                        int prev = __shfl_up_sync(C3_FULL, hv.w, 1);
                        if (lane == 0) {
                            const int jc = g0 - 1;
                            prev = (g00 > 0 && jc >= pb && jc <= pe) ? reinterpret_cast<const int *>(pr)[jc - pb] : C3_NEG_INF;
                        }
                        m[0] = max(m[0], prev); m[1] = max(m[1], hv.x); m[2] = max(m[2], hv.y); m[3] = max(m[3], hv.z);
                        x1[0] = max(x1[0], ev1.x); x1[1] = max(x1[1], ev1.y); x1[2] = max(x1[2], ev1.z); x1[3] = max(x1[3], ev1.w);
                        x2[0] = max(x2[0], ev2.x); x2[1] = max(x2[1], ev2.y); x2[2] = max(x2[2], ev2.z); x2[3] = max(x2[3], ev2.w);
                    }
                    // scores of the 4 columns (int8 profile; g0 is a multiple of 4)
                    int sw = 0;
                    if (gact && nbase < 4) sw = *reinterpret_cast<const int *>(qprow + g0);
                    // Cells past `end` exist only in the last group of a row whose band stops at qlen (the
                    // band end is a multiple of 4 otherwise): those rows mask explicitly; on all other rows
                    // inactive lanes carry exact NEG_INF inputs and a zero score, so no per-cell masking is needed.
                    const int nact = gact ? min(4, end - g0 + 1) : 0;      // active cells in this group
                    int hme[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int sc = (int)(int8_t)(sw >> (8 * k));
                        hme[k] = __vimax3_s32(m[k] + sc, x1[k], x2[k]);
                    }
                    if (partial_row) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) hme[k] = (k < nact) ? hme[k] : C3_NEG_INF;
                    }
                    // F: in-lane recurrence, warp prefix-max of the lane aggregates, combine
                    int ga[4], gb[4];
                    ga[0] = C3_NEG_INF; gb[0] = C3_NEG_INF;
#pragma unroll
                    for (int k = 1; k < 4; ++k) {
                        ga[k] = __viaddmax_s32(ga[k - 1], -e1, hme[k - 1] - oe1);
                        gb[k] = __viaddmax_s32(gb[k - 1], -e2, hme[k - 1] - oe2);
                    }
                    int t1 = __viaddmax_s32(ga[3], -e1, hme[3] - oe1) + 4 * e1 * lane;
                    int t2 = __viaddmax_s32(gb[3], -e2, hme[3] - oe2) + 4 * e2 * lane;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        // lanes below d get their own value back from shfl_up: max() leaves them unchanged
                        t1 = max(t1, __shfl_up_sync(C3_FULL, t1, d)); t2 = max(t2, __shfl_up_sync(C3_FULL, t2, d));
                    }
                    int c1 = __shfl_up_sync(C3_FULL, t1, 1), c2 = __shfl_up_sync(C3_FULL, t2, 1);
                    const int tot1 = __shfl_sync(C3_FULL, t1, 31), tot2 = __shfl_sync(C3_FULL, t2, 31);
                    c1 = (lane == 0) ? C3_NEG_INF : c1 - 4 * e1 * (lane - 1);
                    c2 = (lane == 0) ? C3_NEG_INF : c2 - 4 * e2 * (lane - 1);
                    c1 = max(c1, carry1 - 4 * e1 * lane);
                    c2 = max(c2, carry2 - 4 * e2 * lane);
                    carry1 = max(tot1 - 4 * e1 * 31, carry1 - 4 * e1 * 32);
                    carry2 = max(tot2 - 4 * e2 * 31, carry2 - 4 * e2 * 32);
                    int hh[4], n1v[4], n2v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int f1 = max(ga[k], c1 - k * e1), f2 = max(gb[k], c2 - k * e2);
                        hh[k] = __vimax3_s32(hme[k], f1, f2);
                        n1v[k] = __viaddmax_s32(hh[k], -oe1, x1[k] - e1);
                        n2v[k] = __viaddmax_s32(hh[k], -oe2, x2[k] - e2);
                    }
                    if (partial_row) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const bool cact = k < nact;
                            hh[k] = cact ? hh[k] : C3_NEG_INF; n1v[k] = cact ? n1v[k] : C3_NEG_INF; n2v[k] = cact ? n2v[k] : C3_NEG_INF;
                        }
                    }
                    const int lmax = gact ? max(max(hh[0], hh[1]), max(hh[2], hh[3])) : C3_NEG_INF;
                    if (gact) {
                        const int4 vh = make_int4(hh[0], hh[1], hh[2], hh[3]);
                        const int4 v1 = make_int4(n1v[0], n1v[1], n1v[2], n1v[3]);
                        const int4 v2 = make_int4(n2v[0], n2v[1], n2v[2], n2v[3]);
                        rowv[gl] = vh; rowv[ng + gl] = v1; rowv[2 * ng + gl] = v2;
                        if (to_ring) { ringv[gl] = vh; ringv[32 + gl] = v1; ringv[64 + gl] = v2; }
                    }
                    // simd_abpoa_ada_max_i: row max; ties -> lowest SIMD lane, then last vector, then earliest.
                    // The 4 columns of a group share one SIMD vector; the SIMD lane grows with k.
                    const int pm = __reduce_max_sync(C3_FULL, lmax);
                    if (pm >= bestv) {
                        unsigned pr_ = 0;
                        if (lmax == pm && gact) {
                            const int kf = hh[0] == pm ? 0 : hh[1] == pm ? 1 : hh[2] == pm ? 2 : 3;
                            const int sl = (g0 & (pn - 1)) + kf, sn = g0 >> pn_shift;
                            const unsigned vp = (sn == end_sn) ? 0xfffu : (0xffeu - (unsigned)(sn - beg_sn));
                            pr_ = ((unsigned)(pn - 1 - sl) << 12) | vp;
                        }
                        pr_ = __reduce_max_sync(C3_FULL, pr_);
                        if (pm > bestv) { bestv = pm; bestp = pr_; } else bestp = max(bestp, pr_);
                    }
                }
                int best_i = -1;
                if (bestv >= C3_NEG_HALF) {
                    const int sl = pn - 1 - (int)(bestp >> 12);
                    const unsigned vp = bestp & 0xfffu;
                    const int sn = (vp == 0xfffu) ? end_sn : beg_sn + (int)(0xffeu - vp);
                    best_i = (sn << pn_shift) + sl;
                }
                if (lane == 0) {
                    c3_prow ri; ri.off = off; ri.beg = (uint16_t)beg; ri.end = (uint16_t)end; ri.mp = (uint16_t)(best_i + 1);
                    ri.in0 = (uint16_t)C3_N_IN0(nd); ri.base = (uint8_t)nbase; ri.npre = (uint8_t)npre;
                    ri.link = (uint16_t)rcount;
                    W.rows[v] = ri;
                    rrec[slot] = ri; rid[slot] = to_ring ? v : -1;
                    ri.link = (uint16_t)v; ri.mp = r0.link;          // r0 is rows[]-style: link = position of in0
                    W.ord[rcount] = ri;
                }
                ++rcount;
                __syncwarp();
                v = vnext; nd = nd_next; hrv = hr_next;
            }
            if (err) break;

            // ---- backtrack + merge (warp-cooperative, see c3_bt_merge) ----
            err = c3_bt_merge(A, P, W, q, qlen, sq, lane, node_n, pool_n);
            __syncwarp();
        }

        // ---------------- two MSA rows (pairwise path) or heaviest-bundling consensus ----------------
        int cons_len = 0;
        const bool do_msa = A.msa2 && nseq == 2;
        if (!err) {
            char *co = A.cons + (int64_t)item * A.cons_cap;
            if (do_msa) { const int r = c3_emit_msa(A, W, node_n, lane, co); if (r < 0) err = r; else cons_len = r; }
            else {
                int r = 0;
                if (lane == 0) r = c3_consensus(A, W, co);
                r = __shfl_sync(C3_FULL, r, 0);
                if (r < 0) err = r; else cons_len = r;
            }
        }
        err = __shfl_sync(C3_FULL, err, 0);
        if (lane == 0) {
            const int64_t o = (int64_t)item * A.out_stride;
            A.status[o] = err ? err : (do_msa ? A.ok_status : 0);
            A.cons_len[o] = err ? 0 : cons_len;
            A.nodes_out[o] = node_n;
            *(long long *)((int32_t *)A.cells_out + (int64_t)item * A.cells_stride) = cells_total;
        }
        __syncwarp();
    }
}
