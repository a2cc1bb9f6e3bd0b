// poa.cuh -- stage 3b: adaptive-banded partial-order alignment + heaviest-bundling
// consensus with abPOA 1.0.5 semantics (convex gap, global mode, extra_b/extra_f band).
//
// Replaces poa.msa_aligner(match=5).msa(subreads, out_cons=True, out_msa=True)
//   (/root/reference/bin/determine_consensus.py:30-47).
//
// One warp per read (persistent grid, atomic work counter).  Per added subread:
//   prepare   (warp-parallel)  reset band bookkeeping, heaviest successor per node,
//                              "remaining path length" by pointer jumping
//   DP        (warp-parallel)  rows = graph nodes in a maintained topological list
//                              order, lanes = band columns; the horizontal (F) gap
//                              dependency is a warp prefix-max over H+e*j
//   backtrack (lane 0)         value-based, abPOA's M -> E1 -> E2 -> F1 -> F2 order with
//                              the op-mask state machine
//   merge     (lane 0)         graph update; new nodes are spliced into the list so that
//                              aligned groups stay contiguous (a valid topological
//                              order of the graph without re-sorting; every quantity
//                              the DP derives is order-independent)
// then heaviest bundling (lane 0) and the consensus walk.
//
// The graph, DP rows and cigar live in a per-warp HBM workspace (L1/L2 cached).
#pragma once
#include "common.cuh"

#define C3_POA_THREADS 128
#define C3_NONE 0xffffu
#define C3_SRC 0
#define C3_SINK 1
#define C3_NEG_INF (-(1 << 29))
#define C3_NEG_HALF (-(1 << 28))
#define C3_MAXPRE 64

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
    uint16_t next, prev;          // maintained topological list order
    uint16_t in0, out0;           // first in / out neighbour (C3_NONE when absent)
    uint16_t w0, in_more;         // weight of out0; pool index of the 2nd in edge
    uint16_t out_more, mpl;       // pool index of the 2nd out edge; max_pos_left
    uint16_t mpr, aln0;           // max_pos_right; aligned node ids (insertion order)
    uint16_t aln1, aln2;
    uint16_t aln3, max_out;       // heaviest-bundling successor
    uint8_t base, in_n, out_n, aln_n;
};
static_assert(sizeof(c3_pnode) == 32, "node record must be 32 bytes");

struct c3_pedge { uint16_t id, w, next, pad; };              // overflow edge (in or out list)
struct c3_prow { int32_t off; uint16_t beg, end; };          // banded row: cells at off, columns beg..end

struct c3_poa_para_dev {
    int match, mismatch, o1, e1, o2, e2, wb, simd_bits;
    double wf;
};

struct c3_poa_args {
    const uint8_t *codes;          // base codes of all sequences
    const int64_t *item_base;      // [n_items] offset of the item's sequence block in codes
    const int32_t *bounds;         // [n_items][max_seqs][2] (start,end) relative to item_base
    const int32_t *n_seqs;         // [n_items] (via stride, see n_seqs_stride)
    int n_seqs_stride;             // in int32 units (lets n_seqs alias c3_read_result.n_sub)
    int n_items, max_seqs, min_seqs;
    c3_poa_para_dev P;
    // per-warp workspace
    uint8_t *ws; int64_t ws_stride;
    int node_cap, pool_cap, cell_cap, cigar_cap;
    // outputs
    char *cons; int cons_cap;
    int32_t *status; int32_t *cons_len; int32_t *nodes_out; long long *cells_out;
    int out_stride, cells_stride;  // strides of the int32 outputs / of cells_out, in int32 units
    unsigned *counter;
};

struct c3_poa_ws {
    c3_pnode *nodes; c3_pedge *pool; c3_prow *rows; uint32_t *hr; int32_t *cells; unsigned long long *cigar;
};

__host__ __device__ inline int64_t c3_poa_ws_bytes(int node_cap, int pool_cap, int cell_cap, int cigar_cap)
{
    int64_t b = 0;
    b += (int64_t)node_cap * 32; b += (int64_t)pool_cap * 8; b += (int64_t)node_cap * 8;
    b += (int64_t)node_cap * 4; b += (int64_t)cell_cap * 4; b += (int64_t)cigar_cap * 8;
    return (b + 255) & ~(int64_t)255;
}

__device__ __forceinline__ c3_poa_ws c3_poa_ws_carve(uint8_t *base, int node_cap, int pool_cap, int cell_cap)
{
    c3_poa_ws w;
    w.nodes = (c3_pnode *)base; base += (int64_t)node_cap * 32;
    w.pool = (c3_pedge *)base; base += (int64_t)pool_cap * 8;
    w.rows = (c3_prow *)base; base += (int64_t)node_cap * 8;
    w.hr = (uint32_t *)base; base += (int64_t)node_cap * 4;
    w.cells = (int32_t *)base; base += (int64_t)cell_cap * 4;
    w.cigar = (unsigned long long *)base;
    return w;
}

__device__ __forceinline__ long long c3_mkkey(int v, unsigned prio)
{
    return (long long)(((unsigned long long)(unsigned)v << 32) | (unsigned long long)prio);
}

__device__ __forceinline__ int c3_score(const c3_poa_para_dev &P, int a, int b)
{
    return (a >= 4 || b >= 4) ? 0 : (a == b ? P.match : -P.mismatch);
}

__device__ __forceinline__ uint16_t c3_aln_get(const c3_pnode &n, int k)
{
    return k == 0 ? n.aln0 : k == 1 ? n.aln1 : k == 2 ? n.aln2 : n.aln3;
}
__device__ __forceinline__ void c3_aln_push(c3_pnode *n, uint16_t id)
{
    const int k = n->aln_n;
    if (k == 0) n->aln0 = id; else if (k == 1) n->aln1 = id; else if (k == 2) n->aln2 = id; else if (k == 3) n->aln3 = id; else return;
    n->aln_n = (uint8_t)(k + 1);
}

// ---------------- graph mutation (lane 0 only) ----------------
struct c3_graph { c3_pnode *nodes; c3_pedge *pool; int node_n, pool_n, node_cap, pool_cap, err; };

__device__ __forceinline__ int c3_g_add_node(c3_graph &g, uint8_t base)
{
    if (g.node_n >= g.node_cap) { g.err = C3_E_NODES; return 0; }
    c3_pnode n;
    n.next = n.prev = n.in0 = n.out0 = n.in_more = n.out_more = C3_NONE;
    n.w0 = 0; n.mpl = 0; n.mpr = 0; n.aln0 = n.aln1 = n.aln2 = n.aln3 = C3_NONE; n.max_out = C3_NONE;
    n.base = base; n.in_n = n.out_n = n.aln_n = 0;
    g.nodes[g.node_n] = n;
    return g.node_n++;
}

// abpoa_add_graph_edge: find (optional) else append at the END of both lists
__device__ void c3_g_add_edge(c3_graph &g, int from, int to, int check)
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

__device__ __forceinline__ void c3_list_insert_before(c3_graph &g, int x, int y)
{
    const uint16_t p = g.nodes[y].prev;
    g.nodes[x].prev = p; g.nodes[x].next = (uint16_t)y;
    g.nodes[p].next = (uint16_t)x; g.nodes[y].prev = (uint16_t)x;
}
__device__ __forceinline__ void c3_list_insert_after(c3_graph &g, int x, int a)
{
    const uint16_t nx = g.nodes[a].next;
    g.nodes[x].prev = (uint16_t)a; g.nodes[x].next = nx;
    g.nodes[a].next = (uint16_t)x; g.nodes[nx].prev = (uint16_t)x;
}
// last list element of the contiguous aligned block that contains `a`, looking forward
__device__ int c3_group_tail(const c3_graph &g, int a)
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
__global__ void __launch_bounds__(C3_POA_THREADS) c3_poa_kernel(c3_poa_args A)
{
    __shared__ int s_poff[C3_POA_THREADS / 32][C3_MAXPRE];
    __shared__ int s_pbe[C3_POA_THREADS / 32][C3_MAXPRE];     // beg | end << 16

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + wib;
    const c3_poa_ws W = c3_poa_ws_carve(A.ws + (int64_t)gwarp * A.ws_stride, A.node_cap, A.pool_cap, A.cell_cap);
    const c3_poa_para_dev P = A.P;
    const int o1 = P.o1, e1 = P.e1, o2 = P.o2, e2 = P.e2, oe1 = o1 + e1, oe2 = o2 + e2;
    int *poff = s_poff[wib], *pbe = s_pbe[wib];

    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(A.counter, 1u);
        item = __shfl_sync(C3_FULL, item, 0);
        if (item >= A.n_items) break;
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
                    n.in_more = n.out_more = C3_NONE; n.mpl = 0; n.mpr = 0;
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
            const int pn = (max_score <= 32767 - P.mismatch - o1 - e1) ? P.simd_bits / 16 : P.simd_bits / 32;
            const int w = P.wb < 0 ? qlen : P.wb + (int)(P.wf * (double)qlen);

            // ---- prepare: band bookkeeping reset, heaviest successor, remain by pointer jumping ----
            for (int v = lane; v < n; v += 32) {
                c3_pnode *nd = &W.nodes[v];
                nd->mpl = (uint16_t)n; nd->mpr = 0;
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

            // ---- DP ----
            int cell_used = 0;
            // source row
            {
                if (lane == 0) {
                    c3_pnode *s = &W.nodes[C3_SRC];
                    s->mpl = 0; s->mpr = 0;
                    W.nodes[s->out0].mpl = 1; W.nodes[s->out0].mpr = 1;
                    uint16_t e = s->out_more;
                    while (e != C3_NONE) { const c3_pedge pe = W.pool[e]; W.nodes[pe.id].mpl = 1; W.nodes[pe.id].mpr = 1; e = pe.next; }
                }
                const int rem = (int)(W.hr[C3_SRC] >> 16) - 1;
                const int rr = qlen - rem;
                const int beg = max(0, min(0, rr) - w);
                const int end = min(qlen, max(0, rr) + w);
                const int beg_sn = beg / pn, end_sn = end / pn;
                const int b0 = beg_sn * pn, e0 = min(qlen, (end_sn + 1) * pn - 1);
                const int wd = e0 - b0 + 1;
                if (5 * wd > A.cell_cap) { err = C3_E_CELLS; break; }
                if (lane == 0) { c3_prow ri; ri.off = 0; ri.beg = (uint16_t)b0; ri.end = (uint16_t)e0; W.rows[C3_SRC] = ri; }
                int32_t *H = W.cells, *E1 = H + wd, *E2 = E1 + wd, *F1 = E2 + wd, *F2 = F1 + wd;
                for (int c = lane; c < wd; c += 32) {
                    int h = C3_NEG_INF, x1 = C3_NEG_INF, x2 = C3_NEG_INF, f1 = C3_NEG_INF, f2 = C3_NEG_INF;
                    if (b0 == 0) {
                        if (c == 0) { h = 0; x1 = -oe1; x2 = -oe2; }
                        else { f1 = -(o1 + e1 * c); f2 = -(o2 + e2 * c); h = max(f1, f2); }
                    }
                    H[c] = h; E1[c] = x1; E2[c] = x2; F1[c] = f1; F2[c] = f2;
                }
                cell_used = 5 * wd;
                __syncwarp();
            }
            int v = W.nodes[C3_SRC].next;
            while (v != C3_SINK) {
                const c3_pnode nd = W.nodes[v];
                const int rem = (int)(W.hr[v] >> 16) - 1;
                const int rr = qlen - rem;
                int beg = max(0, min((int)nd.mpl, rr) - w);
                int end = min(qlen, max((int)nd.mpr, rr) + w);
                int beg_sn = beg / pn, end_sn = end / pn;
                // predecessors (in-edge order), cached in shared memory
                const int npre = nd.in_n;
                if (npre > C3_MAXPRE) { err = C3_E_PRE; break; }
                int min_pre_beg = 0x7fffffff;
                {
                    uint16_t e = nd.in_more;
                    for (int k = 0; k < npre; ++k) {
                        int p;
                        if (k == 0) p = nd.in0; else { const c3_pedge pe = W.pool[e]; p = pe.id; e = pe.next; }
                        const c3_prow ri = W.rows[p];
                        if (lane == 0) { poff[k] = ri.off; pbe[k] = (int)ri.beg | ((int)ri.end << 16); }
                        min_pre_beg = min(min_pre_beg, (int)ri.beg);
                    }
                }
                if (beg_sn < min_pre_beg / pn) beg_sn = min_pre_beg / pn;
                if (end_sn < beg_sn) end_sn = beg_sn;
                beg = beg_sn * pn; end = min(qlen, (end_sn + 1) * pn - 1);
                const int wd = end - beg + 1;
                if (wd <= 0) { err = C3_E_BAND; break; }
                if (cell_used + 5 * wd > A.cell_cap) { err = C3_E_CELLS; break; }
                const int off = cell_used; cell_used += 5 * wd; cells_total += wd;
                if (lane == 0) { c3_prow ri; ri.off = off; ri.beg = (uint16_t)beg; ri.end = (uint16_t)end; W.rows[v] = ri; }
                __syncwarp();
                int32_t *H = W.cells + off, *E1 = H + wd, *E2 = E1 + wd, *F1 = E2 + wd, *F2 = F1 + wd;
                const int base = nd.base;
                int carry1 = C3_NEG_INF, carry2 = C3_NEG_INF;
                long long bestkey = c3_mkkey(C3_NEG_INF, 0u);           // value | priority
                for (int c0 = 0; c0 < wd; c0 += 32) {
                    const int c = c0 + lane, j = beg + c;
                    const bool act = c < wd;
                    int M = C3_NEG_INF, x1 = C3_NEG_INF, x2 = C3_NEG_INF;
                    for (int k = 0; k < npre; ++k) {
                        const int po = poff[k], pb = pbe[k] & 0xffff, pe = (pbe[k] >> 16) & 0xffff, pw = pe - pb + 1;
                        const int32_t *pH = W.cells + po;
                        if (act && j - 1 >= max(pb, beg) && j - 1 <= pe) M = max(M, pH[j - 1 - pb]);
                        if (act && j >= pb && j <= pe) { x1 = max(x1, pH[pw + j - pb]); x2 = max(x2, pH[2 * pw + j - pb]); }
                    }
                    const int s = (act && j > 0) ? c3_score(P, base, q[j - 1]) : 0;
                    const int m = M + s;
                    int hme = max(m, max(x1, x2));
                    if (!act) hme = C3_NEG_INF;
                    int a1 = hme + e1 * j, a2 = hme + e2 * j;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t1 = __shfl_up_sync(C3_FULL, a1, d), t2 = __shfl_up_sync(C3_FULL, a2, d);
                        if (lane >= d) { a1 = max(a1, t1); a2 = max(a2, t2); }
                    }
                    int p1 = __shfl_up_sync(C3_FULL, a1, 1), p2 = __shfl_up_sync(C3_FULL, a2, 1);
                    if (lane == 0) { p1 = C3_NEG_INF; p2 = C3_NEG_INF; }
                    p1 = max(p1, carry1); p2 = max(p2, carry2);
                    carry1 = max(carry1, __shfl_sync(C3_FULL, a1, 31));
                    carry2 = max(carry2, __shfl_sync(C3_FULL, a2, 31));
                    const int f1 = p1 - o1 - e1 * j, f2 = p2 - o2 - e2 * j;
                    const int h = max(hme, max(f1, f2));
                    if (act) {
                        H[c] = h; F1[c] = f1; F2[c] = f2;
                        E1[c] = max(h - oe1, x1 - e1);
                        E2[c] = max(h - oe2, x2 - e2);
                        // simd_abpoa_ada_max_i tie-break: lowest SIMD lane, then last vector, then earliest vector
                        const int sl = c % pn, sn = j / pn;
                        const unsigned vp = (sn == end_sn) ? 0xfffffu : (0xffffeu - (unsigned)(sn - beg_sn));
                        const unsigned prio = ((unsigned)(pn - 1 - sl) << 20) | vp;
                        const long long key = c3_mkkey(h, prio);
                        if (key > bestkey) bestkey = key;
                    }
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const long long o = __shfl_xor_sync(C3_FULL, bestkey, d);
                    if (o > bestkey) bestkey = o;
                }
                int best_i = -1;
                {
                    const int bv = (int)(bestkey >> 32);
                    if (bv >= C3_NEG_HALF) {
                        const unsigned prio = (unsigned)(bestkey & 0xffffffffll);
                        const int sl = pn - 1 - (int)(prio >> 20);
                        const unsigned vp = prio & 0xfffffu;
                        const int sn = (vp == 0xfffffu) ? end_sn : beg_sn + (int)(0xffffeu - vp);
                        best_i = sn * pn + sl;
                    }
                }
                if (lane == 0) {
                    const int mp = best_i + 1;
                    uint16_t e = nd.out_more;
                    for (int k = 0; k < nd.out_n; ++k) {
                        int o;
                        if (k == 0) o = nd.out0; else { const c3_pedge pe = W.pool[e]; o = pe.id; e = pe.next; }
                        c3_pnode *on = &W.nodes[o];
                        if (mp > (int)on->mpr) on->mpr = (uint16_t)mp;
                        if (mp < (int)on->mpl) on->mpl = (uint16_t)mp;
                    }
                }
                __syncwarp();
                v = nd.next;
            }
            if (err) break;

            // ---- best end cell over the sink's predecessors + backtrack + merge (lane 0) ----
            int n_new_nodes = node_n, n_new_pool = pool_n;
            if (lane == 0) {
                const c3_pnode sk = W.nodes[C3_SINK];
                int best_score = -0x7fffffff - 1, bi = -1, bj = -1;
                {
                    uint16_t e = sk.in_more;
                    for (int k = 0; k < sk.in_n; ++k) {
                        int p;
                        if (k == 0) p = sk.in0; else { const c3_pedge pe = W.pool[e]; p = pe.id; e = pe.next; }
                        const c3_prow ri = W.rows[p];
                        const int en = min(qlen, (int)ri.end);
                        const int val = W.cells[ri.off + en - ri.beg];
                        if (val > best_score) { best_score = val; bi = p; bj = en; }
                    }
                }
                int nc = 0;
                unsigned long long *cg = W.cigar;
                if (bi < 0) err = C3_E_BEST;
                int i = bi, j = bj;
                if (!err) {
                    if (qlen - bj + 8 > A.cigar_cap) err = C3_E_CIGAR;
                    else for (int t = qlen; t > bj; --t) cg[nc++] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(t - 1) << 32);
                }
                int cur_op = C3_OP_ALL;
                while (!err && i != C3_SRC && j > 0) {
                    const c3_pnode nd = W.nodes[i];
                    const c3_prow ri = W.rows[i];
                    const int b = ri.beg, wd = (int)ri.end - b + 1;
                    const int32_t *H = W.cells + ri.off, *E1 = H + wd, *E2 = E1 + wd, *F1 = E2 + wd, *F2 = F1 + wd;
                    if (j < b || j > (int)ri.end) { err = C3_E_BT; break; }
                    const int s = c3_score(P, nd.base, q[j - 1]);
                    const int hij = H[j - b];
                    int hit = 0;
                    if (cur_op & C3_OP_M) {
                        uint16_t e = nd.in_more;
                        for (int k = 0; k < nd.in_n; ++k) {
                            int p;
                            if (k == 0) p = nd.in0; else { const c3_pedge pe = W.pool[e]; p = pe.id; e = pe.next; }
                            const c3_prow pr = W.rows[p];
                            if (j - 1 < max((int)pr.beg, b) || j - 1 > (int)pr.end) continue;
                            if (W.cells[pr.off + j - 1 - pr.beg] + s == hij) {
                                cg[nc++] = C3_CG_MATCH | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                                i = p; --j; hit = 1; cur_op = C3_OP_ALL;
                                break;
                            }
                        }
                    }
                    if (!hit && (cur_op & C3_OP_E)) {
                        uint16_t e = nd.in_more;
                        for (int k = 0; k < nd.in_n; ++k) {
                            int p;
                            if (k == 0) p = nd.in0; else { const c3_pedge pe = W.pool[e]; p = pe.id; e = pe.next; }
                            const c3_prow pr = W.rows[p];
                            if (j < (int)pr.beg || j > (int)pr.end) continue;
                            const int pw = (int)pr.end - pr.beg + 1, pc = j - pr.beg;
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
                                cg[nc++] = C3_CG_DEL | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                                i = p;
                                break;
                            }
                        }
                    }
                    if (!hit && (cur_op & C3_OP_F)) {
                        if (j - 1 >= b) {
                            const int hl = H[j - 1 - b];
                            if (cur_op & C3_OP_F1) {
                                const int f = F1[j - b], fl = F1[j - 1 - b];
                                if (!(cur_op & C3_OP_M) || hij == f) {
                                    if (hl - oe1 == f) { cur_op = C3_OP_M | C3_OP_E; hit = 1; }
                                    else if (fl - e1 == f) { cur_op = C3_OP_F1; hit = 1; }
                                }
                            }
                            if (!hit && (cur_op & C3_OP_F2)) {
                                const int f = F2[j - b], fl = F2[j - 1 - b];
                                if (!(cur_op & C3_OP_M) || hij == f) {
                                    if (hl - oe2 == f) { cur_op = C3_OP_M | C3_OP_E; hit = 1; }
                                    else if (fl - e2 == f) { cur_op = C3_OP_F2; hit = 1; }
                                }
                            }
                        }
                        if (hit) { cg[nc++] = C3_CG_INS | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32); --j; }
                    }
                    if (!hit) { err = C3_E_BT; break; }
                    if (nc + j + 8 > A.cigar_cap) { err = C3_E_CIGAR; break; }
                }
                if (!err) for (; j > 0; --j) cg[nc++] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(j - 1) << 32);

                // ---- merge (abpoa_add_graph_alignment), cigar walked from its tail = forward order ----
                if (!err) {
                    c3_graph g; g.nodes = W.nodes; g.pool = W.pool; g.node_n = node_n; g.pool_n = pool_n;
                    g.node_cap = A.node_cap; g.pool_cap = A.pool_cap; g.err = 0;
                    int last_id = C3_SRC, last_new = 0;
                    for (int t = nc - 1; t >= 0 && !g.err; --t) {
                        const unsigned long long op = cg[t];
                        const int kind = (int)(op & 0xff), node_id = (int)((op >> 8) & 0xffff), qpos = (int)(op >> 32);
                        if (kind == (int)C3_CG_MATCH) {
                            const uint8_t bq = q[qpos];
                            const c3_pnode nm = g.nodes[node_id];
                            if (nm.base != bq) {
                                int al = -1;
                                for (int k = 0; k < nm.aln_n; ++k) {
                                    const int a = c3_aln_get(nm, k);
                                    if (g.nodes[a].base == bq) { al = a; break; }
                                }
                                if (al != -1) {
                                    c3_g_add_edge(g, last_id, al, 1 - last_new);
                                    last_id = al; last_new = 0;
                                } else {
                                    const int id = c3_g_add_node(g, bq);
                                    if (g.err) break;
                                    c3_list_insert_before(g, id, node_id);
                                    c3_g_add_edge(g, last_id, id, 0);
                                    last_id = id; last_new = 1;
                                    // abpoa_add_graph_aligned_node
                                    for (int k = 0; k < nm.aln_n; ++k) {
                                        const int a = c3_aln_get(nm, k);
                                        c3_aln_push(&g.nodes[a], (uint16_t)id);
                                        c3_aln_push(&g.nodes[id], (uint16_t)a);
                                    }
                                    c3_aln_push(&g.nodes[node_id], (uint16_t)id);
                                    c3_aln_push(&g.nodes[id], (uint16_t)node_id);
                                }
                            } else {
                                c3_g_add_edge(g, last_id, node_id, 1 - last_new);
                                last_id = node_id; last_new = 0;
                            }
                        } else if (kind == (int)C3_CG_INS) {
                            const int id = c3_g_add_node(g, q[qpos]);
                            if (g.err) break;
                            c3_list_insert_after(g, id, c3_group_tail(g, last_id));
                            c3_g_add_edge(g, last_id, id, 0);
                            last_id = id; last_new = 1;
                        }
                    }
                    if (!g.err) c3_g_add_edge(g, last_id, C3_SINK, 1 - last_new);
                    err = g.err;
                    n_new_nodes = g.node_n; n_new_pool = g.pool_n;
                }
            }
            err = __shfl_sync(C3_FULL, err, 0);
            node_n = __shfl_sync(C3_FULL, n_new_nodes, 0);
            pool_n = __shfl_sync(C3_FULL, n_new_pool, 0);
            __syncwarp();
        }

        // ---------------- heaviest bundling + consensus walk (lane 0) ----------------
        int cons_len = 0;
        if (!err && lane == 0) {
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
            char *co = A.cons + (int64_t)item * A.cons_cap;
            int id = W.nodes[C3_SRC].max_out;
            while (id != C3_SINK) {
                if (id == C3_NONE || cons_len >= A.cons_cap) { err = C3_E_CONS; break; }
                const c3_pnode nd = W.nodes[id];
                co[cons_len++] = "ACGTN"[nd.base];
                id = nd.max_out;
            }
        }
        err = __shfl_sync(C3_FULL, err, 0);
        if (lane == 0) {
            const int64_t o = (int64_t)item * A.out_stride;
            A.status[o] = err;
            A.cons_len[o] = err ? 0 : cons_len;
            A.nodes_out[o] = node_n;
            *(long long *)((int32_t *)A.cells_out + (int64_t)item * A.cells_stride) = cells_total;
        }
        __syncwarp();
    }
}
