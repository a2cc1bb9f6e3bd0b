// poa_graph.cuh -- stage 3b, the serial phases of the group POA: one THREAD per read.
//
// Between two DP launches (poa_grp.cuh) a read needs: backtrack of the alignment just computed, merge of the
// path into the graph, the new topological order, and the row descriptors of the next alignment; before the first
// one: the first sequence as a linear graph; after the last one: heaviest bundling and the consensus walk.
// All of that is pointer chasing with a few thousand dependent steps per read and no parallelism inside a read
// worth the name -- measured: with 8 or 32 lanes per read those phases ran ONE read per warp instruction and cost
// more than the DP.  So here every thread runs the plain scalar algorithm on its own read, all reads of the wave
// at once (a B200 holds >300 000 threads): the latency of each dependent load is hidden by the other reads, and
// the instruction stream is shared by the 32 reads of a warp.  The data is the same position-ordered layout the
// DP kernel reads and writes (poa_grp.cuh); the next row record / descriptor / cell of the usual path are
// requested one step ahead.
//
// Same results as poa.cuh (abPOA 1.0.5 semantics; /root/reference/bin/determine_consensus.py:30-47).
// Everything here is plain host+device code: tests/emul/grp_emul.cu calls it directly.
#pragma once
#include "poa_grp.cuh"

#ifndef C3S_PFK
#define C3S_PFK 0                     // kind of the sweeps' software prefetch: 0 none, 1 towards L1, 2 towards L2
#endif
#if C3S_PFK == 2
#define C3S_PREFETCH(p) C3L_PREFETCH2(p)
#elif C3S_PFK == 1
#define C3S_PREFETCH(p) C3L_PREFETCH(p)
#else
#define C3S_PREFETCH(p) do { (void)(p); } while (0)
#endif
#ifndef C3S_PF
#define C3S_PF 6                      // look-ahead (loop steps) of the software prefetches in the sweeps
#endif

// Sequential streams of small elements are read and written in 16-byte pieces kept in registers: with several hundred
// threads per SM, each on its own read, a line does not survive in L1 (nor reliably in L2) from one element to the next,
// so an element-wise sweep would fetch the same sector up to 16 times.
struct c3s_q4 { const uint8_t *q; uintptr_t wa; uint32_t w; };       // byte reader over the query codes, one word cached
C3_HD __forceinline__ void c3s_q4_init(c3s_q4 &c, const uint8_t *q) { c.q = q; c.wa = ~(uintptr_t)0; c.w = 0u; }
C3_HD __forceinline__ int c3s_q4_get(c3s_q4 &c, const int idx)
{
    const uintptr_t a = (uintptr_t)(c.q + idx), wa = a & ~(uintptr_t)3;
    if (wa != c.wa) { c.wa = wa; c.w = *reinterpret_cast<const uint32_t *>(wa); }
    return (int)((c.w >> (8 * (int)(a & 3))) & 0xffu);
}
C3_HD __forceinline__ int c3s_u16_of(const uint4 v, const int k)      // halfword k (0..7) of a 16-byte piece
{
    const uint32_t w = (k & 4) ? ((k & 2) ? v.w : v.z) : ((k & 2) ? v.y : v.x);
    return (int)((w >> ((k & 1) * 16)) & 0xffffu);
}
struct c3s_out8 { uint16_t *dst; unsigned long long lo, hi; int n; };  // sequential uint16 writer, 8 per store
C3_HD __forceinline__ void c3s_out8_push(c3s_out8 &o, const int v)
{
    o.lo = (o.lo >> 16) | (o.hi << 48);
    o.hi = (o.hi >> 16) | ((unsigned long long)(unsigned)v << 48);
    if ((++o.n & 7) == 0)
        *reinterpret_cast<uint4 *>(o.dst + o.n - 8) = make_uint4((uint32_t)o.lo, (uint32_t)(o.lo >> 32), (uint32_t)o.hi, (uint32_t)(o.hi >> 32));
}
C3_HD __forceinline__ void c3s_out8_flush(c3s_out8 &o)                // the last, partial piece: element-wise
{
    const int r = o.n & 7;
    if (!r) return;
    // the r newest values sit in the top r halfwords of (hi:lo)
    for (int k = 0; k < r; ++k) {
        const int sh = (8 - r + k) * 16;
        const unsigned long long w = sh >= 64 ? (o.hi >> (sh - 64)) : (o.lo >> sh);
        o.dst[o.n - r + k] = (uint16_t)(w & 0xffffu);
    }
}

// First sequence -> linear graph, order = SRC, 2, 3, ..., L+1, SINK, and the row descriptors of the first alignment
// (a chain: predecessor = position - 1, hops to the sink = n - 1 - position).  `tid` of `nthr` workers share the
// node loop: the init kernel gives a read a whole CTA (coalesced stores), the emulation one worker.
C3_HD inline void c3s_item_begin(c3g_grp &G, const c3_poa_args &A, const c3_poa_para_dev &P, const c3g_ws &W, const int item,
                                 const int tid, const int nthr)
{
    G.item = item; G.sq = 1; G.err = 0; G.nseq = 0; G.node_n = 0; G.pool_n = 0; G.cells_total = 0; G.ob = 0;
    G.qlen = 0; G.n = 0; G.w = 0;
    const int nseq = A.n_seqs[(int64_t)item * A.n_seqs_stride];
    if (nseq < A.min_seqs || nseq > A.max_seqs || nseq < 1) { C3G_DECLINE(); G.err = C3G_E_RETRY; return; }
    G.ibase = A.codes + A.item_base[item];
    G.bnd = A.bounds + (int64_t)item * A.max_seqs * 2;
    G.nseq = nseq;
    const uint8_t *q = G.ibase + G.bnd[0];
    const int L = G.bnd[1] - G.bnd[0];
    if (L <= 0 || L > 65000 || L + 2 > A.node_cap) { C3G_DECLINE(); G.err = C3G_E_RETRY; return; }
    G.node_n = L + 2;
    if (nseq > 1) {                                // what c3s_prepare checks and sets, for the second sequence
        const int qlen = G.bnd[3] - G.bnd[2], n = L + 2;
        const int len = qlen > n ? qlen : n;
        const int max_score = max(qlen * 5, len * P.e1 + P.o1);
        const int pn = (max_score <= 32767 - P.mismatch - P.o1 - P.e1) ? P.simd_bits / 16 : P.simd_bits / 32;
        if (qlen <= 0 || qlen > 65000 || qlen + 32 > A.qp_stride || pn != 16 || P.wb < 0) { C3G_DECLINE(); G.err = C3G_E_RETRY; return; }
        G.q = G.ibase + G.bnd[2]; G.qlen = qlen; G.n = n; G.w = P.wb + (int)(P.wf * (double)qlen);
    }
    uint16_t *ord = W.order[0];
    for (int i = tid; i < L + 2; i += nthr) {
        c3_pnode n;
        n.in_more = n.out_more = C3_NONE; n.rmask = 1; n.spare = 0;
        n.aln0 = n.aln1 = n.aln2 = n.aln3 = C3_NONE; n.max_out = C3_NONE; n.aln_n = 0;
        n.prev = n.next = C3_NONE;
        int pos;
        if (i == C3_SRC) {
            n.base = 4; n.in_n = 0; n.out_n = 1; n.in0 = C3_NONE; n.out0 = 2; n.w0 = 1; pos = 0;
        } else if (i == C3_SINK) {
            n.base = 4; n.in_n = 1; n.out_n = 0; n.in0 = (uint16_t)(L + 1); n.out0 = C3_NONE; n.w0 = 0; pos = L + 1;
        } else {
            n.base = q[i - 2]; n.in_n = 1; n.out_n = 1; n.w0 = 1;
            n.in0 = (uint16_t)(i == 2 ? C3_SRC : i - 1);
            n.out0 = (uint16_t)(i == L + 1 ? C3_SINK : i + 1);
            pos = i - 1;
        }
        W.nodes[i] = n;
        ord[pos] = (uint16_t)i; W.posof[i] = (uint16_t)pos;
        const int hops = L + 1 - pos;
        W.desc[pos] = make_uint4((uint32_t)i | ((uint32_t)(pos == 0 ? C3_NONE : pos - 1) << 16), (uint32_t)C3_NONE | ((uint32_t)hops << 16),
                                 (uint32_t)n.base | ((uint32_t)n.in_n << 8), 0u);
    }
}

// score mode, band half-width, score profile, row descriptors by position (reverse sweep) with the remaining path
// length along the heaviest out-edges (hops to the sink; a successor sits at a higher position: already written)
C3_HD inline void c3s_prepare(c3g_grp &G, const c3_poa_args &A, const c3_poa_para_dev &P, const c3g_ws &W)
{
    const int sq = G.sq;
    const uint8_t *q = G.ibase + G.bnd[2 * sq];
    const int qlen = G.bnd[2 * sq + 1] - G.bnd[2 * sq];
    const int n = G.node_n;
    {
        const int len = qlen > n ? qlen : n;
        const int max_score = max(qlen * 5, len * P.e1 + P.o1);
        const int pn = (max_score <= 32767 - P.mismatch - P.o1 - P.e1) ? P.simd_bits / 16 : P.simd_bits / 32;
        if (qlen <= 0 || qlen > 65000 || qlen + 32 > A.qp_stride || pn != 16 || P.wb < 0) { C3G_DECLINE(); G.err = C3G_E_RETRY; return; }
    }
    G.q = q; G.qlen = qlen; G.n = n; G.w = P.wb + (int)(P.wf * (double)qlen);
    const uint16_t *ord = W.order[G.ob];
    int xbase = 0;
    int prev_id = -1, prev_hops = 0;              // the node handled one step earlier (position p + 1)
    // 8 positions per block: their order entries are one 16-byte piece, their node records are requested together
    // (one memory round trip for the block); the entry before the block comes from the next block's piece
    uint4 ocn = *reinterpret_cast<const uint4 *>(ord + ((n - 1) & ~7));
    for (int pc = (n - 1) & ~7; pc >= 0; pc -= 8) {
        const uint4 oc = ocn;
        if (pc > 0) ocn = *reinterpret_cast<const uint4 *>(ord + pc - 8);
        c3_nrec rk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) if (pc + k < n) rk[k] = c3_ld_node(&W.nodes[c3s_u16_of(oc, k)]);
#pragma unroll
        for (int k = 7; k >= 0; --k) {
        const int p = pc + k;
        if (p >= n) continue;
        const int id = c3s_u16_of(oc, k);
        const c3_nrec nd = rk[k];
        const int idn = k > 0 ? c3s_u16_of(oc, k - 1) : c3s_u16_of(ocn, 7);     // id at position p - 1 (unused when p = 0)
        const int base = C3_N_BASE(nd), in_n = C3_N_INN(nd);
        int p0 = C3_NONE, p1 = C3_NONE;
        const int xo = xbase;
        // (the usual first predecessor is the node before this one in the order, the usual heaviest successor the
        // node after it: no look-up then)
        if (in_n > 0) p0 = (p > 0 && C3_N_IN0(nd) == idn) ? p - 1 : (int)W.posof[C3_N_IN0(nd)];
        if (in_n > 1) {
            c3_pedge pe = W.pool[C3_N_INMORE(nd)];
            p1 = W.posof[pe.id];
            int e = pe.next;
            for (int k = 2; k < in_n && e != (int)C3_NONE; ++k) {
                pe = W.pool[e];
                if (xbase >= A.pool_cap) { C3G_DECLINE(); G.err = C3G_E_RETRY; return; }
                W.xpred[xbase++] = W.posof[pe.id]; e = pe.next;
            }
        }
        int hops = 0;
        if (id != C3_SINK) {
            int best_w = C3_N_W0(nd), best = C3_N_OUT0(nd);
            if (C3_N_OUTN(nd) > 1) {
                int e = W.nodes[id].out_more;
                while (e != (int)C3_NONE) {
                    const c3_pedge pe = W.pool[e];
                    if ((int)pe.w > best_w) { best_w = pe.w; best = pe.id; }
                    e = pe.next;
                }
            }
            hops = (best == prev_id ? prev_hops : C3G_D_HOPS(W.desc[W.posof[best]])) + 1;
        }
        prev_id = id; prev_hops = hops;
        if (in_n > C3_MAXPRE || hops > 65535) { C3G_DECLINE(); G.err = C3G_E_RETRY; return; }
        W.desc[p] = make_uint4((uint32_t)id | ((uint32_t)p0 << 16), (uint32_t)p1 | ((uint32_t)hops << 16),
                               (uint32_t)base | ((uint32_t)in_n << 8) | ((uint32_t)xo << 16), 0u);
        }
    }
}

// ---------------------------------------------------------------------------
// backtrack: abPOA's M -> E1 -> E2 -> F1 -> F2 order and op-mask state machine on the arena's (H, H-E1, H-E2) cells.
// The record, descriptor and cell of the first predecessor -- where the path goes next nine times out of ten -- are
// requested together, so a match/mismatch move costs one memory round trip.  Returns the number of cigar ops or <0.
// ---------------------------------------------------------------------------
#ifndef C3S_BK
#define C3S_BK 4
#endif
#ifndef C3S_MK
#define C3S_MK 4                      // ops per fast block of the merge
#endif
C3_HD inline int c3s_backtrack(const c3g_grp &G, const c3g_args &L, const c3_poa_para_dev &P, const c3g_ws &W, const uint4 *arena)
{
    const int e1 = P.e1, e2 = P.e2, oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const int vs = L.vs_shift;
    const uint8_t *q = G.q; const int qlen = G.qlen, n = G.n;
    const int cap = L.A.cigar_cap;
    unsigned long long *cg = W.cigar;
    const bool eager = L.eager != 0, cs = !eager;       // small wave: eager gap-test loads; large wave: streaming cell loads
    int nc = 0, j, pos, hij;
    {
        const uint4 ds = W.desc[n - 1];
        int best = -0x7fffffff - 1, bj = -1, bk = -1;
        const int skn = C3G_D_NPRE(ds);
        for (int k = 0; k < skn; ++k) {
            const int pk = c3g_pred_pos(W, ds, k);
            const uint2 rp = W.rowrec[pk];
            const int en = min(qlen, C3G_R_END(rp) * 16 + 15);
            const int val = c3g_cell_h(arena, vs, pk, en, cs);
            if (val > best) { best = val; bj = en; bk = pk; }
        }
        if (bk < 0 || qlen - bj + 8 > cap) { C3G_DECLINE(); return C3G_E_RETRY; }
        for (int t = qlen; t > bj; --t)
            cg[qlen - t] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(t - 1) << 32);
        nc = qlen - bj; j = bj; pos = bk; hij = best;
    }
    int cur_op = C3_OP_ALL;
    c3s_q4 qc; c3s_q4_init(qc, q);
    uint4 d = W.desc[pos];
    uint2 rt = W.rowrec[pos];
    while (pos != 0 && j > 0) {
        // Fast block: the path mostly runs (pos, j) -> (pos - 1, j - 1) by match/mismatch moves through first
        // predecessors that sit one position back.  The addresses of the next C3S_BK such steps do not depend on any
        // value, so everything they need is requested at once (one memory round trip instead of C3S_BK); each step is
        // then accepted only if the generic step below would have made exactly this move (M is tested first, the
        // first predecessor first), and the block stops at the first step that is anything else.
        int fast = 0;
        if (cur_op & C3_OP_M) {
            uint4 dk[C3S_BK]; uint2 rk[C3S_BK]; int hk[C3S_BK], qk[C3S_BK];
#pragma unroll
            for (int k = 0; k < C3S_BK; ++k) {
                const int pp = max(pos - 1 - k, 0), jj = max(j - 1 - k, 0);
                dk[k] = W.desc[pp]; rk[k] = W.rowrec[pp];
                hk[k] = c3g_cell_h(arena, vs, pp, jj, cs); qk[k] = c3s_q4_get(qc, jj);
            }
            bool ok = true;
#pragma unroll
            for (int k = 0; k < C3S_BK; ++k) {
                if (ok) {
                    const int b = C3G_R_BEG(rt) * 16, en = min(qlen, C3G_R_END(rt) * 16 + 15);
                    const int pbeg = C3G_R_BEG(rk[k]) * 16, pend = min(qlen, C3G_R_END(rk[k]) * 16 + 15);
                    ok = pos != 0 && j > 0 && C3G_D_NPRE(d) > 0 && C3G_D_P0(d) == pos - 1 && j >= b && j <= en &&
                         j - 1 >= max(pbeg, b) && j - 1 <= pend &&
                         hk[k] + c3_score(P, C3G_D_BASE(d), qk[k]) == hij && nc + j + 8 <= cap;
                    if (ok) {
                        cg[nc++] = C3_CG_MATCH | ((unsigned long long)C3G_D_ID(d) << 8) | ((unsigned long long)(j - 1) << 32);
                        --pos; --j; hij = hk[k]; rt = rk[k]; d = dk[k]; cur_op = C3_OP_ALL; ++fast;
                    }
                }
            }
        }
        if (fast == C3S_BK || pos == 0 || j <= 0) continue;
        const int i = C3G_D_ID(d);
        const int b = C3G_R_BEG(rt) * 16, en = min(qlen, C3G_R_END(rt) * 16 + 15);
        if (j < b || j > en) { C3G_DECLINE(); return C3G_E_RETRY; }
        const int s = c3_score(P, C3G_D_BASE(d), c3s_q4_get(qc, j - 1));
        const int npre = C3G_D_NPRE(d);
        int hit = 0;
        unsigned long long opw = 0;
        // first predecessor: record, descriptor and both cells the tests can ask for, all requested at once
        const int p0 = C3G_D_P0(d);
        const uint2 r0 = W.rowrec[p0];
        const uint4 d0 = W.desc[p0];
        const int ph0 = c3g_cell_h(arena, vs, p0, j - 1, cs);           // (inside the read's arena whatever the band is)
        // ... and what the gap tests of this cell would ask for next (its own E byte, the first predecessor's cell above
        // it), and the row's H vectors the F test rebuilds F from: all addresses are known now, so a deletion or an
        // insertion costs one more round trip instead of two or three
        // (eager: waves small enough to be bound by the chain of round trips, not by DRAM transactions -- at 100 000 reads
        // the extra sectors cost 3 %, at 12 500 the saved round trips gain 6 %)
        int ceb0 = 0, ph0j = 0, peb0j = 0;
        if (eager) {
            ceb0 = c3g_cell_eb(arena, vs, pos, j, cs);
            ph0j = c3g_cell_h(arena, vs, p0, j, cs); peb0j = c3g_cell_eb(arena, vs, p0, j, cs);
            if (cur_op & C3_OP_F)
                for (int c0 = b; c0 < j; c0 += 16) C3L_PREFETCH(arena + (((int64_t)pos << vs) + ((c0 >> 4) & ((1 << vs) - 1))) * 3);
        }
        if (cur_op & C3_OP_M) {
            for (int k = 0; k < npre; ++k) {
                const int pk = k == 0 ? p0 : c3g_pred_pos(W, d, k);
                const uint2 pr = k == 0 ? r0 : W.rowrec[pk];
                const int pbeg = C3G_R_BEG(pr) * 16, pend = min(qlen, C3G_R_END(pr) * 16 + 15);
                if (j - 1 < max(pbeg, b) || j - 1 > pend) continue;
                const int ph = k == 0 ? ph0 : c3g_cell_h(arena, vs, pk, j - 1, cs);
                if (ph + s == hij) {
                    opw = C3_CG_MATCH | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                    pos = pk; --j; hit = 1; cur_op = C3_OP_ALL; hij = ph;
                    rt = pr; d = k == 0 ? d0 : W.desc[pk];
                    break;
                }
            }
        }
        if (!hit && (cur_op & C3_OP_E)) {
            const int ceb = eager ? ceb0 : c3g_cell_eb(arena, vs, pos, j, cs);
            const int ce1 = hij - (ceb & 7), ce2 = hij - (ceb >> 3);
            for (int k = 0; k < npre; ++k) {
                const int pk = k == 0 ? p0 : c3g_pred_pos(W, d, k);
                const uint2 pr = k == 0 ? r0 : W.rowrec[pk];
                const int pbeg = C3G_R_BEG(pr) * 16, pend = min(qlen, C3G_R_END(pr) * 16 + 15);
                if (j < pbeg || j > pend) continue;
                const int ph = (k == 0 && eager) ? ph0j : c3g_cell_h(arena, vs, pk, j, cs);
                int pe1, pe2;
                if (pk == 0) { pe1 = j == 0 ? -oe1 : C3_NEG_INF; pe2 = j == 0 ? -oe2 : C3_NEG_INF; }
                else { const int peb = (k == 0 && eager) ? peb0j : c3g_cell_eb(arena, vs, pk, j, cs); pe1 = ph - (peb & 7); pe2 = ph - (peb >> 3); }
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
                    pos = pk; hij = ph;
                    rt = pr; d = k == 0 ? d0 : W.desc[pk];
                    break;
                }
            }
        }
        if (!hit && (cur_op & C3_OP_F)) {
            int hl = C3_NEG_INF;
            if (j - 1 >= b) {
                // F is not stored: rebuild F[j] and F[j-1] of this row from its H, 16 columns per load
                // (F[c+1] = max(F[c]-e, H[c]-o-e); equal to the DP's F wherever F decides)
                int f1 = C3_NEG_INF, f2 = C3_NEG_INF, f1l = C3_NEG_INF, f2l = C3_NEG_INF;
                for (int c0 = b; c0 < j; c0 += 16) {
                    const uint4 *hp = arena + (((int64_t)pos << vs) + ((c0 >> 4) & ((1 << vs) - 1))) * 3;
                    int t16[16];
                    { int t8[8]; c3l_unpack8(cs ? C3G_LDCS(hp) : hp[0], t8);
#pragma unroll
                      for (int k = 0; k < 8; ++k) t16[k] = t8[k];
                      c3l_unpack8(cs ? C3G_LDCS(hp + 1) : hp[1], t8);
#pragma unroll
                      for (int k = 0; k < 8; ++k) t16[8 + k] = t8[k]; }
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        if (c0 + k < j) {
                            hl = c3l_map(t16[k]);
                            f1l = f1; f2l = f2;
                            f1 = max(f1 - e1, hl - oe1); f2 = max(f2 - e2, hl - oe2);
                        }
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
            }
            if (hit) { opw = C3_CG_INS | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32); --j; hij = hl; }
        }
        if (!hit) { C3G_DECLINE(); return C3G_E_RETRY; }
        cg[nc] = opw;
        ++nc;
        if (nc + j + 8 > cap) { C3G_DECLINE(); return C3G_E_RETRY; }
    }
    for (int t = j; t > 0; --t)
        cg[nc + j - t] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(t - 1) << 32);
    nc += j;
    return nc;
}

// ---------------------------------------------------------------------------
// merge (abpoa_add_graph_alignment): the cigar is walked from its tail = forward order.  A new node is not linked
// into a list: its place in the order is the gap of the OLD order it falls into (W.gaps, non-decreasing in creation
// order), see c3s_reorder.
// ---------------------------------------------------------------------------
C3_HD inline int c3s_tail_gap(const uint16_t *ord, const int n_old, const c3_pnode &na, const int p_start)
{
    int t = p_start;
    for (;;) {
        if (t + 1 >= n_old) break;
        const int nx = ord[t + 1];
        bool in_group = false;
        for (int k = 0; k < na.aln_n; ++k) in_group |= (c3_aln_get(na, k) == nx);
        if (!in_group) break;
        ++t;
    }
    return t + 1;
}

C3_HD inline int c3s_merge(c3g_grp &G, const c3g_args &L, const c3g_ws &W, const int nc)
{
    const uint8_t *q = G.q;
    const unsigned long long *cg = W.cigar;
    const uint16_t *ord = W.order[G.ob];
    const int n_old = G.n;
    c3_graph g; g.nodes = W.nodes; g.pool = W.pool; g.node_n = G.node_n; g.pool_n = G.pool_n;
    g.node_cap = L.A.node_cap; g.pool_cap = L.A.pool_cap; g.err = 0;
    int last_id = C3_SRC, last_new = 0;
    // items whose two MSA rows are wanted (c3s_emit_msa) record which sequence passes through which node (rmask, as
    // poa.cuh): they take the generic code for every op
    const bool track = L.A.msa2 && G.nseq == 2;
    const uint16_t rbit = (uint16_t)(1u << (G.sq < 16 ? G.sq : 15));
    int last_gap = 0, last_p = 0;                // gap of the last new node / old position its group walk starts at
    c3s_q4 qc; c3s_q4_init(qc, q);
    int t = nc - 1;
    while (t >= 0 && !g.err) {
        // Fast block: most ops are deletions (nothing to do) or matches onto a node with the read's base whose edge from
        // the previous node already exists as that node's first out-edge (one weight to bump).  The next C3S_MK ops and
        // the hot halves of their node records are requested together -- two memory round trips for the block instead
        // of two per op -- and consumed while they are of that kind; the first op that is anything else is left to the
        // generic code below, which reads everything afresh.
        if (!track) {
            unsigned long long ok_[C3S_MK]; c3_nrec rk[C3S_MK];
#pragma unroll
            for (int k = 0; k < C3S_MK; ++k) ok_[k] = t - k >= 0 ? cg[t - k] : C3_CG_DEL;
            c3_nrec prev = c3_ld_node(&g.nodes[last_id]);
#pragma unroll
            for (int k = 0; k < C3S_MK; ++k) {
                const int nn = (int)((ok_[k] >> 8) & 0xffff);
                rk[k] = c3_ld_node(&g.nodes[(ok_[k] & 0xff) == C3_CG_MATCH ? nn : 0]);
            }
            bool ok = true;
            int used = 0;
#pragma unroll
            for (int k = 0; k < C3S_MK; ++k) {
                if (ok && t - k >= 0) {
                    const int kc = (int)(ok_[k] & 0xff), nid = (int)((ok_[k] >> 8) & 0xffff), qp = (int)(ok_[k] >> 32);
                    if (kc == (int)C3_CG_DEL) ++used;
                    else if (kc == (int)C3_CG_MATCH && !last_new && C3_N_BASE(rk[k]) == c3s_q4_get(qc, qp) &&
                             C3_N_OUTN(prev) > 0 && C3_N_OUT0(prev) == nid) {
                        g.nodes[last_id].w0 = (uint16_t)(C3_N_W0(prev) + 1);
                        last_id = nid; prev = rk[k]; ++used;
                    } else ok = false;
                }
            }
            t -= used;
            if (ok) continue;
        }
        const unsigned long long opc = cg[t];
        --t;
        const int kc = (int)(opc & 0xff), nid = (int)((opc >> 8) & 0xffff), qp = (int)(opc >> 32);
        if (kc == (int)C3_CG_DEL) continue;
        if (kc == (int)C3_CG_MATCH) {
            const uint8_t bq = (uint8_t)c3s_q4_get(qc, qp);
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
                    if (track) g.nodes[al].rmask |= rbit;
                } else {
                    const int id = c3_g_add_node(g, bq);
                    if (g.err) break;
                    if (track) g.nodes[id].rmask = rbit;
                    last_p = W.posof[nid]; last_gap = last_p;          // placed right before nid
                    W.gaps[id - n_old] = (uint16_t)last_gap;
                    c3_g_add_edge(g, last_id, id, 0);
                    last_id = id; last_new = 1;
                    for (int k = 0; k < nm.aln_n; ++k) {               // abpoa_add_graph_aligned_node
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
                if (track) g.nodes[nid].rmask |= rbit;
            }
        } else {                                                     // insertion: right after the aligned block of last_id
            const int id = c3_g_add_node(g, (uint8_t)c3s_q4_get(qc, qp));
            if (g.err) break;
            if (track) g.nodes[id].rmask = rbit;
            const c3_pnode nl = g.nodes[last_id];
            int gap;
            if (last_id >= n_old) gap = nl.aln_n ? c3s_tail_gap(ord, n_old, nl, last_p) : last_gap;
            else gap = c3s_tail_gap(ord, n_old, nl, W.posof[last_id]);
            last_gap = gap;
            W.gaps[id - n_old] = (uint16_t)gap;
            c3_g_add_edge(g, last_id, id, 0);
            last_id = id; last_new = 1;
        }
    }
    if (!g.err) c3_g_add_edge(g, last_id, C3_SINK, 1 - last_new);
    if (g.err) { C3G_DECLINE(); return C3G_E_RETRY; }
    G.node_n = g.node_n; G.pool_n = g.pool_n;
    return 0;
}

// new order = stable merge of the old order with the new nodes by gap: the t-th new node (id n_old + t, gap g)
// lands at g + t, an old node at position p moves up by the number of new nodes with gap <= p
C3_HD inline void c3s_reorder(c3g_grp &G, const c3g_ws &W)
{
    const int n_old = G.n, m = G.node_n - n_old;
    const uint16_t *oo = W.order[G.ob];
    c3s_out8 on; on.dst = W.order[G.ob ^ 1]; on.lo = on.hi = 0ull; on.n = 0;
    int t = 0;
    int gt = m > 0 ? (int)W.gaps[0] : 0x7fffffff;
    for (int p0 = 0; p0 < n_old; p0 += 8) {
        const uint4 oc = *reinterpret_cast<const uint4 *>(oo + p0);          // (the arrays are padded to 32 entries)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int p = p0 + k;
            if (p < n_old) {
                while (gt <= p) {
                    W.posof[n_old + t] = (uint16_t)on.n; c3s_out8_push(on, n_old + t); ++t;
                    gt = t < m ? (int)W.gaps[t] : 0x7fffffff;
                }
                const int id = c3s_u16_of(oc, k);
                W.posof[id] = (uint16_t)on.n; c3s_out8_push(on, id);
            }
        }
    }
    c3s_out8_flush(on);
    G.ob ^= 1;
}

// heaviest bundling (abpoa_heaviest_bundling) + consensus walk.  Returns the length or a negative code.
C3_HD inline int c3s_consensus(const c3g_grp &G, const c3_poa_args &A, const c3g_ws &W, char *co)
{
    int32_t *score = reinterpret_cast<int32_t *>(W.desc);
    const uint16_t *ord = W.order[G.ob];
    for (int p = G.node_n - 1; p >= 0; --p) {
        const int v = ord[p];
        c3_pnode *nd = &W.nodes[v];
        if (p >= C3S_PF) C3S_PREFETCH(&W.nodes[ord[p - C3S_PF]]);
        if ((p & 15) == 0 && p >= C3S_PF + 80) C3S_PREFETCH(&ord[p - C3S_PF - 80]);
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
    }
    int cons_len = 0;
    int id = W.nodes[C3_SRC].max_out;
    while (id != C3_SINK) {
        if (id == C3_NONE || cons_len >= A.cons_cap) { C3G_DECLINE(); return C3G_E_RETRY; }
        const c3_pnode nd = W.nodes[id];
        co[cons_len++] = "ACGTN"[nd.base];
        id = nd.max_out;
    }
    return cons_len;
}

C3_HD inline int c3s_consensus_small(const c3g_grp &G, const c3_poa_args &A, const c3g_ws &W, char *co)
{
    // The variant for small waves (bound by the chain of round trips; at 100 000 reads its two extra streams cost 4 %).
    // Reverse sweep in blocks of 8 positions like c3s_prepare (order entries as one 16-byte piece, node records requested
    // together).  The heaviest successor and the base of every node go to two dense arrays that are free by now (posof,
    // gaps: 2 bytes per node), so the node records stay clean and the walk along the consensus path reads 16 nodes per
    // sector instead of chasing 32-byte records through DRAM one round trip at a time.
    int32_t *score = reinterpret_cast<int32_t *>(W.desc);
    const uint16_t *ord = W.order[G.ob];
    uint16_t *nxt = W.posof, *nbase = W.gaps;
    const int n = G.node_n;
    for (int pc = (n - 1) & ~7; pc >= 0; pc -= 8) {
        const uint4 oc = *reinterpret_cast<const uint4 *>(ord + pc);
        c3_nrec rk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) if (pc + k < n) rk[k] = c3_ld_node(&W.nodes[c3s_u16_of(oc, k)]);
#pragma unroll
        for (int k = 7; k >= 0; --k) {
            if (pc + k >= n) continue;
            const int v = c3s_u16_of(oc, k);
            const c3_nrec nd = rk[k];
            const int out_n = C3_N_OUTN(nd);
            int max_id = -1;
            if (v == C3_SINK) score[v] = 0;
            else if (v == C3_SRC) {
                int path_score = -1, path_w = -1;
                uint16_t e = out_n > 1 ? W.nodes[v].out_more : (uint16_t)C3_NONE;
                for (int t = 0; t < out_n; ++t) {
                    int o, wv;
                    if (t == 0) { o = C3_N_OUT0(nd); wv = C3_N_W0(nd); } else { const c3_pedge pe = W.pool[e]; o = pe.id; wv = pe.w; e = pe.next; }
                    if (wv > path_w || (wv == path_w && score[o] > path_score)) { max_id = o; path_score = score[o]; path_w = wv; }
                }
            } else {
                int max_w = -0x7fffffff - 1;
                uint16_t e = out_n > 1 ? W.nodes[v].out_more : (uint16_t)C3_NONE;
                for (int t = 0; t < out_n; ++t) {
                    int o, wv;
                    if (t == 0) { o = C3_N_OUT0(nd); wv = C3_N_W0(nd); } else { const c3_pedge pe = W.pool[e]; o = pe.id; wv = pe.w; e = pe.next; }
                    if (max_w < wv) { max_w = wv; max_id = o; }
                    else if (max_w == wv && score[max_id] <= score[o]) max_id = o;
                }
                score[v] = max_w + score[max_id];
            }
            nxt[v] = (uint16_t)max_id;                              // (-1 -> C3_NONE)
            nbase[v] = (uint16_t)C3_N_BASE(nd);
        }
    }
    int cons_len = 0;
    int id = nxt[C3_SRC];
    while (id != C3_SINK) {
        if (id == C3_NONE || cons_len >= A.cons_cap) { C3G_DECLINE(); return C3G_E_RETRY; }
        co[cons_len++] = "ACGTN"[nbase[id]];
        id = nxt[id];
    }
    return cons_len;
}

// The two MSA rows of a 2-sequence item (abpoa_generate_rc_msa: ranks by the LIFO traversal in which a node is pushed once
// its in-edges and those of its aligned nodes are used up; same code as poa.cuh's c3_emit_msa, one thread).  Rows go to
// co as [row0 | row1]; returns the number of columns or a negative code.  rank / in-degree / stack live in the
// descriptor array (no alignment follows).
C3_HD inline int c3s_emit_msa(const c3g_grp &G, const c3_poa_args &A, const c3g_ws &W, char *co)
{
    const int node_n = G.node_n;
    int32_t *rank = reinterpret_cast<int32_t *>(W.desc), *indeg = rank + A.node_cap, *stk = indeg + A.node_cap;
    for (int v = 0; v < node_n; ++v) { rank[v] = 0; indeg[v] = W.nodes[v].in_n; }
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
                if (top + 5 > A.node_cap) { C3G_DECLINE(); return C3G_E_RETRY; }
                stk[top++] = o; rank[o] = -1;
                for (int a = 0; a < on.aln_n; ++a) { const int al = c3_aln_get(on, a); stk[top++] = al; rank[al] = -1; }
            }
        }
    }
    const int msa_len = ok ? rank[C3_SINK] - 1 : -1;
    if (msa_len < 0 || 2 * msa_len > A.cons_cap) { C3G_DECLINE(); return C3G_E_RETRY; }
    for (int c = 0; c < 2 * msa_len; ++c) co[c] = '-';
    for (int v = 2; v < node_n; ++v) {
        const c3_pnode nd = W.nodes[v];
        int rk = rank[v];
        for (int k = 0; k < nd.aln_n; ++k) rk = max(rk, rank[c3_aln_get(nd, k)]);
        const char ch = "ACGTN"[nd.base];
        if (nd.rmask & 1) co[rk - 1] = ch;
        if (nd.rmask & 2) co[msa_len + rk - 1] = ch;
    }
    return msa_len;
}

// last sequence merged: heaviest bundling, consensus (or the two MSA rows) and the per-read outputs
C3_HD inline void c3s_finish(c3g_grp &G, const c3g_args &L, const c3g_ws &W)
{
    const c3_poa_args &A = L.A;
    char *co = A.cons + (int64_t)G.item * A.cons_cap;
    const bool do_msa = A.msa2 && G.nseq == 2;
    const int r = do_msa ? c3s_emit_msa(G, A, W, co) : L.eager ? c3s_consensus_small(G, A, W, co) : c3s_consensus(G, A, W, co);
    if (r >= 0) {
        const int64_t o = (int64_t)G.item * A.out_stride;
        A.status[o] = do_msa ? A.ok_status : 0;
        A.cons_len[o] = r;
        A.nodes_out[o] = G.node_n;
        *(long long *)((int32_t *)A.cells_out + (int64_t)G.item * A.cells_stride) = G.cells_total;
        L.done[G.item] = 1;
    } else G.err = C3G_E_RETRY;
}

// ---------------------------------------------------------------------------
// one read's turn between two DP launches.  The 32 threads of a warp are brought
// back together after every phase (C3S_SYNC): the phases are loops of data-dependent length with early exits, and
// without it a thread that leaves one loop early runs all later phases alone.
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define C3S_SYNC() __syncwarp()
#else
#define C3S_SYNC() do { } while (0)
#endif
C3_HD inline void c3s_graph_step(const c3g_args &L, const int it, const bool valid)
{
    const c3_poa_args &A = L.A;
    const c3_poa_para_dev P = A.P;
    c3g_state *S = L.state + (valid ? it : 0);
    const c3g_ws W = c3g_ws_carve(L.ws + (int64_t)(valid ? it : 0) * L.ws_stride, A.node_cap, A.pool_cap, A.cigar_cap);
    const uint4 *arena = L.arena + (int64_t)(valid ? it : 0) * L.arena_stride4;
    c3g_grp G;
    G.err = 0; G.sq = 0; G.nseq = 0;
    bool run = valid;
#if defined(C3L_PROF) && defined(__CUDA_ARCH__)
    const int lane = threadIdx.x & 31;
#endif
    C3L_TICK_INIT;
    {
        if (run) { c3g_state_load(G, S, A); if (G.err || G.sq >= G.nseq) run = false; }   // declined earlier / finished earlier
        int nc = 0;
        C3S_SYNC();
        C3L_TICK(6);
        if (run) { nc = c3s_backtrack(G, L, P, W, arena); if (nc < 0) G.err = C3G_E_RETRY; }
        C3S_SYNC();
        C3L_TICK(7);
        if (run && !G.err && c3s_merge(G, L, W, nc)) G.err = C3G_E_RETRY;
        C3S_SYNC();
        C3L_TICK(8);
        if (run && !G.err) { c3s_reorder(G, W); ++G.sq; }
    }
    C3S_SYNC();
    C3L_TICK(9);
    if (run && !G.err && G.sq < G.nseq) c3s_prepare(G, A, P, W);
    C3S_SYNC();
    C3L_TICK(10);
    if (run && !G.err && G.sq >= G.nseq) c3s_finish(G, L, W);
    C3S_SYNC();
    C3L_TICK(11);
    if (run) c3g_state_store(G, S);
}

#ifdef __CUDACC__
#ifndef C3S_THREADS
#define C3S_THREADS 64
#endif
#ifndef C3S_MINB
#define C3S_MINB 12                    // 768 threads per SM: a 100k-read wave is resident at once
#endif
__global__ void __launch_bounds__(C3S_THREADS, C3S_MINB) c3_poa_graph_kernel(c3g_args L)
{
    const int it = blockIdx.x * C3S_THREADS + threadIdx.x;
    c3s_graph_step(L, it, it < L.A.n_work);
}
// first launch of a wave: one CTA per read writes the first sequence's linear graph (coalesced stores); a read of a
// single sequence has nothing to align and is finished here
__global__ void __launch_bounds__(128) c3_poa_graph_init_kernel(c3g_args L)
{
    const int it = blockIdx.x;
    const c3_poa_args &A = L.A;
    const c3g_ws W = c3g_ws_carve(L.ws + (int64_t)it * L.ws_stride, A.node_cap, A.pool_cap, A.cigar_cap);
    c3g_grp G;
    c3s_item_begin(G, A, A.P, W, A.order ? A.order[it] : it, threadIdx.x, 128);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!G.err && G.sq >= G.nseq) c3s_finish(G, L, W);
        c3g_state_store(G, L.state + it);
    }
}
#endif
