// poa_grp.cuh -- stage 3b, group kernel: 8 lanes per read, 4 reads per warp.
//
// Same algorithm and the same results as poa.cuh / poa_lane.cuh (abPOA 1.0.5 semantics; replaces
// poa.msa_aligner(match=5).msa(...), /root/reference/bin/determine_consensus.py:30-47), third mapping:
//
//   poa.cuh       one warp per read: half the lanes idle on a ~70-column band and every row pays the warp-uniform
//                 bookkeeping 32 times over.
//   poa_lane.cuh  one thread per read: 240 KB of thread-private graph state per read and 6 B per DP cell; nothing
//                 is cacheable, every pointer-chasing step costs a DRAM round trip.
//   poa_grp.cuh   one 8-lane group per read.  Everything a read owns is laid out BY POSITION in the topological
//                 order, so all accesses of the DP are sequential and coalesced:
//     * order[] / posof[] arrays instead of a linked list; after a merge the new nodes (already sorted by the
//       gap of the old order they fall into) are merged into the order in parallel;
//     * `prepare` turns the graph into one 16-byte row descriptor per position (node id, base, positions of the
//       first two predecessors, remaining path length along the heaviest out-edges);
//     * DP: each lane owns one 16-column vector of the band (abPOA's int16 SIMD granule), lane = vector index
//       mod 8, cells packed two per register (VIADDMNMX.S16x2 / VIMNMX3.S16x2); the horizontal gap is an
//       in-lane chain plus a 3-step (max,+) scan over the 8 lanes; predecessor rows come from a 4-row
//       shared-memory ring (H, E1, E2 as int16) or, further back, from the arena;
//     * arena: 3 bytes per cell -- H as int16 plus one byte holding H-E1 (3 bits) and H-E2 (5 bits), stored
//       complemented (that is what one packed add of ~H yields); both
//       differences are bounded by the gap-open costs, so H, E1 and E2 are recovered exactly and the backtrack
//       stays abPOA's value-based one (M -> E1 -> E2 -> F1 -> F2, op-mask state machine).  Rows have a fixed
//       stride of VS vectors, vector v of a row at slot v mod VS: a cell's address needs no row record;
//     * backtrack: runs of match/mismatch moves along consecutive rows are verified 7 at a time.
//
// Scope: int16 score mode with the 256-bit granule, banded, default-sized gap costs (o1+e1 <= 7, o2+e2 <= 31),
// consensus output.  Anything else -- and any capacity overflow or a row that fails the int16 exactness guard --
// leaves the item not-done; the caller then runs it through c3_poa_kernel, which owns all error reporting.
//
// With -DC3G_EMUL the kernel body is host code as well: tests/emul/grp_emul.cu runs it on the fiber warp
// emulator (tests/emul/warp_emu.cpp) against the oracle, no GPU needed.
#pragma once
#include "poa_lane.cuh"

#define C3G_R 4                       // ring slots (rows) per group
#define C3G_GL 8                      // lanes per group
#define C3G_E_RETRY (-298)
// graph kernel: GL lanes per read (32: one warp per read, constant member masks; 8: four reads per warp)
#define C3G_HWIN(GL) ((GL) == 32 ? 1024 : 512)   // window of remaining-length words kept in shared memory during prepare
#define C3G_BTK(GL) ((GL) == 32 ? 64 : 16)       // rows per backtrack window
#define C3G_BTV(GL) ((GL) == 32 ? 31 : 7)        // rows below the current one that a verification trip wants in the window
#define C3G_GRAPH_SMEM(GL) ((GL) == 32 ? 7680 : 2048)   // shared memory per read of the graph kernel: max(HWIN * 2, BTK * 120)
#define C3G_LMASK(GL) ((GL) == 32 ? 0xffffffffu : ((1u << ((GL) & 31)) - 1u))
#define C3G_LOG2(GL) ((GL) == 32 ? 5 : 3)

#if defined(C3G_EMUL) && !defined(__CUDA_ARCH__)
extern "C" int c3emu_shfl(unsigned mask, int v, int src, int tag);
extern "C" unsigned c3emu_ballot(unsigned mask, int pred, int tag);
extern "C" void c3emu_sync(unsigned mask, int tag);
#define C3G_SHFL(m, v, s) c3emu_shfl((m), (int)(v), (s), __LINE__)
#define C3G_BALLOT(m, p) c3emu_ballot((m), (p), __LINE__)
#define C3G_SYNC(m) c3emu_sync((m), __LINE__)
#define C3G_ATOMIC_INC(p) ((*(p))++)
#define C3G_FFS(x) __builtin_ffs((int)(x))
#define C3G_CLZ(x) ((x) ? __builtin_clz((unsigned)(x)) : 32)
#else
#define C3G_SHFL(m, v, s) __shfl_sync((m), (int)(v), (s))
#define C3G_BALLOT(m, p) __ballot_sync((m), (p))
#define C3G_SYNC(m) __syncwarp(m)
#define C3G_ATOMIC_INC(p) atomicAdd((p), 1u)
#define C3G_FFS(x) __ffs((int)(x))
#define C3G_CLZ(x) __clz((int)(x))
#endif
#ifdef C3G_EMUL
#define C3G_FN __host__ __device__ inline
#else
#define C3G_FN __device__ __forceinline__
#endif
#define C3G_ANYG(m, p) (C3G_BALLOT((m), (p)) != 0u)
#if defined(__CUDA_ARCH__)
#define C3G_VADD2(a, b) __vadd2((a), (b))            // per-halfword add, no carry between the halves: VIADD.16x2
#else
#define C3G_VADD2(a, b) ((uint32_t)(((((uint32_t)(a)) & 0xffffu) + (((uint32_t)(b)) & 0xffffu)) & 0xffffu) | (uint32_t)(((((uint32_t)(a)) >> 16) + (((uint32_t)(b)) >> 16)) << 16))
#endif
#if defined(C3G_EMUL) && !defined(__CUDA_ARCH__)
extern "C" void c3g_emul_note(int line);            // test hook: why an item was declined
#define C3G_DECLINE() c3g_emul_note(__LINE__)
#else
#define C3G_DECLINE() do { } while (0)
#endif

struct c3g_state;
struct c3g_args {
    c3_poa_args A;                    // inputs / outputs / parameters; order + n_work: items this kernel covers;
                                      // node_cap, pool_cap, cigar_cap, qp_stride: per-group capacities
    uint8_t *ws; long long ws_stride; // per-read graph workspace
    uint4 *arena; long long arena_stride4;   // per-read DP arena, in uint4: node_cap rows x VS vectors x 3
    int vs_shift;                     // log2 VS: vectors per arena row
    int rv_shift;                     // log2 RV: vectors per shared-memory ring slot (RV <= VS)
    int32_t *done;                    // [n_items] 1 = finished here
    struct c3g_state *state;          // per read of the wave
    int first;                        // graph kernel: 1 = first launch of a wave (first sequence -> graph)
};

// per-group workspace
struct c3g_ws {
    c3_pnode *nodes; c3_pedge *pool;
    uint4 *desc;                      // by position: x = id | p0 << 16, y = p1 | hops << 16, z = base | npre << 8 | xofs << 16
    uint2 *rowrec;                    // by position: x = beg_sn | end_sn << 16, y = mp (arg-max column + 1)
    uint16_t *order[2];               // position -> node id (double buffered across merges)
    uint16_t *posof;                  // node id -> position
    uint16_t *xpred;                  // positions of the third and further predecessors (in-edge order)
    uint16_t *gaps;                   // new nodes of the running merge: old position they are inserted before
    unsigned long long *cigar;
    int8_t *qp;                       // substitution scores: 5 rows (node base; N: zeros) x qp_stride, index = column
};

__host__ __device__ inline int64_t c3g_ws_bytes(int node_cap, int pool_cap, int cigar_cap, int qp_stride)
{
    int64_t b = 0;
    b += (int64_t)node_cap * 32 + (int64_t)pool_cap * 8 + (int64_t)node_cap * 16 + (int64_t)node_cap * 8;
    b += (int64_t)node_cap * 2 * 2 + (int64_t)node_cap * 2 + (int64_t)pool_cap * 2 + (int64_t)node_cap * 2;
    b += (int64_t)cigar_cap * 8 + (int64_t)qp_stride * 5 + 256;    // + 256: lanes past the band read scores beyond qlen
    return (b + 255) & ~(int64_t)255;
}

C3_HD __forceinline__ c3g_ws c3g_ws_carve(uint8_t *p, int node_cap, int pool_cap, int cigar_cap)
{
    c3g_ws w;
    w.nodes = (c3_pnode *)p; p += (int64_t)node_cap * 32;
    w.desc = (uint4 *)p; p += (int64_t)node_cap * 16;
    w.pool = (c3_pedge *)p; p += (int64_t)pool_cap * 8;
    w.rowrec = (uint2 *)p; p += (int64_t)node_cap * 8;
    w.cigar = (unsigned long long *)p; p += (int64_t)cigar_cap * 8;
    w.order[0] = (uint16_t *)p; p += (int64_t)node_cap * 2;
    w.order[1] = (uint16_t *)p; p += (int64_t)node_cap * 2;
    w.posof = (uint16_t *)p; p += (int64_t)node_cap * 2;
    w.gaps = (uint16_t *)p; p += (int64_t)node_cap * 2;
    w.xpred = (uint16_t *)p; p += (int64_t)pool_cap * 2;
    w.qp = (int8_t *)p;
    return w;
}

// shared memory per group: ring cells, then C3G_R row records
C3_HD __forceinline__ int c3g_smem_group_bytes(int rv_shift) { return (C3G_R * 6 * 16 << rv_shift) + C3G_R * 8 + 32; }

struct c3g_grp {                       // group-uniform state (replicated in the 8 lanes)
    int item, sq, nseq, node_n, pool_n, err, ob;
    long long cells_total;
    const uint8_t *ibase; const int32_t *bnd;
    const uint8_t *q; int qlen, n, w;
};

#define C3G_D_ID(d) ((int)((d).x & 0xffffu))
#define C3G_D_P0(d) ((int)((d).x >> 16))
#define C3G_D_P1(d) ((int)((d).y & 0xffffu))
#define C3G_D_HOPS(d) ((int)((d).y >> 16))
#define C3G_D_BASE(d) ((int)((d).z & 0xffu))
#define C3G_D_NPRE(d) ((int)(((d).z >> 8) & 0xffu))
#define C3G_D_XOFS(d) ((int)((d).z >> 16))
#define C3G_R_BEG(r) ((int)((r).x & 0xffffu))
#define C3G_R_END(r) ((int)((r).x >> 16))
#define C3G_R_MP(r) ((int)((r).y & 0xffffu))

// ---------------------------------------------------------------------------
// item start: first sequence -> linear graph, order = SRC, 2, 3, ..., L+1, SINK
// ---------------------------------------------------------------------------
template <int GL>
C3G_FN void c3g_item_begin(c3g_grp &G, const c3_poa_args &A, const c3g_ws &W, const int item, const int li)
{
    G.item = item; G.sq = 1; G.err = 0; G.nseq = 0; G.node_n = 0; G.pool_n = 0; G.cells_total = 0; G.ob = 0;
    const int nseq = A.n_seqs[(int64_t)item * A.n_seqs_stride];
    if (nseq < A.min_seqs || nseq > A.max_seqs || nseq < 1 || (A.msa2 && nseq == 2)) { C3G_DECLINE(); G.err = C3G_E_RETRY; return; }
    G.ibase = A.codes + A.item_base[item];
    G.bnd = A.bounds + (int64_t)item * A.max_seqs * 2;
    G.nseq = nseq;
    const uint8_t *q = G.ibase + G.bnd[0];
    const int L = G.bnd[1] - G.bnd[0];
    if (L <= 0 || L > 65000 || L + 2 > A.node_cap) { C3G_DECLINE(); G.err = C3G_E_RETRY; return; }
    uint16_t *ord = W.order[0];
    for (int i = li; i < L + 2; i += GL) {
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
    }
    G.node_n = L + 2;
}

// ---------------------------------------------------------------------------
// prepare: score mode, band half-width, score profile, row descriptors by position (reverse sweep, 8 positions
// per step) with the remaining path length along the heaviest out-edges.  hw: shared-memory window of hop counts.
// ---------------------------------------------------------------------------
template <int GL>
C3G_FN void c3g_prepare(c3g_grp &G, const c3_poa_args &A, const c3_poa_para_dev &P, const c3g_ws &W, uint16_t *hw,
                        const int li, const int gbase, const unsigned gmask)
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
    // profile: qp[b][j] = score of node base b against column j (= q[j-1]); j = 0 and the padding score 0
    {
        const int qs = A.qp_stride;
        for (int j0 = 4 * li; j0 < qs; j0 += 4 * GL) {
            uint32_t wv[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = j0 + k;
                if (j >= 1 && j <= qlen) {
                    const int qc = q[j - 1];
#pragma unroll
                    for (int b = 0; b < 4; ++b) wv[b] |= (uint32_t)(uint8_t)(int8_t)c3_score(P, b, qc) << (8 * k);
                }
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) *reinterpret_cast<uint32_t *>(W.qp + b * qs + j0) = wv[b];
            *reinterpret_cast<uint32_t *>(W.qp + 4 * qs + j0) = 0u;        // an N node scores 0 against everything
        }
    }
    const uint16_t *ord = W.order[G.ob];
    int xbase = 0;
    const int nb = (n + GL - 1) / GL;
    for (int bi = nb - 1; bi >= 0; --bi) {
        const int pb = bi * GL, p = pb + li;
        const bool valid = p < n;
        int id = C3_SINK, base = 4, in_n = 0, p0 = C3_NONE, p1 = C3_NONE, e_more = C3_NONE;
        int tgt = p, hops = 0, fin = 1;
        if (valid) {
            id = ord[p];
            const c3_nrec nd = c3_ld_node(&W.nodes[id]);
            base = C3_N_BASE(nd); in_n = C3_N_INN(nd);
            if (in_n > 0) p0 = W.posof[C3_N_IN0(nd)];
            if (in_n > 1) {
                const c3_pedge pe = W.pool[C3_N_INMORE(nd)];
                p1 = W.posof[pe.id]; e_more = pe.next;
            }
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
                tgt = W.posof[best]; hops = 1; fin = 0;
            }
        }
        // third and further predecessors: positions appended to xpred (exclusive scan of the counts over the group)
        const int nx = in_n > 2 ? in_n - 2 : 0;
        int incl = nx;
#pragma unroll
        for (int d = 1; d < GL; d <<= 1) {
            const int v = C3G_SHFL(gmask, incl, gbase + ((li - d) & (GL - 1)));
            if (li >= d) incl += v;
        }
        const int xo = xbase + incl - nx;
        xbase += C3G_SHFL(gmask, incl, gbase + GL - 1);
        if (xbase > A.pool_cap) { C3G_DECLINE(); G.err = C3G_E_RETRY; }          // uniform
        if (nx > 0 && !G.err) {
            int e = e_more;
            for (int k = 0; k < nx && e != (int)C3_NONE; ++k) {
                const c3_pedge pe = W.pool[e];
                W.xpred[xo + k] = W.posof[pe.id]; e = pe.next;
            }
        }
        // hops to the sink: pointer doubling inside the batch, then one look-up above it
#pragma unroll
        for (int r = 0; r < C3G_LOG2(GL); ++r) {
            const int src = gbase + ((tgt - pb) & (GL - 1));
            const int t2 = C3G_SHFL(gmask, tgt, src), h2 = C3G_SHFL(gmask, hops, src), f2 = C3G_SHFL(gmask, fin, src);
            if (!fin && tgt >= pb && tgt < pb + GL) { hops += h2; if (f2) fin = 1; else tgt = t2; }
        }
        if (!fin) {
            if (tgt - pb < C3G_HWIN(GL) - GL) hops += hw[tgt & (C3G_HWIN(GL) - 1)];
            else hops += C3G_D_HOPS(W.desc[tgt]);
        }
        C3G_SYNC(gmask);
        if (valid) {
            hw[p & (C3G_HWIN(GL) - 1)] = (uint16_t)hops;
            if (in_n > C3_MAXPRE || hops > 65535) { C3G_DECLINE(); G.err = C3G_E_RETRY; }
            W.desc[p] = make_uint4((uint32_t)id | ((uint32_t)p0 << 16), (uint32_t)p1 | ((uint32_t)hops << 16),
                                   (uint32_t)base | ((uint32_t)in_n << 8) | ((uint32_t)xo << 16), 0u);
        }
        G.err = C3G_ANYG(gmask, G.err != 0) ? C3G_E_RETRY : 0;
        C3G_SYNC(gmask);
        if (G.err) return;
    }
}

// ---------------------------------------------------------------------------
// DP
// ---------------------------------------------------------------------------
// Shared-memory ring of one group: C3G_R slots (row = position mod C3G_R) x 6 quarter-rows (H, E1, E2 x low / high
// 8 columns) x RV vectors; vector sn sits at index sn mod RV, so the 8 lanes of a group read and write 128
// consecutive bytes.  A row only writes the vectors of its band: readers check the row's record first.
#define C3G_RING_PTR(ring, rvs, slot, sn) ((ring) + (((slot) * 6) << (rvs)) + ((sn) & ((1 << (rvs)) - 1)))

C3G_FN uint2 c3g_get_rowrec(const uint2 *srr, const uint2 *rowrec, const int pos, const int pk)
{
    uint2 r;
    if (pos - pk <= C3G_R) r = srr[pk & (C3G_R - 1)]; else r = rowrec[pk];
    return r;
}

// this lane's vector `sn` of the row at position pk (record rk) from the arena: E1 = H - d1, E2 = H - d2 with
// byte = ~(d1 | d2 << 3); the source row keeps E only at column 0
C3G_FN void c3g_arena_pred(uint32_t (&h)[8], uint32_t (&x1)[8], uint32_t (&x2)[8], const uint4 *arena, const int vs_shift,
                           const int pk, const int sn)
{
    const uint4 *s = arena + (((int64_t)pk << vs_shift) + (sn & ((1 << vs_shift) - 1))) * 3;
    const uint4 a0 = s[0], a1 = s[1], eb = s[2];
    h[0] = a0.x; h[1] = a0.y; h[2] = a0.z; h[3] = a0.w; h[4] = a1.x; h[5] = a1.y; h[6] = a1.z; h[7] = a1.w;
    if (pk == 0) {
#pragma unroll
        for (int t = 0; t < 8; ++t) x1[t] = x2[t] = C3L_FLOOR2;
        if (sn == 0) {
            const uint32_t b0 = ~eb.x & 0xffu;
            x1[0] = C3L_PACK2((int)(int16_t)(h[0] & 0xffffu) - (int)(b0 & 7u), C3L_FLOOR);
            x2[0] = C3L_PACK2((int)(int16_t)(h[0] & 0xffffu) - (int)(b0 >> 3), C3L_FLOOR);
        }
        return;
    }
    // (byte & 7) - 7 = -d1, ((byte >> 3) & 31) - 31 = -d2, added per halfword
    const uint32_t ew[4] = {eb.x, eb.y, eb.z, eb.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const uint32_t two = C3L_PRMT(ew[t >> 1], 0u, (t & 1) ? 0x4342u : 0x4140u);    // bytes -> halfwords
        x1[t] = C3G_VADD2(h[t], C3G_VADD2(two & 0x00070007u, 0xfff9fff9u));
        x2[t] = C3G_VADD2(h[t], C3G_VADD2((two >> 3) & 0x001f001fu, 0xffe1ffe1u));
    }
}

// this lane's vector `sn` of a predecessor row: from the ring when the row is at most ring_delta positions back and
// fits a slot, else from the arena; floor outside the row's band.  `always`: lanes outside this row's band take
// the plain path whatever they would read (their results are dropped).
template <int RVS>
C3G_FN void c3g_fetch_pred(uint32_t (&h)[8], uint32_t (&x1)[8], uint32_t (&x2)[8], const uint4 *ring, const uint4 *arena,
                           const int vs_shift, const int pos, const int pk, const uint2 rk, const int sn, const int ring_delta,
                           const bool always)
{
    const int pb = C3G_R_BEG(rk), pe = C3G_R_END(rk);
    const bool in_ring = pos - pk <= ring_delta && pe - pb < (1 << RVS);
    if (in_ring && (always || (sn >= pb && sn <= pe))) {
        const uint4 *s = C3G_RING_PTR(ring, RVS, pk & (C3G_R - 1), sn);
        constexpr int st = 1 << RVS;
        const uint4 a0 = s[0], a1 = s[st], b0 = s[2 * st], b1 = s[3 * st], c0 = s[4 * st], c1 = s[5 * st];
        h[0] = a0.x; h[1] = a0.y; h[2] = a0.z; h[3] = a0.w; h[4] = a1.x; h[5] = a1.y; h[6] = a1.z; h[7] = a1.w;
        x1[0] = b0.x; x1[1] = b0.y; x1[2] = b0.z; x1[3] = b0.w; x1[4] = b1.x; x1[5] = b1.y; x1[6] = b1.z; x1[7] = b1.w;
        x2[0] = c0.x; x2[1] = c0.y; x2[2] = c0.z; x2[3] = c0.w; x2[4] = c1.x; x2[5] = c1.y; x2[6] = c1.z; x2[7] = c1.w;
    } else if (!always && sn >= pb && sn <= pe) {
        c3g_arena_pred(h, x1, x2, arena, vs_shift, pk, sn);
    } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) h[t] = x1[t] = x2[t] = C3L_FLOOR2;
    }
}

// source row: cells, ring slot 0, arena, record
template <int RVS>
C3G_FN void c3g_source_row(c3g_grp &G, const c3g_args &L, const c3_poa_para_dev &P, const c3g_ws &W, uint4 *ring, uint2 *srr,
                           uint4 *arena, uint2 &rec_out, const int li, const unsigned gmask)
{
    const int oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const int qlen = G.qlen;
    const int rem = C3G_D_HOPS(W.desc[0]) - 1;
    const int rr = qlen - rem;
    const int end = min(qlen, max(0, rr) + G.w);
    const int end_sn = end >> 4, e0 = min(qlen, end_sn * 16 + 15);
    const int nvec = end_sn + 1;
    if (nvec > (1 << L.vs_shift)) { C3G_DECLINE(); G.err = C3G_E_RETRY; return; }
    const int vsm = (1 << L.vs_shift) - 1;
    bool low = false;
    for (int sn = li; sn <= end_sn; sn += C3G_GL) {
        uint32_t h[8], x1[8], x2[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            int hv[2], a1[2], a2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c = sn * 16 + 2 * t + u;
                hv[u] = a1[u] = a2[u] = C3L_FLOOR;
                if (c <= e0) {
                    if (c == 0) { hv[u] = 0; a1[u] = -oe1; a2[u] = -oe2; }
                    else hv[u] = max(-(P.o1 + P.e1 * c), -(P.o2 + P.e2 * c));
                    if (hv[u] < C3L_FLOOR + 1024) low = true;
                }
            }
            h[t] = C3L_PACK2(hv[0], hv[1]); x1[t] = C3L_PACK2(a1[0], a1[1]); x2[t] = C3L_PACK2(a2[0], a2[1]);
        }
        if (nvec <= (1 << RVS)) {
            uint4 *s = C3G_RING_PTR(ring, RVS, 0, sn);
            constexpr int st = 1 << RVS;
            s[0] = make_uint4(h[0], h[1], h[2], h[3]); s[st] = make_uint4(h[4], h[5], h[6], h[7]);
            s[2 * st] = make_uint4(x1[0], x1[1], x1[2], x1[3]); s[3 * st] = make_uint4(x1[4], x1[5], x1[6], x1[7]);
            s[4 * st] = make_uint4(x2[0], x2[1], x2[2], x2[3]); s[5 * st] = make_uint4(x2[4], x2[5], x2[6], x2[7]);
        }
        uint4 *d = arena + (int64_t)(sn & vsm) * 3;
        d[0] = make_uint4(h[0], h[1], h[2], h[3]); d[1] = make_uint4(h[4], h[5], h[6], h[7]);
        d[2] = make_uint4(sn == 0 ? ~(uint32_t)(oe1 | (oe2 << 3)) : ~0u, ~0u, ~0u, ~0u);
    }
    if (C3G_ANYG(gmask, low)) { C3G_DECLINE(); G.err = C3G_E_RETRY; return; }
    const uint2 rec = make_uint2((uint32_t)end_sn << 16, 1u);      // successors of the source start at column 1
    if (li == 0) { W.rowrec[0] = rec; srr[0] = rec; }
    rec_out = rec;
}

// One DP row (position pos, descriptor d; rprev: record of the row at pos - 1, replaced by this row's).  Returns the
// band width in columns, 0 after an error.  Lanes whose vector lies outside the band run along with whatever they
// load -- nothing they compute is stored or enters the row's arg-max.  Columns past qlen inside the last vector are
// computed like any other (nothing at or left of qlen depends on them) and only kept out of the arg-max.
// MULTI = false: the arena rows hold 8 vectors, so a row is one pass of the 8 lanes.
template <int RVS, bool MULTI>
C3G_FN int c3g_row(c3g_grp &G, const c3g_args &L, const c3_poa_para_dev &P, const c3g_ws &W, uint4 *ring, uint2 *srr,
                   uint4 *arena, const int pos, const uint4 d, uint2 &rprev, const bool live, const int li, const int gbase)
{
    // All 32 lanes of the warp run every row step together and every collective names the full warp (a member mask
    // that differs between the groups costs a MATCH + vote per shuffle).  A group without a row (`live` false) runs
    // along on harmless inputs and stores nothing.
    const int e1 = P.e1, e2 = P.e2, oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const int qlen = G.qlen;
    const int npre = live ? C3G_D_NPRE(d) : 1, nbase = live ? C3G_D_BASE(d) : 0;
    const int p0 = live ? C3G_D_P0(d) : pos - 1, p1 = C3G_D_P1(d);
    const uint2 r0 = (p0 == pos - 1) ? rprev : c3g_get_rowrec(srr, W.rowrec, pos, p0);
    uint2 r1 = make_uint2(1u, 0u);                                // empty band
    int mpl = min(G.n, C3G_R_MP(r0)), mpr = C3G_R_MP(r0), minb = C3G_R_BEG(r0);
    if (npre > 1) {
        r1 = (p1 == pos - 1) ? rprev : c3g_get_rowrec(srr, W.rowrec, pos, p1);
        mpl = min(mpl, C3G_R_MP(r1)); mpr = max(mpr, C3G_R_MP(r1)); minb = min(minb, C3G_R_BEG(r1));
        for (int k = 2; k < npre; ++k) {
            const int pk = W.xpred[C3G_D_XOFS(d) + k - 2];
            const uint2 rk = c3g_get_rowrec(srr, W.rowrec, pos, pk);
            mpl = min(mpl, C3G_R_MP(rk)); mpr = max(mpr, C3G_R_MP(rk)); minb = min(minb, C3G_R_BEG(rk));
        }
    }
    const int rr = qlen - (C3G_D_HOPS(d) - 1);
    int beg_sn = max(max(0, min(mpl, rr) - G.w) >> 4, minb);
    int end_sn = max(min(qlen, max(mpr, rr) + G.w) >> 4, beg_sn);
    if (!live) { beg_sn = 0; end_sn = 0; }
    bool bad = false;
    if (end_sn - beg_sn + 1 > (MULTI ? (1 << L.vs_shift) : C3G_GL)) { bad = true; end_sn = beg_sn; }   // wider than an arena row
    const int nvec = end_sn - beg_sn + 1;
    const bool to_ring = MULTI ? nvec <= (1 << RVS) : true;
    const uint32_t ne1 = C3L_PACK2(-e1, -e1), ne2 = C3L_PACK2(-e2, -e2), noe1 = C3L_PACK2(-oe1, -oe1), noe2 = C3L_PACK2(-oe2, -oe2);
    const int d21 = oe1 - oe2;
    int pc1 = C3L_FLOOR, pc2 = C3L_FLOOR;                          // F1, F2' entering the pass (chain domain, see below)
    uint32_t mcarry = C3L_FLOOR2;                                  // merged predecessor H of the vector before the pass
    unsigned bestkey = 0u;
    int h_first = 0x7fff;
    // everything read from the ring must be in registers before any lane overwrites the slot of row pos - C3G_R:
    // a row of several passes stores its first vectors before it has read the last ones, so it does not use that slot
    const int ring_delta = (!MULTI || nvec <= C3G_GL) ? C3G_R : C3G_R - 1;
    int sn0 = beg_sn;
    do {
        const int l = (li - sn0) & 7, sn = sn0 + l;
        const bool act = sn <= end_sn;
        const int j0 = sn << 4;
        // scores of the 16 columns against the node base (row 4 of the profile: zeros, an N node)
        const uint4 s4 = *reinterpret_cast<const uint4 *>(W.qp + nbase * L.A.qp_stride + j0);
        uint32_t h[8], x1[8], x2[8];
        c3g_fetch_pred<RVS>(h, x1, x2, ring, arena, L.vs_shift, pos, p0, r0, sn, ring_delta, !act || !live);
        for (int k = 1; k < npre; ++k) {
            int pk = p1; uint2 rk = r1;
            if (k > 1) { pk = W.xpred[C3G_D_XOFS(d) + k - 2]; rk = c3g_get_rowrec(srr, W.rowrec, pos, pk); }
            uint32_t th[8], t1[8], t2[8];
            c3g_fetch_pred<RVS>(th, t1, t2, ring, arena, L.vs_shift, pos, pk, rk, sn, ring_delta, !act || !live);
#pragma unroll
            for (int u = 0; u < 8; ++u) { h[u] = C3L_VMAX2(h[u], th[u]); x1[u] = C3L_VMAX2(x1[u], t1[u]); x2[u] = C3L_VMAX2(x2[u], t2[u]); }
        }
        // M: merged predecessor H one column to the left (the first cell of the band never takes M)
        uint32_t top = (uint32_t)C3G_SHFL(C3_FULL, h[7], gbase + ((li - 1) & 7));
        if (l == 0) top = mcarry;
        if (MULTI) mcarry = (uint32_t)C3G_SHFL(C3_FULL, h[7], gbase + ((sn0 + 7) & 7));      // (every lane of the warp: no group-dependent condition around a collective)
        uint32_t hme[8];
        {
            const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t sc = C3L_PRMT(sw[t >> 1], 0u, (t & 1) ? 0xb3a2u : 0x9180u);
                const uint32_t mw = C3L_PRMT(t == 0 ? top : h[t - 1], h[t], 0x5432u);
                hme[t] = C3L_VMAX3_2(C3L_VADDMAX2(mw, sc, C3L_FLOOR2), x1[t], x2[t]);
            }
        }
        // horizontal gap, chain domain: g1 = F1 at the column, g2 = F2 at the column + (oe2 - oe1); both take
        // a = hme - oe1 per column.  Local chain from "nothing enters the lane" first, then the 8-lane scan.
        int g1 = C3L_FLOOR, g2 = C3L_FLOOR;
        uint32_t fmp[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const uint32_t hs = C3L_VADDMAX2(hme[t], noe1, 0x80008000u);
            const int a0 = (int)(int16_t)(hs & 0xffffu), a1 = (int)hs >> 16;
            const int f0 = C3L_ADDMAX(g2, d21, g1);
            g1 = C3L_ADDMAX(g1, -e1, a0); g2 = C3L_ADDMAX(g2, -e2, a0);
            const int f1 = C3L_ADDMAX(g2, d21, g1);
            g1 = C3L_ADDMAX(g1, -e1, a1); g2 = C3L_ADDMAX(g2, -e2, a1);
            fmp[t] = C3L_PACK2(f0, f1);
        }
        // scan over the lanes, both chains packed in one word (low half: chain 1): value entering lane l =
        // max(pass carry decayed, local outputs of the lanes before, decayed).  All values are int16-sized.
        const uint32_t dec = C3L_PACK2(16 * e1, 16 * e2);
        uint32_t tt = C3G_VADD2(C3L_PACK2(g1, g2), dec * (uint32_t)l);
#pragma unroll
        for (int dd = 1; dd < C3G_GL; dd <<= 1) {
            const uint32_t v = (uint32_t)C3G_SHFL(C3_FULL, tt, gbase + ((li - dd) & 7));
            if (l >= dd) tt = C3L_VMAX2(tt, v);
        }
        const uint32_t ex = (uint32_t)C3G_SHFL(C3_FULL, tt, gbase + ((li - 1) & 7));
        int c1 = (int)(int16_t)(ex & 0xffffu) - 16 * e1 * (l - 1), c2 = ((int)ex >> 16) - 16 * e2 * (l - 1);
        if (MULTI) {
            c1 = l == 0 ? pc1 : max(c1, pc1 - 16 * e1 * l);
            c2 = l == 0 ? pc2 : max(c2, pc2 - 16 * e2 * l);
            const uint32_t tot = (uint32_t)C3G_SHFL(C3_FULL, tt, gbase + ((sn0 + 7) & 7));
            pc1 = max((int)(int16_t)(tot & 0xffffu) - 16 * e1 * 7, pc1 - 16 * e1 * 8);
            pc2 = max(((int)tot >> 16) - 16 * e2 * 7, pc2 - 16 * e2 * 8);
        } else if (l == 0) { c1 = C3L_FLOOR; c2 = C3L_FLOOR; }
        c1 = max(c1, C3L_FLOOR); c2 = max(c2 + d21, C3L_FLOOR);    // F1, F2 entering the vector
        const uint32_t c1p = C3L_PACK2(c1, c1), c2p = C3L_PACK2(c2, c2);
        uint32_t hh[8], n1[8], n2[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const uint32_t rp1 = C3L_PACK2(-e1 * 2 * t, -e1 * (2 * t + 1)), rp2 = C3L_PACK2(-e2 * 2 * t, -e2 * (2 * t + 1));
            const uint32_t ff = C3L_VADDMAX2(c2p, rp2, C3L_VADDMAX2(c1p, rp1, fmp[t]));
            hh[t] = C3L_VMAX2(hme[t], ff);
            n1[t] = C3L_VADDMAX2(x1[t], ne1, C3L_VADDMAX2(hh[t], noe1, C3L_FLOOR2));
            n2[t] = C3L_VADDMAX2(x2[t], ne2, C3L_VADDMAX2(hh[t], noe2, C3L_FLOOR2));
        }
        if (sn == beg_sn) h_first = (int)(int16_t)(hh[0] & 0xffffu);
        // simd_abpoa_ada_max_i as one packed max: value (biased to unsigned) in the high half, tie-break priority in
        // the low half (lowest SIMD lane, then the last vector, then the earliest vector)
        if (act) {
            unsigned lk = 0u;
            const int lim = min(qlen, end_sn * 16 + 15) - j0;        // last column of this vector that exists (>= 15: all)
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                unsigned klo = C3L_PRMT(hh[t], (uint32_t)((15 - 2 * t) << 12), 0x1054u) ^ 0x80000000u;   // lo half -> high half, priority below
                unsigned khi = C3L_PRMT(hh[t], (uint32_t)((14 - 2 * t) << 12), 0x3254u) ^ 0x80000000u;
                if (lim < 15) { if (2 * t > lim) klo = 0u; if (2 * t + 1 > lim) khi = 0u; }
                lk = max(lk, max(klo, khi));
            }
            const unsigned vp = (sn == end_sn) ? 0xfffu : (unsigned)(0xffe - (sn - beg_sn));
            bestkey = max(bestkey, lk | vp);
        }
        C3G_SYNC(C3_FULL);
        if (act && live) {
            const uint4 oa0 = make_uint4(hh[0], hh[1], hh[2], hh[3]), oa1 = make_uint4(hh[4], hh[5], hh[6], hh[7]);
            if (to_ring) {
                uint4 *s = C3G_RING_PTR(ring, RVS, pos & (C3G_R - 1), sn);
                constexpr int st = 1 << RVS;
                s[0] = oa0; s[st] = oa1;
                s[2 * st] = make_uint4(n1[0], n1[1], n1[2], n1[3]); s[3 * st] = make_uint4(n1[4], n1[5], n1[6], n1[7]);
                s[4 * st] = make_uint4(n2[0], n2[1], n2[2], n2[3]); s[5 * st] = make_uint4(n2[4], n2[5], n2[6], n2[7]);
            }
            // E byte: n + ~H = -(H - n) - 1 per halfword, whose low bits are the complement of H - n; bits 0-2 from E1,
            // bits 3-7 from E2 (what spills above bit 7 is dropped by the byte pack below)
            uint32_t eb[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t nh = ~hh[t];
                const uint32_t s1 = C3G_VADD2(n1[t], nh), s2 = C3G_VADD2(n2[t], nh) << 3;
                eb[t] = (s1 & 0x00070007u) | (s2 & ~0x00070007u);
            }
            uint4 *dst = arena + (((int64_t)pos << L.vs_shift) + (sn & ((1 << L.vs_shift) - 1))) * 3;
            dst[0] = oa0; dst[1] = oa1;
            dst[2] = make_uint4(C3L_PRMT(eb[0], eb[1], 0x6420u), C3L_PRMT(eb[2], eb[3], 0x6420u),
                                C3L_PRMT(eb[4], eb[5], 0x6420u), C3L_PRMT(eb[6], eb[7], 0x6420u));
        }
        sn0 += C3G_GL;
    } while (MULTI && C3G_ANYG(C3_FULL, sn0 <= end_sn));
    // exactness of the int16 form (see poa_lane.cuh): the first band cell comfortably above the floor means that
    // every cell of the row is reachable and was never clamped.  The verdict rides on the arg-max reduction.
    {
        const int end = min(qlen, end_sn * 16 + 15);
        const int D = max(min(oe1, oe2), max(e1, e2));
        if (bad || h_first < C3L_FLOOR + C3L_LOW_GUARD + oe2 + D * (end - (beg_sn << 4) + 1)) bestkey = 0xffffffffu;
    }
#pragma unroll
    for (int dd = 1; dd < C3G_GL; dd <<= 1) bestkey = max(bestkey, (unsigned)C3G_SHFL(C3_FULL, bestkey, gbase + (li ^ dd)));
    if (bestkey == 0xffffffffu && live) { C3G_DECLINE(); G.err = C3G_E_RETRY; }
    int best_i = -1;
    if ((int)(bestkey >> 16) - 32768 > C3L_FLOOR) {
        const int sl = 15 - (int)((bestkey >> 12) & 15u);
        const int vp = (int)(bestkey & 0xfffu);
        const int snb = (vp == 0xfff) ? end_sn : beg_sn + (0xffe - vp);
        best_i = (snb << 4) + sl;
    }
    const uint2 rec = make_uint2((uint32_t)beg_sn | ((uint32_t)end_sn << 16), (uint32_t)(best_i + 1));
    if (li == 0 && live) { W.rowrec[pos] = rec; srr[pos & (C3G_R - 1)] = rec; }
    if (live) rprev = rec;
    C3G_SYNC(C3_FULL);
    return min(qlen, end_sn * 16 + 15) - (beg_sn << 4) + 1;
}

// ---------------------------------------------------------------------------
// backtrack: abPOA's M -> E1 -> E2 -> F1 -> F2 order and op-mask state machine on the arena's (H, H-E1, H-E2)
// cells.  Group-uniform; runs of match/mismatch moves along consecutive rows are verified 7 at a time.
// Returns the number of cigar ops or a negative code.
// ---------------------------------------------------------------------------
C3_HD __forceinline__ int c3g_cell_h(const uint4 *arena, const int vs_shift, const int pos, const int j)
{
    const int16_t *p = reinterpret_cast<const int16_t *>(arena + (((int64_t)pos << vs_shift) + ((j >> 4) & ((1 << vs_shift) - 1))) * 3);
    return c3l_map((int)p[j & 15]);
}
C3_HD __forceinline__ int c3g_cell_eb(const uint4 *arena, const int vs_shift, const int pos, const int j)
{
    const uint8_t *p = reinterpret_cast<const uint8_t *>(arena + (((int64_t)pos << vs_shift) + ((j >> 4) & ((1 << vs_shift) - 1))) * 3 + 2);
    return (int)(~p[j & 15] & 0xff);               // (H - E1) | (H - E2) << 3
}
C3_HD __forceinline__ int c3g_pred_pos(const c3g_ws &W, const uint4 d, const int k)
{
    return k == 0 ? C3G_D_P0(d) : k == 1 ? C3G_D_P1(d) : (int)W.xpred[C3G_D_XOFS(d) + k - 2];
}

// Backtrack window: the rows top .. top - C3G_BTK + 1 staged in shared memory (the idle ring) with ONE round of
// independent loads: per row the two 16-column vectors (H and E bytes) around the column the path is expected
// to cross it at (j0 - r: one column per row along the diagonal, +- 8 columns of drift), its descriptor and its
// record.  Whatever falls outside is read from global memory.
struct c3g_btw { const uint4 *wH, *wE, *wD; const uint2 *wR; int top, bot, j0; };

template <int GL>
C3G_FN void c3g_bt_fill(c3g_btw &B, uint8_t *sg, const c3g_ws &W, const uint4 *arena, const int vs, const int top, const int j0,
                        const int li, const unsigned gmask)
{
    constexpr int K = C3G_BTK(GL);
    uint4 *wH = reinterpret_cast<uint4 *>(sg), *wE = wH + K * 4, *wD = wE + K * 2;
    uint2 *wR = reinterpret_cast<uint2 *>(wD + K);
    const int vsm = (1 << vs) - 1;
    C3G_SYNC(gmask);                                          // nobody still reads the previous window
#pragma unroll
    for (int u = 0; u < (K + GL - 1) / GL; ++u) {
        const int r = li + u * GL, p = top - r;
        if (r < K && p >= 0) {
            const int vlo = max(0, (j0 - r - 8) >> 4);
            const uint4 *s0 = arena + (((int64_t)p << vs) + (vlo & vsm)) * 3, *s1 = arena + (((int64_t)p << vs) + ((vlo + 1) & vsm)) * 3;
            const uint4 a0 = s0[0], a1 = s0[1], a2 = s0[2], b0 = s1[0], b1 = s1[1], b2 = s1[2];
            const uint4 dd = W.desc[p]; const uint2 rr = W.rowrec[p];
            wH[r * 4] = a0; wH[r * 4 + 1] = a1; wH[r * 4 + 2] = b0; wH[r * 4 + 3] = b1;
            wE[r * 2] = a2; wE[r * 2 + 1] = b2; wD[r] = dd; wR[r] = rr;
        }
    }
    C3G_SYNC(gmask);
    B.wH = wH; B.wE = wE; B.wD = wD; B.wR = wR; B.top = top; B.bot = max(0, top - K + 1); B.j0 = j0;
}
C3G_FN uint4 c3g_bt_desc(const c3g_btw &B, const c3g_ws &W, const int p)
{
    return (p <= B.top && p >= B.bot) ? B.wD[B.top - p] : W.desc[p];
}
C3G_FN uint2 c3g_bt_rec(const c3g_btw &B, const c3g_ws &W, const int p)
{
    return (p <= B.top && p >= B.bot) ? B.wR[B.top - p] : W.rowrec[p];
}
C3G_FN int c3g_bt_h(const c3g_btw &B, const uint4 *arena, const int vs, const int p, const int j)
{
    if (p <= B.top && p >= B.bot) {
        const int r = B.top - p, v = (j >> 4) - max(0, (B.j0 - r - 8) >> 4);
        if ((unsigned)v < 2u) return c3l_map((int)reinterpret_cast<const int16_t *>(B.wH + r * 4 + v * 2)[j & 15]);
    }
    return c3g_cell_h(arena, vs, p, j);
}
C3G_FN int c3g_bt_eb(const c3g_btw &B, const uint4 *arena, const int vs, const int p, const int j)
{
    if (p <= B.top && p >= B.bot) {
        const int r = B.top - p, v = (j >> 4) - max(0, (B.j0 - r - 8) >> 4);
        if ((unsigned)v < 2u) return (int)(~reinterpret_cast<const uint8_t *>(B.wE + r * 2 + v)[j & 15] & 0xff);
    }
    return c3g_cell_eb(arena, vs, p, j);
}

template <int GL>
C3G_FN int c3g_backtrack(c3g_grp &G, const c3g_args &L, const c3_poa_para_dev &P, const c3g_ws &W, const uint4 *arena,
                         uint8_t *sg, const int li, const int gbase, const unsigned gmask)
{
    const int e1 = P.e1, e2 = P.e2, oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const int vs = L.vs_shift;
    const uint8_t *q = G.q; const int qlen = G.qlen, n = G.n;
    const int cap = L.A.cigar_cap;
    unsigned long long *cg = W.cigar;
    int nc = 0, j, pos, hij;
    {
        const uint4 ds = W.desc[n - 1];
        int best = -0x7fffffff - 1, bj = -1, bk = -1;
        const int skn = C3G_D_NPRE(ds);
        for (int k = 0; k < skn; ++k) {
            const int pk = c3g_pred_pos(W, ds, k);
            const uint2 rp = W.rowrec[pk];
            const int en = min(qlen, C3G_R_END(rp) * 16 + 15);
            const int val = c3g_cell_h(arena, vs, pk, en);
            if (val > best) { best = val; bj = en; bk = pk; }
        }
        if (bk < 0 || qlen - bj + 8 > cap) { C3G_DECLINE(); return C3G_E_RETRY; }
        for (int t = qlen - li; t > bj; t -= GL)
            cg[qlen - t] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(t - 1) << 32);
        nc = qlen - bj; j = bj; pos = bk; hij = best;
    }
    c3g_btw B;
    B.top = -1; B.bot = 0; B.j0 = 0; B.wH = B.wE = B.wD = nullptr; B.wR = nullptr;
    int cur_op = C3_OP_ALL;
    while (pos != 0 && j > 0) {
        if (pos > B.top || (pos - C3G_BTV(GL) < B.bot && B.bot > 0)) c3g_bt_fill<GL>(B, sg, W, arena, vs, pos, j, li, gmask);
        if (cur_op == C3_OP_ALL) {
            // Window slot w = lane <-> position pos - w.  The chain of FIRST predecessors inside the slots is resolved
            // by pointer doubling over shuffles (it runs on through the bubbles of the graph); lane t then looks at the
            // t-th row of the chain at column j - t, and the leading run of verified match/mismatch moves is taken at once.
            const int pw = pos - li;
            uint4 dl = make_uint4(0u, 0u, 0u, 0u); uint2 rl = make_uint2(1u, 0u);
            if (pw >= 0) { dl = c3g_bt_desc(B, W, pw); rl = c3g_bt_rec(B, W, pw); }
            int f = GL;                                        // slot of this row's first predecessor; GL = outside / none
            if (pw >= 1) { const int dlt = pos - C3G_D_P0(dl); if (dlt < GL) f = dlt; }
            int tbl[C3G_LOG2(GL)];
            tbl[0] = f;
#pragma unroll
            for (int b2 = 1; b2 < C3G_LOG2(GL); ++b2) {
                const int prev = tbl[b2 - 1];
                const int nx = C3G_SHFL(gmask, prev, gbase + (prev & (GL - 1)));
                tbl[b2] = prev < GL ? nx : GL;
            }
            int sl = 0;                                        // slot of the t-th row of the chain (t = lane)
#pragma unroll
            for (int b2 = 0; b2 < C3G_LOG2(GL); ++b2) {
                const int nx = C3G_SHFL(gmask, tbl[b2], gbase + (sl & (GL - 1)));
                if ((li >> b2) & 1) sl = sl < GL ? nx : GL;
            }
            const bool have = sl < GL;
            const int srcl = gbase + (sl & (GL - 1));
            const uint32_t cdx = (uint32_t)C3G_SHFL(gmask, dl.x, srcl), cdz = (uint32_t)C3G_SHFL(gmask, dl.z, srcl);
            const uint32_t crx = (uint32_t)C3G_SHFL(gmask, rl.x, srcl);
            const int pc = pos - sl, jt = j - li;              // chain row t: position, column
            const int bl = (int)(crx & 0xffffu) * 16, el = min(qlen, (int)(crx >> 16) * 16 + 15);
            const bool inb = have && jt >= 1 && jt >= bl && jt <= el;
            int ht = C3_NEG_INF;
            if (inb) ht = c3g_bt_h(B, arena, vs, pc, jt);
            const int nxl = gbase + ((li + 1) & (GL - 1));
            const int hn = C3G_SHFL(gmask, ht, nxl), bn = C3G_SHFL(gmask, bl, nxl), en = C3G_SHFL(gmask, el, nxl);
            const int haven = C3G_SHFL(gmask, (int)have, nxl);
            bool ok = li < GL - 1 && inb && haven && pc >= 1;
            if (ok) {
                const int st = c3_score(P, (int)(cdz & 0xffu), q[jt - 1]);
                ok = jt - 1 >= max(bn, bl) && jt - 1 <= en && ht == hn + st;
            }
            const unsigned okm = (C3G_BALLOT(gmask, ok) >> gbase) & C3G_LMASK(GL);
            int Lr = C3G_FFS(~okm) - 1;
            Lr = min(Lr, cap - 8 - j - nc);
            if (Lr > 0) {
                if (li < Lr) cg[nc + li] = C3_CG_MATCH | ((unsigned long long)(cdx & 0xffffu) << 8) | ((unsigned long long)(jt - 1) << 32);
                nc += Lr; j -= Lr;
                pos -= C3G_SHFL(gmask, sl, gbase + Lr);
                hij = C3G_SHFL(gmask, ht, gbase + Lr);
                continue;
            }
        }
        // generic single step
        const uint4 d = c3g_bt_desc(B, W, pos);
        const uint2 rt = c3g_bt_rec(B, W, pos);
        const int i = C3G_D_ID(d);
        const int b = C3G_R_BEG(rt) * 16, en = min(qlen, C3G_R_END(rt) * 16 + 15);
        if (j < b || j > en) { C3G_DECLINE(); return C3G_E_RETRY; }
        const int s = c3_score(P, C3G_D_BASE(d), q[j - 1]);
        const int npre = C3G_D_NPRE(d);
        int hit = 0;
        unsigned long long opw = 0;
        if (cur_op & C3_OP_M) {
            for (int k = 0; k < npre; ++k) {
                const int pk = c3g_pred_pos(W, d, k);
                const uint2 pr = c3g_bt_rec(B, W, pk);
                const int pbeg = C3G_R_BEG(pr) * 16, pend = min(qlen, C3G_R_END(pr) * 16 + 15);
                if (j - 1 < max(pbeg, b) || j - 1 > pend) continue;
                const int ph = c3g_bt_h(B, arena, vs, pk, j - 1);
                if (ph + s == hij) {
                    opw = C3_CG_MATCH | ((unsigned long long)i << 8) | ((unsigned long long)(j - 1) << 32);
                    pos = pk; --j; hit = 1; cur_op = C3_OP_ALL; hij = ph;
                    break;
                }
            }
        }
        if (!hit && (cur_op & C3_OP_E)) {
            const int ceb = c3g_bt_eb(B, arena, vs, pos, j);
            const int ce1 = hij - (ceb & 7), ce2 = hij - (ceb >> 3);
            for (int k = 0; k < npre; ++k) {
                const int pk = c3g_pred_pos(W, d, k);
                const uint2 pr = c3g_bt_rec(B, W, pk);
                const int pbeg = C3G_R_BEG(pr) * 16, pend = min(qlen, C3G_R_END(pr) * 16 + 15);
                if (j < pbeg || j > pend) continue;
                const int ph = c3g_bt_h(B, arena, vs, pk, j);
                int pe1, pe2;
                if (pk == 0) { pe1 = j == 0 ? -oe1 : C3_NEG_INF; pe2 = j == 0 ? -oe2 : C3_NEG_INF; }
                else { const int peb = c3g_bt_eb(B, arena, vs, pk, j); pe1 = ph - (peb & 7); pe2 = ph - (peb >> 3); }
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
                    break;
                }
            }
        }
        if (!hit && (cur_op & C3_OP_F)) {
            int hl = C3_NEG_INF;
            if (j - 1 >= b) {
                // F is not stored: rebuild F[j] and F[j-1] of this row from its H, 8 lanes x 16 columns with the same
                // (max,+) scan as the DP; only the lane that holds column j - 1 decides
                const int jm = j - 1, sb = b >> 4;
                int f1 = C3_NEG_INF, f2 = C3_NEG_INF, f1l = C3_NEG_INF, f2l = C3_NEG_INF;
                int pc1 = C3_NEG_INF, pc2 = C3_NEG_INF;            // F entering the pass
                for (int sn0 = sb; sn0 <= (jm >> 4); sn0 += GL) {
                    const int sn = sn0 + li;
                    const int16_t *hp = reinterpret_cast<const int16_t *>(arena + (((int64_t)pos << vs) + (sn & ((1 << vs) - 1))) * 3);
                    int hv[16];
                    if (sn <= (jm >> 4)) {
                        const uint4 a0 = reinterpret_cast<const uint4 *>(hp)[0], a1 = reinterpret_cast<const uint4 *>(hp)[1];
                        int t8[8];
                        c3l_unpack8(a0, t8);
#pragma unroll
                        for (int k = 0; k < 8; ++k) hv[k] = c3l_map(t8[k]);
                        c3l_unpack8(a1, t8);
#pragma unroll
                        for (int k = 0; k < 8; ++k) hv[8 + k] = c3l_map(t8[k]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; ++k) hv[k] = C3_NEG_INF;
                    }
                    // local outputs with nothing entering: value leaving the vector
                    int o1 = C3_NEG_INF, o2 = C3_NEG_INF;
#pragma unroll
                    for (int k = 0; k < 16; ++k) { o1 = max(o1 - e1, hv[k] - oe1); o2 = max(o2 - e2, hv[k] - oe2); }
                    int t1 = o1 + 16 * e1 * li, t2 = o2 + 16 * e2 * li;
#pragma unroll
                    for (int dd = 1; dd < GL; dd <<= 1) {
                        const int v1 = C3G_SHFL(gmask, t1, gbase + ((li - dd) & (GL - 1))), v2 = C3G_SHFL(gmask, t2, gbase + ((li - dd) & (GL - 1)));
                        if (li >= dd) { t1 = max(t1, v1); t2 = max(t2, v2); }
                    }
                    int c1 = C3G_SHFL(gmask, t1, gbase + ((li - 1) & (GL - 1))), c2 = C3G_SHFL(gmask, t2, gbase + ((li - 1) & (GL - 1)));
                    const int tot1 = C3G_SHFL(gmask, t1, gbase + GL - 1), tot2 = C3G_SHFL(gmask, t2, gbase + GL - 1);
                    c1 = li == 0 ? pc1 : max(c1 - 16 * e1 * (li - 1), pc1 - 16 * e1 * li);
                    c2 = li == 0 ? pc2 : max(c2 - 16 * e2 * (li - 1), pc2 - 16 * e2 * li);
                    pc1 = max(tot1 - 16 * e1 * (GL - 1), pc1 - 16 * e1 * GL);
                    pc2 = max(tot2 - 16 * e2 * (GL - 1), pc2 - 16 * e2 * GL);
                    if (sn == (jm >> 4)) {                         // this lane holds column j - 1: F[j-1] and F[j]
                        int a1 = c1, a2 = c2, hlast = C3_NEG_INF, p1 = C3_NEG_INF, p2 = C3_NEG_INF;
                        const int tm = jm & 15;
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            if (k <= tm) { hlast = hv[k]; p1 = a1; p2 = a2; a1 = max(a1 - e1, hlast - oe1); a2 = max(a2 - e2, hlast - oe2); }
                        }
                        f1 = a1; f2 = a2; f1l = p1; f2l = p2; hl = hlast;
                    }
                }
                const int own = gbase + (((jm >> 4) - sb) & (GL - 1));
                f1 = C3G_SHFL(gmask, f1, own); f2 = C3G_SHFL(gmask, f2, own); f1l = C3G_SHFL(gmask, f1l, own);
                f2l = C3G_SHFL(gmask, f2l, own); hl = C3G_SHFL(gmask, hl, own);
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
        if (li == 0) cg[nc] = opw;
        ++nc;
        if (nc + j + 8 > cap) { C3G_DECLINE(); return C3G_E_RETRY; }
    }
    for (int t = j - li; t > 0; t -= GL)
        cg[nc + j - t] = C3_CG_INS | ((unsigned long long)C3_NONE << 8) | ((unsigned long long)(t - 1) << 32);
    nc += j;
    C3G_SYNC(gmask);
    return nc;
}

// ---------------------------------------------------------------------------
// merge (abpoa_add_graph_alignment): the cigar is walked from its tail = forward order, 8 ops at a time; ops that
// only bump the weight of an existing edge between two matched nodes are applied by all lanes at once, the rest
// (new nodes / edges) goes through lane 0 in order.  A new node is not linked into a list: its place in the order
// is the gap of the OLD order it falls into (W.gaps, non-decreasing in creation order), see c3g_reorder.
// ---------------------------------------------------------------------------
C3G_FN int c3g_tail_gap(const c3g_ws &W, const uint16_t *ord, const int n_old, const c3_pnode &na, const int p_start)
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

template <int GL>
C3G_FN int c3g_merge(c3g_grp &G, const c3g_args &L, const c3g_ws &W, const int nc, const int li, const int gbase, const unsigned gmask)
{
    const uint8_t *q = G.q;
    const unsigned long long *cg = W.cigar;
    const uint16_t *ord = W.order[G.ob];
    const int n_old = G.n;
    c3_graph g; g.nodes = W.nodes; g.pool = W.pool; g.node_n = G.node_n; g.pool_n = G.pool_n;
    g.node_cap = L.A.node_cap; g.pool_cap = L.A.pool_cap; g.err = 0;
    int last_id = C3_SRC, last_new = 0;                     // uniform
    int last_gap = 0, last_p = 0;                           // lane 0: gap of the last new node / old position its group walk starts at
    for (int tb = nc - 1; tb >= 0; tb -= GL) {
        const int t = tb - li;
        const bool have = t >= 0;
        const unsigned long long op = have ? cg[t] : C3_CG_DEL;
        const int kind = (int)(op & 0xff), node_id = (int)((op >> 8) & 0xffff), qpos = (int)(op >> 32);
        const bool is_match = have && kind == (int)C3_CG_MATCH;
        bool eq = false;
        if (is_match) eq = W.nodes[node_id].base == q[qpos];
        const unsigned m_nondel = (C3G_BALLOT(gmask, have && kind != (int)C3_CG_DEL) >> gbase) & C3G_LMASK(GL);
        const unsigned m_eq = (C3G_BALLOT(gmask, eq) >> gbase) & C3G_LMASK(GL);
        const unsigned lower = m_nondel & ((1u << li) - 1u);
        const int pl = lower ? 31 - C3G_CLZ(lower) : -1;    // lane of the previous non-deletion op
        const int pred_node = C3G_SHFL(gmask, node_id, gbase + (pl < 0 ? 0 : pl));
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
        const unsigned m_cx = m_nondel & ~((C3G_BALLOT(gmask, done) >> gbase) & C3G_LMASK(GL));
        C3G_SYNC(gmask);
        if (m_cx) {
            if (li == 0) {
                unsigned mc = m_cx;
                while (mc && !g.err) {
                    const int c = C3G_FFS(mc) - 1; mc &= mc - 1;
                    const unsigned lowc = m_nondel & ((1u << c) - 1u);
                    const int pc = lowc ? 31 - C3G_CLZ(lowc) : -1;
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
                            } else {
                                const int id = c3_g_add_node(g, bq);
                                if (g.err) break;
                                last_p = W.posof[nid]; last_gap = last_p;          // placed right before nid
                                W.gaps[id - n_old] = (uint16_t)last_gap;
                                c3_g_add_edge(g, last_id, id, 0);
                                last_id = id; last_new = 1;
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
                    } else {                                     // insertion: right after the aligned block of last_id
                        const int id = c3_g_add_node(g, q[qp]);
                        if (g.err) break;
                        const c3_pnode nl = g.nodes[last_id];
                        int gap;
                        if (last_id >= n_old) gap = nl.aln_n ? c3g_tail_gap(W, ord, n_old, nl, last_p) : last_gap;
                        else gap = c3g_tail_gap(W, ord, n_old, nl, W.posof[last_id]);
                        last_gap = gap;
                        W.gaps[id - n_old] = (uint16_t)gap;
                        c3_g_add_edge(g, last_id, id, 0);
                        last_id = id; last_new = 1;
                    }
                }
            }
            g.err = C3G_SHFL(gmask, g.err, gbase);
            g.node_n = C3G_SHFL(gmask, g.node_n, gbase);
            g.pool_n = C3G_SHFL(gmask, g.pool_n, gbase);
            last_id = C3G_SHFL(gmask, last_id, gbase);
            last_new = C3G_SHFL(gmask, last_new, gbase);
        }
        if (m_nondel) {                                          // state after the chunk
            const int ln = 31 - C3G_CLZ(m_nondel);
            if ((m_eq >> ln) & 1u) { last_id = C3G_SHFL(gmask, node_id, gbase + ln); last_new = 0; }
        }
        C3G_SYNC(gmask);
        if (g.err) break;
    }
    if (!g.err && li == 0) c3_g_add_edge(g, last_id, C3_SINK, 1 - last_new);
    g.err = C3G_SHFL(gmask, g.err, gbase);
    g.pool_n = C3G_SHFL(gmask, g.pool_n, gbase);
    if (g.err) { C3G_DECLINE(); return C3G_E_RETRY; }
    G.node_n = g.node_n; G.pool_n = g.pool_n;
    C3G_SYNC(gmask);
    return 0;
}

// new order = stable merge of the old order with the new nodes by gap: the t-th new node (id n_old + t, gap g)
// lands at g + t, an old node at position p moves up by the number of new nodes with gap <= p
template <int GL>
C3G_FN void c3g_reorder(c3g_grp &G, const c3g_ws &W, const int li, const unsigned gmask)
{
    const int n_old = G.n, m = G.node_n - n_old;
    const uint16_t *oo = W.order[G.ob];
    uint16_t *on = W.order[G.ob ^ 1];
    const int cs = (n_old + GL - 1) / GL;
    const int ps = li * cs, pe = min(n_old, ps + cs);
    int t = 0;
    if (ps > 0) {                                                // first t with gaps[t] > ps - 1
        int lo = 0, hi = m;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)W.gaps[mid] <= ps - 1) lo = mid + 1; else hi = mid; }
        t = lo;
    }
    for (int p = ps; p < pe; ++p) {
        while (t < m && (int)W.gaps[t] <= p) { on[p + t] = (uint16_t)(n_old + t); W.posof[n_old + t] = (uint16_t)(p + t); ++t; }
        const int id = oo[p];
        on[p + t] = (uint16_t)id; W.posof[id] = (uint16_t)(p + t);
    }
    C3G_SYNC(gmask);
    G.ob ^= 1;
}

// heaviest bundling (abpoa_heaviest_bundling) + consensus walk; one lane.  Returns the length or a negative code.
C3G_FN int c3g_consensus(const c3g_grp &G, const c3_poa_args &A, const c3g_ws &W, char *co)
{
    int32_t *score = reinterpret_cast<int32_t *>(W.desc);
    const uint16_t *ord = W.order[G.ob];
    for (int p = G.node_n - 1; p >= 0; --p) {
        const int v = ord[p];
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

// ---------------------------------------------------------------------------
// Two kernels per alignment, all reads of a wave resident in HBM (state, graph workspace and arena per read):
//   graph kernel  per read: [first launch: first sequence -> graph] or [backtrack + merge + reorder of the alignment
//                 the DP kernel just finished]; then `prepare` for the next sequence.  One warp per read (constant
//                 member masks), few registers, many resident warps.
//   finish kernel once per wave: heaviest bundling + consensus walk + outputs (c3g_finish_body).
//   DP kernel     per read: source row + all rows of the prepared alignment.  One flat loop: a group that finishes
//                 its read fetches the next one while the other groups of the warp keep computing rows.
// Host: graph, (DP, graph) x (most sequences of a read - 1), finish; every launch has its own work counter.
// ---------------------------------------------------------------------------
struct c3g_state {                    // per read (work index), 48 bytes
    int32_t item, sq, nseq, node_n, pool_n, err, ob, qlen, n, w;
    long long cells_total;
};

C3G_FN void c3g_state_load(c3g_grp &G, const c3g_state *S, const c3_poa_args &A)
{
    G.item = S->item; G.sq = S->sq; G.nseq = S->nseq; G.node_n = S->node_n; G.pool_n = S->pool_n; G.err = S->err;
    G.ob = S->ob; G.qlen = S->qlen; G.n = S->n; G.w = S->w; G.cells_total = S->cells_total;
    G.ibase = A.codes + A.item_base[G.item];
    G.bnd = A.bounds + (int64_t)G.item * A.max_seqs * 2;
    G.q = G.ibase + G.bnd[2 * (G.sq < G.nseq ? G.sq : 0)];
}
C3G_FN void c3g_state_store(const c3g_grp &G, c3g_state *S)
{
    S->item = G.item; S->sq = G.sq; S->nseq = G.nseq; S->node_n = G.node_n; S->pool_n = G.pool_n; S->err = G.err;
    S->ob = G.ob; S->qlen = G.qlen; S->n = G.n; S->w = G.w; S->cells_total = G.cells_total;
}

template <int GL>
C3G_FN void c3g_graph_body(const c3g_args &L, uint8_t *smem_warp, const int lane)
{
    const c3_poa_args &A = L.A;
    const c3_poa_para_dev P = A.P;
    const int li = lane & (GL - 1), gbase = lane & ~(GL - 1) & 31, grp = lane / GL;
    const unsigned gmask = C3G_LMASK(GL) << gbase;
    uint8_t *sg = smem_warp + (size_t)grp * C3G_GRAPH_SMEM(GL);
    for (;;) {
        int it = 0;
        if (li == 0) it = (int)C3G_ATOMIC_INC(A.counter);
        it = C3G_SHFL(gmask, it, gbase);
        if (it >= A.n_work) break;
        c3g_state *S = L.state + it;
        const c3g_ws W = c3g_ws_carve(L.ws + (int64_t)it * L.ws_stride, A.node_cap, A.pool_cap, A.cigar_cap);
        uint4 *arena = L.arena + (int64_t)it * L.arena_stride4;
        c3g_grp G;
        if (L.first) {
            c3g_item_begin<GL>(G, A, W, A.order ? A.order[it] : it, li);
            C3G_SYNC(gmask);
        } else {
            c3g_state_load(G, S, A);
            if (G.err || G.sq >= G.nseq) continue;                 // declined earlier / finished earlier
            const int nc = c3g_backtrack<GL>(G, L, P, W, arena, sg, li, gbase, gmask);
            if (nc < 0) { C3G_DECLINE(); G.err = C3G_E_RETRY; }
            else if (c3g_merge<GL>(G, L, W, nc, li, gbase, gmask)) { C3G_DECLINE(); G.err = C3G_E_RETRY; }
            else { c3g_reorder<GL>(G, W, li, gmask); ++G.sq; }
        }
        if (!G.err && G.sq < G.nseq) c3g_prepare<GL>(G, A, P, W, reinterpret_cast<uint16_t *>(sg), li, gbase, gmask);
        if (li == 0) c3g_state_store(G, S);
        C3G_SYNC(gmask);
    }
}

// ---------------------------------------------------------------------------
// finish kernel: heaviest bundling (abpoa_heaviest_bundling) + consensus walk + outputs, once per wave.  8 lanes per
// read, 4 reads per warp in one flat loop with full-warp collectives (like the DP kernel).  Scores are resolved by
// position in reverse order, 8 positions per step: every lane evaluates its node's out-edges each of the 8 turns, the
// turn's lane is final (all its successors sit at higher positions) and its score is handed to the lanes before it.
// ---------------------------------------------------------------------------
#define C3G_FIN_WIN 512               // window of scores by position kept in shared memory (int32)
#define C3G_FIN_SMEM (C3G_FIN_WIN * 4)
#define C3G_FIN_EDGES 8               // out-edges of a node held in registers (more: the read goes to the warp kernel)

C3G_FN void c3g_finish_body(const c3g_args &L, uint8_t *smem_warp, const int lane)
{
    const c3_poa_args &A = L.A;
    const int li = lane & 7, gbase = lane & 24, grp = lane >> 3;
    const unsigned gmask = 0xffu << gbase;
    int32_t *sw = reinterpret_cast<int32_t *>(smem_warp + (size_t)grp * C3G_FIN_SMEM);
    c3g_grp G;
    G.item = -1; G.err = 0; G.node_n = 2; G.ob = 0; G.cells_total = 0;
    c3g_ws W = c3g_ws_carve(L.ws, A.node_cap, A.pool_cap, A.cigar_cap);
    c3g_state *S = L.state;
    bool have = false, exhausted = false;
    int pb = 0;
    for (;;) {
        if (!have && !exhausted) {
            int it = 0;
            if (li == 0) it = (int)C3G_ATOMIC_INC(A.counter);
            it = C3G_SHFL(gmask, it, gbase);
            if (it >= A.n_work) exhausted = true;
            else {
                S = L.state + it;
                c3g_state_load(G, S, A);
                if (!G.err && G.sq >= G.nseq) {
                    W = c3g_ws_carve(L.ws + (int64_t)it * L.ws_stride, A.node_cap, A.pool_cap, A.cigar_cap);
                    have = true;
                    pb = ((G.node_n - 1) / C3G_GL) * C3G_GL;
                }
            }
        }
        if (!C3G_ANYG(C3_FULL, have)) {
            if (!C3G_ANYG(C3_FULL, !exhausted)) break;
            continue;
        }
        // ---- one step of 8 positions (a group without a read runs along on its last inputs and stores nothing) ----
        const bool live = have;
        const int n = G.node_n;
        const uint16_t *ord = W.order[G.ob];
        int32_t *gscore = reinterpret_cast<int32_t *>(W.desc);        // by node id
        uint32_t *nxt = reinterpret_cast<uint32_t *>(W.rowrec);       // by position: position of the chosen successor | base << 16
        const int p = pb + li;
        const bool valid = live && p < n;
        int v = C3_SINK, nout = 0, base = 4;
        int o[C3G_FIN_EDGES], wt[C3G_FIN_EDGES], tp[C3G_FIN_EDGES], sc[C3G_FIN_EDGES];
#pragma unroll
        for (int k = 0; k < C3G_FIN_EDGES; ++k) { o[k] = C3_SINK; wt[k] = -1; tp[k] = -1; sc[k] = 0; }
        if (valid) {
            v = ord[p];
            const c3_nrec nd = c3_ld_node(&W.nodes[v]);
            nout = C3_N_OUTN(nd); base = C3_N_BASE(nd);
            if (nout > C3G_FIN_EDGES) { G.err = C3G_E_RETRY; nout = C3G_FIN_EDGES; }
            if (nout > 0) { o[0] = C3_N_OUT0(nd); wt[0] = C3_N_W0(nd); }
            if (nout > 1) {
                int e = W.nodes[v].out_more;
#pragma unroll
                for (int k = 1; k < C3G_FIN_EDGES; ++k) {
                    if (k < nout && e != (int)C3_NONE) { const c3_pedge pe = W.pool[e]; o[k] = pe.id; wt[k] = pe.w; e = pe.next; }
                }
            }
#pragma unroll
            for (int k = 0; k < C3G_FIN_EDGES; ++k) {
                if (k < nout) {
                    tp[k] = W.posof[o[k]];
                    if (tp[k] >= pb + C3G_GL) sc[k] = (tp[k] - pb < C3G_FIN_WIN - C3G_GL) ? sw[tp[k] & (C3G_FIN_WIN - 1)] : gscore[o[k]];
                }
            }
        }
        int myscore = 0, mybest = 0;
#pragma unroll
        for (int r = C3G_GL - 1; r >= 0; --r) {
            // abpoa_heaviest_bundling: the heaviest out-edge, ties to the later edge whose target scores at least as
            // much; the source takes the heaviest with strictly larger score on ties
            int bk = 0, bw = wt[0], bs = sc[0];
            if (v == C3_SRC) {
#pragma unroll
                for (int k = 1; k < C3G_FIN_EDGES; ++k)
                    if (k < nout && (wt[k] > bw || (wt[k] == bw && sc[k] > bs))) { bk = k; bw = wt[k]; bs = sc[k]; }
            } else {
#pragma unroll
                for (int k = 1; k < C3G_FIN_EDGES; ++k)
                    if (k < nout && (bw < wt[k] || (bw == wt[k] && bs <= sc[k]))) { bk = k; bw = wt[k]; bs = sc[k]; }
            }
            const int mine = (v == C3_SINK || nout == 0) ? 0 : bw + bs;
            if (li == r) { myscore = mine; mybest = bk; }
            const int sr = C3G_SHFL(C3_FULL, mine, gbase + r);
#pragma unroll
            for (int k = 0; k < C3G_FIN_EDGES; ++k) if (tp[k] == pb + r) sc[k] = sr;
        }
        if (valid) {
            int bo = o[0], bt = tp[0];
#pragma unroll
            for (int k = 1; k < C3G_FIN_EDGES; ++k) if (mybest == k) { bo = o[k]; bt = tp[k]; }
            if (v == C3_SINK || nout == 0) { bo = C3_NONE; bt = 0xffff; }
            sw[p & (C3G_FIN_WIN - 1)] = myscore;
            gscore[v] = myscore;
            W.nodes[v].max_out = (uint16_t)bo;
            nxt[p] = (uint32_t)(bt & 0xffff) | ((uint32_t)base << 16);
        }
        G.err = (C3G_BALLOT(C3_FULL, G.err != 0) & gmask) ? C3G_E_RETRY : 0;
        C3G_SYNC(C3_FULL);
        if (have) {
            pb -= C3G_GL;
            if (pb < 0 || G.err) {
                // ---- consensus walk along the chosen successors + outputs (this group alone) ----
                int cons_len = 0;
                if (!G.err && li == 0) {
                    char *co = A.cons + (int64_t)G.item * A.cons_cap;
                    int cur = (int)(nxt[0] & 0xffffu);
                    while (cur != n - 1) {
                        if (cur >= n || cons_len >= A.cons_cap) { cons_len = -1; break; }
                        const uint32_t x = nxt[cur];
                        co[cons_len++] = "ACGTN"[(x >> 16) & 7u];
                        cur = (int)(x & 0xffffu);
                    }
                    if (cons_len >= 0) {
                        const int64_t oo = (int64_t)G.item * A.out_stride;
                        A.status[oo] = 0;
                        A.cons_len[oo] = cons_len;
                        A.nodes_out[oo] = G.node_n;
                        *(long long *)((int32_t *)A.cells_out + (int64_t)G.item * A.cells_stride) = G.cells_total;
                        L.done[G.item] = 1;
                    }
                }
                have = false;
            }
        }
    }
}

template <int RVS, bool MULTI>
C3G_FN void c3g_dp_body(const c3g_args &L, uint8_t *smem_warp, const int lane)
{
    const c3_poa_args &A = L.A;
    const c3_poa_para_dev P = A.P;
    const int li = lane & 7, gbase = lane & 24, grp = lane >> 3;
    const unsigned gmask = 0xffu << gbase;
    uint8_t *sg = smem_warp + (size_t)grp * c3g_smem_group_bytes(RVS);
    uint4 *ring = reinterpret_cast<uint4 *>(sg);
    uint2 *srr = reinterpret_cast<uint2 *>(sg + (C3G_R * 6 * 16 << RVS));
#if defined(__CUDA_ARCH__)
    __builtin_assume(__isShared(ring)); __builtin_assume(__isShared(srr));
#endif
    c3g_grp G;
    G.item = -1; G.err = 0; G.n = 0; G.qlen = 0; G.w = 0; G.cells_total = 0;
    c3g_ws W = c3g_ws_carve(L.ws, A.node_cap, A.pool_cap, A.cigar_cap);
    uint4 *arena = L.arena;
    c3g_state *S = L.state;
    bool have = false, exhausted = false;
    int pos = 0;
    uint2 rprev = make_uint2(0u, 0u);
    uint4 dn = make_uint4(0u, 0u, 0u, 0u);
    for (;;) {
        if (!have && !exhausted) {
            int it = 0;
            if (li == 0) it = (int)C3G_ATOMIC_INC(A.counter);
            it = C3G_SHFL(gmask, it, gbase);
            if (it >= A.n_work) exhausted = true;
            else {
                S = L.state + it;
                c3g_state_load(G, S, A);
                if (!G.err && G.sq < G.nseq) {
                    W = c3g_ws_carve(L.ws + (int64_t)it * L.ws_stride, A.node_cap, A.pool_cap, A.cigar_cap);
                    arena = L.arena + (int64_t)it * L.arena_stride4;
                    c3g_source_row<RVS>(G, L, P, W, ring, srr, arena, rprev, li, gmask);
                    C3G_SYNC(gmask);
                    if (G.err) { if (li == 0) S->err = G.err; }
                    else { have = true; pos = 1; dn = W.desc[1]; }
                }
            }
        }
        if (!C3G_ANYG(C3_FULL, have)) {
            if (!C3G_ANYG(C3_FULL, !exhausted)) break;
            continue;
        }
        {
            const uint4 d = dn;
            if (have) dn = W.desc[pos + 1];
            const int wd = c3g_row<RVS, MULTI>(G, L, P, W, ring, srr, arena, pos, d, rprev, have, li, gbase);
            if (have) {
                G.cells_total += wd;
                ++pos;
                if (G.err || pos >= G.n - 1) {
                    if (li == 0) { S->err = G.err; S->cells_total = G.cells_total; }
                    have = false;
                }
            }
        }
    }
}

#ifdef __CUDACC__
#ifndef C3G_THREADS
#define C3G_THREADS 128
#endif
#ifndef C3G_MINB
#define C3G_MINB 4
#endif
#ifndef C3G_GRAPH_MINB
#define C3G_GRAPH_MINB 6
#endif
template <int RVS, bool MULTI>
__global__ void __launch_bounds__(C3G_THREADS, C3G_MINB) c3_poa_grp_dp_kernel(c3g_args L)
{
    extern __shared__ uint4 c3g_smem[];
    const int wib = threadIdx.x >> 5;
    uint8_t *sw = reinterpret_cast<uint8_t *>(c3g_smem) + (size_t)wib * 4 * c3g_smem_group_bytes(RVS);
    c3g_dp_body<RVS, MULTI>(L, sw, threadIdx.x & 31);
}
__global__ void __launch_bounds__(C3G_THREADS, 4) c3_poa_grp_finish_kernel(c3g_args L)
{
    extern __shared__ uint4 c3g_smem[];
    const int wib = threadIdx.x >> 5;
    c3g_finish_body(L, reinterpret_cast<uint8_t *>(c3g_smem) + (size_t)wib * 4 * C3G_FIN_SMEM, threadIdx.x & 31);
}
#define C3G_GRAPH_GL 32
__global__ void __launch_bounds__(C3G_THREADS, C3G_GRAPH_MINB) c3_poa_grp_graph_kernel(c3g_args L)
{
    extern __shared__ uint4 c3g_smem[];
    const int wib = threadIdx.x >> 5;
    uint8_t *sw = reinterpret_cast<uint8_t *>(c3g_smem) + (size_t)wib * (32 / C3G_GRAPH_GL) * C3G_GRAPH_SMEM(C3G_GRAPH_GL);
    c3g_graph_body<C3G_GRAPH_GL>(L, sw, threadIdx.x & 31);
}
#endif
