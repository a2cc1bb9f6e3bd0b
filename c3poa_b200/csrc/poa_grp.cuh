// poa_grp.cuh -- stage 3b, group DP kernel: 8 lanes per read, 4 reads per warp.
//
// Same algorithm and the same results as poa.cuh / poa_lane.cuh (abPOA 1.0.5 semantics; replaces
// poa.msa_aligner(match=5).msa(...), /root/reference/bin/determine_consensus.py:30-47), third mapping:
//
//   poa.cuh       one warp per read: half the lanes idle on a ~70-column band and every row pays the warp-uniform
//                 bookkeeping 32 times over.
//   poa_lane.cuh  one thread per read for everything: 240 KB of thread-private state per read and 6 B per DP cell; the
//                 DP (the parallel part) runs at thread speed.
//   poa_grp.cuh + poa_graph.cuh   the work is split by its nature, two kernels per alignment, all reads of a wave
//                 resident in HBM (state, graph workspace and DP arena per read: a B200 holds a whole 100k-read batch):
//     * DP (this file): one 8-lane group per read.  Everything a read owns is laid out BY POSITION in the
//       topological order, so the DP's accesses are sequential and coalesced.  Each lane owns one 16-column vector of
//       the band (abPOA's int16 SIMD granule), lane = vector index mod 8, cells packed two per register
//       (VIADDMNMX.S16x2 / VIMNMX3.S16x2); the horizontal gap is an in-lane chain plus a 3-step (max,+) scan over
//       the 8 lanes; predecessor rows come from a 4-row shared-memory ring (H, E1, E2 as int16) or, further back,
//       from the arena.  One flat loop: a group that finishes its read fetches the next one while the other groups
//       of the warp keep computing rows; every collective names the full warp.
//     * arena: 3 bytes per cell -- H as int16 plus one byte holding H-E1 (3 bits) and H-E2 (5 bits), stored
//       complemented (that is what one packed add of ~H yields); both differences are bounded by the gap-open
//       costs, so H, E1 and E2 are recovered exactly and the backtrack stays abPOA's value-based one.  Rows have a
//       fixed stride of VS vectors, vector v of a row at slot v mod VS: a cell's address needs no row record.
//     * serial phases (poa_graph.cuh): backtrack, merge, new order, row descriptors of the next alignment, heaviest
//       bundling -- one THREAD per read, all reads at once.
//
// Scope: int16 score mode with the 256-bit granule, banded, default-sized gap costs (o1+e1 <= 7, o2+e2 <= 31),
// consensus output.  Anything else -- and any capacity overflow or a row that fails the int16 exactness guard --
// leaves the item not-done; the caller then runs it through c3_poa_kernel, which owns all error reporting.
//
// With -DC3G_EMUL the DP body is host code as well: tests/emul/grp_emul.cu runs it on the fiber warp emulator
// (tests/emul/warp_emu.cpp) and calls the scalar phases directly, against the oracle, no GPU needed.
#pragma once
#include "poa_lane.cuh"

#define C3G_R 4                       // ring slots (rows) per group
#define C3G_GL 8                      // lanes per group (the usual instantiation; GL = 4 for short sequences, see c3g_dp_body)
#define C3G_E_RETRY (-298)
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
    int eager;                        // graph kernel: 1 = small wave (bound by round trips): the backtrack asks early for what its gap tests may need
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

// score profile of the read's current sequence: qp[b][j] = score of node base b against column j (= q[j-1]); column 0
// and the padding score 0; row 4 (an N node) scores 0 against everything.  The 8 lanes of the group write 32 columns
// per step.
template <int GL>
C3G_FN void c3g_build_qp(const c3g_grp &G, const c3_poa_args &A, const c3_poa_para_dev &P, const c3g_ws &W, const int li)
{
    const int qs = A.qp_stride, qlen = G.qlen;
    const uint8_t *q = G.q;
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
        *reinterpret_cast<uint32_t *>(W.qp + 4 * qs + j0) = 0u;
    }
}

// source row: cells, ring slot 0, arena, record
template <int RVS, int GL>
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
    for (int sn = li; sn <= end_sn; sn += GL) {
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
template <int RVS, bool MULTI, int GL>
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
    if (end_sn - beg_sn + 1 > (MULTI ? (1 << L.vs_shift) : GL)) { bad = true; end_sn = beg_sn; }   // wider than an arena row
    const int nvec = end_sn - beg_sn + 1;
    const bool to_ring = MULTI ? nvec <= (1 << RVS) : true;
    const uint32_t ne1 = C3L_PACK2(-e1, -e1), ne2 = C3L_PACK2(-e2, -e2), noe1 = C3L_PACK2(-oe1, -oe1), noe2 = C3L_PACK2(-oe2, -oe2);
    const int d21 = oe1 - oe2;
    int pc1 = C3L_FLOOR, pc2 = C3L_FLOOR;                          // F1, F2' entering the pass (chain domain, see below)
    uint32_t mcarry = C3L_FLOOR2;                                  // merged predecessor H of the vector before the pass
    int bestkey = -0x7fffffff - 1;      // row arg-max key, compared as a signed word (see below)
    int h_first = 0x7fff;
    // everything read from the ring must be in registers before any lane overwrites the slot of row pos - C3G_R:
    // a row of several passes stores its first vectors before it has read the last ones, so it does not use that slot
    const int ring_delta = (!MULTI || nvec <= GL) ? C3G_R : C3G_R - 1;
    int sn0 = beg_sn;
    do {
        const int l = (li - sn0) & (GL - 1), sn = sn0 + l;
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
        uint32_t top = (uint32_t)C3G_SHFL(C3_FULL, h[7], gbase + ((li - 1) & (GL - 1)));
        if (l == 0) top = mcarry;
        if (MULTI) mcarry = (uint32_t)C3G_SHFL(C3_FULL, h[7], gbase + ((sn0 + GL - 1) & (GL - 1)));      // (every lane of the warp: no group-dependent condition around a collective)
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
        for (int dd = 1; dd < GL; dd <<= 1) {
            const uint32_t v = (uint32_t)C3G_SHFL(C3_FULL, tt, gbase + ((li - dd) & (GL - 1)));
            if (l >= dd) tt = C3L_VMAX2(tt, v);
        }
        const uint32_t ex = (uint32_t)C3G_SHFL(C3_FULL, tt, gbase + ((li - 1) & (GL - 1)));
        int c1 = (int)(int16_t)(ex & 0xffffu) - 16 * e1 * (l - 1), c2 = ((int)ex >> 16) - 16 * e2 * (l - 1);
        if (MULTI) {
            c1 = l == 0 ? pc1 : max(c1, pc1 - 16 * e1 * l);
            c2 = l == 0 ? pc2 : max(c2, pc2 - 16 * e2 * l);
            const uint32_t tot = (uint32_t)C3G_SHFL(C3_FULL, tt, gbase + ((sn0 + GL - 1) & (GL - 1)));
            pc1 = max((int)(int16_t)(tot & 0xffffu) - 16 * e1 * (GL - 1), pc1 - 16 * e1 * GL);
            pc2 = max(((int)tot >> 16) - 16 * e2 * (GL - 1), pc2 - 16 * e2 * GL);
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
        // simd_abpoa_ada_max_i as one packed max: the value in the high half, tie-break priority in the low half (lowest
        // SIMD lane, then the last vector, then the earliest vector); compared as SIGNED 32-bit words: the high half is
        // the signed value, and between equal values the low halves compare as the unsigned numbers they are
        if (act) {
            int lk = -0x7fffffff - 1;
            const int lim = min(qlen, end_sn * 16 + 15) - j0;        // last column of this vector that exists (>= 15: all)
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int klo = (int)C3L_PRMT(hh[t], (uint32_t)((15 - 2 * t) << 12), 0x1054u);   // lo half -> high half, priority below
                const int khi = (int)C3L_PRMT(hh[t], (uint32_t)((14 - 2 * t) << 12), 0x3254u);
                lk = max(lk, max(klo, khi));
            }
            if (lim < 15) {                                          // the row's last vector when it reaches past qlen: again, masked
                lk = -0x7fffffff - 1;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    int klo = (int)C3L_PRMT(hh[t], (uint32_t)((15 - 2 * t) << 12), 0x1054u);
                    int khi = (int)C3L_PRMT(hh[t], (uint32_t)((14 - 2 * t) << 12), 0x3254u);
                    if (2 * t > lim) klo = -0x7fffffff - 1;
                    if (2 * t + 1 > lim) khi = -0x7fffffff - 1;
                    lk = max(lk, max(klo, khi));
                }
            }
            const int vp = (sn == end_sn) ? 0xfff : (0xffe - (sn - beg_sn));
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
        sn0 += GL;
    } while (MULTI && C3G_ANYG(C3_FULL, sn0 <= end_sn));
    // exactness of the int16 form (see poa_lane.cuh): the first band cell comfortably above the floor means that
    // every cell of the row is reachable and was never clamped.  The verdict rides on the arg-max reduction.
    {
        const int end = min(qlen, end_sn * 16 + 15);
        const int D = max(min(oe1, oe2), max(e1, e2));
        if (bad || h_first < C3L_FLOOR + C3L_LOW_GUARD + oe2 + D * (end - (beg_sn << 4) + 1)) bestkey = 0x7fffffff;
    }
#pragma unroll
    for (int dd = 1; dd < GL; dd <<= 1) bestkey = max(bestkey, (int)C3G_SHFL(C3_FULL, bestkey, gbase + (li ^ dd)));
    if (bestkey == 0x7fffffff && live) { C3G_DECLINE(); G.err = C3G_E_RETRY; }
    int best_i = -1;
    if ((bestkey >> 16) > C3L_FLOOR) {
        const int sl = 15 - ((bestkey >> 12) & 15);
        const int vp = bestkey & 0xfff;
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
// `cs`: streaming (evict-first) loads of the arena cells.  A backtrack touches a cell's sector once; in a wave large
// enough to be bound by DRAM transactions (100 000 reads: -3.4 %) that keeps the sequential streams (descriptors, row
// records, cigar) in the few L1 lines a thread can count on; in smaller waves it costs (40 000 reads: +5 %)
#if defined(__CUDA_ARCH__)
#define C3G_LDCS(p) __ldcs(p)
#else
#define C3G_LDCS(p) (*(p))
#endif
C3_HD __forceinline__ int c3g_cell_h(const uint4 *arena, const int vs_shift, const int pos, const int j, const bool cs = false)
{
    const int16_t *p = reinterpret_cast<const int16_t *>(arena + (((int64_t)pos << vs_shift) + ((j >> 4) & ((1 << vs_shift) - 1))) * 3);
    return c3l_map(cs ? (int)C3G_LDCS(p + (j & 15)) : (int)p[j & 15]);
}
C3_HD __forceinline__ int c3g_cell_eb(const uint4 *arena, const int vs_shift, const int pos, const int j, const bool cs = false)
{
    const uint8_t *p = reinterpret_cast<const uint8_t *>(arena + (((int64_t)pos << vs_shift) + ((j >> 4) & ((1 << vs_shift) - 1))) * 3 + 2);
    return (int)(~(cs ? C3G_LDCS(p + (j & 15)) : p[j & 15]) & 0xff);               // (H - E1) | (H - E2) << 3
}
C3_HD __forceinline__ int c3g_pred_pos(const c3g_ws &W, const uint4 d, const int k)
{
    return k == 0 ? C3G_D_P0(d) : k == 1 ? C3G_D_P1(d) : (int)W.xpred[C3G_D_XOFS(d) + k - 2];
}

// ---------------------------------------------------------------------------
// Per-read state carried between the launches of a wave.
// Host: graph kernel (first), then (DP kernel, graph kernel) x (most sequences of a read - 1).
// ---------------------------------------------------------------------------
struct c3g_state {                    // per read (work index), 48 bytes
    int32_t item, sq, nseq, node_n, pool_n, err, ob, qlen, n, w;
    long long cells_total;
};

C3_HD inline void c3g_state_load(c3g_grp &G, const c3g_state *S, const c3_poa_args &A)
{
    G.item = S->item; G.sq = S->sq; G.nseq = S->nseq; G.node_n = S->node_n; G.pool_n = S->pool_n; G.err = S->err;
    G.ob = S->ob; G.qlen = S->qlen; G.n = S->n; G.w = S->w; G.cells_total = S->cells_total;
    G.ibase = A.codes + A.item_base[G.item];
    G.bnd = A.bounds + (int64_t)G.item * A.max_seqs * 2;
    G.q = G.ibase + G.bnd[2 * (G.sq < G.nseq ? G.sq : 0)];
}
C3_HD inline void c3g_state_store(const c3g_grp &G, c3g_state *S)
{
    S->item = G.item; S->sq = G.sq; S->nseq = G.nseq; S->node_n = G.node_n; S->pool_n = G.pool_n; S->err = G.err;
    S->ob = G.ob; S->qlen = G.qlen; S->n = G.n; S->w = G.w; S->cells_total = G.cells_total;
}

// GL lanes per read, 32 / GL reads per warp.  GL = 8 is the usual shape (a ~70-column band spans 4-6 vectors); GL = 4 with
// a 4-vector ring (RVS = 2) serves short sequences, whose ~40-column bands span 3-4 vectors: eight reads per warp instead
// of four at the same instruction count per row step (a row of 5-8 vectors takes a second pass).
template <int RVS, bool MULTI, int GL = C3G_GL>
C3G_FN void c3g_dp_body(const c3g_args &L, uint8_t *smem_warp, const int lane)
{
    const c3_poa_args &A = L.A;
    const c3_poa_para_dev P = A.P;
    const int li = lane & (GL - 1), gbase = lane & (32 - GL), grp = lane / GL;
    const unsigned gmask = ((1u << GL) - 1u) << gbase;
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
                    c3g_build_qp<GL>(G, A, P, W, li);
                    c3g_source_row<RVS, GL>(G, L, P, W, ring, srr, arena, rprev, li, gmask);
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
            const int wd = c3g_row<RVS, MULTI, GL>(G, L, P, W, ring, srr, arena, pos, d, rprev, have, li, gbase);
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
template <int RVS, bool MULTI, int GL = C3G_GL>
__global__ void __launch_bounds__(C3G_THREADS, C3G_MINB) c3_poa_grp_dp_kernel(c3g_args L)
{
    extern __shared__ uint4 c3g_smem[];
    const int wib = threadIdx.x >> 5;
    uint8_t *sw = reinterpret_cast<uint8_t *>(c3g_smem) + (size_t)wib * (32 / GL) * c3g_smem_group_bytes(RVS);
    c3g_dp_body<RVS, MULTI, GL>(L, sw, threadIdx.x & 31);
}
#endif
