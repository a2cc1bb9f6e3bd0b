// grp_emul.cu -- test infrastructure: runs the warp body of c3poa_b200/csrc/poa_grp.cuh (the same code the GPU
// runs, compiled as host code with -DC3G_EMUL) on the fiber warp emulator (warp_emu.cpp), one warp after the
// other, so the group kernel's logic can be checked against the oracle without a GPU.
#define C3G_EMUL 1
#include "../../c3poa_b200/csrc/poa_graph.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" int c3emu_run_warp(void (*body)(void *, int), void *arg);
extern "C" int c3emu_lane(void);
extern "C" void c3g_emul_note(int line)
{
    static const char *v = getenv("C3G_EMUL_VERBOSE");
    if (v && (c3emu_lane() & 7) == 0) fprintf(stderr, "grp_emul: decline at poa_grp.cuh:%d (lane %d)\n", line, c3emu_lane());
}

namespace {
struct WarpArg { const c3g_args *L; uint8_t *smem; };
void warp_lane(void *p, int lane)
{
    const WarpArg *a = (const WarpArg *)p;
    if (a->L->rv_shift == 2) c3g_dp_body<2, true, 4>(*a->L, a->smem, lane);        // 4 lanes per read, 8 reads per warp
    else if (a->L->vs_shift == 3) c3g_dp_body<3, false>(*a->L, a->smem, lane);
    else if (a->L->rv_shift == 3) c3g_dp_body<3, true>(*a->L, a->smem, lane);
    else c3g_dp_body<4, true>(*a->L, a->smem, lane);
}
}

extern "C" int c3g_emul_batch(int n_items, const uint8_t *codes, const int64_t *item_base, const int32_t *bounds,
                              const int32_t *n_seqs, int max_seqs, int min_seqs, int msa2,
                              int match, int mismatch, int o1, int e1, int o2, int e2, int wb, double wf, int simd_bits,
                              int node_cap, int cigar_cap, int qp_stride, int vs_shift, int rv_shift, int n_warps, int graph_gl,
                              char *cons, int cons_cap, int32_t *status, int32_t *cons_len, int32_t *nodes_out,
                              long long *cells_out, int32_t *done)
{
    c3g_args L;
    memset(&L, 0, sizeof(L));
    c3_poa_args &A = L.A;
    A.codes = codes; A.item_base = item_base; A.bounds = bounds; A.n_seqs = n_seqs; A.n_seqs_stride = 1;
    A.n_items = n_items; A.max_seqs = max_seqs; A.min_seqs = min_seqs; A.msa2 = msa2; A.ok_status = 0;
    A.P.match = match; A.P.mismatch = mismatch; A.P.o1 = o1; A.P.e1 = e1; A.P.o2 = o2; A.P.e2 = e2;
    A.P.wb = wb; A.P.wf = wf; A.P.simd_bits = simd_bits;
    A.node_cap = node_cap; A.pool_cap = node_cap; A.cell_cap = 0; A.cigar_cap = cigar_cap; A.qp_stride = qp_stride;
    A.cons = cons; A.cons_cap = cons_cap; A.status = status; A.cons_len = cons_len; A.nodes_out = nodes_out;
    A.cells_out = cells_out; A.out_stride = 1; A.cells_stride = 2;
    A.n_work = n_items;
    int max_nseq = 1;
    for (int i = 0; i < n_items; ++i) if (n_seqs[i] > max_nseq) max_nseq = n_seqs[i];
    std::vector<unsigned> counters((size_t)2 * max_nseq + 3, 0u);
    const int64_t ws_bytes = c3g_ws_bytes(node_cap, node_cap, cigar_cap, qp_stride);
    const int64_t arena4 = ((int64_t)node_cap << vs_shift) * 3;
    uint8_t *ws = (uint8_t *)aligned_alloc(256, (size_t)ws_bytes * n_items);
    uint4 *arena = (uint4 *)aligned_alloc(256, (size_t)arena4 * 16 * n_items);
    size_t smw = (size_t)(rv_shift == 2 ? 8 : 4) * c3g_smem_group_bytes(rv_shift);
    uint8_t *smem = (uint8_t *)aligned_alloc(256, (smw * n_warps + 255) & ~(size_t)255);
    std::vector<c3g_state> state((size_t)n_items);
    if (!ws || !arena || !smem) return -1;
    memset(ws, 0xA5, (size_t)ws_bytes * n_items);               // poison: nothing may depend on zeroed memory
    memset(arena, 0xA5, (size_t)arena4 * 16 * n_items);
    memset(smem, 0xA5, smw * n_warps);
    memset(state.data(), 0xA5, state.size() * sizeof(c3g_state));
    L.ws = ws; L.ws_stride = ws_bytes; L.arena = arena; L.arena_stride4 = arena4;
    L.vs_shift = vs_shift; L.rv_shift = rv_shift; L.done = done; L.state = state.data();
    L.eager = graph_gl & 1;            // (the former graph_gl argument: odd = the small-wave variant of the backtrack)
    int rc = 0, launch = 0;
    // the host's launch sequence: init kernel, then (DP kernel, graph kernel) per further sequence.  Init and graph
    // kernel are scalar code, one worker per read: called directly.  The DP kernel's warps run on the fiber emulator
    // one after the other (they only share the launch's work counter).
    auto init = [&]() {
        for (int it = 0; it < n_items; ++it) {
            const c3g_ws W = c3g_ws_carve(L.ws + (int64_t)it * L.ws_stride, A.node_cap, A.pool_cap, A.cigar_cap);
            c3g_grp G;
            c3s_item_begin(G, A, A.P, W, it, 0, 1);
            if (!G.err && G.sq >= G.nseq) c3s_finish(G, L, W);
            c3g_state_store(G, L.state + it);
        }
    };
    auto graph = [&]() { for (int it = 0; it < n_items; ++it) c3s_graph_step(L, it, true); };
    auto dp = [&]() {
        L.A.counter = &counters[launch++];
        for (int w = 0; w < n_warps && !rc; ++w) {
            WarpArg a{&L, smem + smw * w};
            rc = c3emu_run_warp(warp_lane, &a);
        }
    };
    init();
    for (int sq = 1; sq < max_nseq && !rc; ++sq) { dp(); graph(); }
    free(ws); free(arena); free(smem);
    return rc;
}
