// grp_emul.cu -- test infrastructure: runs the warp body of c3poa_b200/csrc/poa_grp.cuh (the same code the GPU
// runs, compiled as host code with -DC3G_EMUL) on the fiber warp emulator (warp_emu.cpp), one warp after the
// other, so the group kernel's logic can be checked against the oracle without a GPU.
#define C3G_EMUL 1
#include "../../c3poa_b200/csrc/poa_grp.cuh"
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
struct WarpArg { const c3g_args *L; uint8_t *smem; int gwarp; };
void warp_lane(void *p, int lane)
{
    const WarpArg *a = (const WarpArg *)p;
    c3g_warp_body(*a->L, a->smem, a->gwarp, lane);
}
}

extern "C" int c3g_emul_batch(int n_items, const uint8_t *codes, const int64_t *item_base, const int32_t *bounds,
                              const int32_t *n_seqs, int max_seqs, int min_seqs, int msa2,
                              int match, int mismatch, int o1, int e1, int o2, int e2, int wb, double wf, int simd_bits,
                              int node_cap, int cigar_cap, int qp_stride, int vs_shift, int rv_shift, int n_warps,
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
    unsigned counter = 0;
    A.counter = &counter;
    const int n_groups = n_warps * 4;
    const int64_t ws_bytes = c3g_ws_bytes(node_cap, node_cap, cigar_cap, qp_stride);
    const int64_t arena4 = ((int64_t)node_cap << vs_shift) * 3;
    uint8_t *ws = (uint8_t *)aligned_alloc(256, (size_t)ws_bytes * n_groups);
    uint4 *arena = (uint4 *)aligned_alloc(256, (size_t)arena4 * 16 * n_groups);
    const size_t smw = (size_t)4 * c3g_smem_group_bytes(rv_shift);
    uint8_t *smem = (uint8_t *)aligned_alloc(256, (smw * n_warps + 255) & ~(size_t)255);
    if (!ws || !arena || !smem) return -1;
    memset(ws, 0xA5, (size_t)ws_bytes * n_groups);              // poison: nothing may depend on zeroed memory
    memset(arena, 0xA5, (size_t)arena4 * 16 * n_groups);
    memset(smem, 0xA5, smw * n_warps);
    L.ws = ws; L.ws_stride = ws_bytes; L.arena = arena; L.arena_stride4 = arena4;
    L.vs_shift = vs_shift; L.rv_shift = rv_shift; L.done = done;
    int rc = 0;
    // warps run one after the other; they only share the work counter
    for (int w = 0; w < n_warps && !rc; ++w) {
        WarpArg a{&L, smem + smw * w, w};
        rc = c3emu_run_warp(warp_lane, &a);
    }
    free(ws); free(arena); free(smem);
    return rc;
}
