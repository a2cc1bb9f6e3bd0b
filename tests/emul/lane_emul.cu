// lane_emul.cu -- test infrastructure: runs the per-thread phases of c3poa_b200/csrc/poa_lane.cuh on the
// CPU, 32 states in lockstep exactly as c3_poa_lane_kernel sequences them, so the lane kernel's logic can
// be checked against the oracle without a GPU.  Built by tests/test_lane_emul.py with nvcc (host code only
// is executed).
#include "../../c3poa_b200/csrc/poa_lane.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" int c3l_emul_batch(int n_items, const uint8_t *codes, const int64_t *item_base, const int32_t *bounds,
                              const int32_t *n_seqs, int max_seqs, int min_seqs, int msa2,
                              int match, int mismatch, int o1, int e1, int o2, int e2, int wb, double wf, int simd_bits,
                              int node_cap, int cigar_cap, int qp_stride, int arena_cap4, int sm_vec,
                              char *cons, int cons_cap, int32_t *status, int32_t *cons_len, int32_t *nodes_out,
                              long long *cells_out, int32_t *done)
{
    c3_poa_args A;
    memset(&A, 0, sizeof(A));
    A.codes = codes; A.item_base = item_base; A.bounds = bounds; A.n_seqs = n_seqs; A.n_seqs_stride = 1;
    A.n_items = n_items; A.max_seqs = max_seqs; A.min_seqs = min_seqs; A.msa2 = msa2; A.ok_status = 0;
    A.P.match = match; A.P.mismatch = mismatch; A.P.o1 = o1; A.P.e1 = e1; A.P.o2 = o2; A.P.e2 = e2;
    A.P.wb = wb; A.P.wf = wf; A.P.simd_bits = simd_bits;
    A.node_cap = node_cap; A.pool_cap = node_cap; A.cell_cap = 0; A.cigar_cap = cigar_cap; A.qp_stride = qp_stride;
    A.cons = cons; A.cons_cap = cons_cap; A.status = status; A.cons_len = cons_len; A.nodes_out = nodes_out;
    A.cells_out = cells_out; A.out_stride = 1; A.cells_stride = 2;
    A.n_work = n_items;
    const int64_t ws_bytes = c3_poa_ws_bytes(node_cap, node_cap, 0, cigar_cap, qp_stride);
    uint8_t *ws = (uint8_t *)aligned_alloc(256, (size_t)ws_bytes * 32);
    int32_t *ar = (int32_t *)aligned_alloc(256, (size_t)arena_cap4 * 16);
    std::vector<uint4> smbuf((size_t)(sm_vec > 0 ? sm_vec : 1) * C3L_RSLOT * 32);
    uint4 *sm = smbuf.data();
    if (!ws || !ar) return -1;
    memset(ws, 0, (size_t)ws_bytes * 32);
    A.ws = ws; A.ws_stride = ws_bytes;
    const c3_poa_para_dev P = A.P;
    c3l_state S[32];
    c3_poa_ws W[32];
    for (int l = 0; l < 32; ++l) W[l] = c3_poa_ws_carve(ws + (int64_t)l * ws_bytes, node_cap, node_cap, 0, cigar_cap);
    for (int first = 0; first < n_items; first += 32) {
        int max_nseq = 0;
        for (int l = 0; l < 32; ++l) {
            const bool have = first + l < n_items;
            c3l_item_begin(S[l], A, W[l], have ? first + l : 0, have);
            if (S[l].on && !S[l].err && S[l].nseq > max_nseq) max_nseq = S[l].nseq;
        }
        for (int sq = 1; sq < max_nseq; ++sq) {
            int mv = 0;
            for (int l = 0; l < 32; ++l) { const int nv = c3l_align_begin(S[l], A, P, W[l], sq); if (nv > mv) mv = nv; }
            if (mv == 0) continue;
            int used4 = mv * C3L_VSTRIDE;
            if (used4 > arena_cap4) { for (int l = 0; l < 32; ++l) if (S[l].aligning) { S[l].err = C3L_E_RETRY; S[l].aligning = 0; } continue; }
            for (int l = 0; l < 32; ++l) c3l_source_row(S[l], P, W[l], ar, l);
            for (;;) {
                mv = 0;
                for (int l = 0; l < 32; ++l) { const int nv = c3l_row_setup(S[l], A, W[l]); if (nv > mv) mv = nv; }
                if (mv == 0) break;
                if (used4 + mv * C3L_VSTRIDE > arena_cap4) { for (int l = 0; l < 32; ++l) if (S[l].aligning) S[l].err = C3L_E_RETRY; break; }
                for (int l = 0; l < 32; ++l) c3l_row_compute(S[l], A, P, W[l], ar, used4, l, sm, sm_vec);
                used4 += mv * C3L_VSTRIDE;
            }
            for (int l = 0; l < 32; ++l) c3l_align_end(S[l], A, P, W[l], ar, l, sq);
        }
        for (int l = 0; l < 32; ++l) c3l_item_end(S[l], A, W[l], done);
    }
    free(ws); free(ar);
    return 0;
}
