// c3poa_gpu.cu -- C ABI (include/c3poa_gpu.h) over the sm_100a kernels.
// No torch types, no CPU fallback: every entry point needs a working device.
#include "../../include/c3poa_gpu.h"
#include "common.cuh"
#include "conk.cuh"
#include "peaks.cuh"
#include "poa.cuh"
#include "poa_lane.cuh"
#include "poa_grp.cuh"
#include "poa_graph.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <vector>

#define C3_VERSION "c3poa_b200 0.1.0 (sm_100a)"

#define C3_PIPE_CHUNKS 8               // chunks of the fused call's read upload (copy stream under the conk kernel)

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct c3_handle {
    int device = 0, sm_count = 0;
    size_t total_mem = 0;
    cudaStream_t stream = nullptr;
    char err[512] = {0};
    cudaEvent_t ev[8] = {nullptr};
    // Host waits go through an event created with cudaEventBlockingSync: the waiting thread sleeps instead of spinning, so
    // a handle does not burn a core per batch in flight (the driver's reader / writer threads want them).
    cudaEvent_t ev_wait = nullptr;
    cudaStream_t stream2 = nullptr; cudaEvent_t ev_chunk[C3_PIPE_CHUNKS] = {nullptr};     // copy stream of the fused call
    cudaError_t sync()
    {
        if (!ev_wait) return cudaStreamSynchronize(stream);
        cudaError_t e = cudaEventRecord(ev_wait, stream);
        return e != cudaSuccess ? e : cudaEventSynchronize(ev_wait);
    }
    c3_timings tim{};
    // per-launch timing of the POA kernels: event pairs recorded around every launch, summed by kind after the call's
    // final synchronisation (kt_collect)
    std::vector<cudaEvent_t> kt_ev; std::vector<int> kt_kind; size_t kt_used = 0;
    void kt_begin() { kt_mark(-1); }
    void kt_end(int kind) { kt_mark(kind); }
    void kt_mark(int kind)
    {
        if (kt_used == kt_ev.size()) { cudaEvent_t e; cudaEventCreate(&e); kt_ev.push_back(e); kt_kind.push_back(0); }
        cudaEventRecord(kt_ev[kt_used], stream); kt_kind[kt_used] = kind; ++kt_used;
    }
    void kt_collect()
    {
        for (size_t k = 1; k < kt_used; ++k) {
            if (kt_kind[k] < 0) continue;
            float t = 0.f; cudaEventElapsedTime(&t, kt_ev[k - 1], kt_ev[k]);
            switch (kt_kind[k]) {
            case 0: tim.poa_dp_ms += t; tim.poa_dp_launches++; break;
            case 1: tim.poa_graph_ms += t; tim.poa_graph_launches++; break;
            case 2: tim.poa_warp_ms += t; break;
            default: tim.poa_lane_ms += t; break;
            }
        }
        kt_used = 0;
    }
    // staged batch
    int n_reads = 0, n_splints = 0, max_lr = 0, max_ls = 0, max_peaks = 0, cons_cap = 0;
    int64_t total_bases = 0, total_sp = 0;
    bool staged = false, ran = false;
    DevBuf d_ascii, d_codes, d_off, d_sp_ascii, d_sp_codes, d_sp_off, d_sp_idx;
    DevBuf d_prof, d_brow, d_counter, d_coef, d_pk_scratch, d_smoothed, d_median;
    DevBuf d_peaks, d_npk, d_sub, d_dang, d_res, d_stats, d_cons, d_ws;
    // B3 staging
    DevBuf d_item_base, d_bounds, d_nseq, d_status, d_clen, d_nodes, d_cells, d_order;
    // POA kernel choice: 0 auto (group kernel for every eligible read), 1 warp kernel only, 2 lane kernel whenever
    // eligible, 3 group kernel whenever eligible (= auto); the warp kernel always takes what the others leave
    int poa_mode = 0;
    double dbg_sync_ms = 0; std::chrono::steady_clock::time_point dbg_t1;     // C3POA_GRP_TIMING only
    int grp_bps[3] = {0, 0, 0}, warp_bps = 0, grp_grow_idx = 0, grp_ask_wait = 0;
    int sw_int8_lanes = 0, sw_end_clamp = 0;   // c3_set_abpoa_switches: only the warp kernel implements them
    DevBuf d_order_lane, d_done, d_order_grp, d_ws_grp, d_order_scratch, d_pairs;
    int n_work_grp = 0, grp_max_nseq = 0, grp_max_q = 0; int64_t grp_max_total = 0;
    std::vector<int32_t> grp_nseq;   // sequences per item of the group kernels' list (host copy: launches per wave)
    int n_work_lane = 0, lane_items = 0, lane_n_items = 0;
    int64_t lane_max_total = 0; int lane_max_nseq = 0, lane_max_q = 0;
    double lane_cost_spread = 1.0;       // estimated cost of the largest lane item / of the median one
};

static int fail(c3_handle *h, int code, const char *fmt, ...)
{
    if (h) {
        va_list ap; va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(h, -1, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" const char *c3_version(void) { return C3_VERSION; }

extern "C" int c3_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" void c3_default_poa_params(c3_poa_params *p)
{
    p->match = 5; p->mismatch = 4; p->gap_open1 = 4; p->gap_ext1 = 2; p->gap_open2 = 24; p->gap_ext2 = 1;
    p->wb = 10; p->simd_bits = 256; p->wf = 0.01;
}

extern "C" int c3_init(int device, c3_handle **out)
{
    if (!out) return -1;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return -2;   // no CPU fallback
    c3_handle *h = new c3_handle();
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return -3; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return -3; }
    h->sm_count = prop.multiProcessorCount;
    h->total_mem = prop.totalGlobalMem;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return -4; }
    for (int i = 0; i < 8; ++i) cudaEventCreate(&h->ev[i]);
    if (cudaEventCreateWithFlags(&h->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) { h->ev_wait = nullptr; (void)cudaGetLastError(); }
    *out = h;
    return 0;
}

extern "C" void c3_destroy(c3_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    h->sync();
    DevBuf *bufs[] = {&h->d_ascii, &h->d_codes, &h->d_off, &h->d_sp_ascii, &h->d_sp_codes, &h->d_sp_off, &h->d_sp_idx,
                      &h->d_prof, &h->d_brow, &h->d_counter, &h->d_coef, &h->d_pk_scratch, &h->d_smoothed, &h->d_median,
                      &h->d_peaks, &h->d_npk, &h->d_sub, &h->d_dang, &h->d_res, &h->d_stats, &h->d_cons, &h->d_ws,
                      &h->d_item_base, &h->d_bounds, &h->d_nseq, &h->d_status, &h->d_clen, &h->d_nodes, &h->d_cells, &h->d_order,
                      &h->d_order_lane, &h->d_done, &h->d_order_grp, &h->d_ws_grp, &h->d_order_scratch, &h->d_pairs};
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < 8; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (cudaEvent_t e : h->kt_ev) cudaEventDestroy(e);
    if (h->ev_wait) cudaEventDestroy(h->ev_wait);
    for (int k = 0; k < C3_PIPE_CHUNKS; ++k) if (h->ev_chunk[k]) cudaEventDestroy(h->ev_chunk[k]);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" void *c3_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}

extern "C" void c3_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" const char *c3_last_error(const c3_handle *h) { return h ? h->err : "null handle"; }

extern "C" int c3_set_poa_mode(c3_handle *h, int32_t mode)
{
    if (!h) return -1;
    if (mode < 0 || mode > 3) return fail(h, -5, "poa mode must be 0 (auto), 1 (warp kernel), 2 (lane kernel) or 3 (group kernel)");
    h->poa_mode = mode;
    return 0;
}

// reads of the last B3/B4 call that the lane kernel was given / finished (the rest went to the warp kernel)
extern "C" int c3_set_abpoa_switches(c3_handle *h, int32_t int8_lanes, int32_t end_clamp)
{
    if (!h) return -1;
    h->sw_int8_lanes = int8_lanes ? 1 : 0; h->sw_end_clamp = end_clamp ? 1 : 0;
    return 0;
}

extern "C" int c3_lane_counts(c3_handle *h, int32_t *out_given, int32_t *out_done)
{
    if (!h || !out_given || !out_done) return -1;
    *out_given = h->lane_items; *out_done = 0;
    if (h->lane_items <= 0) return 0;
    CK(cudaSetDevice(h->device));
    std::vector<int32_t> d(h->d_done.cap / 4);
    const size_t n = std::min(d.size(), (size_t)h->lane_n_items);
    CK(cudaMemcpy(d.data(), h->d_done.p, n * 4, cudaMemcpyDeviceToHost));
    int c = 0;
    for (size_t i = 0; i < n; ++i) c += d[i] != 0;
    *out_done = c;
    return 0;
}

extern "C" int c3_get_timings(const c3_handle *h, c3_timings *out)
{
    if (!h || !out) return -1;
    *out = h->tim;
    return 0;
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
static int launch_encode(c3_handle *h, const void *in, void *out, int64_t n)
{
    if (n <= 0) return 0;
    int blocks = (int)std::min<int64_t>((n / 16 + 255) / 256 + 1, (int64_t)h->sm_count * 8);
    c3_encode_kernel<<<blocks, 256, 0, h->stream>>>((const uint8_t *)in, (uint8_t *)out, n);
    CK(cudaGetLastError());
    h->tim.kernel_launches++;
    return 0;
}

// reads [r0, r1) of the staged batch: the offsets are absolute, so a range is a pointer shift; `slot`: its work counter
template <int R>
static cudaError_t launch_conk_R(c3_handle *h, int grid, int penalty, int32_t *brow, int64_t brow_stride, int r0, int r1, int slot)
{
    c3_conk_kernel<R><<<grid, C3_CONK_THREADS, 0, h->stream>>>(
        h->d_codes.as<uint8_t>(), h->d_off.as<int64_t>() + r0, r1 - r0, h->d_sp_codes.as<uint8_t>(),
        h->d_sp_off.as<int32_t>(), h->d_sp_idx.as<int32_t>() + r0, penalty, h->d_prof.as<int32_t>(), brow, brow_stride,
        h->d_counter.as<unsigned>() + slot);
    return cudaGetLastError();
}

// packed conk (two reads per warp, conk.cuh): pair list of the range on the device, then the kernel; no readback
template <int R>
static cudaError_t launch_conk2_R(c3_handle *h, int grid, int penalty, const int32_t *pairs, const int *n_pairs, int slot)
{
    c3_conk2_kernel<R><<<grid, C3_CONK_THREADS, 0, h->stream>>>(
        h->d_codes.as<uint8_t>(), h->d_off.as<int64_t>(), pairs, n_pairs, h->d_sp_codes.as<uint8_t>(),
        h->d_sp_off.as<int32_t>(), h->d_sp_idx.as<int32_t>(), penalty, h->d_prof.as<int32_t>(), h->d_counter.as<unsigned>() + slot);
    return cudaGetLastError();
}

static int launch_conk2(c3_handle *h, int penalty, int r0, int r1, int slot, int R)
{
    const int nkeys = h->n_splints * 32;
    const int nr = r1 - r0;
    // per slot: hist, start, fill (nkeys each), n_pairs; the pair list of the range sits at r0 + slot * nkeys
    const size_t meta = (size_t)3 * C3_CONK2_MAXKEYS + 4;
    CK(h->d_pairs.ensure(((size_t)h->n_reads + (size_t)C3_PIPE_CHUNKS * C3_CONK2_MAXKEYS + C3_PIPE_CHUNKS * meta + 64) * 4));
    int32_t *base = h->d_pairs.as<int32_t>();
    unsigned *hist = reinterpret_cast<unsigned *>(base) + (size_t)slot * meta, *start = hist + C3_CONK2_MAXKEYS, *fill = start + C3_CONK2_MAXKEYS;
    int *n_pairs = reinterpret_cast<int *>(fill + C3_CONK2_MAXKEYS);
    int32_t *pairs = base + C3_PIPE_CHUNKS * meta + (size_t)r0 + (size_t)slot * C3_CONK2_MAXKEYS;
    pairs = reinterpret_cast<int32_t *>((reinterpret_cast<uintptr_t>(pairs) + 7) & ~(uintptr_t)7);
    CK(cudaMemsetAsync(hist, 0, meta * 4, h->stream));
    CK(cudaMemsetAsync(pairs, 0xff, ((size_t)nr + nkeys + 2) * 4, h->stream));
    c3_conk2_count_kernel<<<(nr + 255) / 256, 256, 0, h->stream>>>(r0, r1, h->d_off.as<int64_t>(), h->d_sp_idx.as<int32_t>(), hist);
    c3_conk2_scan_kernel<<<1, 1024, 0, h->stream>>>(nkeys, hist, start, n_pairs);
    c3_conk2_scatter_kernel<<<(nr + 255) / 256, 256, 0, h->stream>>>(r0, r1, h->d_off.as<int64_t>(), h->d_sp_idx.as<int32_t>(), start, fill, pairs);
    CK(cudaGetLastError());
    h->tim.kernel_launches += 3;
    const int warps_per_block = C3_CONK_THREADS / 32;
    const int grid = std::max(1, std::min(h->sm_count * 4, ((nr + 1) / 2 + nkeys + warps_per_block - 1) / warps_per_block));
    cudaError_t e;
    switch (R) {
#define C3_CASE(r) case r: e = launch_conk2_R<r>(h, grid, penalty, pairs, n_pairs, slot); break;
        C3_CASE(1) C3_CASE(2) C3_CASE(3) C3_CASE(4) C3_CASE(5) C3_CASE(6) C3_CASE(7) C3_CASE(8)
        C3_CASE(9) C3_CASE(10) C3_CASE(11) C3_CASE(12) C3_CASE(13) C3_CASE(14) C3_CASE(15)
#undef C3_CASE
        default: e = cudaErrorInvalidValue;
    }
    CK(e);
    h->tim.kernel_launches++;
    return 0;
}

static int launch_conk(c3_handle *h, int penalty, int r0 = 0, int r1 = -1, int slot = 0)
{
    if (r1 < 0) r1 = h->n_reads;
    if (r1 <= r0) return 0;
    CK(h->d_prof.ensure((size_t)h->total_bases * 4 + 16));
    CK(h->d_counter.ensure(64));
    if (slot == 0) CK(cudaMemsetAsync(h->d_counter.p, 0, 64, h->stream));
    int R = (h->max_ls + 31) / 32;
    if (R < 1) R = 1;
    if (R > C3_CONK_MAXR) R = C3_CONK_MAXR;
    // two reads per warp when everything fits 16-bit halves: one pass over the splint (<= 480 rows), a sane penalty,
    // a pairing key space that the scan kernel covers, and enough reads to fill the grid with pairs (C3POA_CONK_PACKED_MIN:
    // tests lower the bound to run small cases through the packed kernel; C3POA_CONK_INT32 forces the one-read kernel)
    if (R <= 15 && h->max_ls <= 32 * R && penalty >= 0 && penalty <= 8000 && h->n_splints * 32 <= C3_CONK2_MAXKEYS &&
        r1 - r0 >= (getenv("C3POA_CONK_PACKED_MIN") ? atoi(getenv("C3POA_CONK_PACKED_MIN")) : 4096) && !getenv("C3POA_CONK_INT32"))
        return launch_conk2(h, penalty, r0, r1, slot, R);
    const int warps_per_block = C3_CONK_THREADS / 32;
    int grid = h->sm_count * 4;
    grid = std::max(1, std::min(grid, (r1 - r0 + warps_per_block - 1) / warps_per_block));
    int32_t *brow = nullptr; int64_t bstride = 0;
    if (h->max_ls > 32 * R) {
        bstride = ((int64_t)h->max_lr + 31) & ~31ll;
        CK(h->d_brow.ensure((size_t)grid * warps_per_block * 2 * bstride * 4));
        brow = h->d_brow.as<int32_t>();
    }
    cudaError_t e;
    switch (R) {
#define C3_CASE(r) case r: e = launch_conk_R<r>(h, grid, penalty, brow, bstride, r0, r1, slot); break;
        C3_CASE(1) C3_CASE(2) C3_CASE(3) C3_CASE(4) C3_CASE(5) C3_CASE(6) C3_CASE(7) C3_CASE(8)
        C3_CASE(9) C3_CASE(10) C3_CASE(11) C3_CASE(12) C3_CASE(13) C3_CASE(14) C3_CASE(15) C3_CASE(16)
#undef C3_CASE
        default: e = cudaErrorInvalidValue;
    }
    CK(e);
    h->tim.kernel_launches++;
    return 0;
}

static int launch_peaks(c3_handle *h, const int32_t *d_prof, const int64_t *d_off, int n, int max_len,
                        const double *coef, int window, int iters, int min_dist, double hm, double gm,
                        bool want_smoothed, bool want_median, int max_peaks, int64_t total)
{
    if (window < 1 || window > C3_PK_MAXWIN || !(window & 1)) return fail(h, -5, "window must be odd and <= %d", C3_PK_MAXWIN);
    // (Smoothing a read in place in shared memory -- the whole read resident, no scratch round trip -- was measured: 33.9
    // vs 29.2 ms per 100k reads.  60 KB per CTA leave two CTAs per SM, and the kernel's time is barriers and latency in
    // the select / maxima phases, not the FIR: the scratch stays, it is L2-resident.)
    int bps = 4;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, c3_peaks_kernel, C3_PK_THREADS, 0) != cudaSuccess || bps < 1) bps = 4;
    int grid = std::max(1, std::min(h->sm_count * bps, n));
    int64_t stride = ((int64_t)max_len + 63) & ~63ll;
    CK(h->d_pk_scratch.ensure((size_t)grid * 2 * stride * 8));
    CK(h->d_peaks.ensure((size_t)n * max_peaks * 4 + 16));
    CK(h->d_npk.ensure((size_t)n * 4 + 16));
    if (want_smoothed) CK(h->d_smoothed.ensure((size_t)total * 8 + 16));
    if (want_median) CK(h->d_median.ensure((size_t)n * 8 + 16));
    CK(h->d_counter.ensure(64));
    CK(cudaMemsetAsync(h->d_counter.p, 0, 64, h->stream));
    CK(cudaMemsetAsync(h->d_peaks.p, 0, (size_t)n * max_peaks * 4, h->stream));
    c3_peaks_args A;
    A.prof = d_prof; A.off = d_off; A.n = n; A.window = window; A.iters = iters;
    for (int k = 0; k < C3_PK_MAXWIN; ++k) A.coefv[k] = k < window ? coef[k] : 0.0;
    A.min_dist = min_dist; A.height_mult = hm; A.gate_mult = gm; A.scratch = h->d_pk_scratch.as<double>();
    A.scratch_stride = stride; A.out_smoothed = want_smoothed ? h->d_smoothed.as<double>() : nullptr;
    A.out_median = want_median ? h->d_median.as<double>() : nullptr; A.out_peaks = h->d_peaks.as<int32_t>();
    A.out_n_peaks = h->d_npk.as<int32_t>(); A.max_peaks = max_peaks; A.counter = h->d_counter.as<unsigned>();
    c3_peaks_kernel<<<grid, C3_PK_THREADS, 0, h->stream>>>(A);
    CK(cudaGetLastError());
    h->tim.kernel_launches++;
    return 0;
}

static void to_dev_para(const c3_handle *h, const c3_poa_params *p, c3_poa_para_dev *d)
{
    d->match = p->match; d->mismatch = p->mismatch; d->o1 = p->gap_open1; d->e1 = p->gap_ext1;
    d->o2 = p->gap_open2; d->e2 = p->gap_ext2; d->wb = p->wb; d->simd_bits = p->simd_bits; d->wf = p->wf;
    d->int8_lanes = h->sw_int8_lanes; d->end_clamp = h->sw_end_clamp;
}

// auto mode: what goes to the group kernel (poa_grp.cuh + poa_graph.cuh); measured on B200, see profiles/README.md
#define C3_GRP_AUTO_MAX_LEN 2600       // mean subread length of a read
#define C3_GRP_AUTO_MAX_NSEQ 32
#define C3_GRP_AUTO_MIN_READS 12000
// ---------------------------------------------------------------------------
// POA work order.  Shared by the host builder (B3: the inputs are host arrays) and the device builder (B4: the per-read
// repeat counts and subread bases are on the device, nothing is copied back but 64 bytes of totals).
// ---------------------------------------------------------------------------
#define C3_ORD_NB (64 * 16)            // 1/16-octave cost classes
struct c3_order_rule {                 // what decides list membership, by value into the kernels
    double wb, wf;
    int min_seqs, msa2, mode;
    int lane_ok, grp_ok;               // parameter sets the fast kernels cover
    long long lim16;                   // int16 score bound
    int e1, o1;
};
// cost class of an item: cost ~ alignments x mean length x band width
__host__ __device__ inline int c3_order_class(const int nseq, const long long total, const c3_order_rule &R)
{
    const double L = (double)total / nseq;
    const double cost = (double)(nseq - 1 > 1 ? nseq - 1 : 1) * L * (2.0 * (R.wb + R.wf * L) + 48.0) + 1.0;
    int e; const double m = frexp(cost, &e);                      // cost = m * 2^e, m in [0.5, 1)
    int b = e * 16 + (int)((m - 0.5) * 32.0);
    return b < 0 ? 0 : (b > C3_ORD_NB - 1 ? C3_ORD_NB - 1 : b);
}
// bit 0: lane kernel's list, bit 1: group path's list
__host__ __device__ inline int c3_order_flags(const int nseq, const long long total, const c3_order_rule &R, int *grp_q)
{
    int f = 0;
    const bool rows = R.msa2 && nseq == 2;                        // MSA rows wanted: the group path and the warp kernel emit them
    const long long L = total / nseq;
    if (R.lane_ok && !rows && L <= 5000 && nseq <= 40) f |= 1;    // long / deep reads: per-thread arenas would not fit
    if (R.grp_ok) {
        // int16 score mode of an alignment: max(qlen * 5, max(qlen, nodes) * e1 + o1) <= lim, with the usual graph
        // growth (the kernel re-checks per alignment and declines what turns out larger)
        const long long Lm = L * 5 / 4 + 1;
        const long long nodes = 2 + Lm + (long long)(nseq - 1) * (Lm * 35 / 100 + 16);
        const bool fits = !(Lm * 5 > R.lim16 || (Lm > nodes ? Lm : nodes) * R.e1 + R.o1 > R.lim16 || nodes > 65000);
        // auto: the bulk of short subreads only.  The serial phases of the group path run one thread per read with a
        // latency floor per launch that grows with the sequence length, and a read's workspace grows with length x
        // depth: long or very deep reads are better off in the warp kernel (profiles/README.md)
        const bool bulk = R.mode != 0 || (L <= C3_GRP_AUTO_MAX_LEN && nseq <= C3_GRP_AUTO_MAX_NSEQ);
        if (fits && bulk) { f |= 2; *grp_q = R.mode == 0 ? (int)(Lm < 65000 ? Lm : 65000) : 65000; }
    }
    return f;
}
static c3_order_rule make_order_rule(const c3_handle *h, const c3_poa_args &A, int min_seqs, const c3_poa_params *pp)
{
    c3_order_rule R;
    R.wb = pp->wb; R.wf = pp->wf; R.min_seqs = min_seqs; R.msa2 = A.msa2; R.mode = h->poa_mode;
    R.lane_ok = pp->simd_bits == 256 && pp->wb >= 0;
    R.grp_ok = pp->simd_bits == 256 && pp->wb >= 0 && pp->gap_open1 + pp->gap_ext1 <= 7 &&
               pp->gap_open2 + pp->gap_ext2 <= 31 && pp->gap_ext1 >= 0 && pp->gap_ext2 >= 0 &&
               pp->gap_ext1 <= 16 && pp->gap_ext2 <= 16 && pp->match >= 0 && pp->mismatch >= 0 &&
               pp->match <= 100 && pp->mismatch <= 100 && pp->gap_open1 >= 0 && pp->gap_open2 >= 0;
    R.lim16 = 32767 - pp->mismatch - pp->gap_open1 - pp->gap_ext1; R.e1 = pp->gap_ext1; R.o1 = pp->gap_open1;
    return R;
}

// device builder: histogram by class for the three lists, descending scan, scatter.  Inside a class the order is the
// order of arrival (it only decides which reads share a warp; every read's result is its own).
struct c3_order_tot {                  // 64 bytes copied back
    int n_work, n_lane, n_grp, grp_max_nseq, grp_max_q, lane_max_nseq, lane_cls_first, lane_cls_median;
    long long grp_max_total, lane_max_total;
    int pad[4];
};
__global__ void c3_order_classify_kernel(int n, const c3_read_result_dev *res, c3_order_rule R, int32_t *cls, unsigned *hist,
                                         c3_order_tot *tot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ns = res[i].status >= 0 ? res[i].n_sub : 0;
    const long long total = ns >= 2 ? res[i].poa_cells : 0;        // the split kernel leaves the subread bases there
    if (ns < R.min_seqs || ns < 1) { cls[i] = -1; return; }
    const int b = c3_order_class(ns, total, R);
    int gq = 0;
    const int f = c3_order_flags(ns, total, R, &gq);
    cls[i] = b | (f << 16);
    atomicAdd(&hist[b], 1u);
    if (f & 1) {
        atomicAdd(&hist[C3_ORD_NB + b], 1u);
        atomicMax((unsigned long long *)&tot->lane_max_total, (unsigned long long)total); atomicMax(&tot->lane_max_nseq, ns);
    }
    if (f & 2) {
        atomicAdd(&hist[2 * C3_ORD_NB + b], 1u);
        atomicMax((unsigned long long *)&tot->grp_max_total, (unsigned long long)total); atomicMax(&tot->grp_max_nseq, ns);
        atomicMax(&tot->grp_max_q, gq);
    }
}
__global__ void __launch_bounds__(32) c3_order_scan_kernel(const unsigned *hist, unsigned *start, c3_order_tot *tot)
{
    const int l = threadIdx.x;                                      // lanes 0..2: one list each, classes from the top
    if (l >= 3) return;
    unsigned acc = 0;
    for (int b = C3_ORD_NB - 1; b >= 0; --b) { start[l * C3_ORD_NB + b] = acc; acc += hist[l * C3_ORD_NB + b]; }
    if (l == 0) tot->n_work = (int)acc;
    if (l == 2) tot->n_grp = (int)acc;
    if (l == 1) {
        tot->n_lane = (int)acc;
        int first = -1, med = -1; unsigned seen = 0;
        for (int b = C3_ORD_NB - 1; b >= 0 && med < 0; --b) {
            const unsigned c = hist[C3_ORD_NB + b];
            if (c && first < 0) first = b;
            seen += c;
            if (c && seen > acc / 2) med = b;
        }
        tot->lane_cls_first = first; tot->lane_cls_median = med;
    }
}
__global__ void c3_order_scatter_kernel(int n, const int32_t *cls, const unsigned *start, unsigned *fill, int32_t *order,
                                        int32_t *lane, int32_t *grp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cls[i];
    if (c < 0) return;
    const int b = c & 0xffff, f = c >> 16;
    order[start[b] + atomicAdd(&fill[b], 1u)] = i;
    if (f & 1) lane[start[C3_ORD_NB + b] + atomicAdd(&fill[C3_ORD_NB + b], 1u)] = i;
    if (f & 2) grp[start[2 * C3_ORD_NB + b] + atomicAdd(&fill[2 * C3_ORD_NB + b], 1u)] = i;
}

static int build_poa_order_device(c3_handle *h, c3_poa_args &A, int n, const c3_read_result_dev *res, int min_seqs,
                                  const c3_poa_params *pp)
{
    const c3_order_rule R = make_order_rule(h, A, min_seqs, pp);
    CK(h->d_order.ensure((size_t)std::max(n, 1) * 4));
    CK(h->d_order_lane.ensure((size_t)std::max(n, 1) * 4));
    CK(h->d_order_grp.ensure((size_t)std::max(n, 1) * 4));
    const size_t hb = (size_t)3 * C3_ORD_NB * 4;
    CK(h->d_order_scratch.ensure((size_t)n * 4 + 3 * hb + sizeof(c3_order_tot) + 256));
    uint8_t *sc = h->d_order_scratch.as<uint8_t>();
    unsigned *hist = reinterpret_cast<unsigned *>(sc), *start = hist + 3 * C3_ORD_NB, *fill = start + 3 * C3_ORD_NB;
    c3_order_tot *tot = reinterpret_cast<c3_order_tot *>(sc + 3 * hb);
    int32_t *cls = reinterpret_cast<int32_t *>(sc + 3 * hb + 64 + 64);
    CK(cudaMemsetAsync(sc, 0, 3 * hb + sizeof(c3_order_tot), h->stream));
    c3_order_classify_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(n, res, R, cls, hist, tot);
    c3_order_scan_kernel<<<1, 32, 0, h->stream>>>(hist, start, tot);
    c3_order_scatter_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(n, cls, start, fill, h->d_order.as<int32_t>(),
                                                                    h->d_order_lane.as<int32_t>(), h->d_order_grp.as<int32_t>());
    CK(cudaGetLastError());
    h->tim.kernel_launches += 3;
    c3_order_tot T;
    const auto dbg0 = std::chrono::steady_clock::now();
    CK(cudaMemcpyAsync(&T, tot, sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    CK(h->sync());
    h->dbg_sync_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - dbg0).count();
    h->dbg_t1 = std::chrono::steady_clock::now();                          // 64 bytes: list sizes and maxima size the workspaces
    A.order = h->d_order.as<int32_t>(); A.n_work = T.n_work;
    h->n_work_lane = T.n_lane; h->lane_max_total = T.lane_max_total; h->lane_max_nseq = T.lane_max_nseq; h->lane_max_q = 0;
    h->lane_cost_spread = T.n_lane > 0 ? exp2((double)(T.lane_cls_first - T.lane_cls_median) / 16.0) : 1.0;
    h->n_work_grp = T.n_grp; h->grp_max_total = T.grp_max_total; h->grp_max_nseq = T.grp_max_nseq; h->grp_max_q = T.grp_max_q;
    h->grp_nseq.clear();                                           // waves of the group path use the list's maximum
    if (h->poa_mode == 0 && h->n_work_grp < C3_GRP_AUTO_MIN_READS) h->n_work_grp = 0;     // too few to amortise the launches
    return 0;
}

// Work order for the persistent POA grid: items with >= min_seqs sequences, largest estimated DP cost first
// (LPT scheduling: the longest reads start first, so a batch ends with short ones).  cost ~ alignments x
// mean length x band width.  O(n) bucket sort on 1/16-octave cost classes.
static int upload_poa_order(c3_handle *h, c3_poa_args &A, const std::vector<int32_t> &nseq, const std::vector<int64_t> &total,
                            int min_seqs, const c3_poa_params *pp)
{
    const c3_order_rule R = make_order_rule(h, A, min_seqs, pp);
    const int n = (int)nseq.size();
    std::vector<int32_t> cls((size_t)n, -1);
    std::vector<int32_t> cnt(C3_ORD_NB + 1, 0);
    int n_work = 0;
    for (int i = 0; i < n; ++i) {
        if (nseq[i] < min_seqs || nseq[i] < 1) continue;
        cls[i] = c3_order_class(nseq[i], total[i], R); ++cnt[cls[i]]; ++n_work;
    }
    std::vector<int32_t> start(C3_ORD_NB + 1, 0);
    for (int b = C3_ORD_NB - 1, acc = 0; b >= 0; --b) { start[b] = acc; acc += cnt[b]; }   // descending classes
    std::vector<int32_t> order((size_t)std::max(n_work, 1));
    for (int i = 0; i < n; ++i) if (cls[i] >= 0) order[start[cls[i]]++] = i;
    // the fast kernels' shares: same order (neighbours become the threads / groups of a warp, so they are of similar
    // size); their workspaces are sized from these items alone
    std::vector<int32_t> lane((size_t)std::max(n_work, 1)), grp((size_t)std::max(n_work, 1));
    int nl = 0, ng = 0;
    h->lane_max_total = 0; h->lane_max_nseq = 0; h->lane_max_q = 0;
    h->grp_nseq.clear(); h->grp_max_total = 0; h->grp_max_nseq = 0; h->grp_max_q = 0;
    for (int k = 0; k < n_work; ++k) {
        const int i = order[k];
        int gq = 0;
        const int f = c3_order_flags(nseq[i], total[i], R, &gq);
        if (f & 1) {
            lane[nl++] = i;
            h->lane_max_total = std::max(h->lane_max_total, total[i]); h->lane_max_nseq = std::max(h->lane_max_nseq, nseq[i]);
        }
        if (f & 2) {
            grp[ng++] = i;
            h->grp_nseq.push_back(nseq[i]);
            h->grp_max_total = std::max(h->grp_max_total, total[i]); h->grp_max_nseq = std::max(h->grp_max_nseq, nseq[i]);
            h->grp_max_q = std::max(h->grp_max_q, gq);
        }
    }
    if (h->poa_mode == 0 && ng < C3_GRP_AUTO_MIN_READS) { ng = 0; h->grp_nseq.clear(); }    // too few to amortise the launches
    h->lane_cost_spread = nl > 0 ? exp2((double)(cls[lane[0]] - cls[lane[nl / 2]]) / 16.0) : 1.0;
    CK(h->d_order.ensure((size_t)std::max(n_work, 1) * 4));
    CK(h->d_order_grp.ensure((size_t)std::max(ng, 1) * 4));
    CK(h->d_order_lane.ensure((size_t)std::max(nl, 1) * 4));
    CK(cudaMemcpyAsync(h->d_order.p, order.data(), (size_t)n_work * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_order_grp.p, grp.data(), (size_t)ng * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_order_lane.p, lane.data(), (size_t)nl * 4, cudaMemcpyHostToDevice, h->stream));
    CK(h->sync());                          // the lists are locals
    A.order = h->d_order.as<int32_t>(); A.n_work = n_work;
    h->n_work_lane = nl; h->n_work_grp = ng;
    return 0;
}

// Lane kernel pass (poa_lane.cuh) ahead of the warp kernel: covers the bulk of a large batch; whatever it
// declines or cannot fit stays not-done and is picked up by c3_poa_kernel afterwards.
static int launch_poa_lane(c3_handle *h, c3_poa_args &A, int max_q, const c3_poa_params *pp)
{
    h->lane_items = 0;
    A.done = nullptr;
    const int nl = h->n_work_lane;
    if (h->poa_mode == 1 || nl <= 0 || h->sw_int8_lanes || h->sw_end_clamp) return 0;
    const int max_nseq = h->lane_max_nseq;
    const int64_t max_total = h->lane_max_total;
    max_q = (int)std::min<int64_t>(max_q, max_total);
    // per-thread graph workspace: same node estimate as the warp kernel
    int64_t est = 2 + (int64_t)max_q + (int64_t)(max_nseq - 1) * ((int64_t)max_q * 35 / 100 + 16);
    int64_t node_cap = std::min<int64_t>(std::min<int64_t>(est, max_total + 2), 65504);
    node_cap = std::max<int64_t>((node_cap + 31) & ~31ll, 64);
    const int pool_cap = (int)node_cap;
    const int cigar_cap = (int)((max_q + node_cap + 64 + 1) & ~1ll);
    const int qp_stride = (max_q + 48) & ~15;
    const int64_t ws_bytes = c3_poa_ws_bytes((int)node_cap, pool_cap, 0, cigar_cap, qp_stride);
    // per-warp DP arena: rows of one alignment x vectors per row step (the widest band among 32 threads).
    // Sized for the usual case, not the worst: an overflow only sends those reads to the warp kernel.
    const int w = pp->wb + (int)(pp->wf * max_q);
    const int64_t rows_est = 2 + (int64_t)max_q + (int64_t)(max_nseq - 1) * ((int64_t)max_q * 20 / 100 + 16);
    const int64_t vec_est = (2 * w + 48) / 16 + 1;
    int64_t arena4 = std::min<int64_t>(rows_est, node_cap) * vec_est * C3L_VSTRIDE;
    if (arena4 > 0x7fffff00ll / 4) return 0;                       // int32 cell indices
    const int64_t warp_bytes = 32 * ws_bytes + arena4 * 16;
    const int wpb = C3L_THREADS / 32;
    // shared-memory ring of the row just computed: vec_est slots of 3.5 KB per warp, cut down if C3L_MINB CTAs
    // would not fit the SM (rows wider than the ring are simply not mirrored)
    int sm_vec = (int)vec_est;
    {
        const int per_cta_max = (C3L_SMEM_KB * 1024) / C3L_MINB;
        sm_vec = std::max(1, std::min(sm_vec, per_cta_max / (wpb * C3L_RSLOT * 512)));
    }
    const size_t sm_bytes = (size_t)wpb * sm_vec * C3L_RSLOT * 512;
    CK(cudaFuncSetAttribute(c3_poa_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_bytes));
    int bps = 4;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, c3_poa_lane_kernel, C3L_THREADS, sm_bytes) != cudaSuccess || bps < 1) bps = 4;
    bps = std::min(bps, C3L_MINB);
    int64_t warps = (int64_t)h->sm_count * bps * wpb;
    if (const char *lim = getenv("C3POA_LANE_WARPS_PER_SM")) {       // tuning only: occupancy sweeps
        const int w = atoi(lim);
        if (w >= wpb) warps = std::min<int64_t>(warps, (int64_t)h->sm_count * (w / wpb) * wpb);
    }
    // auto mode: 32 reads advance in lockstep per warp, so one warp item takes as long as ~20 reads in the warp kernel;
    // that pays only when the batch fills most of the grid (measured break-even: about half a wave) and the reads are
    // of similar size (the largest items set the latency of the whole launch).  Otherwise the warp kernel is faster.
    if (h->poa_mode == 0 && ((int64_t)nl * 4 < warps * 32 * 3 || h->lane_cost_spread > 4.0)) return 0;
    warps = std::min<int64_t>(warps, ((int64_t)nl + 31) / 32);
    size_t free_b = 0, tot_b = 0;
    CK(cudaMemGetInfo(&free_b, &tot_b));
    const int64_t budget = (int64_t)((double)(free_b + h->d_ws.cap) * 0.8);
    const int64_t fit = budget / warp_bytes / wpb * wpb;
    if (h->poa_mode == 0 && fit * 2 < warps) return 0;             // auto: not worth it with under half of the grid resident
    warps = std::min<int64_t>(warps, fit);
    if (warps < wpb) return 0;                                     // does not fit: the warp kernel takes everything
    CK(h->d_ws.ensure((size_t)(warps * warp_bytes)));
    CK(h->d_done.ensure((size_t)A.n_items * 4));
    CK(cudaMemsetAsync(h->d_done.p, 0, (size_t)A.n_items * 4, h->stream));
    CK(h->d_counter.ensure(64));
    CK(cudaMemsetAsync(h->d_counter.p, 0, 64, h->stream));
    c3_lane_args L;
    L.A = A;
    L.A.ws = h->d_ws.as<uint8_t>(); L.A.ws_stride = ws_bytes;
    L.A.node_cap = (int)node_cap; L.A.pool_cap = pool_cap; L.A.cell_cap = 0; L.A.cigar_cap = cigar_cap; L.A.qp_stride = qp_stride;
    L.A.counter = h->d_counter.as<unsigned>();
    L.A.order = h->d_order_lane.as<int32_t>(); L.A.n_work = nl;
    L.arena = reinterpret_cast<int4 *>(h->d_ws.as<uint8_t>() + warps * 32 * ws_bytes);
    L.arena_stride4 = arena4; L.arena_cap4 = (int)arena4; L.sm_vec = sm_vec;
    L.done = h->d_done.as<int32_t>();
    h->kt_begin();
    c3_poa_lane_kernel<<<(int)(warps / wpb), C3L_THREADS, sm_bytes, h->stream>>>(L);
    h->kt_end(3);
    CK(cudaGetLastError());
    h->tim.kernel_launches++;
    h->lane_items = nl; h->lane_n_items = A.n_items;
    A.done = h->d_done.as<int32_t>();
    return 0;
}

// Group kernels (poa_grp.cuh) ahead of the warp kernel: every read the host knows to be in their scope, in waves of as
// many reads as fit the device memory (state + graph workspace + DP arena per read, all resident: a B200 holds a whole
// 100k-read batch).  Per wave: graph kernel, then (DP kernel, graph kernel) per further sequence.  Whatever is declined
// (capacity, exactness guard) stays not-done and is picked up by c3_poa_kernel afterwards.
static int launch_poa_grp(c3_handle *h, c3_poa_args &A, int max_q, const c3_poa_params *pp)
{
    h->lane_items = 0;
    A.done = nullptr;
    const int ng = h->n_work_grp;
    if ((h->poa_mode != 0 && h->poa_mode != 3) || ng <= 0 || h->sw_int8_lanes || h->sw_end_clamp) return 0;
    const int max_nseq = h->grp_max_nseq;
    const int64_t max_total = h->grp_max_total;
    max_q = (int)std::min<int64_t>(std::min<int64_t>(max_q, max_total), std::max(h->grp_max_q, 64));
    // Per-read capacity: nodes = first sequence + a share of every further one.  The serial phases run one thread per
    // read and are bound by the length of their dependent chains, not by the number of reads, so one wave for the whole
    // batch is worth a tighter node estimate (what outgrows it is declined and goes to the warp kernel with its own,
    // roomier workspace): the largest growth share of {35, 25, 20} % that lets the batch fit in one wave, else 20 %.
    int budget_pct = 70, grow_env = 0;
    if (const char *e = getenv("C3POA_GRP_GROW_PCT")) grow_env = std::max(1, atoi(e));           // tuning only
    if (const char *e = getenv("C3POA_GRP_BUDGET_PCT")) budget_pct = std::max(1, std::min(95, atoi(e)));
    // cudaMemGetInfo is a driver round trip that takes anything from 0.1 to 100 ms once >100 GB are allocated (measured: it
    // was the whole gap between the mid-pipeline readback and the first POA kernel, 5-60 ms in one step out of three, and
    // the run-to-run spread of round 1's numbers): asked only when the workspace this handle owns does not do
    int64_t budget = (int64_t)h->d_ws_grp.cap;
    bool asked = false;
    auto ask = [&]() -> int {
        if (asked) return 0;
        size_t free_b = 0, tot_b = 0;
        CK(cudaMemGetInfo(&free_b, &tot_b));
        budget = (int64_t)((double)(free_b + h->d_ws_grp.cap) * (budget_pct / 100.0));
        asked = true;
        return 0;
    };
    const int w = pp->wb + (int)(pp->wf * max_q);
    const int need = (2 * w + 1 + 48) / 16 + 2;
    int vs_shift = 3;
    while ((1 << vs_shift) < need && vs_shift < 8) ++vs_shift;
    // lanes per read: short sequences (band half-width <= 20: bands of 3-4 vectors) run 8 reads per warp with 4 lanes each
    // and a 4-vector ring; everything else 4 reads per warp with 8 lanes each
    int gl = (vs_shift == 3 && w <= 20) ? 4 : 8;
    if (const char *e = getenv("C3POA_GRP_GL")) gl = (atoi(e) == 4 && vs_shift == 3) ? 4 : 8;          // tests / tuning
    const int rv_shift = gl == 4 ? 2 : (vs_shift == 3 ? 3 : 4);
    const int qp_stride = (max_q + 48) & ~15;
    int64_t node_cap = 0, ws_bytes = 0, arena4 = 0, read_bytes = 0;
    int cigar_cap = 0;
    static const int grow_try[3] = {35, 25, 20};
    auto size_for = [&](const int grow_pct) {
        const int64_t est = 2 + (int64_t)max_q + (int64_t)(max_nseq - 1) * ((int64_t)max_q * grow_pct / 100 + 16);
        node_cap = std::min<int64_t>(std::min<int64_t>(est, max_total + 2), 65504);
        node_cap = std::max<int64_t>((node_cap + 31) & ~31ll, 64);
        cigar_cap = (int)((max_q + node_cap + 64 + 1) & ~1ll);
        ws_bytes = c3g_ws_bytes((int)node_cap, (int)node_cap, cigar_cap, qp_stride);
        arena4 = (node_cap << vs_shift) * 3;
        read_bytes = ws_bytes + arena4 * 16 + (int64_t)sizeof(c3g_state);
    };
    // steady state (batch after batch of the same shape): the growth share chosen last time, inside the workspace the
    // handle already owns -- no driver query.  Otherwise ask once and choose again.
    size_for(grow_env ? grow_env : grow_try[h->grp_grow_idx]);
    // (a batch that needs several waves does not ask on every call, unless what the handle owns holds no decent wave)
    const bool may_ask = h->d_ws_grp.cap == 0 || h->grp_ask_wait <= 0 || budget / read_bytes < std::min(ng, 12000);
    if (h->grp_ask_wait > 0) --h->grp_ask_wait;
    if (!grow_env && budget / read_bytes < ng && may_ask) {
        h->grp_ask_wait = 64;
        if (int rc = ask()) return rc;
        for (int t = 0; t < 3; ++t) {
            size_for(grow_try[t]);
            h->grp_grow_idx = t;
            if (budget / read_bytes >= ng) break;
        }
    } else if (grow_env && budget / read_bytes < ng && may_ask) { h->grp_ask_wait = 64; if (int rc = ask()) return rc; }
    const int pool_cap = (int)node_cap;
    const int wpb = C3G_THREADS / 32;
    const size_t sm_dp = (size_t)wpb * (32 / gl) * c3g_smem_group_bytes(rv_shift);
    void (*kdp)(c3g_args) = gl == 4 ? c3_poa_grp_dp_kernel<2, true, 4> : vs_shift == 3 ? c3_poa_grp_dp_kernel<3, false> : c3_poa_grp_dp_kernel<4, true>;
    // function attributes and occupancy once per handle and instantiation: driver calls on the path between the batch's
    // only mid-pipeline synchronisation and the first POA launch were seen to stall for milliseconds now and then
    const int kv = gl == 4 ? 2 : vs_shift == 3 ? 0 : 1;
    if (!h->grp_bps[kv]) {
        CK(cudaFuncSetAttribute(kdp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_dp));
        int b = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kdp, C3G_THREADS, sm_dp) != cudaSuccess || b < 1) b = 1;
        h->grp_bps[kv] = b;
    }
    int bps_dp = h->grp_bps[kv];
    if (const char *lim = getenv("C3POA_GRP_DP_CTAS")) bps_dp = std::max(1, std::min(bps_dp, atoi(lim)));     // tuning only
    int64_t wave = std::min<int64_t>(ng, budget / read_bytes);
    if (wave < std::min<int64_t>(ng, 64)) return 0;                // does not fit: the warp kernel takes everything
    wave = (ng + (ng + wave - 1) / wave - 1) / ((ng + wave - 1) / wave);    // several waves: equal sizes
    const int n_counters = 2 * max_nseq + 3;
    // several handles of one device (the driver keeps batches in flight) size themselves from the same free figure: an
    // allocation that fails is not an error of the batch -- the warp kernel, with its small workspace, takes the reads
    if (h->d_ws_grp.ensure((size_t)(wave * read_bytes) + (size_t)n_counters * 4 + 256) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    CK(h->d_done.ensure((size_t)A.n_items * 4));
    CK(cudaMemsetAsync(h->d_done.p, 0, (size_t)A.n_items * 4, h->stream));
    uint8_t *base = h->d_ws_grp.as<uint8_t>();
    c3g_args L;
    memset(&L, 0, sizeof(L));
    L.A = A;
    L.A.node_cap = (int)node_cap; L.A.pool_cap = pool_cap; L.A.cell_cap = 0; L.A.cigar_cap = cigar_cap; L.A.qp_stride = qp_stride;
    L.ws = base; L.ws_stride = ws_bytes;
    L.arena = reinterpret_cast<uint4 *>(base + wave * ws_bytes);
    L.arena_stride4 = arena4; L.vs_shift = vs_shift; L.rv_shift = rv_shift;
    L.state = reinterpret_cast<c3g_state *>(base + wave * (ws_bytes + arena4 * 16));
    unsigned *counters = reinterpret_cast<unsigned *>(base + wave * read_bytes + 128);
    L.done = h->d_done.as<int32_t>();
    const bool timing = getenv("C3POA_GRP_TIMING") != nullptr;
    for (int64_t w0 = 0; w0 < ng; w0 += wave) {
        const int nw = (int)std::min<int64_t>(wave, ng - w0);
        int wave_nseq = 1;
        if (h->grp_nseq.empty()) wave_nseq = std::max(1, max_nseq);      // order built on the device: the list's maximum
        else for (int k = 0; k < nw; ++k) wave_nseq = std::max(wave_nseq, h->grp_nseq[(size_t)(w0 + k)]);
        CK(cudaMemsetAsync(counters, 0, (size_t)n_counters * 4, h->stream));
        L.A.order = h->d_order_grp.as<int32_t>() + w0; L.A.n_work = nw;
        L.eager = nw < 48000;
        const int w_dp = (nw + 32 / gl - 1) / (32 / gl);            // warps that can be busy
        const int grid_dp = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)h->sm_count * bps_dp, (w_dp + wpb - 1) / wpb));
        const int grid_gr = (nw + C3S_THREADS - 1) / C3S_THREADS;  // one thread per read
        int launch = 0;
        if (timing) fprintf(stderr, "host: order sync %.2f ms, sync->first POA launch %.2f ms\n", h->dbg_sync_ms,
                            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h->dbg_t1).count());
        h->kt_begin();
        c3_poa_graph_init_kernel<<<nw, 128, 0, h->stream>>>(L);
        h->kt_end(1);
        h->tim.kernel_launches++;
        for (int sq = 1; sq < wave_nseq; ++sq) {
            L.A.counter = counters + launch++;
            kdp<<<grid_dp, C3G_THREADS, sm_dp, h->stream>>>(L);
            h->kt_end(0);
            c3_poa_graph_kernel<<<grid_gr, C3S_THREADS, 0, h->stream>>>(L);
            h->kt_end(1);
            h->tim.kernel_launches += 2;
        }
        CK(cudaGetLastError());
        if (timing) fprintf(stderr, "grp wave %d reads (node_cap %lld, %lld B/read, %d alignments)\n", nw, (long long)node_cap, (long long)read_bytes, wave_nseq - 1);
    }
    h->lane_items = ng; h->lane_n_items = A.n_items;
    A.done = h->d_done.as<int32_t>();
    return 0;
}

// workspace sizing from the batch maxima (longest sequence, most sequences, largest total)
static int launch_poa(c3_handle *h, c3_poa_args &A, int max_q, int max_nseq, int64_t max_total, const c3_poa_params *pp)
{
    to_dev_para(h, pp, &A.P);
    if (A.P.simd_bits != 128 && A.P.simd_bits != 256 && A.P.simd_bits != 512) return fail(h, -5, "simd_bits must be 128/256/512");
    if (!A.order) { h->n_work_lane = 0; h->n_work_grp = 0; }
    int fast_items = 0;                                            // reads handed to a fast kernel ahead of this one
    {
        int rc = launch_poa_grp(h, A, max_q, pp);
        if (rc) return rc;
        if (!A.done && h->poa_mode == 2 && (rc = launch_poa_lane(h, A, max_q, pp))) return rc;
        if (A.done) fast_items = h->lane_items;
    }
    int64_t est = 2 + (int64_t)max_q + (int64_t)(max_nseq - 1) * ((int64_t)max_q * 35 / 100 + 16);
    int64_t node_cap = std::min<int64_t>(std::min<int64_t>(est, max_total + 2), 65534);
    node_cap = std::max<int64_t>((node_cap + 31) & ~31ll, 64);
    if (node_cap > 65534) node_cap = 65534;
    int pool_cap = (int)std::min<int64_t>(node_cap, 65534);
    int w = pp->wb < 0 ? max_q : pp->wb + (int)(pp->wf * max_q);
    int64_t width = std::min<int64_t>(2ll * w + 64 + 16, (int64_t)max_q + 4);
    width = (width + 3) & ~3ll;
    int64_t cell_cap = std::min<int64_t>(3 * node_cap * width, 0x7ffffff0ll / 4) & ~3ll;
    int cigar_cap = (int)((max_q + node_cap + 64 + 1) & ~1ll);
    int qp_stride = (max_q + 8) & ~3;
    int64_t ws_bytes = c3_poa_ws_bytes((int)node_cap, pool_cap, (int)cell_cap, cigar_cap, qp_stride);
    // one warp per read.  (A two-reads-per-warp variant, 16 lanes each, was measured in round 1: no faster --
    // the kernel is bound by per-row dependency latency, not by the lanes left idle -- and was dropped.)
    const int threads = C3_POA_THREADS;
    const int rpb = threads / 32;                                     // reads per block
    if (!h->warp_bps) {
        int b = 4;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, c3_poa_kernel, threads, 0) != cudaSuccess || b < 1) b = 4;
        h->warp_bps = b;
    }
    const int bps = h->warp_bps;
    int grid = h->sm_count * bps;
    if (!A.order) A.n_work = A.n_items;
    grid = std::max(1, std::min(grid, (A.n_work + rpb - 1) / rpb));
    // what a fast kernel was given comes back only if it declined it: one CTA per SM is kept for those
    if (fast_items > 0 && h->poa_mode != 2)
        grid = std::max(1, std::min(grid, (A.n_work - fast_items + rpb - 1) / rpb + h->sm_count));
    int64_t budget = (int64_t)h->d_ws.cap;
    if ((int64_t)grid * rpb * ws_bytes > budget) {                 // (see launch_poa_grp: asked only when the buffer must grow)
        size_t free_b = 0, tot_b = 0;
        CK(cudaMemGetInfo(&free_b, &tot_b));
        budget = (int64_t)((double)(free_b + h->d_ws.cap) * 0.8);
    }
    int64_t max_slots = budget / ws_bytes;
    if (max_slots < rpb) return fail(h, -6, "POA workspace of %lld bytes per read does not fit", (long long)ws_bytes);
    if ((int64_t)grid * rpb > max_slots) grid = (int)std::max<int64_t>(1, max_slots / rpb);
    CK(h->d_ws.ensure((size_t)grid * rpb * ws_bytes));
    CK(h->d_counter.ensure(64));
    CK(cudaMemsetAsync(h->d_counter.p, 0, 64, h->stream));
    A.ws = h->d_ws.as<uint8_t>(); A.ws_stride = ws_bytes;
    A.node_cap = (int)node_cap; A.pool_cap = pool_cap; A.cell_cap = (int)cell_cap; A.cigar_cap = cigar_cap; A.qp_stride = qp_stride;
    A.counter = h->d_counter.as<unsigned>();
    h->kt_begin();
    c3_poa_kernel<<<grid, threads, 0, h->stream>>>(A);
    h->kt_end(2);
    CK(cudaGetLastError());
    h->tim.kernel_launches++;
    return 0;
}

// ---------------------------------------------------------------------------
// staging of reads + splints (shared by B1 and B4)
// ---------------------------------------------------------------------------
static int stage_reads(c3_handle *h, int32_t n_reads, const char *reads, const int64_t *read_off,
                       int32_t n_splints, const char *splints, const int32_t *splint_off, const int32_t *splint_idx,
                       bool defer_reads = false)
{
    if (!h) return -1;
    if (n_reads <= 0 || !reads || !read_off || n_splints <= 0 || !splints || !splint_off || !splint_idx)
        return fail(h, -5, "bad arguments");
    CK(cudaSetDevice(h->device));
    int64_t total = read_off[n_reads] - read_off[0];
    if (read_off[0] != 0) return fail(h, -5, "read_off[0] must be 0");
    int max_lr = 0;
    for (int i = 0; i < n_reads; ++i) {
        int64_t L = read_off[i + 1] - read_off[i];
        if (L <= 0 || L > 0x3fffffff) return fail(h, -5, "read %d has invalid length %lld", i, (long long)L);
        max_lr = std::max(max_lr, (int)L);
        if (splint_idx[i] < 0 || splint_idx[i] >= n_splints) return fail(h, -5, "splint_idx[%d] out of range", i);
    }
    int max_ls = 0;
    for (int i = 0; i < n_splints; ++i) {
        int L = splint_off[i + 1] - splint_off[i];
        if (L <= 0) return fail(h, -5, "splint %d is empty", i);
        max_ls = std::max(max_ls, L);
    }
    int64_t total_sp = splint_off[n_splints];
    h->n_reads = n_reads; h->n_splints = n_splints; h->max_lr = max_lr; h->max_ls = max_ls;
    h->total_bases = total; h->total_sp = total_sp;
    CK(h->d_ascii.ensure((size_t)total + 64));
    CK(h->d_codes.ensure((size_t)total + 64));
    CK(h->d_off.ensure((size_t)(n_reads + 1) * 8));
    CK(h->d_sp_ascii.ensure((size_t)total_sp + 64));
    CK(h->d_sp_codes.ensure((size_t)total_sp + 64));
    CK(h->d_sp_off.ensure((size_t)(n_splints + 1) * 4));
    CK(h->d_sp_idx.ensure((size_t)n_reads * 4));
    // defer_reads (fused call): the read bytes follow in chunks on the copy stream, under the conk kernel of the chunk before
    if (!defer_reads) CK(cudaMemcpyAsync(h->d_ascii.p, reads, (size_t)total, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_off.p, read_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_sp_ascii.p, splints, (size_t)total_sp, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_sp_off.p, splint_off, (size_t)(n_splints + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_sp_idx.p, splint_idx, (size_t)n_reads * 4, cudaMemcpyHostToDevice, h->stream));
    if (!defer_reads) CK(h->sync());
    h->staged = true; h->ran = false;
    return 0;
}

static int encode_staged(c3_handle *h)
{
    int rc = launch_encode(h, h->d_ascii.p, h->d_codes.p, h->total_bases);
    if (rc) return rc;
    return launch_encode(h, h->d_sp_ascii.p, h->d_sp_codes.p, h->total_sp);
}

// ---------------------------------------------------------------------------
// B1
// ---------------------------------------------------------------------------
extern "C" int c3_conk_batch(c3_handle *h, int32_t n_reads, const char *reads, const int64_t *read_off,
                             int32_t n_splints, const char *splints, const int32_t *splint_off,
                             const int32_t *splint_idx, int32_t penalty, int32_t *out_profile)
{
    int rc = stage_reads(h, n_reads, reads, read_off, n_splints, splints, splint_off, splint_idx);
    if (rc) return rc;
    if (!out_profile) return fail(h, -5, "out_profile is null");
    h->tim = c3_timings{}; h->kt_used = 0;
    CK(cudaEventRecord(h->ev[0], h->stream));
    if ((rc = encode_staged(h))) return rc;
    CK(cudaEventRecord(h->ev[1], h->stream));
    if ((rc = launch_conk(h, penalty))) return rc;
    CK(cudaEventRecord(h->ev[2], h->stream));
    CK(cudaMemcpyAsync(out_profile, h->d_prof.p, (size_t)h->total_bases * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h->sync());
    cudaEventElapsedTime(&h->tim.encode_ms, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->tim.conk_ms, h->ev[1], h->ev[2]);
    h->tim.total_ms = h->tim.encode_ms + h->tim.conk_ms;
    return 0;
}

// ---------------------------------------------------------------------------
// f-3: splint assignment (replaces the BLAT pre-step, bin/preprocess.py:12-45,74-76, with the conk kernel)
// ---------------------------------------------------------------------------
__global__ void c3_profile_max_kernel(int n, const int32_t *__restrict__ prof, const int64_t *__restrict__ off,
                                      int32_t *__restrict__ out)
{
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= n) return;
    const int64_t a = off[r], b = off[r + 1];
    int m = 0;
    for (int64_t i = a + lane; i < b; i += 32) m = max(m, prof[i]);
    m = __reduce_max_sync(C3_FULL, m);
    if (lane == 0) out[r] = m;
}

__global__ void c3_fill_i32_kernel(int32_t *p, int n, int v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int c3_assign_splints(c3_handle *h, int32_t n_reads, const char *reads, const int64_t *read_off,
                                 int32_t n_cands, const char *cands, const int32_t *cand_off, int32_t penalty,
                                 int32_t *out_best, int32_t *out_scores)
{
    if (!h) return -1;
    if (n_cands <= 0 || !out_best) return fail(h, -5, "bad arguments");
    std::vector<int32_t> zero((size_t)std::max(n_reads, 1), 0);
    int rc = stage_reads(h, n_reads, reads, read_off, n_cands, cands, cand_off, zero.data());
    if (rc) return rc;
    h->tim = c3_timings{}; h->kt_used = 0;
    CK(h->d_status.ensure((size_t)n_cands * n_reads * 4));          // scores [n_cands][n_reads]
    CK(cudaEventRecord(h->ev[0], h->stream));
    if ((rc = encode_staged(h))) return rc;
    for (int c = 0; c < n_cands; ++c) {
        c3_fill_i32_kernel<<<(n_reads + 255) / 256, 256, 0, h->stream>>>(h->d_sp_idx.as<int32_t>(), n_reads, c);
        CK(cudaGetLastError());
        if ((rc = launch_conk(h, penalty))) return rc;
        c3_profile_max_kernel<<<(n_reads + 3) / 4, 128, 0, h->stream>>>(n_reads, h->d_prof.as<int32_t>(), h->d_off.as<int64_t>(),
                                                                        h->d_status.as<int32_t>() + (size_t)c * n_reads);
        CK(cudaGetLastError());
        h->tim.kernel_launches += 2;
    }
    CK(cudaEventRecord(h->ev[1], h->stream));
    std::vector<int32_t> sc((size_t)n_cands * n_reads);
    CK(cudaMemcpyAsync(sc.data(), h->d_status.p, sc.size() * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h->sync());
    cudaEventElapsedTime(&h->tim.conk_ms, h->ev[0], h->ev[1]);
    h->tim.total_ms = h->tim.conk_ms;
    for (int r = 0; r < n_reads; ++r) {
        int best = 0, bv = sc[r];
        for (int c = 1; c < n_cands; ++c) { const int v = sc[(size_t)c * n_reads + r]; if (v > bv) { bv = v; best = c; } }
        out_best[r] = best;
    }
    if (out_scores) memcpy(out_scores, sc.data(), sc.size() * 4);
    h->staged = false;
    return 0;
}

// ---------------------------------------------------------------------------
// B2
// ---------------------------------------------------------------------------
extern "C" int c3_peaks_batch(c3_handle *h, int32_t n, const int32_t *profile, const int64_t *off,
                              const double *coef, int32_t window, int32_t iters, int32_t min_dist,
                              double height_mult, double gate_mult, double *out_smoothed,
                              double *out_median, int32_t *out_peaks, int32_t max_peaks, int32_t *out_n_peaks)
{
    if (!h) return -1;
    if (n <= 0 || !profile || !off || !coef || !out_peaks || !out_n_peaks || max_peaks <= 0)
        return fail(h, -5, "bad arguments");
    CK(cudaSetDevice(h->device));
    int64_t total = off[n];
    int max_len = 0;
    for (int i = 0; i < n; ++i) max_len = std::max<int64_t>(max_len, off[i + 1] - off[i]);
    CK(h->d_prof.ensure((size_t)total * 4 + 16));
    CK(h->d_off.ensure((size_t)(n + 1) * 8));
    CK(cudaMemcpyAsync(h->d_prof.p, profile, (size_t)total * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_off.p, off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    h->tim = c3_timings{}; h->kt_used = 0;
    CK(cudaEventRecord(h->ev[0], h->stream));
    int rc = launch_peaks(h, h->d_prof.as<int32_t>(), h->d_off.as<int64_t>(), n, max_len, coef, window, iters, min_dist,
                          height_mult, gate_mult, out_smoothed != nullptr, out_median != nullptr, max_peaks, total);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(cudaMemcpyAsync(out_peaks, h->d_peaks.p, (size_t)n * max_peaks * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(out_n_peaks, h->d_npk.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (out_smoothed) CK(cudaMemcpyAsync(out_smoothed, h->d_smoothed.p, (size_t)total * 8, cudaMemcpyDeviceToHost, h->stream));
    if (out_median) CK(cudaMemcpyAsync(out_median, h->d_median.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h->sync());
    cudaEventElapsedTime(&h->tim.peaks_ms, h->ev[0], h->ev[1]);
    h->tim.total_ms = h->tim.peaks_ms;
    h->staged = false;
    return 0;
}

// ---------------------------------------------------------------------------
// B3
// ---------------------------------------------------------------------------
extern "C" int c3_poa_batch(c3_handle *h, int32_t n_groups, const char *seqs, const int64_t *seq_off,
                            const int32_t *group_off, const c3_poa_params *params, char *out_cons,
                            int32_t cons_cap, int32_t *out_cons_len, int64_t *out_cells,
                            int32_t *out_nodes, int32_t *out_status, char *out_msa, int32_t msa_cap,
                            int32_t *out_msa_len)
{
    if (!h) return -1;
    if (n_groups <= 0 || !seqs || !seq_off || !group_off || !params || !out_cons || cons_cap <= 0 || !out_cons_len || !out_status)
        return fail(h, -5, "bad arguments");
    const bool want_msa = out_msa != nullptr;
    if (want_msa && (msa_cap <= 0 || !out_msa_len || cons_cap < 2 * msa_cap)) return fail(h, -5, "out_msa needs msa_cap > 0, out_msa_len and cons_cap >= 2*msa_cap");
    CK(cudaSetDevice(h->device));
    const int n_seqs = group_off[n_groups];
    const int64_t total = seq_off[n_seqs];
    int max_nseq = 0, max_q = 0; int64_t max_total = 0;
    std::vector<int64_t> item_base(n_groups);
    std::vector<int32_t> nseq(n_groups);
    for (int g = 0; g < n_groups; ++g) {
        nseq[g] = group_off[g + 1] - group_off[g];
        max_nseq = std::max(max_nseq, nseq[g]);
    }
    if (max_nseq <= 0) return fail(h, -5, "empty groups");
    std::vector<int32_t> bounds((size_t)n_groups * max_nseq * 2, 0);
    for (int g = 0; g < n_groups; ++g) {
        const int s0 = group_off[g];
        item_base[g] = seq_off[s0];
        int64_t tot = 0;
        for (int k = 0; k < nseq[g]; ++k) {
            const int64_t a = seq_off[s0 + k] - item_base[g], b = seq_off[s0 + k + 1] - item_base[g];
            if (b - a > 65000 || b > 0x7fffffff) return fail(h, -5, "sequence too long in group %d", g);
            bounds[((size_t)g * max_nseq + k) * 2] = (int32_t)a;
            bounds[((size_t)g * max_nseq + k) * 2 + 1] = (int32_t)b;
            max_q = std::max<int>(max_q, (int)(b - a));
            tot += b - a;
        }
        max_total = std::max(max_total, tot);
    }
    CK(h->d_ascii.ensure((size_t)total + 64));
    CK(h->d_codes.ensure((size_t)total + 64));
    CK(h->d_item_base.ensure((size_t)n_groups * 8));
    CK(h->d_bounds.ensure(bounds.size() * 4));
    CK(h->d_nseq.ensure((size_t)n_groups * 4));
    CK(h->d_status.ensure((size_t)n_groups * 4));
    CK(h->d_clen.ensure((size_t)n_groups * 4));
    CK(h->d_nodes.ensure((size_t)n_groups * 4));
    CK(h->d_cells.ensure((size_t)n_groups * 8));
    CK(h->d_cons.ensure((size_t)n_groups * cons_cap));
    CK(cudaMemcpyAsync(h->d_ascii.p, seqs, (size_t)total, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_item_base.p, item_base.data(), (size_t)n_groups * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_bounds.p, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_nseq.p, nseq.data(), (size_t)n_groups * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_status.p, 0, (size_t)n_groups * 4, h->stream));
    CK(cudaMemsetAsync(h->d_clen.p, 0, (size_t)n_groups * 4, h->stream));
    CK(cudaMemsetAsync(h->d_nodes.p, 0, (size_t)n_groups * 4, h->stream));
    CK(cudaMemsetAsync(h->d_cells.p, 0, (size_t)n_groups * 8, h->stream));
    CK(cudaMemsetAsync(h->d_cons.p, 0, (size_t)n_groups * cons_cap, h->stream));
    h->tim = c3_timings{}; h->kt_used = 0;
    CK(cudaEventRecord(h->ev[0], h->stream));
    int rc = launch_encode(h, h->d_ascii.p, h->d_codes.p, total);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev[1], h->stream));
    c3_poa_args A;
    memset(&A, 0, sizeof(A));
    A.codes = h->d_codes.as<uint8_t>(); A.item_base = h->d_item_base.as<int64_t>(); A.bounds = h->d_bounds.as<int32_t>();
    A.n_seqs = h->d_nseq.as<int32_t>(); A.n_seqs_stride = 1; A.n_items = n_groups; A.max_seqs = max_nseq; A.min_seqs = 1;
    A.cons = h->d_cons.as<char>(); A.cons_cap = cons_cap; A.status = h->d_status.as<int32_t>();
    A.cons_len = h->d_clen.as<int32_t>(); A.nodes_out = h->d_nodes.as<int32_t>(); A.cells_out = h->d_cells.as<long long>();
    A.out_stride = 1; A.cells_stride = 2; A.msa2 = want_msa ? 1 : 0; A.ok_status = 0;
    {
        std::vector<int64_t> tot((size_t)n_groups);
        for (int g = 0; g < n_groups; ++g) tot[g] = seq_off[group_off[g + 1]] - seq_off[group_off[g]];
        if ((rc = upload_poa_order(h, A, nseq, tot, 1, params))) return rc;
    }
    if ((rc = launch_poa(h, A, max_q, max_nseq, max_total, params))) return rc;
    CK(cudaEventRecord(h->ev[2], h->stream));
    CK(cudaMemcpyAsync(out_cons, h->d_cons.p, (size_t)n_groups * cons_cap, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(out_cons_len, h->d_clen.p, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(out_status, h->d_status.p, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, h->stream));
    if (out_cells) CK(cudaMemcpyAsync(out_cells, h->d_cells.p, (size_t)n_groups * 8, cudaMemcpyDeviceToHost, h->stream));
    if (out_nodes) CK(cudaMemcpyAsync(out_nodes, h->d_nodes.p, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h->sync());
    if (want_msa) {      // 2-sequence groups: the consensus slot holds [row0 | row1], cons_len = columns
        memset(out_msa, '-', (size_t)n_seqs * msa_cap);
        for (int g = 0; g < n_groups; ++g) {
            out_msa_len[g] = 0;
            if (nseq[g] != 2 || out_status[g] != 0) continue;
            const int L = out_cons_len[g];
            if (L > msa_cap) { out_status[g] = C3_E_CONS; continue; }
            const char *src = out_cons + (size_t)g * cons_cap;
            memcpy(out_msa + (size_t)group_off[g] * msa_cap, src, (size_t)L);
            memcpy(out_msa + (size_t)(group_off[g] + 1) * msa_cap, src + L, (size_t)L);
            out_msa_len[g] = L; out_cons_len[g] = 0;
        }
    }
    h->kt_collect();
    cudaEventElapsedTime(&h->tim.encode_ms, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->tim.poa_ms, h->ev[1], h->ev[2]);
    h->tim.total_ms = h->tim.encode_ms + h->tim.poa_ms;
    h->tim.poa_items = n_groups;
    h->staged = false;
    return 0;
}

// repeats == 1: the consensus is the subread itself (bin/determine_consensus.py:31-32)
__global__ void c3_copy_single_kernel(int n, const uint8_t *__restrict__ ascii, const int64_t *__restrict__ off,
                                      const int32_t *__restrict__ sub, int max_peaks, char *cons, int cons_cap,
                                      c3_read_result_dev *res)
{
    const int r = blockIdx.x;
    if (r >= n) return;
    if (res[r].status != 0 || res[r].n_sub != 1) return;
    const int32_t *sb = sub + (int64_t)r * max_peaks * 2;
    const int L = sb[1] - sb[0];
    if (L > cons_cap) { if (threadIdx.x == 0) res[r].status = C3_E_CONS; return; }
    const uint8_t *src = ascii + off[r] + sb[0];
    char *dst = cons + (int64_t)r * cons_cap;
    for (int i = threadIdx.x; i < L; i += blockDim.x) dst[i] = (char)src[i];
    if (threadIdx.x == 0) res[r].cons_len = L;
}

// ---------------------------------------------------------------------------
// B4
// ---------------------------------------------------------------------------
extern "C" int c3_stage(c3_handle *h, int32_t n_reads, const char *reads, const int64_t *read_off,
                        int32_t n_splints, const char *splints, const int32_t *splint_off, const int32_t *splint_idx)
{
    return stage_reads(h, n_reads, reads, read_off, n_splints, splints, splint_off, splint_idx);
}

// pipe_reads / pipe_off != null (fused call): the reads' bytes are still on the host; they are copied in up to 8 chunks on
// a second stream, and chunk k's encode + conk run while chunk k + 1 is on its way (PCIe under the conk kernel).
static int run_impl(c3_handle *h, int32_t penalty, const double *coef, int32_t window, int32_t iters,
                    int32_t min_dist, const c3_poa_params *params, int32_t max_peaks, int32_t cons_cap,
                    const char *pipe_reads, const int64_t *pipe_off)
{
    if (!h) return -1;
    if (!h->staged) return fail(h, -8, "c3_run without c3_stage");
    if (!coef || !params || max_peaks <= 0 || max_peaks > C3_SPLIT_MAXP || cons_cap <= 0) return fail(h, -5, "bad arguments");
    CK(cudaSetDevice(h->device));
    const int n = h->n_reads;
    h->max_peaks = max_peaks; h->cons_cap = cons_cap;
    h->tim = c3_timings{}; h->kt_used = 0;
    CK(h->d_sub.ensure((size_t)n * max_peaks * 2 * 4 + 16));
    CK(h->d_dang.ensure((size_t)n * 4 * 4 + 16));
    CK(h->d_res.ensure((size_t)n * sizeof(c3_read_result)));
    CK(h->d_stats.ensure(64));
    CK(h->d_cons.ensure((size_t)n * cons_cap + 16));
    CK(cudaMemsetAsync(h->d_stats.p, 0, 64, h->stream));
    // outputs are only partly written by the kernels (first n_peaks / n_sub / cons_len entries): define the rest
    CK(cudaMemsetAsync(h->d_sub.p, 0, (size_t)n * max_peaks * 2 * 4, h->stream));
    CK(cudaMemsetAsync(h->d_cons.p, 0, (size_t)n * cons_cap, h->stream));
    int rc;
    CK(cudaEventRecord(h->ev[0], h->stream));
    if (!pipe_reads) {
        if ((rc = encode_staged(h))) return rc;
        CK(cudaEventRecord(h->ev[1], h->stream));
        if ((rc = launch_conk(h, penalty))) return rc;
    } else {
        if (!h->stream2) {
            CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
            for (int k = 0; k < C3_PIPE_CHUNKS; ++k) CK(cudaEventCreateWithFlags(&h->ev_chunk[k], cudaEventDisableTiming));
        }
        if ((rc = launch_encode(h, h->d_sp_ascii.p, h->d_sp_codes.p, h->total_sp))) return rc;
        CK(cudaEventRecord(h->ev[1], h->stream));                 // (encode_ms then only covers the splints)
        // (chunks of at least 4 096 reads: what the packed conk kernel wants to fill the grid with pairs)
        const int nchunk = h->total_bases >= (64ll << 20) ? std::max(1, std::min(C3_PIPE_CHUNKS, n / 4096)) : 1;
        int rb[C3_PIPE_CHUNKS + 1]; int64_t bb[C3_PIPE_CHUNKS + 1];
        for (int k = 0; k <= nchunk; ++k) {
            // chunk borders at reads holding equal shares of the bytes; byte ranges start 16-aligned (the encode kernel
            // works on 16-byte words) and overlap by at most one word, which is copied and encoded twice
            int r = k == nchunk ? n : (int)(std::lower_bound(pipe_off, pipe_off + n, h->total_bases * k / nchunk) - pipe_off);
            rb[k] = std::min(r, n);
            bb[k] = pipe_off[rb[k]];
        }
        // the copy stream may only overwrite d_ascii once everything queued on the launch stream so far has passed
        CK(cudaEventRecord(h->ev_chunk[0], h->stream));
        CK(cudaStreamWaitEvent(h->stream2, h->ev_chunk[0], 0));
        for (int k = 0; k < nchunk; ++k) {
            const int64_t b0 = bb[k] & ~15ll, b1 = std::min<int64_t>(h->total_bases, (bb[k + 1] + 15) & ~15ll);
            if (b1 > b0) CK(cudaMemcpyAsync(h->d_ascii.as<uint8_t>() + b0, pipe_reads + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, h->stream2));
            CK(cudaEventRecord(h->ev_chunk[k], h->stream2));
        }
        for (int k = 0; k < nchunk; ++k) {
            const int64_t b0 = bb[k] & ~15ll, b1 = std::min<int64_t>(h->total_bases, (bb[k + 1] + 15) & ~15ll);
            CK(cudaStreamWaitEvent(h->stream, h->ev_chunk[k], 0));
            if ((rc = launch_encode(h, h->d_ascii.as<uint8_t>() + b0, h->d_codes.as<uint8_t>() + b0, b1 - b0))) return rc;
            if ((rc = launch_conk(h, penalty, rb[k], rb[k + 1], k))) return rc;
        }
    }
    CK(cudaEventRecord(h->ev[2], h->stream));
    if ((rc = launch_peaks(h, h->d_prof.as<int32_t>(), h->d_off.as<int64_t>(), n, h->max_lr, coef, window, iters,
                           min_dist, 3.0, 6.0, false, false, max_peaks, h->total_bases))) return rc;
    CK(cudaEventRecord(h->ev[3], h->stream));
    c3_split_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(
        n, h->d_off.as<int64_t>(), h->d_sp_off.as<int32_t>(), h->d_sp_idx.as<int32_t>(), h->d_peaks.as<int32_t>(),
        h->d_npk.as<int32_t>(), max_peaks, h->d_sub.as<int32_t>(), h->d_dang.as<int32_t>(),
        h->d_res.as<c3_read_result_dev>(), h->d_stats.as<int32_t>());
    CK(cudaGetLastError());
    h->tim.kernel_launches++;
    c3_copy_single_kernel<<<n, 128, 0, h->stream>>>(n, h->d_ascii.as<uint8_t>(), h->d_off.as<int64_t>(),
                                                     h->d_sub.as<int32_t>(), max_peaks, h->d_cons.as<char>(), cons_cap,
                                                     h->d_res.as<c3_read_result_dev>());
    CK(cudaGetLastError());
    h->tim.kernel_launches++;
    CK(cudaEventRecord(h->ev[4], h->stream));
    // The host's only look at the batch between the kernels: 16 bytes of batch maxima from the split kernel and 64 bytes
    // of list sizes from the work-order kernels, one synchronisation (inside build_poa_order_device); they size the POA
    // workspaces.  The per-read repeat counts and subread bases (left in poa_cells by the split kernel) stay on the
    // device: the work order is a histogram by cost class + scan + scatter there.
    int32_t stats[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(stats, h->d_stats.p, 16, cudaMemcpyDeviceToHost, h->stream));
    {
        c3_poa_args A;
        memset(&A, 0, sizeof(A));
        c3_read_result *res = h->d_res.as<c3_read_result>();
        A.codes = h->d_codes.as<uint8_t>(); A.item_base = h->d_off.as<int64_t>(); A.bounds = h->d_sub.as<int32_t>();
        A.n_seqs = &res->n_sub; A.n_seqs_stride = sizeof(c3_read_result) / 4; A.n_items = n; A.max_seqs = max_peaks; A.min_seqs = 2;
        A.msa2 = 1; A.ok_status = 2;    // 2-repeat reads: [row0 | row1] of the pairwise MSA in the consensus slot
        if ((rc = build_poa_order_device(h, A, n, h->d_res.as<c3_read_result_dev>(), 2, params))) return rc;
        h->tim.poa_items = stats[3];
        A.cons = h->d_cons.as<char>(); A.cons_cap = cons_cap; A.status = &res->status; A.cons_len = &res->cons_len;
        A.nodes_out = &res->poa_nodes; A.cells_out = (long long *)&res->poa_cells;
        A.out_stride = sizeof(c3_read_result) / 4; A.cells_stride = sizeof(c3_read_result) / 4;
        if (stats[3] > 0 && (rc = launch_poa(h, A, stats[0], stats[1], stats[2], params))) return rc;
    }
    CK(cudaEventRecord(h->ev[5], h->stream));
    CK(h->sync());
    cudaEventElapsedTime(&h->tim.encode_ms, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->tim.conk_ms, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&h->tim.peaks_ms, h->ev[2], h->ev[3]);
    cudaEventElapsedTime(&h->tim.split_ms, h->ev[3], h->ev[4]);
    if (getenv("C3POA_GRP_TIMING") && h->kt_used > 1) {          // tuning only: where the POA stage's non-kernel time sits
        float head = 0.f, tail = 0.f, span = 0.f;
        cudaEventElapsedTime(&head, h->ev[4], h->kt_ev[0]);
        cudaEventElapsedTime(&tail, h->kt_ev[h->kt_used - 1], h->ev[5]);
        cudaEventElapsedTime(&span, h->kt_ev[0], h->kt_ev[h->kt_used - 1]);
        fprintf(stderr, "poa stage: %.2f ms before the first POA kernel, %.2f ms first..last kernel, %.2f ms after\n", head, span, tail);
    }
    h->kt_collect();
    cudaEventElapsedTime(&h->tim.poa_ms, h->ev[4], h->ev[5]);
    cudaEventElapsedTime(&h->tim.total_ms, h->ev[0], h->ev[5]);
    h->ran = true;
    return 0;
}

extern "C" int c3_run(c3_handle *h, int32_t penalty, const double *coef, int32_t window, int32_t iters,
                      int32_t min_dist, const c3_poa_params *params, int32_t max_peaks, int32_t cons_cap)
{
    return run_impl(h, penalty, coef, window, iters, min_dist, params, max_peaks, cons_cap, nullptr, nullptr);
}

extern "C" int c3_fetch(c3_handle *h, int32_t *out_peaks, int32_t *out_sub_bounds, int32_t *out_dang_bounds,
                        char *out_cons, c3_read_result *out_results)
{
    if (!h) return -1;
    if (!h->ran) return fail(h, -8, "c3_fetch without c3_run");
    if (!out_results) return fail(h, -5, "out_results is null");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->n_reads;
    if (out_peaks) CK(cudaMemcpyAsync(out_peaks, h->d_peaks.p, n * h->max_peaks * 4, cudaMemcpyDeviceToHost, h->stream));
    if (out_sub_bounds) CK(cudaMemcpyAsync(out_sub_bounds, h->d_sub.p, n * h->max_peaks * 8, cudaMemcpyDeviceToHost, h->stream));
    if (out_dang_bounds) CK(cudaMemcpyAsync(out_dang_bounds, h->d_dang.p, n * 16, cudaMemcpyDeviceToHost, h->stream));
    if (out_cons) CK(cudaMemcpyAsync(out_cons, h->d_cons.p, n * h->cons_cap, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(out_results, h->d_res.p, n * sizeof(c3_read_result), cudaMemcpyDeviceToHost, h->stream));
    CK(h->sync());
    return 0;
}

extern "C" int c3_consensus_batch(c3_handle *h, int32_t n_reads, const char *reads, const int64_t *read_off,
                                  int32_t n_splints, const char *splints, const int32_t *splint_off,
                                  const int32_t *splint_idx, int32_t penalty, const double *coef,
                                  int32_t window, int32_t iters, int32_t min_dist,
                                  const c3_poa_params *params, int32_t max_peaks, int32_t cons_cap,
                                  int32_t *out_peaks, int32_t *out_sub_bounds, int32_t *out_dang_bounds,
                                  char *out_cons, c3_read_result *out_results)
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    int rc = stage_reads(h, n_reads, reads, read_off, n_splints, splints, splint_off, splint_idx, /*defer_reads=*/true);
    if (rc) return rc;
    if ((rc = run_impl(h, penalty, coef, window, iters, min_dist, params, max_peaks, cons_cap, reads, read_off))) {
        // the deferred staging queued copies from the caller's buffers without waiting: nothing may still read them
        // once this call has returned
        (void)cudaStreamSynchronize(h->stream);
        if (h->stream2) (void)cudaStreamSynchronize(h->stream2);
        return rc;
    }
    if ((rc = c3_fetch(h, out_peaks, out_sub_bounds, out_dang_bounds, out_cons, out_results))) return rc;
    return 0;
}

// ---------------------------------------------------------------------------
// integer-pipe peak micro-benchmark (roofline denominator for the DP kernels)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) c3_intpeak_kernel(int iters, int *sink)
{
    int a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const int b = blockIdx.x | 1, c = -7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = __viaddmax_s32(a0, c, b); a1 = __viaddmax_s32(a1, c, b); a2 = __viaddmax_s32(a2, c, b); a3 = __viaddmax_s32(a3, c, b);
            a4 = __viaddmax_s32(a4, c, b); a5 = __viaddmax_s32(a5, c, b); a6 = __viaddmax_s32(a6, c, b); a7 = __viaddmax_s32(a7, c, b);
        }
    }
    if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x7fffffff) *sink = a0;
}

#ifdef C3_POA_STATS
extern "C" int c3_debug_stats(unsigned long long *out16, int reset)
{
    if (out16) cudaMemcpyFromSymbol(out16, c3_poa_stats, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(c3_poa_stats, z, sizeof(z)); }
    return 0;
}
#endif

#ifdef C3L_PROF
extern "C" int c3_debug_lane_prof(unsigned long long *out24, int reset)
{
    if (out24) cudaMemcpyFromSymbol(out24, c3l_prof, sizeof(unsigned long long) * 24);
    if (reset) { unsigned long long z[24] = {0}; cudaMemcpyToSymbol(c3l_prof, z, sizeof(z)); }
    return 0;
}
#endif

extern "C" int c3_measure_int_peak(c3_handle *h, double *out_ops_per_s)
{
    if (!h || !out_ops_per_s) return -1;
    CK(cudaSetDevice(h->device));
    CK(h->d_stats.ensure(64));
    const int iters = 4096, grid = h->sm_count * 8;
    double best = 0;
    for (int rep = 0; rep < 12; ++rep) {
        CK(cudaEventRecord(h->ev[6], h->stream));
        c3_intpeak_kernel<<<grid, 256, 0, h->stream>>>(iters, h->d_stats.as<int>());
        CK(cudaEventRecord(h->ev[7], h->stream));
        CK(h->sync());
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]);
        // one VIADDMNMX = 2 integer ops (add + max)
        double ops = 2.0 * 64.0 * iters * 256.0 * grid;
        if (rep >= 2 && ms > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    *out_ops_per_s = best;
    return 0;
}
