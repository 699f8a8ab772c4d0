// blu_capi.cu -- C ABI of libbluest_b200.so (see include/bluest_b200.h).
// Host-side orchestration only: argument checks, HBM allocation, kernel launches on the
// context's stream, host<->device copies.  All arithmetic of the path runs in the kernels of
// blu_invert.cuh / blu_phi.cuh / blu_grad.cuh / blu_hess.cuh / blu_gram.cuh / blu_level1.cuh.
// There is no CPU fallback: without a CUDA device every entry point fails with BLU_ERR_NODEVICE.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "blu_common.cuh"
#include "blu_jacobi.cuh"
#include "blu_stream.cuh"
#include "blu_launch.h"
#include "blu_phi.cuh"
#include "blu_grad.cuh"
#include "blu_soa_types.h"
#include "blu_hess.cuh"
#include "blu_matvec.cuh"
#include "blu_gram.cuh"
#include "blu_intproj.cuh"
#include "blu_level1.cuh"
#include "blu_kkt.cuh"
#include "blu_batch.cuh"

// --------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(e_ == cudaErrorMemoryAllocation ? BLU_ERR_NOMEM : BLU_ERR_CUDA,        \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define KERNEL_CHECK(ctx)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess)                                                                 \
            return fail(BLU_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
        (ctx)->launches++;                                                                     \
    } while (0)

struct blu_ctx {
    int device = 0, N = 0, K = 0, NP = 0, NCH = 0, nsm = 148;
    long long L = 0, Lpad = 0, ldH = 0;
    std::vector<long long> sizes;      // K entries
    std::vector<BluClass> cls;         // non-empty classes only
    std::vector<int> cls_of_k;         // k -> index in cls or -1
    int Tmax = 1;
    BluClass *d_cls = nullptr;
    uint8_t *d_gidx = nullptr;
    unsigned *d_gmask = nullptr;
    uint16_t *d_lut = nullptr;
    unsigned *d_plut = nullptr;          // bank-aware run table of the Phi kernel (blu_phi.cuh)
    BluChunk *d_wchunks = nullptr;       // Phi kernel: its chunk list laid out warp by warp ...
    int *d_wstart = nullptr;             // ... and the first chunk of every warp (+ end sentinel)
    int phi_stages = 2, phi_ns = 2, phi_sd = 0, idsd = 64;
    double *d_cinv = nullptr;
    long long cinv_len = 0, gidx_len = 0;
    double *d_C = nullptr, *d_m = nullptr, *d_part = nullptr, *d_phi = nullptr, *d_pinv = nullptr;
    double *d_x = nullptr, *d_S = nullptr, *d_grad = nullptr, *d_U = nullptr, *d_V = nullptr, *d_H = nullptr;
    long long H_rows = 0;              // rows allocated in d_H
    BluEvalHeader *d_hdr = nullptr, *h_hdr = nullptr;
    int grid_phi = 1, grid_grad = 1;
    double *d_soa = nullptr;           // group-interleaved (SoA tile) copy of the inverses, blu_soa.cuh
    long long *d_soff = nullptr;       // per class: offset of its tiles in d_soa
    std::vector<long long> soff;
    long long soa_len = 0;
    bool soa_valid = false, tiles_valid = false, use_soa = true;
    BluTile *d_tiles = nullptr;
    int ntiles = 0, grid_soa = 1;
    BluChunk *d_chunks = nullptr;      // work list of the owned slice (blu_stream.cuh)
    int nchunks = 0, lutlen = 0, plutlen = 0, part_rows = 0, phi_warps = BLU_PHI_WARPS;
    int sd = BLU_CHUNK_DOUBLES + 4;    // stage size (doubles) of the streaming kernels
    bool have_inv = false;
    bool hess_attr_done[3] = {false, false, false};
    bool hess_onebuf = true;           // symmetric Hessian kernel: single staging buffer, 4 CTAs per SM (blu_hess.cuh)
    std::vector<char> inv_set;         // per class: inverses present
    long long lo = 0, hi = 0;          // owned slice of the flat enumeration
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool timed = false;
    int launches = 0;
    double *h_m = nullptr, *h_grad = nullptr;   // pinned staging for m and the gradient (async begin/end pair)
    double *pend_hess = nullptr;       // host destination of the pending evaluation's Hessian
    bool pending = false;
    std::vector<cudaEvent_t> panel_ev; // one event per Hessian row panel (symmetric download)
    int mirror_threads = 0;            // 0: automatic (see download_hessian_symmetric)
    bool sym_download = true;          // dense Hessian to the host: upper block-triangle over PCIe, lower mirrored by host threads
    int sym_full_rows_pct = 10;        // share (%) of the lower triangle that still travels by DMA (the bottom rows, in full)
    bool sym_pct_auto = true;          // adapt that share to the host: more DMA when the mirroring threads lag, less when they idle
    int sym_calm = 0;                  // consecutive downloads in which the mirror finished with the DMA
    BluXchg *d_xchg = nullptr;         // this rank's exchange buffer (CUDA IPC shared)
    BluPeers peers{};                  // peers as mapped here; world == 0: not connected
    std::vector<void *> ipc_opened;
    double *d_Sop = nullptr;           // S = 2 pinv(Phi) of the evaluation the resident U factor belongs to (the operator's own copy)
    double *d_hvpart = nullptr, *d_hvp = nullptr, *d_hvout = nullptr;   // Hessian mat-vec: CTA partials of t, staged p and H p
    void *kkt_ws[4] = {nullptr, nullptr, nullptr, nullptr};            // blu_kkt_solve workspace (Bs | small | vectors | partial tiles), kept between solves
    size_t kkt_cap[4] = {0, 0, 0, 0};
    int hv_grid = 1;
    bool uv_ready = false, v_ready = false;   // U (and V) hold the factors of the last want_hess evaluation
    std::vector<cudaEvent_t> evlog;    // optional per-evaluation event log (4 events per evaluation)
    int evlog_n = 0;
    bool capturing = false;            // between blu_ctx_graph_begin and _end: the stream records instead of running
    std::vector<cudaGraphExec_t> graphs;
    double *d_grad_out = nullptr;      // where the gradient kernels write (default: d_grad)
    bool is_clone = false;             // shares the read-only tables / inverses / work lists of another context (blu_ctx_clone)
};

static int use(blu_ctx *c)
{
    if (!c) return fail(BLU_ERR_ARG, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    return BLU_OK;
}

extern "C" const char *blu_last_error(void) { return g_err.c_str(); }
extern "C" const char *blu_version(void) { return "bluest_b200 0.1 (sm_100a, fp64)"; }

extern "C" int blu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int need_device(int device)
{
    int n = blu_device_count();
    if (n <= 0) return fail(BLU_ERR_NODEVICE, "no CUDA device visible: bluest_b200 has no CPU fallback");
    if (device < 0 || device >= n) return fail(BLU_ERR_ARG, "device %d out of range (%d visible)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    return BLU_OK;
}

extern "C" int blu_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(BLU_ERR_ARG, "null out");
    if (blu_device_count() <= 0) return fail(BLU_ERR_NODEVICE, "no CUDA device");
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return BLU_OK;
}
extern "C" int blu_host_free(void *p)
{
    if (p) CUDA_TRY(cudaFreeHost(p));
    return BLU_OK;
}

// --------------------------------------------------------------------------------------------
// context
// --------------------------------------------------------------------------------------------
extern "C" int blu_ctx_destroy(blu_ctx *c)
{
    if (!c) return BLU_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (!c->is_clone) {                // a clone borrows these from its parent
        cudaFree(c->d_cls); cudaFree(c->d_gidx); cudaFree(c->d_gmask); cudaFree(c->d_lut); cudaFree(c->d_plut); cudaFree(c->d_cinv);
        cudaFree(c->d_C); cudaFree(c->d_chunks); cudaFree(c->d_wchunks); cudaFree(c->d_wstart); cudaFree(c->d_soa); cudaFree(c->d_soff); cudaFree(c->d_tiles);
    }
    cudaFree(c->d_m); cudaFree(c->d_part); cudaFree(c->d_phi); cudaFree(c->d_pinv);
    cudaFree(c->d_x); cudaFree(c->d_S); cudaFree(c->d_grad); cudaFree(c->d_U); cudaFree(c->d_V); cudaFree(c->d_H);
    cudaFree(c->d_hdr);
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    for (auto g : c->graphs) if (g) cudaGraphExecDestroy(g);
    cudaFree(c->d_xchg);
    cudaFree(c->d_hvpart); cudaFree(c->d_hvp); cudaFree(c->d_hvout); cudaFree(c->d_Sop);
    for (int q = 0; q < 4; ++q) { if (c->kkt_ws[q]) cudaFree(c->kkt_ws[q]); c->kkt_ws[q] = nullptr; c->kkt_cap[q] = 0; }
    if (c->h_hdr) cudaFreeHost(c->h_hdr);
    if (c->h_m) cudaFreeHost(c->h_m);
    if (c->h_grad) cudaFreeHost(c->h_grad);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->evlog) if (e) cudaEventDestroy(e);
    for (auto &e : c->panel_ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return BLU_OK;
}

// Work list of the streaming kernels: chunks of consecutive groups of one class inside [lo,hi).
static int build_chunks(blu_ctx *c)
{
    std::vector<BluChunk> ch;
    // chunk payload: up to a full stage, but small enough that every resident warp gets ~3 chunks
    long long total = 0;
    for (const BluClass &ci : c->cls) {
        const long long i0 = std::max<long long>(c->lo - ci.goff, 0), i1 = std::min<long long>(c->hi - ci.goff, ci.Lk);
        if (i1 > i0) total += (i1 - i0) * ci.T;
    }
    // stage payload: at least the largest group block, at most ~2 KB beyond it (small stages = more CTAs/SM)
    // (big problems amortise the per-chunk overhead better with full 4 KB chunks than they gain from occupancy)
    const long long share = total / ((long long)c->nsm * BLU_STREAM_WARPS * 2 * 3);
    const long long maxpay = std::min<long long>(BLU_CHUNK_DOUBLES, std::max<long long>(std::max<long long>(((long long)c->Tmax + 15) / 16 * 16, 272), share));
    const long long cap = std::max<long long>(std::max<long long>(96, c->Tmax), std::min<long long>(maxpay, total / ((long long)c->nsm * BLU_STREAM_WARPS * 2 * 3)));
    c->sd = (int)(((maxpay + 4) + 1) / 2 * 2);
    for (size_t ic = 0; ic < c->cls.size(); ++ic) {
        const BluClass &ci = c->cls[ic];
        const long long i0 = std::max<long long>(c->lo - ci.goff, 0), i1 = std::min<long long>(c->hi - ci.goff, ci.Lk);
        if (i1 <= i0) continue;
        const int G = (int)std::max<long long>(1, std::min<long long>(32, cap / ci.T));
        for (long long i = i0; i < i1; i += G) {
            BluChunk b; b.cls = (int)ic; b.g = (int)std::min<long long>(G, i1 - i); b.i0 = i;
            ch.push_back(b);
        }
    }
    if (c->d_chunks) { CUDA_TRY(cudaStreamSynchronize(c->stream)); CUDA_TRY(cudaFree(c->d_chunks)); c->d_chunks = nullptr; }
    c->nchunks = (int)ch.size();
    CUDA_TRY(cudaMalloc(&c->d_chunks, sizeof(BluChunk) * std::max<size_t>(ch.size(), 1)));
    if (!ch.empty()) CUDA_TRY(cudaMemcpyAsync(c->d_chunks, ch.data(), sizeof(BluChunk) * ch.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // launch geometry: ~4 chunks per warp, at most two CTAs per SM for the Phi kernel (its partial
    // tiles are reduced by one CTA afterwards), a few more for the gradient kernels
    c->grid_grad = (int)std::min<long long>(std::max<long long>(1, ((long long)c->nchunks + BLU_STREAM_WARPS - 1) / BLU_STREAM_WARPS), (long long)c->nsm * 5);
    {   // Phi kernel (blu_phi.cuh): its own chunk list (chunk payload and ring depth differ from the gradient kernels'),
        // laid out warp by warp -- every warp walks a few CONTIGUOUS runs of the list (the register run-length
        // accumulation carries on across chunks), the runs cut at equal cost and dealt round-robin.
        const int ns = c->phi_stages == 4 ? 4 : 2;
        const long long paycap = ns == 4 ? 272 : BLU_CHUNK_DOUBLES;
        const long long pmaxpay = std::max<long long>(((long long)c->Tmax + 15) / 16 * 16, std::min<long long>(paycap, std::max<long long>(272, share)));
        const long long pcap = std::max<long long>(std::max<long long>(96, c->Tmax), std::min<long long>(pmaxpay, total / ((long long)c->nsm * BLU_PHI_WARPS * 3)));
        std::vector<BluChunk> pch;
        long long idmax = 16;
        for (size_t ic = 0; ic < c->cls.size(); ++ic) {
            const BluClass &ci = c->cls[ic];
            const long long i0 = std::max<long long>(c->lo - ci.goff, 0), i1 = std::min<long long>(c->hi - ci.goff, ci.Lk);
            if (i1 <= i0) continue;
            const int G = (int)std::max<long long>(1, std::min<long long>(32, pcap / ci.T));
            idmax = std::max<long long>(idmax, (long long)G * ci.k);
            for (long long i = i0; i < i1; i += G) {
                BluChunk b; b.cls = (int)ic; b.g = (int)std::min<long long>(G, i1 - i); b.i0 = i;
                pch.push_back(b);
            }
        }
        c->phi_ns = ns;
        c->phi_sd = (int)(((pmaxpay + 4) + 1) / 2 * 2);
        c->idsd = (int)((idmax + 15) / 16 * 16 + 48);            // G*k bytes + 16-byte skew, rounded; slack for idle lanes
        c->phi_warps = blu_phi_smem_bytes(ns, c->phi_sd, c->idsd, c->N, (int)c->cls.size(), BLU_PHI_WARPS) <= 200 * 1024 ? BLU_PHI_WARPS : 8;
        int per_sm = BLU_PHI_WARPS / c->phi_warps;                // CTAs per SM: what registers and shared memory allow
        {
            int occ = 0;
            const size_t smem = std::max(blu_phi_smem_bytes(ns, c->phi_sd, c->idsd, c->N, (int)c->cls.size(), c->phi_warps), (size_t)BLU_FIN_SCRATCH_BYTES);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blu_phi_partial_kernel, c->phi_warps * 32, smem) == cudaSuccess && occ > 0) per_sm = occ;
            else (void)cudaGetLastError();
        }
        c->grid_phi = (int)std::min<long long>(std::max<long long>(1, ((long long)pch.size() + c->phi_warps * 3 - 1) / (c->phi_warps * 3)),
                                               (long long)c->nsm * per_sm);
        c->grid_phi = std::min(c->grid_phi, BLU_PHI_GROUP * BLU_PHI_MAXGROUPS);
        const int nW = c->grid_phi * c->phi_warps;
        const int R = pch.size() >= (size_t)nW * 36 ? 12 : (pch.size() >= (size_t)nW * 12 ? 4 : 1);      // runs per warp
        const size_t nRuns = (size_t)nW * R;
        std::vector<long long> pre(pch.size() + 1, 0);
        for (size_t i = 0; i < pch.size(); ++i) pre[i + 1] = pre[i] + (long long)pch[i].g * (c->cls[(size_t)pch[i].cls].psteps + 1) + 6;
        std::vector<size_t> rs(nRuns + 1, pch.size());
        rs[0] = 0;
        {
            size_t pos = 0;
            for (size_t j = 1; j < nRuns; ++j) {
                const long long want_cost = (long long)((__int128)pre[pch.size()] * (long long)j / (long long)nRuns);
                while (pos < pch.size() && pre[pos] < want_cost) ++pos;
                rs[j] = pos;
            }
        }
        std::vector<BluChunk> wch;
        wch.reserve(pch.size());
        std::vector<int> wst((size_t)nW + 1, 0);
        for (int w = 0; w < nW; ++w) {
            wst[(size_t)w] = (int)wch.size();
            for (int r = 0; r < R; ++r) {
                const size_t j = (size_t)r * nW + ((r & 1) ? nW - 1 - w : w);   // serpentine: errors of the cost model that grow along the list cancel
                for (size_t q = rs[j]; q < rs[j + 1]; ++q) wch.push_back(pch[q]);
            }
        }
        wst[(size_t)nW] = (int)wch.size();
        if (c->d_wchunks) CUDA_TRY(cudaFree(c->d_wchunks));
        if (c->d_wstart) CUDA_TRY(cudaFree(c->d_wstart));
        c->d_wchunks = nullptr; c->d_wstart = nullptr;
        CUDA_TRY(cudaMalloc(&c->d_wchunks, sizeof(BluChunk) * std::max<size_t>(wch.size(), 1)));
        CUDA_TRY(cudaMalloc(&c->d_wstart, sizeof(int) * wst.size()));
        if (!wch.empty()) CUDA_TRY(cudaMemcpy(c->d_wchunks, wch.data(), sizeof(BluChunk) * wch.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(c->d_wstart, wst.data(), sizeof(int) * wst.size(), cudaMemcpyHostToDevice));
    }
    if (c->grid_phi > c->part_rows) {
        if (c->d_part) CUDA_TRY(cudaFree(c->d_part));
        c->d_part = nullptr;
        CUDA_TRY(cudaMalloc(&c->d_part, sizeof(double) * (size_t)c->N * c->N * (c->grid_phi + BLU_PHI_MAXGROUPS + 1)));   // + the group sums of the in-kernel reduction
        c->part_rows = c->grid_phi;
    }
    return BLU_OK;
}

extern "C" int blu_ctx_create(int device, int N, int K, const int64_t *sizes, const int64_t *groups_flat,
                              blu_ctx **out)
{
    if (!out) return fail(BLU_ERR_ARG, "null out");
    *out = nullptr;
    if (N < 1 || N > BLU_MAX_MODELS) return fail(BLU_ERR_ARG, "N=%d outside [1,%d]", N, BLU_MAX_MODELS);
    if (K < 1 || K > N) return fail(BLU_ERR_ARG, "K=%d outside [1,N=%d]", K, N);
    if (!sizes) return fail(BLU_ERR_ARG, "null sizes");
    int rc = need_device(device);
    if (rc) return rc;

    blu_ctx *c = new (std::nothrow) blu_ctx();
    if (!c) return fail(BLU_ERR_NOMEM, "host allocation failed");
    c->device = device; c->N = N; c->K = K;
    c->NCH = (N + 3) / 4; c->NP = 4 * c->NCH;
    c->sizes.assign(sizes, sizes + K);
    c->cls_of_k.assign(K + 1, -1);
    long long L = 0, ioff = 0, coff = 0; int lutoff = 0;
    for (int k = 1; k <= K; ++k) {
        long long Lk = sizes[k - 1];
        if (Lk < 0) { delete c; return fail(BLU_ERR_ARG, "negative class size"); }
        if (Lk > 0) {
            BluClass ci{};
            ci.k = k; ci.T = blu_tri(k); ci.Lk = Lk; ci.goff = L; ci.ioff = ioff; ci.coff = coff; ci.lutoff = lutoff;
            c->cls_of_k[k] = (int)c->cls.size();
            c->cls.push_back(ci);
            ioff += Lk * k;
            coff += ((Lk * ci.T + 15) / 16) * 16;           // class blocks start on 128-byte lines
            lutoff += ci.T;
            c->Tmax = std::max(c->Tmax, ci.T);
        }
        L += Lk;
    }
    if (L == 0) { delete c; return fail(BLU_ERR_ARG, "no groups"); }
    if (L > 0 && !groups_flat) { delete c; return fail(BLU_ERR_ARG, "null groups"); }
    c->L = L; c->gidx_len = ioff; c->cinv_len = coff; c->lo = 0; c->hi = L;
    c->Lpad = ((L + BLU_HT - 1) / BLU_HT) * BLU_HT + BLU_HT;
    c->ldH = ((L + 15) / 16) * 16;

    // host-side tables: byte ids, masks, (j,l) look-up table
    std::vector<uint8_t> gidx((size_t)ioff);
    std::vector<unsigned> gmask((size_t)L);
    std::vector<uint16_t> lut((size_t)std::max(lutoff, 1));
    {
        const int64_t *g = groups_flat;
        for (const BluClass &ci : c->cls) {
            for (long long i = 0; i < ci.Lk; ++i) {
                unsigned mask = 0; long long prev = -1;
                for (int j = 0; j < ci.k; ++j) {
                    long long v = g[i * ci.k + j];
                    if (v < 0 || v >= N || v <= prev) {
                        delete c;
                        return fail(BLU_ERR_ARG, "group %lld of class %d is not a strictly increasing subset of [0,%d)", i, ci.k, N);
                    }
                    prev = v; mask |= 1u << v;
                    gidx[(size_t)(ci.ioff + i * ci.k + j)] = (uint8_t)v;
                }
                gmask[(size_t)(ci.goff + i)] = mask;
            }
            g += ci.Lk * ci.k;
            int e = 0;
            for (int j = 0; j < ci.k; ++j)
                for (int l = j; l < ci.k; ++l) lut[(size_t)(ci.lutoff + e++)] = (uint16_t)((j << 8) | l);
        }
    }
    // Run table of the Phi kernel (blu_phi_chunk_run): the packed entries of a class sorted by l DESCENDING (entries
    // whose target moves most often first) and first-fit packed into half-warps of 16 entries with 16 DIFFERENT
    // shared-memory banks (packed position mod 16), so a half-warp's staged loads are conflict-free.  Two half-warps
    // = one 32-lane step.  Word: position | j << 10 | l << 15 | valid << 20.
    std::vector<unsigned> plut;
    for (BluClass &ci : c->cls) {
        struct Ent { int l, j, pos; };
        std::vector<Ent> ents;
        for (int j = 0; j < ci.k; ++j)
            for (int l = j; l < ci.k; ++l) ents.push_back({l, j, blu_pk(ci.k, j, l)});
        std::stable_sort(ents.begin(), ents.end(), [](const Ent &a, const Ent &b) { return a.l > b.l; });
        std::vector<std::vector<Ent>> halves;
        std::vector<unsigned> used;
        for (const Ent &en : ents) {
            const unsigned bit = 1u << (en.pos & 15);
            size_t h = 0;
            while (h < halves.size() && (used[h] & bit)) ++h;
            if (h == halves.size()) { halves.emplace_back(); used.push_back(0u); }
            halves[h].push_back(en); used[h] |= bit;
        }
        ci.plutoff = (int)plut.size();
        ci.psteps = (int)(halves.size() + 1) / 2;
        plut.resize(plut.size() + (size_t)ci.psteps * 32, 0u);
        for (size_t h = 0; h < halves.size(); ++h)
            for (size_t q = 0; q < halves[h].size(); ++q) {
                const Ent &en = halves[h][q];
                plut[(size_t)ci.plutoff + (h / 2) * 32 + (h & 1) * 16 + q] = (unsigned)en.pos | ((unsigned)en.j << 10) | ((unsigned)en.l << 15) | (1u << 20);
            }
    }
    if (plut.empty()) plut.push_back(0u);
    c->inv_set.assign(c->cls.size(), 0);

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->nsm = prop.multiProcessorCount;
#define CTX_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int code_ = fail(e_ == cudaErrorMemoryAllocation ? BLU_ERR_NOMEM : BLU_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); blu_ctx_destroy(c); return code_; } } while (0)
    CTX_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto &e : c->ev) CTX_TRY(cudaEventCreate(&e));
    CTX_TRY(cudaMalloc(&c->d_cls, sizeof(BluClass) * c->cls.size()));
    CTX_TRY(cudaMalloc(&c->d_gidx, std::max<long long>(ioff, 1) + 32));                 // + slack for 16-byte rounded bulk copies
    CTX_TRY(cudaMalloc(&c->d_gmask, sizeof(unsigned) * L));
    CTX_TRY(cudaMalloc(&c->d_lut, sizeof(uint16_t) * lut.size()));
    CTX_TRY(cudaMalloc(&c->d_plut, sizeof(unsigned) * plut.size()));
    CTX_TRY(cudaMalloc(&c->d_cinv, sizeof(double) * (std::max<long long>(coff, 1) + 16)));   // + slack for 16-byte rounded bulk copies
    CTX_TRY(cudaMemsetAsync(c->d_cinv, 0, sizeof(double) * (std::max<long long>(coff, 1) + 16), c->stream));
    CTX_TRY(cudaMemcpyAsync(c->d_cls, c->cls.data(), sizeof(BluClass) * c->cls.size(), cudaMemcpyHostToDevice, c->stream));
    CTX_TRY(cudaMemcpyAsync(c->d_gidx, gidx.data(), gidx.size(), cudaMemcpyHostToDevice, c->stream));
    CTX_TRY(cudaMemcpyAsync(c->d_gmask, gmask.data(), sizeof(unsigned) * L, cudaMemcpyHostToDevice, c->stream));
    CTX_TRY(cudaMemcpyAsync(c->d_lut, lut.data(), sizeof(uint16_t) * lut.size(), cudaMemcpyHostToDevice, c->stream));

    CTX_TRY(cudaMemcpyAsync(c->d_plut, plut.data(), sizeof(unsigned) * plut.size(), cudaMemcpyHostToDevice, c->stream));
    c->lutlen = (int)lut.size();
    c->plutlen = (int)plut.size();

    const size_t NN = (size_t)N * N;
    CTX_TRY(cudaMalloc(&c->d_C, sizeof(double) * NN));
    CTX_TRY(cudaMalloc(&c->d_m, sizeof(double) * L));
    CTX_TRY(cudaMalloc(&c->d_phi, sizeof(double) * (NN + 40)));   // + SUM-reducible support / non-tiny indicators
    CTX_TRY(cudaMalloc(&c->d_pinv, sizeof(double) * NN));
    CTX_TRY(cudaMalloc(&c->d_x, sizeof(double) * BLU_MAX_MODELS));
    CTX_TRY(cudaMalloc(&c->d_S, sizeof(double) * NN));
    CTX_TRY(cudaMalloc(&c->d_grad, sizeof(double) * L));
    CTX_TRY(cudaMalloc(&c->d_hdr, sizeof(BluEvalHeader)));
    CTX_TRY(cudaMemsetAsync(c->d_hdr, 0, sizeof(BluEvalHeader), c->stream));
    CTX_TRY(cudaHostAlloc(&c->h_hdr, sizeof(BluEvalHeader), cudaHostAllocDefault));
    memset(c->h_hdr, 0, sizeof(BluEvalHeader));
    // opt in to large dynamic shared memory where needed
    CTX_TRY(cudaFuncSetAttribute(blu_phi_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BLU_FIN_SEG * 1024 * 8));
    {
        const int ncl = (int)c->cls.size();
        CTX_TRY(cudaFuncSetAttribute(blu_phi_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     220 * 1024));
        CTX_TRY(cudaFuncSetAttribute(blu_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)blu_stream_smem_bytes(BLU_CHUNK_DOUBLES + 4, 0, 32, 6000)));
        CTX_TRY(cudaFuncSetAttribute(blu_gradu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)blu_stream_smem_bytes(BLU_CHUNK_DOUBLES + 4, 1024, 32, 6000)));
        CTX_TRY(cudaFuncSetAttribute(blu_ysum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)blu_stream_smem_bytes(BLU_CHUNK_DOUBLES + 4, BLU_STREAM_WARPS * 32, 32, 6000)));
        (void)ncl;
    }
    CTX_TRY(cudaStreamSynchronize(c->stream));
#undef CTX_TRY
    {
        int rc2 = build_chunks(c);
        if (rc2) { blu_ctx_destroy(c); return rc2; }
    }
    *out = c;
    return BLU_OK;
}

static int ensure_soa(blu_ctx *c);

// A second evaluation LANE on the same problem: the clone shares the parent's read-only data (group tables,
// packed inverses in both layouts, work lists of the current slice) and owns only the small per-evaluation state
// (stream, partial tiles, Phi, pinv, x, gradient, status block, exchange inbox).  Two lanes on two streams let the
// serial tail of one evaluation (fold, peer exchange, N x N inverse: one CTA) overlap the streaming kernels of the
// next -- independent evaluations of a sweep, or of a solver that evaluates several trial points.
// The parent must have its inverses set; it must outlive the clone and keep its slice while the clone lives.
extern "C" int blu_ctx_clone(blu_ctx *p, blu_ctx **out)
{
    if (!out) return fail(BLU_ERR_ARG, "null out");
    *out = nullptr;
    int rc = use(p);
    if (rc) return rc;
    if (!p->have_inv) return fail(BLU_ERR_STATE, "clone: the parent's inverses are not set");
    if (p->hi - p->lo >= (long long)p->nsm * 2 * 32 && p->use_soa) { if ((rc = ensure_soa(p))) return rc; }
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    blu_ctx *c = new (std::nothrow) blu_ctx(*p);      // scalars, host tables and the shared device pointers
    if (!c) return fail(BLU_ERR_NOMEM, "host allocation failed");
    c->is_clone = true;
    c->stream = nullptr;
    for (auto &e : c->ev) e = nullptr;
    c->evlog.clear(); c->evlog_n = 0; c->panel_ev.clear(); c->graphs.clear(); c->ipc_opened.clear();
    c->d_m = c->d_part = c->d_phi = c->d_pinv = c->d_x = c->d_S = c->d_grad = c->d_U = c->d_V = c->d_H = nullptr;
    c->d_hdr = nullptr; c->h_hdr = nullptr; c->h_m = c->h_grad = nullptr; c->pend_hess = nullptr; c->pending = false;
    c->d_xchg = nullptr; c->peers = BluPeers{}; c->d_Sop = c->d_hvpart = c->d_hvp = c->d_hvout = nullptr;
    for (int q = 0; q < 4; ++q) { c->kkt_ws[q] = nullptr; c->kkt_cap[q] = 0; }
    c->H_rows = 0; c->uv_ready = c->v_ready = false; c->capturing = false; c->d_grad_out = nullptr; c->timed = false;
#define CL_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int code_ = fail(e_ == cudaErrorMemoryAllocation ? BLU_ERR_NOMEM : BLU_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); blu_ctx_destroy(c); return code_; } } while (0)
    const size_t NN = (size_t)c->N * c->N;
    CL_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto &e : c->ev) CL_TRY(cudaEventCreate(&e));
    CL_TRY(cudaMalloc(&c->d_m, sizeof(double) * c->L));
    CL_TRY(cudaMalloc(&c->d_part, sizeof(double) * NN * (c->part_rows + BLU_PHI_MAXGROUPS + 1)));
    CL_TRY(cudaMalloc(&c->d_phi, sizeof(double) * (NN + 40)));
    CL_TRY(cudaMalloc(&c->d_pinv, sizeof(double) * NN));
    CL_TRY(cudaMalloc(&c->d_x, sizeof(double) * BLU_MAX_MODELS));
    CL_TRY(cudaMalloc(&c->d_S, sizeof(double) * NN));
    CL_TRY(cudaMalloc(&c->d_grad, sizeof(double) * c->L));
    CL_TRY(cudaMalloc(&c->d_hdr, sizeof(BluEvalHeader)));
    CL_TRY(cudaMemsetAsync(c->d_hdr, 0, sizeof(BluEvalHeader), c->stream));
    CL_TRY(cudaHostAlloc(&c->h_hdr, sizeof(BluEvalHeader), cudaHostAllocDefault));
    memset(c->h_hdr, 0, sizeof(BluEvalHeader));
    CL_TRY(cudaStreamSynchronize(c->stream));
#undef CL_TRY
    *out = c;
    return BLU_OK;
}

// --------------------------------------------------------------------------------------------
// kernel (1): inverses
// --------------------------------------------------------------------------------------------
extern "C" int blu_ctx_set_covariance(blu_ctx *c, const double *C, double pivot_rtol, int64_t *n_fallback)
{
    int rc = use(c);
    if (rc) return rc;
    if (!C) return fail(BLU_ERR_ARG, "null covariance");
    if (c->is_clone) return fail(BLU_ERR_STATE, "a clone shares its parent's inverses: set them on the parent");
    if (!(pivot_rtol >= 0.0)) pivot_rtol = 1e-10;
    const size_t NN = (size_t)c->N * c->N;
    CUDA_TRY(cudaMemcpyAsync(c->d_C, C, sizeof(double) * NN, cudaMemcpyHostToDevice, c->stream));
    struct DevBuf {                                  // freed on every return path
        void *p = nullptr;
        ~DevBuf() { if (p) cudaFree(p); }
    } flagbuf, todobuf;
    CUDA_TRY(cudaMalloc(&flagbuf.p, (size_t)c->L));
    unsigned char *d_flag = (unsigned char *)flagbuf.p;
    CUDA_TRY(cudaMemsetAsync(d_flag, 0, (size_t)c->L, c->stream));
    for (const BluClass &ci : c->cls) {
        blu_launch_invert_class(ci.k, c->nsm, c->stream, c->d_C, c->N, c->d_gidx + ci.ioff, ci.Lk, c->d_cinv + ci.coff, d_flag + ci.goff, pivot_rtol);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(BLU_ERR_CUDA, "invert kernel (k=%d): %s", ci.k, cudaGetErrorString(e));
        c->launches++;
    }
    std::vector<unsigned char> flag((size_t)c->L);
    CUDA_TRY(cudaMemcpyAsync(flag.data(), d_flag, (size_t)c->L, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // slow path for flagged groups
    long long nbad = 0;
    for (const BluClass &ci : c->cls) {
        std::vector<long long> todo;
        for (long long i = 0; i < ci.Lk; ++i) if (flag[(size_t)(ci.goff + i)]) todo.push_back(i);
        if (todo.empty()) continue;
        nbad += (long long)todo.size();
        if (todobuf.p) { cudaFree(todobuf.p); todobuf.p = nullptr; }
        CUDA_TRY(cudaMalloc(&todobuf.p, sizeof(long long) * todo.size()));
        long long *d_todo = (long long *)todobuf.p;
        CUDA_TRY(cudaMemcpyAsync(d_todo, todo.data(), sizeof(long long) * todo.size(), cudaMemcpyHostToDevice, c->stream));
        blu_launch_pinv_groups((unsigned)todo.size(), c->stream, c->d_C, c->N, ci.k, c->d_gidx + ci.ioff, d_todo, c->d_cinv + ci.coff, 1.0e-15);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return fail(BLU_ERR_CUDA, "pinv kernel (k=%d): %s", ci.k, cudaGetErrorString(e));
        c->launches++;
    }
    if (n_fallback) *n_fallback = nbad;
    std::fill(c->inv_set.begin(), c->inv_set.end(), 1);
    c->have_inv = true;
    c->soa_valid = false;
    return BLU_OK;
}

extern "C" int blu_ctx_set_invcovs(blu_ctx *c, int k, const double *invcovs_k)
{
    int rc = use(c);
    if (rc) return rc;
    if (k < 1 || k > c->K) return fail(BLU_ERR_ARG, "class k=%d outside [1,%d]", k, c->K);
    if (c->is_clone) return fail(BLU_ERR_STATE, "a clone shares its parent's inverses: set them on the parent");
    const int idx = c->cls_of_k[k];
    if (idx < 0) return BLU_OK;                       // empty class: nothing to ingest (sap.py:79)
    if (!invcovs_k) return fail(BLU_ERR_ARG, "null invcovs");
    const BluClass &ci = c->cls[idx];
    const size_t n = (size_t)ci.Lk * k * k;
    double *d_full = nullptr;
    CUDA_TRY(cudaMalloc(&d_full, sizeof(double) * n));
    cudaError_t e = cudaMemcpyAsync(d_full, invcovs_k, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        const long long total = ci.Lk * ci.T;
        const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)c->nsm * 8));
        blu_launch_pack_invcovs(grid, c->stream, d_full, k, ci.Lk, c->d_cinv + ci.coff);
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_full);
    if (e != cudaSuccess) return fail(BLU_ERR_CUDA, "set_invcovs: %s", cudaGetErrorString(e));
    c->inv_set[idx] = 1;
    c->soa_valid = false;
    c->have_inv = std::all_of(c->inv_set.begin(), c->inv_set.end(), [](char v) { return v != 0; });
    return BLU_OK;
}

extern "C" int blu_ctx_get_invcovs(blu_ctx *c, int k, double *invcovs_k)
{
    int rc = use(c);
    if (rc) return rc;
    if (k < 1 || k > c->K) return fail(BLU_ERR_ARG, "class k=%d outside [1,%d]", k, c->K);
    const int idx = c->cls_of_k[k];
    if (idx < 0) return BLU_OK;
    if (!c->inv_set[idx]) return fail(BLU_ERR_STATE, "inverses of class %d not set", k);
    if (!invcovs_k) return fail(BLU_ERR_ARG, "null invcovs");
    const BluClass &ci = c->cls[idx];
    const size_t n = (size_t)ci.Lk * k * k;
    double *d_full = nullptr;
    CUDA_TRY(cudaMalloc(&d_full, sizeof(double) * n));
    const int grid = (int)std::max<long long>(1, std::min<long long>(((long long)n + 255) / 256, (long long)c->nsm * 8));
    blu_launch_unpack_invcovs(grid, c->stream, c->d_cinv + ci.coff, k, ci.Lk, d_full);
    cudaError_t e = cudaGetLastError();
    c->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(invcovs_k, d_full, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_full);
    if (e != cudaSuccess) return fail(BLU_ERR_CUDA, "get_invcovs: %s", cudaGetErrorString(e));
    return BLU_OK;
}

extern "C" int blu_ctx_assemble_psi(blu_ctx *c, double *psi)
{
    int rc = use(c);
    if (rc) return rc;
    if (!c->have_inv) return fail(BLU_ERR_STATE, "inverses not set");
    if (!psi) return fail(BLU_ERR_ARG, "null psi");
    const long long NN = (long long)c->N * c->N;
    // column panels of at most ~256 MB so the device footprint stays bounded at large L
    const long long panel = std::max<long long>(1, std::min<long long>(c->L, (256ll << 20) / (8 * NN)));
    double *d_psi = nullptr;
    CUDA_TRY(cudaMalloc(&d_psi, sizeof(double) * NN * panel));
    cudaError_t e = cudaSuccess;
    for (long long col0 = 0; col0 < c->L && e == cudaSuccess; col0 += panel) {
        const long long nc = std::min(panel, c->L - col0);
        e = cudaMemsetAsync(d_psi, 0, sizeof(double) * NN * nc, c->stream);
        if (e != cudaSuccess) break;
        blu_psi_kernel<<<c->nsm * 4, 256, 0, c->stream>>>(c->d_cls, (int)c->cls.size(), c->N, col0, nc, c->d_gidx, c->d_cinv, d_psi);
        e = cudaGetLastError();
        c->launches++;
        if (e != cudaSuccess) break;
        // (NN, nc) panel -> columns [col0, col0+nc) of the (NN, L) host matrix
        e = cudaMemcpy2DAsync(psi + col0, sizeof(double) * c->L, d_psi, sizeof(double) * nc, sizeof(double) * nc, NN,
                              cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    }
    cudaFree(d_psi);
    if (e != cudaSuccess) return fail(BLU_ERR_CUDA, "assemble_psi: %s", cudaGetErrorString(e));
    return BLU_OK;
}

// --------------------------------------------------------------------------------------------
// evaluation pipeline
// --------------------------------------------------------------------------------------------
static int launch_phi(blu_ctx *c, const double *d_m, double delta, int mode)
{
    // ONE launch: the streaming pass over the packed inverses, the in-kernel two-level reduction of the CTA
    // tiles by the last CTAs to arrive, and the finish step (mode 0/1/2, or 3 = NVLink peer exchange) in that
    // same last CTA (blu_phi.cuh).
    const size_t smem = std::max(blu_phi_smem_bytes(c->phi_ns, c->phi_sd, c->idsd, c->N, (int)c->cls.size(), c->phi_warps), (size_t)BLU_FIN_SCRATCH_BYTES);
    blu_phi_partial_kernel<<<c->grid_phi, c->phi_warps * 32, smem, c->stream>>>(
        c->d_cls, (int)c->cls.size(), c->N, c->d_wchunks, c->d_wstart, c->phi_ns, c->phi_sd, c->idsd, c->d_cinv, c->d_gidx, c->d_lut, c->d_plut,
        c->d_gmask, d_m, c->d_part, c->d_hdr,
        mode, delta, c->d_phi, c->d_pinv, c->d_x, c->d_S, c->peers);
    KERNEL_CHECK(c);
    return BLU_OK;
}

static int ensure_uv(blu_ctx *c)
{
    if (c->d_U) return BLU_OK;
    const size_t n = (size_t)c->Lpad * c->NP;
    CUDA_TRY(cudaMalloc(&c->d_U, sizeof(double) * n));
    CUDA_TRY(cudaMalloc(&c->d_V, sizeof(double) * n));
    CUDA_TRY(cudaMemsetAsync(c->d_U, 0, sizeof(double) * n, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d_V, 0, sizeof(double) * n, c->stream));
    CUDA_TRY(cudaMalloc(&c->d_Sop, sizeof(double) * (size_t)c->N * c->N));
    return BLU_OK;
}

// SoA-tile copy of the inverses + tile work list of the owned slice (lazy; see blu_soa.cuh).
static int ensure_soa(blu_ctx *c)
{
    if (!c->d_soa) {
        c->soff.clear();
        long long off = 0;
        for (const BluClass &ci : c->cls) { c->soff.push_back(off); off += ((ci.Lk + 31) / 32) * 32 * ci.T; }
        c->soa_len = off;
        CUDA_TRY(cudaMalloc(&c->d_soa, sizeof(double) * std::max<long long>(off, 1)));
        CUDA_TRY(cudaMalloc(&c->d_soff, sizeof(long long) * c->cls.size()));
        CUDA_TRY(cudaMemcpyAsync(c->d_soff, c->soff.data(), sizeof(long long) * c->cls.size(), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    if (!c->soa_valid) {
        for (size_t ic = 0; ic < c->cls.size(); ++ic) {
            const BluClass &ci = c->cls[ic];
            const long long total = ((ci.Lk + 31) / 32) * 32 * ci.T;
            const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)c->nsm * 16));
            CUDA_TRY(blu_launch_soa_build(grid, c->stream, c->d_cinv + ci.coff, ci.Lk, ci.T, c->d_soa + c->soff[ic]));
            c->launches++;
        }
        c->soa_valid = true;
    }
    if (!c->tiles_valid) {
        std::vector<BluTile> tl;
        for (size_t ic = 0; ic < c->cls.size(); ++ic) {
            const BluClass &ci = c->cls[ic];
            const long long i0 = std::max<long long>(c->lo - ci.goff, 0), i1 = std::min<long long>(c->hi - ci.goff, ci.Lk);
            if (i1 <= i0) continue;
            for (long long t = i0 / 32; t <= (i1 - 1) / 32; ++t) {
                BluTile b; b.cls = (int)ic; b.nsub = (ci.T + BLU_SOA_E - 1) / BLU_SOA_E; b.t = t;
                tl.push_back(b);
            }
        }
        if (c->d_tiles) { CUDA_TRY(cudaStreamSynchronize(c->stream)); CUDA_TRY(cudaFree(c->d_tiles)); c->d_tiles = nullptr; }
        c->ntiles = (int)tl.size();
        CUDA_TRY(cudaMalloc(&c->d_tiles, sizeof(BluTile) * std::max<size_t>(tl.size(), 1)));
        if (!tl.empty()) CUDA_TRY(cudaMemcpyAsync(c->d_tiles, tl.data(), sizeof(BluTile) * tl.size(), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->grid_soa = (int)std::max<long long>(1, std::min<long long>(((long long)c->ntiles + BLU_SOA_WARPS - 1) / BLU_SOA_WARPS, (long long)c->nsm * 2));
        c->tiles_valid = true;
    }
    return BLU_OK;
}

// V rows of the owned slice from its U rows (V = U S, S = 2 pinv(Phi)): needed by the dense Hessian
// kernel and the sharded row panels, not by the operator.
static int launch_v_from_u(blu_ctx *c)
{
    const long long total = (c->hi - c->lo) * c->NP;
    if (total <= 0) return BLU_OK;
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)c->nsm * 8));
    CUDA_TRY(blu_launch_v_from_u(grid, c->stream, c->d_U, c->d_S, c->N, c->NP, c->lo, c->hi, c->d_V));
    c->launches++;
    return BLU_OK;
}

// uv: 0 = gradient only, 1 = gradient + U and V rows, 2 = gradient + U rows only (Hessian operator).
static int launch_grad(blu_ctx *c, int uv)
{
    const bool want_uv = uv != 0;
    double *gout = c->d_grad_out ? c->d_grad_out : c->d_grad;
    if (want_uv) {
        int rc0 = ensure_uv(c);
        if (rc0) return rc0;
        c->uv_ready = true; c->v_ready = (uv == 1);
        // The operator H p = U S U^T p outlives this evaluation (a trust-region solver keeps multiplying by the
        // Hessian of the last ACCEPTED iterate while it evaluates trial points): it gets its own copy of S.
        CUDA_TRY(cudaMemcpyAsync(c->d_Sop, c->d_S, sizeof(double) * (size_t)c->N * c->N, cudaMemcpyDeviceToDevice, c->stream));
    }
    // One group per lane needs at least a couple of 32-group tiles per SM to fill the machine; smaller
    // problems (latency-bound anyway) keep the entry-per-lane kernels, which spread a group over a warp.
    if (c->use_soa && c->hi - c->lo >= (long long)c->nsm * 2 * 32) {
        int rc = ensure_soa(c);
        if (rc) return rc;
        if (want_uv && (rc = ensure_uv(c))) return rc;
        CUDA_TRY(blu_launch_grad_soa(want_uv, c->grid_soa, c->stream, c->d_cls, (int)c->cls.size(), c->N, c->NP, c->K, c->d_tiles, c->ntiles,
                                     c->d_soa, c->d_soff, c->d_gmask, c->d_x, c->lo, c->hi, gout, want_uv ? c->d_U : nullptr));
        c->launches++;
        return uv == 1 ? launch_v_from_u(c) : BLU_OK;
    }
    if (!want_uv) {
        blu_grad_kernel<<<c->grid_grad, BLU_STREAM_WARPS * 32, blu_stream_smem_bytes(c->sd, 0, (int)c->cls.size(), c->lutlen), c->stream>>>(
            c->d_cls, (int)c->cls.size(), c->N, c->d_chunks, c->nchunks, c->sd, c->d_cinv, c->d_lut, c->lutlen, c->d_gmask, c->d_x, gout);
        KERNEL_CHECK(c);
        return BLU_OK;
    }
    int rc = ensure_uv(c);
    if (rc) return rc;
    c->v_ready = true;                  // the entry-per-lane kernel writes both factors
    blu_gradu_kernel<<<c->grid_grad, BLU_STREAM_WARPS * 32, blu_stream_smem_bytes(c->sd, c->N * c->N, (int)c->cls.size(), c->lutlen), c->stream>>>(
        c->d_cls, (int)c->cls.size(), c->N, c->NP, c->d_chunks, c->nchunks, c->sd, c->d_cinv, c->d_lut, c->lutlen, c->d_gmask, c->d_x, c->d_S,
        gout, c->d_U, c->d_V);
    KERNEL_CHECK(c);
    return BLU_OK;
}

template <int NCH>
static void launch_hess_t(blu_ctx *c, bool sym, const double *Ua, long long Lrows, double *H)
{
    const int which = sym ? (c->hess_onebuf ? 2 : 1) : 0;
    if (!c->hess_attr_done[which]) {          // per context: function attributes are per device
        if (which == 2) cudaFuncSetAttribute(blu_hess_kernel<NCH, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLU_HESS_SMEM1);
        else if (which == 1) cudaFuncSetAttribute(blu_hess_kernel<NCH, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLU_HESS_SMEM2);
        else cudaFuncSetAttribute(blu_hess_kernel<NCH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLU_HESS_SMEM1);
        c->hess_attr_done[which] = true;
    }
    const int nTc = (int)((c->L + BLU_HT - 1) / BLU_HT);
    if (sym) {
        const int nB = (nTc + BLU_HSB - 1) / BLU_HSB;
        dim3 grid(BLU_HSB * BLU_HSB, (unsigned)((long long)nB * (nB + 1) / 2));
        if (c->hess_onebuf) blu_hess_kernel<NCH, true, true><<<grid, 128, BLU_HESS_SMEM1, c->stream>>>(Ua, c->d_V, Lrows, c->L, c->ldH, H, nTc, 0);
        else blu_hess_kernel<NCH, true, false><<<grid, 128, BLU_HESS_SMEM2, c->stream>>>(Ua, c->d_V, Lrows, c->L, c->ldH, H, nTc, 0);
    } else {
        const int nTr = (int)((Lrows + BLU_HT - 1) / BLU_HT);
        dim3 grid((unsigned)nTc, (unsigned)nTr);
        // row panels stage only the normal orientation: half the shared memory, more CTAs per SM
        blu_hess_kernel<NCH, false><<<grid, 128, BLU_HESS_SMEM1, c->stream>>>(Ua, c->d_V, Lrows, c->L, c->ldH, H, nTc, 0);
    }
}

static int launch_hess(blu_ctx *c, bool sym, long long rlo = 0, long long rhi = 0)
{
    const long long rows = sym ? c->L : (rhi - rlo);
    if (rows <= 0) return BLU_OK;
    if (!c->d_H || c->H_rows < rows) {
        if (c->d_H) { CUDA_TRY(cudaStreamSynchronize(c->stream)); CUDA_TRY(cudaFree(c->d_H)); c->d_H = nullptr; }
        CUDA_TRY(cudaMalloc(&c->d_H, sizeof(double) * (size_t)rows * (size_t)c->ldH));
        c->H_rows = rows;
    }
    const double *Ua = c->d_U + (sym ? 0 : rlo * c->NP);
    switch (c->NCH) {
        case 1: launch_hess_t<1>(c, sym, Ua, rows, c->d_H); break;
        case 2: launch_hess_t<2>(c, sym, Ua, rows, c->d_H); break;
        case 3: launch_hess_t<3>(c, sym, Ua, rows, c->d_H); break;
        case 4: launch_hess_t<4>(c, sym, Ua, rows, c->d_H); break;
        case 5: launch_hess_t<5>(c, sym, Ua, rows, c->d_H); break;
        case 6: launch_hess_t<6>(c, sym, Ua, rows, c->d_H); break;
        case 7: launch_hess_t<7>(c, sym, Ua, rows, c->d_H); break;
        default: launch_hess_t<8>(c, sym, Ua, rows, c->d_H); break;
    }
    KERNEL_CHECK(c);
    return BLU_OK;
}

extern "C" int blu_eval_device(blu_ctx *c, const double *d_m, double delta, int want_grad, int want_hess)
{
    int rc = use(c);
    if (rc) return rc;
    if (!c->have_inv) return fail(BLU_ERR_STATE, "inverses not set: call blu_ctx_set_covariance / blu_ctx_set_invcovs first");
    if (c->lo != 0 || c->hi != c->L) return fail(BLU_ERR_STATE, "context owns a slice: use the blu_shard_* calls");
    if (!d_m) d_m = c->d_m;
    c->launches = 0;
    if (c->capturing) {                 // inside a graph: kernels only (phase events belong to eager runs)
        if ((rc = launch_phi(c, d_m, delta, 1))) return rc;
        if (want_grad || want_hess) { if ((rc = launch_grad(c, want_hess == 0 ? 0 : (want_hess == 3 ? 2 : 1)))) return rc; }
        if (want_hess == 1) { if ((rc = launch_hess(c, true))) return rc; }
        return BLU_OK;
    }
    cudaEvent_t *ev = c->ev;
    if ((size_t)(c->evlog_n + 1) * 4 <= c->evlog.size()) { ev = c->evlog.data() + (size_t)c->evlog_n * 4; c->evlog_n++; }
    CUDA_TRY(cudaEventRecord(ev[0], c->stream));
    rc = launch_phi(c, d_m, delta, 1);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ev[1], c->stream));
    if (want_grad || want_hess) { rc = launch_grad(c, want_hess == 0 ? 0 : (want_hess == 3 ? 2 : 1)); if (rc) return rc; }
    CUDA_TRY(cudaEventRecord(ev[2], c->stream));
    if (want_hess == 1) { rc = launch_hess(c, true); if (rc) return rc; }      // 2: U,V only (row panels follow via blu_shard_hess)
    CUDA_TRY(cudaEventRecord(ev[3], c->stream));
    c->timed = (ev == c->ev);
    return BLU_OK;
}

// Per-evaluation event log: after blu_ctx_timing_log(ctx, cap) the next `cap` calls of
// blu_eval_device record their four phase events into a log instead of the single "last" slot, so
// a benchmark can time every kernel of a long run without a host sync inside the timed region.
extern "C" int blu_ctx_timing_log(blu_ctx *c, int capacity)
{
    int rc = use(c);
    if (rc) return rc;
    if (capacity < 0) return fail(BLU_ERR_ARG, "negative capacity");
    for (auto &e : c->evlog) if (e) cudaEventDestroy(e);
    c->evlog.assign((size_t)capacity * 4, nullptr);
    for (auto &e : c->evlog) CUDA_TRY(cudaEventCreate(&e));
    c->evlog_n = 0;
    return BLU_OK;
}

// ms: (n,4) floats  [phi+pinv, grad/U, Hessian, total] per logged evaluation; *n = evaluations logged.
extern "C" int blu_ctx_timing_read(blu_ctx *c, float *ms, int *n)
{
    int rc = use(c);
    if (rc) return rc;
    if (!ms || !n) return fail(BLU_ERR_ARG, "null argument");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < c->evlog_n; ++i) {
        cudaEvent_t *ev = c->evlog.data() + (size_t)i * 4;
        CUDA_TRY(cudaEventElapsedTime(&ms[4 * i + 0], ev[0], ev[1]));
        CUDA_TRY(cudaEventElapsedTime(&ms[4 * i + 1], ev[1], ev[2]));
        CUDA_TRY(cudaEventElapsedTime(&ms[4 * i + 2], ev[2], ev[3]));
        CUDA_TRY(cudaEventElapsedTime(&ms[4 * i + 3], ev[0], ev[3]));
    }
    *n = c->evlog_n;
    return BLU_OK;
}

extern "C" int blu_ctx_sync(blu_ctx *c)
{
    int rc = use(c);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return BLU_OK;
}

extern "C" int blu_ctx_stream(blu_ctx *c, void **s)
{
    if (!c || !s) return fail(BLU_ERR_ARG, "null argument");
    *s = (void *)c->stream;
    return BLU_OK;
}

static int fetch_header(blu_ctx *c)
{
    CUDA_TRY(cudaMemcpyAsync(c->h_hdr, c->d_hdr, sizeof(BluEvalHeader), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return BLU_OK;
}

extern "C" int blu_ctx_last_result(blu_ctx *c, double *var, unsigned *flags)
{
    int rc = use(c);
    if (rc) return rc;
    rc = fetch_header(c);
    if (rc) return rc;
    if (var) *var = c->h_hdr->scal[0];
    if (flags) *flags = c->h_hdr->flags;
    return BLU_OK;
}

extern "C" int blu_ctx_last_timing(blu_ctx *c, float *ms)
{
    int rc = use(c);
    if (rc) return rc;
    if (!ms) return fail(BLU_ERR_ARG, "null ms");
    if (!c->timed) return fail(BLU_ERR_STATE, "no timed evaluation yet");
    CUDA_TRY(cudaEventSynchronize(c->ev[3]));
    CUDA_TRY(cudaEventElapsedTime(&ms[0], c->ev[0], c->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&ms[1], c->ev[1], c->ev[2]));
    CUDA_TRY(cudaEventElapsedTime(&ms[2], c->ev[2], c->ev[3]));
    CUDA_TRY(cudaEventElapsedTime(&ms[3], c->ev[0], c->ev[3]));
    return BLU_OK;
}

extern "C" int blu_ctx_last_launches(blu_ctx *c) { return c ? c->launches : 0; }

// Profiling aid: %globaltimer stamps (ns) the last CTA of the fused Phi kernel left at its milestones
// ([1] own stream done, [2] last of its group, [3] final fold starts, [4] rank sums complete, [5] peer exchange
// complete, [6] Phi ready, [7] pinv written, [9] done).  Synchronises.
extern "C" int blu_ctx_last_stamps(blu_ctx *c, unsigned long long *stamps16)
{
    int rc = use(c);
    if (rc) return rc;
    if (!stamps16) return fail(BLU_ERR_ARG, "null stamps");
    rc = fetch_header(c);
    if (rc) return rc;
    memcpy(stamps16, c->h_hdr->stamp, sizeof(unsigned long long) * 16);
    return BLU_OK;
}

#ifdef BLU_PHI_PROFILE
extern "C" int blu_prof_warp_read(long long *out, int nwarps)
{
    return cudaMemcpyFromSymbol(out, blu_prof_warp, sizeof(long long) * 4 * (size_t)nwarps) == cudaSuccess ? 0 : 1;
}
#endif

// Options: "soa" (default 1) -- gradient / U,V kernels on the group-interleaved copy of the inverses
// (blu_soa.cuh); 0 selects the entry-per-lane kernels of blu_grad.cuh on the group-major copy.
//          "sym_download" (default 1) -- copy only the upper block-triangle of the dense Hessian
// over PCIe and mirror it on the host with threads (streaming stores, blu_hostmirror.cpp) while later
// panels are in flight.  On the B200 box (16 host cores, PCIe 57 GB/s) one N = 15 evaluation end to
// end takes 97-125 ms instead of 152-170 ms; the limit is the host's memory bandwidth (4.3 GB of DMA
// writes + 8.6 GB of mirror traffic).  0 = one plain 2-D copy of all 8 L^2 bytes.
//          "mirror_threads" (default 0 = automatic: min(16, cores) / LOCAL_WORLD_SIZE, at least 2).
extern "C" int blu_ctx_set_option(blu_ctx *c, const char *name, int value)
{
    if (!c || !name) return fail(BLU_ERR_ARG, "null argument");
    if (!strcmp(name, "sym_download")) { c->sym_download = value != 0; return BLU_OK; }
    if (!strcmp(name, "soa")) { c->use_soa = value != 0; return BLU_OK; }
    if (!strcmp(name, "mirror_threads")) { c->mirror_threads = value; return BLU_OK; }
    if (!strcmp(name, "sym_full_rows_pct")) {          // a value >= 0 pins the share; -1 returns to the adaptive default
        if (value < 0) { c->sym_pct_auto = true; c->sym_full_rows_pct = 10; c->sym_calm = 0; }
        else { c->sym_pct_auto = false; c->sym_full_rows_pct = value; }
        return BLU_OK;
    }
    if (!strcmp(name, "hess_onebuf")) { c->hess_onebuf = value != 0; return BLU_OK; }
    if (!strcmp(name, "phi_stages")) {          // ring depth of the Phi kernel: 2 (4 KB chunks) or 4 (2 KB chunks)
        if (value != 2 && value != 4) return fail(BLU_ERR_ARG, "phi_stages must be 2 or 4");
        if (c->is_clone) return fail(BLU_ERR_STATE, "set phi_stages on the parent context");
        int rc = use(c);
        if (rc) return rc;
        c->phi_stages = value;
        return build_chunks(c);
    }
    return fail(BLU_ERR_ARG, "unknown option %s", name);
}

extern "C" int blu_ctx_get_option(blu_ctx *c, const char *name, int *value)
{
    if (!c || !name || !value) return fail(BLU_ERR_ARG, "null argument");
    if (!strcmp(name, "sym_download")) { *value = c->sym_download ? 1 : 0; return BLU_OK; }
    if (!strcmp(name, "soa")) { *value = c->use_soa ? 1 : 0; return BLU_OK; }
    if (!strcmp(name, "mirror_threads")) { *value = c->mirror_threads; return BLU_OK; }
    if (!strcmp(name, "sym_full_rows_pct")) { *value = c->sym_full_rows_pct; return BLU_OK; }   // the CURRENT share (adaptive or pinned)
    if (!strcmp(name, "hess_onebuf")) { *value = c->hess_onebuf ? 1 : 0; return BLU_OK; }
    if (!strcmp(name, "phi_stages")) { *value = c->phi_stages; return BLU_OK; }
    return fail(BLU_ERR_ARG, "unknown option %s", name);
}

extern "C" int blu_ctx_device_ptr(blu_ctx *c, int which, void **ptr, int64_t *nbytes)
{
    if (!c || !ptr) return fail(BLU_ERR_ARG, "null argument");
    const int64_t NN = (int64_t)c->N * c->N;
    void *p = nullptr; int64_t n = 0;
    switch (which) {
        case BLU_BUF_M: p = c->d_m; n = 8 * c->L; break;
        case BLU_BUF_PHI: p = c->d_phi; n = 8 * (NN + 40); break;     // + support / non-tiny indicators (sharded mode)
        case BLU_BUF_PINV: p = c->d_pinv; n = 8 * NN; break;
        case BLU_BUF_GRAD: p = c->d_grad; n = 8 * c->L; break;
        case BLU_BUF_U: { int rc = use(c); if (rc) return rc; rc = ensure_uv(c); if (rc) return rc; p = c->d_U; n = 8 * c->Lpad * c->NP; break; }
        case BLU_BUF_V: { int rc = use(c); if (rc) return rc; rc = ensure_uv(c); if (rc) return rc; p = c->d_V; n = 8 * c->Lpad * c->NP; break; }
        case BLU_BUF_HESS: p = c->d_H; n = 8 * c->H_rows * c->ldH; break;
        case BLU_BUF_CINV: p = c->d_cinv; n = 8 * c->cinv_len; break;
        case BLU_BUF_SCAL: p = c->d_hdr->scal; n = 64; break;
        default: return fail(BLU_ERR_ARG, "unknown buffer id %d", which);
    }
    *ptr = p;
    if (nbytes) *nbytes = n;
    return BLU_OK;
}

// ---- CUDA graphs ---------------------------------------------------------------------------
// Everything the device-resident calls (blu_eval_device, blu_shard_eval_fused, blu_hess_matvec_device,
// blu_ctx_save_result) enqueue between _begin and _end is recorded into a graph instead of running; replaying
// it costs one launch for the whole chain (small problems and sharded evaluations are launch-latency bound).
// Buffers must exist already: run the same calls once eagerly first (lazy allocations cannot be captured).
extern "C" int blu_ctx_graph_begin(blu_ctx *c)
{
    int rc = use(c);
    if (rc) return rc;
    if (c->capturing) return fail(BLU_ERR_STATE, "a graph is already being recorded on this context");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    c->capturing = true;
    return BLU_OK;
}

extern "C" int blu_ctx_graph_end(blu_ctx *c, int *graph_id)
{
    int rc = use(c);
    if (rc) return rc;
    if (!c->capturing) return fail(BLU_ERR_STATE, "no graph is being recorded");
    c->capturing = false;
    cudaGraph_t g = nullptr;
    CUDA_TRY(cudaStreamEndCapture(c->stream, &g));
    cudaGraphExec_t ge = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(BLU_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    c->graphs.push_back(ge);
    if (graph_id) *graph_id = (int)c->graphs.size() - 1;
    return BLU_OK;
}

extern "C" int blu_ctx_graph_launch(blu_ctx *c, int graph_id, int times)
{
    int rc = use(c);
    if (rc) return rc;
    if (graph_id < 0 || graph_id >= (int)c->graphs.size()) return fail(BLU_ERR_ARG, "unknown graph %d", graph_id);
    if (c->capturing) return fail(BLU_ERR_STATE, "cannot launch while recording");
    for (int t = 0; t < times; ++t) CUDA_TRY(cudaGraphLaunch(c->graphs[graph_id], c->stream));
    return BLU_OK;
}

// Keep the scalar results of the evaluation just enqueued (stream ordered, capturable): *d_var = variance,
// *d_flags = BLU_FLAG_* -- the context's own status block is overwritten by the next evaluation.
extern "C" int blu_ctx_save_result(blu_ctx *c, double *d_var, unsigned *d_flags)
{
    int rc = use(c);
    if (rc) return rc;
    if (d_var) CUDA_TRY(cudaMemcpyAsync(d_var, &c->d_hdr->scal[0], sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    if (d_flags) CUDA_TRY(cudaMemcpyAsync(d_flags, &c->d_hdr->flags, sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));
    return BLU_OK;
}

// Redirect the gradient of the following evaluations to d_grad (L doubles on the context's device; NULL: back to
// the context's own buffer): a batch of evaluations keeps every gradient without a copy.
extern "C" int blu_ctx_set_grad_output(blu_ctx *c, double *d_grad)
{
    if (!c) return fail(BLU_ERR_ARG, "null context");
    c->d_grad_out = d_grad;
    return BLU_OK;
}

// ---- host-pointer closures ----------------------------------------------------------------
static int upload_m(blu_ctx *c, const double *m)
{
    if (!m) return fail(BLU_ERR_ARG, "null m");
    CUDA_TRY(cudaMemcpyAsync(c->d_m, m, sizeof(double) * c->L, cudaMemcpyHostToDevice, c->stream));
    return BLU_OK;
}

extern "C" int blu_get_phi(blu_ctx *c, const double *m, double delta, double *phi)
{
    int rc = use(c);
    if (rc) return rc;
    if (!c->have_inv) return fail(BLU_ERR_STATE, "inverses not set");
    if (!phi) return fail(BLU_ERR_ARG, "null phi");
    if ((rc = upload_m(c, m))) return rc;
    c->launches = 0;
    if ((rc = launch_phi(c, c->d_m, delta, 0))) return rc;
    CUDA_TRY(cudaMemcpyAsync(phi, c->d_phi, sizeof(double) * c->N * c->N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return BLU_OK;
}

extern "C" int blu_variance_GH_begin(blu_ctx *c, const double *m, double delta, int want_grad, double *hess);
extern "C" int blu_variance_GH_end(blu_ctx *c, double *var, double *grad, unsigned *flags);

extern "C" int blu_variance(blu_ctx *c, const double *m, double delta, double *var, unsigned *flags)
{
    int rc = blu_variance_GH_begin(c, m, delta, 0, nullptr);
    if (rc) return rc;
    return blu_variance_GH_end(c, var, nullptr, flags);
}

// --------------------------------------------------------------------------------------------
// Dense Hessian to the host.  H is exactly symmetric, so only the block-upper triangle crosses
// PCIe (row panels, columns >= the panel's first row); host threads mirror each panel into the
// lower triangle (blocked 64 x 64 transposes) while the next panel is still in flight.  Halves the
// PCIe bytes of the one transfer that dominates the end-to-end evaluation (8 L^2 bytes at 57 GB/s).
// --------------------------------------------------------------------------------------------
static int download_hessian_symmetric(blu_ctx *c, double *hess)
{
    const long long L = c->L;
    const long long PH = 1024;                                   // panel height (rows)
    const int npan = (int)((L + PH - 1) / PH);
    while ((int)c->panel_ev.size() < npan) { cudaEvent_t e; CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->panel_ev.push_back(e); }
    // The host mirror (4.3 GB of reads + 4.3 GB of streaming stores next to the DMA's own 4.3 GB of writes) is the
    // slower of the two parties: the bottom panels -- whose lower-triangle part is the longest -- cross PCIe as FULL
    // rows, so the host mirrors only into the rows above them.  With the measured rates (DMA 57 GB/s, mirror ~39 GB/s
    // while the DMA runs) on the pool's hosts the optimum is flat between 0 and ~19 % (host memory bandwidth is shared by both): 10 % by default.
    const double dma_share = std::min(100, std::max(0, c->sym_full_rows_pct)) / 100.0;
    long long cs = (long long)std::floor((double)L * std::sqrt(1.0 - dma_share));
    cs = std::min<long long>(L, ((cs + PH - 1) / PH) * PH);       // first fully transferred row, panel aligned
    for (int p = 0; p < npan; ++p) {
        const long long r0 = p * PH, r1 = std::min<long long>(L, r0 + PH);
        const long long c0 = r0 >= cs ? 0 : r0;                   // full rows below the split, columns >= r0 above it
        CUDA_TRY(cudaMemcpy2DAsync(hess + r0 * L + c0, sizeof(double) * L, c->d_H + r0 * c->ldH + c0, sizeof(double) * c->ldH,
                                   sizeof(double) * (L - c0), (size_t)(r1 - r0), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaEventRecord(c->panel_ev[p], c->stream));
    }
    // host threads for the mirroring: all cores (at most 16), shared fairly when torchrun runs one rank per GPU
    int nthreads = c->mirror_threads;
    if (nthreads <= 0) {
        nthreads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        const char *lws = getenv("LOCAL_WORLD_SIZE");
        const int ranks = lws ? atoi(lws) : 1;
        if (ranks > 1) nthreads = std::max(2, nthreads / ranks);
    }
    std::vector<std::atomic<int>> ready(npan), next(npan);
    for (int p = 0; p < npan; ++p) { ready[p].store(0); next[p].store(0); }
    const long long CW = 512;                                     // columns per work item
    auto worker = [&]() {
        double *scratch = blu_host_mirror_scratch_alloc();           // owned by this worker for the whole download
        for (int p = 0; p < npan; ++p) {
            const long long r0 = p * PH, r1 = std::min<long long>(L, r0 + PH);
            const long long cend = std::min(L, cs);                   // rows >= cs arrive whole
            const int nitems = cend > r1 ? (int)((cend - r1 + CW - 1) / CW) : 0;
            if (nitems <= 0) continue;
            while (ready[p].load(std::memory_order_acquire) == 0) std::this_thread::yield();
            for (;;) {
                const int it = next[p].fetch_add(1);
                if (it >= nitems) break;
                const long long c0 = r1 + it * CW, c1 = std::min<long long>(cend, c0 + CW);
                blu_host_mirror_block(hess, L, r0, r1, c0, c1, scratch);
            }
        }
        blu_host_mirror_scratch_free(scratch);
        blu_host_store_fence();                                     // streaming stores visible before join
    };
    const bool dbg = getenv("BLU_DEBUG_TIMING") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    try {
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker);
    } catch (...) {                                               // thread creation failed: mirror with what we have (or inline)
    }
    const bool inline_mirror = pool.empty();
    cudaError_t err = cudaSuccess;
    for (int p = 0; p < npan; ++p) {
        if (err == cudaSuccess) err = cudaEventSynchronize(c->panel_ev[p]);
        ready[p].store(1, std::memory_order_release);              // on error too: never leave the workers spinning
        if (dbg && (p == 0 || p == npan / 2 || p == npan - 1))
            fprintf(stderr, "[blu] panel %d arrived at %.1f ms\n", p, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    }
    const auto t_arrived = std::chrono::steady_clock::now();      // the last panel is on the host
    if (inline_mirror) worker();                                  // no helper thread could be started
    for (auto &t : pool) t.join();
    const auto t_done = std::chrono::steady_clock::now();
    if (dbg) fprintf(stderr, "[blu] mirror done at %.1f ms (DMA share %d %%)\n", std::chrono::duration<double, std::milli>(t_done - t_start).count(), c->sym_full_rows_pct);
    if (c->sym_pct_auto && err == cudaSuccess && !inline_mirror && npan >= 8) {
        // Balance the two parties for THIS host: the download ends when the slower of DMA and mirror ends.  The mirror
        // lagging behind the last panel means the host threads are the bottleneck -> send more rows whole over PCIe;
        // finishing together with it (twice in a row) means they have slack -> mirror more, move fewer bytes.
        const double total = std::chrono::duration<double, std::milli>(t_done - t_start).count();
        const double lag = std::chrono::duration<double, std::milli>(t_done - t_arrived).count();
        if (lag > 0.04 * total) { c->sym_full_rows_pct = std::min(40, c->sym_full_rows_pct + 5); c->sym_calm = 0; }
        else if (lag < 0.01 * total) { if (++c->sym_calm >= 2 && c->sym_full_rows_pct > 0) { c->sym_full_rows_pct -= 5; c->sym_calm = 0; } }
        else c->sym_calm = 0;
    }
    if (err != cudaSuccess) return fail(BLU_ERR_CUDA, "Hessian download: %s", cudaGetErrorString(err));
    // the diagonal blocks arrived whole; the strictly-lower part of each diagonal block was copied too
    return BLU_OK;
}

// Split evaluation: _begin queues H2D of m (through pinned staging), the kernels and the D2H of the
// small results on the context's stream and returns at once; _end waits and hands the results over.
// Contexts have their own streams, so several outputs (MOSAP) or instances overlap on the device.
extern "C" int blu_variance_GH_begin(blu_ctx *c, const double *m, double delta, int want_grad, double *hess)
{
    int rc = use(c);
    if (rc) return rc;
    if (!m) return fail(BLU_ERR_ARG, "null m");
    if (c->pending) return fail(BLU_ERR_STATE, "an evaluation is already pending on this context");
    if (!c->h_m) {
        CUDA_TRY(cudaHostAlloc(&c->h_m, sizeof(double) * c->L, cudaHostAllocDefault));
        CUDA_TRY(cudaHostAlloc(&c->h_grad, sizeof(double) * c->L, cudaHostAllocDefault));
    }
    memcpy(c->h_m, m, sizeof(double) * c->L);
    CUDA_TRY(cudaMemcpyAsync(c->d_m, c->h_m, sizeof(double) * c->L, cudaMemcpyHostToDevice, c->stream));
    if ((rc = blu_eval_device(c, c->d_m, delta, want_grad || hess != nullptr, hess != nullptr))) return rc;
    // small results first; the Hessian D2H (the long pole) is queued behind them on the same stream
    CUDA_TRY(cudaMemcpyAsync(c->h_hdr, c->d_hdr, sizeof(BluEvalHeader), cudaMemcpyDeviceToHost, c->stream));
    if (want_grad || hess) CUDA_TRY(cudaMemcpyAsync(c->h_grad, c->d_grad, sizeof(double) * c->L, cudaMemcpyDeviceToHost, c->stream));
    if (hess && !(c->L >= 4096 && c->sym_download))
        CUDA_TRY(cudaMemcpy2DAsync(hess, sizeof(double) * c->L, c->d_H, sizeof(double) * c->ldH, sizeof(double) * c->L,
                                   (size_t)c->L, cudaMemcpyDeviceToHost, c->stream));
    c->pend_hess = hess;
    c->pending = true;
    return BLU_OK;
}

extern "C" int blu_variance_GH_end(blu_ctx *c, double *var, double *grad, unsigned *flags)
{
    int rc = use(c);
    if (rc) return rc;
    if (!c->pending) return fail(BLU_ERR_STATE, "no pending evaluation");
    c->pending = false;
    double *hess = c->pend_hess;
    c->pend_hess = nullptr;
    if (hess && c->L >= 4096 && c->sym_download) {
        if ((rc = download_hessian_symmetric(c, hess))) return rc;
    } else {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    const unsigned fl = c->h_hdr->flags;
    if (var) *var = c->h_hdr->scal[0];
    if (flags) *flags = fl;
    if (grad) {
        if (fl & BLU_FLAG_TINY) for (long long i = 0; i < c->L; ++i) grad[i] = std::numeric_limits<double>::infinity();
        else memcpy(grad, c->h_grad, sizeof(double) * c->L);
    }
    return BLU_OK;
}

extern "C" int blu_variance_GH(blu_ctx *c, const double *m, double delta, double *var, double *grad,
                               double *hess, unsigned *flags)
{
    if (!grad) return fail(BLU_ERR_ARG, "null grad");
    int rc = blu_variance_GH_begin(c, m, delta, 1, hess);
    if (rc) return rc;
    return blu_variance_GH_end(c, var, grad, flags);
}

// --------------------------------------------------------------------------------------------
// Hessian as an operator: H p = V (U^T p)  (blu_matvec.cuh).  The dense (L,L) matrix of
// misc.py:497-503 is never formed; U, V are the factors the gradient pass leaves in HBM.
// --------------------------------------------------------------------------------------------
static int ensure_hv(blu_ctx *c)
{
    if (c->d_hvpart) return BLU_OK;
    // CTA partials (32 doubles each) | folded t (32 doubles) | ticket counter
    CUDA_TRY(cudaMalloc(&c->d_hvpart, sizeof(double) * (32 * (size_t)c->nsm * 8 + 32 + 2)));
    CUDA_TRY(cudaMemsetAsync(c->d_hvpart, 0, sizeof(double) * (32 * (size_t)c->nsm * 8 + 32 + 2), c->stream));
    CUDA_TRY(cudaMalloc(&c->d_hvp, sizeof(double) * c->L));
    CUDA_TRY(cudaMalloc(&c->d_hvout, sizeof(double) * c->L));
    return BLU_OK;
}

// t (32 doubles at d_t) = sum over the owned rows of p_i u_i
// Second-generation kernels (blu_matvec.cuh): 8 x 16 bytes in flight per thread, two resident CTAs per SM, one wave;
// tools/lab/hv_lab.cu measured 58.7 us per product at 20 models against 89.3 us for the first generation.
#define BLU_HV2_UN 8
#define BLU_HV2_MINB 2
#define BLU_HVA2_WARPS 8
#define BLU_HVA2_MINB 4
static int launch_hv_reduce(blu_ctx *c, const double *d_p, double *d_t)
{
    const long long rows = c->hi - c->lo;
    const int rpp = BLU_HV_THREADS / (c->NP / 2);
    const int grid = (int)std::max<long long>(1, std::min<long long>((rows + (long long)rpp * BLU_HV2_UN - 1) / ((long long)rpp * BLU_HV2_UN), (long long)c->nsm * BLU_HV2_MINB));
    unsigned *ticket = reinterpret_cast<unsigned *>(c->d_hvpart + 32 * (size_t)c->nsm * 8 + 32);
    switch (c->NP) {
#define HV_CASE(P) case P: blu_hv_reduce_kernel<P, BLU_HV2_UN, BLU_HV2_MINB><<<grid, BLU_HV_THREADS, 0, c->stream>>>(c->d_U, d_p, c->lo, c->hi, c->d_hvpart, ticket, d_t); break;
        HV_CASE(4) HV_CASE(8) HV_CASE(12) HV_CASE(16) HV_CASE(20) HV_CASE(24) HV_CASE(28) HV_CASE(32)
#undef HV_CASE
    }
    KERNEL_CHECK(c);
    return BLU_OK;
}

static int launch_hv_apply(blu_ctx *c, const double *d_t, double *d_out)
{
    const long long rows = c->hi - c->lo;
    const long long nblk = (rows + 31) / 32;
    const int grid = (int)std::max<long long>(1, std::min<long long>((nblk + BLU_HVA2_WARPS - 1) / BLU_HVA2_WARPS, (long long)c->nsm * BLU_HVA2_MINB));
    switch (c->NP) {
#define HV_CASE(P) case P: blu_hv_apply_kernel<P, BLU_HVA2_WARPS, BLU_HVA2_MINB><<<grid, BLU_HVA2_WARPS * 32, 0, c->stream>>>(c->d_U, c->d_Sop, c->N, d_t, c->lo, c->hi, d_out, 1); break;
        HV_CASE(4) HV_CASE(8) HV_CASE(12) HV_CASE(16) HV_CASE(20) HV_CASE(24) HV_CASE(28) HV_CASE(32)
#undef HV_CASE
    }
    KERNEL_CHECK(c);
    return BLU_OK;
}

static int hv_ready(blu_ctx *c)
{
    if (!c->uv_ready || !c->d_U) return fail(BLU_ERR_STATE, "no Hessian factors: evaluate with blu_variance_GH_factored (or blu_eval_device, want_hess != 0) first");
    return ensure_hv(c);
}

// d_out = H d_p on the device (both (L,) doubles in HBM), asynchronous on the context's stream.
extern "C" int blu_hess_matvec_device(blu_ctx *c, const double *d_p, double *d_out)
{
    int rc = use(c);
    if (rc) return rc;
    if (!d_p || !d_out) return fail(BLU_ERR_ARG, "null argument");
    if (c->lo != 0 || c->hi != c->L) return fail(BLU_ERR_STATE, "context owns a slice: use blu_shard_hv_partial / blu_shard_hv_apply");
    if ((rc = hv_ready(c))) return rc;
    double *d_t = c->d_hvpart + 32 * (size_t)c->nsm * 8;
    if ((rc = launch_hv_reduce(c, d_p, d_t))) return rc;
    return launch_hv_apply(c, d_t, d_out);
}

// Host pointers: out[v*L + i] = (H p_v)_i for nvec vectors stored one after the other.
extern "C" int blu_hess_matvec(blu_ctx *c, const double *p, int nvec, double *out)
{
    int rc = use(c);
    if (rc) return rc;
    if (!p || !out) return fail(BLU_ERR_ARG, "null argument");
    if (nvec < 0) return fail(BLU_ERR_ARG, "negative nvec");
    c->launches = 0;
    for (int v = 0; v < nvec; ++v) {
        if ((rc = hv_ready(c))) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->d_hvp, p + (size_t)v * c->L, sizeof(double) * c->L, cudaMemcpyHostToDevice, c->stream));
        if ((rc = blu_hess_matvec_device(c, c->d_hvp, c->d_hvout))) return rc;
        CUDA_TRY(cudaMemcpyAsync(out + (size_t)v * c->L, c->d_hvout, sizeof(double) * c->L, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));     // d_hvp / d_hvout are reused by the next vector
    }
    return BLU_OK;
}

// variance + gradient + Hessian FACTORS (U, V stay in HBM): the evaluation behind a Hessian operator.
extern "C" int blu_variance_GH_factored(blu_ctx *c, const double *m, double delta, double *var, double *grad, unsigned *flags)
{
    int rc = use(c);
    if (rc) return rc;
    if (!m || !grad) return fail(BLU_ERR_ARG, "null argument");
    if (c->pending) return fail(BLU_ERR_STATE, "an evaluation is already pending on this context");
    if ((rc = upload_m(c, m))) return rc;
    if ((rc = blu_eval_device(c, c->d_m, delta, 1, 3))) return rc;     // U rows only: the operator never reads V
    CUDA_TRY(cudaMemcpyAsync(grad, c->d_grad, sizeof(double) * c->L, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = fetch_header(c))) return rc;
    const unsigned fl = c->h_hdr->flags;
    if (var) *var = c->h_hdr->scal[0];
    if (flags) *flags = fl;
    if (fl & BLU_FLAG_TINY) {
        c->uv_ready = false;                                       // early-out (misc.py:484): there is no Hessian
        for (long long i = 0; i < c->L; ++i) grad[i] = std::numeric_limits<double>::infinity();
    }
    return BLU_OK;
}

// Group-sharded operator: every rank sums its own rows into t (32 doubles, zero beyond N), the host
// side all-reduces t (N doubles: the only exchange), then every rank applies it to its own rows.
extern "C" int blu_shard_hv_partial(blu_ctx *c, const double *d_p, double *d_t)
{
    int rc = use(c);
    if (rc) return rc;
    if (!d_p || !d_t) return fail(BLU_ERR_ARG, "null argument");
    if ((rc = hv_ready(c))) return rc;
    return launch_hv_reduce(c, d_p, d_t);
}

extern "C" int blu_shard_hv_apply(blu_ctx *c, const double *d_t, double *d_out)
{
    int rc = use(c);
    if (rc) return rc;
    if (!d_t || !d_out) return fail(BLU_ERR_ARG, "null argument");
    if ((rc = hv_ready(c))) return rc;
    return launch_hv_apply(c, d_t, d_out);
}

extern "C" int blu_cleanup_matrix(blu_ctx *c, const double *m, double delta, int mode, double *X, unsigned *flags)
{
    int rc = use(c);
    if (rc) return rc;
    if (!X) return fail(BLU_ERR_ARG, "null X");
    if (mode != 0 && mode != 1) return fail(BLU_ERR_ARG, "mode must be 0 (reference) or 1 (corrected)");
    if ((rc = upload_m(c, m))) return rc;
    if ((rc = blu_eval_device(c, c->d_m, delta, 0, 0))) return rc;
    unsigned fl = 0;
    if ((rc = blu_ctx_last_result(c, nullptr, &fl))) return rc;
    if (flags) *flags = fl;
    if (fl & BLU_FLAG_TINY) return BLU_OK;
    double *d_X = nullptr;
    const size_t n = (size_t)c->N * c->L;
    CUDA_TRY(cudaMalloc(&d_X, sizeof(double) * n));
    cudaError_t e = cudaMemsetAsync(d_X, 0, sizeof(double) * n, c->stream);
    if (e == cudaSuccess) {
        blu_cleanup_kernel<<<c->nsm * 4, 256, 0, c->stream>>>(c->d_cls, (int)c->cls.size(), c->N, c->L, c->d_gidx, c->d_cinv,
                                                              c->d_x, mode, d_X);
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(X, d_X, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_X);
    if (e != cudaSuccess) return fail(BLU_ERR_CUDA, "cleanup_matrix: %s", cudaGetErrorString(e));
    return BLU_OK;
}

// BLUE estimator (compute_BLUE_estimator sap.py:99-119 + PHIinvY0 misc.py:518-544)
extern "C" int blu_blue_estimator(blu_ctx *c, const double *samples, const double *sums_flat, double *mu, double *var,
                                  double *y_out, unsigned *flags)
{
    int rc = use(c);
    if (rc) return rc;
    if (!sums_flat) return fail(BLU_ERR_ARG, "null sums");
    if ((rc = upload_m(c, samples))) return rc;
    if ((rc = blu_eval_device(c, c->d_m, 0.0, 0, 0))) return rc;
    double *d_sums = nullptr, *d_part = nullptr, *d_y = nullptr;
    const size_t ns = (size_t)std::max<long long>(c->gidx_len, 1);
    CUDA_TRY(cudaMalloc(&d_sums, sizeof(double) * ns));
    cudaError_t e = cudaMalloc(&d_part, sizeof(double) * 32 * (size_t)c->grid_grad);
    if (e == cudaSuccess) e = cudaMalloc(&d_y, sizeof(double) * 32);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_sums, sums_flat, sizeof(double) * (size_t)c->gidx_len, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        blu_ysum_kernel<<<c->grid_grad, BLU_STREAM_WARPS * 32, blu_stream_smem_bytes(c->sd, BLU_STREAM_WARPS * 32, (int)c->cls.size(), c->lutlen), c->stream>>>(
            c->d_cls, (int)c->cls.size(), c->N, c->d_chunks, c->nchunks, c->sd, c->d_cinv, c->d_lut, c->lutlen, c->d_gmask, d_sums, d_part);
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) {
        blu_ysum_finish_kernel<<<1, 32, 0, c->stream>>>(d_part, c->grid_grad, c->N, d_y, c->d_hdr);
        e = cudaGetLastError();
        c->launches++;
    }
    std::vector<double> y(32, 0.0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(y.data(), d_y, sizeof(double) * 32, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_hdr, c->d_hdr, sizeof(BluEvalHeader), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_sums); cudaFree(d_part); cudaFree(d_y);
    if (e != cudaSuccess) return fail(BLU_ERR_CUDA, "blue_estimator: %s", cudaGetErrorString(e));
    if (flags) *flags = c->h_hdr->flags;
    if (var) *var = c->h_hdr->scal[0];
    if (mu) *mu = (c->h_hdr->flags & BLU_FLAG_TINY) ? std::numeric_limits<double>::infinity() : c->h_hdr->scal[5];
    if (y_out) memcpy(y_out, y.data(), sizeof(double) * c->N);
    return BLU_OK;
}

// Integer projection: batched candidate variances (misc.py:368-369)
extern "C" int blu_candidate_variances(blu_ctx *c, const double *basephi, int LL, const int64_t *idx, const int64_t *ms,
                                       int64_t ncand, double rcond, double *Vs)
{
    int rc = use(c);
    if (rc) return rc;
    if (!c->have_inv) return fail(BLU_ERR_STATE, "inverses not set");
    if (!basephi || !idx || !ms || !Vs) return fail(BLU_ERR_ARG, "null argument");
    if (LL < 1 || LL > BLU_IP_MAXLL) return fail(BLU_ERR_ARG, "LL=%d outside [1,%d] (misc.py:320 brute-forces at most 24 dimensions)", LL, BLU_IP_MAXLL);
    if (ncand < 1) return BLU_OK;
    for (int t = 0; t < LL; ++t)
        if (idx[t] < 0 || idx[t] >= c->L) return fail(BLU_ERR_ARG, "group id %lld outside [0,%lld)", (long long)idx[t], c->L);
    const int N = c->N, NN = N * N;
    const size_t smem = sizeof(double) * ((size_t)LL * NN + NN + (size_t)BLU_IP_WARPS * N * (N + 1));
    if (smem > 220 * 1024) return fail(BLU_ERR_ARG, "N=%d with LL=%d candidates groups exceeds the shared-memory budget", N, LL);
    double *d_base = nullptr, *d_psis = nullptr, *d_V = nullptr;
    long long *d_idx = nullptr, *d_ms = nullptr;
    cudaError_t e = cudaMalloc(&d_base, sizeof(double) * NN);
    if (e == cudaSuccess) e = cudaMalloc(&d_psis, sizeof(double) * (size_t)LL * NN);
    if (e == cudaSuccess) e = cudaMalloc(&d_V, sizeof(double) * (size_t)ncand);
    if (e == cudaSuccess) e = cudaMalloc(&d_idx, sizeof(long long) * LL);
    if (e == cudaSuccess) e = cudaMalloc(&d_ms, sizeof(long long) * (size_t)LL * (size_t)ncand);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_base, basephi, sizeof(double) * NN, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, idx, sizeof(long long) * LL, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ms, ms, sizeof(long long) * (size_t)LL * (size_t)ncand, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        blu_ip_expand_kernel<<<LL, 128, 0, c->stream>>>(c->d_cls, (int)c->cls.size(), N, LL, d_idx, c->d_gidx, c->d_cinv, d_psis);
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(blu_ip_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        const long long warps = (ncand + 31) / 32;
        const int grid = (int)std::max<long long>(1, std::min<long long>((warps + BLU_IP_WARPS - 1) / BLU_IP_WARPS, (long long)c->nsm * 8));
        blu_ip_candidates_kernel<<<grid, BLU_IP_WARPS * 32, smem, c->stream>>>(N, LL, d_base, d_psis, d_ms, ncand, rcond, d_V);
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(Vs, d_V, sizeof(double) * (size_t)ncand, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_base); cudaFree(d_psis); cudaFree(d_V); cudaFree(d_idx); cudaFree(d_ms);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? BLU_ERR_NOMEM : BLU_ERR_CUDA, "candidate_variances: %s", cudaGetErrorString(e));
    return BLU_OK;
}

// --------------------------------------------------------------------------------------------
// Row f1: structure-exploiting KKT solve of the SDP of sap.py:242-307 (blu_kkt.cuh)
// --------------------------------------------------------------------------------------------
static bool small_inverse(const double *A, int M, double *inv)       // Gauss-Jordan with partial pivoting, M <= 33
{
    std::vector<double> a(A, A + (size_t)M * M);
    for (int i = 0; i < M; ++i) for (int j = 0; j < M; ++j) inv[i * M + j] = (i == j) ? 1.0 : 0.0;
    for (int p = 0; p < M; ++p) {
        int piv = p;
        for (int r = p + 1; r < M; ++r) if (fabs(a[r * M + p]) > fabs(a[piv * M + p])) piv = r;
        if (a[piv * M + p] == 0.0) return false;
        if (piv != p) for (int c = 0; c < M; ++c) { std::swap(a[p * M + c], a[piv * M + c]); std::swap(inv[p * M + c], inv[piv * M + c]); }
        const double dinv = 1.0 / a[p * M + p];
        for (int c = 0; c < M; ++c) { a[p * M + c] *= dinv; inv[p * M + c] *= dinv; }
        for (int r = 0; r < M; ++r) {
            if (r == p) continue;
            const double f = a[r * M + p];
            if (f == 0.0) continue;
            for (int c = 0; c < M; ++c) { a[r * M + c] -= f * a[p * M + c]; inv[r * M + c] -= f * inv[p * M + c]; }
        }
    }
    return true;
}

extern "C" int blu_kkt_solve(blu_ctx *c, int has_t, double scales, int nlin, const double *Gx, const double *d, const double *r,
                             const double *bx, const double *bz, double *ux, double *uz, float *device_ms)
{
    int rc = use(c);
    if (rc) return rc;
    if (!c->have_inv) return fail(BLU_ERR_STATE, "inverses not set");
    if (c->lo != 0 || c->hi != c->L) return fail(BLU_ERR_STATE, "context owns a slice");
    if (!d || !r || !bx || !bz || !ux || !uz || (nlin > 0 && !Gx)) return fail(BLU_ERR_ARG, "null argument");
    if (has_t != 0 && has_t != 1) return fail(BLU_ERR_ARG, "has_t must be 0 or 1");
    const int N = c->N, M = N + 1, MM = M * M;
    const long long L = c->L, n = L + has_t;
    const int Qs = M * (M + 1) / 2, Q = Qs + nlin, QP = ((Q + 1 + 7) / 8) * 8;      // + the right-hand-side column
    if (nlin < 0 || Q > 255) return fail(BLU_ERR_ARG, "capacitance matrix of order %d exceeds 255 (N=%d, %d dense rows)", Q, N, nlin);
    if (M > 22) return fail(BLU_ERR_ARG, "kkt_solve: N = %d is beyond the rows kernel's register tiles (N <= 21)", N);
    for (long long t = 0; t < n + nlin; ++t) if (!(d[t] > 0.0)) return fail(BLU_ERR_ARG, "scaling d[%lld] is not positive", t);
    // host: r^-1, Lam = r^-T r^-1, Wm = Lam Z Lam  (Z = the 's' part of bz, a symmetric matrix)
    std::vector<double> rinv(MM), Lam(MM), tmp(MM), Wm(MM);
    if (!small_inverse(r, M, rinv.data())) return fail(BLU_ERR_ARG, "the scaling matrix r is singular");
    for (int a = 0; a < M; ++a) for (int b = 0; b < M; ++b) { double s = 0.0; for (int q = 0; q < M; ++q) s += rinv[q * M + a] * rinv[q * M + b]; Lam[a * M + b] = s; }
    const double *Z = bz + n + nlin;
    auto sandwich = [&](const double *X, double *out) {            // out = Lam X Lam
        for (int a = 0; a < M; ++a) for (int b = 0; b < M; ++b) { double s = 0.0; for (int q = 0; q < M; ++q) s += Lam[a * M + q] * X[q * M + b]; tmp[a * M + b] = s; }
        for (int a = 0; a < M; ++a) for (int b = 0; b < M; ++b) { double s = 0.0; for (int q = 0; q < M; ++q) s += tmp[a * M + q] * Lam[q * M + b]; out[a * M + b] = s; }
    };
    sandwich(Z, Wm.data());
    // device buffers
    // workspace: an interior-point solve calls this once per iteration with the same sizes, so the buffers (37.7 MB of Bs at 15
    // models) stay with the context instead of being allocated and freed per call (that alone was > 1 ms of a 2.4 ms call)
    struct DevBuf { void *p = nullptr; };
    DevBuf bBs, bsmall, bvec, bpart;
    auto reserve = [&](int slot, size_t bytes, DevBuf &out) -> cudaError_t {
        if (c->kkt_cap[slot] < bytes) {
            if (c->kkt_ws[slot]) { cudaStreamSynchronize(c->stream); cudaFree(c->kkt_ws[slot]); c->kkt_ws[slot] = nullptr; c->kkt_cap[slot] = 0; }
            cudaError_t e_ = cudaMalloc(&c->kkt_ws[slot], bytes);
            if (e_ != cudaSuccess) return e_;
            c->kkt_cap[slot] = bytes;
        }
        out.p = c->kkt_ws[slot];
        return cudaSuccess;
    };
    const int LDC = ((Q + 1 + 3) / 4) * 4;                                        // leading dimension of the capacitance matrix (+ the rhs row)
    const size_t nsmall = (size_t)2 * MM + (size_t)Q * LDC + 256 + 8;             // rinv | Wm | cap (column-major, rhs as row Q) | y | info
    const size_t nvec = (size_t)(n + nlin) + (size_t)nlin * n + 5 * (size_t)n + nlin;   // d | Gx | bx | bz0 | g1tw | rhs | ux
    CUDA_TRY(reserve(0, sizeof(double) * (size_t)n * QP, bBs));
    CUDA_TRY(reserve(1, sizeof(double) * nsmall, bsmall));
    CUDA_TRY(reserve(2, sizeof(double) * nvec, bvec));
    // Gram kernel: ny CTAs share the 4 x 4 tile blocks of one row range (one block per warp), nsplit row ranges = partial tile sets; one wave
    const int ny = (blu_kkt_syrk_blocks(QP) + BLU_SYRK_WARPS - 1) / BLU_SYRK_WARPS;
    const int nsplit = (int)std::max<long long>(1, std::min<long long>(c->nsm / ny, (n + 63) / 64));
    CUDA_TRY(reserve(3, sizeof(double) * 64 * (size_t)nsplit * (size_t)((QP / 8) * (QP / 8 + 1) / 2), bpart));
    double *d_part = (double *)bpart.p;
    double *d_Bs = (double *)bBs.p;
    double *d_rinv = (double *)bsmall.p, *d_Wm = d_rinv + MM, *d_cap = d_Wm + MM, *d_y = d_cap + (size_t)Q * LDC;
    int *d_info = (int *)(d_y + 256);
    double *d_d = (double *)bvec.p, *d_Gx = d_d + (n + nlin), *d_bx = d_Gx + (size_t)nlin * n, *d_bz0 = d_bx + n;
    double *d_g1tw = d_bz0 + (n + nlin), *d_rhs = d_g1tw + n, *d_ux = d_rhs + n;
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMemcpyAsync(d_rinv, rinv.data(), sizeof(double) * MM, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_Wm, Wm.data(), sizeof(double) * MM, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_d, d, sizeof(double) * (n + nlin), cudaMemcpyHostToDevice, st));
    if (nlin) CUDA_TRY(cudaMemcpyAsync(d_Gx, Gx, sizeof(double) * (size_t)nlin * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_bx, bx, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_bz0, bz, sizeof(double) * (n + nlin), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_info, 0xff, sizeof(int), st));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    c->launches = 0;
    {
        // rows kernel: LW lanes per group (16 while M <= 16: two groups per warp); per sub-warp scratch = expanded inverse / staging row |
        // gathered r^-1 columns | Tt | member ids
        int KMAX = 1;
        for (const BluClass &ci : c->cls) KMAX = std::max(KMAX, ci.k);
        const int LW = (M <= 16) ? 16 : 32, GPW = 32 / LW;
        const int KSF = (LW == 16) ? 16 : 22;                                     // blu_kkt_rows_kernel: 2 * NH
        const int SCR = ((KSF * KSF + 2 * KMAX * LW + 16) + 1) & ~1;
        const size_t smem = sizeof(double) * ((size_t)2 * M * (M | 1) + (size_t)BLU_KKT_WARPS * GPW * SCR) + sizeof(uint16_t) * (size_t)((c->lutlen + 7) & ~7);
        const int grid = (int)std::max<long long>(1, std::min<long long>((n + BLU_KKT_WARPS * GPW - 1) / (BLU_KKT_WARPS * GPW), (long long)c->nsm * 2));
        if (LW == 16) {
            CUDA_TRY(cudaFuncSetAttribute(blu_kkt_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            blu_kkt_rows_kernel<16><<<grid, BLU_KKT_WARPS * 32, smem, st>>>(c->d_cls, (int)c->cls.size(), N, L, has_t, scales, c->d_gidx, c->d_cinv, c->d_lut, c->lutlen,
                                                                            d_rinv, d_Wm, d_d, nlin, d_Gx, d_d + n, Q, QP, KMAX, SCR, d_Bs, d_g1tw);
        } else {
            CUDA_TRY(cudaFuncSetAttribute(blu_kkt_rows_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            blu_kkt_rows_kernel<32><<<grid, BLU_KKT_WARPS * 32, smem, st>>>(c->d_cls, (int)c->cls.size(), N, L, has_t, scales, c->d_gidx, c->d_cinv, c->d_lut, c->lutlen,
                                                                            d_rinv, d_Wm, d_d, nlin, d_Gx, d_d + n, Q, QP, KMAX, SCR, d_Bs, d_g1tw);
        }
        KERNEL_CHECK(c);
        blu_kkt_rhs_kernel<<<(int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)c->nsm * 8)), 256, 0, st>>>(
            n, nlin, d_bx, d_bz0, d_d, d_Gx, d_d + n, d_g1tw, Q, QP, d_rhs, d_Bs);
        KERNEL_CHECK(c);
        const int NTQ = QP / 8, npairs = NTQ * (NTQ + 1) / 2;
        const size_t ssm = blu_kkt_syrk_smem(QP);
        CUDA_TRY(cudaFuncSetAttribute(blu_kkt_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
        blu_kkt_syrk_kernel<<<dim3((unsigned)nsplit, (unsigned)ny), BLU_SYRK_WARPS * 32, ssm, st>>>(d_Bs, n, QP, d_part);
        KERNEL_CHECK(c);
        blu_kkt_capfold_kernel<<<npairs * 2, 256, 0, st>>>(d_part, npairs, nsplit, Q, QP, LDC, d_cap);
        KERNEL_CHECK(c);
        blu_kkt_chol_kernel<<<1, BLU_CHOL_T, 0, st>>>(d_cap, Q, LDC, d_y, d_info);
        KERNEL_CHECK(c);
        blu_kkt_apply_kernel<<<(int)std::max<long long>(1, std::min<long long>((n + BLU_KKT_WARPS - 1) / BLU_KKT_WARPS, (long long)c->nsm * 4)), BLU_KKT_WARPS * 32, 0, st>>>(
            d_Bs, n, Q, QP, d_y, d_d, d_rhs, d_ux);
        KERNEL_CHECK(c);
    }
    // G1 ux needs Phi(ux_m): the ordinary Phi kernel on the solution's group part
    int rcp = launch_phi(c, d_ux + has_t, 0.0, 0);
    cudaEventRecord(e1, st);
    std::vector<double> phi((size_t)N * N);
    int info = -1;
    cudaError_t e = (rcp == BLU_OK) ? cudaSuccess : cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaMemcpyAsync(ux, d_ux, sizeof(double) * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(phi.data(), c->d_phi, sizeof(double) * N * N, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (device_ms && e == cudaSuccess) cudaEventElapsedTime(device_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rcp != BLU_OK) return rcp;
    if (e != cudaSuccess) return fail(BLU_ERR_CUDA, "kkt_solve: %s", cudaGetErrorString(e));
    if (info != 0) return fail(BLU_ERR_ARG, "kkt_solve: the capacitance matrix is not positive definite (Cholesky stopped at column %d)", info);
    // uz0 = d^-2 (G0 ux - bz0),  G0 = [-I; Gx]
    for (long long t = 0; t < n; ++t) uz[t] = (-ux[t] - bz[t]) / (d[t] * d[t]);
    for (int q = 0; q < nlin; ++q) {
        double s = 0.0;
        for (long long t = 0; t < n; ++t) s += Gx[(size_t)q * n + t] * ux[t];
        uz[n + q] = (s - bz[n + q]) / (d[n + q] * d[n + q]);
    }
    // uz1 = vec(Lam (G1 ux - Z) Lam),  G1 ux = -scales pad(Phi(ux_m)) - ux_t E_NN
    std::vector<double> X(MM, 0.0), out(MM);
    for (int a = 0; a < N; ++a) for (int b = 0; b < N; ++b) X[a * M + b] = -scales * phi[(size_t)a * N + b];
    if (has_t) X[N * M + N] -= ux[0];
    for (int t = 0; t < MM; ++t) X[t] -= Z[t];
    sandwich(X.data(), out.data());
    memcpy(uz + n + nlin, out.data(), sizeof(double) * MM);
    return BLU_OK;
}

// --------------------------------------------------------------------------------------------
// Batched evaluation of small problems: P problems x B sample vectors in one launch (blu_batch.cuh)
// --------------------------------------------------------------------------------------------
struct blu_batch {
    int device = 0, P = 0, Nmax = 0, capB = 0;
    long long Lm = 0, gtot = 0, Lmax = 0, capC = 0, capG = 0;
    bool mapped = false, resident = false;
    std::vector<blu_ctx *> ctxs;
    std::vector<BluBatchProb> probs;
    std::vector<long long> pre;                 // prefix sums of the problems' group counts
    std::vector<long long *> d_maps;
    BluBatchProb *d_probs = nullptr;
    BluEvalHeader *d_hdrs = nullptr;
    double *d_scratch = nullptr, *d_m = nullptr, *d_var = nullptr, *d_grad = nullptr;
    unsigned *d_flags = nullptr;
    double *h_m = nullptr, *h_var = nullptr, *h_grad = nullptr;
    unsigned *h_flags = nullptr;
    cudaStream_t stream = nullptr;
    size_t smem = 0;
};

static void batch_free_buffers(blu_batch *b)
{
    cudaFree(b->d_hdrs); cudaFree(b->d_scratch); cudaFree(b->d_m); cudaFree(b->d_var); cudaFree(b->d_grad); cudaFree(b->d_flags);
    if (b->h_m) cudaFreeHost(b->h_m);
    if (b->h_var) cudaFreeHost(b->h_var);
    if (b->h_grad) cudaFreeHost(b->h_grad);
    if (b->h_flags) cudaFreeHost(b->h_flags);
    b->d_hdrs = nullptr; b->d_scratch = b->d_m = b->d_var = b->d_grad = nullptr; b->d_flags = nullptr;
    b->h_m = b->h_var = b->h_grad = nullptr; b->h_flags = nullptr;
    b->capB = 0;
}

extern "C" int blu_batch_destroy(blu_batch *b)
{
    if (!b) return BLU_OK;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    batch_free_buffers(b);
    for (auto p : b->d_maps) cudaFree(p);
    cudaFree(b->d_probs);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
    return BLU_OK;
}

// ctxs: P contexts on one device with their inverses set.  maps == NULL: a sample vector is the concatenation of the
// problems' own vectors (Lm = sum of their L is implied).  maps != NULL: all problems read ONE shared vector of length
// Lm through maps[p] (L_p indices), the `m[mappings[n]]` of mosap.py:88,95.
extern "C" int blu_batch_create(blu_ctx **ctxs, int P, const int64_t *const *maps, int64_t Lm, blu_batch **out)
{
    if (!out) return fail(BLU_ERR_ARG, "null out");
    *out = nullptr;
    if (!ctxs || P < 1) return fail(BLU_ERR_ARG, "no problems");
    for (int p = 0; p < P; ++p) {
        if (!ctxs[p]) return fail(BLU_ERR_ARG, "null context %d", p);
        if (ctxs[p]->device != ctxs[0]->device) return fail(BLU_ERR_ARG, "all problems of a batch must live on one device");
        if (!ctxs[p]->have_inv) return fail(BLU_ERR_STATE, "inverses of problem %d not set", p);
        if (ctxs[p]->lo != 0 || ctxs[p]->hi != ctxs[p]->L) return fail(BLU_ERR_STATE, "problem %d owns a slice", p);
        if (maps && !maps[p]) return fail(BLU_ERR_ARG, "null map %d", p);
    }
    int rc = use(ctxs[0]);
    if (rc) return rc;
    blu_batch *b = new (std::nothrow) blu_batch();
    if (!b) return fail(BLU_ERR_NOMEM, "host allocation failed");
    b->device = ctxs[0]->device; b->P = P; b->mapped = maps != nullptr;
    b->ctxs.assign(ctxs, ctxs + P);
    b->pre.assign(P + 1, 0);
    for (int p = 0; p < P; ++p) { b->pre[p + 1] = b->pre[p] + ctxs[p]->L; b->Nmax = std::max(b->Nmax, ctxs[p]->N); }
    b->gtot = b->pre[P];
    b->Lm = maps ? Lm : b->gtot;
    if (b->Lm < 1) { delete b; return fail(BLU_ERR_ARG, "empty sample vector"); }
#define BT_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int code_ = fail(e_ == cudaErrorMemoryAllocation ? BLU_ERR_NOMEM : BLU_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); blu_batch_destroy(b); return code_; } } while (0)
    BT_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    b->probs.resize(P);
    for (int p = 0; p < P; ++p) {
        blu_ctx *c = ctxs[p];
        BluBatchProb &q = b->probs[p];
        q.cls = c->d_cls; q.gidx = c->d_gidx; q.cinv = c->d_cinv; q.gmask = c->d_gmask; q.lut = c->d_lut; q.map = nullptr;
        q.cinv_len = c->cinv_len; q.gidx_len = c->gidx_len; q.lutlen = c->lutlen; q.pad = 0;
        q.L = c->L; q.moff = b->pre[p]; q.goff = b->pre[p]; q.ncls = (int)c->cls.size(); q.N = c->N;
        if (maps) {
            for (long long i = 0; i < c->L; ++i)
                if (maps[p][i] < 0 || maps[p][i] >= Lm) { blu_batch_destroy(b); return fail(BLU_ERR_ARG, "map %d entry %lld outside [0,%lld)", p, i, (long long)Lm); }
            long long *dm = nullptr;
            BT_TRY(cudaMalloc(&dm, sizeof(long long) * c->L));
            b->d_maps.push_back(dm);
            BT_TRY(cudaMemcpyAsync(dm, maps[p], sizeof(long long) * c->L, cudaMemcpyHostToDevice, b->stream));
            q.map = dm;
        }
        CUDA_TRY(cudaStreamSynchronize(c->stream));       // the contexts' setup work (inversion) is complete
    }
    BT_TRY(cudaMalloc(&b->d_probs, sizeof(BluBatchProb) * P));
    BT_TRY(cudaMemcpyAsync(b->d_probs, b->probs.data(), sizeof(BluBatchProb) * P, cudaMemcpyHostToDevice, b->stream));
    long long Lmax = 0, cmax = 0, gmax = 0; int lutmax = 0;
    for (int p = 0; p < P; ++p) {
        Lmax = std::max(Lmax, ctxs[p]->L); cmax = std::max(cmax, ctxs[p]->cinv_len); gmax = std::max(gmax, ctxs[p]->gidx_len);
        lutmax = std::max(lutmax, ctxs[p]->lutlen);
    }
    b->Lmax = Lmax;
    const size_t fixed = sizeof(double) * ((size_t)BLU_BATCH_WARPS * b->Nmax * b->Nmax + 32 + (size_t)Lmax) + BLU_FIN_SCRATCH_BYTES
                         + sizeof(unsigned short) * (size_t)((lutmax + 7) / 8 * 8) + 16;      // + alignment slack of the staged inverses
    // resident mode: the whole problem (inverses + ids) in shared memory next to the fixed part; else 32 KB chunks
    const size_t res_bytes = sizeof(double) * (size_t)cmax + (size_t)((gmax + 15) / 16 * 16);
    b->resident = fixed + res_bytes <= 220 * 1024;
    b->capC = b->resident ? cmax : BLU_BATCH_CHUNK;
    b->capG = b->resident ? (gmax + 15) / 16 * 16 : BLU_BATCH_CHUNK;
    b->smem = fixed + sizeof(double) * (size_t)b->capC + (size_t)b->capG;
    if (b->smem > 220 * 1024) { blu_batch_destroy(b); return fail(BLU_ERR_ARG, "problems too large for the one-CTA-per-evaluation batch (%lld groups): use the per-context calls", Lmax); }
    BT_TRY(cudaFuncSetAttribute(blu_batch_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
    BT_TRY(cudaStreamSynchronize(b->stream));
#undef BT_TRY
    *out = b;
    return BLU_OK;
}

static int batch_reserve(blu_batch *b, int B)
{
    if (B <= b->capB) return BLU_OK;
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    batch_free_buffers(b);
    const size_t PB = (size_t)b->P * B, NN = (size_t)b->Nmax * b->Nmax;
    CUDA_TRY(cudaMalloc(&b->d_hdrs, sizeof(BluEvalHeader) * PB));
    CUDA_TRY(cudaMemsetAsync(b->d_hdrs, 0, sizeof(BluEvalHeader) * PB, b->stream));
    CUDA_TRY(cudaMalloc(&b->d_scratch, sizeof(double) * PB * (3 * NN + 40 + 32)));
    CUDA_TRY(cudaMalloc(&b->d_m, sizeof(double) * (size_t)B * b->Lm));
    CUDA_TRY(cudaMalloc(&b->d_var, sizeof(double) * PB));
    CUDA_TRY(cudaMalloc(&b->d_flags, sizeof(unsigned) * PB));
    CUDA_TRY(cudaMalloc(&b->d_grad, sizeof(double) * (size_t)B * b->gtot));
    // pinned AND mapped: small batches are evaluated straight on these buffers (see blu_batch_eval)
    CUDA_TRY(cudaHostAlloc(&b->h_m, sizeof(double) * (size_t)B * b->Lm, cudaHostAllocMapped));
    CUDA_TRY(cudaHostAlloc(&b->h_var, sizeof(double) * PB, cudaHostAllocMapped));
    CUDA_TRY(cudaHostAlloc(&b->h_flags, sizeof(unsigned) * PB, cudaHostAllocMapped));
    CUDA_TRY(cudaHostAlloc(&b->h_grad, sizeof(double) * (size_t)B * b->gtot, cudaHostAllocMapped));
    b->capB = B;
    return BLU_OK;
}

// m: (B, Lm) host, row-major.  var, flags: (P, B).  grad (optional): for problem p a (B, L_p) block at offset
// B * (L_0 + ... + L_{p-1}); rows of evaluations with BLU_FLAG_TINY are filled with inf (misc.py:484).
extern "C" int blu_batch_eval(blu_batch *b, const double *m, int B, double delta, double *var, unsigned *flags, double *grad)
{
    if (!b) return fail(BLU_ERR_ARG, "null batch");
    if (!m || !var || B < 1) return fail(BLU_ERR_ARG, "bad batch arguments");
    CUDA_TRY(cudaSetDevice(b->device));
    int rc = batch_reserve(b, B);
    if (rc) return rc;
    const size_t PB = (size_t)b->P * B;
    memcpy(b->h_m, m, sizeof(double) * (size_t)B * b->Lm);
    // Small batches (<= 128 KB each way): the kernel reads the sample vectors from and writes variance / flags / gradient to the
    // MAPPED pinned host buffers directly -- one launch and one synchronisation instead of one H2D and three D2H copy operations
    // in front of and behind a 30 us kernel (each ~6 us of DMA set-up on the stream).  Larger batches keep the explicit copies.
    const bool zc = (size_t)B * b->Lm <= 16384 && (size_t)B * b->gtot <= 16384 && !getenv("BLU_BATCH_NO_ZEROCOPY");
    double *k_m = b->d_m, *k_var = b->d_var, *k_grad = b->d_grad;
    unsigned *k_flags = b->d_flags;
    if (zc) {
        CUDA_TRY(cudaHostGetDevicePointer((void **)&k_m, b->h_m, 0));
        CUDA_TRY(cudaHostGetDevicePointer((void **)&k_var, b->h_var, 0));
        CUDA_TRY(cudaHostGetDevicePointer((void **)&k_flags, b->h_flags, 0));
        CUDA_TRY(cudaHostGetDevicePointer((void **)&k_grad, b->h_grad, 0));
    } else
        CUDA_TRY(cudaMemcpyAsync(b->d_m, b->h_m, sizeof(double) * (size_t)B * b->Lm, cudaMemcpyHostToDevice, b->stream));
    const bool dbg = getenv("BLU_DEBUG_TIMING") != nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (dbg) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, b->stream); }
    dim3 grid((unsigned)b->P, (unsigned)B);
    blu_batch_eval_kernel<<<grid, BLU_BATCH_WARPS * 32, b->smem, b->stream>>>(b->d_probs, k_m, b->Lm, delta, grad ? 1 : 0, b->Lmax, b->capC, b->capG, b->resident ? 1 : 0, b->d_hdrs, b->d_scratch,
                                                                             k_var, k_flags, k_grad);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(BLU_ERR_CUDA, "batch kernel launch failed: %s", cudaGetErrorString(e));
    if (dbg) cudaEventRecord(e1, b->stream);
    if (!zc) {
        CUDA_TRY(cudaMemcpyAsync(b->h_var, b->d_var, sizeof(double) * PB, cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaMemcpyAsync(b->h_flags, b->d_flags, sizeof(unsigned) * PB, cudaMemcpyDeviceToHost, b->stream));
        if (grad) CUDA_TRY(cudaMemcpyAsync(b->h_grad, b->d_grad, sizeof(double) * (size_t)B * b->gtot, cudaMemcpyDeviceToHost, b->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (dbg) {
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        BluEvalHeader h0;
        cudaMemcpy(&h0, b->d_hdrs, sizeof(BluEvalHeader), cudaMemcpyDeviceToHost);
        auto us = [&](int i) { return (double)(h0.stamp[i] - h0.stamp[0]) * 1e-3; };
        fprintf(stderr, "[blu] batch kernel %d x %d: %.1f us | CTA 0: staged +%.1f, Phi summed +%.1f, Phi ready +%.1f, pinv +%.1f, finish +%.1f, grad +%.1f us\n",
                b->P, B, ms * 1e3f, us(1), us(4), us(6), us(7), us(8), grad ? us(9) : 0.0);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    memcpy(var, b->h_var, sizeof(double) * PB);
    if (flags) memcpy(flags, b->h_flags, sizeof(unsigned) * PB);
    if (grad) {
        memcpy(grad, b->h_grad, sizeof(double) * (size_t)B * b->gtot);
        for (int p = 0; p < b->P; ++p)
            for (int v = 0; v < B; ++v)
                if (b->h_flags[(size_t)p * B + v] & BLU_FLAG_TINY) {
                    double *row = grad + b->pre[p] * B + (long long)v * b->ctxs[p]->L;
                    for (long long i = 0; i < b->ctxs[p]->L; ++i) row[i] = std::numeric_limits<double>::infinity();
                }
    }
    return BLU_OK;
}

// ---- group-sharded phases -----------------------------------------------------------------
extern "C" int blu_ctx_set_slice(blu_ctx *c, int64_t lo, int64_t hi)
{
    if (!c) return fail(BLU_ERR_ARG, "null context");
    if (lo < 0 || hi > c->L || lo > hi) return fail(BLU_ERR_ARG, "slice [%lld,%lld) outside [0,%lld]", (long long)lo, (long long)hi, c->L);
    int rc = use(c);
    if (rc) return rc;
    if (c->is_clone) return fail(BLU_ERR_STATE, "a clone keeps its parent's slice: set the slice on the parent, then clone");
    c->lo = lo; c->hi = hi;
    c->tiles_valid = false;
    c->uv_ready = false; c->v_ready = false;      // factors of another slice are not this slice's factors
    return build_chunks(c);
}

extern "C" int blu_shard_phi(blu_ctx *c, const double *d_m)
{
    int rc = use(c);
    if (rc) return rc;
    if (!c->have_inv) return fail(BLU_ERR_STATE, "inverses not set");
    if (!d_m) d_m = c->d_m;
    c->launches = 0;
    return launch_phi(c, d_m, 0.0, 2);
}

extern "C" int blu_shard_finish(blu_ctx *c, double delta, int want_grad, int want_uv)
{
    int rc = use(c);
    if (rc) return rc;
    const int NN = c->N * c->N;
    // BLU_BUF_PHI now holds the all-reduced upper-triangle sum; supp / max|m| were reduced by the host
    blu_phi_finish_kernel<<<1, BLU_FIN_THREADS, sizeof(double) * NN * BLU_FIN_SEG, c->stream>>>(
        c->N, 0, c->d_part, delta, 1, c->d_phi, c->d_pinv, c->d_x, c->d_S, c->d_hdr, c->peers);
    KERNEL_CHECK(c);
    if (want_grad || want_uv) return launch_grad(c, want_uv);
    return BLU_OK;
}

extern "C" int blu_shard_hess(blu_ctx *c, int64_t row_lo, int64_t row_hi)
{
    int rc = use(c);
    if (rc) return rc;
    if (!c->d_U || !c->v_ready) return fail(BLU_ERR_STATE, "U/V not computed (evaluate with want_uv = 1)");
    if (row_lo < 0 || row_hi > c->L || row_lo > row_hi) return fail(BLU_ERR_ARG, "row panel [%lld,%lld) outside [0,%lld]", (long long)row_lo, (long long)row_hi, c->L);
    return launch_hess(c, false, row_lo, row_hi);
}

// ---- fused multi-GPU path: all-reduce of the partial Phi over NVLink peer memory ---------------
extern "C" int blu_ctx_peer_handle(blu_ctx *c, void *handle64)
{
    int rc = use(c);
    if (rc) return rc;
    if (!handle64) return fail(BLU_ERR_ARG, "null handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!c->d_xchg) CUDA_TRY(cudaMalloc(&c->d_xchg, sizeof(BluXchg)));
    // a fresh inbox for every connection: stale flags of an earlier one must never satisfy a wait
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemset(c->d_xchg, 0, sizeof(BluXchg)));
    CUDA_TRY(cudaMemset(&c->d_hdr->epoch, 0, sizeof(unsigned long long)));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->d_xchg));
    memcpy(handle64, &h, 64);
    return BLU_OK;
}

// Every rank calls blu_ctx_peer_handle, the handles are all-gathered on the host (that exchange is also the
// barrier that orders every rank's inbox reset before anybody's first message), then blu_ctx_peer_connect.
extern "C" int blu_ctx_peer_connect(blu_ctx *c, int world, int rank, const void *handles)
{
    int rc = use(c);
    if (rc) return rc;
    if (world < 1 || world > BLU_MAX_PEERS || rank < 0 || rank >= world) return fail(BLU_ERR_ARG, "bad world/rank %d/%d", world, rank);
    if (!c->d_xchg) return fail(BLU_ERR_STATE, "call blu_ctx_peer_handle first");
    if (world > 1 && !handles) return fail(BLU_ERR_ARG, "null handles");
    if (c->N * c->N + 40 > BLU_XCHG_DOUBLES) return fail(BLU_ERR_ARG, "N too large for the exchange slot");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);      // mappings of an earlier connection
    c->ipc_opened.clear();
    BluPeers p{};
    p.world = world; p.rank = rank;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { p.peer[r] = c->d_xchg; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)r, 64);
        void *ptr = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->ipc_opened.push_back(ptr);
        p.peer[r] = (BluXchg *)ptr;
    }
    c->peers = p;
    return BLU_OK;
}

extern "C" int blu_shard_eval_fused(blu_ctx *c, const double *d_m, double delta, int want_grad, int want_uv)
{
    int rc = use(c);
    if (rc) return rc;
    if (!c->have_inv) return fail(BLU_ERR_STATE, "inverses not set");
    if (c->peers.world < 1) return fail(BLU_ERR_STATE, "peers not connected: call blu_ctx_peer_connect first");
    if (!d_m) d_m = c->d_m;
    c->launches = 0;                  // the evaluation count (epoch) lives on the device: every rank's finish step increments its own
    rc = launch_phi(c, d_m, delta, 3);
    if (rc) return rc;
    if (want_grad || want_uv) return launch_grad(c, want_uv);
    return BLU_OK;
}

// --------------------------------------------------------------------------------------------
// kernel (4): pilot covariance
// --------------------------------------------------------------------------------------------
extern "C" int blu_pilot_covariance(int device, const double *Y, int64_t n, int N, int y_on_device,
                                    double *s1, double *S2, double *C_hat, float *kernel_ms)
{
    int rc = need_device(device);
    if (rc) return rc;
    if (!Y || n < 1 || N < 1 || N > BLU_MAX_MODELS) return fail(BLU_ERR_ARG, "bad pilot-sample arguments");
    return blu_gram_run(Y, n, N, y_on_device, s1, S2, C_hat, kernel_ms, g_err);
}

extern "C" int blu_pilot_sums(int device, const double *Y, int64_t n, int N, int n_out, int64_t ystride, int y_on_device,
                              int telescoped, void *producer_stream, double *sums, int sums_on_device, float *kernel_ms)
{
    int rc = need_device(device);
    if (rc) return rc;
    if (!Y || !sums || n < 1 || N < 1 || N > BLU_MAX_MODELS || n_out < 1) return fail(BLU_ERR_ARG, "bad pilot-sample arguments");
    if (ystride != 0 && ystride < n * N) return fail(BLU_ERR_ARG, "output stride smaller than one sample matrix");
    return blu_gram_sums(Y, n, N, n_out, ystride, y_on_device, telescoped, (cudaStream_t)producer_stream, sums, sums_on_device, kernel_ms, g_err);
}

extern "C" int blu_pilot_finalize(const double *sums, int64_t n_total, int N, int telescoped, double *s1, double *S2, double *C_hat,
                                  double *d1, double *d2, double *dV)
{
    if (!sums || n_total < 1 || N < 1 || N > BLU_MAX_MODELS) return fail(BLU_ERR_ARG, "bad pilot-sum arguments");
    blu_gram_finalize(sums, n_total, N, telescoped, s1, S2, C_hat, d1, d2, dV);
    return BLU_OK;
}

// --------------------------------------------------------------------------------------------
// Level 1
// --------------------------------------------------------------------------------------------
extern "C" int blu_assemble_psi_c(double *psi, int N, int k, int Lk, const int64_t *groupsk, const double *invcovsk)
{
    int rc = need_device(0);
    if (rc) return rc;
    return blu_l1_psi(psi, N, k, Lk, groupsk, invcovsk, g_err);
}
extern "C" int blu_objectiveK_c(double *PHI, int N, int k, int Lk, const double *mk, const int64_t *groupsk, const double *invcovsk)
{
    int rc = need_device(0);
    if (rc) return rc;
    return blu_l1_phi(PHI, N, k, Lk, mk, nullptr, groupsk, invcovsk, g_err);
}
extern "C" int blu_objectiveK_c_i64(double *PHI, int N, int k, int Lk, const int64_t *mk, const int64_t *groupsk, const double *invcovsk)
{
    int rc = need_device(0);
    if (rc) return rc;
    return blu_l1_phi(PHI, N, k, Lk, nullptr, mk, groupsk, invcovsk, g_err);
}
extern "C" int blu_cleanupK_c(double *X, int N, int k, int Lk, const int64_t *groupsk, const double *invcovsk, const double *invPHI_0)
{
    int rc = need_device(0);
    if (rc) return rc;
    return blu_l1_cleanup(X, N, k, Lk, groupsk, invcovsk, invPHI_0, g_err);
}
extern "C" int blu_gradK_c(double *grad, int N, int k, int Lk, const int64_t *groupsk, const double *invcovsk, const double *invPHI_0)
{
    int rc = need_device(0);
    if (rc) return rc;
    return blu_l1_grad(grad, N, k, Lk, groupsk, invcovsk, invPHI_0, g_err);
}
extern "C" int blu_hessKQ_c(double *hess, int N, int k, int q, int Lk, int Lq, const int64_t *groupsk, const int64_t *groupsq,
                            const double *invcovsk, const double *invcovsq, const double *invPHI)
{
    int rc = need_device(0);
    if (rc) return rc;
    return blu_l1_hess(hess, N, k, q, Lk, Lq, groupsk, groupsq, invcovsk, invcovsq, invPHI, g_err);
}
