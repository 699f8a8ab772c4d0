// blu_gram.cuh -- kernel (4): pilot-sample covariance as an FP64 tensor-core Gram contraction.
//
// Replaces the per-sample Python accumulation of blue_fn.py:159-167 (sumse[i] += P_i,
// sumsc[j,i] += P_i P_j) and the formula of blue_models.py:333,
//     C_hat = sumsc / n - outer(sumse, sumse) / n^2      (biased, one pass).
// Y is the (n, N) row-major sample matrix.  The model axis is padded to NT tiles of 8 columns with
// one extra column of ones at index N, so the column sums come out of the same contraction:
// G = [Y 1]^T [Y 1]  =>  S2 = G[:N,:N], s1 = G[:N, N].
// A warp walks its contiguous slab of samples (bulk-async staged through shared memory) 4 at a time: the fragment of tile t held by a lane
// (sample lane&3, column 8t + lane>>2) is at once the A operand (row = column index, k = sample)
// and the B operand (k = sample, col = column index) of mma.m8n8k4.f64, so every loaded double is
// used NT times.  Only tile pairs ti <= tj are accumulated.  Reduction: warps -> CTA in shared
// memory in warp order, CTAs -> result in block order by a second kernel: no atomics,
// bit-reproducible.  The kernel is an HBM stream of 8 n N bytes.
#pragma once
#include <string>
#include "blu_common.cuh"
#include "blu_hess.cuh"
#include "blu_stream.cuh"

#define BLU_GRAM_WARPS 8

#define BLU_GRAM_STAGE_DOUBLES 544           // >= 16 samples x 32 models + skew/round-up slack

// Each warp owns a contiguous slab of samples and streams it through a private two-stage
// shared-memory ring with bulk asynchronous copies (cp.async.bulk + mbarrier, as blu_stream.cuh):
// a stage holds `spc` samples (spc*N doubles, one contiguous span of Y).  Fragments are read
// from shared memory; nothing in the MMA loop waits on a global load.
template <int NT>
__global__ void __launch_bounds__(BLU_GRAM_WARPS * 32)
blu_gram_kernel(const double *__restrict__ Y, long long n, int N, long long slab, int spc, double *__restrict__ part)
{
    constexpr int NPG = 8 * NT;
    constexpr int NPAIR = NT * (NT + 1) / 2;
    extern __shared__ __align__(16) double gsm_raw[];   // [WARPS][2][STAGE] stages | [WARPS][NPG*NPG] reduction | barriers
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *stage0 = gsm_raw + (size_t)(2 * w) * BLU_GRAM_STAGE_DOUBLES;
    double *stage1 = stage0 + BLU_GRAM_STAGE_DOUBLES;
    double *sred = gsm_raw + (size_t)2 * BLU_GRAM_WARPS * BLU_GRAM_STAGE_DOUBLES + (size_t)w * NPG * NPG;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(gsm_raw + (size_t)2 * BLU_GRAM_WARPS * BLU_GRAM_STAGE_DOUBLES
                                                                      + (size_t)BLU_GRAM_WARPS * NPG * NPG) + 2 * w;
    if (lane == 0) { blu_mbar_init(&bars[0], 1); blu_mbar_init(&bars[1], 1); blu_mbar_fence_init(); }
    __syncwarp();
    const int ks = lane & 3, cq = lane >> 2;
    const long long gw = (long long)blockIdx.x * BLU_GRAM_WARPS + w;
    const long long s0 = gw * slab;
    long long s1 = s0 + slab;
    if (s1 > n) s1 = n;
    double acc[NPAIR][2];
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) { acc[p][0] = 0.0; acc[p][1] = 0.0; }

    auto issue = [&](long long sb, int st) -> int {       // copy samples [sb, sb+cnt) into stage st; returns the skew
        const long long cnt = (s1 - sb) < spc ? (s1 - sb) : spc;
        const unsigned long long addr = (unsigned long long)(Y + sb * N);
        const int skew = (int)((addr & 15ull) >> 3);
        if (lane == 0) {
            const unsigned bytes = (unsigned)(((skew + cnt * N) * 8 + 15) & ~15ll);
            blu_mbar_expect_tx(&bars[st], bytes);
            blu_bulk_g2s(st ? stage1 : stage0, (const void *)(addr & ~15ull), bytes, &bars[st]);
        }
        return skew;
    };
    int skew_cur = 0, skew_nxt = 0;
    if (s0 < s1) skew_cur = issue(s0, 0);
    int it = 0;
    for (long long sb = s0; sb < s1; sb += spc, ++it) {
        const int st = it & 1;
        if (sb + spc < s1) skew_nxt = issue(sb + spc, st ^ 1);
        blu_mbar_wait(&bars[st], (unsigned)((it >> 1) & 1));
        const double *base = (st ? stage1 : stage0) + skew_cur;
        const int cnt = (int)((s1 - sb) < spc ? (s1 - sb) : spc);
        for (int r0 = 0; r0 < cnt; r0 += 4) {
            const int row = r0 + ks;
            const bool rok = row < cnt;
            double f[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int col = 8 * t + cq;
                f[t] = rok ? ((col < N) ? base[row * N + col] : (col == N ? 1.0 : 0.0)) : 0.0;
            }
            int p = 0;
#pragma unroll
            for (int ti = 0; ti < NT; ++ti)
#pragma unroll
                for (int tj = ti; tj < NT; ++tj) { blu_dmma(acc[p][0], acc[p][1], f[ti], f[tj]); ++p; }
        }
        __syncwarp();                                   // stage consumed before it is refilled
        skew_cur = skew_nxt;
    }
    // warp tile -> shared (C fragment: row lane>>2, cols 2*(lane&3)+{0,1})
    {
        int p = 0;
#pragma unroll
        for (int ti = 0; ti < NT; ++ti)
#pragma unroll
            for (int tj = ti; tj < NT; ++tj) {
                const int r = 8 * ti + cq, c = 8 * tj + 2 * ks;
                sred[r * NPG + c] = acc[p][0];
                sred[r * NPG + c + 1] = acc[p][1];
                ++p;
            }
    }
    __syncthreads();
    const double *sall = gsm_raw + (size_t)2 * BLU_GRAM_WARPS * BLU_GRAM_STAGE_DOUBLES;
    for (int t = threadIdx.x; t < NPG * NPG; t += blockDim.x) {
        const int r = t / NPG, c = t - r * NPG;
        if ((r >> 3) > (c >> 3)) continue;                  // lower tiles are never produced
        double sum = 0.0;
#pragma unroll
        for (int ww = 0; ww < BLU_GRAM_WARPS; ++ww) sum += sall[(size_t)ww * NPG * NPG + t];
        part[(long long)blockIdx.x * NPG * NPG + t] = sum;
    }
}

// Fixed-order reduction over CTAs + covariance formula, spread over E/32 CTAs: CTA b owns the 32
// entries [32b, 32b+32) of the Gram tile; its 8 warps sum interleaved subsets of the partial tiles
// (coalesced 256-byte rows, independent loads in flight) and are combined in warp order, so the
// association is fixed.  The CTA that finishes last (a ticket counter: control flow only, no
// arithmetic through atomics) turns G into s1, S2 and C_hat.  A single-CTA version of this reduction
// took as long as the streaming kernel itself (1.4 MB of partials behind one SM's load queue).
#define BLU_GRAM_FIN_WARPS 8
__global__ void __launch_bounds__(BLU_GRAM_FIN_WARPS * 32)
blu_gram_finish_kernel(const double *__restrict__ part, int nparts, int NPG, int N, long long n,
                       double *__restrict__ G, unsigned *__restrict__ ticket,
                       double *__restrict__ s1, double *__restrict__ S2, double *__restrict__ Chat)
{
    __shared__ double sh[BLU_GRAM_FIN_WARPS][32];
    __shared__ bool last;
    const int E = NPG * NPG;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * 32 + lane;
    double sum = 0.0;
    if (e < E) {
        const int r = e / NPG, c = e - r * NPG;
        if ((r >> 3) <= (c >> 3)) {                              // lower tiles are never produced
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int p = w;
            for (; p + 3 * BLU_GRAM_FIN_WARPS < nparts; p += 4 * BLU_GRAM_FIN_WARPS) {
                a0 += part[(long long)p * E + e];
                a1 += part[(long long)(p + BLU_GRAM_FIN_WARPS) * E + e];
                a2 += part[(long long)(p + 2 * BLU_GRAM_FIN_WARPS) * E + e];
                a3 += part[(long long)(p + 3 * BLU_GRAM_FIN_WARPS) * E + e];
            }
            for (; p < nparts; p += BLU_GRAM_FIN_WARPS) a0 += part[(long long)p * E + e];
            sum = (a0 + a1) + (a2 + a3);
        }
    }
    sh[w][lane] = sum;
    __syncthreads();
    if (w == 0 && e < E) {
        double g = 0.0;
#pragma unroll
        for (int ww = 0; ww < BLU_GRAM_FIN_WARPS; ++ww) g += sh[ww][lane];
        G[e] = g;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0u;                          // ready for the next call on this buffer
    const double dn = (double)n;
    for (int t = threadIdx.x; t < N * N; t += blockDim.x) {
        const int r = t / N, c = t - r * N;
        const int lo = r < c ? r : c, hi = r < c ? c : r;
        const double g = __ldcg(G + lo * NPG + hi);              // upper triangle holds the sums
        const double a = __ldcg(G + r * NPG + N), b = __ldcg(G + c * NPG + N);
        S2[t] = g;
        Chat[t] = g / dn - (a * b) / (dn * dn);
    }
    for (int t = threadIdx.x; t < N; t += blockDim.x) s1[t] = __ldcg(G + t * NPG + N);
}

template <int NT>
static cudaError_t blu_gram_launch(const double *dY, long long n, int N, int grid, long long slab, double *d_part, cudaStream_t st)
{
    const int spc = (512 / N) & ~3;                      // samples per stage: multiple of 4, <= 512 doubles
    const size_t smem = sizeof(double) * ((size_t)2 * BLU_GRAM_WARPS * BLU_GRAM_STAGE_DOUBLES + (size_t)BLU_GRAM_WARPS * (8 * NT) * (8 * NT))
                        + sizeof(unsigned long long) * 2 * BLU_GRAM_WARPS;
    cudaError_t e = cudaFuncSetAttribute(blu_gram_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    blu_gram_kernel<NT><<<grid, BLU_GRAM_WARPS * 32, smem, st>>>(dY, n, N, slab, spc, d_part);
    return cudaGetLastError();
}

static int blu_gram_run(const double *Y, long long n, int N, int y_on_device, double *s1, double *S2, double *C_hat,
                        float *kernel_ms, std::string &err)
{
    const int NT = (N + 1 + 7) / 8, NPG = 8 * NT;
    cudaDeviceProp prop; int dev = 0;
    cudaGetDevice(&dev);
    cudaGetDeviceProperties(&prop, dev);
    const long long warps_wanted = (n + 255) / 256;                       // >= 256 samples per warp
    int grid = (int)std::max<long long>(1, std::min<long long>((warps_wanted + BLU_GRAM_WARPS - 1) / BLU_GRAM_WARPS,
                                                                  (long long)prop.multiProcessorCount * 2));
    const long long nwarps = (long long)grid * BLU_GRAM_WARPS;
    long long slab = (n + nwarps - 1) / nwarps;
    slab = ((slab + 15) / 16) * 16;
    double *dY = nullptr, *d_part = nullptr, *d_out = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaSuccess;
    auto done = [&](int code) {
        if (!y_on_device) cudaFree(dY);
        cudaFree(d_part); cudaFree(d_out);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (st) cudaStreamDestroy(st);
        if (code) err = std::string("pilot covariance: ") + cudaGetErrorString(e);
        return code;
    };
    if ((e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return done(BLU_ERR_CUDA);
    if (y_on_device) dY = const_cast<double *>(Y);
    else {
        if ((e = cudaMalloc(&dY, sizeof(double) * n * N)) != cudaSuccess) return done(BLU_ERR_NOMEM);
        if ((e = cudaMemcpyAsync(dY, Y, sizeof(double) * n * N, cudaMemcpyHostToDevice, st)) != cudaSuccess) return done(BLU_ERR_CUDA);
    }
    if ((e = cudaMalloc(&d_part, sizeof(double) * NPG * NPG * grid)) != cudaSuccess) return done(BLU_ERR_NOMEM);
    if ((e = cudaMalloc(&d_out, sizeof(double) * (2 * N * N + N + NPG * NPG + 2))) != cudaSuccess) return done(BLU_ERR_NOMEM);
    double *d_G = d_out + 2 * N * N + N;                          // reduced Gram tile, then the ticket counter
    unsigned *d_ticket = reinterpret_cast<unsigned *>(d_G + NPG * NPG);
    if ((e = cudaMemsetAsync(d_ticket, 0, sizeof(unsigned), st)) != cudaSuccess) return done(BLU_ERR_CUDA);
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    switch (NT) {
        case 1: e = blu_gram_launch<1>(dY, n, N, grid, slab, d_part, st); break;
        case 2: e = blu_gram_launch<2>(dY, n, N, grid, slab, d_part, st); break;
        case 3: e = blu_gram_launch<3>(dY, n, N, grid, slab, d_part, st); break;
        case 4: e = blu_gram_launch<4>(dY, n, N, grid, slab, d_part, st); break;
        default: e = blu_gram_launch<5>(dY, n, N, grid, slab, d_part, st); break;
    }
    if (e != cudaSuccess) return done(BLU_ERR_CUDA);
    blu_gram_finish_kernel<<<(NPG * NPG + 31) / 32, BLU_GRAM_FIN_WARPS * 32, 0, st>>>(d_part, grid, NPG, N, n, d_G, d_ticket,
                                                                                     d_out, d_out + N, d_out + N + N * N);
    if ((e = cudaGetLastError()) != cudaSuccess) return done(BLU_ERR_CUDA);
    cudaEventRecord(e1, st);
    std::vector<double> h((size_t)(2 * N * N + N));
    if ((e = cudaMemcpyAsync(h.data(), d_out, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return done(BLU_ERR_CUDA);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(BLU_ERR_CUDA);
    if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
    if (s1) memcpy(s1, h.data(), sizeof(double) * N);
    if (S2) memcpy(S2, h.data() + N, sizeof(double) * N * N);
    if (C_hat) memcpy(C_hat, h.data() + N + N * N, sizeof(double) * N * N);
    return done(BLU_OK);
}
