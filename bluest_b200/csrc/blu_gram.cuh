// blu_gram.cuh -- kernel (4): the pilot-sample sums as an FP64 tensor-core Gram contraction.
//
// Replaces the per-sample Python accumulation of blue_fn.py:147-167:
//     sumse[n][i] += P_i,   sumsc[n][j,i] += P_i P_j,
//     sumsd1[n][i][j] += P_i - P_j,   sumsd2[n][i][j] += (P_i - P_j)^2     (compute_mlmc_differences)
// for n_out outputs in one launch (grid.y = output), and feeds the formulas of blue_models.py:333,339
//     C_hat = sumsc / n - outer(sumse, sumse) / n^2,      dV = sumsd2 / n - (sumsd1 / n)^2
// (evaluated on the host from the (N^2 + N) sums per output: blu_pilot_finalize).
//
// Y[o] is the (n, N) row-major sample matrix of output o.  The model axis is padded to NT tiles of 8 columns with
// one extra column of ones at index N, so the column sums come out of the same contraction:
//     G = [X 1]^T [X 1]   =>   sumsc = G[:N,:N], sumse = G[:N, N].
// TELE = false: X = Y.  TELE = true: X = Z, the TELESCOPED samples Z_0 = Y_0, Z_j = Y_j - Y_{j-1}, formed in
// registers.  Y_i - Y_j is then a sum of consecutive Z columns, so the difference sums of the MLMC pairs are sums
// of Gram entries OF DIFFERENCES: sumsd2[i][j] = sum_{l,l' in (i,j]} G^Z[l,l'] has no cancellation against the
// (much larger) variances of the models themselves, which sumsc[i,i] - 2 sumsc[i,j] + sumsc[j,j] would have for the
// strongly coupled models a multilevel hierarchy consists of; sumse and sumsc are prefix sums of the same G^Z.
//
// A warp walks its contiguous slab of samples (bulk-async staged through a private 3-stage ring) 4 at a time: the
// fragment of tile t held by a lane (sample lane&3, column 8t + lane>>2) is at once the A operand (row = column
// index, k = sample) and the B operand (k = sample, col = column index) of mma.m8n8k4.f64, so every loaded double
// is used NT times.  Only tile pairs ti <= tj are accumulated.  Reduction: warps -> CTA in shared memory in warp
// order; CTAs -> result by the last CTAs to arrive (two levels, fixed order, as in blu_phi.cuh): no atomics on
// data, no second launch, bit-reproducible.  The kernel is an HBM stream of 8 n N bytes per output.
#pragma once
#include <string>
#include <vector>
#include "blu_common.cuh"
#include "blu_hess.cuh"
#include "blu_stream.cuh"

#define BLU_GRAM_WARPS 8
#define BLU_GRAM_NS 2                        // ring depth per warp: two CTAs (16 warps) per SM keep the FP64 tensor pipe fed --
                                             // with 3 stages and one CTA per SM the DMMA pipe was 57 % busy, waiting on its own accumulators
#define BLU_GRAM_STAGE_DOUBLES 544           // >= 16 samples x 32 models + skew/round-up slack
#define BLU_GRAM_GROUP 16                    // CTAs per group of the in-kernel reduction
#define BLU_GRAM_MAXGROUPS 40

// explicit shared-space load of one double (32-bit address)
__device__ __forceinline__ double blu_gram_lds(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

template <int NT, bool TELE>
__global__ void __launch_bounds__(BLU_GRAM_WARPS * 32)
blu_gram_kernel(const double *__restrict__ Yall, long long ystride, long long n, int N, long long slab, int spc,
                double *__restrict__ part_all, double *__restrict__ G_all, unsigned *__restrict__ tickets_all)
{
    constexpr int NPG = 8 * NT;
    constexpr int E = NPG * NPG;
    constexpr int NPAIR = NT * (NT + 1) / 2;
    extern __shared__ __align__(16) double gsm_raw[];   // [WARPS][NS][STAGE] stages | [WARPS][E] reduction | barriers
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.y;
    const double *__restrict__ Y = Yall + (long long)o * ystride;
    double *__restrict__ part = part_all + (size_t)o * (gridDim.x + BLU_GRAM_MAXGROUPS) * E;
    double *__restrict__ G = G_all + (size_t)o * E;
    unsigned *tickets = tickets_all + (size_t)o * (BLU_GRAM_MAXGROUPS + 1);
    double *stages = gsm_raw + (size_t)(BLU_GRAM_NS * w) * BLU_GRAM_STAGE_DOUBLES;
    double *sall = gsm_raw + (size_t)BLU_GRAM_NS * BLU_GRAM_WARPS * BLU_GRAM_STAGE_DOUBLES;
    double *sred = sall + (size_t)w * E;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sall + (size_t)BLU_GRAM_WARPS * E) + BLU_GRAM_NS * w;
    if (lane == 0) {
        for (int s = 0; s < BLU_GRAM_NS; ++s) blu_mbar_init(&bars[s], 1);
        blu_mbar_fence_init();
    }
    __syncwarp();
    const int ks = lane & 3, cq = lane >> 2;
    const long long gw = (long long)blockIdx.x * BLU_GRAM_WARPS + w;
    const long long s0 = gw * slab;
    long long s1 = s0 + slab;
    if (s1 > n) s1 = n;
    // two accumulator sets, alternating between consecutive blocks of 4 samples: a DMMA depends on the one two blocks
    // back, not on the previous one (the FP64 tensor pipe stalled on that dependency)
    double acc[2][NPAIR][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int p = 0; p < NPAIR; ++p) { acc[h][p][0] = 0.0; acc[h][p][1] = 0.0; }

    auto issue = [&](long long sb, int st) {              // copy samples [sb, sb+cnt) into stage st
        const long long cnt = (s1 - sb) < spc ? (s1 - sb) : spc;
        const unsigned long long addr = (unsigned long long)(Y + sb * N);
        const int skew = (int)((addr & 15ull) >> 3);
        if (lane == 0) {
            const unsigned bytes = (unsigned)(((skew + cnt * N) * 8 + 15) & ~15ll);
            blu_mbar_expect_tx(&bars[st], bytes);
            blu_bulk_g2s(stages + (size_t)st * BLU_GRAM_STAGE_DOUBLES, (const void *)(addr & ~15ull), bytes, &bars[st]);
        }
    };
    // A lane's fragment of tile t is (sample ks, column 8 t + cq) of every block of 4 samples: a FIXED byte offset inside the block
    // plus a per-row stride -- or, for the column of ones and the padding columns, a constant slot with stride 0.  Addresses are
    // 32-bit shared-space values advanced by additions only: no predicates, no index arithmetic, no selects in the block loop
    // (the first version spent 4.5 instructions per DMMA on them and kept the ring's skew in local memory).
    __shared__ double s_const[2];
    if (threadIdx.x == 0) { s_const[0] = 0.0; s_const[1] = 1.0; }
    __syncthreads();
    const unsigned zero_a = blu_smem_u32(&s_const[0]), one_a = blu_smem_u32(&s_const[1]);
    const unsigned stage_a = blu_smem_u32(stages);
    unsigned off1[NT], str1[NT], off2[NT], str2[NT];      // off: relative to the chunk's first sample (str != 0) or absolute (str == 0)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int col = 8 * t + cq;
        if (col < N) { off1[t] = (unsigned)((ks * N + col) * 8); str1[t] = (unsigned)(N * 8); }
        else { off1[t] = (col == N) ? one_a : zero_a; str1[t] = 0u; }
        if (TELE && col >= 1 && col < N) { off2[t] = off1[t] - 8u; str2[t] = str1[t]; }     // Z_j = Y_j - Y_{j-1}
        else { off2[t] = zero_a; str2[t] = 0u; }
    }
    long long sissue = s0;                                // next chunk to issue
    int nissued = 0;
    for (; nissued < BLU_GRAM_NS - 1 && sissue < s1; ++nissued, sissue += spc) issue(sissue, nissued);
    int it = 0;
    for (long long sb = s0; sb < s1; sb += spc, ++it) {
        const int st = it % BLU_GRAM_NS;
        if (sissue < s1) {                                // refill the stage consumed one step ago
            issue(sissue, (it + BLU_GRAM_NS - 1) % BLU_GRAM_NS);
            sissue += spc;
        }
        blu_mbar_wait(&bars[st], (unsigned)((it / BLU_GRAM_NS) & 1));
        const int cnt = (int)((s1 - sb) < spc ? (s1 - sb) : spc);
        const int skew = (int)((((unsigned long long)(Y + sb * N)) & 15ull) >> 3);
        if (cnt & 7) {                                    // last chunk of the slab: the missing samples of its last block of 8 count as zeros
            double *sp = stages + (size_t)st * BLU_GRAM_STAGE_DOUBLES + skew + (size_t)cnt * N;
            const int nz = (((cnt + 7) & ~7) - cnt) * N;
            for (int z = lane; z < nz; z += 32) sp[z] = 0.0;
            __syncwarp();
        }
        const unsigned base = stage_a + (unsigned)((st * BLU_GRAM_STAGE_DOUBLES + skew) * 8);
        unsigned pa[NT], pb[NT], qa[NT], qb[NT];          // block h = 0 / h = 1 (4 samples further down)
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            pa[t] = str1[t] ? base + off1[t] : off1[t];
            pb[t] = pa[t] + 4u * str1[t];
            qa[t] = str2[t] ? base + off2[t] : off2[t];
            qb[t] = qa[t] + 4u * str2[t];
        }
        const int nb8 = (cnt + 7) >> 3;
        for (int b8 = 0; b8 < nb8; ++b8) {
            double f0[NT], f1[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) { f0[t] = blu_gram_lds(pa[t]); f1[t] = blu_gram_lds(pb[t]); }
            if (TELE) {
#pragma unroll
                for (int t = 0; t < NT; ++t) { f0[t] -= blu_gram_lds(qa[t]); f1[t] -= blu_gram_lds(qb[t]); }
            }
            int p = 0;
#pragma unroll
            for (int ti = 0; ti < NT; ++ti)
#pragma unroll
                for (int tj = ti; tj < NT; ++tj) { blu_dmma(acc[0][p][0], acc[0][p][1], f0[ti], f0[tj]); ++p; }
            p = 0;
#pragma unroll
            for (int ti = 0; ti < NT; ++ti)
#pragma unroll
                for (int tj = ti; tj < NT; ++tj) { blu_dmma(acc[1][p][0], acc[1][p][1], f1[ti], f1[tj]); ++p; }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                pa[t] += 8u * str1[t]; pb[t] += 8u * str1[t];
                if (TELE) { qa[t] += 8u * str2[t]; qb[t] += 8u * str2[t]; }
            }
        }
        __syncwarp();                                   // stage consumed before it is refilled
    }
    // warp tile -> shared (C fragment: row lane>>2, cols 2*(lane&3)+{0,1})
    {
        int p = 0;
#pragma unroll
        for (int ti = 0; ti < NT; ++ti)
#pragma unroll
            for (int tj = ti; tj < NT; ++tj) {
                const int r = 8 * ti + cq, c = 8 * tj + 2 * ks;
                sred[r * NPG + c] = acc[0][p][0] + acc[1][p][0];
                sred[r * NPG + c + 1] = acc[0][p][1] + acc[1][p][1];
                ++p;
            }
    }
    __syncthreads();
    // Only the upper tile pairs exist.  All three reduction levels walk them as 16-byte pairs: NPAIR * 32 double2 per tile set, one
    // per thread at 20 models (192 <= 256 threads), so the last CTA of a group has its 16 loads -- and the last group its <= 20 --
    // in flight ONCE instead of once per 256 elements of the padded square (three dependent L2 round trips per level before).
    constexpr int NE2 = NPAIR * 32;
    auto elem = [&](int t) -> int {                       // double2 index -> element offset (row-major NPG x NPG, even column)
        const int p = t >> 5, q = t & 31;
        int ti = 0, rem = p;
        while (rem >= NT - ti) { rem -= NT - ti; ++ti; }
        return (8 * ti + (q >> 2)) * NPG + 8 * (ti + rem) + 2 * (q & 3);
    };
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int t = tid; t < NE2; t += nthr) {
        const int e = elem(t);
        double2 sum = make_double2(0.0, 0.0);
#pragma unroll
        for (int ww = 0; ww < BLU_GRAM_WARPS; ++ww) {
            const double2 v = *reinterpret_cast<const double2 *>(sall + (size_t)ww * E + e);
            sum.x += v.x; sum.y += v.y;
        }
        *reinterpret_cast<double2 *>(part + (size_t)blockIdx.x * E + e) = sum;
    }

    // ---- fused reduction over CTAs: last CTA of a group folds the group (CTA order), last group folds the groups ----
    __shared__ int s_last;
    const int grp = blockIdx.x / BLU_GRAM_GROUP;
    const int ngrp = (gridDim.x + BLU_GRAM_GROUP - 1) / BLU_GRAM_GROUP;
    const int members = min(BLU_GRAM_GROUP, (int)gridDim.x - grp * BLU_GRAM_GROUP);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(&tickets[1 + grp], 1u);
        s_last = (t == (unsigned)(members - 1));
        if (s_last) tickets[1 + grp] = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double *part2 = part + (size_t)gridDim.x * E;
    for (int t = tid; t < NE2; t += nthr) {
        const int e = elem(t);
        double2 v[BLU_GRAM_GROUP];
        const double *pp = part + (size_t)grp * BLU_GRAM_GROUP * E + e;
#pragma unroll
        for (int cc = 0; cc < BLU_GRAM_GROUP; ++cc)
            v[cc] = (cc < members) ? __ldcg(reinterpret_cast<const double2 *>(pp + (size_t)cc * E)) : make_double2(0.0, 0.0);
        double2 sum = make_double2(0.0, 0.0);
#pragma unroll
        for (int cc = 0; cc < BLU_GRAM_GROUP; ++cc) { sum.x += v[cc].x; sum.y += v[cc].y; }
        *reinterpret_cast<double2 *>(part2 + (size_t)grp * E + e) = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(&tickets[0], 1u);
        s_last = (t == (unsigned)(ngrp - 1));
        if (s_last) tickets[0] = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int t = tid; t < NE2; t += nthr) {
        const int e = elem(t);
        double2 sum = make_double2(0.0, 0.0);
        for (int g0 = 0; g0 < ngrp; g0 += 20) {          // 296 CTAs = 19 groups: one batch of loads in flight
            double2 v[20];
#pragma unroll
            for (int u = 0; u < 20; ++u)
                v[u] = (g0 + u < ngrp) ? __ldcg(reinterpret_cast<const double2 *>(part2 + (size_t)(g0 + u) * E + e)) : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < 20; ++u) { sum.x += v[u].x; sum.y += v[u].y; }
        }
        *reinterpret_cast<double2 *>(G + e) = sum;
    }
}

// G tiles (upper tile pairs) of every output -> packed sums: per output N*N Gram entries (full, symmetric) then the
// N column sums.  One small CTA per output.
__global__ void blu_gram_pack_kernel(const double *__restrict__ G_all, int NPG, int N, double *__restrict__ sums)
{
    const double *G = G_all + (size_t)blockIdx.x * NPG * NPG;
    double *out = sums + (size_t)blockIdx.x * (N * N + N);
    for (int t = threadIdx.x; t < N * N; t += blockDim.x) {
        const int r = t / N, c = t - r * N;
        const int lo = r < c ? r : c, hi = r < c ? c : r;
        out[t] = G[lo * NPG + hi];
    }
    for (int t = threadIdx.x; t < N; t += blockDim.x) out[N * N + t] = G[t * NPG + N];
}

template <int NT>
static cudaError_t blu_gram_launch(bool tele, const double *dY, long long ystride, long long n, int N, int n_out, int grid, long long slab,
                                   double *d_part, double *d_G, unsigned *d_tickets, cudaStream_t st)
{
    const int spc = (512 / N) & ~7;                      // samples per stage: multiple of 8 (the block loop consumes 8 at a time), <= 512 doubles
    const size_t smem = sizeof(double) * ((size_t)BLU_GRAM_NS * BLU_GRAM_WARPS * BLU_GRAM_STAGE_DOUBLES + (size_t)BLU_GRAM_WARPS * (8 * NT) * (8 * NT))
                        + sizeof(unsigned long long) * BLU_GRAM_NS * BLU_GRAM_WARPS;
    cudaError_t e;
    dim3 g((unsigned)grid, (unsigned)n_out);
    if (tele) {
        if ((e = cudaFuncSetAttribute(blu_gram_kernel<NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        blu_gram_kernel<NT, true><<<g, BLU_GRAM_WARPS * 32, smem, st>>>(dY, ystride, n, N, slab, spc, d_part, d_G, d_tickets);
    } else {
        if ((e = cudaFuncSetAttribute(blu_gram_kernel<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        blu_gram_kernel<NT, false><<<g, BLU_GRAM_WARPS * 32, smem, st>>>(dY, ystride, n, N, slab, spc, d_part, d_G, d_tickets);
    }
    return cudaGetLastError();
}

// Sums of n_out sample matrices.  Y: (n_out blocks of (n, N) row-major, `ystride` doubles apart; host or device.
// sums: n_out x (N*N + N) doubles, host or device.  producer: CUDA stream that wrote a device-resident Y (the
// kernel waits for the work queued on it so far), or NULL.
static int blu_gram_sums(const double *Y, long long n, int N, int n_out, long long ystride, int y_on_device, int telescoped,
                         cudaStream_t producer, double *sums, int sums_on_device, float *kernel_ms, std::string &err)
{
    const int NT = (N + 1 + 7) / 8, NPG = 8 * NT, E = NPG * NPG;
    if (ystride <= 0) ystride = n * N;
    cudaDeviceProp prop; int dev = 0;
    cudaGetDevice(&dev);
    cudaGetDeviceProperties(&prop, dev);
    const long long warps_wanted = (n + 255) / 256;                       // >= 256 samples per warp
    const int ctas_per_sm = 2;                                            // 2-stage rings: two CTAs of 8 warps per SM
    int grid = (int)std::max<long long>(1, std::min<long long>((warps_wanted + BLU_GRAM_WARPS - 1) / BLU_GRAM_WARPS,
                                                                  std::max<long long>(1, (long long)prop.multiProcessorCount * ctas_per_sm / n_out)));
    grid = std::min(grid, BLU_GRAM_GROUP * BLU_GRAM_MAXGROUPS);
    const long long nwarps = (long long)grid * BLU_GRAM_WARPS;
    long long slab = (n + nwarps - 1) / nwarps;
    slab = ((slab + 15) / 16) * 16;
    double *dY = nullptr, *d_part = nullptr, *d_G = nullptr, *d_sums = nullptr;
    unsigned *d_tickets = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, ep = nullptr;
    cudaError_t e = cudaSuccess;
    const size_t nsums = (size_t)n_out * (N * N + N);
    auto done = [&](int code) {
        if (!y_on_device) cudaFree(dY);
        cudaFree(d_part); cudaFree(d_G); cudaFree(d_tickets);
        if (!sums_on_device) cudaFree(d_sums);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (ep) cudaEventDestroy(ep);
        if (st) cudaStreamDestroy(st);
        if (code) err = std::string("pilot sums: ") + cudaGetErrorString(e);
        return code;
    };
    if ((e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return done(BLU_ERR_CUDA);
    if (y_on_device) {
        dY = const_cast<double *>(Y);
        if (producer) {                                   // order behind the stream that produced Y
            if ((e = cudaEventCreateWithFlags(&ep, cudaEventDisableTiming)) != cudaSuccess) return done(BLU_ERR_CUDA);
            if ((e = cudaEventRecord(ep, producer)) != cudaSuccess) return done(BLU_ERR_CUDA);
            if ((e = cudaStreamWaitEvent(st, ep, 0)) != cudaSuccess) return done(BLU_ERR_CUDA);
        }
    } else {
        const size_t tot = (size_t)((n_out - 1) * ystride + n * N);
        if ((e = cudaMalloc(&dY, sizeof(double) * tot)) != cudaSuccess) return done(BLU_ERR_NOMEM);
        if ((e = cudaMemcpyAsync(dY, Y, sizeof(double) * tot, cudaMemcpyHostToDevice, st)) != cudaSuccess) return done(BLU_ERR_CUDA);
    }
    if ((e = cudaMalloc(&d_part, sizeof(double) * E * (size_t)(grid + BLU_GRAM_MAXGROUPS) * n_out)) != cudaSuccess) return done(BLU_ERR_NOMEM);
    if ((e = cudaMalloc(&d_G, sizeof(double) * E * (size_t)n_out)) != cudaSuccess) return done(BLU_ERR_NOMEM);
    if ((e = cudaMalloc(&d_tickets, sizeof(unsigned) * (BLU_GRAM_MAXGROUPS + 1) * (size_t)n_out)) != cudaSuccess) return done(BLU_ERR_NOMEM);
    if (sums_on_device) d_sums = sums;
    else if ((e = cudaMalloc(&d_sums, sizeof(double) * nsums)) != cudaSuccess) return done(BLU_ERR_NOMEM);
    if ((e = cudaMemsetAsync(d_tickets, 0, sizeof(unsigned) * (BLU_GRAM_MAXGROUPS + 1) * (size_t)n_out, st)) != cudaSuccess) return done(BLU_ERR_CUDA);
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    const bool tele = telescoped != 0;
    switch (NT) {
        case 1: e = blu_gram_launch<1>(tele, dY, ystride, n, N, n_out, grid, slab, d_part, d_G, d_tickets, st); break;
        case 2: e = blu_gram_launch<2>(tele, dY, ystride, n, N, n_out, grid, slab, d_part, d_G, d_tickets, st); break;
        case 3: e = blu_gram_launch<3>(tele, dY, ystride, n, N, n_out, grid, slab, d_part, d_G, d_tickets, st); break;
        case 4: e = blu_gram_launch<4>(tele, dY, ystride, n, N, n_out, grid, slab, d_part, d_G, d_tickets, st); break;
        default: e = blu_gram_launch<5>(tele, dY, ystride, n, N, n_out, grid, slab, d_part, d_G, d_tickets, st); break;
    }
    if (e != cudaSuccess) return done(BLU_ERR_CUDA);
    cudaEventRecord(e1, st);
    blu_gram_pack_kernel<<<n_out, 256, 0, st>>>(d_G, NPG, N, d_sums);
    if ((e = cudaGetLastError()) != cudaSuccess) return done(BLU_ERR_CUDA);
    if (!sums_on_device && (e = cudaMemcpyAsync(sums, d_sums, sizeof(double) * nsums, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return done(BLU_ERR_CUDA);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(BLU_ERR_CUDA);
    if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
    return done(BLU_OK);
}

// Host arithmetic on the (all-reduced) sums of ONE output: the reference's outputs from either form.
//   sums = [N*N Gram | N column sums] of Y (telescoped = 0) or of Z (telescoped = 1), n = total number of samples.
// Outputs (each may be NULL): s1 (N) = sumse, S2 (N,N) = sumsc, C_hat (N,N) (blue_models.py:333),
// d1, d2 (N,N): sumsd1[i][j], sumsd2[i][j] for i < j, zero elsewhere (blue_fn.py:147-157),
// dV (N,N): sumsd2/n - (sumsd1/n)^2 for i < j, NaN elsewhere (blue_models.py:339).
static void blu_gram_finalize(const double *sums, long long n, int N, int telescoped, double *s1, double *S2, double *C_hat,
                              double *d1, double *d2, double *dV)
{
    const double *G = sums, *cs = sums + (size_t)N * N;
    std::vector<double> a1(N), A2((size_t)N * N), D1((size_t)N * N, 0.0), D2((size_t)N * N, 0.0);
    if (!telescoped) {
        for (int i = 0; i < N; ++i) a1[i] = cs[i];
        for (int t = 0; t < N * N; ++t) A2[t] = G[t];
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j) {
                D1[(size_t)i * N + j] = a1[i] - a1[j];
                D2[(size_t)i * N + j] = (G[(size_t)i * N + i] - 2.0 * G[(size_t)i * N + j]) + G[(size_t)j * N + j];
            }
    } else {
        // Y_j = Z_0 + ... + Z_j: prefix sums of the column sums, two-dimensional prefix sums of the Gram matrix
        double run = 0.0;
        for (int i = 0; i < N; ++i) { run += cs[i]; a1[i] = run; }
        std::vector<double> R((size_t)N * N);              // R[l][j] = sum_{l' <= j} G[l][l']
        for (int l = 0; l < N; ++l) { double r = 0.0; for (int j = 0; j < N; ++j) { r += G[(size_t)l * N + j]; R[(size_t)l * N + j] = r; } }
        for (int j = 0; j < N; ++j) { double r = 0.0; for (int i = 0; i < N; ++i) { r += R[(size_t)i * N + j]; A2[(size_t)i * N + j] = r; } }
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j) {
                double t1 = 0.0, t2 = 0.0;                  // Y_i - Y_j = -(Z_{i+1} + ... + Z_j)
                for (int l = i + 1; l <= j; ++l) {
                    t1 += cs[l];
                    for (int lp = i + 1; lp <= j; ++lp) t2 += G[(size_t)l * N + lp];
                }
                D1[(size_t)i * N + j] = -t1;
                D2[(size_t)i * N + j] = t2;
            }
    }
    const double dn = (double)n;
    for (int i = 0; i < N; ++i) if (s1) s1[i] = a1[i];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            const size_t t = (size_t)i * N + j;
            if (S2) S2[t] = A2[t];
            if (C_hat) C_hat[t] = A2[t] / dn - (a1[i] * a1[j]) / (dn * dn);
            if (d1) d1[t] = D1[t];
            if (d2) d2[t] = D2[t];
            if (dV) dV[t] = (j > i) ? D2[t] / dn - (D1[t] / dn) * (D1[t] / dn) : nan("");
        }
}

// blu_pilot_covariance (include/bluest_b200.h): one output, plain Gram, host results.
static int blu_gram_run(const double *Y, long long n, int N, int y_on_device, double *s1, double *S2, double *C_hat,
                        float *kernel_ms, std::string &err)
{
    std::vector<double> sums((size_t)N * N + N);
    int rc = blu_gram_sums(Y, n, N, 1, 0, y_on_device, 0, nullptr, sums.data(), 0, kernel_ms, err);
    if (rc) return rc;
    blu_gram_finalize(sums.data(), n, N, 0, s1, S2, C_hat, nullptr, nullptr, nullptr);
    return BLU_OK;
}
