// blu_soa.cuh -- lane-per-group gradient / U,V kernels on a group-interleaved copy of the inverses.
//
// Second HBM layout of the packed inverses ("SoA tiles"): the groups of a size class are cut into
// tiles of 32 consecutive groups and, inside a tile, stored entry-major:
//     soa[soff + (t*T + e)*32 + g]  =  packed entry e of group 32 t + g      (zero beyond the class)
// so that the 32 lanes of a warp, one GROUP per lane, read 256 contiguous bytes per packed entry --
// perfectly coalesced, and bank-conflict free once staged in shared memory.  A tile is one contiguous
// span of 32 T doubles; warps stream their tiles through the same two-stage bulk-async ring as the
// other kernels (blu_stream.cuh helpers), 16 entries (4 KB) per stage.
//
// Why a second layout: with one group per lane the quadratic forms need no shuffles, no warp
// reductions and no per-group bookkeeping -- about 4 instructions per 32 group-entries instead of
// ~14 plus ~30 per group in the entry-per-lane form.  The Phi accumulation cannot use it (its
// scatter targets differ from lane to lane and collide), so the AoS copy stays for blu_phi.cuh; each
// kernel streams its own copy once, the HBM traffic per evaluation is unchanged.
#pragma once
#include "blu_common.cuh"
#include "blu_stream.cuh"

#define BLU_SOA_WARPS 8
#define BLU_SOA_E 16                              // packed entries per stage
#define BLU_SOA_STAGE (BLU_SOA_E * 32)            // doubles per stage

struct BluTile {
    int cls;          // class index
    int nsub;         // ceil(T / BLU_SOA_E)
    long long t;      // tile index inside the class (groups 32 t .. 32 t + 31)
};

// AoS (group-major) -> SoA tiles for one class.
__global__ void blu_soa_build_kernel(const double *__restrict__ cinv, long long Lk, int T, double *__restrict__ soa)
{
    const long long ntile = (Lk + 31) / 32;
    const long long total = ntile * T * 32;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(p & 31);
        const long long te = p >> 5;
        const long long t = te / T;
        const int e = (int)(te - t * T);
        const long long i = t * 32 + g;
        soa[p] = i < Lk ? cinv[i * T + e] : 0.0;
    }
}

// Shared-memory carve-up: [stages WARPS x 2 x STAGE][xg WARPS x K x 32][y WARPS x K x 32 (U kernel only)]
//                         [extra][class table][LUT][barriers]
struct BluSoaSmem {
    double *stages, *xg, *y, *extra;
    BluClass *cls;
    unsigned short *lut;
    unsigned long long *bars;
};
__host__ __device__ __forceinline__ size_t blu_soa_smem_bytes(int K, bool withy, int extra_doubles, int ncls, int lutlen)
{
    size_t d = (size_t)BLU_SOA_WARPS * 2 * BLU_SOA_STAGE + (size_t)BLU_SOA_WARPS * K * 32 * (withy ? 2 : 1) + extra_doubles;
    size_t b = sizeof(double) * d + sizeof(BluClass) * ncls + ((sizeof(unsigned short) * lutlen + 7) / 8) * 8;
    return b + sizeof(unsigned long long) * BLU_SOA_WARPS * 2;
}
__device__ __forceinline__ BluSoaSmem blu_soa_carve(unsigned char *raw, int K, bool withy, int extra_doubles, int ncls, int lutlen)
{
    BluSoaSmem s;
    s.stages = reinterpret_cast<double *>(raw);
    s.xg = s.stages + (size_t)BLU_SOA_WARPS * 2 * BLU_SOA_STAGE;
    s.y = s.xg + (size_t)BLU_SOA_WARPS * K * 32;
    s.extra = s.y + (withy ? (size_t)BLU_SOA_WARPS * K * 32 : 0);
    s.cls = reinterpret_cast<BluClass *>(s.extra + extra_doubles);
    s.lut = reinterpret_cast<unsigned short *>(s.cls + ncls);
    s.bars = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(s.lut) + ((sizeof(unsigned short) * lutlen + 7) / 8) * 8);
    return s;
}

// Per-warp producer/consumer cursor over the flattened (tile, sub-chunk) sequence of the warp.
struct BluSoaCursor {
    int q;            // index into the tile list (advances by the number of warps)
    int s;            // sub-chunk inside the tile
    BluTile d;        // descriptor of tile q
};

// V rows from U rows: v_i = S u_i with S = 2 pinv(Phi) (N x N, symmetric), for rows [lo,hi).  One thread
// per (row, column): a warp covers 32 consecutive doubles of V (coalesced), S lives in shared memory,
// the U row is read through L1.  8 NP L bytes in, 8 NP L bytes out.
__global__ void __launch_bounds__(256)
blu_v_from_u_kernel(const double *__restrict__ U, const double *__restrict__ S, int N, int NP, long long lo, long long hi,
                    double *__restrict__ V)
{
    __shared__ double sS[BLU_MAX_MODELS_C * BLU_MAX_MODELS_C];
    for (int t = threadIdx.x; t < N * N; t += blockDim.x) sS[t] = S[t];
    __syncthreads();
    const long long total = (hi - lo) * NP;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const long long r = q / NP;
        const int a = (int)(q - r * NP);
        const double *ur = U + (lo + r) * NP;
        double va = 0.0;
        if (a < N) {
#pragma unroll 4
            for (int b = 0; b < N; ++b) va = fma(sS[a * N + b], __ldg(ur + b), va);
        }
        V[(lo + r) * NP + a] = va;
    }
}

// WITHU = false: gradient only.  WITHU = true: gradient + U rows.
template <bool WITHU>
__global__ void __launch_bounds__(BLU_SOA_WARPS * 32)
blu_grad_soa_kernel(const BluClass *__restrict__ cls, int ncls, int N, int NP, int K,
                    const BluTile *__restrict__ tiles, int ntiles, const double *__restrict__ soa,
                    const long long *__restrict__ soff, const unsigned short *__restrict__ lut, int lutlen,
                    const unsigned *__restrict__ gmask, const double *__restrict__ xrow,
                    long long lo, long long hi, double *__restrict__ grad, double *__restrict__ U)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const BluSoaSmem sm = blu_soa_carve(smraw, K, WITHU, N, ncls, lutlen);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // block prologue
    for (int t = threadIdx.x; t < ncls; t += blockDim.x) sm.cls[t] = cls[t];
    for (int t = threadIdx.x; t < lutlen; t += blockDim.x) sm.lut[t] = lut[t];
    double *sx = sm.extra;                         // x (N)
    for (int t = threadIdx.x; t < N; t += blockDim.x) sx[t] = xrow[t];
    double *stage[2] = {sm.stages + (size_t)(2 * w) * BLU_SOA_STAGE, sm.stages + (size_t)(2 * w + 1) * BLU_SOA_STAGE};
    unsigned long long *bar[2] = {sm.bars + 2 * w, sm.bars + 2 * w + 1};
    double *xg = sm.xg + (size_t)w * K * 32;       // xg[j*32 + lane] = x[id_j(group of lane)]
    double *yv = sm.y + (size_t)w * K * 32;        // y[j*32 + lane]  (U kernel)
    if (lane == 0) { blu_mbar_init(bar[0], 1); blu_mbar_init(bar[1], 1); blu_mbar_fence_init(); }
    __syncthreads();

    const int gw = blockIdx.x * BLU_SOA_WARPS + w;
    const int nw = gridDim.x * BLU_SOA_WARPS;
    if (gw >= ntiles) return;

    // ---- producer side ------------------------------------------------------------------------
    BluSoaCursor pr;
    pr.q = gw; pr.s = 0; pr.d = tiles[gw];
    unsigned mask_next = 0u;
    auto issue = [&](int st) {                     // copy sub-chunk (pr.q, pr.s) into stage st, advance the cursor
        const BluClass ci = sm.cls[pr.d.cls];
        const int e0 = pr.s * BLU_SOA_E;
        const int cnt = (ci.T - e0) < BLU_SOA_E ? (ci.T - e0) : BLU_SOA_E;
        if (lane == 0) {
            const double *src = soa + soff[pr.d.cls] + ((long long)pr.d.t * ci.T + e0) * 32;
            blu_mbar_expect_tx(bar[st], (unsigned)(cnt * 256));
            blu_bulk_g2s(stage[st], src, (unsigned)(cnt * 256), bar[st]);
        }
        if (pr.s == 0) {                           // first sub-chunk of a tile: fetch the lane's membership mask too
            const long long gi = ci.goff + pr.d.t * 32 + lane;
            mask_next = (pr.d.t * 32 + lane < ci.Lk) ? gmask[gi] : 0u;
        }
        if (++pr.s >= pr.d.nsub) {
            pr.s = 0; pr.q += nw;
            if (pr.q < ntiles) pr.d = tiles[pr.q];
        }
    };
    issue(0);
    // ---- consumer side ------------------------------------------------------------------------
    BluTile cd = tiles[gw];
    int cq = gw, cs = 0;
    unsigned mask = 0u;
    double acc = 0.0;
    int k = 0, T = 0;
    const unsigned short *lt = sm.lut;
    for (int it = 0; cq < ntiles; ++it) {
        const int st = it & 1;
        const unsigned mask_this = mask_next;      // valid when cs == 0 (set by the issue of this very sub-chunk)
        if (pr.q < ntiles) issue(st ^ 1);
        if (cs == 0) {                             // ---- tile start ----
            const BluClass ci = sm.cls[cd.cls];
            k = ci.k; T = ci.T; lt = sm.lut + ci.lutoff;
            mask = mask_this;
            unsigned mk = mask;
            for (int j = 0; j < k; ++j) {
                const int b = __ffs(mk) - 1;
                xg[j * 32 + lane] = sx[b < 0 ? 0 : b];
                if (WITHU) yv[j * 32 + lane] = 0.0;
                mk &= mk - 1u;
            }
            acc = 0.0;
            __syncwarp();
        }
        blu_mbar_wait(bar[st], (unsigned)((it >> 1) & 1));
        const double *sp = stage[st];
        const int e0 = cs * BLU_SOA_E;
        const int cnt = (T - e0) < BLU_SOA_E ? (T - e0) : BLU_SOA_E;
        if (!WITHU) {
#pragma unroll 4
            for (int e = 0; e < cnt; ++e) {
                const unsigned jl = lt[e0 + e];
                const int j = jl >> 8, l = jl & 255u;
                const double c = sp[e * 32 + lane];
                const double xj = xg[j * 32 + lane], xl = xg[l * 32 + lane];
                acc += ((j == l) ? 1.0 : 2.0) * (xj * c * xl);
            }
        } else {
            for (int e = 0; e < cnt; ++e) {
                const unsigned jl = lt[e0 + e];
                const int j = jl >> 8, l = jl & 255u;
                const double c = sp[e * 32 + lane];
                const double xj = xg[j * 32 + lane], xl = xg[l * 32 + lane];
                yv[j * 32 + lane] = fma(c, xl, yv[j * 32 + lane]);
                if (j != l) yv[l * 32 + lane] = fma(c, xj, yv[l * 32 + lane]);
            }
        }
        __syncwarp();                              // stage consumed before it is refilled
        if (++cs >= cd.nsub) {                     // ---- tile end ----
            const BluClass ci = sm.cls[cd.cls];
            const long long gi = ci.goff + cd.t * 32 + lane;          // flat group index of this lane
            const bool live = (cd.t * 32 + lane < ci.Lk) && gi >= lo && gi < hi;
            if (!WITHU) {
                if (live) grad[gi] = -acc;
            } else {
                double gsum = 0.0;
                for (int j = 0; j < k; ++j) gsum = fma(xg[j * 32 + lane], yv[j * 32 + lane], gsum);
                if (live) grad[gi] = -gsum;
                // U rows: y scattered to model slots.  The tile's 32 groups are consecutive flat indices, so
                // their rows are ONE contiguous span of 32 NP doubles: element idx = r NP + a is produced by
                // lane idx & 31 from row r's mask (shuffle) and y (shared), and every store instruction
                // writes 256 contiguous bytes.  (V = U S is a separate pass over U: blu_v_from_u_kernel.)
                __syncwarp();
                const long long g0 = ci.goff + cd.t * 32;                   // flat index of the tile's first group
                const int nrow = (int)((ci.Lk - cd.t * 32) < 32 ? (ci.Lk - cd.t * 32) : 32);
                double *ub = U + g0 * NP;
                for (int idx = lane; idx < 32 * NP; idx += 32) {
                    const int r = idx / NP, a = idx - r * NP;
                    const unsigned mr = __shfl_sync(BLU_FULL, mask, r);
                    double ua = 0.0;
                    if ((mr >> a) & 1u) ua = yv[__popc(mr & ((1u << a) - 1u)) * 32 + r];
                    const long long gr = g0 + r;
                    if (r < nrow && gr >= lo && gr < hi) ub[idx] = ua;
                }
            }
            cs = 0; cq += nw;
            if (cq < ntiles) cd = tiles[cq];
            __syncwarp();
        }
    }
}
