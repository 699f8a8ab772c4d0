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
// reductions and no per-group bookkeeping, and -- the entry loop being unrolled per group size -- the
// lane's x and y vectors stay in registers: one shared load and one or two FMAs per 32 group-entries
// instead of ~14 instructions plus ~30 per group in the entry-per-lane form.  The Phi accumulation cannot use it (its
// scatter targets differ from lane to lane and collide), so the AoS copy stays for blu_phi.cuh; each
// kernel streams its own copy once, the HBM traffic per evaluation is unchanged.
#pragma once
#include "blu_common.cuh"
#include "blu_stream.cuh"

#include "blu_soa_types.h"

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

// Shared-memory carve-up: [stages WARPS x NS x STAGE][y WARPS x K x 32 (U kernel only)][x (N)][class table][barriers]
struct BluSoaSmem {
    double *stages, *y, *extra;
    BluClass *cls;
    unsigned long long *bars;
};
__host__ __device__ __forceinline__ size_t blu_soa_smem_bytes(int K, bool withy, int extra_doubles, int ncls, int ns)
{
    size_t d = (size_t)BLU_SOA_WARPS * ns * BLU_SOA_STAGE + (withy ? (size_t)BLU_SOA_WARPS * K * 32 : 0) + extra_doubles;
    return sizeof(double) * d + sizeof(BluClass) * ncls + sizeof(unsigned long long) * BLU_SOA_WARPS * ns;
}
__device__ __forceinline__ BluSoaSmem blu_soa_carve(unsigned char *raw, int K, bool withy, int extra_doubles, int ncls, int ns)
{
    BluSoaSmem s;
    s.stages = reinterpret_cast<double *>(raw);
    s.y = s.stages + (size_t)BLU_SOA_WARPS * ns * BLU_SOA_STAGE;
    s.extra = s.y + (withy ? (size_t)BLU_SOA_WARPS * K * 32 : 0);
    s.cls = reinterpret_cast<BluClass *>(s.extra + extra_doubles);
    s.bars = reinterpret_cast<unsigned long long *>(s.cls + ncls);
    return s;
}

// V rows from U rows: v_i = S u_i with S = 2 pinv(Phi) (N x N, symmetric), for rows [lo,hi).  One thread
// per (row, column): a warp covers 32 consecutive doubles of V (coalesced), S lives in shared memory,
// the U row is read through L1.  Only the dense Hessian (N <= 17) and its row panels read V.
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

// Per-warp pipeline state: producer cursor over the flattened (tile, sub-chunk) sequence of the warp,
// an NS-deep ring of stages with their barriers, and the consumer's sub-chunk count.  The producer
// runs NS-1 sub-chunks (4 KB each) ahead of the consumer: with two stages a warp had one 4 KB copy in
// flight, ~64 KB per SM -- about what one SM's share of HBM bandwidth needs to cover ONE memory
// latency, so the stream stalled on every hand-off.
template <int NS>
struct BluSoaPipe {
    const BluTile *tiles;
    const double *soa;
    const long long *soff;
    const BluClass *cls;          // class table in shared memory
    double *stage0;               // NS stages of BLU_SOA_STAGE doubles
    unsigned long long *bar0;     // NS barriers
    int ntiles, nw;
    int q, s;                     // producer: tile list index / sub-chunk inside the tile
    BluTile d;                    // descriptor of tile q
    int it;                       // consumer: sub-chunks consumed so far

    // copy sub-chunk (q, s) into stage st, advance the cursor
    __device__ __forceinline__ void issue(int st, int lane)
    {
        const BluClass ci = cls[d.cls];
        const int e0 = s * BLU_SOA_E;
        const int cnt = (ci.T - e0) < BLU_SOA_E ? (ci.T - e0) : BLU_SOA_E;
        if (lane == 0) {
            const double *src = soa + soff[d.cls] + ((long long)d.t * ci.T + e0) * 32;
            blu_mbar_expect_tx(bar0 + st, (unsigned)(cnt * 256));
            blu_bulk_g2s(stage0 + (size_t)st * BLU_SOA_STAGE, src, (unsigned)(cnt * 256), bar0 + st);
        }
        if (++s >= d.nsub) {
            s = 0; q += nw;
            if (q < ntiles) d = tiles[q];
        }
    }
    __device__ __forceinline__ void prime(int lane)
    {
        it = 0;
        for (int st = 0; st < NS - 1 && q < ntiles; ++st) issue(st, lane);
    }
    // hand the consumer the next staged sub-chunk.  The sub-chunk consumed before this call is done with
    // (the caller synchronised the warp), so its stage -- (it-1) mod NS -- is refilled first.
    __device__ __forceinline__ const double *next(int lane)
    {
        const int st = it % NS;
        if (q < ntiles) issue((it + NS - 1) % NS, lane);
        blu_mbar_wait(bar0 + st, (unsigned)((it / NS) & 1));
        ++it;
        return stage0 + (size_t)st * BLU_SOA_STAGE;
    }
};

// A warp's RUN of tiles of one size class (32 groups of size K per tile, one group per lane) with
// everything in REGISTERS: the (j,l) sequence of the packed entries is the same for every lane, so with
// the entry loop fully unrolled for a fixed K the lane's x_j and y_j are statically indexed register
// arrays and a packed entry costs one shared load (the staged value) and one or two FMAs -- no index
// table, no x/y traffic through shared memory.  Stage hand-offs happen at compile-time-known entry
// counts (every BLU_SOA_E entries).
//
// NOT inlined into the kernel: 2 x 32 unrolled routines in one function body make ptxas take ten
// minutes; as separate functions they compile in well under one.  One call covers all consecutive tiles
// of the warp's sequence that belong to the same class (thousands, for the big classes), so the call
// and the copy of the pipeline state into registers are paid once per run, not once per tile.
struct BluSoaCursor {
    int cq;               // index of the current tile in the tile list
    BluTile cd;           // its descriptor
    unsigned mask;        // the lane's membership mask in that tile
};

template <int K, bool WITHU, int NS>
__device__ __noinline__ void blu_soa_run(BluSoaPipe<NS> &pipe, BluSoaCursor &cur, const unsigned *__restrict__ gmask,
                                         const double *__restrict__ sx, double *__restrict__ yv, int NP, long long lo, long long hi,
                                         double *__restrict__ grad, double *__restrict__ U, int lane)
{
    // these generic pointers all point into shared memory: say so, or the staged values are read with
    // generic LD.E.64 instead of LDS
    __builtin_assume(__isShared(sx));
    __builtin_assume(__isShared(yv));
    BluSoaPipe<NS> p = pipe;
    int cq = cur.cq;
    BluTile cd = cur.cd;
    unsigned mask = cur.mask;
    const int cls0 = cd.cls;
    const BluClass ci = p.cls[cls0];
    for (;;) {
        // descriptor and membership mask of the warp's next tile, one tile ahead (an L2 hit hidden behind a whole tile)
        BluTile cn = cd; unsigned mask_n = 0u;
        if (cq + p.nw < p.ntiles) {
            cn = p.tiles[cq + p.nw];
            const BluClass cin = p.cls[cn.cls];
            mask_n = (cn.t * 32 + lane < cin.Lk) ? gmask[cin.goff + cn.t * 32 + lane] : 0u;
        }
        double x[K], y[WITHU ? K : 1];
        {
            unsigned mk = mask;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int b = __ffs(mk) - 1;
                x[j] = sx[b < 0 ? 0 : b];
                if (WITHU) y[j] = 0.0;
                mk &= mk - 1u;
            }
        }
        double acc = 0.0;
        const double *sp = nullptr;
        int e = 0;                                     // compile-time after unrolling
#pragma unroll
        for (int j = 0; j < K; ++j) {
#pragma unroll
            for (int l = j; l < K; ++l) {
                if (e % BLU_SOA_E == 0) {
                    if (e > 0) __syncwarp();           // previous stage fully consumed before it is refilled
                    sp = p.next(lane);
                    __builtin_assume(__isShared(sp));
                }
                const double c = sp[(e % BLU_SOA_E) * 32 + lane];
                if (!WITHU) {
                    acc += ((j == l) ? 1.0 : 2.0) * (x[j] * c * x[l]);
                } else {
                    y[j] = fma(c, x[l], y[j]);
                    if (j != l) y[l] = fma(c, x[j], y[l]);
                }
                ++e;
            }
        }
        __syncwarp();
        const long long gi = ci.goff + cd.t * 32 + lane;          // flat group index of this lane
        const bool live = (cd.t * 32 + lane < ci.Lk) && gi >= lo && gi < hi;
        if (!WITHU) {
            if (live) grad[gi] = -acc;
        } else {
            double gsum = 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) { gsum = fma(x[j], y[j], gsum); yv[j * 32 + lane] = y[j]; }
            if (live) grad[gi] = -gsum;
            __syncwarp();
            // U rows: y scattered to model slots.  The tile's 32 groups are consecutive flat indices, so their
            // rows are ONE contiguous span of 32 NP doubles: element idx = r NP + a is produced by lane idx & 31
            // from row r's mask (shuffle) and y (shared); every store instruction writes 256 contiguous bytes.
            const long long g0 = ci.goff + cd.t * 32;
            const int nrow = (int)((ci.Lk - cd.t * 32) < 32 ? (ci.Lk - cd.t * 32) : 32);
            double *ub = U + g0 * NP;
            const int dr = 32 / NP, da = 32 - dr * NP;
            int r = lane / NP, a = lane - r * NP;
            for (int idx = lane; idx < 32 * NP; idx += 32) {
                const unsigned mr = __shfl_sync(BLU_FULL, mask, r);
                double ua = 0.0;
                if ((mr >> a) & 1u) ua = yv[__popc(mr & ((1u << a) - 1u)) * 32 + r];
                const long long gr = g0 + r;
                if (r < nrow && gr >= lo && gr < hi) ub[idx] = ua;
                r += dr; a += da;
                if (a >= NP) { a -= NP; ++r; }
            }
            __syncwarp();
        }
        cq += p.nw; cd = cn; mask = mask_n;
        if (cq >= p.ntiles || cd.cls != cls0) break;
    }
    pipe = p;
    cur.cq = cq; cur.cd = cd; cur.mask = mask;
}

// WITHU = false: gradient only.  WITHU = true: gradient + U rows (V = U S is a separate pass over U,
// blu_v_from_u_kernel, run only where the dense Hessian needs it).
#define BLU_SOA_NS_GRAD 3          // two CTAs per SM: 16 warps x 2 copies in flight
#define BLU_SOA_NS_U 4             // one CTA per SM (registers): 8 warps x 3 copies in flight
template <bool WITHU>
__global__ void __launch_bounds__(BLU_SOA_WARPS * 32)
blu_grad_soa_kernel(const BluClass *__restrict__ cls, int ncls, int N, int NP, int K,
                    const BluTile *__restrict__ tiles, int ntiles, const double *__restrict__ soa,
                    const long long *__restrict__ soff, const unsigned *__restrict__ gmask, const double *__restrict__ xrow,
                    long long lo, long long hi, double *__restrict__ grad, double *__restrict__ U)
{
    constexpr int NS = WITHU ? BLU_SOA_NS_U : BLU_SOA_NS_GRAD;
    extern __shared__ __align__(16) unsigned char smraw[];
    const BluSoaSmem sm = blu_soa_carve(smraw, K, WITHU, N, ncls, NS);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < ncls; t += blockDim.x) sm.cls[t] = cls[t];
    double *sx = sm.extra;                         // x (N)
    for (int t = threadIdx.x; t < N; t += blockDim.x) sx[t] = xrow[t];
    BluSoaPipe<NS> p;
    p.tiles = tiles; p.soa = soa; p.soff = soff; p.cls = sm.cls;
    p.stage0 = sm.stages + (size_t)(NS * w) * BLU_SOA_STAGE;
    p.bar0 = sm.bars + NS * w;
    p.ntiles = ntiles; p.nw = gridDim.x * BLU_SOA_WARPS;
    double *yv = sm.y + (size_t)w * K * 32;        // y[j*32 + lane], staging for the U rows (WITHU only)
    if (lane == 0) {
        for (int st = 0; st < NS; ++st) blu_mbar_init(p.bar0 + st, 1);
        blu_mbar_fence_init();
    }
    __syncthreads();

    const int gw = blockIdx.x * BLU_SOA_WARPS + w;
    if (gw >= ntiles) return;
    p.q = gw; p.s = 0; p.d = tiles[gw];
    p.prime(lane);
    BluSoaCursor cur;
    cur.cq = gw; cur.cd = tiles[gw];
    {
        const BluClass ci = sm.cls[cur.cd.cls];
        cur.mask = (cur.cd.t * 32 + lane < ci.Lk) ? gmask[ci.goff + cur.cd.t * 32 + lane] : 0u;
    }
    while (cur.cq < ntiles) {
        switch (sm.cls[cur.cd.cls].k) {
#define BLU_SOA_CASE(KK) case KK: blu_soa_run<KK, WITHU, NS>(p, cur, gmask, sx, yv, NP, lo, hi, grad, U, lane); break;
            BLU_SOA_CASE(1) BLU_SOA_CASE(2) BLU_SOA_CASE(3) BLU_SOA_CASE(4) BLU_SOA_CASE(5) BLU_SOA_CASE(6) BLU_SOA_CASE(7) BLU_SOA_CASE(8)
            BLU_SOA_CASE(9) BLU_SOA_CASE(10) BLU_SOA_CASE(11) BLU_SOA_CASE(12) BLU_SOA_CASE(13) BLU_SOA_CASE(14) BLU_SOA_CASE(15) BLU_SOA_CASE(16)
            BLU_SOA_CASE(17) BLU_SOA_CASE(18) BLU_SOA_CASE(19) BLU_SOA_CASE(20) BLU_SOA_CASE(21) BLU_SOA_CASE(22) BLU_SOA_CASE(23) BLU_SOA_CASE(24)
            BLU_SOA_CASE(25) BLU_SOA_CASE(26) BLU_SOA_CASE(27) BLU_SOA_CASE(28) BLU_SOA_CASE(29) BLU_SOA_CASE(30) BLU_SOA_CASE(31) BLU_SOA_CASE(32)
#undef BLU_SOA_CASE
            default: return;                           // unreachable: 1 <= k <= 32
        }
    }
}
