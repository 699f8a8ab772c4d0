// tools/lab/phi_lab.cu -- lab harness for the lane-per-group Phi accumulation on the SoA tiles
// (round 2; not part of the product).  Variants:
//   stream   pure read of the tile stream (ceiling of direct LDG.64 streaming at this occupancy)
//   phi<R>   lane = group, per-warp accumulators acc[target][R slots] in shared memory (conflict free:
//            lane -> slot lane % R, the 32/R lanes sharing a slot take turns), values by direct LDG.64
//            with UB loads in flight per lane, optional bulk L2 prefetch D tiles ahead
// Checked against an atomicAdd reference.  N = 20, all 2^20 - 1 groups by default.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o phi_lab tools/lab/phi_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cmath>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e_=(x); if(e_!=cudaSuccess){printf("CUDA error %s at line %d\n",cudaGetErrorString(e_),__LINE__); exit(1);} }while(0)

struct Cls { int k, T; long long Lk, goff, soff; };
struct Tile { int cls; int pad; long long t; };

__device__ __forceinline__ double lds64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void l2_prefetch(const void *p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------------
template <int UB>
__global__ void __launch_bounds__(256)
stream_kernel(const Cls *__restrict__ cls, const Tile *__restrict__ tiles, int ntiles, const double *__restrict__ soa, double *out, int D)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwc = blockDim.x >> 5;
    const int gw = blockIdx.x * nwc + w, nw = gridDim.x * nwc;
    double s = 0.0;
    for (int q = gw; q < ntiles; q += nw) {
        const Tile tl = tiles[q];
        const Cls ci = cls[tl.cls];
        if (D > 0 && lane == 0 && q + D * nw < ntiles) {
            const Tile tp = tiles[q + D * nw];
            const Cls cp = cls[tp.cls];
            l2_prefetch(soa + cp.soff + tp.t * cp.T * 32, (unsigned)(cp.T * 256));
        }
        const double *src = soa + ci.soff + tl.t * ci.T * 32 + lane;
        for (int e0 = 0; e0 < ci.T; e0 += UB) {
            double v[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) v[u] = (e0 + u < ci.T) ? src[(long long)(e0 + u) * 32] : 0.0;
#pragma unroll
            for (int u = 0; u < UB; ++u) s += v[u];
        }
    }
    if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------------------------------------
// lane = group.  acc[(tgt * R + slot)] per warp; upper-triangle target index tgt(a,b) = a N - a(a-1)/2 + (b-a).
template <int R, int UB>
__global__ void __launch_bounds__(512)
phi_kernel(const Cls *__restrict__ cls, int N, const Tile *__restrict__ tiles, int ntiles, const double *__restrict__ soa,
           const unsigned *__restrict__ gmask, const double *__restrict__ m, double *__restrict__ part, int D)
{
    constexpr int P = 32 / R;
    extern __shared__ double sacc[];
    const int NT = N * (N + 1) / 2;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwc = blockDim.x >> 5;
    double *acc = sacc + (size_t)w * NT * R;
    for (int t = lane; t < NT * R; t += 32) acc[t] = 0.0;
    __syncwarp();
    const int slot = lane % R, phase = lane / R;
    const unsigned acc_s = (unsigned)__cvta_generic_to_shared(acc) + 8u * slot;
    const int gw = blockIdx.x * nwc + w, nw = gridDim.x * nwc;
    for (int q = gw; q < ntiles; q += nw) {
        const Tile tl = tiles[q];
        const Cls ci = cls[tl.cls];
        if (D > 0 && lane == 0 && q + D * nw < ntiles) {
            const Tile tp = tiles[q + D * nw];
            const Cls cp = cls[tp.cls];
            l2_prefetch(soa + cp.soff + tp.t * cp.T * 32, (unsigned)(cp.T * 256));
        }
        const long long gi = tl.t * 32 + lane;
        const bool ok = gi < ci.Lk;
        const unsigned mask = ok ? gmask[ci.goff + gi] : ((1u << ci.k) - 1u);
        const double mv = ok ? m[ci.goff + gi] : 0.0;
        if (!__ballot_sync(0xffffffffu, mv != 0.0)) continue;
        const double *src = soa + ci.soff + tl.t * ci.T * 32 + lane;
        const int T = ci.T, k = ci.k;
        // (j,l) walk: row j, columns l = j..k-1; mj = members from j on, ml = members from l on
        unsigned mj = mask, ml = mask;
        int a = __ffs(mj) - 1;
        int rowoff = a * N - a * (a - 1) / 2 - a;
        int left = k;                                    // entries left in the current row (warp uniform)
        for (int e0 = 0; e0 < T; e0 += UB) {
            double v[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) v[u] = (e0 + u < T) ? src[(long long)(e0 + u) * 32] : 0.0;
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                if (e0 + u < T) {
                    const int b = __ffs(ml) - 1;
                    ml &= ml - 1u;
                    const unsigned at = acc_s + 8u * (unsigned)((rowoff + b) * R);
#pragma unroll
                    for (int ph = 0; ph < P; ++ph) {
                        if (P == 1 || phase == ph) sts64(at, fma(mv, v[u], lds64(at)));
                        if (P > 1) __syncwarp();
                    }
                    if (--left == 0) {
                        mj &= mj - 1u; ml = mj;
                        a = __ffs(mj) - 1; if (a < 0) a = 0;
                        rowoff = a * N - a * (a - 1) / 2 - a;
                        left = __popc(mj);
                    }
                }
            }
        }
    }
    __syncwarp();
    __syncthreads();
    // fixed-order reduce: slots, then warps -> CTA partial in (N,N) upper-triangle layout
    for (int t = threadIdx.x; t < NT; t += blockDim.x) {
        double s = 0.0;
        for (int ww = 0; ww < nwc; ++ww) {
            const double *aw = sacc + (size_t)ww * NT * R + (size_t)t * R;
            double sw = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) sw += aw[r];
            s += sw;
        }
        part[(long long)blockIdx.x * NT + t] = s;
    }
}


// ---------------------------------------------------------------------------------------------
// Second form: entries walked in REVERSE packed order (one FLO per entry peels the mask from the top), in
// branch-free batches of UB entries: all UB accumulator addresses first, then per phase UB independent LDS,
// UB DFMA, UB STS (targets inside a group are distinct: no hazard inside a batch).  Loads of the next batch
// are issued before the current one is consumed.
template <int UB>
struct PhiBatch { double v[UB]; };

template <int UB>
__device__ __forceinline__ void phi_load(PhiBatch<UB> &b, const double *__restrict__ src, int e1)
{
#pragma unroll
    for (int u = 0; u < UB; ++u) {
        const int e = e1 - 1 - u;
        const double x = src[(long long)(e < 0 ? 0 : e) * 32];
        b.v[u] = e < 0 ? 0.0 : x;
    }
}

template <int R, int UB, int WARPS, int ABL = 0>      // ABL 1: no global loads of the values; 2: no shared RMW
__global__ void __launch_bounds__(WARPS * 32)
phi_kernel2(const Cls *__restrict__ cls, int N, const Tile *__restrict__ tiles, int ntiles, const double *__restrict__ soa,
            const unsigned *__restrict__ gmask, const double *__restrict__ m, double *__restrict__ part, int D)
{
    constexpr int P = 32 / R;
    extern __shared__ double sacc[];
    const int NT = N * (N + 1) / 2;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwc = blockDim.x >> 5;
    double *acc = sacc + (size_t)w * (NT + 1) * R;       // + one scratch target for the padding steps of the last batch
    for (int t = lane; t < (NT + 1) * R; t += 32) acc[t] = 0.0;
    __syncwarp();
    const int slot = lane % R, phase = lane / R;
    const unsigned acc_s = (unsigned)__cvta_generic_to_shared(acc) + 8u * slot;
    const int gw = blockIdx.x * nwc + w, nw = gridDim.x * nwc;
    double dummy = 0.0;
    for (int q = gw; q < ntiles; q += nw) {
        const Tile tl = tiles[q];
        const Cls ci = cls[tl.cls];
        const long long gi = tl.t * 32 + lane;
        const bool ok = gi < ci.Lk;
        const unsigned mask = ok ? gmask[ci.goff + gi] : ((1u << ci.k) - 1u);
        const double mv = ok ? m[ci.goff + gi] : 0.0;
        if (!__ballot_sync(0xffffffffu, mv != 0.0)) continue;
        const double *src = soa + ci.soff + tl.t * ci.T * 32 + lane;
        const int T = ci.T, k = ci.k;
        // reverse walk: row j = k-1 .. 0, inside a row l = k-1 .. j
        int a = 31 - __clz(mask);
        unsigned mrow = 1u << a;                         // members at positions >= j
        unsigned ml = mrow;                              // members of the row not yet consumed
        int rowoff = a * N - a * (a - 1) / 2 - a;
        int left = 1, rowlen = 1;                        // entries left in the row / length of the row (warp uniform)
        PhiBatch<UB> cur, nxt;
        if (ABL == 1) { for (int u = 0; u < UB; ++u) { cur.v[u] = mv + u; nxt.v[u] = mv - u; } }
        else phi_load<UB>(cur, src, T);
        for (int e1 = T; e1 > 0; e1 -= UB) {
            if (ABL != 1 && e1 - UB > 0) phi_load<UB>(nxt, src, e1 - UB);
            unsigned at[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int b = 31 - __clz(ml | 1u);
                ml &= ~(1u << b);
                at[u] = acc_s + 8u * (unsigned)((rowoff + b) * R);
                if (--left == 0) {                       // next row (one position lower); uniform
                    const unsigned lower = mask & ((1u << a) - 1u);
                    if (rowlen < k) {
                        a = 31 - __clz(lower);
                        mrow |= 1u << a; ml = mrow;
                        rowoff = a * N - a * (a - 1) / 2 - a;
                        left = ++rowlen;
                    } else { ml = 0u; left = 1 << 30; rowoff = NT; }
                }
            }
#pragma unroll
            if (ABL == 2) {
#pragma unroll
                for (int u = 0; u < UB; ++u) dummy += cur.v[u] * (double)at[u];
            } else
#pragma unroll
            for (int ph = 0; ph < P; ++ph) {
                if (P == 1 || phase == ph) {
                    double o[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) o[u] = lds64(at[u]);
#pragma unroll
                    for (int u = 0; u < UB; ++u) o[u] = fma(mv, cur.v[u], o[u]);
#pragma unroll
                    for (int u = 0; u < UB; ++u) sts64(at[u], o[u]);
                }
                if (P > 1) __syncwarp();
            }
            cur = nxt;
        }
    }
    __syncwarp();
    __syncthreads();
    for (int t = threadIdx.x; t < NT; t += blockDim.x) {
        double s = 0.0;
        for (int ww = 0; ww < nwc; ++ww) {
            const double *aw = sacc + (size_t)ww * (NT + 1) * R + (size_t)t * R;
            double sw = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) sw += aw[r];
            s += sw;
        }
        part[(long long)blockIdx.x * NT + t] = s;
    }
    if (ABL == 2 && dummy == 1.2345) part[0] = dummy;
}

// ---------------------------------------------------------------------------------------------
// Fourth form: the accumulator target of every entry is PRECOMPUTED (one byte per entry, same tile layout
// as the values): no mask walk, no per-class code.  Per 32 entries: LDG.64 value, LDG.U8 target, LEA, RMW.
template <int R, int UB, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
phi_kernel4(const Cls *__restrict__ cls, int N, const Tile *__restrict__ tiles, int ntiles, const double *__restrict__ soa,
            const unsigned char *__restrict__ tgt, const double *__restrict__ m, double *__restrict__ part)
{
    constexpr int P = 32 / R;
    extern __shared__ double sacc[];
    const int NT = N * (N + 1) / 2;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *acc = sacc + (size_t)w * NT * R;
    for (int t = lane; t < NT * R; t += 32) acc[t] = 0.0;
    __syncwarp();
    const int slot = lane % R, phase = lane / R;
    const unsigned acc_s = (unsigned)__cvta_generic_to_shared(acc) + 8u * slot;
    // contiguous run of tiles per warp
    const int gw = blockIdx.x * WARPS + w, nw = gridDim.x * WARPS;
    const int q0 = (int)((long long)ntiles * gw / nw), q1 = (int)((long long)ntiles * (gw + 1) / nw);
    for (int q = q0; q < q1; ++q) {
        const Tile tl = tiles[q];
        const Cls ci = cls[tl.cls];
        const long long gi = tl.t * 32 + lane;
        const double mv = gi < ci.Lk ? m[ci.goff + gi] : 0.0;
        if (!__ballot_sync(0xffffffffu, mv != 0.0)) continue;
        const long long row0 = ci.soff / 32 + tl.t * ci.T;             // first row of the tile
        const double *src = soa + row0 * 32 + lane;
        const unsigned char *tsrc = tgt + row0 * 32 + lane;
        const int T = ci.T;
        double va[UB], vb[UB]; unsigned ta[UB], tb[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) { const int e = u < T ? u : T - 1; va[u] = src[e * 32]; ta[u] = tsrc[e * 32]; }
        for (int e0 = 0; e0 < T; e0 += 2 * UB) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int eb = e0 + half * UB;                          // batch being consumed
                if (eb >= T) break;
                double *cv = half ? vb : va; unsigned *ct = half ? tb : ta;
                double *nv = half ? va : vb; unsigned *nt = half ? ta : tb;
                if (eb + UB < T) {
#pragma unroll
                    for (int u = 0; u < UB; ++u) { int e = eb + UB + u; e = e < T ? e : T - 1; nv[u] = src[e * 32]; nt[u] = tsrc[e * 32]; }
                }
                unsigned at[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u) at[u] = acc_s + 8u * R * ct[u];
#pragma unroll
                for (int ph = 0; ph < P; ++ph) {
                    if (P == 1 || phase == ph) {
                        double o[UB];
#pragma unroll
                        for (int u = 0; u < UB; ++u) if (eb + u < T) o[u] = lds64(at[u]);
#pragma unroll
                        for (int u = 0; u < UB; ++u) if (eb + u < T) o[u] = fma(mv, cv[u], o[u]);
#pragma unroll
                        for (int u = 0; u < UB; ++u) if (eb + u < T) sts64(at[u], o[u]);
                    }
                    if (P > 1) __syncwarp();
                }
            }
        }
    }
    __syncwarp();
    __syncthreads();
    for (int t = threadIdx.x; t < NT; t += blockDim.x) {
        double s = 0.0;
        for (int ww = 0; ww < WARPS; ++ww) {
            const double *aw = sacc + (size_t)ww * NT * R + (size_t)t * R;
            double sw = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) sw += aw[r];
            s += sw;
        }
        part[(long long)blockIdx.x * NT + t] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Fifth form: as the fourth, but round-robin tiles (balanced), exact batches (no clamping arithmetic),
// L1-bypassing streaming loads, tail entries one at a time.
__device__ __forceinline__ double ldg_stream_f64(const double *p)
{
    double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}
__device__ __forceinline__ unsigned ldg_stream_u8(const unsigned char *p)
{
    unsigned v; asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
template <int R, int UB, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
phi_kernel5(const Cls *__restrict__ cls, int N, const Tile *__restrict__ tiles, int ntiles, const double *__restrict__ soa,
            const unsigned char *__restrict__ tgt, const double *__restrict__ m, double *__restrict__ part)
{
    constexpr int P = 32 / R;
    extern __shared__ double sacc[];
    const int NT = N * (N + 1) / 2;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *acc = sacc + (size_t)w * NT * R;
    for (int t = lane; t < NT * R; t += 32) acc[t] = 0.0;
    __syncwarp();
    const int slot = lane % R, phase = lane / R;
    const unsigned acc_s = (unsigned)__cvta_generic_to_shared(acc) + 8u * slot;
    const int gw = blockIdx.x * WARPS + w, nw = gridDim.x * WARPS;
    for (int q = gw; q < ntiles; q += nw) {
        const Tile tl = tiles[q];
        const Cls ci = cls[tl.cls];
        const long long gi = tl.t * 32 + lane;
        const double mv = gi < ci.Lk ? m[ci.goff + gi] : 0.0;
        if (!__ballot_sync(0xffffffffu, mv != 0.0)) continue;
        const long long row0 = ci.soff / 32 + tl.t * ci.T;
        const double *src = soa + row0 * 32 + lane;
        const unsigned char *tsrc = tgt + row0 * 32 + lane;
        const int nb = ci.T / UB;
        double va[UB], vb[UB]; unsigned ta[UB], tb[UB];
        if (nb > 0) {
#pragma unroll
            for (int u = 0; u < UB; ++u) { va[u] = ldg_stream_f64(src + u * 32); ta[u] = ldg_stream_u8(tsrc + u * 32); }
        }
        for (int b = 0; b < nb; b += 2) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (b + half >= nb) break;
                double *cv = half ? vb : va; unsigned *ct = half ? tb : ta;
                double *nv = half ? va : vb; unsigned *nt = half ? ta : tb;
                src += UB * 32; tsrc += UB * 32;
                if (b + half + 1 < nb) {
#pragma unroll
                    for (int u = 0; u < UB; ++u) { nv[u] = ldg_stream_f64(src + u * 32); nt[u] = ldg_stream_u8(tsrc + u * 32); }
                }
                unsigned at[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u) at[u] = acc_s + 8u * R * ct[u];
#pragma unroll
                for (int ph = 0; ph < P; ++ph) {
                    if (P == 1 || phase == ph) {
                        double o[UB];
#pragma unroll
                        for (int u = 0; u < UB; ++u) o[u] = lds64(at[u]);
#pragma unroll
                        for (int u = 0; u < UB; ++u) o[u] = fma(mv, cv[u], o[u]);
#pragma unroll
                        for (int u = 0; u < UB; ++u) sts64(at[u], o[u]);
                    }
                    if (P > 1) __syncwarp();
                }
            }
        }
        // tail: T - nb*UB entries (src, tsrc point at them)
        const int rem = ci.T - nb * UB;
        for (int e = 0; e < rem; ++e) {
            const double v = ldg_stream_f64(src + e * 32);
            const unsigned at1 = acc_s + 8u * R * ldg_stream_u8(tsrc + e * 32);
#pragma unroll
            for (int ph = 0; ph < P; ++ph) {
                if (P == 1 || phase == ph) sts64(at1, fma(mv, v, lds64(at1)));
                if (P > 1) __syncwarp();
            }
        }
    }
    __syncwarp();
    __syncthreads();
    for (int t = threadIdx.x; t < NT; t += blockDim.x) {
        double s = 0.0;
        for (int ww = 0; ww < WARPS; ++ww) {
            const double *aw = sacc + (size_t)ww * NT * R + (size_t)t * R;
            double sw = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) sw += aw[r];
            s += sw;
        }
        part[(long long)blockIdx.x * NT + t] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Sixth form: precomputed byte targets (as the fourth/fifth), but the rows are STAGED by the bulk copy engine into a
// deep per-warp ring (bytes in flight independent of registers and warps), few warps with R = 16 private slots.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned long long *bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{ asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nW6:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D6;\nbra W6;\nD6:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int R, int UB, int WARPS, int NS>
__global__ void __launch_bounds__(WARPS * 32)
phi_kernel6(const Cls *__restrict__ cls, int N, const Tile *__restrict__ tiles, int ntiles, const double *__restrict__ soa,
            const unsigned char *__restrict__ tgt, const double *__restrict__ m, double *__restrict__ part)
{
    constexpr int P = 32 / R;
    constexpr int STAGE = UB * 288;                       // UB rows: 256 B of values + 32 B of targets each
    extern __shared__ __align__(128) unsigned char s6[];
    const int NT = N * (N + 1) / 2;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *acc = reinterpret_cast<double *>(s6) + (size_t)w * NT * R;
    unsigned char *ring = s6 + sizeof(double) * (size_t)WARPS * NT * R + (size_t)w * NS * STAGE;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(s6 + sizeof(double) * (size_t)WARPS * NT * R + (size_t)WARPS * NS * STAGE) + w * NS;
    for (int t = lane; t < NT * R; t += 32) acc[t] = 0.0;
    if (lane == 0) { for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const int slot = lane % R, phase = lane / R;
    const unsigned acc_s = smem_u32(acc) + 8u * slot;
    const int gw = blockIdx.x * WARPS + w, nw = gridDim.x * WARPS;
    // producer cursor over (tile, sub-chunk of UB rows)
    int pq = gw, ps = 0;                                  // tile list index, sub-chunk
    Tile pt = pq < ntiles ? tiles[pq] : Tile{0, 0, 0};
    int pT = pq < ntiles ? cls[pt.cls].T : 0;
    long long prow0 = pq < ntiles ? cls[pt.cls].soff / 32 + pt.t * pT : 0;
    auto issue = [&](int st) {
        const int e0 = ps * UB;
        const int cnt = (pT - e0) < UB ? (pT - e0) : UB;
        if (lane == 0) {
            mbar_expect(&bars[st], (unsigned)(cnt * 288));
            bulk_g2s(ring + (size_t)st * STAGE, soa + (prow0 + e0) * 32, (unsigned)(cnt * 256), &bars[st]);
            bulk_g2s(ring + (size_t)st * STAGE + UB * 256, tgt + (prow0 + e0) * 32, (unsigned)(cnt * 32), &bars[st]);
        }
        if ((ps + 1) * UB >= pT) {
            ps = 0; pq += nw;
            if (pq < ntiles) { pt = tiles[pq]; pT = cls[pt.cls].T; prow0 = cls[pt.cls].soff / 32 + pt.t * pT; }
        } else ++ps;
    };
    int nissued = 0;
    for (; nissued < NS - 1 && pq < ntiles; ++nissued) issue(nissued);
    int it = 0;
    for (int q = gw; q < ntiles; q += nw) {
        const Tile tl = tiles[q];
        const Cls ci = cls[tl.cls];
        const long long gi = tl.t * 32 + lane;
        const double mv = gi < ci.Lk ? m[ci.goff + gi] : 0.0;
        const int T = ci.T;
        for (int e0 = 0; e0 < T; e0 += UB, ++it) {
            const int st = it % NS;
            if (pq < ntiles) issue((it + NS - 1) % NS);
            mbar_wait(&bars[st], (unsigned)((it / NS) & 1));
            const int cnt = (T - e0) < UB ? (T - e0) : UB;
            const unsigned vs = smem_u32(ring + (size_t)st * STAGE) + 8u * lane;
            const unsigned ts = smem_u32(ring + (size_t)st * STAGE + UB * 256) + lane;
            unsigned at[UB]; double v[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                unsigned tg;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(tg) : "r"(ts + 32u * u) : "memory");
                v[u] = lds64(vs + 256u * u);
                at[u] = acc_s + 8u * R * tg;
            }
#pragma unroll
            for (int ph = 0; ph < P; ++ph) {
                if (P == 1 || phase == ph) {
                    double o[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) if (u < cnt) o[u] = lds64(at[u]);
#pragma unroll
                    for (int u = 0; u < UB; ++u) if (u < cnt) o[u] = fma(mv, v[u], o[u]);
#pragma unroll
                    for (int u = 0; u < UB; ++u) if (u < cnt) sts64(at[u], o[u]);
                }
                if (P > 1) __syncwarp();
            }
            __syncwarp();
        }
    }
    __syncwarp();
    __syncthreads();
    const double *sacc6 = reinterpret_cast<const double *>(s6);
    for (int t = threadIdx.x; t < NT; t += blockDim.x) {
        double s = 0.0;
        for (int ww = 0; ww < WARPS; ++ww) {
            const double *aw = sacc6 + (size_t)ww * NT * R + (size_t)t * R;
            double sw = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) sw += aw[r];
            s += sw;
        }
        part[(long long)blockIdx.x * NT + t] = s;
    }
}

__global__ void ref_kernel(const Cls *cls, int ncls, int N, const double *soa, const unsigned *gmask, const double *m, double *phi)
{
    for (int ic = 0; ic < ncls; ++ic) {
        const Cls ci = cls[ic];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ci.Lk; i += (long long)gridDim.x * blockDim.x) {
            const unsigned mask = gmask[ci.goff + i];
            const double mv = m[ci.goff + i];
            int ids[32]; int c = 0;
            for (int b = 0; b < 32; ++b) if (mask >> b & 1u) ids[c++] = b;
            const long long t = i / 32; const int g = (int)(i % 32);
            int e = 0;
            for (int j = 0; j < ci.k; ++j)
                for (int l = j; l < ci.k; ++l, ++e) {
                    const double v = soa[ci.soff + (t * ci.T + e) * 32 + g];
                    const int a = ids[j], b = ids[l];
                    atomicAdd(&phi[a * N - a * (a - 1) / 2 + (b - a)], mv * v);
                }
        }
    }
}

__global__ void fold_kernel(const double *part, int nparts, int NT, double *out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NT) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(long long)p * NT + t];
    out[t] = s;
}

static void next_comb(std::vector<int> &c, int N)
{
    const int k = (int)c.size();
    int i = k - 1;
    while (i >= 0 && c[i] == N - k + i) --i;
    if (i < 0) return;
    ++c[i];
    for (int j = i + 1; j < k; ++j) c[j] = c[j - 1] + 1;
}

template <typename F>
static float timeit(F f, int reps)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps * 1e3f;
}

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 20;
    const int Kmax = argc > 2 ? atoi(argv[2]) : N;
    std::vector<Cls> cls; std::vector<unsigned> gmask; std::vector<Tile> tiles;
    long long goff = 0, soff = 0;
    for (int k = 1; k <= Kmax; ++k) {
        Cls c; c.k = k; c.T = k * (k + 1) / 2; c.goff = goff; c.soff = soff;
        std::vector<int> comb(k); for (int j = 0; j < k; ++j) comb[j] = j;
        long long Lk = 1; for (int j = 0; j < k; ++j) Lk = Lk * (N - j) / (j + 1);
        c.Lk = Lk;
        for (long long i = 0; i < Lk; ++i) { unsigned mk = 0; for (int v : comb) mk |= 1u << v; gmask.push_back(mk); next_comb(comb, N); }
        const long long nt = (Lk + 31) / 32;
        for (long long t = 0; t < nt; ++t) { Tile tl; tl.cls = (int)cls.size(); tl.pad = 0; tl.t = t; tiles.push_back(tl); }
        goff += Lk; soff += nt * 32 * c.T;
        cls.push_back(c);
    }
    const long long L = goff, S = soff;
    const int NT = N * (N + 1) / 2;
    printf("N=%d K=%d L=%lld tiles=%zu soa=%.1f MB\n", N, Kmax, L, tiles.size(), 8.0 * S / 1e6);
    std::vector<double> hsoa((size_t)S), hm((size_t)L);
    srand(1);
    for (auto &v : hsoa) v = rand() / (double)RAND_MAX - 0.3;
    for (auto &v : hm) v = 1.0 + 10.0 * rand() / (double)RAND_MAX;
    Cls *d_cls; Tile *d_tiles; unsigned *d_mask; double *d_soa, *d_m, *d_part, *d_phi, *d_ref, *d_out;
    CK(cudaMalloc(&d_cls, sizeof(Cls) * cls.size())); CK(cudaMemcpy(d_cls, cls.data(), sizeof(Cls) * cls.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_tiles, sizeof(Tile) * tiles.size())); CK(cudaMemcpy(d_tiles, tiles.data(), sizeof(Tile) * tiles.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_mask, 4 * L)); CK(cudaMemcpy(d_mask, gmask.data(), 4 * L, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_soa, 8 * S)); CK(cudaMemcpy(d_soa, hsoa.data(), 8 * S, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_m, 8 * L)); CK(cudaMemcpy(d_m, hm.data(), 8 * L, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_part, 8 * (size_t)NT * 1024)); CK(cudaMalloc(&d_phi, 8 * NT)); CK(cudaMalloc(&d_ref, 8 * NT)); CK(cudaMalloc(&d_out, 64));
    CK(cudaMemset(d_ref, 0, 8 * NT));
    ref_kernel<<<296, 256>>>(d_cls, (int)cls.size(), N, d_soa, d_mask, d_m, d_ref);
    CK(cudaDeviceSynchronize());
    std::vector<double> href(NT), hphi(NT);
    CK(cudaMemcpy(href.data(), d_ref, 8 * NT, cudaMemcpyDeviceToHost));
    double refmax = 0; for (double v : href) refmax = std::max(refmax, fabs(v));
    int nsm = 148; { cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); nsm = p.multiProcessorCount; }
    const int ntiles = (int)tiles.size();
    const double bytes = 8.0 * S;

    auto report = [&](const char *name, float us) { printf("%-44s %8.1f us  %7.1f GB/s\n", name, us, bytes / us / 1e3); fflush(stdout); };
    // ---- stream ceilings
    for (int cps : {1, 2, 4})
        for (int D : {0}) {
            char nm[128]; snprintf(nm, sizeof nm, "stream UB=16 ctas/sm=%d warps=8 D=%d", cps, D);
            report(nm, timeit([&] { stream_kernel<16><<<nsm * cps, 256>>>(d_cls, d_tiles, ntiles, d_soa, d_out, D); }, 20));
        }
    report("stream UB=8 ctas/sm=1 D=4", timeit([&] { stream_kernel<8><<<nsm, 256>>>(d_cls, d_tiles, ntiles, d_soa, d_out, 4); }, 20));
    report("stream UB=8 ctas/sm=1 D=16", timeit([&] { stream_kernel<8><<<nsm, 256>>>(d_cls, d_tiles, ntiles, d_soa, d_out, 16); }, 20));

    // ---- phi variants
    auto run_phi = [&](auto kern, int R, int warps, int D, const char *tag) {
        const size_t smem = 8ull * warps * (NT + 1) * R;
        if (smem > 227 * 1024) { printf("%-44s skipped (smem %zu)\n", tag, smem); return; }
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = nsm;
        auto go = [&] { kern<<<grid, warps * 32, smem>>>(d_cls, N, d_tiles, ntiles, d_soa, d_mask, d_m, d_part, D); };
        const float us = timeit(go, 20);
        CK(cudaGetLastError());
        fold_kernel<<<(NT + 127) / 128, 128>>>(d_part, grid, NT, d_phi);
        CK(cudaMemcpy(hphi.data(), d_phi, 8 * NT, cudaMemcpyDeviceToHost));
        double err = 0; for (int t = 0; t < NT; ++t) err = std::max(err, fabs(hphi[t] - href[t]));
        char nm[160]; snprintf(nm, sizeof nm, "%s R=%d warps=%d D=%d (err %.1e)", tag, R, warps, D, err / refmax);
        report(nm, us);
    };
    // target table (one byte per entry, tile layout)
    std::vector<unsigned char> htgt((size_t)S, 0);
    for (size_t ic = 0; ic < cls.size(); ++ic) {
        const Cls &c = cls[ic];
        for (long long i = 0; i < c.Lk; ++i) {
            const unsigned mk = gmask[(size_t)(c.goff + i)];
            int ids[32], n = 0; for (int b = 0; b < 32; ++b) if (mk >> b & 1u) ids[n++] = b;
            const long long t = i / 32; const int g = (int)(i % 32);
            int e = 0;
            for (int j = 0; j < c.k; ++j) for (int l = j; l < c.k; ++l, ++e)
                htgt[(size_t)(c.soff + (t * c.T + e) * 32 + g)] = (unsigned char)(ids[j] * N - ids[j] * (ids[j] - 1) / 2 + (ids[l] - ids[j]));
        }
    }
    unsigned char *d_tgt; CK(cudaMalloc(&d_tgt, S)); CK(cudaMemcpy(d_tgt, htgt.data(), S, cudaMemcpyHostToDevice));
    const int only = argc > 3 ? atoi(argv[3]) : -1;
    int vidx = 0;
    auto run4 = [&](auto kern, int R, int warps, int cps, const char *tag) {
        if (only >= 0 && vidx++ != only) return;
        const size_t smem = 8ull * warps * NT * R;
        if (smem * cps > 227 * 1024) { printf("%-44s skipped (smem %zu)\n", tag, smem); return; }
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = nsm * cps;
        auto go = [&] { kern<<<grid, warps * 32, smem>>>(d_cls, N, d_tiles, ntiles, d_soa, d_tgt, d_m, d_part); };
        const float us = timeit(go, 20);
        CK(cudaGetLastError());
        fold_kernel<<<(NT + 127) / 128, 128>>>(d_part, grid, NT, d_phi);
        CK(cudaMemcpy(hphi.data(), d_phi, 8 * NT, cudaMemcpyDeviceToHost));
        double err = 0; for (int t = 0; t < NT; ++t) err = std::max(err, fabs(hphi[t] - href[t]));
        char nm[160]; snprintf(nm, sizeof nm, "%s R=%d warps=%d ctas/sm=%d (err %.1e)", tag, R, warps, cps, err / refmax);
        report(nm, us);
    };
    run4(phi_kernel5<8, 8, 8>, 8, 8, 2, "phi5 UB=8");
    auto run6 = [&](auto kern, int R, int warps, int UB, int NS, const char *tag) {
        const size_t smem = 8ull * warps * NT * R + (size_t)warps * NS * UB * 288 + 8ull * warps * NS;
        if (smem > 227 * 1024) { printf("%-44s skipped (smem %zu)\n", tag, smem); return; }
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = nsm;
        auto go = [&] { kern<<<grid, warps * 32, smem>>>(d_cls, N, d_tiles, ntiles, d_soa, d_tgt, d_m, d_part); };
        const float us = timeit(go, 20);
        CK(cudaGetLastError());
        fold_kernel<<<(NT + 127) / 128, 128>>>(d_part, grid, NT, d_phi);
        CK(cudaMemcpy(hphi.data(), d_phi, 8 * NT, cudaMemcpyDeviceToHost));
        double err = 0; for (int t = 0; t < NT; ++t) err = std::max(err, fabs(hphi[t] - href[t]));
        char nm[160]; snprintf(nm, sizeof nm, "%s R=%d warps=%d UB=%d NS=%d smem=%zuK (err %.1e)", tag, R, warps, UB, NS, smem >> 10, err / refmax);
        report(nm, us);
    };
    run6(phi_kernel6<16, 16, 4, 6>, 16, 4, 16, 6, "phi6");
    run6(phi_kernel6<16, 8, 4, 12>, 16, 4, 8, 12, "phi6");
    run6(phi_kernel6<16, 16, 5, 3>, 16, 5, 16, 3, "phi6");
    run6(phi_kernel6<8, 16, 8, 3>, 8, 8, 16, 3, "phi6");
    run6(phi_kernel6<8, 8, 8, 6>, 8, 8, 8, 6, "phi6");
    run6(phi_kernel6<8, 16, 6, 5>, 8, 6, 16, 5, "phi6");
    run6(phi_kernel6<32, 16, 2, 8>, 32, 2, 16, 8, "phi6");
    run6(phi_kernel6<4, 8, 16, 3>, 4, 16, 8, 3, "phi6");
    return 0;
}
