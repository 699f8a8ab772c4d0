// blu_matvec.cuh -- the Hessian of misc.py:497-503 applied to a vector without forming it.
//
// The reference fills hess (L,L) block by block (hessKQ_c, cmisc.cpp:74-97) and doubles it with
// `hess += hess.T`; with u_i = R_i^T Cinv_i R_i x and S = 2 pinv(Phi) (u_i: the rows of the U factor
// the gradient pass leaves in HBM) that matrix is H[i][j] = u_i . S u_j, so
//
//     (H p)_i = u_i . (S t),      t = sum_j p_j u_j
//                                                    (two passes over the L x NP doubles of U; V is not needed)
// which is all scipy's trust-constr (projected CG, sap.py:410) or a truncated-Newton step ever asks
// of the Hessian.  At N = 15 the dense matrix is 8.59 GB (151 ms of PCIe per evaluation); the two
// factor passes read 8.4 MB out of L2.  At N = 20 the dense matrix (8.8 TB) does not exist and this
// is the only form of the Hessian.
//
//   blu_hv_reduce_kernel<NP,UN,MINB> : per-CTA partial sums of t over a row range in a fixed association (UN 16-byte loads
//                              in flight per thread, MINB resident CTAs per SM, ONE wave); the CTA that finishes last
//                              (ticket counter: control flow only) folds the partials in a fixed order into t[0..31]
//   blu_hv_apply_kernel<NP,WARPS,MINB> : s = S t, then one dot product per row (16-byte loads, partial products parked in
//                              shared memory with an odd row pitch)
// Both are HBM/L2 streams of the U factor (8*NP*L bytes each); no arithmetic through atomics,
// bit-reproducible.  20 models (2 x 168 MB): 58.7 us per product = 0.92 of the measured copy peak (the apply pass walks U
// backwards and finds the tail of the reduce pass in L2); the first generation (kept in tools/lab/hv_lab.cu for the
// comparison) took 89.3 us.
#pragma once
#include "blu_common.cuh"

#define BLU_HV_THREADS 256
#define BLU_HV_FOLD 8                 // sub-sums per column when folding the CTA partials

// What the first generation lost (profiles/r01c_stream_kernels_N20.md: 0.51 / 0.69 of the copy peak at 20 models):
// reduce -- 1184 CTAs at 3 resident per SM = 2.67 waves, and a last-CTA fold of 1184 partial vectors with one load in flight per
// thread; apply -- twice the load and shared-memory instructions a 16-byte mapping needs.

// Fold with 8 loads in flight per thread (fixed association: BLU_HV_FOLD strided sub-sums per column, combined in
// order; warps beyond BLU_HV_FOLD idle, so it does not depend on the block size).
__device__ __forceinline__ void blu_hv_fold(const double *part, int nparts, double *sh, double *t)
{
    const int tid = threadIdx.x;
    const int c = tid & 31, q = tid >> 5;
    double s = 0.0;
    if (q < BLU_HV_FOLD) {
        for (int b = q; b < nparts; b += 8 * BLU_HV_FOLD) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (b + u * BLU_HV_FOLD < nparts) ? __ldcg(part + (size_t)(b + u * BLU_HV_FOLD) * 32 + c) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        sh[tid] = s;
    }
    __syncthreads();
    if (tid < 32) {
        double a = 0.0;
#pragma unroll
        for (int qq = 0; qq < BLU_HV_FOLD; ++qq) a += sh[qq * 32 + tid];
        t[tid] = a;
    }
    __syncthreads();
}

template <int NP, int UN, int MINB>
__global__ void __launch_bounds__(BLU_HV_THREADS, MINB)
blu_hv_reduce_kernel(const double *__restrict__ U, const double *__restrict__ p, long long lo, long long hi,
                      double *__restrict__ part, unsigned *__restrict__ ticket, double *__restrict__ t_out)
{
    constexpr int HC = NP / 2;
    constexpr int RPP = BLU_HV_THREADS / HC;
    __shared__ double sh[BLU_HV_THREADS * 2];
    __shared__ double tfin[32];
    __shared__ bool last;
    const int tid = threadIdx.x;
    const int r = tid / HC, c2 = tid - r * HC;
    double a0[UN], a1[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) { a0[u] = 0.0; a1[u] = 0.0; }
    if (r < RPP) {
        // the CTAs sweep U front to back together: the order the backwards-walking apply pass relies on for its L2 hits
        // (a contiguous span per CTA measured 2 us slower per product)
        const long long b1 = hi;
        const long long stride = (long long)gridDim.x * RPP;
        long long row = lo + (long long)blockIdx.x * RPP + r;
        for (; row + (UN - 1) * stride < b1; row += UN * stride) {
            double2 v[UN]; double pv[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const long long q = row + u * stride;
                v[u] = __ldg(reinterpret_cast<const double2 *>(U + q * NP) + c2);
                pv[u] = __ldg(p + q);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) { a0[u] = fma(pv[u], v[u].x, a0[u]); a1[u] = fma(pv[u], v[u].y, a1[u]); }
        }
        for (; row < b1; row += stride) {
            const double2 v = __ldg(reinterpret_cast<const double2 *>(U + row * NP) + c2);
            const double pv = __ldg(p + row);
            a0[0] = fma(pv, v.x, a0[0]); a1[0] = fma(pv, v.y, a1[0]);
        }
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int u = 0; u < UN; ++u) { s0 += a0[u]; s1 += a1[u]; }
    sh[2 * tid] = s0;
    sh[2 * tid + 1] = s1;
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        if (tid < NP)
            for (int q = 0; q < RPP; ++q) s += sh[2 * (q * HC + (tid >> 1)) + (tid & 1)];
        part[(size_t)blockIdx.x * 32 + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (tid == 0) *ticket = 0u;
    blu_hv_fold(part, (int)gridDim.x, sh, tfin);
    if (tid < 32) t_out[tid] = tfin[tid];
}

// Apply pass with 16-byte loads: a warp owns 32 consecutive rows = 32 HC double2; lane l loads the double2 with flat index
// it*32 + l (fully coalesced 512-byte requests), multiplies by the matching pair of s = S t, parks the partial product in
// shared memory (row pitch HC + 1: conflict free) and lane r adds the HC partials of row r.
template <int NP, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
blu_hv_apply_kernel(const double *__restrict__ U, const double *__restrict__ S, int N, const double *__restrict__ t_in,
                     long long lo, long long hi, double *__restrict__ out, int reverse)
{
    constexpr int HC = NP / 2;
    __shared__ __align__(16) double t[32];
    __shared__ double prod[WARPS][32 * (HC + 1)];
    const int tid = threadIdx.x;
    if (tid < 32) {
        double s = 0.0;
        if (tid < N)
            for (int b = 0; b < N; ++b) s = fma(__ldg(S + tid * N + b), __ldcg(t_in + b), s);
        t[tid] = s;
    }
    __syncthreads();
    const int w = tid >> 5, lane = tid & 31;
    double *pw = prod[w];
    const double2 *t2 = reinterpret_cast<const double2 *>(t);
    const long long nblk = (hi - lo + 31) / 32;
    for (long long blk = (long long)blockIdx.x * WARPS + w; blk < nblk; blk += (long long)gridDim.x * WARPS) {
        const long long r0 = lo + (reverse ? (nblk - 1 - blk) : blk) * 32;
        const int nrow = (int)((hi - r0) < 32 ? (hi - r0) : 32);
        const double2 *ub = reinterpret_cast<const double2 *>(U + r0 * NP);
        if (nrow == 32) {
            double2 v[HC];
#pragma unroll
            for (int it = 0; it < HC; ++it) v[it] = __ldg(ub + it * 32 + lane);
#pragma unroll
            for (int it = 0; it < HC; ++it) {
                const int idx = it * 32 + lane;
                const double2 sv = t2[idx % HC];
                pw[idx + idx / HC] = fma(v[it].x, sv.x, v[it].y * sv.y);
            }
        } else {
            for (int idx = lane; idx < nrow * HC; idx += 32) {
                const double2 v = __ldg(ub + idx);
                const double2 sv = t2[idx % HC];
                pw[idx + idx / HC] = fma(v.x, sv.x, v.y * sv.y);
            }
        }
        __syncwarp();
        if (lane < nrow) {
            const double *pr = pw + lane * (HC + 1);
            double b0 = 0.0, b1 = 0.0;
#pragma unroll
            for (int c = 0; c < HC; c += 2) { b0 += pr[c]; b1 += pr[c + 1]; }
            out[r0 + lane] = b0 + b1;
        }
        __syncwarp();
    }
}
