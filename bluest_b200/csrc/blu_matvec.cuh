// blu_matvec.cuh -- the Hessian of misc.py:497-503 applied to a vector without forming it.
//
// The reference fills hess (L,L) block by block (hessKQ_c, cmisc.cpp:74-97) and doubles it with
// `hess += hess.T`; with u_i = R_i^T Cinv_i R_i x and v_i = 2 pinv(Phi) u_i (the rows of the U and V
// factors the gradient pass leaves in HBM) that matrix is H[i][j] = v_i . u_j, so
//
//     (H p)_i = v_i . t = u_i . (S t),      t = sum_j p_j u_j,  S = 2 pinv(Phi)
//                                                    (two passes over the L x NP doubles of U; V is not needed)
//
// which is all scipy's trust-constr (projected CG, sap.py:410) or a truncated-Newton step ever asks
// of the Hessian.  At N = 15 the dense matrix is 8.59 GB (151 ms of PCIe per evaluation); the two
// factor passes read 8.4 MB out of L2.  At N = 20 the dense matrix (8.8 TB) does not exist and this
// is the only form of the Hessian.
//
//   blu_hv_reduce_kernel : per-CTA partial sums of t over a row range, fixed association
//   blu_hv_apply_kernel  : fixed-order sum of the partials, s = S t, then one dot product per row
// Both are HBM/L2 streams of one factor (8*NP*L bytes each); no atomics, bit-reproducible.
#pragma once
#include "blu_common.cuh"

#define BLU_HV_THREADS 256
#define BLU_HV_UNROLL 4

// part[b*32 + c] = sum over the rows r of CTA b of p[r] * U[r][c]      (c < NP <= 32)
__global__ void __launch_bounds__(BLU_HV_THREADS)
blu_hv_reduce_kernel(const double *__restrict__ U, const double *__restrict__ p, long long lo, long long hi, int NP,
                     double *__restrict__ part)
{
    __shared__ double sh[BLU_HV_THREADS];
    const int rpp = BLU_HV_THREADS / NP;                    // rows per pass of one CTA
    const int tid = threadIdx.x;
    const int r = tid / NP, c = tid - r * NP;
    double acc[BLU_HV_UNROLL];
#pragma unroll
    for (int u = 0; u < BLU_HV_UNROLL; ++u) acc[u] = 0.0;
    if (r < rpp) {
        const long long stride = (long long)gridDim.x * rpp;
        long long row = lo + (long long)blockIdx.x * rpp + r;
        for (; row + (BLU_HV_UNROLL - 1) * stride < hi; row += BLU_HV_UNROLL * stride) {
#pragma unroll
            for (int u = 0; u < BLU_HV_UNROLL; ++u) {
                const long long q = row + u * stride;
                acc[u] = fma(__ldg(p + q), __ldg(U + q * NP + c), acc[u]);
            }
        }
        for (; row < hi; row += stride) acc[0] = fma(__ldg(p + row), __ldg(U + row * NP + c), acc[0]);
    }
    sh[tid] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        if (tid < NP)
            for (int q = 0; q < rpp; ++q) s += sh[q * NP + tid];
        part[(size_t)blockIdx.x * 32 + tid] = s;
    }
}

// t = sum of `nparts` partial vectors (32 doubles each): 8 strided sub-sums per column combined in a
// fixed order, so every CTA of every run forms the identical t.
__device__ __forceinline__ void blu_hv_fold(const double *__restrict__ part, int nparts, double *sh, double *t)
{
    const int tid = threadIdx.x;
    const int c = tid & 31, q = tid >> 5;
    double s = 0.0;
    for (int b = q; b < nparts; b += BLU_HV_THREADS / 32) s += __ldg(part + (size_t)b * 32 + c);
    sh[tid] = s;
    __syncthreads();
    if (tid < 32) {
        double a = 0.0;
#pragma unroll
        for (int qq = 0; qq < BLU_HV_THREADS / 32; ++qq) a += sh[qq * 32 + tid];
        t[tid] = a;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(BLU_HV_THREADS)
blu_hv_fold_kernel(const double *__restrict__ part, int nparts, double *__restrict__ t_out)
{
    __shared__ double sh[BLU_HV_THREADS];
    __shared__ double t[32];
    blu_hv_fold(part, nparts, sh, t);
    if (threadIdx.x < 32) t_out[threadIdx.x] = t[threadIdx.x];
}

// out[i] = u_i . s for i in [lo,hi), s = S t (S = 2 pinv(Phi), N x N), t folded from `nparts` partial
// vectors: H p = U S U^T p needs the U factor only.
__global__ void __launch_bounds__(BLU_HV_THREADS)
blu_hv_apply_kernel(const double *__restrict__ U, const double *__restrict__ S, int N, const double *__restrict__ part, int nparts,
                    long long lo, long long hi, int NP, double *__restrict__ out)
{
    __shared__ double sh[BLU_HV_THREADS];
    __shared__ double tt[32];
    __shared__ double t[32];
    const int tid = threadIdx.x;
    blu_hv_fold(part, nparts, sh, tt);
    if (tid < 32) {
        double s = 0.0;
        if (tid < N)
            for (int b = 0; b < N; ++b) s = fma(__ldg(S + tid * N + b), tt[b], s);
        t[tid] = s;
    }
    __syncthreads();
    const int nq = NP >> 1;
    for (long long i = lo + (long long)blockIdx.x * BLU_HV_THREADS + tid; i < hi; i += (long long)gridDim.x * BLU_HV_THREADS) {
        const double2 *row = reinterpret_cast<const double2 *>(U + i * NP);
        double a0 = 0.0, a1 = 0.0;
        for (int q = 0; q < nq; ++q) {
            const double2 v = __ldg(row + q);
            a0 = fma(v.x, t[2 * q], a0);
            a1 = fma(v.y, t[2 * q + 1], a1);
        }
        out[i] = a0 + a1;
    }
}
