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
//   blu_hv_reduce_kernel<NP> : per-CTA partial sums of t over a row range in a fixed association;
//                              the CTA that finishes last (ticket counter: control flow only) folds the
//                              partials in a fixed order into t[0..31]
//   blu_hv_apply_kernel<NP>  : s = S t, then one dot product per row
// Both are HBM/L2 streams of the U factor (8*NP*L bytes each); no arithmetic through atomics,
// bit-reproducible.
#pragma once
#include "blu_common.cuh"

#define BLU_HV_THREADS 256
#define BLU_HV_UNROLL 4
#define BLU_HV_FOLD 8                 // sub-sums per column when folding the CTA partials
#define BLU_HVA_THREADS 128           // apply kernel: 4 warps x (32 rows x (NP+1)) doubles of shared staging

// t = sum of `nparts` partial vectors (32 doubles each): BLU_HV_FOLD strided sub-sums per column combined
// in a fixed order (warps beyond BLU_HV_FOLD idle, so the association does not depend on the block size).
__device__ __forceinline__ void blu_hv_fold(const double *part, int nparts, double *sh, double *t)
{
    const int tid = threadIdx.x;
    const int c = tid & 31, q = tid >> 5;
    double s = 0.0;
    if (q < BLU_HV_FOLD)
        for (int b = q; b < nparts; b += BLU_HV_FOLD) s += __ldcg(part + (size_t)b * 32 + c);
    if (q < BLU_HV_FOLD) sh[tid] = s;
    __syncthreads();
    if (tid < 32) {
        double a = 0.0;
#pragma unroll
        for (int qq = 0; qq < BLU_HV_FOLD; ++qq) a += sh[qq * 32 + tid];
        t[tid] = a;
    }
    __syncthreads();
}

// part[b*32 + c] = sum over the rows r of CTA b of p[r] * U[r][c]   (c < NP <= 32); t_out = their sum.
// A thread owns two adjacent columns of a fixed row slot (16-byte loads, BLU_HV_UNROLL rows in flight).
template <int NP>
__global__ void __launch_bounds__(BLU_HV_THREADS)
blu_hv_reduce_kernel(const double *__restrict__ U, const double *__restrict__ p, long long lo, long long hi,
                     double *__restrict__ part, unsigned *__restrict__ ticket, double *__restrict__ t_out)
{
    constexpr int HC = NP / 2;                              // column pairs per row
    constexpr int RPP = BLU_HV_THREADS / HC;                // rows per pass of one CTA
    __shared__ double sh[BLU_HV_THREADS * 2];
    __shared__ double tfin[32];
    __shared__ bool last;
    const int tid = threadIdx.x;
    const int r = tid / HC, c2 = tid - r * HC;
    double a0[BLU_HV_UNROLL], a1[BLU_HV_UNROLL];
#pragma unroll
    for (int u = 0; u < BLU_HV_UNROLL; ++u) { a0[u] = 0.0; a1[u] = 0.0; }
    if (r < RPP) {
        const long long stride = (long long)gridDim.x * RPP;
        long long row = lo + (long long)blockIdx.x * RPP + r;
        for (; row + (BLU_HV_UNROLL - 1) * stride < hi; row += BLU_HV_UNROLL * stride) {
            double2 v[BLU_HV_UNROLL]; double pv[BLU_HV_UNROLL];
#pragma unroll
            for (int u = 0; u < BLU_HV_UNROLL; ++u) {
                const long long q = row + u * stride;
                v[u] = __ldg(reinterpret_cast<const double2 *>(U + q * NP) + c2);
                pv[u] = __ldg(p + q);
            }
#pragma unroll
            for (int u = 0; u < BLU_HV_UNROLL; ++u) { a0[u] = fma(pv[u], v[u].x, a0[u]); a1[u] = fma(pv[u], v[u].y, a1[u]); }
        }
        for (; row < hi; row += stride) {
            const double2 v = __ldg(reinterpret_cast<const double2 *>(U + row * NP) + c2);
            const double pv = __ldg(p + row);
            a0[0] = fma(pv, v.x, a0[0]); a1[0] = fma(pv, v.y, a1[0]);
        }
    }
    sh[2 * tid] = (a0[0] + a0[1]) + (a0[2] + a0[3]);
    sh[2 * tid + 1] = (a1[0] + a1[1]) + (a1[2] + a1[3]);
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
    if (tid == 0) *ticket = 0u;                              // ready for the next product
    blu_hv_fold(part, (int)gridDim.x, sh, tfin);
    if (tid < 32) t_out[tid] = tfin[tid];
}

// out[i] = u_i . s for i in [lo,hi), s = S t (S = 2 pinv(Phi), N x N; t: 32 doubles, zero beyond N).
// A warp owns 32 consecutive rows = one contiguous span of 32 NP doubles: coalesced loads, products
// staged in shared memory with row pitch NP + 1, then lane r sums row r.
template <int NP>
__global__ void __launch_bounds__(BLU_HVA_THREADS)
blu_hv_apply_kernel(const double *__restrict__ U, const double *__restrict__ S, int N, const double *__restrict__ t_in,
                    long long lo, long long hi, double *__restrict__ out, int reverse)
{
    __shared__ double t[32];
    __shared__ double prod[BLU_HVA_THREADS / 32][32 * (NP + 1)];
    const int tid = threadIdx.x;
    if (tid < 32) {
        double s = 0.0;
        if (tid < N)
            for (int b = 0; b < N; ++b) s = fma(__ldg(S + tid * N + b), __ldcg(t_in + b), s);
        t[tid] = s;
    }
    __syncthreads();
    const int w = tid >> 5, lane = tid & 31;
    constexpr int NWARP = BLU_HVA_THREADS / 32;
    double *pw = prod[w];
    // reverse: walk the 32-row blocks from the END of the range.  The reduce pass that precedes this kernel streamed U
    // front to back, so the last ~100 MB of it are still in the 126 MB L2: reading backwards turns most of this
    // pass's DRAM traffic into L2 hits when U (168 MB at 20 models) does not fit -- and leaves the FRONT of U in L2
    // for the next product's reduce pass.
    const long long nblk = (hi - lo + 31) / 32;
    for (long long blk = (long long)blockIdx.x * NWARP + w; blk < nblk; blk += (long long)gridDim.x * NWARP) {
        const long long r0 = lo + (reverse ? (nblk - 1 - blk) : blk) * 32;
        const int nrow = (int)((hi - r0) < 32 ? (hi - r0) : 32);
        const double *ub = U + r0 * NP;
        if (nrow == 32) {
#pragma unroll
            for (int it = 0; it < NP; ++it) {
                const int idx = it * 32 + lane;
                pw[idx + idx / NP] = __ldg(ub + idx) * t[idx % NP];
            }
        } else {
            for (int idx = lane; idx < nrow * NP; idx += 32) pw[idx + idx / NP] = __ldg(ub + idx) * t[idx % NP];
        }
        __syncwarp();
        if (lane < nrow) {
            const double *pr = pw + lane * (NP + 1);
            double b0 = 0.0, b1 = 0.0;
#pragma unroll
            for (int c = 0; c < NP; c += 2) { b0 += pr[c]; b1 += pr[c + 1]; }
            out[r0 + lane] = b0 + b1;
        }
        __syncwarp();
    }
}
