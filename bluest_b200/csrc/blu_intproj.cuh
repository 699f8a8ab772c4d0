// blu_intproj.cuh -- "next" row f2: the batched candidate evaluation inside the integer projection
// (best_closest_integer_solution_BLUE, misc.py:313-382):
//     phis = basephi + psi[:, idx] @ ms            (n_cand matrices N x N, misc.py:368)
//     Vs   = pinv(phis, hermitian=True, rcond=1e-10)[:, 0, 0]                (misc.py:369)
// where idx are the LL <= 24 groups whose sample count is rounded down or up and each column of ms
// is one floor/ceil combination.  One warp evaluates 32 consecutive candidates, one after the other:
// the 32 coefficient columns are loaded coalesced (lane = candidate), the LL dense Psi_t = R_t^T
// Cinv_t R_t matrices sit in shared memory (expanded once per CTA from the packed inverses), lane r
// assembles row r of Phi_c, and the (0,0) entry of the pseudo-inverse is the reciprocal of the Schur
// complement of the other models onto model 0 -- Gauss elimination from the last model down, rows
// in the warp's shared-memory tile.  Pivots that are exactly zero (models no candidate group
// touches: whole zero rows/columns) or below rcond x the largest diagonal entry are skipped, which
// reproduces pinv on structurally singular matrices; for numerically near-singular ones the result
// can differ from the eigenvalue cut-off of numpy (documented in DESIGN.md).
#pragma once
#include "blu_common.cuh"

#define BLU_IP_WARPS 4
#define BLU_IP_MAXLL 24

// psis: (LL, N*N) dense, built by blu_ip_expand_kernel.  ms: (LL, ncand) int64 row-major.
__global__ void __launch_bounds__(BLU_IP_WARPS * 32)
blu_ip_candidates_kernel(int N, int LL, const double *__restrict__ basephi, const double *__restrict__ psis,
                         const long long *__restrict__ ms, long long ncand, double rcond, double *__restrict__ Vs)
{
    extern __shared__ double ism[];
    const int NN = N * N;
    double *sPsi = ism;                          // LL * NN
    double *sBase = sPsi + (size_t)LL * NN;      // NN
    double *sA = sBase + NN;                     // BLU_IP_WARPS * N * (N+1)
    for (int t = threadIdx.x; t < LL * NN; t += blockDim.x) sPsi[t] = psis[t];
    for (int t = threadIdx.x; t < NN; t += blockDim.x) sBase[t] = basephi[t];
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ld = N + 1;
    double *A = sA + (size_t)w * N * ld;
    const long long gw = (long long)blockIdx.x * BLU_IP_WARPS + w;
    const long long nw = (long long)gridDim.x * BLU_IP_WARPS;
    for (long long c0 = gw * 32; c0 < ncand; c0 += nw * 32) {
        // lane = candidate: its LL coefficients
        double coef[BLU_IP_MAXLL];
#pragma unroll
        for (int t = 0; t < BLU_IP_MAXLL; ++t)
            coef[t] = (t < LL && c0 + lane < ncand) ? (double)ms[(long long)t * ncand + c0 + lane] : 0.0;
        double myV = 0.0;
        const int nc = (int)((ncand - c0) < 32 ? (ncand - c0) : 32);
        for (int j = 0; j < nc; ++j) {
            // row `lane` of Phi_c
            if (lane < N)
                for (int col = 0; col < N; ++col) A[lane * ld + col] = sBase[lane * N + col];
#pragma unroll
            for (int t = 0; t < BLU_IP_MAXLL; ++t) {
                if (t < LL) {
                    const double ct = blu_shfl(coef[t], j);
                    if (ct != 0.0 && lane < N) {
                        const double *P = sPsi + (size_t)t * NN + lane * N;
                        for (int col = 0; col < N; ++col) A[lane * ld + col] = fma(ct, P[col], A[lane * ld + col]);
                    }
                }
            }
            __syncwarp();
            double dmax = 0.0;
            for (int a = 0; a < N; ++a) dmax = fmax(dmax, fabs(A[a * ld + a]));
            const double thr = rcond * dmax;
            // eliminate models N-1 .. 1; what is left at (0,0) is the Schur complement
            for (int p = N - 1; p >= 1; --p) {
                const double piv = A[p * ld + p];
                if (fabs(piv) > thr) {
                    if (lane < p) {
                        const double f = A[lane * ld + p] / piv;
                        for (int col = 0; col < p; ++col) A[lane * ld + col] = fma(-f, A[p * ld + col], A[lane * ld + col]);
                    }
                }
                __syncwarp();
            }
            const double s00 = A[0];
            const double v = (fabs(s00) > thr && s00 != 0.0) ? 1.0 / s00 : 0.0;
            if (lane == j) myV = v;
            __syncwarp();
        }
        if (c0 + lane < ncand) Vs[c0 + lane] = myV;
    }
}

// Dense Psi_t (N x N) for the listed flat group ids, from the packed inverses.
__global__ void blu_ip_expand_kernel(const BluClass *__restrict__ cls, int ncls, int N, int LL, const long long *__restrict__ idx,
                                     const uint8_t *__restrict__ gidx, const double *__restrict__ cinv, double *__restrict__ psis)
{
    const int t = blockIdx.x;
    if (t >= LL) return;
    const long long gi = idx[t];
    int c = 0;
    while (c + 1 < ncls && cls[c + 1].goff <= gi) ++c;
    const BluClass ci = cls[c];
    const long long i = gi - ci.goff;
    const int k = ci.k;
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) psis[(size_t)t * N * N + e] = 0.0;
    __syncthreads();
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
        const int j = e / k, l = e - j * k;
        const int a = j < l ? j : l, b = j < l ? l : j;
        const uint8_t *g = gidx + ci.ioff + i * k;
        psis[(size_t)t * N * N + (int)g[j] * N + (int)g[l]] = cinv[ci.coff + i * ci.T + blu_pk(k, a, b)];
    }
}
