// blu_jacobi.cuh -- block-parallel cyclic Jacobi eigen-solver and symmetric pseudo-inverse for the
// small dense matrices of the path: Phi (N x N, N <= 32) once per evaluation, and the rare
// ill-conditioned group covariance at setup.
//
// Replaces numpy.linalg.pinv at misc.py:487,490 / sap.py:74 with the same semantics:
// singular values below 1e-15 * sigma_max are dropped (numpy's default rcond).  For a symmetric
// matrix the singular values are |lambda_i|, so pinv = sum_{|lambda_i| > cutoff} v_i v_i^T / lambda_i.
//
// Ordering: round-robin ("chess tournament") -- in each of the n-1 rounds of a sweep the n/2 index
// pairs are disjoint, so all rotations of a round are applied together: column phase
// (A <- A J, V <- V J), then row phase (A <- J^T A).  n must be even; odd sizes are padded with
// a zero row/column, which never rotates.
#pragma once
#include "blu_common.cuh"

#define BLU_JLD 33          // shared-memory leading dimension (odd: no bank aliasing between rows)
#define BLU_JMAX 32
#define BLU_JACOBI_MAX_SWEEPS 40

struct BluJacobiScratch {
    double c[BLU_JMAX / 2], s[BLU_JMAX / 2];
    int p[BLU_JMAX / 2], q[BLU_JMAX / 2];
    double lam[BLU_JMAX], inv[BLU_JMAX];
    int rotated;
    int sweeps;
    double lmax;
};

// All threads of the block call this.  A (n x n, ld BLU_JLD, symmetric) is destroyed (its diagonal
// ends up holding the eigenvalues), V receives the eigenvectors as columns.
__device__ __forceinline__ void blu_jacobi_eig(double *A, double *V, int n, BluJacobiScratch *js,
                                               int tid, int nthr)
{
    const int h = n >> 1;
    for (int t = tid; t < n * n; t += nthr) {
        int r = t / n, c = t - r * n;
        V[r * BLU_JLD + c] = (r == c) ? 1.0 : 0.0;
    }
    if (tid == 0) js->sweeps = 0;
    __syncthreads();
    for (int sweep = 0; sweep < BLU_JACOBI_MAX_SWEEPS; ++sweep) {
        if (tid == 0) js->rotated = 0;
        __syncthreads();
        for (int rnd = 0; rnd < n - 1; ++rnd) {
            if (tid < h) {
                int p, q;
                if (tid == 0) { p = n - 1; q = rnd; }
                else { p = (rnd + tid) % (n - 1); q = (rnd - tid + (n - 1)) % (n - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                const double app = A[p * BLU_JLD + p], aqq = A[q * BLU_JLD + q], apq = A[p * BLU_JLD + q];
                double c = 1.0, s = 0.0;
                const double mag = fabs(apq);
                if (mag > 1e-16 * sqrt(fabs(app * aqq)) && mag > 1e-300) {
                    const double tau = (aqq - app) / (2.0 * apq);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    s = t * c;
                    js->rotated = 1;
                }
                js->p[tid] = p; js->q[tid] = q; js->c[tid] = c; js->s[tid] = s;
            }
            __syncthreads();
            // column phase: M[:,p] <- c M[:,p] - s M[:,q] ; M[:,q] <- s M[:,p] + c M[:,q]  for M in {A, V}
            for (int t = tid; t < 2 * n * h; t += nthr) {
                const int which = t / (n * h);
                const int u = t - which * (n * h);
                const int r = u / h, i = u - r * h;
                double *M = which ? V : A;
                const int p = js->p[i], q = js->q[i];
                const double c = js->c[i], s = js->s[i];
                const double mp = M[r * BLU_JLD + p], mq = M[r * BLU_JLD + q];
                M[r * BLU_JLD + p] = c * mp - s * mq;
                M[r * BLU_JLD + q] = s * mp + c * mq;
            }
            __syncthreads();
            // row phase: A[p,:] <- c A[p,:] - s A[q,:] ; A[q,:] <- s A[p,:] + c A[q,:]
            for (int t = tid; t < n * h; t += nthr) {
                const int cc = t / h, i = t - cc * h;
                const int p = js->p[i], q = js->q[i];
                const double c = js->c[i], s = js->s[i];
                const double ap = A[p * BLU_JLD + cc], aq = A[q * BLU_JLD + cc];
                A[p * BLU_JLD + cc] = c * ap - s * aq;
                A[q * BLU_JLD + cc] = s * ap + c * aq;
            }
            __syncthreads();
        }
        const int rotated = js->rotated;
        __syncthreads();
        if (tid == 0) js->sweeps = sweep + 1;
        if (!rotated) break;
    }
    __syncthreads();
}

// pinv of the symmetric n x n matrix held in A (destroyed); result written to P (ld ldp, may be
// global or shared), exactly symmetric.  rcond follows numpy (1e-15).  n even (padded).
__device__ __forceinline__ void blu_sym_pinv(double *A, double *V, int n, BluJacobiScratch *js,
                                             double *P, int ldp, int nout, double rcond, int tid, int nthr)
{
    blu_jacobi_eig(A, V, n, js, tid, nthr);
    if (tid < n) js->lam[tid] = A[tid * BLU_JLD + tid];
    __syncthreads();
    if (tid == 0) {
        double lmax = 0.0;
        for (int i = 0; i < n; ++i) lmax = fmax(lmax, fabs(js->lam[i]));
        js->lmax = lmax;
        for (int i = 0; i < n; ++i)
            js->inv[i] = (fabs(js->lam[i]) > rcond * lmax) ? 1.0 / js->lam[i] : 0.0;
    }
    __syncthreads();
    for (int t = tid; t < nout * nout; t += nthr) {
        const int r = t / nout, c = t - r * nout;
        const int lo = r < c ? r : c, hi = r < c ? c : r;     // same operand order for (r,c) and (c,r)
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += V[lo * BLU_JLD + i] * js->inv[i] * V[hi * BLU_JLD + i];
        P[r * ldp + c] = acc;
    }
    __syncthreads();
}
