// blu_invert.cuh -- kernel (1): one-time batched inversion of the per-group covariance
// sub-matrices C[g_i, g_i]  (replaces the L Python-level np.linalg.pinv calls of sap.py:69-75).
//
// Fast path  blu_invert_groups_kernel<K>: a slice of SUB = pow2ceil(K) lanes owns one group
// (32/SUB groups per warp), lane r keeps row r of the k x k block in K registers, and an in-place
// Gauss-Jordan sweep runs entirely on warp shuffles: pivot row p is broadcast from lane p, one
// element at a time.  C is SPD on every admissible group, so no pivoting is needed; each pivot is
// compared with its original diagonal entry and the group is flagged when it falls below
// `pivtol` x diagonal (rank-deficient / badly conditioned block).
//
// Slow path  blu_pinv_groups_kernel: flagged groups are redone by one CTA each with the Jacobi
// eigen-solver and numpy's pinv cutoff, so the semantics of np.linalg.pinv are kept where they
// matter (hard part 1).
//
// Output: packed upper triangle, see blu_common.cuh.
#pragma once
#include "blu_common.cuh"
#include "blu_jacobi.cuh"

template <int K> struct BluSub { static constexpr int value = K <= 1 ? 1 : K <= 2 ? 2 : K <= 4 ? 4 : K <= 8 ? 8 : K <= 16 ? 16 : 32; };

template <int K>
__global__ void __launch_bounds__(128)
blu_invert_groups_kernel(const double *__restrict__ C, int N, const uint8_t *__restrict__ gidx,
                         long long Lk, double *__restrict__ cinv, unsigned char *__restrict__ flag,
                         double pivtol)
{
    constexpr int SUB = BluSub<K>::value;
    constexpr int G = 32 / SUB;
    constexpr int T = K * (K + 1) / 2;
    __shared__ double sC[BLU_MAX_MODELS_C * BLU_MAX_MODELS_C];
    for (int t = threadIdx.x; t < N * N; t += blockDim.x) sC[t] = C[t];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int sub = lane / SUB, r = lane % SUB;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const unsigned submask = (SUB == 32) ? 0xffffffffu : (((1u << SUB) - 1u) << (sub * SUB));

    for (long long base = warp * G; base < Lk; base += nwarps * G) {
        const long long grp = base + sub;
        const bool live = grp < Lk;
        const bool valid = live && r < K;
        const int gr = valid ? (int)gidx[grp * K + r] : 0;
        double a[K];
        double diag0 = 1.0;
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const int gc = __shfl_sync(BLU_FULL, gr, c, SUB);
            a[c] = valid ? sC[gr * N + gc] : ((c == r) ? 1.0 : 0.0);
            if (c == r) diag0 = a[c];
        }
        bool bad = false;
#pragma unroll
        for (int p = 0; p < K; ++p) {
            const double app = __shfl_sync(BLU_FULL, a[p], p, SUB);
            const double d0 = __shfl_sync(BLU_FULL, diag0, p, SUB);
            if (!(app > pivtol * d0)) bad = true;          // also catches NaN and non-positive pivots
            const double d = 1.0 / app;
            const double f = a[p] * d;
#pragma unroll
            for (int c = 0; c < K; ++c) {
                const double apc = __shfl_sync(BLU_FULL, a[c], p, SUB);
                if (c != p) a[c] = (r == p) ? apc * d : fma(-f, apc, a[c]);
            }
            a[p] = (r == p) ? d : -f;
        }
        const unsigned badmask = __ballot_sync(BLU_FULL, bad && valid);
        if (valid) {
            double *out = cinv + grp * T + (r * K - r * (r - 1) / 2);
#pragma unroll
            for (int c = 0; c < K; ++c)
                if (c >= r) out[c - r] = a[c];
            if (r == 0) flag[grp] = (badmask & submask) ? 1 : 0;
        }
    }
}

// One CTA per flagged group: gather the block, Jacobi pseudo-inverse, store packed upper triangle.
__global__ void __launch_bounds__(256)
blu_pinv_groups_kernel(const double *__restrict__ C, int N, int k, const uint8_t *__restrict__ gidx,
                       const long long *__restrict__ todo, double *__restrict__ cinv, double rcond)
{
    __shared__ double A[BLU_JMAX * BLU_JLD], V[BLU_JMAX * BLU_JLD], P[BLU_JMAX * BLU_JLD];
    __shared__ BluJacobiScratch js;
    __shared__ int g[BLU_JMAX];
    const long long grp = todo[blockIdx.x];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n = k + (k & 1);
    if (tid < k) g[tid] = gidx[grp * k + tid];
    __syncthreads();
    for (int t = tid; t < n * n; t += nthr) {
        const int r = t / n, c = t - r * n;
        A[r * BLU_JLD + c] = (r < k && c < k) ? C[g[r] * N + g[c]] : 0.0;
    }
    __syncthreads();
    blu_sym_pinv(A, V, n, &js, P, BLU_JLD, k, rcond, tid, nthr);
    const int T = blu_tri(k);
    for (int t = tid; t < k * k; t += nthr) {
        const int r = t / k, c = t - r * k;
        if (c >= r) cinv[grp * T + blu_pk(k, r, c)] = P[r * BLU_JLD + c];
    }
}

// Reference-layout ingest: (Lk,k,k) full inverses -> symmetrised packed upper triangle.
__global__ void blu_pack_invcovs_kernel(const double *__restrict__ full, int k, long long Lk,
                                        double *__restrict__ cinv)
{
    const int T = blu_tri(k);
    const long long total = Lk * T;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / T;
        int e = (int)(t - i * T);
        int j = 0;
        while (e >= k - j) { e -= k - j; ++j; }
        const int l = j + e;
        const double *blk = full + i * k * k;
        cinv[t] = 0.5 * (blk[j * k + l] + blk[l * k + j]);
    }
}

// Packed -> reference layout (both triangles).
__global__ void blu_unpack_invcovs_kernel(const double *__restrict__ cinv, int k, long long Lk,
                                          double *__restrict__ full)
{
    const long long total = Lk * k * k;
    const int T = blu_tri(k);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / (k * k);
        const int rem = (int)(t - i * k * k);
        const int j = rem / k, l = rem - j * k;
        const int lo = j < l ? j : l, hi = j < l ? l : j;
        full[t] = cinv[i * T + blu_pk(k, lo, hi)];
    }
}
