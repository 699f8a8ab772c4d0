// blu_grad.cuh -- kernel (3a): per-group quadratic forms with x = first row of pinv(Phi).
//
//   grad_i = - x[g_i]^T Cinv_i x[g_i]                         (gradK_c, cmisc.cpp:58-72; misc.py:493)
//   u_i    = R_i^T Cinv_i x[g_i]  (scattered to model slots)   (intended cleanup matrix, misc.py:507-516)
//   v_i    = S u_i,  S = pinv(Phi) + pinv(Phi)^T = 2 pinv(Phi)
// so that the Hessian of misc.py:497-503 is  H = U S U^T = U V^T  (SURVEY.md 0.3), which
// blu_hess.cuh forms with FP64 tensor-core MMAs.
//
// blu_grad_kernel   (no Hessian wanted): lane-per-packed-entry streaming, one shuffle-tree sum per
//                   group -- a pure HBM stream of the packed inverses.
// blu_gradu_kernel  (Hessian / U wanted): the group's packed block is staged in the warp's shared
//                   memory slice, lane j forms row j of Cinv_i x[g_i], rows are scattered to model
//                   positions via the membership mask and multiplied by S; U and V rows (NP doubles,
//                   one 128-byte line at N <= 16) are written coalesced.
// x and S are broadcast once per CTA into registers / shared memory.
#pragma once
#include "blu_common.cuh"

#define BLU_GRAD_WARPS 8

__global__ void __launch_bounds__(BLU_GRAD_WARPS * 32)
blu_grad_kernel(const BluClass *__restrict__ cls, int ncls, int N, const uint8_t *__restrict__ gidx,
                const double *__restrict__ cinv, const uint16_t *__restrict__ lut,
                const double *__restrict__ xrow, long long lo, long long hi, double *__restrict__ grad)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double xv = lane < N ? xrow[lane] : 0.0;
    const long long gw = (long long)blockIdx.x * BLU_GRAD_WARPS + w;
    const long long nw = (long long)gridDim.x * BLU_GRAD_WARPS;
    for (int c = 0; c < ncls; ++c) {
        const BluClass ci = cls[c];
        const int k = ci.k, T = ci.T;
        const uint16_t *lt = lut + ci.lutoff;
        long long i0 = lo > ci.goff ? lo - ci.goff : 0;
        long long i1 = hi < ci.goff + ci.Lk ? hi - ci.goff : ci.Lk;
        for (long long i = i0 + gw; i < i1; i += nw) {
            const int gv = lane < k ? (int)gidx[ci.ioff + i * k + lane] : 0;
            const double xg = blu_shfl(xv, gv);
            const double *cp = cinv + ci.coff + i * T;
            double sum = 0.0;
            for (int e0 = 0; e0 < T; e0 += 32) {
                const int e = e0 + lane;
                const bool ok = e < T;
                const unsigned jl = ok ? lt[e] : 0u;
                const double v = ok ? cp[e] : 0.0;
                const int j = jl >> 8, l = jl & 255u;
                const double xa = blu_shfl(xg, j), xb = blu_shfl(xg, l);
                const double wgt = (j == l) ? 1.0 : 2.0;
                sum += wgt * (xa * v * xb);
            }
            sum = blu_warp_sum(sum);
            if (lane == 0) grad[ci.goff + i] = -sum;
        }
    }
}

// dynamic shared memory: N*N doubles (S) + BLU_GRAD_WARPS * Tmax doubles (staging)
__global__ void __launch_bounds__(BLU_GRAD_WARPS * 32)
blu_gradu_kernel(const BluClass *__restrict__ cls, int ncls, int N, int NP, int Tmax,
                 const uint8_t *__restrict__ gidx, const unsigned *__restrict__ gmask,
                 const double *__restrict__ cinv, const double *__restrict__ xrow,
                 const double *__restrict__ S, long long lo, long long hi,
                 double *__restrict__ grad, double *__restrict__ U, double *__restrict__ V)
{
    extern __shared__ double sm[];
    double *sS = sm;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *st = sm + N * N + w * Tmax;
    for (int t = threadIdx.x; t < N * N; t += blockDim.x) sS[t] = S[t];
    __syncthreads();
    const double xv = lane < N ? xrow[lane] : 0.0;
    const long long gw = (long long)blockIdx.x * BLU_GRAD_WARPS + w;
    const long long nw = (long long)gridDim.x * BLU_GRAD_WARPS;
    for (int c = 0; c < ncls; ++c) {
        const BluClass ci = cls[c];
        const int k = ci.k, T = ci.T;
        long long i0 = lo > ci.goff ? lo - ci.goff : 0;
        long long i1 = hi < ci.goff + ci.Lk ? hi - ci.goff : ci.Lk;
        for (long long i = i0 + gw; i < i1; i += nw) {
            const double *cp = cinv + ci.coff + i * T;
            for (int e = lane; e < T; e += 32) st[e] = cp[e];
            const int gv = lane < k ? (int)gidx[ci.ioff + i * k + lane] : 0;
            const unsigned mask = gmask[ci.goff + i];
            const double xg = blu_shfl(xv, gv);
            __syncwarp();
            const int j = lane < k ? lane : k - 1;
            double y = 0.0;
            for (int l = 0; l < k; ++l) {
                const double xl = blu_shfl(xg, l);
                const int a = j < l ? j : l, b = j < l ? l : j;
                y += st[blu_pk(k, a, b)] * xl;
            }
            const double gs = blu_warp_sum(lane < k ? xg * y : 0.0);
            // scatter rows of Cinv_i x[g_i] to model slots
            const bool in = (mask >> lane) & 1u;
            const int pos = __popc(mask & ((1u << lane) - 1u));
            const double ysrc = blu_shfl(y, pos);
            const double ua = in ? ysrc : 0.0;
            double va = 0.0;
            for (int b = 0; b < N; ++b) {
                const double ub = blu_shfl(ua, b);
                if (lane < N) va += sS[lane * N + b] * ub;
            }
            const long long row = ci.goff + i;
            if (lane == 0) grad[row] = -gs;
            if (lane < NP) {
                U[row * NP + lane] = ua;
                V[row * NP + lane] = lane < N ? va : 0.0;
            }
            __syncwarp();
        }
    }
}

// get_cleanup_matrix (misc.py:507-516).  X is (N, L) row-major on the device.
// mode 0: the reference's assignment semantics, X[g[j], i] = Cinv_i[j,k-1] * x[g[k-1]]
//         (cmisc.cpp:51: "=" inside the l loop, the last l wins).
// mode 1: the intended matrix, X[:, i] = u_i.
__global__ void blu_cleanup_kernel(const BluClass *__restrict__ cls, int ncls, int N, long long L,
                                   const uint8_t *__restrict__ gidx, const double *__restrict__ cinv,
                                   const double *__restrict__ xrow, int mode, double *__restrict__ X)
{
    for (int c = 0; c < ncls; ++c) {
        const BluClass ci = cls[c];
        const int k = ci.k, T = ci.T;
        const long long total = ci.Lk * k;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
             t += (long long)gridDim.x * blockDim.x) {
            const long long i = t / k;
            const int j = (int)(t - i * k);
            const uint8_t *g = gidx + ci.ioff + i * k;
            const double *cp = cinv + ci.coff + i * T;
            double val;
            if (mode == 0) {
                val = cp[blu_pk(k, j, k - 1)] * xrow[g[k - 1]];
            } else {
                val = 0.0;
                for (int l = 0; l < k; ++l) {
                    const int a = j < l ? j : l, b = j < l ? l : j;
                    val += cp[blu_pk(k, a, b)] * xrow[g[l]];
                }
            }
            X[(long long)g[j] * L + ci.goff + i] = val;
        }
    }
}

// Dense psi (N*N, L): psi[N*g[j]+g[l], i] = Cinv_i[j,l]   (assemble_psi_c, cmisc.cpp:10-23; sap.py:129).
// psi must be zero-filled; columns [col0, col0+ncols) of the flat enumeration are produced into a
// (N*N, ncols) row-major panel.
__global__ void blu_psi_kernel(const BluClass *__restrict__ cls, int ncls, int N, long long col0,
                               long long ncols, const uint8_t *__restrict__ gidx,
                               const double *__restrict__ cinv, double *__restrict__ psi)
{
    for (int c = 0; c < ncls; ++c) {
        const BluClass ci = cls[c];
        const int k = ci.k, T = ci.T;
        long long i0 = col0 > ci.goff ? col0 - ci.goff : 0;
        long long i1 = col0 + ncols < ci.goff + ci.Lk ? col0 + ncols - ci.goff : ci.Lk;
        if (i1 <= i0) continue;
        const long long total = (i1 - i0) * k * k;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
             t += (long long)gridDim.x * blockDim.x) {
            const long long i = i0 + t / (k * k);
            const int rem = (int)(t % (k * k));
            const int j = rem / k, l = rem - j * k;
            const uint8_t *g = gidx + ci.ioff + i * k;
            const int a = j < l ? j : l, b = j < l ? l : j;
            psi[((long long)N * g[j] + g[l]) * ncols + (ci.goff + i - col0)] = cinv[ci.coff + i * T + blu_pk(k, a, b)];
        }
    }
}
