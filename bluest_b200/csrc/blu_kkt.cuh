// blu_kkt.cuh -- scope-table row f1: a structure-exploiting KKT solve for the semidefinite programme that
// SAP.cvxopt_solve hands to cvxopt.solvers.sdp (sap.py:242-307; cvxopt calls its built-in dense solver there).
//
// Variables x in R^n (n = L + 1 with the epigraph variable t in front for a budget, sap.py:259-275; n = L for a
// tolerance, :276-287).  Inequalities: G0 x <=_l h0 with G0 = [-I_n ; Gx] (Gx: the nlin dense rows -- cost, coverage,
// sample caps) and one (N+1) x (N+1) semidefinite block G1 x <=_s h1 whose column of group i is
// -scales * vec(pad(Psi_i)) and whose t column (budget) is -E_NN.  With the Nesterov-Todd scaling
// W = blkdiag(diag(d), W_s), W_s vec(U) = vec(r^T U r), every interior-point iteration solves
//     [ 0   G^T   ] [ux]   [bx]
//     [ G  -W^T W ] [uz] = [bz].
// Eliminating uz:  M ux = bx + G^T (W^T W)^-1 bz,  M = diag(d_I^-2) + Gx^T diag(d_lin^-2) Gx + A^T A  with
// A = (r^-1 (x) r^-1) G1: a DIAGONAL plus a matrix of rank Q = (N+1)(N+2)/2 + nlin.  Woodbury:
//     ux = D^-1 rhs - D^-1 B (I + B^T D^-1 B)^-1 B^T D^-1 rhs,     B = [A^T | Gx^T diag(1/d_lin)],  D = diag(d_I^-2)
// so an iteration costs one weighted Gram contraction over the groups (Q^2 L multiply-adds on the FP64 tensor cores)
// and a Q x Q Cholesky instead of the (L+1)^3/3 of the dense factorisation.
//
// Device pipeline (all FP64):
//   blu_kkt_rows_kernel   one warp per column: a_col = svec(-scales r^-1 pad(Psi_i) r^-T) from the packed inverse and
//                         the member columns of r^-1, the nlin constraint entries, everything times d_col (= D^-1/2),
//                         plus the column's entry of G1^T vec(Lam Z Lam) (right-hand side) -> row of Bs (n x QP)
//   blu_kkt_syrk_kernel   cap = I + Bs^T Bs and v = Bs^T (d * rhs) in one pass (the rhs rides along as column Q):
//                         mma.m8n8k4.f64, one CTA per 8 x 8 tile pair, warps combined in a fixed order
//   blu_kkt_chol_kernel   Cholesky of the Q x Q capacitance matrix + two triangular solves, one CTA
//   blu_kkt_apply_kernel  ux_col = d_col^2 rhs_col - d_col (Bs_col . y)
// then Phi(ux) through the ordinary Phi kernel gives G1 ux, and uz follows on the host (O(n + N^2) work).
#pragma once
#include "blu_common.cuh"
#include "blu_hess.cuh"

#define BLU_KKT_WARPS 8

// svec index of (a,b), a <= b, of an M x M symmetric matrix stored by rows of the upper triangle
__host__ __device__ __forceinline__ int blu_svec(int M, int a, int b) { return a * M - a * (a - 1) / 2 + (b - a); }

// One warp per column of the reduced system.
//   col < has_t            : the t column, X = -E_NN
//   otherwise group i      : X = -scales * pad(Psi_i)
// Bs[col][q] = d[col] * b_col[q],  q < Qs: svec(r^-1 X r^-T) (off-diagonal entries times sqrt 2 so that the
// Euclidean inner product of svecs equals the trace inner product), Qs <= q < Q: Gx[q-Qs][col] / dlin[q-Qs],
// Bs[col][Q] = d[col] * rhs_col (filled by blu_kkt_rhs_kernel), zero padding up to QP.
// g1tw[col] = <X_col, Wm>  (Wm = Lam Z Lam, (N+1) x (N+1)): the column's entry of G1^T vec(Wm).
__global__ void __launch_bounds__(BLU_KKT_WARPS * 32)
blu_kkt_rows_kernel(const BluClass *__restrict__ cls, int ncls, int N, long long L, int has_t, double scales,
                    const uint8_t *__restrict__ gidx, const double *__restrict__ cinv, const double *__restrict__ rinv,
                    const double *__restrict__ Wm, const double *__restrict__ d, int nlin, const double *__restrict__ Gx,
                    const double *__restrict__ dlin, int Q, int QP, double *__restrict__ Bs, double *__restrict__ g1tw)
{
    const int M = N + 1, Qs = M * (M + 1) / 2;
    const int MP = M | 1;                                 // odd leading dimensions: rows of r^-1 / Tt land in different banks
    extern __shared__ double kraw[];                      // [M*MP r^-1][M*MP Wm][WARPS x M*(N|1) Tt]
    double *sR = kraw, *sW = kraw + M * MP;
    __shared__ BluClass scls[BLU_MAX_MODELS_C];
    for (int t = threadIdx.x; t < M * M; t += blockDim.x) { sR[(t / M) * MP + (t % M)] = rinv[t]; sW[(t / M) * MP + (t % M)] = Wm[t]; }
    for (int t = threadIdx.x; t < ncls; t += blockDim.x) scls[t] = cls[t];
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n = L + has_t;
    const double rt2 = 1.4142135623730951;
    for (long long col = (long long)blockIdx.x * BLU_KKT_WARPS + w; col < n; col += (long long)gridDim.x * BLU_KKT_WARPS) {
        double *row = Bs + col * QP;
        const double dc = d[col];
        double dotw = 0.0;
        if (col < has_t) {
            // X = -E_NN: r^-1 X r^-T = -(column N of r^-1)(column N of r^-1)^T
            for (int q = lane; q < Qs; q += 32) {
                int a = 0, rem = q;
                while (rem >= M - a) { rem -= M - a; ++a; }
                const int b = a + rem;
                const double v = -(sR[a * MP + N] * sR[b * MP + N]);
                row[q] = dc * (a == b ? v : rt2 * v);
            }
            if (lane == 0) dotw = -sW[N * MP + N];
        } else {
            const long long i = col - has_t;
            int ic = 0;
            while (ic + 1 < ncls && i >= scls[ic + 1].goff) ++ic;
            const BluClass ci = scls[ic];
            const long long il = i - ci.goff;
            const int k = ci.k;
            const uint8_t *g = gidx + ci.ioff + il * k;
            const double *C = cinv + ci.coff + il * ci.T;
            const int KP = k | 1;
            double *Tt = kraw + 2 * M * MP + (size_t)w * M * (N | 1);    // Tt[a][l] = sum_j r^-1[a][g_j] C[j][l], leading dimension KP
            // Tt[a][l] = sum_j r^-1[a][g_j] C[j][l]
            for (int t = lane; t < M * k; t += 32) {
                const int a = t / k, l = t - a * k;
                double s = 0.0;
                for (int j = 0; j < k; ++j) {
                    const int lo = j < l ? j : l, hi = j < l ? l : j;
                    s = fma(sR[a * MP + g[j]], C[blu_pk(k, lo, hi)], s);
                }
                Tt[a * KP + l] = s;
            }
            __syncwarp();
            for (int q = lane; q < Qs; q += 32) {
                int a = 0, rem = q;
                while (rem >= M - a) { rem -= M - a; ++a; }
                const int b = a + rem;
                double s = 0.0;
                for (int l = 0; l < k; ++l) s = fma(Tt[a * KP + l], sR[b * MP + g[l]], s);
                s *= -scales;
                row[q] = dc * (a == b ? s : rt2 * s);
            }
            // <X_i, Wm> = -scales sum_{j,l} C[j][l] Wm[g_j][g_l]
            for (int e = lane; e < ci.T; e += 32) {
                int j = 0, rem = e;
                while (rem >= k - j) { rem -= k - j; ++j; }
                const int l = j + rem;
                const double wv = (j == l) ? sW[g[j] * MP + g[j]] : (sW[g[j] * MP + g[l]] + sW[g[l] * MP + g[j]]);
                dotw = fma(C[e], wv, dotw);
            }
            dotw *= -scales;
            __syncwarp();
        }
        dotw = blu_warp_sum(dotw);
        for (int q = Qs + lane; q < QP; q += 32) {
            double v = 0.0;
            if (q < Q) v = dc * Gx[(long long)(q - Qs) * n + col] / dlin[q - Qs];
            row[q] = v;                                   // column Q (the right-hand side) is written by blu_kkt_rhs_kernel
        }
        if (lane == 0) g1tw[col] = dotw;
    }
}

// rhs = bx + G0^T (d^-2 bz0) + G1^T vec(Lam Z Lam);  Bs[col][Q] = d[col] * rhs[col].
__global__ void blu_kkt_rhs_kernel(long long n, int nlin, const double *__restrict__ bx, const double *__restrict__ bz0,
                                   const double *__restrict__ d, const double *__restrict__ Gx, const double *__restrict__ dlin,
                                   const double *__restrict__ g1tw, int Q, int QP, double *__restrict__ rhs, double *__restrict__ Bs)
{
    for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < n; col += (long long)gridDim.x * blockDim.x) {
        double v = bx[col] - bz0[col] / (d[col] * d[col]);          // -I rows of G0
        for (int q = 0; q < nlin; ++q) v += Gx[(long long)q * n + col] * (bz0[n + q] / (dlin[q] * dlin[q]));
        v += g1tw[col];
        rhs[col] = v;
        Bs[col * QP + Q] = d[col] * v;
    }
}

// cap = I + Bs^T Bs and v = Bs^T (d * rhs) (column Q of Bs): ONE CTA PER TILE PAIR (ti <= tj) of the QP = 8 NTQ
// columns.  The 8 warps split the rows; a lane's fragment of tile t (row 4 s + (lane & 3), column 8 t + (lane >> 2)) is
// the A operand for t = ti and the B operand for t = tj of mma.m8n8k4.f64; four independent accumulator pairs per
// warp hide the DMMA latency.  Bs (n x QP doubles, tens of MB) is L2 resident, so re-reading two tile columns per
// pair costs L2 bandwidth only.  Warps are combined in warp order: no partial tiles in HBM, no atomics, bit-reproducible.
__global__ void __launch_bounds__(BLU_KKT_WARPS * 32)
blu_kkt_syrk_kernel(const double *__restrict__ Bs, long long n, int QP, double *__restrict__ part)
{
    // grid (tile pairs, row splits): more CTAs per SM = more strided L2 reads in flight (one split was latency bound:
    // long_scoreboard 59 warps per issue, tensor pipe 4 %); partial 8 x 8 tiles, folded in split order by the next kernel
    __shared__ double red[BLU_KKT_WARPS][64];
    const int NTQ = QP >> 3;
    int ti = 0, rem = blockIdx.x;
    while (rem >= NTQ - ti) { rem -= NTQ - ti; ++ti; }
    const int tj = ti + rem;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ks = lane & 3, cq = lane >> 2;
    const long long rows_per = (((n + gridDim.y - 1) / gridDim.y) + 15) / 16 * 16;
    const long long r0 = (long long)blockIdx.y * rows_per;
    long long r1 = r0 + rows_per;
    if (r1 > n) r1 = n;
    double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    const double *pi = Bs + 8 * ti + cq, *pj = Bs + 8 * tj + cq;
    for (long long base = r0 + (long long)w * 32; base < r1; base += (long long)BLU_KKT_WARPS * 32) {
        double fi[8], fj[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long row = base + 4 * u + ks;
            const bool ok = row < r1;
            fi[u] = ok ? __ldg(pi + row * QP) : 0.0;
            fj[u] = ok ? __ldg(pj + row * QP) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) blu_dmma(acc[u & 3][0], acc[u & 3][1], fi[u], fj[u]);
    }
    const double c0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
    const double c1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
    red[w][cq * 8 + 2 * ks] = c0;                        // C fragment: row cq, cols 2 ks + {0,1}
    red[w][cq * 8 + 2 * ks + 1] = c1;
    __syncthreads();
    if (threadIdx.x < 64) {
        double s = 0.0;
#pragma unroll
        for (int ww = 0; ww < BLU_KKT_WARPS; ++ww) s += red[ww][threadIdx.x];
        part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64 + threadIdx.x] = s;
    }
}

// cap = I + sum over the row splits (split order) of the partial tiles, mirrored; v = column Q.
__global__ void blu_kkt_capfold_kernel(const double *__restrict__ part, int npairs, int nsplit, int Q, int QP,
                                       double *__restrict__ cap, double *__restrict__ v)
{
    const int NTQ = QP >> 3;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < npairs * 64; t += gridDim.x * blockDim.x) {
        const int p = t >> 6, e = t & 63;
        int ti = 0, rem = p;
        while (rem >= NTQ - ti) { rem -= NTQ - ti; ++ti; }
        const int tj = ti + rem;
        double s = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) s += part[((size_t)sp * npairs + p) * 64 + e];
        const int r = 8 * ti + (e >> 3), c = 8 * tj + (e & 7);
        if (r < Q && c < Q) {
            if (ti != tj || r <= c) {
                const double val = s + (r == c ? 1.0 : 0.0);
                cap[r * Q + c] = val;
                cap[c * Q + r] = val;
            }
        } else if (c == Q && r < Q) v[r] = s;
    }
}

// Cholesky (lower, in place) of the Q x Q SPD capacitance matrix and the solve cap y = v, one CTA.
__global__ void __launch_bounds__(512)
blu_kkt_chol_kernel(double *__restrict__ cap, int Q, const double *__restrict__ v, double *__restrict__ y, int *__restrict__ info)
{
    extern __shared__ double cs[];                        // Q x (Q + 1)
    const int ld = Q + 1, tid = threadIdx.x, nthr = blockDim.x;
    for (int t = tid; t < Q * Q; t += nthr) cs[(t / Q) * ld + (t % Q)] = cap[t];
    __syncthreads();
    // left-looking: column j = (A[j:, j] - L[j:, :j] L[j, :j]^T) / l_jj, thread i owns row i -- two barriers per column,
    // row reads are conflict free (odd leading dimension), the pivot row is a broadcast
    __shared__ double s_piv;
    for (int j = 0; j < Q; ++j) {
        double sv[1];
        const int i = j + tid;
        sv[0] = 0.0;
        if (i < Q) {
            double acc0 = cs[i * ld + j], acc1 = 0.0;
            const double *ri = cs + i * ld, *rj = cs + j * ld;
            int c = 0;
#pragma unroll 4
            for (; c + 1 < j; c += 2) {                       // independent loads batched: the plain loop waited ~35 cycles per term
                acc0 = fma(-ri[c], rj[c], acc0);
                acc1 = fma(-ri[c + 1], rj[c + 1], acc1);
            }
            if (c < j) acc0 = fma(-ri[c], rj[c], acc0);
            acc0 += acc1;
            sv[0] = acc0;
            if (i == j) s_piv = acc0;
        }
        __syncthreads();
        const double djj = s_piv;
        if (!(djj > 0.0)) { if (tid == 0) *info = j + 1; return; }          // same value in every thread
        const double sj = sqrt(djj);
        if (i < Q) cs[i * ld + j] = (i == j) ? sj : sv[0] / sj;
        __syncthreads();
    }
    // forward and backward substitution by one warp (Q <= 240): lane-strided dot products, shuffle reduction
    __shared__ double ys[256];
    if (tid < 32) {
        for (int i = 0; i < Q; ++i) {
            double s = 0.0;
            for (int c = tid; c < i; c += 32) s = fma(cs[i * ld + c], ys[c], s);
            s = blu_warp_sum(s);
            if (tid == 0) ys[i] = (v[i] - s) / cs[i * ld + i];
            __syncwarp();
        }
        for (int i = Q - 1; i >= 0; --i) {
            double s = 0.0;
            for (int c = i + 1 + tid; c < Q; c += 32) s = fma(cs[c * ld + i], ys[c], s);
            s = blu_warp_sum(s);
            if (tid == 0) ys[i] = (ys[i] - s) / cs[i * ld + i];
            __syncwarp();
        }
        if (tid == 0) *info = 0;
    }
    __syncthreads();
    for (int t = tid; t < Q; t += nthr) y[t] = ys[t];
}

// ux[col] = d[col]^2 rhs[col] - d[col] (Bs[col][0..Q) . y).  One warp per column.
__global__ void __launch_bounds__(BLU_KKT_WARPS * 32)
blu_kkt_apply_kernel(const double *__restrict__ Bs, long long n, int Q, int QP, const double *__restrict__ y,
                     const double *__restrict__ d, const double *__restrict__ rhs, double *__restrict__ ux)
{
    __shared__ double sy[256];
    for (int t = threadIdx.x; t < Q; t += blockDim.x) sy[t] = y[t];
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long col = (long long)blockIdx.x * BLU_KKT_WARPS + w; col < n; col += (long long)gridDim.x * BLU_KKT_WARPS) {
        const double *row = Bs + col * QP;
        double s = 0.0;
        for (int q = lane; q < Q; q += 32) s = fma(row[q], sy[q], s);
        s = blu_warp_sum(s);
        if (lane == 0) ux[col] = d[col] * (d[col] * rhs[col] - s);
    }
}
