// blu_level1.cuh -- Level-1 drop-ins for the five routines of `_cmisc_bluest`
// (bluest/cmisc.cpp:10-97).  Arguments are the reference's: HOST pointers, the reference's
// (Lk,k) int64 group tables and (Lk,k,k) full inverses, caller-allocated outputs that are
// accumulated in place.  Each call uploads its operands (and the current contents of the output,
// to honour "+="), runs a CUDA kernel and downloads the output -- a PCIe-bound convenience for
// code that still calls the per-class routines; the fast path is the Level-2 context.
#pragma once
#include <string>
#include <vector>
#include "blu_common.cuh"

#define BLU_L1_WARPS 8

// psi[Lk*(N*g[j]+g[l]) + i] += Cinv_i[j,l]           (assemble_psi_c, cmisc.cpp:10-23)
__global__ void blu_l1_psi_kernel(double *psi, int N, int k, long long Lk, const long long *g, const double *inv)
{
    const long long total = Lk * k * k;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / (k * k);
        const int rem = (int)(t - i * k * k);
        const int j = rem / k, l = rem - j * k;
        psi[Lk * (N * g[i * k + j] + g[i * k + l]) + i] += inv[t];
    }
}

// per-CTA partial of PHI[N*g[j]+g[l]] += m_i*Cinv_i[j,l]     (objectiveK_c, cmisc.cpp:25-40)
__global__ void __launch_bounds__(BLU_L1_WARPS * 32)
blu_l1_phi_kernel(double *part, int N, int k, long long Lk, const double *md, const long long *mi,
                  const long long *g, const double *inv)
{
    extern __shared__ double sacc[];
    const int NN = N * N, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *acc = sacc + w * NN;
    for (int t = lane; t < NN; t += 32) acc[t] = 0.0;
    __syncwarp();
    const int kk = k * k;
    for (long long i = (long long)blockIdx.x * BLU_L1_WARPS + w; i < Lk; i += (long long)gridDim.x * BLU_L1_WARPS) {
        const double wgt = md ? md[i] : (double)mi[i];
        for (int e = lane; e < kk; e += 32) {
            const int j = e / k, l = e - j * k;
            acc[N * (int)g[i * k + j] + (int)g[i * k + l]] += wgt * inv[i * kk + e];
        }
        __syncwarp();
    }
    __syncthreads();
    for (int t = threadIdx.x; t < NN; t += blockDim.x) {
        double s = 0.0;
        for (int ww = 0; ww < BLU_L1_WARPS; ++ww) s += sacc[ww * NN + t];
        part[(long long)blockIdx.x * NN + t] = s;
    }
}
__global__ void blu_l1_reduce_add_kernel(double *out, const double *part, int nparts, int n)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int p = 0; p < nparts; ++p) s += part[(long long)p * n + t];
        out[t] += s;
    }
}

// grad[i] += sum_{j,l} x[g[j]] Cinv_i[j,l] x[g[l]]            (gradK_c, cmisc.cpp:58-72)
__global__ void __launch_bounds__(BLU_L1_WARPS * 32)
blu_l1_grad_kernel(double *grad, int k, long long Lk, const long long *g, const double *inv, const double *x)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, kk = k * k;
    for (long long i = (long long)blockIdx.x * BLU_L1_WARPS + w; i < Lk; i += (long long)gridDim.x * BLU_L1_WARPS) {
        double s = 0.0;
        for (int e = lane; e < kk; e += 32) {
            const int j = e / k, l = e - j * k;
            s += x[g[i * k + j]] * inv[i * kk + e] * x[g[i * k + l]];
        }
        s = blu_warp_sum(s);
        if (lane == 0) grad[i] += s;
    }
}

// X[Lk*g[j] + i] = Cinv_i[j,k-1] * x[g[k-1]]    (cleanupK_c, cmisc.cpp:42-56: "=" in the l loop)
__global__ void blu_l1_cleanup_kernel(double *X, int k, long long Lk, const long long *g, const double *inv, const double *x)
{
    const long long total = Lk * k;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / k;
        const int j = (int)(t - i * k);
        X[Lk * g[i * k + j] + i] = inv[i * k * k + j * k + (k - 1)] * x[g[i * k + (k - 1)]];
    }
}

// rows (Lg, NP): mode 0  a_i = scatter(Cinv_i^T x[g_i]);  mode 1  w_i = P * scatter(Cinv_i x[g_i])
__global__ void blu_l1_rows_kernel(double *rows, int N, int NP, int k, long long Lg, const long long *g, const double *inv,
                                   const double *P, int mode)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (long long i = (long long)blockIdx.x * wpb + w; i < Lg; i += (long long)gridDim.x * wpb) {
        // lane j < k: y_j
        double y = 0.0;
        if (lane < k) {
            for (int l = 0; l < k; ++l) {
                const double cij = mode == 0 ? inv[i * k * k + l * k + lane] : inv[i * k * k + lane * k + l];
                y += cij * P[g[i * k + l]];                 // x = first row of P
            }
        }
        // scatter to model slots
        unsigned mask = 0;
        for (int j = 0; j < k; ++j) mask |= 1u << (int)g[i * k + j];
        const bool in = (mask >> lane) & 1u;
        const int pos = __popc(mask & ((1u << lane) - 1u));
        const double ys = blu_shfl(y, pos);
        const double ua = in ? ys : 0.0;
        double out = ua;
        if (mode == 1) {
            out = 0.0;
            for (int b = 0; b < N; ++b) {
                const double ub = blu_shfl(ua, b);
                if (lane < N) out += P[lane * N + b] * ub;
            }
        }
        if (lane < NP) rows[i * NP + lane] = lane < N ? out : 0.0;
    }
}
// hess[ik*Lq + iq] += a_ik . w_iq                               (hessKQ_c, cmisc.cpp:74-97)
__global__ void blu_l1_hess_kernel(double *hess, long long Lk, long long Lq, int NP, const double *A, const double *W)
{
    const long long total = Lk * Lq;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long ik = t / Lq, iq = t - ik * Lq;
        double s = 0.0;
        for (int a = 0; a < NP; ++a) s += A[ik * NP + a] * W[iq * NP + a];
        hess[t] += s;
    }
}

// ---- host drivers --------------------------------------------------------------------------
struct BluL1Bufs {
    std::vector<void *> ptrs;
    cudaError_t e = cudaSuccess;
    template <class T> T *up(const T *h, size_t n)
    {
        T *d = nullptr;
        if (e != cudaSuccess) return nullptr;
        e = cudaMalloc(&d, sizeof(T) * (n ? n : 1));
        if (e != cudaSuccess) return nullptr;
        ptrs.push_back(d);
        if (h && n) e = cudaMemcpy(d, h, sizeof(T) * n, cudaMemcpyHostToDevice);
        return d;
    }
    ~BluL1Bufs() { for (void *p : ptrs) cudaFree(p); }
};
static int blu_l1_done(BluL1Bufs &b, std::string &err, const char *what)
{
    if (b.e == cudaSuccess) b.e = cudaGetLastError();
    if (b.e == cudaSuccess) b.e = cudaDeviceSynchronize();
    if (b.e != cudaSuccess) { err = std::string(what) + ": " + cudaGetErrorString(b.e); return BLU_ERR_CUDA; }
    return BLU_OK;
}
static int blu_l1_check(int N, int k, long long Lk, const void *a, const void *b, const void *c, std::string &err)
{
    if (N < 1 || N > BLU_MAX_MODELS || k < 1 || k > N || Lk < 0 || !a || (Lk > 0 && (!b || !c))) {
        err = "bad Level-1 arguments"; return BLU_ERR_ARG;
    }
    return BLU_OK;
}
static int blu_l1_grid(long long items, int per_block) { return (int)std::max<long long>(1, std::min<long long>((items + per_block - 1) / per_block, 148 * 8)); }

static int blu_l1_psi(double *psi, int N, int k, int Lk, const int64_t *g, const double *inv, std::string &err)
{
    int rc = blu_l1_check(N, k, Lk, psi, g, inv, err);
    if (rc || Lk == 0) return rc;
    BluL1Bufs b;
    const size_t np = (size_t)N * N * Lk;
    double *d_psi = b.up(psi, np);
    long long *d_g = b.up((const long long *)g, (size_t)Lk * k);
    double *d_inv = b.up(inv, (size_t)Lk * k * k);
    if (b.e == cudaSuccess) blu_l1_psi_kernel<<<blu_l1_grid((long long)Lk * k * k, 256), 256>>>(d_psi, N, k, Lk, d_g, d_inv);
    rc = blu_l1_done(b, err, "assemble_psi_c");
    if (rc) return rc;
    b.e = cudaMemcpy(psi, d_psi, sizeof(double) * np, cudaMemcpyDeviceToHost);
    return blu_l1_done(b, err, "assemble_psi_c");
}
static int blu_l1_phi(double *PHI, int N, int k, int Lk, const double *md, const int64_t *mi, const int64_t *g, const double *inv, std::string &err)
{
    int rc = blu_l1_check(N, k, Lk, PHI, g, inv, err);
    if (rc || Lk == 0) return rc;
    if (!md && !mi) { err = "null m"; return BLU_ERR_ARG; }
    BluL1Bufs b;
    const int NN = N * N;
    const int grid = blu_l1_grid(Lk, BLU_L1_WARPS * 4);
    double *d_phi = b.up(PHI, (size_t)NN);
    double *d_md = md ? b.up(md, (size_t)Lk) : nullptr;
    long long *d_mi = mi ? b.up((const long long *)mi, (size_t)Lk) : nullptr;
    long long *d_g = b.up((const long long *)g, (size_t)Lk * k);
    double *d_inv = b.up(inv, (size_t)Lk * k * k);
    double *d_part = b.up((const double *)nullptr, (size_t)NN * grid);
    if (b.e == cudaSuccess) {
        blu_l1_phi_kernel<<<grid, BLU_L1_WARPS * 32, sizeof(double) * NN * BLU_L1_WARPS>>>(d_part, N, k, Lk, d_md, d_mi, d_g, d_inv);
        blu_l1_reduce_add_kernel<<<(NN + 255) / 256, 256>>>(d_phi, d_part, grid, NN);
    }
    rc = blu_l1_done(b, err, "objectiveK_c");
    if (rc) return rc;
    b.e = cudaMemcpy(PHI, d_phi, sizeof(double) * NN, cudaMemcpyDeviceToHost);
    return blu_l1_done(b, err, "objectiveK_c");
}
static int blu_l1_grad(double *grad, int N, int k, int Lk, const int64_t *g, const double *inv, const double *x, std::string &err)
{
    int rc = blu_l1_check(N, k, Lk, grad, g, inv, err);
    if (rc || Lk == 0) return rc;
    if (!x) { err = "null invPHI_0"; return BLU_ERR_ARG; }
    BluL1Bufs b;
    double *d_grad = b.up(grad, (size_t)Lk);
    long long *d_g = b.up((const long long *)g, (size_t)Lk * k);
    double *d_inv = b.up(inv, (size_t)Lk * k * k);
    double *d_x = b.up(x, (size_t)N);
    if (b.e == cudaSuccess) blu_l1_grad_kernel<<<blu_l1_grid(Lk, BLU_L1_WARPS), BLU_L1_WARPS * 32>>>(d_grad, k, Lk, d_g, d_inv, d_x);
    rc = blu_l1_done(b, err, "gradK_c");
    if (rc) return rc;
    b.e = cudaMemcpy(grad, d_grad, sizeof(double) * Lk, cudaMemcpyDeviceToHost);
    return blu_l1_done(b, err, "gradK_c");
}
static int blu_l1_cleanup(double *X, int N, int k, int Lk, const int64_t *g, const double *inv, const double *x, std::string &err)
{
    int rc = blu_l1_check(N, k, Lk, X, g, inv, err);
    if (rc || Lk == 0) return rc;
    if (!x) { err = "null invPHI_0"; return BLU_ERR_ARG; }
    BluL1Bufs b;
    double *d_X = b.up(X, (size_t)N * Lk);
    long long *d_g = b.up((const long long *)g, (size_t)Lk * k);
    double *d_inv = b.up(inv, (size_t)Lk * k * k);
    double *d_x = b.up(x, (size_t)N);
    if (b.e == cudaSuccess) blu_l1_cleanup_kernel<<<blu_l1_grid((long long)Lk * k, 256), 256>>>(d_X, k, Lk, d_g, d_inv, d_x);
    rc = blu_l1_done(b, err, "cleanupK_c");
    if (rc) return rc;
    b.e = cudaMemcpy(X, d_X, sizeof(double) * N * Lk, cudaMemcpyDeviceToHost);
    return blu_l1_done(b, err, "cleanupK_c");
}
static int blu_l1_hess(double *hess, int N, int k, int q, int Lk, int Lq, const int64_t *gk, const int64_t *gq,
                       const double *ik, const double *iq, const double *P, std::string &err)
{
    int rc = blu_l1_check(N, k, Lk, hess, gk, ik, err);
    if (rc) return rc;
    rc = blu_l1_check(N, q, Lq, hess, gq, iq, err);
    if (rc || Lk == 0 || Lq == 0) return rc;
    if (!P) { err = "null invPHI"; return BLU_ERR_ARG; }
    BluL1Bufs b;
    const int NP = 4 * ((N + 3) / 4);
    double *d_h = b.up(hess, (size_t)Lk * Lq);
    long long *d_gk = b.up((const long long *)gk, (size_t)Lk * k);
    long long *d_gq = b.up((const long long *)gq, (size_t)Lq * q);
    double *d_ik = b.up(ik, (size_t)Lk * k * k);
    double *d_iq = b.up(iq, (size_t)Lq * q * q);
    double *d_P = b.up(P, (size_t)N * N);
    double *d_A = b.up((const double *)nullptr, (size_t)Lk * NP);
    double *d_W = b.up((const double *)nullptr, (size_t)Lq * NP);
    if (b.e == cudaSuccess) {
        blu_l1_rows_kernel<<<blu_l1_grid(Lk, 8), 256>>>(d_A, N, NP, k, Lk, d_gk, d_ik, d_P, 0);
        blu_l1_rows_kernel<<<blu_l1_grid(Lq, 8), 256>>>(d_W, N, NP, q, Lq, d_gq, d_iq, d_P, 1);
        blu_l1_hess_kernel<<<blu_l1_grid((long long)Lk * Lq, 256), 256>>>(d_h, Lk, Lq, NP, d_A, d_W);
    }
    rc = blu_l1_done(b, err, "hessKQ_c");
    if (rc) return rc;
    b.e = cudaMemcpy(hess, d_h, sizeof(double) * (size_t)Lk * Lq, cudaMemcpyDeviceToHost);
    return blu_l1_done(b, err, "hessKQ_c");
}
