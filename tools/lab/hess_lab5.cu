// tools/lab/hess_lab5.cu -- strip variant: one CTA walks TJ consecutive J tiles of one row of tiles,
// keeping the U operands and prefetching the next V operands during the epilogue (lab only).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "blu_hess.cuh"
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template <int NCH, int TJ>
__global__ void __launch_bounds__(128, 3)
strip_kernel(const double *__restrict__ U, const double *__restrict__ V, long long L, long long ldH, double *__restrict__ H, int nT)
{
    constexpr int NP = 4 * NCH;
    extern __shared__ double hsm[];
    double *sN = hsm;
    double *sT = hsm + BLU_HT * BLU_HLDN;
    const int nB = (nT + BLU_HSB - 1) / BLU_HSB;
    const long long pid = blockIdx.y;
    const double nb = (double)nB;
    int bi = (int)floor(((2.0 * nb + 1.0) - sqrt((2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * (double)pid)) * 0.5);
    if (bi < 0) bi = 0;
    if (bi > nB - 1) bi = nB - 1;
    while ((long long)bi * nB - (long long)bi * (bi - 1) / 2 > pid) --bi;
    while ((long long)(bi + 1) * nB - (long long)(bi + 1) * bi / 2 <= pid) ++bi;
    const int bj = bi + (int)(pid - ((long long)bi * nB - (long long)bi * (bi - 1) / 2));
    const int I = bi * BLU_HSB + (int)(blockIdx.x / (BLU_HSB / TJ));
    int J = bj * BLU_HSB + (int)(blockIdx.x % (BLU_HSB / TJ)) * TJ;
    int Jend = J + TJ; if (Jend > nT) Jend = nT;
    if (J < I) J = I;
    if (I >= nT || J >= Jend) return;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = w >> 1, wx = w & 1;
    const int gq = lane >> 2, s = lane & 3;
    double a[4][NCH], b[4][NCH];
#pragma unroll
    for (int rb = 0; rb < 4; ++rb) {
        const double *up = U + ((long long)I * BLU_HT + wy * 32 + rb * 8 + gq) * NP + s * NCH;
        const double *vp = V + ((long long)J * BLU_HT + wx * 32 + rb * 8 + gq) * NP + s * NCH;
#pragma unroll
        for (int kc = 0; kc < NCH; kc += 2) {
            const double2 t = *reinterpret_cast<const double2 *>(up + kc);
            const double2 r = *reinterpret_cast<const double2 *>(vp + kc);
            a[rb][kc] = t.x; a[rb][kc + 1] = t.y; b[rb][kc] = r.x; b[rb][kc + 1] = r.y;
        }
    }
    for (; J < Jend; ++J) {
        double c[4][4][2];
#pragma unroll
        for (int rb = 0; rb < 4; ++rb)
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) { c[rb][cb][0] = 0.0; c[rb][cb][1] = 0.0; }
#pragma unroll
        for (int kc = 0; kc < NCH; ++kc)
#pragma unroll
            for (int rb = 0; rb < 4; ++rb)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) blu_dmma(c[rb][cb][0], c[rb][cb][1], a[rb][kc], b[cb][kc]);
        if (J + 1 < Jend) {                    // next V operands: in flight during the epilogue
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) {
                const double *vp = V + ((long long)(J + 1) * BLU_HT + wx * 32 + rb * 8 + gq) * NP + s * NCH;
#pragma unroll
                for (int kc = 0; kc < NCH; kc += 2) { const double2 r = *reinterpret_cast<const double2 *>(vp + kc); b[rb][kc] = r.x; b[rb][kc + 1] = r.y; }
            }
        }
        const bool offdiag = (I != J);
#pragma unroll
        for (int rb = 0; rb < 4; ++rb)
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                const int row = wy * 32 + rb * 8 + gq;
                const int col = wx * 32 + cb * 8 + 2 * s;
                *reinterpret_cast<double2 *>(sN + row * BLU_HLDN + col) = make_double2(c[rb][cb][0], c[rb][cb][1]);
                if (offdiag) { sT[col * BLU_HLDT + row] = c[rb][cb][0]; sT[(col + 1) * BLU_HLDT + row] = c[rb][cb][1]; }
            }
        __syncthreads();
        const long long gcolN = (long long)J * BLU_HT + 2 * lane;
        const long long gcolT = (long long)I * BLU_HT + 2 * lane;
#pragma unroll 4
        for (int r = w; r < BLU_HT; r += 4) {
            const long long grow = (long long)I * BLU_HT + r;
            if (grow < L && gcolN < ldH) __stcs(reinterpret_cast<double2 *>(H + grow * ldH + gcolN), *reinterpret_cast<const double2 *>(sN + r * BLU_HLDN + 2 * lane));
        }
        if (offdiag) {
#pragma unroll 4
            for (int r = w; r < BLU_HT; r += 4) {
                const long long grow = (long long)J * BLU_HT + r;
                if (grow < L && gcolT < ldH) __stcs(reinterpret_cast<double2 *>(H + grow * ldH + gcolT), *reinterpret_cast<const double2 *>(sT + r * BLU_HLDT + 2 * lane));
            }
        }
        __syncthreads();                       // staging buffers free for the next tile
    }
}

template <int TJ> void run(const double *U, const double *V, long long L, long long ldH, double *H, int nT, const char *name)
{
    const int nB = (nT + BLU_HSB - 1) / BLU_HSB;
    CK(cudaFuncSetAttribute(strip_kernel<4, TJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLU_HESS_SMEM));
    dim3 grid(BLU_HSB * (BLU_HSB / TJ), nB * (nB + 1) / 2);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) strip_kernel<4, TJ><<<grid, 128, BLU_HESS_SMEM>>>(U, V, L, ldH, H, nT);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) strip_kernel<4, TJ><<<grid, 128, BLU_HESS_SMEM>>>(U, V, L, ldH, H, nT);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 20;
    printf("%s: mean %.4f ms  %.0f GB/s\n", name, ms, 8.0 * L * L / ms * 1e-6);
}

int main()
{
    const long long L = 32767, ldH = 32768; const int NP = 16;
    const int nT = (int)((L + 63) / 64);
    const long long Lpad = (long long)nT * 64 + 64;
    std::vector<double> hU(Lpad * NP), hV(Lpad * NP);
    for (size_t i = 0; i < hU.size(); ++i) { hU[i] = (rand() % 1000) * 1e-3; hV[i] = (rand() % 1000) * 1e-3; }
    double *U, *V, *H, *H2;
    CK(cudaMalloc(&U, hU.size() * 8)); CK(cudaMalloc(&V, hV.size() * 8)); CK(cudaMalloc(&H, (size_t)L * ldH * 8)); CK(cudaMalloc(&H2, (size_t)L * ldH * 8));
    CK(cudaMemcpy(U, hU.data(), hU.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(V, hV.data(), hV.size() * 8, cudaMemcpyHostToDevice));
    const int nB = (nT + BLU_HSB - 1) / BLU_HSB;
    CK(cudaFuncSetAttribute(blu_hess_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLU_HESS_SMEM));
    dim3 grid(BLU_HSB * BLU_HSB, nB * (nB + 1) / 2);
    CK(cudaMemset(H, 0, (size_t)L * ldH * 8)); CK(cudaMemset(H2, 0, (size_t)L * ldH * 8));
    blu_hess_kernel<4, true><<<grid, 128, BLU_HESS_SMEM>>>(U, V, L, L, ldH, H, nT, 0);
    run<4>(U, V, L, ldH, H2, nT, "strip TJ=4 ");
    CK(cudaDeviceSynchronize());
    std::vector<double> r1(ldH), r2(ldH); double maxd = 0;
    for (long long row : {0LL, 63LL, 64LL, 1000LL, 20000LL, 32766LL}) {
        CK(cudaMemcpy(r1.data(), H + row * ldH, ldH * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r2.data(), H2 + row * ldH, ldH * 8, cudaMemcpyDeviceToHost));
        for (long long c = 0; c < L; ++c) if (c / 64 != row / 64) { double d = fabs(r1[c] - r2[c]); if (d > maxd) maxd = d; }
    }
    printf("max abs diff off-diagonal tiles: %g\n", maxd);
    run<1>(U, V, L, ldH, H2, nT, "strip TJ=1 ");
    run<2>(U, V, L, ldH, H2, nT, "strip TJ=2 ");
    run<4>(U, V, L, ldH, H2, nT, "strip TJ=4 ");
    run<8>(U, V, L, ldH, H2, nT, "strip TJ=8 ");
    run<16>(U, V, L, ldH, H2, nT, "strip TJ=16");
    return 0;
}
