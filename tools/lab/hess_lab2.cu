// tools/hess_lab2.cu -- persistent-CTA variant of the Hessian kernel with operand prefetch (lab only).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "blu_hess.cuh"
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void tile_of(long long t, int nT, int &I, int &J)
{
    // t enumerates (super-block pair, tile in block) exactly like the product kernel's 2-D grid
    const int nB = (nT + BLU_HSB - 1) / BLU_HSB;
    const long long pid = t / (BLU_HSB * BLU_HSB);
    const int x = (int)(t % (BLU_HSB * BLU_HSB));
    const double nb = (double)nB;
    int bi = (int)floor(((2.0 * nb + 1.0) - sqrt((2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * (double)pid)) * 0.5);
    if (bi < 0) bi = 0;
    if (bi > nB - 1) bi = nB - 1;
    while ((long long)bi * nB - (long long)bi * (bi - 1) / 2 > pid) --bi;
    while ((long long)(bi + 1) * nB - (long long)(bi + 1) * bi / 2 <= pid) ++bi;
    const int bj = bi + (int)(pid - ((long long)bi * nB - (long long)bi * (bi - 1) / 2));
    I = bi * BLU_HSB + x / BLU_HSB;
    J = bj * BLU_HSB + x % BLU_HSB;
}

template <int NCH>
__device__ __forceinline__ void load_ops(const double *__restrict__ U, const double *__restrict__ V, int I, int J, int wy, int wx, int gq, int s,
                                         double (&a)[4][NCH], double (&b)[4][NCH])
{
    constexpr int NP = 4 * NCH;
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
}

template <int NCH>
__global__ void __launch_bounds__(128, 3)
persist_kernel(const double *__restrict__ U, const double *__restrict__ V, long long L, long long ldH, double *__restrict__ H, int nT, long long ntiles)
{
    extern __shared__ double hsm[];
    double *sN = hsm;
    double *sT = hsm + BLU_HT * BLU_HLDN;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = w >> 1, wx = w & 1;
    const int gq = lane >> 2, s = lane & 3;
    double a[4][NCH], b[4][NCH];
    long long t = blockIdx.x;
    int I = 0, J = 0;
    bool valid = false;
    // find first valid tile
    for (; t < ntiles; t += gridDim.x) { tile_of(t, nT, I, J); if (I < nT && J < nT && I <= J) { valid = true; break; } }
    if (valid) load_ops<NCH>(U, V, I, J, wy, wx, gq, s, a, b);
    while (valid) {
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
        const int Ic = I, Jc = J;
        // next tile + operand prefetch (in flight during staging and stores)
        valid = false;
        for (t += gridDim.x; t < ntiles; t += gridDim.x) { tile_of(t, nT, I, J); if (I < nT && J < nT && I <= J) { valid = true; break; } }
        if (valid) load_ops<NCH>(U, V, I, J, wy, wx, gq, s, a, b);
        const bool offdiag = (Ic != Jc);
        __syncthreads();      // previous tile's readers are done with the staging buffers
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
        const long long gcolN = (long long)Jc * BLU_HT + 2 * lane;
        const long long gcolT = (long long)Ic * BLU_HT + 2 * lane;
#pragma unroll 4
        for (int r = w; r < BLU_HT; r += 4) {
            const long long grow = (long long)Ic * BLU_HT + r;
            if (grow < L && gcolN < ldH) __stcs(reinterpret_cast<double2 *>(H + grow * ldH + gcolN), *reinterpret_cast<const double2 *>(sN + r * BLU_HLDN + 2 * lane));
        }
        if (offdiag) {
#pragma unroll 4
            for (int r = w; r < BLU_HT; r += 4) {
                const long long grow = (long long)Jc * BLU_HT + r;
                if (grow < L && gcolT < ldH) __stcs(reinterpret_cast<double2 *>(H + grow * ldH + gcolT), *reinterpret_cast<const double2 *>(sT + r * BLU_HLDT + 2 * lane));
            }
        }
    }
}

template <class F> float timeit(F f, int reps)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f, sum = 0;
    for (int r = 0; r < reps; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; sum += ms; }
    printf("   (mean %.3f ms) ", sum / reps);
    return best;
}

int main()
{
    const long long L = 32767, ldH = 32768; const int NP = 16;
    const int nT = (int)((L + 63) / 64);
    const long long Lpad = (long long)nT * 64 + 64;
    std::vector<double> hU(Lpad * NP), hV(Lpad * NP);
    for (size_t i = 0; i < hU.size(); ++i) { hU[i] = (rand() % 1000) * 1e-3; hV[i] = (rand() % 1000) * 1e-3; }
    double *U, *V, *H;
    CK(cudaMalloc(&U, hU.size() * 8)); CK(cudaMalloc(&V, hV.size() * 8)); CK(cudaMalloc(&H, (size_t)L * ldH * 8));
    CK(cudaMemcpy(U, hU.data(), hU.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(V, hV.data(), hV.size() * 8, cudaMemcpyHostToDevice));
    const double bytes = 8.0 * L * L;
    const int nB = (nT + BLU_HSB - 1) / BLU_HSB;
    const long long ntiles = (long long)BLU_HSB * BLU_HSB * (nB * (nB + 1) / 2);
    const int smem = (BLU_HT * BLU_HLDN + BLU_HT * BLU_HLDT) * 8;
    CK(cudaFuncSetAttribute(persist_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(blu_hess_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLU_HESS_SMEM));
    {
        dim3 grid(BLU_HSB * BLU_HSB, nB * (nB + 1) / 2);
        float ms = timeit([&] { blu_hess_kernel<4, true><<<grid, 128, BLU_HESS_SMEM>>>(U, V, L, L, ldH, H, nT, 0); }, 10);
        printf("product kernel: best %.3f ms  %.0f GB/s\n", ms, bytes / ms * 1e-6);
    }
    for (int cps = 2; cps <= 3; ++cps) {
        float ms = timeit([&] { persist_kernel<4><<<148 * cps, 128, smem>>>(U, V, L, ldH, H, nT, ntiles); }, 10);
        printf("persistent x%d: best %.3f ms  %.0f GB/s\n", cps, ms, bytes / ms * 1e-6);
    }
    return 0;
}
