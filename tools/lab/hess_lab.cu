// tools/hess_lab.cu -- ablation harness for the dense-Hessian kernel (not part of the product).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Iinclude -Ibluest_b200/csrc -o tools/hess_lab tools/hess_lab.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "blu_hess.cuh"
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

// ABL: 0 = full kernel, 1 = no DMMA (operands loaded, cheap combine), 2 = no loads and no DMMA
template <int NCH, int ABL, int NTHR_MINB>
__global__ void __launch_bounds__(128, NTHR_MINB)
lab_kernel(const double *__restrict__ U, const double *__restrict__ V, long long L, long long ldH, double *__restrict__ H, int nT)
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
    const int I = bi * BLU_HSB + (int)(blockIdx.x / BLU_HSB);
    const int J = bj * BLU_HSB + (int)(blockIdx.x % BLU_HSB);
    if (I >= nT || J >= nT || I > J) return;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = w >> 1, wx = w & 1;
    const int gq = lane >> 2, s = lane & 3;
    double c[4][4][2];
#pragma unroll
    for (int rb = 0; rb < 4; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) { c[rb][cb][0] = (double)(I + rb); c[rb][cb][1] = (double)(J + cb); }
    if (ABL < 2) {
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
        if (ABL == 0) {
#pragma unroll
            for (int kc = 0; kc < NCH; ++kc)
#pragma unroll
                for (int rb = 0; rb < 4; ++rb)
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) blu_dmma(c[rb][cb][0], c[rb][cb][1], a[rb][kc], b[cb][kc]);
        } else {
#pragma unroll
            for (int rb = 0; rb < 4; ++rb)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) { c[rb][cb][0] += a[rb][0] + b[cb][1] + a[rb][2] + b[cb][3]; c[rb][cb][1] += a[rb][1] + b[cb][0] + a[rb][3] + b[cb][2]; }
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
}

// pure tile-pattern writer: no smem, each warp writes rows of a 64x64 tile (normal only), full grid
__global__ void __launch_bounds__(128) tile_writer(long long L, long long ldH, double *__restrict__ H, int nT)
{
    const int I = blockIdx.y, J = blockIdx.x;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gcol = (long long)J * BLU_HT + 2 * lane;
    const double2 v = make_double2((double)I, (double)J);
#pragma unroll 4
    for (int r = w; r < BLU_HT; r += 4) {
        const long long grow = (long long)I * BLU_HT + r;
        if (grow < L && gcol < ldH) __stcs(reinterpret_cast<double2 *>(H + grow * ldH + gcol), v);
    }
}

template <class F> float timeit(F f, int reps)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
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
    dim3 grid(BLU_HSB * BLU_HSB, nB * (nB + 1) / 2);
    const int smem = (BLU_HT * BLU_HLDN + BLU_HT * BLU_HLDT) * 8;
#define RUN(ABL, MINB) { CK(cudaFuncSetAttribute(lab_kernel<4, ABL, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      float ms = timeit([&] { lab_kernel<4, ABL, MINB><<<grid, 128, smem>>>(U, V, L, ldH, H, nT); }, 5); \
      printf("lab ABL=%d minb=%d: %.3f ms  %.0f GB/s\n", ABL, MINB, ms, bytes / ms * 1e-6); }
    RUN(0, 3) RUN(1, 3) RUN(2, 3) RUN(2, 1)
    { float ms = timeit([&] { tile_writer<<<dim3(nT, nT), 128>>>(L, ldH, H, nT); }, 5); printf("tile_writer full grid: %.3f ms %.0f GB/s\n", ms, bytes / ms * 1e-6); }
    { float ms = timeit([&] { CK(cudaMemsetAsync(H, 0, (size_t)L * ldH * 8)); }, 3); printf("memset: %.3f ms %.0f GB/s\n", ms, bytes / ms * 1e-6); }
    return 0;
}
