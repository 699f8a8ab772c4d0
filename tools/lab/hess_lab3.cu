// tools/hess_lab3.cu -- times the PRODUCT Hessian kernel for a given -DBLU_HSB (lab only).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "blu_hess.cuh"
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
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
    const int nB = (nT + BLU_HSB - 1) / BLU_HSB;
    CK(cudaFuncSetAttribute(blu_hess_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLU_HESS_SMEM));
    dim3 grid(BLU_HSB * BLU_HSB, nB * (nB + 1) / 2);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) blu_hess_kernel<4, true><<<grid, 128, BLU_HESS_SMEM>>>(U, V, L, L, ldH, H, nT, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) blu_hess_kernel<4, true><<<grid, 128, BLU_HESS_SMEM>>>(U, V, L, L, ldH, H, nT, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 20;
    printf("BLU_HSB=%d: mean %.4f ms  %.0f GB/s\n", BLU_HSB, ms, 8.0 * L * L / ms * 1e-6);
    return 0;
}
