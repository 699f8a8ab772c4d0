// blu_soa_tu.cu -- separate translation unit for the lane-per-group gradient / U kernels
// (blu_soa.cuh): 2 x 32 fully unrolled per-group-size tile routines.  Host launch wrappers only.
#include "blu_soa.cuh"
#include "blu_launch.h"

cudaError_t blu_launch_grad_soa(bool with_u, int grid, cudaStream_t stream, const BluClass *cls, int ncls, int N, int NP, int K,
                                const BluTile *tiles, int ntiles, const double *soa, const long long *soff, const unsigned *gmask,
                                const double *xrow, long long lo, long long hi, double *grad, double *U)
{
    const size_t smem = blu_soa_smem_bytes(K, with_u, N, ncls, with_u ? BLU_SOA_NS_U : BLU_SOA_NS_GRAD);
    cudaError_t e;
    if (with_u) {
        e = cudaFuncSetAttribute(blu_grad_soa_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        blu_grad_soa_kernel<true><<<grid, BLU_SOA_WARPS * 32, smem, stream>>>(cls, ncls, N, NP, K, tiles, ntiles, soa, soff, gmask, xrow, lo, hi, grad, U);
    } else {
        e = cudaFuncSetAttribute(blu_grad_soa_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        blu_grad_soa_kernel<false><<<grid, BLU_SOA_WARPS * 32, smem, stream>>>(cls, ncls, N, NP, K, tiles, ntiles, soa, soff, gmask, xrow, lo, hi, grad, nullptr);
    }
    return cudaGetLastError();
}

cudaError_t blu_launch_soa_build(int grid, cudaStream_t stream, const double *cinv, long long Lk, int T, double *soa)
{
    blu_soa_build_kernel<<<grid, 256, 0, stream>>>(cinv, Lk, T, soa);
    return cudaGetLastError();
}

cudaError_t blu_launch_v_from_u(int grid, cudaStream_t stream, const double *U, const double *S, int N, int NP, long long lo, long long hi, double *V)
{
    blu_v_from_u_kernel<<<grid, 256, 0, stream>>>(U, S, N, NP, lo, hi, V);
    return cudaGetLastError();
}
