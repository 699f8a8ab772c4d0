// blu_common.cuh -- shared declarations for libbluest_b200 (sm_100a, FP64).
//
// HBM data layout of a context (all sizes for N models, size classes k = 1..K):
//   gidx   uint8   concatenation over classes of the (Lk,k) model-id tables (1 byte per id, N <= 32)
//   gmask  uint32  one membership bitmask per group, flat enumeration order (size-major)
//   cinv   double  per-group inverse covariances, PACKED UPPER TRIANGLE, row-major inside a group:
//                  element (j,l), j<=l, of group i of class k lives at
//                  coff[k] + i*T_k + j*k - j(j-1)/2 + (l-j),  T_k = k(k+1)/2.
//                  Class blocks start on 128-byte boundaries.  Half the bytes of the reference's
//                  (Lk,k,k) arrays (sap.py:78) and contiguous in the order the kernels stream it.
//   U, V   double  (Lpad, NP) row-major, NP = 4*ceil(N/4): row i holds u_i scattered to model
//                  positions (zeros elsewhere) / v_i = 2 pinv(Phi) u_i.  128 B per group at N<=16.
//   H      double  (L, ldH) row-major dense Hessian with row pitch ldH = 16*ceil(L/16) so every
//                  row starts on a 128-byte line.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "bluest_b200.h"

#define BLU_WARP 32
#define BLU_MAX_MODELS_C 32
#define BLU_FULL 0xffffffffu

struct BluClass {
    int k;            // group size of this class
    int T;            // k(k+1)/2 packed entries per group
    long long Lk;     // groups in the class
    long long goff;   // first flat group index of the class
    long long ioff;   // offset of the class in gidx (bytes)
    long long coff;   // offset of the class in cinv (doubles)
    int lutoff;       // offset of the class's (j,l) table in the LUT
    int plutoff;      // offset (32-bit words) of the class's bank-aware run table (Phi kernel, blu_phi.cuh)
    int psteps;       // 32-lane steps per group in that table
    int pad;
};

__host__ __device__ __forceinline__ int blu_tri(int k) { return k * (k + 1) / 2; }
// packed index of (j,l), j<=l, in a k x k upper triangle stored row-major
__host__ __device__ __forceinline__ int blu_pk(int k, int j, int l) { return j * k - j * (j - 1) / 2 + (l - j); }

__device__ __forceinline__ double blu_shfl(double v, int src, int width = 32) {
    return __shfl_sync(BLU_FULL, v, src, width);
}
__device__ __forceinline__ double blu_warp_sum(double v) {
    // fixed xor tree: the same association every run (deterministic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BLU_FULL, v, o);
    return v;
}
