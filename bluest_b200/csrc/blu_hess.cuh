// blu_hess.cuh -- kernel (3b): dense Hessian  H = U S U^T = U V^T   (L x L, FP64).
//
// Replaces the K^2 calls of hessKQ_c (cmisc.cpp:74-97, a six-deep scalar loop with
// (sum_k C(N,k) k^2)^2 inner iterations) plus ``hess += hess.T`` (misc.py:497-503) by a rank-N
// product: 2 N L^2 flops, bounded by WRITING the 8 L^2 bytes of H (8.59 GB at N = 15).
//
// One CTA (4 warps) computes a 64 x 64 tile with FP64 tensor-core MMAs
// (mma.sync.m8n8k4.f64 -- there is no FP64 in tcgen05), operands read straight from the
// L2-resident U / V rows (the K dimension is only NP = 4*NCH <= 32, so there is no K loop to
// pipeline), accumulators staged through shared memory so that every global store instruction
// writes whole 128-byte lines (32 lanes x 16 B = one 512 B row segment), streaming (.cs).
// SYM = true: only tiles I <= J are computed; the tile is staged a second time, transposed, and
// written to (J, I) as well, so the tensor pipe does half the work and H is exactly symmetric
// bit for bit, as the reference's ``hess += hess.T`` makes it.  Diagonal tiles are symmetrised.
// SYM = false: rectangular row panel (U = the rank's own Lrows rows, V = all Lcols rows, H = the
// rank's (Lrows, ldH) panel), for the group-sharded multi-GPU path where a rank owns a row block.
//
// K-slot permutation: within a k-chunk the MMA's k index of lane (lane&3) is mapped to model
// column (lane&3)*NCH + kc, so a lane's NCH operands are contiguous in memory (32 B at N <= 16).
// A and B use the same map, so the contraction is unchanged.
#pragma once
#include "blu_common.cuh"

#define BLU_HT 64              // tile edge
#ifndef BLU_HSB
#define BLU_HSB 16             // super-block edge in tiles (rasterisation)
#endif
#define BLU_HLDN 72            // staging pitch, normal tile   (72*8 B: rows 2 apart never share a bank phase)
#define BLU_HLDT 66            // staging pitch, transposed tile
// Staging variants of the symmetric kernel (template parameter ONEBUF):
//   false: normal and transposed image staged side by side (70.6 KB, 3 CTAs per SM)
//   true : the transposed image is staged in the SAME buffer after the normal one has been stored
//          (36.9 KB and two more barriers per tile, 4 CTAs per SM)
#define BLU_HESS_SMEM1 (BLU_HT * BLU_HLDN * 8)
#define BLU_HESS_SMEM2 ((BLU_HT * BLU_HLDN + BLU_HT * BLU_HLDT) * 8)

__device__ __forceinline__ void blu_dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NCH, bool SYM, bool ONEBUF = false>
__global__ void __launch_bounds__(128, SYM ? (ONEBUF ? 4 : 3) : 5)
blu_hess_kernel(const double *__restrict__ U, const double *__restrict__ V, long long Lrows,
                long long Lcols, long long ldH, double *__restrict__ H, int nT, int tI0)
{
    constexpr int NP = 4 * NCH;
    extern __shared__ double hsm[];
    double *sN = hsm;
    double *sT = ONEBUF ? hsm : hsm + BLU_HT * BLU_HLDN;

    int I, J;
    if (SYM) {
        // Rasterisation: tiles are visited super-block by super-block (BLU_HSB x BLU_HSB tiles =
        // 1024 x 1024 entries), so that BOTH the (I,J) tile and its transposed image (J,I) land in a
        // compact 2-D region that is completely written within a short time window: the L2 then
        // evicts whole multi-KB row runs instead of isolated 512-byte chunks (DRAM page locality).
        // blockIdx.y = super-block pair (bi <= bj, row-major over the triangle), blockIdx.x = tile
        // inside the super-block.
        const int nB = (nT + BLU_HSB - 1) / BLU_HSB;
        const long long pid = blockIdx.y;
        const double nb = (double)nB;
        int bi = (int)floor(((2.0 * nb + 1.0) - sqrt((2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * (double)pid)) * 0.5);
        if (bi < 0) bi = 0;
        if (bi > nB - 1) bi = nB - 1;
        while ((long long)bi * nB - (long long)bi * (bi - 1) / 2 > pid) --bi;
        while ((long long)(bi + 1) * nB - (long long)(bi + 1) * bi / 2 <= pid) ++bi;
        const int bj = bi + (int)(pid - ((long long)bi * nB - (long long)bi * (bi - 1) / 2));
        I = bi * BLU_HSB + (int)(blockIdx.x / BLU_HSB);
        J = bj * BLU_HSB + (int)(blockIdx.x % BLU_HSB);
        if (I >= nT || J >= nT || I > J) return;
    } else {
        J = blockIdx.x;
        I = tI0 + blockIdx.y;
    }

    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = w >> 1, wx = w & 1;
    const int gq = lane >> 2, s = lane & 3;

    double a[4][NCH], b[4][NCH];
#pragma unroll
    for (int rb = 0; rb < 4; ++rb) {
        const double *up = U + ((long long)I * BLU_HT + wy * 32 + rb * 8 + gq) * NP + s * NCH;
        const double *vp = V + ((long long)J * BLU_HT + wx * 32 + rb * 8 + gq) * NP + s * NCH;
        if (NCH % 2 == 0) {
#pragma unroll
            for (int kc = 0; kc < NCH; kc += 2) {
                const double2 t = *reinterpret_cast<const double2 *>(up + kc);
                const double2 r = *reinterpret_cast<const double2 *>(vp + kc);
                a[rb][kc] = t.x; a[rb][kc + 1] = t.y;
                b[rb][kc] = r.x; b[rb][kc + 1] = r.y;
            }
        } else {
#pragma unroll
            for (int kc = 0; kc < NCH; ++kc) { a[rb][kc] = up[kc]; b[rb][kc] = vp[kc]; }
        }
    }
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

    // ---- epilogue: registers -> shared staging -> whole-line streaming stores -------------------
    const bool offdiag = SYM && (I != J);
    const bool diag = SYM && (I == J);
#pragma unroll
    for (int rb = 0; rb < 4; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            const int row = wy * 32 + rb * 8 + gq;
            const int col = wx * 32 + cb * 8 + 2 * s;
            *reinterpret_cast<double2 *>(sN + row * BLU_HLDN + col) = make_double2(c[rb][cb][0], c[rb][cb][1]);
            if (offdiag && !ONEBUF) {
                sT[col * BLU_HLDT + row] = c[rb][cb][0];
                sT[(col + 1) * BLU_HLDT + row] = c[rb][cb][1];
            }
        }
    __syncthreads();
    const long long gcolN = (long long)J * BLU_HT + 2 * lane;
    const long long gcolT = (long long)I * BLU_HT + 2 * lane;
    if (!diag) {
#pragma unroll 4
        for (int r = w; r < BLU_HT; r += 4) {
            const long long grow = (long long)I * BLU_HT + r;
            if (grow < Lrows && gcolN < ldH)
                __stcs(reinterpret_cast<double2 *>(H + grow * ldH + gcolN),
                       *reinterpret_cast<const double2 *>(sN + r * BLU_HLDN + 2 * lane));
        }
    } else {
        // diagonal tile: symmetrise so that H == H^T bit for bit (``hess += hess.T``, misc.py:503)
        for (int r = w; r < BLU_HT; r += 4) {
            const long long grow = (long long)I * BLU_HT + r;
            if (grow < Lrows && gcolN < ldH) {
                double2 v = *reinterpret_cast<const double2 *>(sN + r * BLU_HLDN + 2 * lane);
                v.x = 0.5 * (v.x + sN[(2 * lane) * BLU_HLDN + r]);
                v.y = 0.5 * (v.y + sN[(2 * lane + 1) * BLU_HLDN + r]);
                __stcs(reinterpret_cast<double2 *>(H + grow * ldH + gcolN), v);
            }
        }
    }
    if (offdiag) {
        if (ONEBUF) {
            __syncthreads();                             // normal image read out: the buffer is free
#pragma unroll
            for (int rb = 0; rb < 4; ++rb)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    const int row = wy * 32 + rb * 8 + gq;
                    const int col = wx * 32 + cb * 8 + 2 * s;
                    sT[col * BLU_HLDT + row] = c[rb][cb][0];
                    sT[(col + 1) * BLU_HLDT + row] = c[rb][cb][1];
                }
            __syncthreads();
        }
#pragma unroll 4
        for (int r = w; r < BLU_HT; r += 4) {
            const long long grow = (long long)J * BLU_HT + r;
            if (grow < Lcols && gcolT < ldH)
                __stcs(reinterpret_cast<double2 *>(H + grow * ldH + gcolT),
                       *reinterpret_cast<const double2 *>(sT + r * BLU_HLDT + 2 * lane));
        }
    }
}
