// hv_lab.cu -- lab harness for the Hessian-operator product kernels (blu_matvec.cuh) at 20 models.
// Runs reduce -> apply in the order of a real product (so the L2 state between the passes is the real one) for every
// candidate configuration (measured once more with a contiguous span of rows per CTA in the reduce pass: 63.6 vs 62.2 us), checks t and out against a plain reference kernel, prints the per-kernel event times.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Iinclude -Ibluest_b200/csrc -o tools/lab/bin/hv_lab tools/lab/hv_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <functional>
#include "blu_matvec.cuh"

// ---- first generation of the two kernels (round 1 .. mid round 2), kept here as the baseline of the comparison ----
#define BLU_HV_UNROLL 4
#define BLU_HVA_THREADS 128
// t = sum of `nparts` partial vectors (32 doubles each): BLU_HV_FOLD strided sub-sums per column combined
// in a fixed order (warps beyond BLU_HV_FOLD idle, so the association does not depend on the block size).
__device__ __forceinline__ void blu_hv_fold_gen1(const double *part, int nparts, double *sh, double *t)
{
    const int tid = threadIdx.x;
    const int c = tid & 31, q = tid >> 5;
    double s = 0.0;
    if (q < BLU_HV_FOLD)
        for (int b = q; b < nparts; b += BLU_HV_FOLD) s += __ldcg(part + (size_t)b * 32 + c);
    if (q < BLU_HV_FOLD) sh[tid] = s;
    __syncthreads();
    if (tid < 32) {
        double a = 0.0;
#pragma unroll
        for (int qq = 0; qq < BLU_HV_FOLD; ++qq) a += sh[qq * 32 + tid];
        t[tid] = a;
    }
    __syncthreads();
}

// part[b*32 + c] = sum over the rows r of CTA b of p[r] * U[r][c]   (c < NP <= 32); t_out = their sum.
// A thread owns two adjacent columns of a fixed row slot (16-byte loads, BLU_HV_UNROLL rows in flight).
template <int NP>
__global__ void __launch_bounds__(BLU_HV_THREADS)
blu_hv_reduce_gen1_kernel(const double *__restrict__ U, const double *__restrict__ p, long long lo, long long hi,
                     double *__restrict__ part, unsigned *__restrict__ ticket, double *__restrict__ t_out)
{
    constexpr int HC = NP / 2;                              // column pairs per row
    constexpr int RPP = BLU_HV_THREADS / HC;                // rows per pass of one CTA
    __shared__ double sh[BLU_HV_THREADS * 2];
    __shared__ double tfin[32];
    __shared__ bool last;
    const int tid = threadIdx.x;
    const int r = tid / HC, c2 = tid - r * HC;
    double a0[BLU_HV_UNROLL], a1[BLU_HV_UNROLL];
#pragma unroll
    for (int u = 0; u < BLU_HV_UNROLL; ++u) { a0[u] = 0.0; a1[u] = 0.0; }
    if (r < RPP) {
        const long long stride = (long long)gridDim.x * RPP;
        long long row = lo + (long long)blockIdx.x * RPP + r;
        for (; row + (BLU_HV_UNROLL - 1) * stride < hi; row += BLU_HV_UNROLL * stride) {
            double2 v[BLU_HV_UNROLL]; double pv[BLU_HV_UNROLL];
#pragma unroll
            for (int u = 0; u < BLU_HV_UNROLL; ++u) {
                const long long q = row + u * stride;
                v[u] = __ldg(reinterpret_cast<const double2 *>(U + q * NP) + c2);
                pv[u] = __ldg(p + q);
            }
#pragma unroll
            for (int u = 0; u < BLU_HV_UNROLL; ++u) { a0[u] = fma(pv[u], v[u].x, a0[u]); a1[u] = fma(pv[u], v[u].y, a1[u]); }
        }
        for (; row < hi; row += stride) {
            const double2 v = __ldg(reinterpret_cast<const double2 *>(U + row * NP) + c2);
            const double pv = __ldg(p + row);
            a0[0] = fma(pv, v.x, a0[0]); a1[0] = fma(pv, v.y, a1[0]);
        }
    }
    sh[2 * tid] = (a0[0] + a0[1]) + (a0[2] + a0[3]);
    sh[2 * tid + 1] = (a1[0] + a1[1]) + (a1[2] + a1[3]);
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        if (tid < NP)
            for (int q = 0; q < RPP; ++q) s += sh[2 * (q * HC + (tid >> 1)) + (tid & 1)];
        part[(size_t)blockIdx.x * 32 + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (tid == 0) *ticket = 0u;                              // ready for the next product
    blu_hv_fold_gen1(part, (int)gridDim.x, sh, tfin);
    if (tid < 32) t_out[tid] = tfin[tid];
}

// out[i] = u_i . s for i in [lo,hi), s = S t (S = 2 pinv(Phi), N x N; t: 32 doubles, zero beyond N).
// A warp owns 32 consecutive rows = one contiguous span of 32 NP doubles: coalesced loads, products
// staged in shared memory with row pitch NP + 1, then lane r sums row r.
template <int NP>
__global__ void __launch_bounds__(BLU_HVA_THREADS)
blu_hv_apply_gen1_kernel(const double *__restrict__ U, const double *__restrict__ S, int N, const double *__restrict__ t_in,
                    long long lo, long long hi, double *__restrict__ out, int reverse)
{
    __shared__ double t[32];
    __shared__ double prod[BLU_HVA_THREADS / 32][32 * (NP + 1)];
    const int tid = threadIdx.x;
    if (tid < 32) {
        double s = 0.0;
        if (tid < N)
            for (int b = 0; b < N; ++b) s = fma(__ldg(S + tid * N + b), __ldcg(t_in + b), s);
        t[tid] = s;
    }
    __syncthreads();
    const int w = tid >> 5, lane = tid & 31;
    constexpr int NWARP = BLU_HVA_THREADS / 32;
    double *pw = prod[w];
    // reverse: walk the 32-row blocks from the END of the range.  The reduce pass that precedes this kernel streamed U
    // front to back, so the last ~100 MB of it are still in the 126 MB L2: reading backwards turns most of this
    // pass's DRAM traffic into L2 hits when U (168 MB at 20 models) does not fit -- and leaves the FRONT of U in L2
    // for the next product's reduce pass.
    const long long nblk = (hi - lo + 31) / 32;
    for (long long blk = (long long)blockIdx.x * NWARP + w; blk < nblk; blk += (long long)gridDim.x * NWARP) {
        const long long r0 = lo + (reverse ? (nblk - 1 - blk) : blk) * 32;
        const int nrow = (int)((hi - r0) < 32 ? (hi - r0) : 32);
        const double *ub = U + r0 * NP;
        if (nrow == 32) {
#pragma unroll
            for (int it = 0; it < NP; ++it) {
                const int idx = it * 32 + lane;
                pw[idx + idx / NP] = __ldg(ub + idx) * t[idx % NP];
            }
        } else {
            for (int idx = lane; idx < nrow * NP; idx += 32) pw[idx + idx / NP] = __ldg(ub + idx) * t[idx % NP];
        }
        __syncwarp();
        if (lane < nrow) {
            const double *pr = pw + lane * (NP + 1);
            double b0 = 0.0, b1 = 0.0;
#pragma unroll
            for (int c = 0; c < NP; c += 2) { b0 += pr[c]; b1 += pr[c + 1]; }
            out[r0 + lane] = b0 + b1;
        }
        __syncwarp();
    }
}



#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void ref_reduce(const double *U, const double *p, long long L, int NP, double *t)
{
    // one CTA per column, plain strided sum (reference for the check only)
    const int c = blockIdx.x;
    __shared__ double sh[256];
    double s = 0.0;
    for (long long r = threadIdx.x; r < L; r += blockDim.x) s += p[r] * U[r * NP + c];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) t[c] = sh[0];
}
__global__ void ref_apply(const double *U, const double *S, int N, int NP, const double *t, long long L, double *out)
{
    __shared__ double s[32];
    if (threadIdx.x < 32) { double a = 0.0; if (threadIdx.x < N) for (int b = 0; b < N; ++b) a += S[threadIdx.x * N + b] * t[b]; s[threadIdx.x] = a; }
    __syncthreads();
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < L; r += (long long)gridDim.x * blockDim.x) {
        double a = 0.0;
        for (int c = 0; c < NP; ++c) a += U[r * NP + c] * s[c];
        out[r] = a;
    }
}

struct Cfg { const char *name; std::function<void()> reduce, apply; };

int main(int argc, char **argv)
{
    constexpr int NP = 20;
    const int N = 20;
    const long long L = 1048575;
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int nsm = pr.multiProcessorCount;
    std::vector<double> hU((size_t)L * NP), hp(L), hS(N * N);
    srand(1);
    for (auto &v : hU) v = (rand() / (double)RAND_MAX) - 0.5;
    for (auto &v : hp) v = (rand() / (double)RAND_MAX) - 0.5;
    for (auto &v : hS) v = (rand() / (double)RAND_MAX) - 0.5;
    double *U, *p, *S, *part, *t, *tref, *out, *oref;
    unsigned *ticket;
    CK(cudaMalloc(&U, sizeof(double) * hU.size())); CK(cudaMalloc(&p, sizeof(double) * L)); CK(cudaMalloc(&S, sizeof(double) * N * N));
    CK(cudaMalloc(&part, sizeof(double) * 32 * (size_t)nsm * 16)); CK(cudaMalloc(&t, sizeof(double) * 32)); CK(cudaMalloc(&tref, sizeof(double) * 32));
    CK(cudaMalloc(&out, sizeof(double) * L)); CK(cudaMalloc(&oref, sizeof(double) * L)); CK(cudaMalloc(&ticket, 8));
    CK(cudaMemcpy(U, hU.data(), sizeof(double) * hU.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p, hp.data(), sizeof(double) * L, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(S, hS.data(), sizeof(double) * N * N, cudaMemcpyHostToDevice));
    CK(cudaMemset(ticket, 0, 8)); CK(cudaMemset(tref, 0, 256)); CK(cudaMemset(t, 0, 256));
    ref_reduce<<<NP, 256>>>(U, p, L, NP, tref);
    ref_apply<<<nsm * 8, 256>>>(U, S, N, NP, tref, L, oref);
    CK(cudaDeviceSynchronize());
    std::vector<double> htref(32), horef(L), ht(32), hout(L);
    CK(cudaMemcpy(htref.data(), tref, 256, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(horef.data(), oref, sizeof(double) * L, cudaMemcpyDeviceToHost));

    const int rpp = BLU_HV_THREADS / (NP / 2);
    const int grid_r1 = (int)std::min<long long>((L + (long long)rpp * BLU_HV_UNROLL - 1) / ((long long)rpp * BLU_HV_UNROLL), (long long)nsm * 8);
    const int grid_a1 = (int)std::min<long long>((L + BLU_HVA_THREADS - 1) / BLU_HVA_THREADS, (long long)nsm * 8);
    auto R1 = [&] { blu_hv_reduce_gen1_kernel<NP><<<grid_r1, BLU_HV_THREADS>>>(U, p, 0, L, part, ticket, t); };
    auto A1 = [&] { blu_hv_apply_gen1_kernel<NP><<<grid_a1, BLU_HVA_THREADS>>>(U, S, N, t, 0, L, out, 1); };
#define R2(UN, MB) [&] { blu_hv_reduce_kernel<NP, UN, MB><<<nsm * MB, BLU_HV_THREADS>>>(U, p, 0, L, part, ticket, t); }
#define A2(W, MB, GM) [&] { blu_hv_apply_kernel<NP, W, MB><<<nsm * GM, W * 32>>>(U, S, N, t, 0, L, out, 1); }
    std::vector<Cfg> cfgs = {
        {"gen1 reduce / gen1 apply", R1, A1},
        {"reduce<UN4,MB4> / apply<W4,MB8>", R2(4, 4), A2(4, 8, 8)},
        {"reduce2<UN8,MB2,stride> / apply2<W8,MB4>", R2(8, 2), A2(8, 4, 4)},
        {"reduce2<UN2,MB6,stride> / apply2<W4,MB12>", R2(2, 6), A2(4, 12, 12)},
        {"reduce2<UN4,MB3,stride> / apply2<W4,MB16>", R2(4, 3), A2(4, 16, 16)},
        {"reduce2<UN8,MB3,stride> / apply2<W2,MB16>", R2(8, 3), A2(2, 16, 16)},
        {"reduce2<UN4,MB4,stride> / gen1 apply", R2(4, 4), A1},
        {"gen1 reduce / apply2<W4,MB8>", R1, A2(4, 8, 8)},
    };
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    const int reps = 30;
    for (auto &c : cfgs) {
        for (int w = 0; w < 3; ++w) { c.reduce(); c.apply(); }
        CK(cudaDeviceSynchronize());
        double tr = 0.0, ta = 0.0;
        for (int i = 0; i < reps; ++i) {
            CK(cudaEventRecord(e0)); c.reduce(); CK(cudaEventRecord(e1)); c.apply(); CK(cudaEventRecord(e2));
            CK(cudaEventSynchronize(e2));
            float a, b; CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
            tr += a; ta += b;
        }
        CK(cudaGetLastError());
        // back to back without events in between (what a CG iteration sees)
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) { c.reduce(); c.apply(); }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float tot; CK(cudaEventElapsedTime(&tot, e0, e1));
        CK(cudaMemcpy(ht.data(), t, 256, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hout.data(), out, sizeof(double) * L, cudaMemcpyDeviceToHost));
        double et = 0.0, eo = 0.0, mt = 0.0, mo = 0.0;
        for (int i = 0; i < NP; ++i) { et = fmax(et, fabs(ht[i] - htref[i])); mt = fmax(mt, fabs(htref[i])); }
        for (long long i = 0; i < L; ++i) { eo = fmax(eo, fabs(hout[i] - horef[i])); mo = fmax(mo, fabs(horef[i])); }
        const double bytes = 8.0 * NP * L;
        printf("%-46s reduce %6.1f us (%4.2f TB/s)  apply %6.1f us (%4.2f TB/s)  product %6.1f us  err t %.1e out %.1e\n", c.name,
               tr / reps * 1e3, (bytes + 8.0 * L) / (tr / reps * 1e-3) * 1e-12, ta / reps * 1e3, (bytes + 8.0 * L) / (ta / reps * 1e-3) * 1e-12,
               tot / reps * 1e3, et / mt, eo / mo);
    }
    return 0;
}
