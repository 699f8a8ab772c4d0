// tools/lab/phi_stream_lab.cu -- lab harness (round 2) for the TARGET-MAJOR Phi stream.
// One-time layout: the packed-inverse entries as (value, u16 local group index) pairs sorted by
// (block of G consecutive groups, target (a,b), group).  Phi[t] = sum over a segment of value * m[group]:
// an SpMV with the m block in shared memory, lane-private accumulators acc[t][lane], no scatter.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o phi_stream_lab tools/lab/phi_stream_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cmath>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e_=(x); if(e_!=cudaSuccess){printf("CUDA error %s at line %d\n",cudaGetErrorString(e_),__LINE__); exit(1);} }while(0)

struct Item { int blk, t0, t1, pad; };

template <int UNR, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
phi_stream_kernel(const double *__restrict__ pv, const uint16_t *__restrict__ pi, const long long *__restrict__ seg,
                  const Item *__restrict__ items, const int *__restrict__ cta_off, int NT, int G, long long L,
                  const double *__restrict__ m, double *__restrict__ part)
{
    extern __shared__ double sm[];
    double *acc = sm;                       // [NT][32]
    double *mb = sm + (size_t)NT * 32;      // [2][G]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < NT * 32; t += blockDim.x) acc[t] = 0.0;
    const int i0 = cta_off[blockIdx.x], i1 = cta_off[blockIdx.x + 1];
    if (i0 < i1) {
        const Item it = items[i0];
        for (int g = threadIdx.x; g < G; g += blockDim.x) { const long long gi = (long long)it.blk * G + g; mb[g] = gi < L ? m[gi] : 0.0; }
    }
    __syncthreads();
    for (int ii = i0; ii < i1; ++ii) {
        const Item it = items[ii];
        const double *mcur = mb + (size_t)((ii - i0) & 1) * G;
        double *mnext = mb + (size_t)((ii - i0 + 1) & 1) * G;
        double pre[16];
        const bool more = ii + 1 < i1;
        Item nx = it;
        if (more) {                                          // m block of the next item: loads now, stores after the work
            nx = items[ii + 1];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int g = threadIdx.x + r * WARPS * 32;
                const long long gi = (long long)nx.blk * G + g;
                pre[r] = (g < G && gi < L) ? m[gi] : 0.0;
            }
        }
        const long long *sg = seg + (long long)it.blk * NT;
        for (int t = it.t0 + w; t < it.t1; t += WARPS) {
            const long long s0 = sg[t], s1 = sg[t + 1];
            double a = 0.0;
            long long n = s0 + lane;
            for (; n + 32 * (UNR - 1) < s1; n += 32 * UNR) {
                double v[UNR]; unsigned short ix[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) { v[u] = pv[n + 32 * u]; ix[u] = pi[n + 32 * u]; }
#pragma unroll
                for (int u = 0; u < UNR; ++u) a = fma(v[u], mcur[ix[u]], a);
            }
            for (; n < s1; n += 32) a = fma(pv[n], mcur[pi[n]], a);
            acc[t * 32 + lane] += a;
        }
        if (more) {
#pragma unroll
            for (int r = 0; r < 16; ++r) { const int g = threadIdx.x + r * WARPS * 32; if (g < G) mnext[g] = pre[r]; }
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < NT; t += blockDim.x) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < 32; ++r) s += acc[t * 32 + r];
        part[(long long)blockIdx.x * NT + t] = s;
    }
}

__global__ void fold_kernel(const double *part, int nparts, int NT, double *out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NT) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(long long)p * NT + t];
    out[t] = s;
}

static void next_comb(std::vector<int> &c, int N)
{
    const int k = (int)c.size();
    int i = k - 1;
    while (i >= 0 && c[i] == N - k + i) --i;
    if (i < 0) return;
    ++c[i];
    for (int j = i + 1; j < k; ++j) c[j] = c[j - 1] + 1;
}

template <typename F>
static float timeit(F f, int reps)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps * 1e3f;
}

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 20;
    const int NT = N * (N + 1) / 2;
    std::vector<unsigned> gmask;
    for (int k = 1; k <= N; ++k) {
        std::vector<int> comb(k); for (int j = 0; j < k; ++j) comb[j] = j;
        long long Lk = 1; for (int j = 0; j < k; ++j) Lk = Lk * (N - j) / (j + 1);
        for (long long i = 0; i < Lk; ++i) { unsigned mk = 0; for (int v : comb) mk |= 1u << v; gmask.push_back(mk); next_comb(comb, N); }
    }
    const long long L = (long long)gmask.size();
    std::vector<double> hm((size_t)L);
    srand(1);
    for (auto &v : hm) v = 1.0 + 10.0 * rand() / (double)RAND_MAX;
    int nsm = 148; { cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); nsm = p.multiProcessorCount; }
    double *d_m; CK(cudaMalloc(&d_m, 8 * L)); CK(cudaMemcpy(d_m, hm.data(), 8 * L, cudaMemcpyHostToDevice));

    for (int G : {1024, 2048, 4096}) {
        const int nblk = (int)((L + G - 1) / G);
        std::vector<long long> seg((size_t)nblk * NT + 1, 0);
        for (long long i = 0; i < L; ++i) {
            const unsigned mk = gmask[i]; const int blk = (int)(i / G);
            for (int a = 0; a < N; ++a) if (mk >> a & 1u) for (int b = a; b < N; ++b) if (mk >> b & 1u)
                seg[(size_t)blk * NT + (a * N - a * (a - 1) / 2 + (b - a)) + 1]++;
        }
        for (size_t s = 1; s < seg.size(); ++s) seg[s] += seg[s - 1];
        const long long S = seg.back();
        std::vector<double> pv((size_t)S); std::vector<uint16_t> pi((size_t)S);
        std::vector<long long> cur(seg.begin(), seg.end() - 1);
        std::vector<double> ref(NT, 0.0);
        for (long long i = 0; i < L; ++i) {
            const unsigned mk = gmask[i]; const int blk = (int)(i / G);
            for (int a = 0; a < N; ++a) if (mk >> a & 1u) for (int b = a; b < N; ++b) if (mk >> b & 1u) {
                const int t = a * N - a * (a - 1) / 2 + (b - a);
                const long long p = cur[(size_t)blk * NT + t]++;
                const double val = ((i * 31 + t * 7) % 1000) / 1000.0 - 0.3;
                pv[(size_t)p] = val; pi[(size_t)p] = (uint16_t)(i - (long long)blk * G);
                ref[t] += val * hm[(size_t)i];
            }
        }
        double refmax = 0; for (double v : ref) refmax = std::max(refmax, fabs(v));
        double *d_pv, *d_part, *d_phi; uint16_t *d_pi; long long *d_seg; Item *d_items; int *d_off;
        CK(cudaMalloc(&d_pv, 8 * S)); CK(cudaMemcpy(d_pv, pv.data(), 8 * S, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_pi, 2 * S)); CK(cudaMemcpy(d_pi, pi.data(), 2 * S, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_seg, 8 * seg.size())); CK(cudaMemcpy(d_seg, seg.data(), 8 * seg.size(), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_part, 8 * (size_t)NT * 2048)); CK(cudaMalloc(&d_phi, 8 * NT));
        printf("N=%d G=%d nblk=%d entries=%lld stream=%.1f MB (values %.1f + idx %.1f)\n", N, G, nblk, S, 10.0 * S / 1e6, 8.0 * S / 1e6, 2.0 * S / 1e6);
        for (int nchunk : {1, 2, 4, 8}) {
            // items ordered (chunk, block); contiguous work-balanced ranges per CTA
            std::vector<Item> items; std::vector<long long> wsum;
            long long tot = 0;
            for (int ch = 0; ch < nchunk; ++ch)
                for (int b = 0; b < nblk; ++b) {
                    Item it; it.blk = b; it.t0 = (int)((long long)NT * ch / nchunk); it.t1 = (int)((long long)NT * (ch + 1) / nchunk); it.pad = 0;
                    items.push_back(it);
                    tot += seg[(size_t)b * NT + it.t1] - seg[(size_t)b * NT + it.t0];
                    wsum.push_back(tot);
                }
            CK(cudaMalloc(&d_items, sizeof(Item) * items.size())); CK(cudaMemcpy(d_items, items.data(), sizeof(Item) * items.size(), cudaMemcpyHostToDevice));
            for (int cps : {1, 2, 3}) {
                const size_t smem = 8ull * ((size_t)NT * 32 + 2 * G);
                if (smem * cps > 227 * 1024) continue;
                const int grid = nsm * cps;
                std::vector<int> off(grid + 1, 0);
                for (int c = 1; c <= grid; ++c) {
                    const long long target = tot * c / grid;
                    off[c] = (int)(std::lower_bound(wsum.begin(), wsum.end(), target) - wsum.begin()) + 1;
                    if (off[c] > (int)items.size()) off[c] = (int)items.size();
                    if (off[c] < off[c - 1]) off[c] = off[c - 1];
                }
                off[grid] = (int)items.size();
                CK(cudaMalloc(&d_off, 4 * (grid + 1))); CK(cudaMemcpy(d_off, off.data(), 4 * (grid + 1), cudaMemcpyHostToDevice));
                long long wmax = 0;
                for (int c = 0; c < grid; ++c) { const long long a = off[c] ? wsum[off[c] - 1] : 0, b2 = off[c + 1] ? wsum[off[c + 1] - 1] : 0; wmax = std::max(wmax, b2 - a); }
                auto run = [&](auto kern, const char *tag) {
                    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    auto go = [&] { kern<<<grid, 256, smem>>>(d_pv, d_pi, d_seg, d_items, d_off, NT, G, L, d_m, d_part); };
                    const float us = timeit(go, 20);
                    CK(cudaGetLastError());
                    fold_kernel<<<(NT + 127) / 128, 128>>>(d_part, grid, NT, d_phi);
                    std::vector<double> hphi(NT);
                    CK(cudaMemcpy(hphi.data(), d_phi, 8 * NT, cudaMemcpyDeviceToHost));
                    double err = 0; for (int t = 0; t < NT; ++t) err = std::max(err, fabs(hphi[t] - ref[t]));
                    printf("  G=%d chunks=%d ctas/sm=%d %s imbalance=%.3f err=%.1e : %7.1f us  %7.1f GB/s (10 B/entry)\n", G, nchunk, cps, tag,
                           (double)wmax * grid / tot, err / refmax, us, 10.0 * S / us / 1e3);
                    fflush(stdout);
                };
                run(phi_stream_kernel<4, 8>, "UNR=4");
                run(phi_stream_kernel<8, 8>, "UNR=8");
                CK(cudaFree(d_off));
            }
            CK(cudaFree(d_items));
        }
        CK(cudaFree(d_pv)); CK(cudaFree(d_pi)); CK(cudaFree(d_seg)); CK(cudaFree(d_part)); CK(cudaFree(d_phi));
    }
    return 0;
}
