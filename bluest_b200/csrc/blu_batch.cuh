// blu_batch.cuh -- batched evaluation of SMALL problems: P problems x B sample vectors in ONE launch.
//
// Replaces the Python loops over outputs (mosap.py:86-100: `for n in range(n_outputs): SAPS[n].variance_GH(m[mappings[n]])`)
// and over the instances of a budget / tolerance sweep, where every evaluation is a few microseconds of work behind
// several kernel launches.  One CTA owns one (problem, sample vector) pair from start to end -- Phi assembly over all
// groups of the problem (misc.py:459-461), N x N pseudo-inverse and variance (misc.py:463-477, 487-490), gradient
// (cmisc.cpp:58-72) -- so nothing is exchanged between CTAs and the whole batch is a single wave of the machine.
// The packed inverses of a problem with ~1000 groups are ~100 KB: L2 resident after the first touch.
#pragma once
#include "blu_common.cuh"
#include "blu_phi.cuh"

#define BLU_BATCH_WARPS 32          // 1024 threads: the per-warp chain of dependent shared-memory accesses is the critical path

struct BluBatchProb {              // one problem of the batch: pointers into its context's HBM data
    const BluClass *cls;
    const uint8_t *gidx;
    const double *cinv;
    const unsigned *gmask;
    const unsigned short *lut;     // (j,l) tables of all classes (BluClass::lutoff)
    const long long *map;          // gather map into the shared sample vector (mosap.py:54-67 `mappings[n]`), or NULL
    long long L;
    long long moff;                // map == NULL: offset of this problem's block inside one input vector
    long long goff;                // groups of the problems before this one: its (B, L) gradient block starts at goff * B
    long long cinv_len, gidx_len;  // doubles of packed inverses / bytes of member ids
    int ncls, N, lutlen, pad;
};

#define BLU_BATCH_CHUNK 4096          // streamed mode: doubles of packed inverses staged per chunk (32 KB)
#define BLU_BATCH_MAXG 1024           // groups per chunk at most

// Walk all groups of a problem: body(ci, il, g (k member ids), C (T packed entries), jl ((j,l) table), lane), warp w
// takes every WARPS-th group.
//  resident: the problem's packed inverses, member ids and (j,l) tables already sit in shared memory (sC, sG, sjl
//            hold verbatim copies) -- no barrier, no global load in the loop; both passes of an evaluation (Phi,
//            gradient) read the one copy.
//  streamed: chunks of consecutive groups of one class are staged by the whole CTA with coalesced loads (every
//            thread has independent loads in flight: a warp walking its groups one dependent L2 access at a time would
//            be pure latency), two barriers per chunk.
template <typename Body>
__device__ __forceinline__ void blu_batch_walk(const BluBatchProb &pr, const BluClass *scls, bool resident, double *sC, uint8_t *sG, unsigned short *sjl, Body body)
{
    const int tid = threadIdx.x, nthr = blockDim.x, w = tid >> 5, lane = tid & 31;
    for (int ic = 0; ic < pr.ncls; ++ic) {
        const BluClass ci = scls[ic];                     // the shared-memory copy: pr.cls is an L2 round trip per class
        const int k = ci.k, T = ci.T;
        const unsigned short *jl = sjl + ci.lutoff;
        if (resident) {
            for (long long il = w; il < ci.Lk; il += BLU_BATCH_WARPS) body(ci, il, sG + ci.ioff + il * k, sC + ci.coff + il * T, jl, lane);
            continue;
        }
        int G = BLU_BATCH_CHUNK / T;
        if (G > BLU_BATCH_MAXG) G = BLU_BATCH_MAXG;
        for (long long i0 = 0; i0 < ci.Lk; i0 += G) {
            const int ng = (int)((ci.Lk - i0) < G ? (ci.Lk - i0) : G);
            __syncthreads();                              // previous chunk consumed
            const double *src = pr.cinv + ci.coff + i0 * T;
            for (int t = tid; t < ng * T; t += nthr) sC[t] = src[t];
            const uint8_t *gsrc = pr.gidx + ci.ioff + i0 * k;
            for (int t = tid; t < ng * k; t += nthr) sG[t] = gsrc[t];
            __syncthreads();
            for (int gl = w; gl < ng; gl += BLU_BATCH_WARPS) body(ci, i0 + gl, sG + gl * k, sC + gl * T, jl, lane);
        }
    }
}

// class of flat group i (classes are few: linear scan of the shared-memory table)
__device__ __forceinline__ int blu_batch_class_of(const BluClass *scls, int ncls, long long i)
{
    int ic = 0;
    while (ic + 1 < ncls && i >= scls[ic + 1].goff) ++ic;
    return ic;
}

// grid (P, B), BLU_BATCH_WARPS warps.  Dynamic shared memory:
// [WARPS x N*N warp tiles | finish scratch | x (32) | m of the problem (Lmax) | inverses (capC doubles) | ids (capG bytes) | (j,l) tables]
__global__ void __launch_bounds__(BLU_BATCH_WARPS * 32)
blu_batch_eval_kernel(const BluBatchProb *__restrict__ probs, const double *__restrict__ m_all, long long mstride, double delta,
                      int want_grad, long long Lmax, long long capC, long long capG, int resident,
                      BluEvalHeader *__restrict__ hdrs, double *__restrict__ scratch, double *__restrict__ var_out,
                      unsigned *__restrict__ flags_out, double *__restrict__ grad_out)
{
    extern __shared__ __align__(16) unsigned char braw[];
    const BluBatchProb pr = probs[blockIdx.x];
    const int b = blockIdx.y, B = gridDim.y;
    const int N = pr.N, NN = N * N;
    const int tid = threadIdx.x, nthr = blockDim.x, w = tid >> 5, lane = tid & 31;
    double *tiles = reinterpret_cast<double *>(braw);
    unsigned char *fraw = braw + sizeof(double) * (size_t)BLU_BATCH_WARPS * NN;
    double *sx = reinterpret_cast<double *>(fraw + BLU_FIN_SCRATCH_BYTES);
    double *sm = sx + 32;                                 // the problem's sample vector
    double *sC = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(sm + Lmax) + 15) & ~static_cast<uintptr_t>(15));   // bulk copies need 16-byte aligned destinations
    uint8_t *sG = reinterpret_cast<uint8_t *>(sC + capC);
    unsigned short *sjl = reinterpret_cast<unsigned short *>(sG + capG);
    __shared__ unsigned s_supp;
    __shared__ unsigned long long s_max;
    __shared__ BluClass scls[BLU_MAX_MODELS_C];
    __shared__ __align__(8) unsigned long long s_bar;
    const double *mvec = m_all + (long long)b * mstride + (pr.map ? 0 : pr.moff);
    BluEvalHeader *hdr = hdrs + (size_t)blockIdx.x * B + b;
    double *acc = tiles + (size_t)w * NN;
    BLU_STAMP(hdr, 0);                                   // [0] start
    for (int t = lane; t < NN; t += 32) acc[t] = 0.0;
    if (tid == 0) { s_supp = 0u; s_max = 0ull; blu_mbar_init(&s_bar, 1); blu_mbar_fence_init(); }
    for (int t = tid; t < pr.ncls; t += nthr) scls[t] = pr.cls[t];
    __syncthreads();
    // ---- one burst of independent loads: (j,l) tables, (resident) inverses and ids, the sample vector (gathered
    //      through the map) with the support mask and max|m| on the way ----
    for (int t = tid; t < pr.lutlen; t += nthr) sjl[t] = pr.lut[t];
    if (resident) {
        // the problem's packed inverses: bulk asynchronous copies (1-D TMA) issued by one thread, completion on an
        // mbarrier -- a per-thread load loop took 5.8 us for 113 KB (latency bound), the copy engine streams it
        if (tid == 0) {
            const unsigned long long total = (unsigned long long)pr.cinv_len * 8ull;      // class blocks are 128-byte aligned: multiple of 16
            blu_mbar_expect_tx(&s_bar, (unsigned)total);
            for (unsigned long long off = 0; off < total; off += 32768ull) {
                const unsigned bytes = (unsigned)((total - off) < 32768ull ? (total - off) : 32768ull);
                blu_bulk_g2s(reinterpret_cast<unsigned char *>(sC) + off, reinterpret_cast<const unsigned char *>(pr.cinv) + off, bytes, &s_bar);
            }
        }
        for (long long t = tid; t < pr.gidx_len; t += nthr) sG[t] = pr.gidx[t];
    }
    {
        unsigned supp = 0u;
        double mymax = 0.0;
        for (long long i = tid; i < pr.L; i += nthr) {
            const double mi = pr.map ? mvec[pr.map[i]] : mvec[i];
            sm[i] = mi;
            const double am = fabs(mi);
            mymax = fmax(mymax, am);
            if (am > 1.0e-6) supp |= pr.gmask[i];
        }
        supp = __reduce_or_sync(BLU_FULL, supp);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mymax = fmax(mymax, __shfl_xor_sync(BLU_FULL, mymax, o));
        __syncthreads();
        if (lane == 0) {
            if (supp) atomicOr(&s_supp, supp);
            atomicMax(&s_max, (unsigned long long)__double_as_longlong(mymax));
        }
    }
    if (resident) blu_mbar_wait(&s_bar, 0u);
    BLU_STAMP(hdr, 1);                                   // [1] operands staged
    // ---- Phi: warp per group, lane per packed entry, warp-private tile (targets inside a group are distinct) ----
    // Measured alternatives of this phase at 10 models (12.2 us of the 26 us kernel; tools/batch_stamps.py): operands of the next step
    // fetched before the current read-modify-write: 15.3 us; gather form (a warp per target entry scanning the group masks, sums in
    // registers): 14.6 us -- divergence makes every scan step pay the full path.  The phase is bound by instruction issue on ONE SM
    // (36 k packed entries, half-empty steps for k <= 7); the next step would be a cluster of CTAs per evaluation.
    blu_batch_walk(pr, scls, resident != 0, sC, sG, sjl, [&](const BluClass &ci, long long il, const uint8_t *g, const double *C, const unsigned short *jlt, int ln) {
        const double mi = sm[ci.goff + il];
        if (mi != 0.0) {
            for (int e = ln; e < ci.T; e += 32) {
                const unsigned jl = jlt[e];
                const int a = g[jl >> 8], bb = g[jl & 255u];
                acc[a * N + bb] = fma(mi, C[e], acc[a * N + bb]);
            }
        }
        __syncwarp();
    });
    __syncthreads();
    const BluFinScratch f = blu_fin_carve(fraw);
    for (int e = tid; e < NN; e += nthr) {
        double sum = 0.0;
        for (int ww = 0; ww < BLU_BATCH_WARPS; ++ww) sum += tiles[(size_t)ww * NN + e];      // warp order
        f.Ph[(e / N) * BLU_JLD + (e % N)] = sum;
    }
    if (tid == 0) { hdr->supp = s_supp; hdr->maxbits = s_max; *f.ns = 0; }
    __threadfence_block();
    __syncthreads();
    double *sc = scratch + ((size_t)blockIdx.x * B + b) * (size_t)(3 * NN + 40 + 32);
    double *phi = sc, *pinv = sc + NN + 40, *S = pinv + NN, *xrow = S + NN;
    BluPeers nopeers{};
    BLU_STAMP(hdr, 4);                                   // [4] Phi summed
    blu_finish_body(N, delta, 1, false, phi, pinv, xrow, S, hdr, nopeers, f, tid, nthr);
    __threadfence_block();
    __syncthreads();
    const unsigned fl = hdr->flags;
    if (tid == 0) { var_out[(size_t)blockIdx.x * B + b] = hdr->scal[0]; flags_out[(size_t)blockIdx.x * B + b] = fl; }
    if (!want_grad || (fl & BLU_FLAG_TINY)) return;       // the host fills the gradient of a tiny m with inf (misc.py:484)
    if (tid < 32) sx[tid] = tid < N ? xrow[tid] : 0.0;
    BLU_STAMP(hdr, 8);                                   // [8] finish done
    __syncthreads();
    // ---- gradient: grad_i = - x_g^T Cinv_i x_g ----
    double *gout = grad_out + pr.goff * B + (long long)b * pr.L;
    if (resident) {
        // one group per THREAD over the FLAT group index (no shuffles, independent chains): the data sits in shared memory.
        // (Class by class, thread 0 owned a group of EVERY class -- 220 dependent entry steps at 10 models, 10.6 us -- while
        // three quarters of the threads had none; flat, the longest chain is one group of the largest class: 1.2 us.)
        for (long long i = tid; i < pr.L; i += nthr) {
            const int ic = blu_batch_class_of(scls, pr.ncls, i);
            const int k = scls[ic].k, T = scls[ic].T;
            const long long il = i - scls[ic].goff;
            const uint8_t *g = sG + scls[ic].ioff + il * k;
            const double *C = sC + scls[ic].coff + il * T;
            double s = 0.0;
            int e = 0;
            for (int j = 0; j < k; ++j) {
                const double xj = sx[g[j]];
                double row = 0.5 * C[e] * xj;                 // diagonal entry counts once
                ++e;
                for (int l = j + 1; l < k; ++l, ++e) row = fma(C[e], sx[g[l]], row);
                s = fma(2.0 * xj, row, s);
            }
            gout[i] = -s;
        }
        BLU_STAMP(hdr, 9);                               // [9] gradient written (this thread's share)
        return;
    }
    blu_batch_walk(pr, scls, false, sC, sG, sjl, [&](const BluClass &ci, long long il, const uint8_t *g, const double *C, const unsigned short *jlt, int ln) {
        double s = 0.0;
        for (int e = ln; e < ci.T; e += 32) {
            const unsigned jl = jlt[e];
            const int j = jl >> 8, l = jl & 255u;
            const double xx = sx[g[j]] * sx[g[l]];
            s = fma(j == l ? C[e] : 2.0 * C[e], xx, s);
        }
        s = blu_warp_sum(s);
        if (ln == 0) gout[ci.goff + il] = -s;
    });
}
