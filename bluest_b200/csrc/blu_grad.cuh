// blu_grad.cuh -- kernel (3a): per-group quadratic forms with x = first row of pinv(Phi).
//
//   grad_i = - x[g_i]^T Cinv_i x[g_i]                         (gradK_c, cmisc.cpp:58-72; misc.py:493)
//   u_i    = R_i^T Cinv_i x[g_i]  (scattered to model slots)   (intended cleanup matrix, misc.py:507-516)
//   v_i    = S u_i,  S = pinv(Phi) + pinv(Phi)^T = 2 pinv(Phi)
// so that the Hessian of misc.py:497-503 is  H = U S U^T = U V^T  (SURVEY.md 0.3), which
// blu_hess.cuh forms with FP64 tensor-core MMAs.
//
// Both kernels stream chunks of consecutive groups through the warp-private bulk-async ring of
// blu_stream.cuh.
// blu_grad_kernel   (no Hessian wanted): lane-per-packed-entry, one shuffle-tree sum per group -- a
//                   pure HBM stream of the packed inverses.
// blu_gradu_kernel  (Hessian / U wanted): lane j forms row j of Cinv_i x[g_i] from the staged block,
//                   rows are scattered to model
//                   positions via the membership mask and multiplied by S; U and V rows (NP doubles,
//                   one 128-byte line at N <= 16) are written coalesced.
// x and S are broadcast once per CTA into registers / shared memory.
#pragma once
#include "blu_common.cuh"
#include "blu_stream.cuh"

// Consume one staged chunk for the gradient: lane g keeps -x_g^T Cinv_g x_g of group g.
template <int S>
__device__ __forceinline__ double blu_grad_chunk(const double *__restrict__ base, const unsigned short *__restrict__ lt, int T, int ng,
                                                 const unsigned char *__restrict__ ids, double xv, int lane)
{
    int ja[S], la[S];
    double wg[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int e = s * 32 + lane;
        const unsigned jl = e < T ? lt[e] : 0u;
        ja[s] = jl >> 8; la[s] = jl & 255u;
        wg[s] = e < T ? ((ja[s] == la[s]) ? 1.0 : 2.0) : 0.0;
    }
    double mine = 0.0;
    for (int g = 0; g < ng; ++g) {
        const int gv = ids[g * BLU_IDS_LD + lane];
        const double xg = blu_shfl(xv, gv);
        const double *cp = base + g * T;
        double sum = 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int e = s * 32 + lane;
            const double v = e < T ? cp[e] : 0.0;
            const double xa = blu_shfl(xg, ja[s]), xb = blu_shfl(xg, la[s]);
            sum += wg[s] * (xa * v * xb);
        }
        sum = blu_warp_sum(sum);
        if (lane == g) mine = -sum;
    }
    return mine;
}
__device__ __forceinline__ double blu_grad_chunk_any(const double *__restrict__ base, const unsigned short *__restrict__ lt, int T, int ng,
                                                     const unsigned char *__restrict__ ids, double xv, int lane)
{
    double mine = 0.0;
    for (int g = 0; g < ng; ++g) {
        const int gv = ids[g * BLU_IDS_LD + lane];
        const double xg = blu_shfl(xv, gv);
        const double *cp = base + g * T;
        double sum = 0.0;
        for (int e0 = 0; e0 < T; e0 += 32) {
            const int e = e0 + lane;
            const bool ok = e < T;
            const unsigned jl = ok ? lt[e] : 0u;
            const double v = ok ? cp[e] : 0.0;
            const int j = jl >> 8, l = jl & 255u;
            const double xa = blu_shfl(xg, j), xb = blu_shfl(xg, l);
            sum += ((j == l) ? 1.0 : 2.0) * (xa * v * xb);
        }
        sum = blu_warp_sum(sum);
        if (lane == g) mine = -sum;
    }
    return mine;
}

// extra shared doubles: none
__global__ void __launch_bounds__(BLU_STREAM_WARPS * 32)
blu_grad_kernel(const BluClass *__restrict__ cls, int ncls, int N, const BluChunk *__restrict__ chunks, int nchunks, int sd,
                const double *__restrict__ cinv, const unsigned short *__restrict__ lut, int lutlen,
                const unsigned *__restrict__ gmask, const double *__restrict__ xrow, double *__restrict__ grad)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const BluStreamSmem sm = blu_stream_carve(smraw, sd, 0, ncls, lutlen);
    const BluWarpStream ws = blu_stream_begin(sm, sd, cls, ncls, lut, lutlen);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ids = sm.ids + w * BLU_IDS_BYTES;
    const double xv = lane < N ? xrow[lane] : 0.0;
    const int gw = blockIdx.x * BLU_STREAM_WARPS + w;
    const int nw = gridDim.x * BLU_STREAM_WARPS;

    int c = gw;
    BluChunkRegs cur, nxt;
    BluChunk dnext;
    if (c < nchunks) cur = blu_prefetch_chunk(chunks[c], sm.cls, cinv, nullptr, gmask, ws, 0, lane);
    if (c + nw < nchunks) dnext = chunks[c + nw];
    for (int it = 0; c < nchunks; c += nw, ++it) {
        const int s = it & 1;
        if (c + nw < nchunks) nxt = blu_prefetch_chunk(dnext, sm.cls, cinv, nullptr, gmask, ws, s ^ 1, lane);
        if (c + 2 * nw < nchunks) dnext = chunks[c + 2 * nw];
        const BluClass ci = sm.cls[cur.cls];
        const int T = ci.T;
        const unsigned short *lt = sm.lut + ci.lutoff;
        blu_mbar_wait(ws.bar[s], (unsigned)((it >> 1) & 1));
        const double *base = ws.stage[s] + cur.skew;
        blu_expand_ids(cur.mask, ci.k, ids, lane);
        double mine;                                      // lane g keeps the result of group g
        switch ((T + 31) >> 5) {
            case 1: mine = blu_grad_chunk<1>(base, lt, T, cur.g, ids, xv, lane); break;
            case 2: mine = blu_grad_chunk<2>(base, lt, T, cur.g, ids, xv, lane); break;
            case 3: mine = blu_grad_chunk<3>(base, lt, T, cur.g, ids, xv, lane); break;
            case 4: mine = blu_grad_chunk<4>(base, lt, T, cur.g, ids, xv, lane); break;
            case 5: mine = blu_grad_chunk<5>(base, lt, T, cur.g, ids, xv, lane); break;
            case 6: mine = blu_grad_chunk<6>(base, lt, T, cur.g, ids, xv, lane); break;
            case 7: mine = blu_grad_chunk<7>(base, lt, T, cur.g, ids, xv, lane); break;
            case 8: mine = blu_grad_chunk<8>(base, lt, T, cur.g, ids, xv, lane); break;
            default: mine = blu_grad_chunk_any(base, lt, T, cur.g, ids, xv, lane); break;
        }
        if (lane < cur.g) grad[ci.goff + cur.i0 + lane] = mine;          // coalesced
        __syncwarp();
        cur = nxt;
    }
}

// extra shared doubles: N*N (S)
__global__ void __launch_bounds__(BLU_STREAM_WARPS * 32)
blu_gradu_kernel(const BluClass *__restrict__ cls, int ncls, int N, int NP, const BluChunk *__restrict__ chunks, int nchunks, int sd,
                 const double *__restrict__ cinv, const unsigned short *__restrict__ lut, int lutlen,
                 const unsigned *__restrict__ gmask, const double *__restrict__ xrow, const double *__restrict__ S,
                 double *__restrict__ grad, double *__restrict__ U, double *__restrict__ V)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const BluStreamSmem sm = blu_stream_carve(smraw, sd, N * N, ncls, lutlen);
    double *sS = sm.extra;
    for (int t = threadIdx.x; t < N * N; t += blockDim.x) sS[t] = S[t];
    const BluWarpStream ws = blu_stream_begin(sm, sd, cls, ncls, lut, lutlen);       // ends with __syncthreads
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ids = sm.ids + w * BLU_IDS_BYTES;
    const double xv = lane < N ? xrow[lane] : 0.0;
    const int gw = blockIdx.x * BLU_STREAM_WARPS + w;
    const int nw = gridDim.x * BLU_STREAM_WARPS;

    int c = gw;
    BluChunkRegs cur, nxt;
    BluChunk dnext;
    if (c < nchunks) cur = blu_prefetch_chunk(chunks[c], sm.cls, cinv, nullptr, gmask, ws, 0, lane);
    if (c + nw < nchunks) dnext = chunks[c + nw];
    for (int it = 0; c < nchunks; c += nw, ++it) {
        const int s = it & 1;
        if (c + nw < nchunks) nxt = blu_prefetch_chunk(dnext, sm.cls, cinv, nullptr, gmask, ws, s ^ 1, lane);
        if (c + 2 * nw < nchunks) dnext = chunks[c + 2 * nw];
        const BluClass ci = sm.cls[cur.cls];
        const int k = ci.k, T = ci.T;
        blu_mbar_wait(ws.bar[s], (unsigned)((it >> 1) & 1));
        const double *base = ws.stage[s] + cur.skew;
        blu_expand_ids(cur.mask, k, ids, lane);
        double mine = 0.0;
        for (int g = 0; g < cur.g; ++g) {
            const unsigned mask = __shfl_sync(BLU_FULL, cur.mask, g);
            const int gv = ids[g * BLU_IDS_LD + lane];
            const double xg = blu_shfl(xv, gv);
            const double *st = base + g * T;
            // lane j < k: row j of Cinv_i x[g_i]
            const int j = lane < k ? lane : k - 1;
            double y = 0.0;
            for (int l = 0; l < k; ++l) {
                const double xl = blu_shfl(xg, l);
                const int a = j < l ? j : l, b = j < l ? l : j;
                y += st[blu_pk(k, a, b)] * xl;
            }
            const double gs = blu_warp_sum(lane < k ? xg * y : 0.0);
            if (lane == g) mine = -gs;
            // scatter rows to model slots, then v = S u
            const bool in = (mask >> lane) & 1u;
            const int pos = __popc(mask & ((1u << lane) - 1u));
            const double ysrc = blu_shfl(y, pos);
            const double ua = in ? ysrc : 0.0;
            double va = 0.0;
            for (int b = 0; b < N; ++b) {
                const double ub = blu_shfl(ua, b);
                if (lane < N) va += sS[lane * N + b] * ub;
            }
            const long long row = ci.goff + cur.i0 + g;
            if (lane < NP) {
                U[row * NP + lane] = ua;
                V[row * NP + lane] = lane < N ? va : 0.0;
            }
        }
        if (lane < cur.g) grad[ci.goff + cur.i0 + lane] = mine;
        __syncwarp();
        cur = nxt;
    }
}

// get_cleanup_matrix (misc.py:507-516).  X is (N, L) row-major on the device.
// mode 0: the reference's assignment semantics, X[g[j], i] = Cinv_i[j,k-1] * x[g[k-1]]
//         (cmisc.cpp:51: "=" inside the l loop, the last l wins).
// mode 1: the intended matrix, X[:, i] = u_i.
__global__ void blu_cleanup_kernel(const BluClass *__restrict__ cls, int ncls, int N, long long L,
                                   const uint8_t *__restrict__ gidx, const double *__restrict__ cinv,
                                   const double *__restrict__ xrow, int mode, double *__restrict__ X)
{
    for (int c = 0; c < ncls; ++c) {
        const BluClass ci = cls[c];
        const int k = ci.k, T = ci.T;
        const long long total = ci.Lk * k;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
             t += (long long)gridDim.x * blockDim.x) {
            const long long i = t / k;
            const int j = (int)(t - i * k);
            const uint8_t *g = gidx + ci.ioff + i * k;
            const double *cp = cinv + ci.coff + i * T;
            double val;
            if (mode == 0) {
                val = cp[blu_pk(k, j, k - 1)] * xrow[g[k - 1]];
            } else {
                val = 0.0;
                for (int l = 0; l < k; ++l) {
                    const int a = j < l ? j : l, b = j < l ? l : j;
                    val += cp[blu_pk(k, a, b)] * xrow[g[l]];
                }
            }
            X[(long long)g[j] * L + ci.goff + i] = val;
        }
    }
}

// Dense psi (N*N, L): psi[N*g[j]+g[l], i] = Cinv_i[j,l]   (assemble_psi_c, cmisc.cpp:10-23; sap.py:129).
// psi must be zero-filled; columns [col0, col0+ncols) of the flat enumeration are produced into a
// (N*N, ncols) row-major panel.
__global__ void blu_psi_kernel(const BluClass *__restrict__ cls, int ncls, int N, long long col0,
                               long long ncols, const uint8_t *__restrict__ gidx,
                               const double *__restrict__ cinv, double *__restrict__ psi)
{
    for (int c = 0; c < ncls; ++c) {
        const BluClass ci = cls[c];
        const int k = ci.k, T = ci.T;
        long long i0 = col0 > ci.goff ? col0 - ci.goff : 0;
        long long i1 = col0 + ncols < ci.goff + ci.Lk ? col0 + ncols - ci.goff : ci.Lk;
        if (i1 <= i0) continue;
        const long long total = (i1 - i0) * k * k;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
             t += (long long)gridDim.x * blockDim.x) {
            const long long i = i0 + t / (k * k);
            const int rem = (int)(t % (k * k));
            const int j = rem / k, l = rem - j * k;
            const uint8_t *g = gidx + ci.ioff + i * k;
            const int a = j < l ? j : l, b = j < l ? l : j;
            psi[((long long)N * g[j] + g[l]) * ncols + (ci.goff + i - col0)] = cinv[ci.coff + i * T + blu_pk(k, a, b)];
        }
    }
}

// BLUE estimator right-hand side (compute_BLUE_estimator, sap.py:99-119):
//     y[g_i[j]] += sum_s Cinv_i[j,s] * sums_i[s]      over all groups,
// the same packed-inverse stream as the U factor with the per-group sample sums as the right-hand
// side.  `sums` is flat in group order, k entries per group (same indexing as the id table).
// Per-warp private y tiles (targets inside a group are distinct), fixed-order reductions.
// part: (gridDim.x, 32).
__global__ void __launch_bounds__(BLU_STREAM_WARPS * 32)
blu_ysum_kernel(const BluClass *__restrict__ cls, int ncls, int N, const BluChunk *__restrict__ chunks, int nchunks, int sd,
                const double *__restrict__ cinv, const unsigned short *__restrict__ lut, int lutlen,
                const unsigned *__restrict__ gmask, const double *__restrict__ sums, double *__restrict__ part)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const BluStreamSmem sm = blu_stream_carve(smraw, sd, BLU_STREAM_WARPS * 32, ncls, lutlen);
    const BluWarpStream ws = blu_stream_begin(sm, sd, cls, ncls, lut, lutlen);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ids = sm.ids + w * BLU_IDS_BYTES;
    double *yacc = sm.extra + w * 32;
    yacc[lane] = 0.0;
    __syncwarp();
    const int gw = blockIdx.x * BLU_STREAM_WARPS + w;
    const int nw = gridDim.x * BLU_STREAM_WARPS;
    int c = gw;
    BluChunkRegs cur, nxt;
    BluChunk dnext;
    if (c < nchunks) cur = blu_prefetch_chunk(chunks[c], sm.cls, cinv, nullptr, gmask, ws, 0, lane);
    if (c + nw < nchunks) dnext = chunks[c + nw];
    for (int it = 0; c < nchunks; c += nw, ++it) {
        const int s = it & 1;
        if (c + nw < nchunks) nxt = blu_prefetch_chunk(dnext, sm.cls, cinv, nullptr, gmask, ws, s ^ 1, lane);
        if (c + 2 * nw < nchunks) dnext = chunks[c + 2 * nw];
        const BluClass ci = sm.cls[cur.cls];
        const int k = ci.k, T = ci.T;
        blu_mbar_wait(ws.bar[s], (unsigned)((it >> 1) & 1));
        const double *base = ws.stage[s] + cur.skew;
        blu_expand_ids(cur.mask, k, ids, lane);
        for (int g = 0; g < cur.g; ++g) {
            const int gv = ids[g * BLU_IDS_LD + lane];
            const double sv = lane < k ? sums[ci.ioff + (cur.i0 + g) * k + lane] : 0.0;
            const double *st = base + g * T;
            const int j = lane < k ? lane : k - 1;
            double y = 0.0;
            for (int l = 0; l < k; ++l) {
                const double sl = blu_shfl(sv, l);
                const int a = j < l ? j : l, b = j < l ? l : j;
                y += st[blu_pk(k, a, b)] * sl;
            }
            if (lane < k) yacc[gv] += y;
            __syncwarp();
        }
        __syncwarp();
        cur = nxt;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double sum = 0.0;
#pragma unroll
        for (int ww = 0; ww < BLU_STREAM_WARPS; ++ww) sum += sm.extra[ww * 32 + threadIdx.x];
        part[(long long)blockIdx.x * 32 + threadIdx.x] = sum;
    }
}

// y = fixed-order sum of the CTA partials; mu = xsup . y  (PHIinvY0, misc.py:529-533).  One warp.
__global__ void blu_ysum_finish_kernel(const double *__restrict__ part, int nparts, int N, double *__restrict__ y, BluEvalHeader *hdr)
{
    const int lane = threadIdx.x;
    double sum = 0.0;
    for (int p = 0; p < nparts; ++p) sum += part[(long long)p * 32 + lane];
    if (lane < N) y[lane] = sum;
    double prod = lane < N ? hdr->xsup[lane] * sum : 0.0;
    // sequential, model order (the reference sums j = 0..len(y)-1 in order)
    double mu = 0.0;
    for (int a = 0; a < N; ++a) mu += blu_shfl(prod, a);
    if (lane == 0) hdr->scal[5] = mu;
}
