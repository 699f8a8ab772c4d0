// blu_kkt.cuh -- scope-table row f1: a structure-exploiting KKT solve for the semidefinite programme that
// SAP.cvxopt_solve hands to cvxopt.solvers.sdp (sap.py:242-307; cvxopt calls its built-in dense solver there).
//
// Variables x in R^n (n = L + 1 with the epigraph variable t in front for a budget, sap.py:259-275; n = L for a
// tolerance, :276-287).  Inequalities: G0 x <=_l h0 with G0 = [-I_n ; Gx] (Gx: the nlin dense rows -- cost, coverage,
// sample caps) and one (N+1) x (N+1) semidefinite block G1 x <=_s h1 whose column of group i is
// -scales * vec(pad(Psi_i)) and whose t column (budget) is -E_NN.  With the Nesterov-Todd scaling
// W = blkdiag(diag(d), W_s), W_s vec(U) = vec(r^T U r), every interior-point iteration solves
//     [ 0   G^T   ] [ux]   [bx]
//     [ G  -W^T W ] [uz] = [bz].
// Eliminating uz:  M ux = bx + G^T (W^T W)^-1 bz,  M = diag(d_I^-2) + Gx^T diag(d_lin^-2) Gx + A^T A  with
// A = (r^-1 (x) r^-1) G1: a DIAGONAL plus a matrix of rank Q = (N+1)(N+2)/2 + nlin.  Woodbury:
//     ux = D^-1 rhs - D^-1 B (I + B^T D^-1 B)^-1 B^T D^-1 rhs,     B = [A^T | Gx^T diag(1/d_lin)],  D = diag(d_I^-2)
// so an iteration costs one weighted Gram contraction over the groups (Q^2 L multiply-adds on the FP64 tensor cores)
// and a Q x Q Cholesky instead of the (L+1)^3/3 of the dense factorisation.
//
// Device pipeline (all FP64):
//   blu_kkt_rows_kernel   one warp per column: a_col = svec(-scales r^-1 pad(Psi_i) r^-T) from the packed inverse and
//                         the member columns of r^-1, the nlin constraint entries, everything times d_col (= D^-1/2),
//                         plus the column's entry of G1^T vec(Lam Z Lam) (right-hand side) -> row of Bs (n x QP)
//   blu_kkt_syrk_kernel   cap = I + Bs^T Bs and v = Bs^T (d * rhs) in one pass (the rhs rides along as column Q):
//                         mma.m8n8k4.f64, a CTA per row range with the rows staged once in shared memory, all tile pairs
//                         dealt to its warps; CTA partials folded in a fixed order
//   blu_kkt_chol_kernel   blocked Cholesky of the Q x Q capacitance matrix (in place in L2, 16-column panels in shared
//                         memory, the right-hand side carried as an extra row) + backward substitution, one CTA
//   blu_kkt_apply_kernel  ux_col = d_col^2 rhs_col - d_col (Bs_col . y)
// then Phi(ux) through the ordinary Phi kernel gives G1 ux, and uz follows on the host (O(n + N^2) work).
#pragma once
#include "blu_common.cuh"
#include "blu_hess.cuh"
#include "blu_stream.cuh"

#define BLU_KKT_WARPS 8

// svec index of (a,b), a <= b, of an M x M symmetric matrix stored by rows of the upper triangle
__host__ __device__ __forceinline__ int blu_svec(int M, int a, int b) { return a * M - a * (a - 1) / 2 + (b - a); }

// Rows of the reduced system, LW lanes per column (LW = 16 while M = N + 1 <= 16: a warp works on TWO groups of the same size
// class; LW = 32 otherwise, M <= 22).
//   col < has_t            : the t column, X = -E_NN
//   otherwise group i      : X = -scales * pad(Psi_i)
// Bs[col][q] = d[col] * b_col[q],  q < Qs: svec(r^-1 X r^-T) (off-diagonal entries times sqrt 2 so that the
// Euclidean inner product of svecs equals the trace inner product), Qs <= q < Q: Gx[q-Qs][col] / dlin[q-Qs],
// Bs[col][Q] = d[col] * rhs_col (filled by blu_kkt_rhs_kernel), zero padding up to QP.
// g1tw[col] = <X_col, Wm>  (Wm = Lam Z Lam, (N+1) x (N+1)): the column's entry of G1^T vec(Wm).
// A warp walks its pairs of groups with the NEXT pair's packed inverse, member ids and scaling already in flight (registers),
// so the DRAM round trip hides behind the products of the current pair.  Per group: the packed inverse is expanded to a
// full k x k matrix and the member columns of r^-1 are gathered into shared memory; then lane a owns row a and keeps its
// results in registers:
//   Tt[a][:] = Rg[a][:] C,   Y[a][:] = Tt[a][:] Rg^T,   Rg[a][j] = r^-1[a][g_j]
// one lane-contiguous shared read per outer step and broadcast 16-byte reads of the other operand (1.6 instructions per
// multiply-add); the finished row leaves through a staging row so that the global stores are contiguous.
// History (15 models): first version, global loads and an (a, b) search per entry inside the loops: 152 us; shared-memory
// staging without prefetch and with scalar product loops: 217 us (one exposed DRAM round trip per 16 packed entries).
template <int LW>
__global__ void __launch_bounds__(BLU_KKT_WARPS * 32)
blu_kkt_rows_kernel(const BluClass *__restrict__ cls, int ncls, int N, long long L, int has_t, double scales,
                    const uint8_t *__restrict__ gidx, const double *__restrict__ cinv, const unsigned short *__restrict__ lut, int lutlen,
                    const double *__restrict__ rinv, const double *__restrict__ Wm, const double *__restrict__ d, int nlin,
                    const double *__restrict__ Gx, const double *__restrict__ dlin, int Q, int QP, int KMAX, int SCR,
                    double *__restrict__ Bs, double *__restrict__ g1tw)
{
    constexpr int GPW = 32 / LW;
    constexpr int NH = (LW == 16) ? 8 : 11;               // register pairs per lane: M <= 2 NH, k <= 2 NH
    constexpr int KSF = 2 * NH;                           // row pitch of the expanded inverse
    constexpr int CV = 8;                                 // packed entries per lane: T <= CV * LW
    const int M = N + 1, Qs = M * (M + 1) / 2;
    const int MP = M | 1;                                 // odd leading dimension: rows of r^-1 land in different banks
    extern __shared__ __align__(16) double kraw[];        // [M*MP r^-1][M*MP Wm][sub-warps x SCR scratch][(j,l) tables]
    double *sR = kraw, *sW = kraw + M * MP;
    __shared__ BluClass scls[BLU_MAX_MODELS_C];
    __shared__ long long spair[BLU_MAX_MODELS_C + 1];     // first flat pair index of every class
    unsigned short *slut = reinterpret_cast<unsigned short *>(kraw + 2 * M * MP + (size_t)BLU_KKT_WARPS * GPW * SCR);
    for (int t = threadIdx.x; t < M * M; t += blockDim.x) { sR[(t / M) * MP + (t % M)] = rinv[t]; sW[(t / M) * MP + (t % M)] = Wm[t]; }
    for (int t = threadIdx.x; t < ncls; t += blockDim.x) scls[t] = cls[t];
    for (int t = threadIdx.x; t < lutlen; t += blockDim.x) slut[t] = lut[t];
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int ic = 0; ic < ncls; ++ic) { spair[ic] = acc; acc += (cls[ic].Lk + GPW - 1) / GPW; }
        spair[ncls] = acc;
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / LW, a = lane % LW;
    const long long n = L + has_t;
    const double rt2 = 1.4142135623730951;
    // scratch of this sub-warp: [sCf / srow: KSF*KSF][sRg: KMAX*LW][sT: KMAX*LW][sg: 32 ints]
    double *scr = kraw + 2 * M * MP + (size_t)(w * GPW + sub) * SCR;
    double *sCf = scr, *srow = scr, *sRg = scr + KSF * KSF, *sT = sRg + KMAX * LW;
    int *sg = reinterpret_cast<int *>(sT + KMAX * LW);
    const long long gwarp = (long long)blockIdx.x * BLU_KKT_WARPS + w, nwarps = (long long)gridDim.x * BLU_KKT_WARPS;
    if (has_t && gwarp == 0) {
        // the t column, X = -E_NN: r^-1 X r^-T = -(column N of r^-1)(column N of r^-1)^T
        double *row = Bs;
        const double dc = d[0];
        for (int q = lane; q < Qs; q += 32) {
            int ra = 0, rem = q;
            while (rem >= M - ra) { rem -= M - ra; ++ra; }
            const int rb = ra + rem;
            const double v = -(sR[ra * MP + N] * sR[rb * MP + N]);
            row[q] = dc * (ra == rb ? v : rt2 * v);
        }
        for (int q = Qs + lane; q < QP; q += 32) row[q] = (q < Q) ? dc * Gx[(long long)(q - Qs) * n] / dlin[q - Qs] : 0.0;
        if (lane == 0) g1tw[0] = -sW[N * MP + N];
    }
    const long long total = spair[ncls];
    // operands of one pair, fetched one pair ahead
    double cv[CV], cvn[CV], dc = 0.0, dcn = 0.0;
    int gv = 0, gvn = 0;
    auto fetch = [&](int ic, long long fp, double (&c8)[CV], int &g1, double &d1) {
        const BluClass &ci = scls[ic];
        const long long il = (fp - spair[ic]) * GPW + sub;
        const bool act = il < ci.Lk;
        const double *C = cinv + ci.coff + il * ci.T;
#pragma unroll
        for (int u = 0; u < CV; ++u) c8[u] = (act && a + u * LW < ci.T) ? __ldg(C + a + u * LW) : 0.0;
        g1 = (act && a < ci.k) ? (int)__ldg(gidx + ci.ioff + il * ci.k + a) : 0;
        d1 = act ? __ldg(d + ci.goff + il + has_t) : 0.0;
    };
    long long fp = gwarp;
    int ic = 0;
    while (ic < ncls && fp >= spair[ic + 1]) ++ic;
    if (fp < total) fetch(ic, fp, cv, gv, dc);
    while (fp < total) {
        const BluClass ci = scls[ic];
        const int k = ci.k, T = ci.T;
        const unsigned short *jlt = slut + ci.lutoff;
        const long long il = (fp - spair[ic]) * GPW + sub;
        const bool active = il < ci.Lk;
        const long long col = ci.goff + il + has_t;
        const long long fpn = fp + nwarps;
        int icn = ic;
        while (icn < ncls && fpn >= spair[icn + 1]) ++icn;
        if (fpn < total) fetch(icn, fpn, cvn, gvn, dcn);
        // ---- stage: member ids, expanded inverse, <X, Wm>, member columns of r^-1 ----
        if (a < k) sg[a] = gv;
        __syncwarp();
        double dotw = 0.0;
#pragma unroll
        for (int u = 0; u < CV; ++u) {
            const int e = a + u * LW;
            if (e < T) {
                const unsigned jl = jlt[e];
                const int j = jl >> 8, l = jl & 255u;
                const double c = cv[u];
                sCf[j * KSF + l] = c;
                sCf[l * KSF + j] = c;
                const int gj = sg[j], gl = sg[l];
                const double wv = (j == l) ? sW[gj * MP + gj] : (sW[gj * MP + gl] + sW[gl * MP + gj]);
                dotw = fma(c, wv, dotw);
            }
        }
        if (a < M)
            for (int j = 0; j < k; ++j) sRg[j * LW + a] = sR[a * MP + sg[j]];
        __syncwarp();
        // ---- Tt[a][l] = sum_j Rg[a][j] C[j][l] in registers (entries l >= k are junk and never used) ----
        {
            double tt[2 * NH];
#pragma unroll
            for (int l = 0; l < 2 * NH; ++l) tt[l] = 0.0;
            for (int j = 0; j < k; ++j) {
                const double rg = sRg[j * LW + a];
                const double2 *cr = reinterpret_cast<const double2 *>(sCf + j * KSF);
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const double2 c2 = cr[h];
                    tt[2 * h] = fma(rg, c2.x, tt[2 * h]);
                    tt[2 * h + 1] = fma(rg, c2.y, tt[2 * h + 1]);
                }
            }
#pragma unroll
            for (int l = 0; l < 2 * NH; ++l)
                if (l < k) sT[l * LW + a] = tt[l];
        }
        __syncwarp();                                       // sCf is dead from here on: srow takes its place
        // ---- Y[a][b] = sum_l Tt[a][l] Rg[b][l] for every b (uniform over the lanes: broadcast reads) ----
        {
            double yy[2 * NH];
#pragma unroll
            for (int b = 0; b < 2 * NH; ++b) yy[b] = 0.0;
            for (int l = 0; l < k; ++l) {
                const double t = sT[l * LW + a];
                const double2 *rr = reinterpret_cast<const double2 *>(sRg + l * LW);
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const double2 r2 = rr[h];
                    yy[2 * h] = fma(t, r2.x, yy[2 * h]);
                    yy[2 * h + 1] = fma(t, r2.y, yy[2 * h + 1]);
                }
            }
            if (a < M) {
                double *rowq = srow + blu_svec(M, a, a) - a;
                const double f = -scales * dc;
#pragma unroll
                for (int b = 0; b < 2 * NH; ++b)
                    if (b >= a && b < M) rowq[b] = (a == b ? f : rt2 * f) * yy[b];
            }
        }
        __syncwarp();
        if (active) {
            double *row = Bs + col * QP;
            for (int q = a; q < Qs; q += LW) row[q] = srow[q];
            for (int q = Qs + a; q < QP; q += LW)           // column Q (the right-hand side) is written by blu_kkt_rhs_kernel
                row[q] = (q < Q) ? dc * Gx[(long long)(q - Qs) * n + col] / dlin[q - Qs] : 0.0;
        }
#pragma unroll
        for (int o = LW / 2; o > 0; o >>= 1) dotw += __shfl_xor_sync(BLU_FULL, dotw, o);
        if (active && a == 0) g1tw[col] = -scales * dotw;
        __syncwarp();                                       // scratch consumed before the next pair stages into it
        fp = fpn; ic = icn; gv = gvn; dc = dcn;
#pragma unroll
        for (int u = 0; u < CV; ++u) cv[u] = cvn[u];
    }
}

// rhs = bx + G0^T (d^-2 bz0) + G1^T vec(Lam Z Lam);  Bs[col][Q] = d[col] * rhs[col].
__global__ void blu_kkt_rhs_kernel(long long n, int nlin, const double *__restrict__ bx, const double *__restrict__ bz0,
                                   const double *__restrict__ d, const double *__restrict__ Gx, const double *__restrict__ dlin,
                                   const double *__restrict__ g1tw, int Q, int QP, double *__restrict__ rhs, double *__restrict__ Bs)
{
    for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < n; col += (long long)gridDim.x * blockDim.x) {
        double v = bx[col] - bz0[col] / (d[col] * d[col]);          // -I rows of G0
        for (int q = 0; q < nlin; ++q) v += Gx[(long long)q * n + col] * (bz0[n + q] / (dlin[q] * dlin[q]));
        v += g1tw[col];
        rhs[col] = v;
        Bs[col * QP + Q] = d[col] * v;
    }
}

// cap = I + Bs^T Bs and v = Bs^T (d * rhs) (column Q of Bs) on the FP64 tensor cores.  A CTA owns a contiguous range of
// rows: 32-row chunks are staged in shared memory by bulk asynchronous copies (one 8 QP-byte row per copy, 3 stages, row
// pitch QP + 4 doubles so that the four sample rows of a fragment fall into different banks), so an element of Bs crosses
// L2 -> SM once per CTA that needs it.  The NTQ x NTQ grid of 8 x 8 tiles is cut into 4 x 4 BLOCKS; every block that touches
// the upper triangle belongs to one warp (16 warps per CTA; gridDim.y CTAs share a row range when there are more than 16
// blocks).  Per 4 rows a warp loads 4 + 4 fragments (a lane's fragment of tile t: row 4 s + (lane & 3), column
// 8 t + (lane >> 2) -- the A operand for t = ti, the B operand for t = tj of mma.m8n8k4.f64) with immediate offsets and
// issues 16 DMMAs (10 in a block on the diagonal): no address arithmetic, no predicates in the loop.  Partial tiles per row range, folded in a fixed order
// by blu_kkt_capfold_kernel: no atomics, bit-reproducible.
// History (15 models, 32768 x 144): one CTA per tile pair and row split, fragments straight from L2: 108 us (every tile
// column re-read NTQ + 1 times, 717 MB through L2 for a 37.7 MB matrix); rows staged once but the pairs dealt to the warps
// through a run table: 85 us, 30 instructions per DMMA (per-pair address arithmetic and predicated reloads).
#define BLU_SYRK_WARPS 16
#define BLU_SYRK_RC 32
#define BLU_SYRK_NS 3
__host__ __device__ __forceinline__ size_t blu_kkt_syrk_smem(int QP) { return sizeof(double) * ((size_t)BLU_SYRK_NS * BLU_SYRK_RC * (QP + 4) + 64); }
__host__ __device__ __forceinline__ int blu_kkt_syrk_blocks(int QP) { const int nb = ((QP >> 3) + 3) >> 2; return nb * (nb + 1) / 2; }

// explicit shared-space load (32-bit address, immediate offsets fold into the instruction)
__device__ __forceinline__ double blu_kkt_lds(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(BLU_SYRK_WARPS * 32, 1)
blu_kkt_syrk_kernel(const double *__restrict__ Bs, long long n, int QP, double *__restrict__ part)
{
    extern __shared__ __align__(16) double ssm[];         // [NS][RC][PITCH] + slack for the fragments of padding tiles
    __shared__ __align__(8) unsigned long long bars[BLU_SYRK_NS];
    const int PITCH = QP + 4;
    const int NTQ = QP >> 3, npairs = NTQ * (NTQ + 1) / 2;
    const int nb = (NTQ + 3) >> 2;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int ks = lane & 3, cq = lane >> 2;
    if (tid == 0) {
        for (int s = 0; s < BLU_SYRK_NS; ++s) blu_mbar_init(&bars[s], 1);
        blu_mbar_fence_init();
    }
    __syncthreads();
    // this warp's block (bi <= bj) of the nb x nb block grid, or none
    int bi = 0, rem = blockIdx.y * BLU_SYRK_WARPS + w;
    while (bi < nb && rem >= nb - bi) { rem -= nb - bi; ++bi; }
    const bool have = bi < nb;
    const int bj = bi + rem;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    const long long rows_per = ((n + gridDim.x - 1) / gridDim.x + BLU_SYRK_RC - 1) / BLU_SYRK_RC * BLU_SYRK_RC;
    const long long r0 = (long long)blockIdx.x * rows_per;
    long long r1 = r0 + rows_per;
    if (r1 > n) r1 = n;
    const int nchunk = r1 > r0 ? (int)((r1 - r0 + BLU_SYRK_RC - 1) / BLU_SYRK_RC) : 0;
    auto issue = [&](int ch) {                            // warp 0, all lanes: one row per lane
        const int st = ch % BLU_SYRK_NS;
        const long long rb = r0 + (long long)ch * BLU_SYRK_RC;
        const int cr = (int)((r1 - rb) < BLU_SYRK_RC ? (r1 - rb) : BLU_SYRK_RC);
        if (lane == 0) blu_mbar_expect_tx(&bars[st], (unsigned)(cr * QP * 8));
        __syncwarp();
        if (lane < cr) blu_bulk_g2s(ssm + ((size_t)st * BLU_SYRK_RC + lane) * PITCH, Bs + (rb + lane) * QP, (unsigned)(QP * 8), &bars[st]);
    };
    if (w == 0)
        for (int ch = 0; ch < BLU_SYRK_NS - 1 && ch < nchunk; ++ch) issue(ch);
    const unsigned s_base = blu_smem_u32(ssm);
    // byte offset of this lane's fragment of tile 4 bi (A side) / 4 bj (B side) inside a stage, row ks
    const unsigned offA = (unsigned)((ks * PITCH + cq + 32 * bi) * 8), offB = (unsigned)((ks * PITCH + cq + 32 * bj) * 8);
    for (int ch = 0; ch < nchunk; ++ch) {
        const int st = ch % BLU_SYRK_NS;
        if (w == 0 && ch + BLU_SYRK_NS - 1 < nchunk) issue(ch + BLU_SYRK_NS - 1);      // the stage consumed one iteration ago (barrier below)
        blu_mbar_wait(&bars[st], (unsigned)((ch / BLU_SYRK_NS) & 1));
        const long long rb = r0 + (long long)ch * BLU_SYRK_RC;
        const int cr = (int)((r1 - rb) < BLU_SYRK_RC ? (r1 - rb) : BLU_SYRK_RC);
        if (cr < BLU_SYRK_RC) {                           // last chunk of the range: the missing rows count as zeros
            double *sb = ssm + (size_t)st * BLU_SYRK_RC * PITCH;
            for (int t = tid; t < (BLU_SYRK_RC - cr) * PITCH; t += blockDim.x) sb[cr * PITCH + t] = 0.0;
            __syncthreads();
        }
        if (have) {
            const unsigned sa = s_base + (unsigned)(st * BLU_SYRK_RC * PITCH * 8);
#pragma unroll 2
            for (int r4 = 0; r4 < BLU_SYRK_RC / 4; ++r4) {
                const unsigned ra = sa + (unsigned)(r4 * 4 * PITCH * 8);
                double fa[4], fb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { fa[i] = blu_kkt_lds(ra + offA + 64u * i); fb[i] = blu_kkt_lds(ra + offB + 64u * i); }
                if (bi != bj) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) blu_dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
                } else {                                  // a block on the diagonal: its tiles below the diagonal are never stored
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = i; j < 4; ++j) blu_dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
                }
            }
        }
        __syncthreads();                                  // stage consumed before it is refilled
    }
    // C fragment: row cq, cols 2 ks + {0,1} of the 8 x 8 tile; tiles below the diagonal or in the padding are dropped
    if (have) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ti = 4 * bi + i, tj = 4 * bj + j;
                if (ti <= tj && tj < NTQ) {
                    const int p = ti * NTQ - ti * (ti - 1) / 2 + (tj - ti);
                    double *dst = part + ((size_t)blockIdx.x * npairs + p) * 64 + cq * 8 + 2 * ks;
                    dst[0] = acc[i][j][0];
                    dst[1] = acc[i][j][1];
                }
            }
    }
}

// cap = I + sum over the row ranges (fixed order) of the partial tiles; v = column Q.
// Output in the layout blu_kkt_chol_kernel factors in place: COLUMN-major with leading dimension LD >= Q + 1, element (i, j)
// at cap[j * LD + i]; the right-hand side rides along as the extra ROW Q of the matrix (cap[j * LD + Q] = v_j).
// Eight threads per element, each sums a contiguous eighth of the row ranges with its loads in flight together; the
// eighths are combined by a fixed shuffle tree.
__global__ void __launch_bounds__(256)
blu_kkt_capfold_kernel(const double *__restrict__ part, int npairs, int nsplit, int Q, int QP, int LD, double *__restrict__ cap)
{
    const int NTQ = QP >> 3;
    const int t = blockIdx.x * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7;       // 32 elements (half a tile) per CTA
    const int p = t >> 6, e = t & 63;
    double s = 0.0;
    if (p < npairs) {
        const int nq = (nsplit + 7) >> 3;
        const int s0 = sub * nq, s1 = min(nsplit, s0 + nq);
        const double *src = part + (size_t)p * 64 + e;
        const size_t stride = (size_t)npairs * 64;
        int sp = s0;
        for (; sp + 8 <= s1; sp += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (size_t)(sp + u) * stride);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        if (sp < s1) {                                        // the remainder as one predicated batch
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (sp + u < s1) ? __ldg(src + (size_t)(sp + u) * stride) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
    }
    s += __shfl_xor_sync(BLU_FULL, s, 1);
    s += __shfl_xor_sync(BLU_FULL, s, 2);
    s += __shfl_xor_sync(BLU_FULL, s, 4);
    if (p >= npairs || sub != 0) return;
    int ti = 0, rem = p;
    while (rem >= NTQ - ti) { rem -= NTQ - ti; ++ti; }
    const int tj = ti + rem;
    const int r = 8 * ti + (e >> 3), c = 8 * tj + (e & 7);
    if (r < Q && c < Q) {
        if (ti != tj || r <= c) {
            const double val = s + (r == c ? 1.0 : 0.0);
            cap[(size_t)r * LD + c] = val;
            cap[(size_t)c * LD + r] = val;
        }
    } else if (c == Q && r < Q) cap[(size_t)r * LD + Q] = s;
}

// Cholesky of the Q x Q SPD capacitance matrix and the solve cap y = v: ONE CTA, blocked right-looking, in place in global
// memory (the matrix, <= 0.5 MB, lives in L2; only a 16-column panel is ever in shared memory, so Q is not limited by the
// 227 KB of one SM -- the previous whole-matrix-in-shared version stopped at Q ~ 170, i.e. 16 models).
//   cap   column-major lower triangle, (i, j) at cap[j*LD + i], rows 0..Q: row Q is the right-hand side v.  Carrying v as
//         an extra row makes its factor row L[Q][:] = L^-1 v -- the forward substitution costs nothing extra.
// Per panel of 16 columns: (1) panel -> shared (column-major, conflict-free in the row index); (2) warp 0 factors the 16 x 16
// diagonal block in registers (lane = row, shuffles broadcast the pivot column); (3) one thread per row below applies
// L_d^-T (136 multiply-adds against broadcast shared reads); (4) panel -> global; (5) trailing update
// A[i][j] -= sum_c P[i][c] P[j][c]: a warp task is 64 rows x 8 columns (lane: rows l, l+32; coalesced global
// read-modify-write, panel operands from shared), tasks entirely above the diagonal are skipped.
// Then the backward substitution L^T y = y' by 16-column blocks from the end (warp 0 solves the block, all threads
// propagate it into the remaining right-hand side).
// Measured (tools/lab/chol_lab.cu, Q = 138): 80 us = 155 k cycles -- diagonal blocks 51 k (a 350-cycle dependent chain per
// column: shuffle, rsqrt, 15 shuffles, multiply-add), trailing updates 45 k (bound by ONE SM's FP64 and shared-memory issue,
// 64 x 8 tasks include the waste above the diagonal), rows below 14 k, panel load + store 19 k, backward substitution 26 k.
// Copying the whole matrix into shared memory first (it fits up to 16 models) changed the total by 3 %: L2 latency is not
// what limits it; the next step would be DMMA tile blocks for the trailing update.  Also measured: factoring the next diagonal
// block one panel ahead (warp 0) while the other warps finish the trailing update -- correct, but the KKT solve stayed at
// 0.204-0.209 ms (0.206-0.208 before): the update that can hide behind the chain is too short at these orders.  Not kept.
#define BLU_CHOL_T 512
#ifdef BLU_CHOL_STAMPS                  // lab only (tools/lab/chol_lab.cu): cycles per phase, accumulated by thread 0
__device__ long long blu_chol_cycles[8];
#define BLU_CHOL_STAMP(i) do { if (threadIdx.x == 0) { const long long now_ = clock64(); blu_chol_cycles[i] += now_ - t_last_; t_last_ = now_; } } while (0)
#else
#define BLU_CHOL_STAMP(i) do { } while (0)
#endif
#define BLU_CHOL_PITCH 272            // rows of a panel (<= 256) + the overhang of masked rows in the trailing update
__global__ void __launch_bounds__(BLU_CHOL_T)
blu_kkt_chol_kernel(double *cap, int Q, int LD, double *__restrict__ y, int *__restrict__ info)
{
    __shared__ __align__(16) double sP[16 * BLU_CHOL_PITCH];      // sP[c][r] = panel column c, row j0 + r
    __shared__ double sD[16 * 17];                                // sD[i][s] = L_d[i][s]
    __shared__ double sinv[16];
    __shared__ double sy[272];
    __shared__ double sdinv[272];                                 // 1 / L[j][j]: the backward substitution multiplies
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int R = Q + 1;
#ifdef BLU_CHOL_STAMPS
    long long t_last_ = clock64();
#endif
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int j0 = 0; j0 < Q; j0 += 16) {
        const int w = min(16, Q - j0);
        const int nrow = R - j0;
        for (int t = tid; t < 16 * nrow; t += BLU_CHOL_T) {
            const int c = t / nrow, r = t - c * nrow;
            sP[c * BLU_CHOL_PITCH + r] = (c < w) ? cap[(size_t)(j0 + c) * LD + j0 + r] : 0.0;
        }
        __syncthreads();
        BLU_CHOL_STAMP(0);
        if (wid == 0) {
            double x[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = sP[c * BLU_CHOL_PITCH + (lane & 15)];
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                if (s < w) {
                    const double dss = __shfl_sync(BLU_FULL, x[s], s);
                    if (!(dss > 0.0) && lane == 0 && s_bad == 0) s_bad = j0 + s + 1;
                    const double inv = rsqrt(dss), ls = dss * inv;      // one long-latency operation on the critical chain, not two
                    x[s] = (lane == s) ? ls : x[s] * inv;
#pragma unroll
                    for (int c = s + 1; c < 16; ++c) {
                        const double lcs = __shfl_sync(BLU_FULL, x[s], c);
                        x[c] = fma(-x[s], lcs, x[c]);
                    }
                    if (lane == 0) { sinv[s] = inv; sdinv[j0 + s] = inv; }
                }
            }
            if (lane < w) {
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    const double vv = (s <= lane) ? x[s] : 0.0;
                    sD[lane * 17 + s] = vv;
                    sP[s * BLU_CHOL_PITCH + lane] = vv;
                }
            }
        }
        __syncthreads();
        if (s_bad) { if (tid == 0) *info = s_bad; return; }         // same value in every thread
        BLU_CHOL_STAMP(1);
        if (tid < nrow - w) {
            const int r = w + tid;
            double x[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = sP[c * BLU_CHOL_PITCH + r];
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                if (s < w) {
                    x[s] *= sinv[s];
#pragma unroll
                    for (int c = s + 1; c < 16; ++c) x[c] = fma(-x[s], sD[c * 17 + s], x[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) sP[c * BLU_CHOL_PITCH + r] = (c < w) ? x[c] : 0.0;
        }
        __syncthreads();
        BLU_CHOL_STAMP(2);
        for (int t = tid; t < w * nrow; t += BLU_CHOL_T) {
            const int c = t / nrow, r = t - c * nrow;
            cap[(size_t)(j0 + c) * LD + j0 + r] = sP[c * BLU_CHOL_PITCH + r];
        }
        BLU_CHOL_STAMP(3);
        const int j1 = j0 + 16;
        if (j1 < Q) {
            const int n_r = R - j1, n_c = Q - j1;
            const int nrb = (n_r + 63) >> 6, ncg = (n_c + 7) >> 3;
            for (int task = wid; task < nrb * ncg; task += BLU_CHOL_T / 32) {
                const int rb = task / ncg, cg = task - rb * ncg;
                if (rb * 64 + 63 < 8 * cg) continue;                 // the whole task lies above the diagonal
                const int ia = rb * 64 + lane, ib = ia + 32;
                double acc[2][8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int col = 8 * cg + cc;
                    const size_t base = (size_t)(j1 + col) * LD + j1;
                    acc[0][cc] = (col < n_c && ia < n_r) ? cap[base + ia] : 0.0;
                    acc[1][cc] = (col < n_c && ib < n_r) ? cap[base + ib] : 0.0;
                }
#pragma unroll 4
                for (int c = 0; c < 16; ++c) {
                    const double *pc = sP + c * BLU_CHOL_PITCH + 16;
                    const double a0 = pc[ia], a1 = pc[ib];
                    const double2 *pb = reinterpret_cast<const double2 *>(pc + 8 * cg);
                    const double2 b01 = pb[0], b23 = pb[1], b45 = pb[2], b67 = pb[3];
                    const double b[8] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y, b67.x, b67.y};
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) { acc[0][cc] = fma(-a0, b[cc], acc[0][cc]); acc[1][cc] = fma(-a1, b[cc], acc[1][cc]); }
                }
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int col = 8 * cg + cc;
                    const size_t base = (size_t)(j1 + col) * LD + j1;
                    if (col < n_c && ia < n_r) cap[base + ia] = acc[0][cc];
                    if (col < n_c && ib < n_r) cap[base + ib] = acc[1][cc];
                }
            }
        }
        __syncthreads();
        BLU_CHOL_STAMP(4);
    }
    // backward substitution: L^T y = y' (y' = row Q of the factor)
    for (int t = tid; t < Q; t += BLU_CHOL_T) sy[t] = cap[(size_t)t * LD + Q];
    __syncthreads();
    for (int jb = ((Q - 1) >> 4) << 4; jb >= 0; jb -= 16) {
        const int w = min(16, Q - jb);
        if (wid == 0) {
            double Lc[16];                                        // lane c: column c of the diagonal block, Lc[s] = L[jb+s][jb+c]
#pragma unroll
            for (int s = 0; s < 16; ++s) Lc[s] = (lane < w && s >= lane && s < w) ? cap[(size_t)(jb + lane) * LD + jb + s] : 0.0;
            double rc = (lane < w) ? sy[jb + lane] : 0.0;
            const double dinv = (lane < w) ? sdinv[jb + lane] : 0.0;
#pragma unroll
            for (int s = 15; s >= 0; --s) {
                if (s < w) {
                    const double ys = __shfl_sync(BLU_FULL, rc * dinv, s);
                    if (lane == s) rc = ys;
                    else if (lane < s) rc = fma(-Lc[s], ys, rc);
                }
            }
            if (lane < w) sy[jb + lane] = rc;
        }
        __syncthreads();
        BLU_CHOL_STAMP(5);
        for (int i = tid; i < jb; i += BLU_CHOL_T) {
            const double *col = cap + (size_t)i * LD + jb;
            double lv[16];
#pragma unroll
            for (int s = 0; s < 16; ++s) lv[s] = (s < w) ? col[s] : 0.0;
            double acc = 0.0;
#pragma unroll
            for (int s = 0; s < 16; ++s) acc = fma(lv[s], (s < w) ? sy[jb + s] : 0.0, acc);
            sy[i] -= acc;
        }
        __syncthreads();
        BLU_CHOL_STAMP(6);
    }
    for (int t = tid; t < Q; t += BLU_CHOL_T) y[t] = sy[t];
    if (tid == 0) *info = 0;
}

// ux[col] = d[col]^2 rhs[col] - d[col] (Bs[col][0..Q) . y).  One warp per column.
__global__ void __launch_bounds__(BLU_KKT_WARPS * 32)
blu_kkt_apply_kernel(const double *__restrict__ Bs, long long n, int Q, int QP, const double *__restrict__ y,
                     const double *__restrict__ d, const double *__restrict__ rhs, double *__restrict__ ux)
{
    __shared__ double sy[256];
    for (int t = threadIdx.x; t < Q; t += blockDim.x) sy[t] = y[t];
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long col = (long long)blockIdx.x * BLU_KKT_WARPS + w; col < n; col += (long long)gridDim.x * BLU_KKT_WARPS) {
        const double *row = Bs + col * QP;
        double v[8];                                      // Q <= 255: the whole row in flight at once (one L2 round trip per row, not one per 32 entries)
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (lane + 32 * u < Q) ? __ldg(row + lane + 32 * u) : 0.0;
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s = fma(v[u], (lane + 32 * u < Q) ? sy[lane + 32 * u] : 0.0, s);
        s = blu_warp_sum(s);
        if (lane == 0) ux[col] = d[col] * (d[col] * rhs[col] - s);
    }
}
