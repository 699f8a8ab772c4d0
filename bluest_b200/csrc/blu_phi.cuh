// blu_phi.cuh -- kernel (2): Phi(m) = delta I + sum_i m_i R_i^T Cinv_i R_i   (replaces the dense
// GEMV psi@m of misc.py:459-461 / the sparse loop objectiveK_c, cmisc.cpp:25-40), followed by the
// N x N pseudo-inverse and the variance (misc.py:463-477, 487-490).
//
// blu_phi_partial_kernel: each warp streams chunks of consecutive groups of the packed inverse
// array through its shared-memory ring (bulk async copies completing on mbarriers).  A lane owns packed
// entry (j,l) of every group of a class and sums m_i * Cinv_i[j,l] in a REGISTER for as long as the
// entry's target (g[j], g[l]) stays the same -- in the reference's enumeration order that is a run of
// many groups -- then adds the run into the warp's PRIVATE N x N accumulator tile in shared memory
// ("the streaming pass" below).  Inside one flush all targets are distinct, so the read-modify-write
// needs no atomics.  Only the upper triangle is ever touched (groups are sorted, j <= l).  Warps are
// combined in a fixed order into one partial tile per CTA -- atomic-free and run-to-run deterministic.
// Groups with m_i == 0 are skipped (they contribute exact zeros).
// Also reduced here: the support mask (models touched by groups with |m_i| > 1e-6,
// misc.py:453-457) and max|m| (early-out of misc.py:464,484) -- both by order-independent
// integer atomics (OR, MAX on the IEEE bit pattern), hence deterministic.
// The last CTAs to arrive fold the partial tiles (two levels, fixed order) and the very last one does
// the finish step in the same launch.
//
// blu_phi_finish_kernel (one CTA): fixed-order sum of the CTA partials, mirror, + delta I, then
// pinv(Phi) by parallel Jacobi, x = first row, S = 2 pinv(Phi), and the variance from the support
// sub-block exactly as misc.py:489-490 does.
#pragma once
#include "blu_common.cuh"
#include "blu_jacobi.cuh"
#include "blu_stream.cuh"

#define BLU_PHI_WARPS 16            // most warps per CTA of the streaming pass (8 when the tiles are large)
// dynamic shared memory of blu_phi_partial_kernel (layout: see the kernel)
__host__ __device__ __forceinline__ size_t blu_phi_smem_bytes(int ns, int sd, int idsd, int N, int ncls, int warps)
{
    return sizeof(double) * ((size_t)warps * ns * sd + (size_t)warps * N * N) + sizeof(BluClass) * ncls
           + sizeof(unsigned long long) * warps * 4 + 16 + (size_t)warps * 512 + (size_t)warps * ns * idsd;
}

__device__ __forceinline__ unsigned long long blu_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct BluEvalHeader {          // small device-side status block of a context
    unsigned supp;              // support mask (OR)
    unsigned flags;             // BLU_FLAG_* of the last evaluation
    unsigned long long maxbits; // bit pattern of max|m|
    double scal[8];             // [0] variance [1] max|m| [2] sweeps [3] lambda_max [4] var (full pinv) [5] BLUE mean
    double xsup[32];            // first row of pinv(Phi[idx,idx]) scattered to model slots (PHIinvY0, misc.py:529-533)
    unsigned long long epoch;   // evaluations of the fused multi-GPU path so far (device side: survives CUDA-graph replays)
    unsigned tickets[40];       // arrival counters of the fused Phi reduction: [0] CTA groups, [1 + g] CTAs of group g (self-resetting)
    unsigned long long stamp[16]; // %globaltimer (ns) at the milestones of the last CTA of the fused Phi kernel (profiling aid, see BLU_STAMP)
};
#define BLU_STAMP(hdr, i) do { if (threadIdx.x == 0) (hdr)->stamp[i] = blu_globaltimer(); } while (0)

// Explicit shared-space accesses with 32-bit addresses.  Through the generic pointers of the stream
// structs the compiler emitted generic LD.E.64 for the staged values and rebuilt the shared window base
// for every accumulator access (64-bit address arithmetic per group); these keep it to LDS/STS.
__device__ __forceinline__ double blu_lds_f64(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
// staged values: read-only while a chunk is consumed -- no memory clobber, so the loads of the next group may be
// issued ahead of the tile updates of the current one (volatile keeps them behind the mbarrier wait)
__device__ __forceinline__ double blu_lds_f64_ro(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void blu_sts_f64(unsigned addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// ---- the streaming pass -------------------------------------------------------------------------------------
// Run-length accumulation in registers over a per-warp bulk-copy ring.
//
// In the enumeration order of sap.py:73 (itertools.combinations: last member fastest) consecutive groups of a class
// share their leading members, so packed entry (j,l) -- target Phi[ids[j], ids[l]] -- keeps its target over a RUN of
// groups as long as members 0..l stay put.  A lane therefore sums m_i Cinv_i[j,l] in a REGISTER and touches the
// warp's shared tile only when its target changes (or the warp leaves the class): most of the read-modify-write
// traffic, whose bank conflicts were 41 % of the round-1 kernel's shared-memory wavefronts, is gone.  The lane <->
// entry table (plut, built by the host) sorts the entries by l descending, so only the first step(s) of a group ever
// change target, and packs them into half-warps whose 16 staged values sit in 16 different banks (the staged loads
// stay conflict-free although the lanes no longer read consecutive entries).
//
// A warp walks its own list of chunks (wchunks, laid out per warp by the host: several CONTIGUOUS, cost-balanced
// runs of the chunk list, so the register runs carry on across chunk boundaries) through an NS-stage ring: a stage
// holds the packed inverses of one chunk (bulk copy 1) and the member ids of its groups (bulk copy 2: k bytes per
// group straight from gidx, no mask expansion), both completing on the stage's mbarrier; copies are issued NS-1
// chunks ahead.  Per chunk the warp works out eff_i = leading members shared with the previous SAMPLED group from
// the masks (one group per lane) and leaves 16-byte records {m_i, eff_i} in shared memory for the group loop.
// The order of the additions is fixed by the chunk lists alone: bit-reproducible.  Any group order is handled (a
// target change is detected from the masks, not assumed); orders without shared prefixes just flush every group.
//   plut word: bits 0-9 packed position, 10-14 j, 15-19 l, 20 valid
#define BLU_PHI_RUN_MAXSTEPS 9
#define BLU_PHI_MAXSTAGES 4
#ifdef BLU_PHI_PROFILE
__device__ long long blu_prof_warp[8192][4];     // per warp of the grid: total cycles, group-loop cycles, chunks, groups
#endif

struct BluPhiStream {
    unsigned val_s, ids_s, bar_s;   // shared addresses of stage 0 (values, ids) and mbarrier 0 of this warp
    unsigned sdb, idsd;             // stage pitch in bytes (values, ids)
    unsigned rec_s;                 // the 32 records
};
struct BluPhiChunkRegs {
    int cls, g, skew, skewb;        // skew: bytes to skip in the value stage, skewb: bytes to skip in the id stage
    double m;                       // m of group (i0 + lane), 0 beyond g
    unsigned mask;                  // membership mask of group (i0 + lane)
};

__device__ __forceinline__ void blu_mbar_expect_tx_s(unsigned bar_s, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void blu_bulk_g2s_s(unsigned dst_s, const void *src, unsigned bytes, unsigned bar_s)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}
__device__ __forceinline__ void blu_mbar_wait_s(unsigned bar_s, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BLU_WAIT_S:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BLU_DONE_S;\n"
        "bra BLU_WAIT_S;\n"
        "BLU_DONE_S:\n"
        "}\n" ::"r"(bar_s), "r"(parity) : "memory");
}

__device__ __forceinline__ BluPhiChunkRegs blu_phi_prefetch(const BluChunk ch, const BluClass *__restrict__ scls,
                                                            const double *__restrict__ cinv, const unsigned char *__restrict__ gidx,
                                                            const double *__restrict__ m, const unsigned *__restrict__ gmask,
                                                            const BluPhiStream &ps, const int st, const int lane)
{
    BluPhiChunkRegs r;
    const BluClass ci = scls[ch.cls];
    const unsigned long long addr = (unsigned long long)(cinv + ci.coff + ch.i0 * ci.T);
    const unsigned long long addrb = (unsigned long long)(gidx + ci.ioff + ch.i0 * ci.k);
    r.skew = (int)(addr & 15ull);
    r.skewb = (int)(addrb & 15ull);
    r.cls = ch.cls; r.g = ch.g;
    if (lane == 0) {
        const unsigned bytes = (unsigned)((r.skew + ch.g * ci.T * 8 + 15) & ~15);
        const unsigned bytesb = (unsigned)((r.skewb + ch.g * ci.k + 15) & ~15);
        const unsigned bar = ps.bar_s + 8u * st;
        blu_mbar_expect_tx_s(bar, bytes + bytesb);
        blu_bulk_g2s_s(ps.val_s + ps.sdb * st, (const void *)(addr & ~15ull), bytes, bar);
        blu_bulk_g2s_s(ps.ids_s + ps.idsd * st, (const void *)(addrb & ~15ull), bytesb, bar);
    }
    const long long gi = ci.goff + ch.i0 + lane;
    const bool ok = lane < ch.g;
    r.m = ok ? m[gi] : 0.0;
    r.mask = ok ? gmask[gi] : 0u;
    return r;
}

__device__ __forceinline__ unsigned blu_lds_u8(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// predicated shared loads that leave the destination untouched when the predicate is off (no per-iteration zeroing)
__device__ __forceinline__ void blu_lds_u8_if(unsigned &dst, unsigned addr, int on)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.s32 p, %2, 0;\n@p ld.shared.u8 %0, [%1];\n}\n" : "+r"(dst) : "r"(addr), "r"(on));
}
__device__ __forceinline__ void blu_lds_f64_if(double &dst, unsigned addr, int on)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.s32 p, %2, 0;\n@p ld.shared.f64 %0, [%1];\n}\n" : "+d"(dst) : "r"(addr), "r"(on));
}

// The groups of one landed chunk.  DENSE: every group of the chunk is sampled (no per-group test).
// lrest: largest l among the entries of steps 1..S-1 (warp-uniform): those steps are looked at only when a group
// moves a member that low.  n8 = 8 N.
template <int S, bool DENSE>
__device__ __forceinline__ void blu_phi_groups(const int ng, const unsigned live, unsigned cp_s, unsigned idp_s, unsigned rp_s,
                                               const int T, const int k, const unsigned n8, const int pmin, const int lrest, const unsigned acc_s,
                                               const unsigned (&off)[S], const int (&js)[S], const int (&ls)[S],
                                               unsigned (&pt)[S], double (&ar)[S])
{
#ifdef BLU_PHI_NOCOMPUTE
    return;                                                       // lab switch: the ring alone (wrong results)
#endif
    const int early = ls[0] >= pmin;                              // pmin: smallest eff of the chunk -- lanes below it never move here
    unsigned a0 = 0u, b0 = 0u;
    double old0 = 0.0;
    for (int g = 0; g < ng; ++g, cp_s += 8u * T, idp_s += k, rp_s += 16u) {
        if (!DENSE && !((live >> g) & 1u)) continue;              // m_i == 0 contributes exact zeros
        // every shared load of the group is issued up front (one latency per group, not a chain of them): the
        // record, the staged values, and for step 0 -- whose lanes change target with almost every group -- the
        // candidate member ids and the pending tile entry
        unsigned lo, hi, e, pad;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo), "=r"(hi), "=r"(e), "=r"(pad) : "r"(rp_s) : "memory");
        double v[S];
#pragma unroll
        for (int s = 0; s < S; ++s) v[s] = blu_lds_f64_ro(cp_s + off[s]);
        blu_lds_u8_if(a0, idp_s + js[0], early);
        blu_lds_u8_if(b0, idp_s + ls[0], early);
        blu_lds_f64_if(old0, pt[0], early);
        const double mi = __hiloint2double((int)hi, (int)lo);
        const int pend = (int)e;
        if (ls[0] >= pend) {                                      // this lane's target moves with this group
            blu_sts_f64(pt[0], old0 + ar[0]);
            ar[0] = 0.0;
            pt[0] = acc_s + a0 * n8 + 8u * b0;
        }
        ar[0] = fma(mi, v[0], ar[0]);
        if (S > 1 && lrest >= pend) {                             // warp-uniform (pend comes from the record, lrest from the table)
#pragma unroll
            for (int s = 1; s < S; ++s) {
                if (ls[s] >= pend) {
                    const unsigned a = blu_lds_u8(idp_s + js[s]), b = blu_lds_u8(idp_s + ls[s]);
                    blu_sts_f64(pt[s], blu_lds_f64(pt[s]) + ar[s]);
                    ar[s] = 0.0;
                    pt[s] = acc_s + a * n8 + 8u * b;
                }
            }
        }
#pragma unroll
        for (int s = 1; s < S; ++s) ar[s] = fma(mi, v[s], ar[s]);
        __syncwarp();                                             // tile updates of different groups stay ordered
    }
}

// Loop state of a warp's walk over its chunk list.
template <int NS>
struct BluPhiWalk {
    int c, c1, it;                  // next chunk to consume, end of the list, chunks consumed so far (ring position)
    BluPhiChunkRegs q[NS - 1];      // registers of chunks c .. c+NS-2 (copies in flight or landed)
    BluChunk dnext;                 // descriptor of chunk c + NS - 1, fetched one chunk early
    unsigned supp;
    double mymax;
#ifdef BLU_PHI_PROFILE
    long long tpre, texp, twait, tloop;
#endif
};

// Consume chunks c, c+1, ... of the walk for as long as they belong to class `mycls` (S steps per group).
template <int S, int NS>
__device__ __forceinline__ void blu_phi_class_run(BluPhiWalk<NS> &wk, const int mycls, const BluClass ci, const unsigned *__restrict__ plt,
                                                  const BluChunk *__restrict__ chunks, const BluClass *__restrict__ scls,
                                                  const double *__restrict__ cinv, const unsigned char *__restrict__ gidx,
                                                  const double *__restrict__ m, const unsigned *__restrict__ gmask,
                                                  const BluPhiStream &ps, double *__restrict__ acc, const int N, const int lane)
{
    constexpr int D = NS - 1;
    unsigned off[S], pt[S];
    int js[S], ls[S];
    double ar[S];
    const unsigned acc_s = blu_smem_u32(acc);
    const int T = ci.T, k = ci.k;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const unsigned wd = __ldg(plt + s * 32 + lane);          // global (L1-resident): one coalesced line per step
        const bool valid = (wd >> 20) & 1u;
        off[s] = 8u * (wd & 1023u);
        js[s] = (int)((wd >> 10) & 31u);
        ls[s] = valid ? (int)((wd >> 15) & 31u) : -1;            // -1: idle lane of a partly filled step, never flushes
        ar[s] = 0.0;
        pt[s] = acc_s;                                           // nothing pending: the first flush adds 0.0 to acc[0]
    }
    int lrest = -1;
#pragma unroll
    for (int s = 1; s < S; ++s) lrest = max(lrest, ls[s]);
    lrest = __reduce_max_sync(BLU_FULL, lrest);
    const unsigned n8 = 8u * (unsigned)N;
    unsigned lastmask = 0u;
    bool havelast = false;
    do {
        const int st = wk.it & (NS - 1);
        BluPhiChunkRegs nxt;
#ifdef BLU_PHI_PROFILE
        const long long tp0 = clock64();
#endif
        if (wk.c + D < wk.c1) nxt = blu_phi_prefetch(wk.dnext, scls, cinv, gidx, m, gmask, ps, (wk.it + D) & (NS - 1), lane);
#ifdef BLU_PHI_PROFILE
        const long long tp1 = clock64(); wk.tpre += tp1 - tp0;
#endif
        if (wk.c + D + 1 < wk.c1) wk.dnext = chunks[wk.c + D + 1];
        const BluPhiChunkRegs cur = wk.q[0];
        const double am = fabs(cur.m);
        wk.mymax = fmax(wk.mymax, am);
        if (am > 1.0e-6) wk.supp |= cur.mask;
        const unsigned live = __ballot_sync(BLU_FULL, cur.m != 0.0);
        const unsigned bar = ps.bar_s + 8u * st;
        const unsigned parity = (unsigned)((wk.it / NS) & 1);
        if (live) {
            // members shared with the previous SAMPLED group (lane g: group g): entries with l below that keep their target
            const unsigned below = live & ((1u << lane) - 1u);
            const int pl = 31 - __clz(below);                     // -1: none in this chunk -> the last sampled group before it
            unsigned pm = __shfl_sync(BLU_FULL, cur.mask, pl < 0 ? 0 : pl);
            if (pl < 0) pm = lastmask;
            const unsigned x = cur.mask ^ pm;
            int eff = x ? __popc(cur.mask & ((1u << (__ffs(x) - 1)) - 1u)) : k;
            if (pl < 0 && !havelast) eff = 0;
            lastmask = __shfl_sync(BLU_FULL, cur.mask, 31 - __clz(live));
            havelast = true;
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ps.rec_s + 16u * lane), "r"((unsigned)__double2loint(cur.m)),
                         "r"((unsigned)__double2hiint(cur.m)), "r"((unsigned)eff), "r"(0u) : "memory");
            __syncwarp();
            const int pmin = __reduce_min_sync(BLU_FULL, ((live >> lane) & 1u) ? eff : 32);
#ifdef BLU_PHI_PROFILE
            const long long tw0 = clock64(); wk.texp += tw0 - tp1;
#endif
            blu_mbar_wait_s(bar, parity);                         // everything above overlapped the copies
#ifdef BLU_PHI_PROFILE
            const long long tw1 = clock64(); wk.twait += tw1 - tw0;
#endif
            const unsigned cp_s = ps.val_s + ps.sdb * st + (unsigned)cur.skew;
            const unsigned idp_s = ps.ids_s + ps.idsd * st + (unsigned)cur.skewb;
            if (live == (cur.g >= 32 ? BLU_FULL : (1u << cur.g) - 1u))
                blu_phi_groups<S, true>(cur.g, live, cp_s, idp_s, ps.rec_s, T, k, n8, pmin, lrest, acc_s, off, js, ls, pt, ar);
            else
                blu_phi_groups<S, false>(cur.g, live, cp_s, idp_s, ps.rec_s, T, k, n8, pmin, lrest, acc_s, off, js, ls, pt, ar);
#ifdef BLU_PHI_PROFILE
            wk.tloop += clock64() - tw1;
#endif
        } else {
            blu_mbar_wait_s(bar, parity);                         // the stage must have landed before it is reused
        }
        __syncwarp();                                             // stage and records consumed before they are rewritten
#pragma unroll
        for (int d = 0; d + 1 < D; ++d) wk.q[d] = wk.q[d + 1];
        wk.q[D - 1] = nxt;
        ++wk.c; ++wk.it;
    } while (wk.c < wk.c1 && wk.q[0].cls == mycls);
#pragma unroll
    for (int s = 0; s < S; ++s) {                                 // leaving the class: everything still in registers
        if (ls[s] >= 0) blu_sts_f64(pt[s], blu_lds_f64(pt[s]) + ar[s]);
        __syncwarp();
    }
}

// generic step count (groups of more than 23 members): classic walk, lane = packed entry, tile update per entry
__device__ __forceinline__ void blu_phi_chunk_any(const double *__restrict__ base, const unsigned short *__restrict__ lt, int T, int k, int ng,
                                                  unsigned live, double mreg, const unsigned char *__restrict__ ids,
                                                  double *__restrict__ acc, int N, int lane)
{
    for (int g = 0; g < ng; ++g) {
        if (!((live >> g) & 1u)) continue;
        const double mi = blu_shfl(mreg, g);
        const int gv = lane < k ? ids[g * k + lane] : 0;
        const double *cp = base + g * T;
        for (int e0 = 0; e0 < T; e0 += 32) {
            const int e = e0 + lane;
            const bool ok = e < T;
            const unsigned jl = ok ? __ldg(lt + e) : 0u;
            const double v = ok ? cp[e] : 0.0;
            const int a = __shfl_sync(BLU_FULL, gv, jl >> 8);
            const int b = __shfl_sync(BLU_FULL, gv, jl & 255u);
            if (ok) acc[a * N + b] += mi * v;
        }
        __syncwarp();
    }
}

// The whole streaming pass of one warp: chunks wstart[gw] .. wstart[gw+1]-1 of wchunks.
template <int NS>
__device__ __forceinline__ void blu_phi_walk(const int c0, const int c1, const BluChunk *__restrict__ chunks, const BluClass *__restrict__ scls,
                                             const double *__restrict__ cinv, const unsigned char *__restrict__ gidx,
                                             const unsigned short *__restrict__ lut, const unsigned *__restrict__ plut,
                                             const double *__restrict__ m, const unsigned *__restrict__ gmask,
                                             const BluPhiStream &ps, double *__restrict__ acc, const int N, const int lane,
                                             unsigned &supp_out, double &max_out, BluEvalHeader *blu_prof_hdr)
{
    constexpr int D = NS - 1;
    BluPhiWalk<NS> wk;
    wk.c = c0; wk.c1 = c1; wk.it = 0; wk.supp = 0u; wk.mymax = 0.0;
#ifdef BLU_PHI_PROFILE
    wk.tpre = wk.texp = wk.twait = wk.tloop = 0;
#endif
#pragma unroll
    for (int d = 0; d < D; ++d)
        if (c0 + d < c1) wk.q[d] = blu_phi_prefetch(chunks[c0 + d], scls, cinv, gidx, m, gmask, ps, d, lane);
    if (c0 + D < c1) wk.dnext = chunks[c0 + D];
    while (wk.c < wk.c1) {
        const int mycls = wk.q[0].cls;
        const BluClass ci = scls[mycls];
        const unsigned *plt = plut + ci.plutoff;
#define BLU_PHI_RUN_CASE(S_) case S_: blu_phi_class_run<S_, NS>(wk, mycls, ci, plt, chunks, scls, cinv, gidx, m, gmask, ps, acc, N, lane); break;
        switch (ci.psteps) {
            BLU_PHI_RUN_CASE(1) BLU_PHI_RUN_CASE(2) BLU_PHI_RUN_CASE(3) BLU_PHI_RUN_CASE(4) BLU_PHI_RUN_CASE(5)
            BLU_PHI_RUN_CASE(6) BLU_PHI_RUN_CASE(7) BLU_PHI_RUN_CASE(8) BLU_PHI_RUN_CASE(9)
            default: {                                    // very large groups: one chunk at a time, classic walk
                const int st = wk.it & (NS - 1);
                BluPhiChunkRegs nxt;
                if (wk.c + D < wk.c1) nxt = blu_phi_prefetch(wk.dnext, scls, cinv, gidx, m, gmask, ps, (wk.it + D) & (NS - 1), lane);
                if (wk.c + D + 1 < wk.c1) wk.dnext = chunks[wk.c + D + 1];
                const BluPhiChunkRegs cur = wk.q[0];
                const double am = fabs(cur.m);
                wk.mymax = fmax(wk.mymax, am);
                if (am > 1.0e-6) wk.supp |= cur.mask;
                const unsigned live = __ballot_sync(BLU_FULL, cur.m != 0.0);
                blu_mbar_wait_s(ps.bar_s + 8u * st, (unsigned)((wk.it / NS) & 1));
                if (live) {
                    const double *base = reinterpret_cast<const double *>(__cvta_shared_to_generic(ps.val_s + ps.sdb * st + (unsigned)cur.skew));
                    const unsigned char *idb = reinterpret_cast<const unsigned char *>(__cvta_shared_to_generic(ps.ids_s + ps.idsd * st + (unsigned)cur.skewb));
                    blu_phi_chunk_any(base, lut + ci.lutoff, ci.T, ci.k, cur.g, live, cur.m, idb, acc, N, lane);
                }
                __syncwarp();
#pragma unroll
                for (int d = 0; d + 1 < D; ++d) wk.q[d] = wk.q[d + 1];
                wk.q[D - 1] = nxt;
                ++wk.c; ++wk.it;
            } break;
        }
#undef BLU_PHI_RUN_CASE
    }
    supp_out = wk.supp; max_out = wk.mymax;
#ifdef BLU_PHI_PROFILE
    if (lane == 0) {
        atomicAdd(&blu_prof_hdr->stamp[13], (unsigned long long)wk.tpre); atomicAdd(&blu_prof_hdr->stamp[14], (unsigned long long)wk.texp);
        atomicAdd(&blu_prof_hdr->stamp[10], (unsigned long long)wk.twait); atomicAdd(&blu_prof_hdr->stamp[15], (unsigned long long)wk.tloop);
        const int gwp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (gwp < 8192) { blu_prof_warp[gwp][1] = wk.tloop; blu_prof_warp[gwp][2] = wk.tpre + wk.texp; blu_prof_warp[gwp][3] = wk.twait; }
    }
#endif
}

// Block-parallel in-place Gauss-Jordan inverse of the SPD n x n matrix in A (ld BLU_JLD), ping-pong
// with B; one __syncthreads per pivot.  Returns (uniformly) false as soon as a pivot falls below
// tol x its original diagonal entry (rank-deficient / indefinite): the caller then takes the
// Jacobi pseudo-inverse.  On success the inverse is in the buffer returned through *out.
__device__ __forceinline__ bool blu_block_gj(double *A, double *B, const double *diag0, int n, double tol,
                                             double **out, int tid, int nthr)
{
    double *src = A, *dst = B;
    // thread <-> (column c = tid & 31, rows r0, r0 + nthr/32, ...): no integer division in the pivot loop, the
    // pivot-row element is the same for all of a thread's rows, the pivot-column elements are warp broadcasts
    const int c = tid & 31, r0 = tid >> 5, rstep = nthr >> 5;
    for (int p = 0; p < n; ++p) {
        const double piv = src[p * BLU_JLD + p];
        if (!(piv > tol * diag0[p])) return false;            // same value in every thread
        const double d = __drcp_rn(piv);                      // correctly rounded reciprocal without the slow division path
        if (c < n) {
            const double *__restrict__ sp = src;              // ping-pong buffers: the rows of a thread are independent
            double *__restrict__ dp = dst;
            const double apc = sp[p * BLU_JLD + c];
#pragma unroll 4
            for (int r = r0; r < n; r += rstep) {
                const double arp = sp[r * BLU_JLD + p];
                double v;
                if (r == p) v = (c == p) ? d : apc * d;
                else if (c == p) v = -(arp * d);
                else v = fma(-(arp * d), apc, sp[r * BLU_JLD + c]);
                dp[r * BLU_JLD + c] = v;
            }
        }
        __syncthreads();
        double *tmp = src; src = dst; dst = tmp;
    }
    *out = src;
    return true;
}

// pinv of the sub-block Phi[idx, idx] (idx: ns model ids) into P (ld BLU_JLD, ns x ns).
// Fast path: Gauss-Jordan (the block is SPD and well conditioned in every regular evaluation);
// fallback: Jacobi eigen pseudo-inverse with numpy's cutoff.  Returns the sweeps used (0 = fast path).
__device__ __forceinline__ int blu_block_pinv(const double *phi, int N, const int *idx, int ns, double *A, double *V,
                                              double *P, double *diag0, BluJacobiScratch *js, int tid, int nthr)
{
    for (int t = tid; t < ns * ns; t += nthr) {
        const int r = t / ns, c = t - r * ns;
        A[r * BLU_JLD + c] = phi[idx[r] * N + idx[c]];
    }
    if (tid < ns) diag0[tid] = phi[idx[tid] * N + idx[tid]];
    __syncthreads();
    double *res = nullptr;
    // Measured alternatives, all slower or equal (tools/tail_probe.py; 9.2 us at 20 models / 256 threads, 2.7 us at 10 / 512):
    // a single-warp in-place elimination (20.6 us: the rows of a column are a chain of dependent shared round trips),
    // an element-per-thread mapping (4.6 vs 4.05 us at 15 models), two pivots per barrier with the 2 x 2 block inverse
    // (3.2 vs 2.7 us at 10 models: the step is bounded by the FP64 reciprocal + dependent FMA latency, not by the barrier,
    // and the larger straight-line code is fetched cold -- this CTA runs it once per launch).
    if (blu_block_gj(A, V, diag0, ns, 1.0e-12, &res, tid, nthr)) {
        for (int t = tid; t < ns * ns; t += nthr) {
            const int r = t / ns, c = t - r * ns;
            const int lo = r < c ? r : c, hi = r < c ? c : r;
            P[r * BLU_JLD + c] = res[lo * BLU_JLD + hi];       // exactly symmetric
        }
        __syncthreads();
        return 0;
    }
    __syncthreads();
    const int n2 = ns + (ns & 1);
    for (int t = tid; t < n2 * n2; t += nthr) {
        const int r = t / n2, c = t - r * n2;
        A[r * BLU_JLD + c] = (r < ns && c < ns) ? phi[idx[r] * N + idx[c]] : 0.0;
    }
    __syncthreads();
    blu_sym_pinv(A, V, n2, js, P, BLU_JLD, ns, 1.0e-15, tid, nthr);
    return js->sweeps;
}

#define BLU_FIN_THREADS 512
#define BLU_MAX_PEERS 16
#define BLU_XCHG_DOUBLES 1088                 // N*N + 40 <= 1064 doubles per message
#define BLU_FIN_SEG 8
#define BLU_PHI_GROUP 16                      // CTAs per group of the two-level in-kernel reduction of the partial tiles
#define BLU_PHI_MAXGROUPS 39
#define BLU_PEER_TIMEOUT_NS 4000000000ull     // a rank that never publishes: give up after 4 s instead of spinning forever

// One rank's INBOX for the fused (peer-memory) all-reduce of the partial Phi.  Push model: at evaluation
// number `epoch` every rank stores its partial sums (N*N raw upper-triangle sums + 33 SUM-reducible
// indicators) into slot [epoch & 1][its own rank] of EVERY rank's inbox -- remote stores over NVLink into
// cudaMalloc memory shared through CUDA IPC -- then, after a system-scope fence, the flag
// ready[epoch & 1][its own rank] = epoch.  A rank only ever polls and reads its OWN memory.  Slot parity is
// safe: a rank publishes epoch e+2 only after it has consumed epoch e+1, which every peer published
// after it had finished reading epoch e.
struct BluXchg {
    double data[2][BLU_MAX_PEERS][BLU_XCHG_DOUBLES];
    unsigned long long ready[2][BLU_MAX_PEERS];
};
struct BluPeers {
    BluXchg *peer[BLU_MAX_PEERS];             // peer[r] = rank r's inbox as mapped in THIS process
    int world, rank;
};

// Pre-reduction of the CTA partials for the stand-alone finish kernel (kept for the NCCL path and as the
// reference point of the fused in-kernel reduction): CTA b owns entries [32b, 32b+32) of the N x N tile.
#define BLU_FOLD_WARPS 8
__global__ void __launch_bounds__(BLU_FOLD_WARPS * 32)
blu_phi_fold_kernel(const double *__restrict__ part, int nparts, int NN, double *__restrict__ out)
{
    __shared__ double sh[BLU_FOLD_WARPS][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * 32 + lane;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (e < NN) {
        int p = w;
        for (; p + 3 * BLU_FOLD_WARPS < nparts; p += 4 * BLU_FOLD_WARPS) {
            a0 += part[(long long)p * NN + e];
            a1 += part[(long long)(p + BLU_FOLD_WARPS) * NN + e];
            a2 += part[(long long)(p + 2 * BLU_FOLD_WARPS) * NN + e];
            a3 += part[(long long)(p + 3 * BLU_FOLD_WARPS) * NN + e];
        }
        for (; p < nparts; p += BLU_FOLD_WARPS) a0 += part[(long long)p * NN + e];
    }
    sh[w][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (w == 0 && e < NN) {
        double s = 0.0;
#pragma unroll
        for (int ww = 0; ww < BLU_FOLD_WARPS; ++ww) s += sh[ww][lane];
        out[e] = s;
    }
}

// Shared-memory scratch of the finish step (about 35 KB): carved from static arrays by the stand-alone
// finish kernel and from the (by then idle) streaming ring by the fused Phi kernel.
struct BluFinScratch {
    double *A, *V, *Ph;       // BLU_JMAX x BLU_JLD each; Ph holds the un-mirrored upper-triangle sums on entry
    double *Pm;               // BLU_JMAX x BLU_JMAX: the finished Phi
    double *diag0;            // BLU_JMAX
    BluJacobiScratch *js;
    int *sidx;                // BLU_JMAX
    int *ns;
    unsigned *amask;
};
#define BLU_FIN_SCRATCH_BYTES (sizeof(double) * (3 * BLU_JMAX * BLU_JLD + BLU_JMAX * BLU_JMAX + BLU_JMAX) + sizeof(BluJacobiScratch) + sizeof(int) * (BLU_JMAX + 2) + 64)
__device__ __forceinline__ BluFinScratch blu_fin_carve(unsigned char *raw)
{
    BluFinScratch f;
    double *d = reinterpret_cast<double *>(raw);
    f.A = d; d += BLU_JMAX * BLU_JLD;
    f.V = d; d += BLU_JMAX * BLU_JLD;
    f.Ph = d; d += BLU_JMAX * BLU_JLD;
    f.Pm = d; d += BLU_JMAX * BLU_JMAX;
    f.diag0 = d; d += BLU_JMAX;
    f.js = reinterpret_cast<BluJacobiScratch *>(d);
    int *ip = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(d) + ((sizeof(BluJacobiScratch) + 7) / 8) * 8);
    f.sidx = ip; ip += BLU_JMAX;
    f.ns = ip; ++ip;
    f.amask = reinterpret_cast<unsigned *>(ip);
    return f;
}

// Everything after the local reduction of the partial tiles, for one CTA (any block size):
// f.Ph holds the raw upper-triangle sums of this rank.
//   mode 0: mirror + delta only (get_phi).  mode 1: + pinv, x, S = 2 pinv, variance (misc.py:487-490).
//   mode 2: hand the raw sums + SUM-reducible indicators out in `phi` (partial of a group slice, before an
//           NCCL all-reduce).
//   mode 3: fused multi-GPU path -- push the raw sums into every rank's inbox over NVLink, wait for every
//           rank's message of this epoch, sum the inbox in rank order (bit-identical Phi on every rank),
//           then continue exactly like mode 1.
//   allreduced: the sums in f.Ph and the indicators in phi[NN..NN+32] come from an all-reduce (stand-alone
//           finish kernel after NCCL).
__device__ __forceinline__ void blu_finish_body(int N, double delta, int mode, bool allreduced,
                                                double *__restrict__ phi, double *__restrict__ pinv, double *__restrict__ xrow,
                                                double *__restrict__ S, BluEvalHeader *hdr, const BluPeers &peers,
                                                const BluFinScratch &f, int tid, int nthr)
{
    const int NN = N * N;
    double *Ph = f.Ph, *Pm = f.Pm;
    if (mode == 3) {
        const unsigned long long epoch = hdr->epoch + 1ull;              // same count on every rank
        const int slot = (int)(epoch & 1ull);
        const unsigned sp = hdr->supp;
        const double mx = __longlong_as_double((long long)hdr->maxbits);
        if (tid == 0) *f.ns = 0;
        __syncthreads();
        for (int r = 0; r < peers.world; ++r) {                          // remote stores: fire and forget
            double *dst = peers.peer[r]->data[slot][peers.rank];
            for (int e = tid; e < NN; e += nthr) dst[e] = Ph[(e / N) * BLU_JLD + (e % N)];
            if (tid < 32) dst[NN + tid] = (sp >> tid) & 1u ? 1.0 : 0.0;
            if (tid == 32) dst[NN + 32] = mx >= 0.05 ? 1.0 : 0.0;
        }
        __syncthreads();                                                 // every thread's stores are issued ...
        BluXchg *mine = peers.peer[peers.rank];
        if (tid < peers.world) {
            // ... and ordered before the flag by a system-scope release of the publishing thread (cumulative over the barrier)
            unsigned long long *fl = &peers.peer[tid]->ready[slot][peers.rank];
            asm volatile("fence.acq_rel.sys;\n\tst.relaxed.sys.global.u64 [%0], %1;" ::"l"(fl), "l"(epoch) : "memory");
            // wait for rank tid's message in OUR inbox (local polling), bounded
            const unsigned long long *flag = &mine->ready[slot][tid];
            const unsigned long long t0 = blu_globaltimer();
            unsigned spins = 0;
            for (;;) {
                unsigned long long seen;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
                if (seen >= epoch) break;
                if ((++spins & 1023u) == 0u && blu_globaltimer() - t0 > BLU_PEER_TIMEOUT_NS) { *f.ns = -1; break; }
            }
        }
        if (tid == 0) { hdr->supp = 0u; hdr->maxbits = 0ull; hdr->epoch = epoch; }
        __syncthreads();
        if (*f.ns == -1) {                                               // a peer never arrived: report, do not hang
            if (tid == 0) { hdr->flags = BLU_FLAG_PEER_TIMEOUT; hdr->scal[0] = nan(""); }
            return;
        }
        for (int e = tid; e < NN + 33; e += nthr) {
            double v[BLU_MAX_PEERS];
#pragma unroll
            for (int r = 0; r < BLU_MAX_PEERS; ++r) v[r] = r < peers.world ? __ldcv(&mine->data[slot][r][e]) : 0.0;   // uncached, all in flight
            double sum = 0.0;
#pragma unroll
            for (int r = 0; r < BLU_MAX_PEERS; ++r) sum += v[r];          // rank order: bit-identical Phi on every rank
            if (e < NN) Ph[(e / N) * BLU_JLD + (e % N)] = sum;
            if (e >= NN) phi[e] = sum;                                   // reduced indicators (read back below)
        }
        __threadfence_block();
        __syncthreads();
        allreduced = true;
        mode = 1;
        BLU_STAMP(hdr, 5);                                               // [5] peer exchange complete
    }
    if (mode == 2) {
        for (int e = tid; e < NN; e += nthr) phi[e] = Ph[(e / N) * BLU_JLD + (e % N)];
        const unsigned sp = hdr->supp;
        const double mx = __longlong_as_double((long long)hdr->maxbits);
        __syncthreads();
        if (tid < 32) phi[NN + tid] = (sp >> tid) & 1u ? 1.0 : 0.0;
        if (tid == 32) phi[NN + 32] = mx >= 0.05 ? 1.0 : 0.0;
        if (tid == 0) { hdr->supp = 0u; hdr->maxbits = 0ull; }
        return;
    }
    // (the header cells were written by other SMs' atomics: fetch them now, the L2 round trip overlaps the mirror loop)
    const unsigned supp_hdr = allreduced ? 0u : hdr->supp;
    const unsigned long long maxbits_hdr = allreduced ? 0ull : hdr->maxbits;
    if (tid == 0) *f.amask = 0u;
    __syncthreads();
    // mirror the upper triangle, add delta on the diagonal; rows with any non-zero entry are "active" (see below)
    for (int e = tid; e < NN; e += nthr) {
        const int r = e / N, c = e - r * N;
        const int lo = r < c ? r : c, hi = r < c ? c : r;
        double v = Ph[lo * BLU_JLD + hi];
        if (r == c) v += delta;
        phi[e] = v;
        Pm[e] = v;
        if (v != 0.0) atomicOr(f.amask, 1u << r);             // order-independent: deterministic
    }
    __syncthreads();
    unsigned supp;
    double maxabs;
    if (!allreduced) {
        supp = supp_hdr;
        maxabs = __longlong_as_double((long long)maxbits_hdr);
    } else {                              // SUM-reduced encodings
        supp = 0u;
        for (int a = 0; a < 32; ++a) if (phi[NN + a] > 0.0) supp |= 1u << a;
        maxabs = phi[NN + 32] > 0.0 ? 1.0 : 0.0;
    }
    __syncthreads();
    if (tid == 0) {                      // reset the reduction cells for the next evaluation
        hdr->supp = 0u; hdr->maxbits = 0ull;
        hdr->scal[1] = maxabs;
    }
    if (mode == 0) return;

    unsigned flags = 0u;
    if (maxabs < 0.05) {                 // misc.py:464,484
        if (tid == 0) { hdr->flags = BLU_FLAG_TINY; hdr->scal[0] = INFINITY; }
        return;
    }
    const unsigned all = (N == 32) ? 0xffffffffu : ((1u << N) - 1u);
    if (!(supp & 1u)) flags |= BLU_FLAG_NO_MODEL0;
    if ((supp & all) != all) flags |= BLU_FLAG_PARTIAL;

    // ---- full pseudo-inverse (misc.py:487) ----
    // Rows/columns of Phi that are entirely zero (models in no group with m_i != 0) split off as a
    // zero block: pinv([[A,0],[0,0]]) = [[pinv(A),0],[0,0]].  The remaining "active" block is SPD
    // in every regular evaluation and is inverted by Gauss-Jordan; Jacobi only if that fails.
    {
        const unsigned am = *f.amask;                        // collected by the mirror loop
        if (tid < N && (am >> tid & 1u)) f.sidx[__popc(am & ((1u << tid) - 1u))] = tid;
        if (tid == 0) *f.ns = __popc(am);
    }
    __syncthreads();
    const int na = *f.ns;
    const unsigned active = *f.amask;
    BLU_STAMP(hdr, 6);                                                   // [6] Phi mirrored, support known
    if (tid == 0) f.js->lmax = 0.0;
    int sweeps = blu_block_pinv(Pm, N, f.sidx, na, f.A, f.V, Ph, f.diag0, f.js, tid, nthr);
    BLU_STAMP(hdr, 10);                                                  // [10] active block inverted
    for (int e = tid; e < NN; e += nthr) {
        const int r = e / N, c = e - r * N;
        double v = 0.0;
        if ((active >> r & 1u) && (active >> c & 1u))
            v = Ph[__popc(active & ((1u << r) - 1u)) * BLU_JLD + __popc(active & ((1u << c) - 1u))];
        pinv[e] = v;
        S[e] = 2.0 * v;
        if (r == 0) xrow[c] = v;
    }
    __syncthreads();
    BLU_STAMP(hdr, 7);                                                   // [7] pseudo-inverse written
    if (tid == 0) { hdr->scal[2] = (double)sweeps; hdr->scal[3] = f.js->lmax; hdr->scal[4] = pinv[0]; }

    // ---- variance on the support sub-block (misc.py:489-490) ----
    const unsigned sup = supp & all;
    if (sup == (active & all)) {
        // pinv(Phi[idx,idx])[0,0] is the (first supported model) diagonal entry of the block above
        const int s0 = sup ? __ffs(sup) - 1 : 0;
        if (tid == 0) { hdr->scal[0] = sup ? pinv[s0 * N + s0] : INFINITY; hdr->flags = flags; }
        if (tid < 32) hdr->xsup[tid] = (tid < N && (sup >> tid & 1u)) ? pinv[s0 * N + tid] : 0.0;
        return;
    }
    __syncthreads();
    if (tid == 0) {
        int cnt = 0;
        for (int a = 0; a < N; ++a) if (sup >> a & 1u) f.sidx[cnt++] = a;
        *f.ns = cnt;
    }
    __syncthreads();
    const int nsub = *f.ns;
    if (nsub > 0) blu_block_pinv(Pm, N, f.sidx, nsub, f.A, f.V, Ph, f.diag0, f.js, tid, nthr);
    if (tid == 0) { hdr->scal[0] = nsub > 0 ? Ph[0] : INFINITY; hdr->flags = flags; }
    if (tid < 32) hdr->xsup[tid] = 0.0;
    __syncthreads();
    if (tid < nsub) hdr->xsup[f.sidx[tid]] = Ph[tid];
}

// Stand-alone finish kernel (one CTA): fixed-order sum of `nparts` partial tiles (or, nparts == 0, the
// all-reduced sums already sitting in `phi`), then blu_finish_body.
__global__ void __launch_bounds__(BLU_FIN_THREADS)
blu_phi_finish_kernel(int N, int nparts, const double *__restrict__ part, double delta, int mode,
                      double *__restrict__ phi, double *__restrict__ pinv, double *__restrict__ xrow,
                      double *__restrict__ S, BluEvalHeader *hdr, BluPeers peers)
{
    __shared__ __align__(16) unsigned char fraw[BLU_FIN_SCRATCH_BYTES];
    extern __shared__ double red[];              // BLU_FIN_SEG x N*N staging for the partial sums
    const BluFinScratch f = blu_fin_carve(fraw);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int NN = N * N;
    if (tid == 0) *f.ns = 0;
    if (nparts > 0) {
        for (int t = tid; t < NN * BLU_FIN_SEG; t += nthr) {
            const int seg = t / NN, e = t - seg * NN;
            double s = 0.0;
            int p = seg;
            for (; p + 3 * BLU_FIN_SEG < nparts; p += 4 * BLU_FIN_SEG) {       // 4 loads in flight, same association every run
                const double a0 = part[(long long)p * NN + e], a1 = part[(long long)(p + BLU_FIN_SEG) * NN + e];
                const double a2 = part[(long long)(p + 2 * BLU_FIN_SEG) * NN + e], a3 = part[(long long)(p + 3 * BLU_FIN_SEG) * NN + e];
                s += a0; s += a1; s += a2; s += a3;
            }
            for (; p < nparts; p += BLU_FIN_SEG) s += part[(long long)p * NN + e];
            red[seg * NN + e] = s;
        }
        __syncthreads();
        for (int e = tid; e < NN; e += nthr) {
            double s = 0.0;
#pragma unroll
            for (int seg = 0; seg < BLU_FIN_SEG; ++seg) s += red[seg * NN + e];
            f.Ph[(e / N) * BLU_JLD + (e % N)] = s;
        }
    } else {
        for (int e = tid; e < NN; e += nthr) f.Ph[(e / N) * BLU_JLD + (e % N)] = phi[e];
    }
    __syncthreads();
    blu_finish_body(N, delta, mode, nparts == 0, phi, pinv, xrow, S, hdr, peers, f, tid, nthr);
}

// `wchunks` lists the work of this launch (whole context or the owned slice) warp by warp: warp w of the grid walks
// wchunks[wstart[w] .. wstart[w+1]).  ns: ring stages (2 or 4).  Dynamic shared memory: blu_phi_smem_bytes.
// lab switch: compile for BLU_PHI_LB_C CTAs of 8 warps per SM.  Measured with 3 (80 registers, 24 warps/SM with 2 KB chunks):
// 314 us at 20 models against 142 us -- the register cap spills the group loop (profiles/r02_phi_lab.md)
#ifdef BLU_PHI_LB_C
#define BLU_PHI_LB 256, BLU_PHI_LB_C
#else
#define BLU_PHI_LB BLU_PHI_WARPS * 32
#endif
__global__ void __launch_bounds__(BLU_PHI_LB)
blu_phi_partial_kernel(const BluClass *__restrict__ cls, int ncls, int N, const BluChunk *__restrict__ wchunks, const int *__restrict__ wstart,
                       int ns, int sd, int idsd, const double *__restrict__ cinv, const unsigned char *__restrict__ gidx,
                       const unsigned short *__restrict__ lut, const unsigned *__restrict__ plut,
                       const unsigned *__restrict__ gmask, const double *__restrict__ m,
                       double *__restrict__ part, BluEvalHeader *hdr,
                       int fin_mode, double delta, double *__restrict__ phi, double *__restrict__ pinv, double *__restrict__ xrow,
                       double *__restrict__ S, BluPeers peers)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const int NN = N * N;
    const int nwarps = blockDim.x >> 5;                  // 16 normally, 8 when N is large (shared-memory budget)
    // shared memory: [value stages warps x ns x sd doubles][tiles warps x N*N][class table][mbarriers warps x 4]
    //                [records warps x 32 x 16 B][id stages warps x ns x idsd bytes]      (blu_phi_smem_bytes)
    // The run table (plut) and, for the generic fallback of very large groups, the classic (j,l) table are read from
    // global memory: shared memory decides how many CTAs fit on an SM.
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *stages = reinterpret_cast<double *>(smraw);
    double *tiles = stages + (size_t)nwarps * ns * sd;
    BluClass *scls = reinterpret_cast<BluClass *>(tiles + (size_t)nwarps * NN);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(scls + ncls);
    unsigned char *recs = reinterpret_cast<unsigned char *>(bars + nwarps * BLU_PHI_MAXSTAGES);
    recs += (16 - (blu_smem_u32(recs) & 15u)) & 15u;
    unsigned char *idst = recs + (size_t)nwarps * 512;
    for (int t = threadIdx.x; t < ncls; t += blockDim.x) scls[t] = cls[t];
    BluPhiStream ps;
    ps.sdb = 8u * (unsigned)sd; ps.idsd = (unsigned)idsd;
    ps.val_s = blu_smem_u32(stages + (size_t)w * ns * sd);
    ps.ids_s = blu_smem_u32(idst + (size_t)w * ns * idsd);
    ps.bar_s = blu_smem_u32(bars + w * BLU_PHI_MAXSTAGES);
    ps.rec_s = blu_smem_u32(recs + (size_t)w * 512);
    if (lane == 0) {
        for (int q = 0; q < ns; ++q) blu_mbar_init(bars + w * BLU_PHI_MAXSTAGES + q, 1);
        blu_mbar_fence_init();
    }
    double *acc = tiles + w * NN;                        // this warp's private N x N tile
    for (int t = lane; t < NN; t += 32) acc[t] = 0.0;
    __syncthreads();

    const int gw = blockIdx.x * nwarps + w;
    unsigned wsupp = 0u;
    double mymax = 0.0;
#ifdef BLU_PHI_PROFILE
    const long long prof_t0 = clock64();
#endif
    if (ns == 4) blu_phi_walk<4>(wstart[gw], wstart[gw + 1], wchunks, scls, cinv, gidx, lut, plut, m, gmask, ps, acc, N, lane, wsupp, mymax, hdr);
    else         blu_phi_walk<2>(wstart[gw], wstart[gw + 1], wchunks, scls, cinv, gidx, lut, plut, m, gmask, ps, acc, N, lane, wsupp, mymax, hdr);
#ifdef BLU_PHI_PROFILE
    if (lane == 0) {
        atomicAdd(&hdr->stamp[11], (unsigned long long)(clock64() - prof_t0));
        atomicAdd(&hdr->stamp[12], 1ull);
        atomicMax(&hdr->stamp[8], (unsigned long long)(clock64() - prof_t0));
        if (gw < 8192) blu_prof_warp[gw][0] = clock64() - prof_t0;
    }
#endif
    unsigned supp = __reduce_or_sync(BLU_FULL, wsupp);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mymax = fmax(mymax, __shfl_xor_sync(BLU_FULL, mymax, o));
    if (lane == 0) {
        if (supp) atomicOr(&hdr->supp, supp);
        atomicMax(&hdr->maxbits, (unsigned long long)__double_as_longlong(mymax));
    }
    __syncthreads();
    for (int t = threadIdx.x; t < NN; t += blockDim.x) {
        double sum = 0.0;
        for (int ww = 0; ww < nwarps; ++ww) sum += tiles[ww * NN + t];
        part[(long long)blockIdx.x * NN + t] = sum;
    }
    if (fin_mode < 0) return;                             // partial tiles only (stand-alone finish kernel follows)

    // ---- fused reduction + finish: the LAST CTA to arrive folds the partial tiles and carries on ----------
    // Two levels so that no CTA reads more than BLU_PHI_GROUP + BLU_PHI_MAXGROUPS tiles: the last CTA of each
    // group of BLU_PHI_GROUP consecutive CTAs folds the group's tiles (CTA order), the last group to finish folds
    // the group sums (group order) -- a fixed association whichever CTA happens to do it, so the result is
    // bit-reproducible.  Then the same CTA does the finish step (peer exchange, pinv, variance): one launch
    // instead of partial -> fold -> finish.
    __shared__ int s_last;
    __shared__ unsigned long long s_t0;
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) s_t0 = blu_globaltimer();
    const int grp = blockIdx.x / BLU_PHI_GROUP;
    const int ngrp = (gridDim.x + BLU_PHI_GROUP - 1) / BLU_PHI_GROUP;
    const int members = min(BLU_PHI_GROUP, (int)gridDim.x - grp * BLU_PHI_GROUP);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(&hdr->tickets[1 + grp], 1u);
        s_last = (t == (unsigned)(members - 1));
        if (s_last) hdr->tickets[1 + grp] = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned long long t_g0 = blu_globaltimer();
    double *part2 = part + (size_t)gridDim.x * NN;        // group sums
    if (ngrp == 1) {
        // small problems (<= BLU_PHI_GROUP CTAs): the group sum IS the total -- no second round trip through global memory
        if (tid == 0) { hdr->stamp[1] = s_t0; hdr->stamp[2] = t_g0; }
        BLU_STAMP(hdr, 3);
        const BluFinScratch f = blu_fin_carve(smraw);     // the ring and the accumulator tiles are idle now
        for (int e = tid; e < NN; e += nthr) {
            const int r = e / N, cc0 = e - r * N;
            double v[BLU_PHI_GROUP];
#pragma unroll
            for (int cc = 0; cc < BLU_PHI_GROUP; ++cc) v[cc] = (r <= cc0 && cc < members) ? __ldcg(part + (size_t)cc * NN + e) : 0.0;
            double sum = 0.0;
#pragma unroll
            for (int cc = 0; cc < BLU_PHI_GROUP; ++cc) sum += v[cc];    // CTA order
            f.Ph[r * BLU_JLD + cc0] = sum;
        }
        __syncthreads();
        BLU_STAMP(hdr, 4);
        blu_finish_body(N, delta, fin_mode, false, phi, pinv, xrow, S, hdr, peers, f, tid, nthr);
        BLU_STAMP(hdr, 9);
        return;
    }
    for (int e = tid; e < NN; e += nthr) {
        const double *pp = part + (size_t)grp * BLU_PHI_GROUP * NN + e;
        double v[BLU_PHI_GROUP];
        const bool upper = (e / N) <= (e % N);                // the tiles hold the upper triangle only
#pragma unroll
        for (int cc = 0; cc < BLU_PHI_GROUP; ++cc) v[cc] = (upper && cc < members) ? __ldcg(pp + (size_t)cc * NN) : 0.0;   // all in flight at once
        double sum = 0.0;
#pragma unroll
        for (int cc = 0; cc < BLU_PHI_GROUP; ++cc) sum += v[cc];    // CTA order
        part2[(size_t)grp * NN + e] = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(&hdr->tickets[0], 1u);
        s_last = (t == (unsigned)(ngrp - 1));
        if (s_last) hdr->tickets[0] = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) { hdr->stamp[1] = s_t0; hdr->stamp[2] = t_g0; }   // [1] this CTA's stream done, [2] it was the last of its group
    BLU_STAMP(hdr, 3);                                    // [3] last group: final fold starts
    const BluFinScratch f = blu_fin_carve(smraw);         // the ring and the accumulator tiles are idle now
    for (int e = tid; e < NN; e += nthr) {
        const int r = e / N, cc0 = e - r * N;
        double sum = 0.0;
        if (r <= cc0) {
            for (int g0 = 0; g0 < ngrp; g0 += 20) {             // 20 loads in flight (one batch up to 320 CTAs), group order
                double v[20];
#pragma unroll
                for (int u = 0; u < 20; ++u) v[u] = (g0 + u < ngrp) ? __ldcg(part2 + (size_t)(g0 + u) * NN + e) : 0.0;
#pragma unroll
                for (int u = 0; u < 20; ++u) sum += v[u];
            }
        }
        f.Ph[r * BLU_JLD + cc0] = sum;
    }
    __syncthreads();
    BLU_STAMP(hdr, 4);                                    // [4] sums of this rank complete
    blu_finish_body(N, delta, fin_mode, false, phi, pinv, xrow, S, hdr, peers, f, tid, nthr);
    BLU_STAMP(hdr, 9);                                    // [9] finish step done
}

