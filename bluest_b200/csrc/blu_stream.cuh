// blu_stream.cuh -- the warp-private streaming pipeline shared by the per-group kernels
// (Phi assembly, gradient, U/V factors).
//
// The packed inverses of consecutive groups of one size class are contiguous in HBM, so the unit
// of work is a CHUNK: G consecutive groups = one contiguous span of G*T_k doubles (<= ~4 KB).
// Chunks are dealt round-robin to warps.  Each warp owns a two-stage ring in shared memory and
// moves its chunks with the bulk asynchronous copy engine (cp.async.bulk global->shared, the
// 1-D TMA path; SASS: UBLKCP) completing on a per-stage mbarrier: one elected lane issues the
// copy of chunk c+1 while all 32 lanes consume chunk c from shared memory.  The small per-group
// operands (m_i, membership mask) ride along in registers, one group per lane, loaded coalesced
// at prefetch time.  Nothing in the inner loop waits on a global load.
//
// Alignment: a group block starts on an 8-byte boundary only (T_k may be odd); the copy starts at
// the 16-byte boundary below and the consumer skips `skew` (0/1) doubles.
#pragma once
#include "blu_common.cuh"

#define BLU_CHUNK_DOUBLES 544                       // largest payload of one stage (>= T_32 = 528)
// The stage size is a per-context run-time value `sd` (doubles) = chunk payload + 4 (skew + 16-byte
// round-up): contexts whose largest group block is small get small stages, hence more CTAs per SM.

struct BluChunk {
    int cls;          // index into the class table
    int g;            // groups in this chunk (1..32)
    long long i0;     // first group, class-relative
};

__device__ __forceinline__ unsigned blu_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void blu_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(blu_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void blu_mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void blu_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(blu_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void blu_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(blu_smem_u32(dst)), "l"(src), "r"(bytes), "r"(blu_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void blu_mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BLU_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BLU_DONE;\n"
        "bra BLU_WAIT;\n"
        "BLU_DONE:\n"
        "}\n" ::"r"(blu_smem_u32(bar)), "r"(parity) : "memory");
}

// Per-warp view of the pipeline state.
struct BluWarpStream {
    double *stage[2];
    unsigned long long *bar[2];
};

// What a lane carries for the chunk being consumed / prefetched.
struct BluChunkRegs {
    int cls, g, skew;
    long long i0;
    double m;          // m of group (i0 + lane), 0 beyond g
    unsigned mask;     // membership mask of group (i0 + lane)
};

// Issue the copy of chunk `ch` into stage s and load its per-group operands.  All lanes call.
__device__ __forceinline__ BluChunkRegs blu_prefetch_chunk(const BluChunk ch, const BluClass *__restrict__ scls,
                                                           const double *__restrict__ cinv, const double *__restrict__ m,
                                                           const unsigned *__restrict__ gmask, const BluWarpStream &ws, int s, int lane)
{
    BluChunkRegs r;
    const BluClass ci = scls[ch.cls];
    const double *src = cinv + ci.coff + ch.i0 * ci.T;
    const unsigned long long addr = (unsigned long long)src;
    r.skew = (int)((addr & 15ull) >> 3);
    r.cls = ch.cls; r.g = ch.g; r.i0 = ch.i0;
    if (lane == 0) {
        const unsigned bytes = (unsigned)(((r.skew + ch.g * ci.T) * 8 + 15) & ~15);
        blu_mbar_expect_tx(ws.bar[s], bytes);
        blu_bulk_g2s(ws.stage[s], (const void *)(addr & ~15ull), bytes, ws.bar[s]);
    }
    const long long gi = ci.goff + ch.i0 + lane;
    const bool ok = lane < ch.g;
    r.m = (ok && m) ? m[gi] : 0.0;
    r.mask = ok ? gmask[gi] : 0u;
    return r;
}

// Shared-memory carve-up common to the streaming kernels:
//   [stages  WARPS x 2 x sd doubles][extra doubles (kernel specific)]
//   [class table][(j,l) LUT u16][mbarriers WARPS x 2][member-id scratch WARPS x 32 groups x 32 bytes]
#define BLU_STREAM_WARPS 8
// member-id scratch of one warp: 32 groups x BLU_IDS_LD bytes.  Row pitch 33 (not 32): lane g writes byte j of ITS row
// when the masks are expanded, and with a pitch of 32 the 32 lanes hit 4 banks (8-way conflict per store).
#define BLU_IDS_LD 33
#define BLU_IDS_BYTES 1088                          // 32 * 33 rounded up to 64

struct BluStreamSmem {
    double *stages;
    double *extra;
    BluClass *cls;
    unsigned short *lut;
    unsigned long long *bars;
    unsigned char *ids;
};

__host__ __device__ __forceinline__ size_t blu_stream_smem_bytes(int sd, int extra_doubles, int ncls, int lutlen, int warps = BLU_STREAM_WARPS)
{
    size_t b = sizeof(double) * ((size_t)warps * 2 * sd + extra_doubles);
    b += sizeof(BluClass) * ncls;
    b += ((sizeof(unsigned short) * lutlen + 7) / 8) * 8;
    b += sizeof(unsigned long long) * warps * 2;
    b += (size_t)warps * BLU_IDS_BYTES;
    return b;
}

__device__ __forceinline__ BluStreamSmem blu_stream_carve(unsigned char *raw, int sd, int extra_doubles, int ncls, int lutlen, int warps = BLU_STREAM_WARPS)
{
    BluStreamSmem s;
    s.stages = reinterpret_cast<double *>(raw);
    s.extra = s.stages + (size_t)warps * 2 * sd;
    s.cls = reinterpret_cast<BluClass *>(s.extra + extra_doubles);
    s.lut = reinterpret_cast<unsigned short *>(s.cls + ncls);
    s.bars = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(s.lut) + ((sizeof(unsigned short) * lutlen + 7) / 8) * 8);
    s.ids = reinterpret_cast<unsigned char *>(s.bars + warps * 2);
    return s;
}

// Block-wide prologue: class table + LUT into shared memory, barriers initialised.
__device__ __forceinline__ BluWarpStream blu_stream_begin(const BluStreamSmem &s, int sd, const BluClass *__restrict__ cls, int ncls,
                                                          const unsigned short *__restrict__ lut, int lutlen)
{
    for (int t = threadIdx.x; t < ncls; t += blockDim.x) s.cls[t] = cls[t];
    for (int t = threadIdx.x; t < lutlen; t += blockDim.x) s.lut[t] = lut[t];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    BluWarpStream ws;
    ws.stage[0] = s.stages + (size_t)(2 * w) * sd;
    ws.stage[1] = s.stages + (size_t)(2 * w + 1) * sd;
    ws.bar[0] = s.bars + 2 * w;
    ws.bar[1] = s.bars + 2 * w + 1;
    if (lane == 0) { blu_mbar_init(ws.bar[0], 1); blu_mbar_init(ws.bar[1], 1); blu_mbar_fence_init(); }
    __syncthreads();
    return ws;
}

// Member ids of every group of a chunk: lane g expands its own mask into ids[g*32 + j] = j-th set bit.
// All lanes call (k is warp-uniform); followed by __syncwarp.  Amortises the expansion over the chunk.
__device__ __forceinline__ void blu_expand_ids(unsigned mask, int k, unsigned char *ids, int lane)
{
    unsigned mk = mask;
    for (int j = 0; j < k; ++j) {
        const int b = __ffs(mk) - 1;
        ids[lane * BLU_IDS_LD + j] = (unsigned char)(b < 0 ? 0 : b);
        mk &= mk - 1u;
    }
    __syncwarp();
}
