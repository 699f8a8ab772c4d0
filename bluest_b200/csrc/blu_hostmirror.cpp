// blu_hostmirror.cpp -- host half of the symmetric Hessian download (blu_capi.cu,
// download_hessian_symmetric): only the upper block-triangle of the exactly symmetric (L,L) matrix
// crosses PCIe; this routine writes the lower triangle from it while later panels are in flight.
//
// Destination tile = 8 rows x (r1-r0) doubles: the source is (r1-r0) rows x 8 doubles (one 64-byte
// line per source row), transposed through a 64 KB L2-resident buffer and written as 8 long runs of
// non-temporal stores, so the destination lines are never read (a cached store costs a
// read-for-ownership: 3 x 4.3 GB of host traffic instead of 2 x).  Measured on the B200 box's host
// (16 cores): 53 ms for the 4.3 GB lower triangle of L = 32767 (74 ms without the software prefetch of the
// next tile's lines: source rows are 262 KB apart, the hardware prefetcher never sees a stream),
// against 130-145 ms for 64 x 64 cached tiles and 48 ms for a plain 4.3 GB memcpy (tools/lab/mirror_bench.cpp).
#include <emmintrin.h>
#include <xmmintrin.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include "blu_launch.h"

#define MIRROR_CW 8
#define MIRROR_MAXR 1024

static inline void stream_run(double *dst, const double *src, int n)
{
    int i = 0;
    if (i < n && ((uintptr_t)dst & 15)) { _mm_stream_si64((long long *)dst, *(const long long *)src); i = 1; }
    for (; i + 2 <= n; i += 2) _mm_stream_pd(dst + i, _mm_loadu_pd(src + i));
    if (i < n) _mm_stream_si64((long long *)(dst + i), *(const long long *)(src + i));
}

double *blu_host_mirror_scratch_alloc() { return (double *)aligned_alloc(64, sizeof(double) * MIRROR_CW * MIRROR_MAXR); }
void blu_host_mirror_scratch_free(double *buf) { free(buf); }

// buf: the calling worker's transpose scratch (blu_host_mirror_scratch_alloc); NULL = no scratch could be
// had: plain cached transpose, slower but correct.
void blu_host_mirror_block(double *H, long long L, long long r0, long long r1, long long c0, long long c1, double *buf)
{
    if (!buf) {
        for (long long r = r0; r < r1; ++r)
            for (long long c = c0; c < c1; ++c) H[c * L + r] = H[r * L + c];
        return;
    }
    for (long long rb = r0; rb < r1; rb += MIRROR_MAXR) {
        const int nr = (int)std::min<long long>(MIRROR_MAXR, r1 - rb);
        for (long long cb = c0; cb < c1; cb += MIRROR_CW) {
            const int nc = (int)std::min<long long>(MIRROR_CW, c1 - cb);
            for (int i = 0; i < nr; ++i) {
                const double *src = H + (rb + i) * L + cb;
                _mm_prefetch((const char *)(src + 2 * MIRROR_CW), _MM_HINT_T1);     // the one new line the next column tile needs
                for (int j = 0; j < nc; ++j) buf[j * MIRROR_MAXR + i] = src[j];
            }
            for (int j = 0; j < nc; ++j) stream_run(H + (cb + j) * L + rb, buf + j * MIRROR_MAXR, nr);
        }
    }
}

void blu_host_store_fence() { _mm_sfence(); }
