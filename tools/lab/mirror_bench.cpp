// host-side mirror microbenchmark: H[c][r] = H[r][c] for the strict upper block triangle
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>
#include <sys/mman.h>
static void mirror_a(double *H, long long L, long long r0, long long r1, long long c0, long long c1)
{
    alignas(64) double buf[64][64];
    for (long long rb = r0; rb < r1; rb += 64) {
        const int nr = (int)std::min<long long>(64, r1 - rb);
        for (long long cb = c0; cb < c1; cb += 64) {
            const int nc = (int)std::min<long long>(64, c1 - cb);
            for (int i = 0; i < nr; ++i) {
                const double *src = H + (rb + i) * L + cb;
                for (int j = 0; j < nc; ++j) buf[j][i] = src[j];
            }
            for (int j = 0; j < nc; ++j) memcpy(H + (cb + j) * L + rb, buf[j], sizeof(double) * nr);
        }
    }
}
static inline void stream_row(double *dst, const double *src, int n)
{
    int i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { _mm_stream_si64((long long *)(dst + i), *(const long long *)(src + i)); ++i; }
    for (; i + 4 <= n; i += 4) _mm256_stream_pd(dst + i, _mm256_loadu_pd(src + i));
    for (; i < n; ++i) _mm_stream_si64((long long *)(dst + i), *(const long long *)(src + i));
}
template <int CW>
static void mirror_b(double *H, long long L, long long r0, long long r1, long long c0, long long c1)
{
    // dest tile: CW rows x (r1-r0) doubles; source: (r1-r0) rows x CW doubles
    static thread_local double *buf = nullptr;
    if (!buf) buf = (double *)aligned_alloc(64, sizeof(double) * CW * 1024);
    const int nr = (int)(r1 - r0);
    for (long long cb = c0; cb < c1; cb += CW) {
        const int nc = (int)std::min<long long>(CW, c1 - cb);
        for (int i = 0; i < nr; ++i) {
            const double *src = H + (r0 + i) * L + cb;
            for (int j = 0; j < nc; ++j) buf[j * 1024 + i] = src[j];
        }
        for (int j = 0; j < nc; ++j) stream_row(H + (cb + j) * L + r0, buf + j * 1024, nr);
    }
}
static int g_cw = 8, g_r = 1024, g_pf = 0;
static void mirror_c(double *H, long long L, long long r0, long long r1, long long c0, long long c1)
{
    static thread_local double *buf = nullptr;
    if (!buf) buf = (double *)aligned_alloc(64, sizeof(double) * 64 * 1024);
    const int CW = g_cw, R = g_r;
    for (long long rb = r0; rb < r1; rb += R) {
        const int nr = (int)std::min<long long>(R, r1 - rb);
        for (long long cb = c0; cb < c1; cb += CW) {
            const int nc = (int)std::min<long long>(CW, c1 - cb);
            for (int i = 0; i < nr; ++i) {
                const double *src = H + (rb + i) * L + cb;
                if (g_pf) _mm_prefetch((const char *)(src + CW + 8), _MM_HINT_T1);
                for (int j = 0; j < nc; ++j) buf[j * R + i] = src[j];
            }
            for (int j = 0; j < nc; ++j) stream_row(H + (cb + j) * L + rb, buf + j * R, nr);
        }
    }
}
typedef void (*mfn)(double *, long long, long long, long long, long long, long long);
int main(int argc, char **argv)
{
    long long L = argc > 1 ? atoll(argv[1]) : 32767;
    int nthreads = argc > 2 ? atoi(argv[2]) : 16;
    double *H = (double *)mmap(nullptr, sizeof(double) * L * L, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (H == MAP_FAILED) { perror("mmap"); return 1; }
    {   // parallel first touch
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back([=]() { for (long long r = t; r < L; r += nthreads) for (long long c = 0; c < L; ++c) H[r * L + c] = (double)(r * 3 + c); });
        for (auto &t : pool) t.join();
    }
    const long long PH = 1024, CW = 512;
    struct V { const char *name; mfn f; int cw, r, pf; } var[] = {{"tile64 memcpy", mirror_a, 0, 0, 0}, {"nt 8x1024", mirror_c, 8, 1024, 0}, {"nt 8x1024 pf", mirror_c, 8, 1024, 1},
        {"nt 8x512", mirror_c, 8, 512, 0}, {"nt 8x256", mirror_c, 8, 256, 0}, {"nt 16x512", mirror_c, 16, 512, 0}, {"nt 16x256", mirror_c, 16, 256, 0}, {"nt 16x256 pf", mirror_c, 16, 256, 1},
        {"nt 32x256", mirror_c, 32, 256, 0}, {"nt 32x128", mirror_c, 32, 128, 0}, {"nt 64x128", mirror_c, 64, 128, 0}, {"nt 64x64", mirror_c, 64, 64, 0}};
    for (auto &v : var) {
        g_cw = v.cw; g_r = v.r; g_pf = v.pf;
        for (int rep = 0; rep < 2; ++rep) {
            std::vector<std::pair<long long, long long>> items;
            const int npan = (int)((L + PH - 1) / PH);
            for (int p = 0; p < npan; ++p) { long long r1 = std::min(L, (p + 1) * PH); for (long long c0 = r1; c0 < L; c0 += CW) items.push_back({p, c0}); }
            std::atomic<size_t> next{0};
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads; ++t) pool.emplace_back([&]() {
                for (;;) { size_t i = next.fetch_add(1); if (i >= items.size()) break;
                    long long r0 = items[i].first * PH, r1 = std::min(L, r0 + PH), c0 = items[i].second, c1 = std::min(L, c0 + CW);
                    v.f(H, L, r0, r1, c0, c1); }
                _mm_sfence(); });
            for (auto &t : pool) t.join();
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("%-16s threads %d: %.1f ms  (%.1f GB/s of mirrored payload)\n", v.name, nthreads, s * 1e3, 4.0 * L * L / s / 1e9);
        }
    }
    // verify
    long long bad = 0;
    for (long long r = 0; r < L; r += 97) for (long long c = r + 1025; c < L; c += 89) if (H[c * L + r] != H[r * L + c]) ++bad;
    printf("bad %lld\n", bad);
    // plain memcpy bandwidth reference: copy top half to bottom half
    {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool; const long long half = L * L / 2;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back([=]() { long long n = half / nthreads; memcpy(H + half + t * n, H + t * n, sizeof(double) * n); });
        for (auto &t : pool) t.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("memcpy %.1f GB in %.1f ms = %.1f GB/s payload\n", 8.0 * half / 1e9, s * 1e3, 8.0 * half / s / 1e9);
    }
    return 0;
}
