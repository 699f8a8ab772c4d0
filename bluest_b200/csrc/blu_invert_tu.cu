// blu_invert_tu.cu -- separate translation unit for kernel (1), the 32 register-resident
// instantiations of blu_invert_groups_kernel<K> (they dominate the compile time of the library),
// plus the pack / unpack / pinv-fallback kernels of blu_invert.cuh.  Host launch wrappers only;
// blu_capi.cu owns the context and the error handling.
#include <algorithm>
#include "blu_invert.cuh"
#include "blu_launch.h"

template <int K>
static void launch_invert(int nsm, cudaStream_t stream, const double *d_C, int N, const uint8_t *gidx, long long Lk,
                          double *cinv, unsigned char *flag, double pivtol)
{
    constexpr int G = 32 / BluSub<K>::value;
    const long long warps = (Lk + G - 1) / G;
    const int grid = (int)std::max<long long>(1, std::min<long long>((warps + 3) / 4, (long long)nsm * 16));
    blu_invert_groups_kernel<K><<<grid, 128, 0, stream>>>(d_C, N, gidx, Lk, cinv, flag, pivtol);
}

void blu_launch_invert_class(int k, int nsm, cudaStream_t stream, const double *d_C, int N, const uint8_t *gidx, long long Lk,
                             double *cinv, unsigned char *flag, double pivtol)
{
    switch (k) {
#define CASE(K) case K: launch_invert<K>(nsm, stream, d_C, N, gidx, Lk, cinv, flag, pivtol); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
        CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
        CASE(17) CASE(18) CASE(19) CASE(20) CASE(21) CASE(22) CASE(23) CASE(24)
        CASE(25) CASE(26) CASE(27) CASE(28) CASE(29) CASE(30) CASE(31) CASE(32)
#undef CASE
    }
}

void blu_launch_pinv_groups(unsigned ngroups, cudaStream_t stream, const double *d_C, int N, int k, const uint8_t *gidx,
                            const long long *d_todo, double *cinv, double rcond)
{
    blu_pinv_groups_kernel<<<ngroups, 256, 0, stream>>>(d_C, N, k, gidx, d_todo, cinv, rcond);
}

void blu_launch_pack_invcovs(int grid, cudaStream_t stream, const double *d_full, int k, long long Lk, double *cinv)
{
    blu_pack_invcovs_kernel<<<grid, 256, 0, stream>>>(d_full, k, Lk, cinv);
}

void blu_launch_unpack_invcovs(int grid, cudaStream_t stream, const double *cinv, int k, long long Lk, double *d_full)
{
    blu_unpack_invcovs_kernel<<<grid, 256, 0, stream>>>(cinv, k, Lk, d_full);
}
