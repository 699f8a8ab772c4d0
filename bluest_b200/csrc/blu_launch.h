// blu_launch.h -- host entry points shared between the translation units of libbluest_b200.so
// (internal; the public boundary is include/bluest_b200.h).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// blu_invert_tu.cu
void blu_launch_invert_class(int k, int nsm, cudaStream_t stream, const double *d_C, int N, const uint8_t *gidx, long long Lk,
                             double *cinv, unsigned char *flag, double pivtol);
void blu_launch_pinv_groups(unsigned ngroups, cudaStream_t stream, const double *d_C, int N, int k, const uint8_t *gidx,
                            const long long *d_todo, double *cinv, double rcond);
void blu_launch_pack_invcovs(int grid, cudaStream_t stream, const double *d_full, int k, long long Lk, double *cinv);
void blu_launch_unpack_invcovs(int grid, cudaStream_t stream, const double *cinv, int k, long long Lk, double *d_full);

// blu_soa_tu.cu
struct BluClass;
struct BluTile;
cudaError_t blu_launch_grad_soa(bool with_u, int grid, cudaStream_t stream, const BluClass *cls, int ncls, int N, int NP, int K,
                                const BluTile *tiles, int ntiles, const double *soa, const long long *soff, const unsigned *gmask,
                                const double *xrow, long long lo, long long hi, double *grad, double *U);
cudaError_t blu_launch_soa_build(int grid, cudaStream_t stream, const double *cinv, long long Lk, int T, double *soa);
cudaError_t blu_launch_v_from_u(int grid, cudaStream_t stream, const double *U, const double *S, int N, int NP, long long lo, long long hi, double *V);

// blu_hostmirror.cpp: H[c][r] = H[r][c] for r in [r0,r1), c in [c0,c1) of a dense row-major (L,L) host
// matrix, written with streaming stores (no read-for-ownership of the destination lines).
void blu_host_mirror_block(double *H, long long L, long long r0, long long r1, long long c0, long long c1, double *scratch);
double *blu_host_mirror_scratch_alloc();          // one per worker thread, freed by the worker (NULL on failure: slow path)
void blu_host_mirror_scratch_free(double *scratch);
void blu_host_store_fence();
