// blu_soa_types.h -- host-visible constants and the tile descriptor of the lane-per-group kernels
// (blu_soa.cuh), shared by blu_capi.cu (work-list construction) and blu_soa_tu.cu (the kernels).
#pragma once

#define BLU_SOA_WARPS 8
#define BLU_SOA_E 16                              // packed entries per stage
#define BLU_SOA_STAGE (BLU_SOA_E * 32)            // doubles per stage

struct BluTile {
    int cls;          // class index
    int nsub;         // ceil(T / BLU_SOA_E)
    long long t;      // tile index inside the class (groups 32 t .. 32 t + 31)
};
