// Shared host/device helpers for libd2r_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/d2r_b200.h"

namespace d2r {

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const std::string& msg);
extern thread_local unsigned long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += (unsigned long long)n; }

#define D2R_CUDA(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::d2r::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +    \
                             __FILE__ + ":" + std::to_string(__LINE__) + ")");                     \
            return D2R_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define D2R_REQUIRE(cond, msg)                                                                     \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            ::d2r::set_error(std::string(msg));                                                    \
            return D2R_ERR_INVALID;                                                                \
        }                                                                                          \
    } while (0)

struct DevBuf {        // temporary device allocation, released on every exit path
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};
struct DeviceGuard {   // the library never leaves the caller's current device changed
    int prev = -1;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) cudaSetDevice(device);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- constants: reference include/neural-graphics-primitives/nerf_device.cuh:23-42 -------------
constexpr uint32_t NERF_GRIDSIZE = 128;
constexpr uint32_t NERF_GRID_N_CELLS = 128u * 128u * 128u;
constexpr uint32_t NERF_CASCADES = 8;
constexpr int MAX_LEVELS = 8;
constexpr int N_FEAT = 4;

struct Mat3x4 {     // NGP camera: columns 0..2 = rotation columns, column 3 = origin
    float c[4][3];
};

struct ModelDev {
    // hash grid
    const __half* grid;                // [n_entries, 4]
    uint32_t level_offset[MAX_LEVELS + 1];
    float level_scale[MAX_LEVELS];
    uint32_t level_res[MAX_LEVELS];
    uint32_t level_hashed[MAX_LEVELS];   // 1: coherent-prime hash into a 2^k table, 0: dense (wrap-around) index
    const uint2* level_table[MAX_LEVELS];   // grid + level_offset[l] as 8-byte entries: one IMAD.WIDE per corner address
    uint32_t level_size[MAX_LEVELS];     // entries of the level; hashed levels: a power of two
    // MLP weights, fp16 row-major [out,in]
    const __half* w_d0;  // [64,32]
    const __half* w_d1;  // [16,64]
    const __half* w_c0;  // [64,32]
    const __half* w_c1;  // [64,64]
    const __half* w_c2;  // [16,64]
    const unsigned char* w_umma;       // the five matrices as UMMA B operands (K-major, no swizzle), 20480 bytes: what the march kernels stage
    const uint8_t* bitfield;           // [8 * 128^3 / 8], Morton order inside a cascade (the reference's layout; exported)
    const uint8_t* bitfield_lin;       // same bits, cell (x,y,z) at bit x + 128*y + 128^2*z: what the march reads (no Morton encode)
    float aabb_min[3], aabb_diag[3];
    float raabb_min[3], raabb_max[3];
    float r2l[9];
    int r2l_identity;
    int max_cascade;
    float cone;
    float min_transmittance;
    float depth_scale;
    float occ_min[3], occ_max[3];      // tight box of occupied cells
    float occ_ctr[3], occ_r2;          // bounding sphere of the occupied cells (centre = box centre), radius^2 (< 0: none)
};

}  // namespace d2r

struct d2r_model {
    int device;
    d2r::ModelDev dev;
    void* params_dev;     // fp16 blob
    uint8_t* bitfield_dev;
    uint8_t* bitfield_lin_dev;
    void* w_umma_dev;
    size_t n_params;
    d2r_model_cfg cfg;
};

struct d2r_view {
    int device;
    int W, H;
    float2* dirs_dev;     // [H*W] undistorted camera-plane (x, y), z == 1
    d2r_camera cam;
    // per-column min/max of dirs.x and per-row min/max of dirs.y, [col_lo W | col_hi W | row_lo H | row_hi H]: the host plans
    // every candidate's conservative screen rectangle from them (launch_march), so no launch needs a read-back
    float* ranges_host;
};
