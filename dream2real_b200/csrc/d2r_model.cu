// Model / view set-up: everything that happens once per snapshot or once per (view, resolution).
// Compiled with -ftz=true -prec-div=false -prec-sqrt=false (the arithmetic part of the reference's
// --use_fast_math, instant-ngp CMakeLists.txt:80-82) so divisions lower to the same approximate code.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "d2r_common.cuh"

namespace d2r {

thread_local std::string g_last_error;
thread_local unsigned long long g_launch_count = 0;
void set_error(const std::string& msg) { g_last_error = msg; }

// ---- occupancy bitfield: reference src/testbed_nerf.cu:284-331 + 2355-2373 ----------------------
__host__ __device__ inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
__host__ __device__ inline uint32_t morton3D_invert(uint32_t x) {
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}

// mean over cascade 0 only of max(v,0)/n  (reduce_sum over NERF_GRID_N_CELLS elements)
__global__ void k_grid_mean(const float* __restrict__ grid, uint32_t n, double* __restrict__ out) {
    double acc = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        acc += (double)(fmaxf(grid[i], 0.f) / (float)n);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

__global__ void k_grid_to_bitfield(uint32_t n_elements, uint32_t n_nonzero, const float* __restrict__ grid,
                                   uint8_t* __restrict__ bits, const double* __restrict__ mean_ptr) {
    const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= n_elements) return;
    if (i >= n_nonzero) { bits[i] = 0; return; }
    const float thresh = fminf(0.01f, (float)*mean_ptr);   // NERF_MIN_OPTICAL_THICKNESS
    uint8_t b = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) b |= grid[i * 8 + j] > thresh ? (uint8_t)(1u << j) : 0;
    bits[i] = b;
}

// Morton-ordered cascade bitfields -> x + 128*y + 128^2*z order (one thread per output byte = 8 cells along x)
__global__ void k_bitfield_linearise(uint32_t n_bytes_total, const uint8_t* __restrict__ morton, uint8_t* __restrict__ lin) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bytes_total) return;
    const uint32_t casc = i / (NERF_GRID_N_CELLS / 8), b = i % (NERF_GRID_N_CELLS / 8);
    const uint32_t cell0 = b * 8, x0 = cell0 & 127u, y = (cell0 >> 7) & 127u, z = cell0 >> 14;
    auto expand = [](uint32_t v) {
        v = (v * 0x00010001u) & 0xFF0000FFu;
        v = (v * 0x00000101u) & 0x0F00F00Fu;
        v = (v * 0x00000011u) & 0xC30C30C3u;
        v = (v * 0x00000005u) & 0x49249249u;
        return v;
    };
    const uint8_t* src = morton + (size_t)casc * (NERF_GRID_N_CELLS / 8);
    uint32_t out = 0;
    for (uint32_t j = 0; j < 8; ++j) {
        const uint32_t m = expand(x0 + j) | (expand(y) << 1) | (expand(z) << 2);
        out |= (uint32_t)((src[m >> 3] >> (m & 7)) & 1u) << j;
    }
    lin[i] = (uint8_t)out;
}

__global__ void k_bitfield_max_pool(uint32_t n_elements, const uint8_t* __restrict__ prev, uint8_t* __restrict__ next) {
    const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= n_elements) return;
    uint8_t b = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) b |= prev[i * 8 + j] > 0 ? (uint8_t)(1u << j) : 0;
    const uint32_t x = morton3D_invert(i >> 0) + NERF_GRIDSIZE / 8;
    const uint32_t y = morton3D_invert(i >> 1) + NERF_GRIDSIZE / 8;
    const uint32_t z = morton3D_invert(i >> 2) + NERF_GRIDSIZE / 8;
    next[morton3D(x, y, z)] |= b;
}

// grid_scale / grid_resolution evaluated on the device like kernel_grid does
// (tiny-cuda-nn common_device.h:856-865, called from grid.h:104-105)
__global__ void k_level_geometry(int n_levels, float log2_pls, uint32_t base_res, float* scale_out, uint32_t* res_out) {
    const int l = threadIdx.x;
    if (l >= n_levels) return;
    const float scale = exp2f(l * log2_pls) * base_res - 1.0f;   // ex2.approx under -ftz (fast-math exp2f)
    scale_out[l] = scale;
    res_out[l] = (uint32_t)ceilf(scale) + 1;
}

// ---- per-pixel camera-plane directions: uv_to_ray before the rotation ---------------------------
// reference include/neural-graphics-primitives/common_device.cuh:249-262, 289-333, 393-431
__device__ inline void opencv_delta(const float* p, float u, float v, float* du, float* dv) {
    const float k1 = p[0], k2 = p[1], p1 = p[2], p2 = p[3];
    const float u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2;
    const float radial = k1 * r2 + k2 * r2 * r2;
    *du = u * radial + 2.f * p1 * uv + p2 * (r2 + 2.f * u2);
    *dv = v * radial + 2.f * p2 * uv + p1 * (r2 + 2.f * v2);
}

__global__ void k_view_dirs(int W, int H, float fx, float fy, float scx, float scy, int lens_mode,
                            float k1, float k2, float p1, float p2, float2* __restrict__ out) {
    const int x = threadIdx.x + blockDim.x * blockIdx.x;
    const int y = threadIdx.y + blockDim.y * blockIdx.y;
    if (x >= W || y >= H) return;
    // ld_random_pixel_offset(0) == (0.5, 0.5) exactly (random_val.cuh:320-325)
    const float uvx = ((float)x + 0.5f) / (float)W;
    const float uvy = ((float)y + 0.5f) / (float)H;
    float dx = (uvx - scx) * (float)W / fx;
    float dy = (uvy - scy) * (float)H / fy;
    if (lens_mode == 1) {
        const float params[4] = {k1, k2, p1, p2};
        const float x0 = dx, y0 = dy;
        float xu = dx, xv = dy;
        for (uint32_t it = 0; it < 100; ++it) {
            const float step0 = fmaxf(1.1920929e-07f, fabsf(1e-6f * xu));
            const float step1 = fmaxf(1.1920929e-07f, fabsf(1e-6f * xv));
            float dxu, dxv, b0u, b0v, f0u, f0v, b1u, b1v, f1u, f1v;
            opencv_delta(params, xu, xv, &dxu, &dxv);
            opencv_delta(params, xu - step0, xv, &b0u, &b0v);
            opencv_delta(params, xu + step0, xv, &f0u, &f0v);
            opencv_delta(params, xu, xv - step1, &b1u, &b1v);
            opencv_delta(params, xu, xv + step1, &f1u, &f1v);
            const float J00 = 1 + (f0u - b0u) / (2 * step0);
            const float J10 = (f1u - b1u) / (2 * step1);
            const float J01 = (f0v - b0v) / (2 * step0);
            const float J11 = 1 + (f1v - b1v) / (2 * step1);
            const float ru = xu + dxu - x0, rv = xv + dxv - y0;
            const float inv_det = 1.0f / (J00 * J11 - J10 * J01);      // tcnn inverse(mat2)
            const float su = (J11 * ru - J10 * rv) * inv_det;
            const float sv = (-J01 * ru + J00 * rv) * inv_det;
            xu -= su;
            xv -= sv;
            if (su * su + sv * sv < 1e-10f) break;
        }
        dx = xu;
        dy = xv;
    }
    out[x + W * y] = make_float2(dx, dy);
}

// per-column min/max of the direction table's x and per-row min/max of its y: with them a candidate's conservative screen
// rectangle follows from the projection of the occupied box, lens distortion included (launch_march plans it on the host)
__global__ void k_view_ranges(int W, int H, const float2* __restrict__ dirs, float* col_lo, float* col_hi, float* row_lo, float* row_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < W) {
        float lo = 1e30f, hi = -1e30f;
        for (int y = 0; y < H; ++y) { const float v = dirs[i + W * y].x; lo = fminf(lo, v); hi = fmaxf(hi, v); }
        col_lo[i] = lo; col_hi[i] = hi;
    }
    if (i < H) {
        float lo = 1e30f, hi = -1e30f;
        for (int x = 0; x < W; ++x) { const float v = dirs[x + W * i].y; lo = fminf(lo, v); hi = fmaxf(hi, v); }
        row_lo[i] = lo; row_hi[i] = hi;
    }
}

}  // namespace d2r

using namespace d2r;

extern "C" const char* d2r_last_error(void) { return g_last_error.c_str(); }
extern "C" unsigned long long d2r_launch_count(int reset) {
    unsigned long long v = g_launch_count;
    if (reset) g_launch_count = 0;
    return v;
}
extern "C" const char* d2r_version(void) { return "d2r_b200 0.1 (sm_100a)"; }

static int model_load_impl(const void* params_f16_host, size_t n_params, const float* density_grid_f32_host,
                           size_t n_grid_cells, const d2r_model_cfg* cfg, int device, d2r_model* m) {
    D2R_REQUIRE(cfg->n_levels == 8 && cfg->n_features_per_level == 4,
                "d2r_model_load: only the Dream2Real NGP config (8 levels x 4 features) is supported");
    D2R_REQUIRE(cfg->max_cascade >= 0 && cfg->max_cascade < (int)NERF_CASCADES, "d2r_model_load: bad max_cascade");
    D2R_REQUIRE(n_grid_cells == (size_t)NERF_GRID_N_CELLS * (cfg->max_cascade + 1),
                "Incompatible number of grid cascades.");   // testbed.cu:4811-4814

    // hash-grid offsets: GridEncodingTemplated ctor (tiny-cuda-nn grid.h:693-722), host math
    uint32_t offsets[MAX_LEVELS + 1];
    {
        const float log2_pls = std::log2(cfg->per_level_scale);
        uint32_t offset = 0;
        for (int i = 0; i < cfg->n_levels; ++i) {
            const float scale = exp2f(i * log2_pls) * cfg->base_resolution - 1.0f;
            const uint32_t res = (uint32_t)ceilf(scale) + 1;
            const uint32_t max_params = 0xFFFFFFFFu / 2;
            uint32_t params_in_level = std::pow((float)res, 3) > (float)max_params ? max_params : res * res * res;
            params_in_level = (params_in_level + 7u) / 8u * 8u;
            params_in_level = std::min(params_in_level, 1u << cfg->log2_hashmap_size);
            offsets[i] = offset;
            offset += params_in_level;
        }
        offsets[cfg->n_levels] = offset;
    }
    const size_t n_mlp = 64 * 32 + 16 * 64 + 64 * 32 + 64 * 64 + 16 * 64;
    const size_t n_grid = (size_t)offsets[cfg->n_levels] * N_FEAT;
    if (n_params != n_mlp + n_grid) {
        set_error("d2r_model_load: n_params " + std::to_string(n_params) + " != expected " + std::to_string(n_mlp + n_grid));
        return D2R_ERR_INVALID;
    }

    m->device = device;
    m->cfg = *cfg;
    m->n_params = n_params;
    D2R_CUDA(cudaMalloc(&m->params_dev, n_params * sizeof(__half)));
    D2R_CUDA(cudaMemcpy(m->params_dev, params_f16_host, n_params * sizeof(__half), cudaMemcpyHostToDevice));
    const size_t bf_bytes = (size_t)NERF_GRID_N_CELLS / 8 * NERF_CASCADES;
    D2R_CUDA(cudaMalloc(&m->bitfield_dev, bf_bytes));
    D2R_CUDA(cudaMalloc(&m->bitfield_lin_dev, bf_bytes));

    // occupancy bitfield
    DevBuf grid_buf, mean_buf, scale_buf, res_buf;
    D2R_CUDA(cudaMalloc(&grid_buf.p, n_grid_cells * sizeof(float)));
    D2R_CUDA(cudaMalloc(&mean_buf.p, sizeof(double)));
    float* grid_dev = (float*)grid_buf.p;
    double* mean_dev = (double*)mean_buf.p;
    D2R_CUDA(cudaMemcpy(grid_dev, density_grid_f32_host, n_grid_cells * sizeof(float), cudaMemcpyHostToDevice));
    D2R_CUDA(cudaMemset(mean_dev, 0, sizeof(double)));
    const uint32_t n = NERF_GRID_N_CELLS;
    k_grid_mean<<<256, 256>>>(grid_dev, n, mean_dev);
    k_grid_to_bitfield<<<(n / 8 * NERF_CASCADES + 255) / 256, 256>>>(n / 8 * NERF_CASCADES, n / 8 * (cfg->max_cascade + 1),
                                                                   grid_dev, m->bitfield_dev, mean_dev);
    for (uint32_t level = 1; level < NERF_CASCADES; ++level)
        k_bitfield_max_pool<<<(n / 64 + 255) / 256, 256>>>(n / 64, m->bitfield_dev + (size_t)(level - 1) * (n / 8),
                                                         m->bitfield_dev + (size_t)level * (n / 8));
    k_bitfield_linearise<<<(unsigned)((bf_bytes + 255) / 256), 256>>>((uint32_t)bf_bytes, m->bitfield_dev, m->bitfield_lin_dev);
    count_launch(3 + NERF_CASCADES - 1);
    D2R_CUDA(cudaGetLastError());

    // level geometry on device
    D2R_CUDA(cudaMalloc(&scale_buf.p, MAX_LEVELS * sizeof(float)));
    D2R_CUDA(cudaMalloc(&res_buf.p, MAX_LEVELS * sizeof(uint32_t)));
    float* scale_dev = (float*)scale_buf.p;
    uint32_t* res_dev = (uint32_t*)res_buf.p;
    k_level_geometry<<<1, 32>>>(cfg->n_levels, std::log2(cfg->per_level_scale), (uint32_t)cfg->base_resolution, scale_dev, res_dev);
    count_launch();
    ModelDev& d = m->dev;
    D2R_CUDA(cudaMemcpy(d.level_scale, scale_dev, MAX_LEVELS * sizeof(float), cudaMemcpyDeviceToHost));
    D2R_CUDA(cudaMemcpy(d.level_res, res_dev, MAX_LEVELS * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    memcpy(d.level_offset, offsets, sizeof(offsets));
    for (int l = 0; l < cfg->n_levels; ++l) {
        // the reference's stride loop (tiny-cuda-nn common_device.h:697-713) decides dense vs hashed per level
        const uint32_t size = offsets[l + 1] - offsets[l], res = d.level_res[l];
        uint32_t stride = 1;
        for (int dim = 0; dim < 3 && stride <= size; ++dim) stride *= res;
        d.level_hashed[l] = size < stride ? 1u : 0u;
        if (d.level_hashed[l] && (size & (size - 1)) != 0) {
            set_error("d2r_model_load: hashed level with a non power-of-two table");
            return D2R_ERR_INVALID;
        }
        if (!d.level_hashed[l] && (uint64_t)res * res * res + (uint64_t)res * res + res >= 2ull * size) {
            set_error("d2r_model_load: dense level index range exceeds 2x the table size");
            return D2R_ERR_INVALID;
        }
    }

    {   // MLP weights re-arranged once as UMMA B operands (canonical K-major, no swizzle: 8x8 core matrices of 128 contiguous
        // bytes, core matrices adjacent in K 128 B apart, 8-row groups K/8*128 B apart), in the order the kernels' W_* offsets expect
        const __half* hp = (const __half*)params_f16_host;
        std::vector<__half> blob(20480 / 2);
        const int shp[5][2] = {{64, 32}, {16, 64}, {64, 32}, {64, 64}, {16, 64}};
        size_t src = 0, dst = 0;
        for (int w = 0; w < 5; ++w) {
            const int N = shp[w][0], K = shp[w][1];
            for (int n = 0; n < N; ++n)
                for (int k = 0; k < K; ++k) {
                    const size_t off = (size_t)(n >> 3) * (K / 8) * 128 + (size_t)(k >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(k & 7) * 2;
                    blob[dst + off / 2] = hp[src + (size_t)n * K + k];
                }
            src += (size_t)N * K;
            dst += (size_t)N * K;
        }
        D2R_CUDA(cudaMalloc(&m->w_umma_dev, 20480));
        D2R_CUDA(cudaMemcpy(m->w_umma_dev, blob.data(), 20480, cudaMemcpyHostToDevice));
        d.w_umma = (const unsigned char*)m->w_umma_dev;
    }
    const __half* p = (const __half*)m->params_dev;
    d.w_d0 = p; p += 64 * 32;
    d.w_d1 = p; p += 16 * 64;
    d.w_c0 = p; p += 64 * 32;
    d.w_c1 = p; p += 64 * 64;
    d.w_c2 = p; p += 16 * 64;
    d.grid = p;
    for (int l = 0; l < MAX_LEVELS; ++l) {
        const bool used = l < cfg->n_levels;
        d.level_table[l] = reinterpret_cast<const uint2*>(d.grid) + (used ? offsets[l] : 0);
        d.level_size[l] = used ? offsets[l + 1] - offsets[l] : 1u;
    }
    d.bitfield = m->bitfield_dev;
    d.bitfield_lin = m->bitfield_lin_dev;
    bool ident = true;
    for (int i = 0; i < 3; ++i) {
        d.aabb_min[i] = cfg->aabb_min[i];
        d.aabb_diag[i] = cfg->aabb_max[i] - cfg->aabb_min[i];
        d.raabb_min[i] = cfg->render_aabb_min[i];
        d.raabb_max[i] = cfg->render_aabb_max[i];
    }
    for (int i = 0; i < 9; ++i) {
        d.r2l[i] = cfg->render_aabb_to_local[i];
        ident &= d.r2l[i] == ((i % 4 == 0) ? 1.f : 0.f);
    }
    d.r2l_identity = ident ? 1 : 0;
    d.max_cascade = cfg->max_cascade;
    d.cone = cfg->cone_angle_constant;
    d.min_transmittance = cfg->min_transmittance;
    d.depth_scale = cfg->depth_scale;

    // tight box of occupied cells (host pass over the bitfield of cascades 0..max_cascade)
    {
        std::vector<uint8_t> bits((size_t)n / 8 * (cfg->max_cascade + 1));
        D2R_CUDA(cudaMemcpy(bits.data(), m->bitfield_dev, bits.size(), cudaMemcpyDeviceToHost));
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (int c = 0; c <= cfg->max_cascade; ++c) {
            uint32_t cmin[3] = {128, 128, 128}, cmax[3] = {0, 0, 0};
            bool any = false;
            const uint8_t* b = bits.data() + (size_t)c * (n / 8);
            for (uint32_t byte = 0; byte < n / 8; ++byte) {
                if (!b[byte]) continue;
                for (int j = 0; j < 8; ++j) {
                    if (!(b[byte] & (1u << j))) continue;
                    const uint32_t idx = byte * 8 + j;
                    const uint32_t xyz[3] = {morton3D_invert(idx), morton3D_invert(idx >> 1), morton3D_invert(idx >> 2)};
                    for (int a = 0; a < 3; ++a) { cmin[a] = std::min(cmin[a], xyz[a]); cmax[a] = std::max(cmax[a], xyz[a]); }
                    any = true;
                }
            }
            if (!any) continue;
            const float size = ldexpf(1.0f, c);           // cascade c spans [0.5 - size/2, 0.5 + size/2]
            for (int a = 0; a < 3; ++a) {
                lo[a] = std::min(lo[a], 0.5f - 0.5f * size + size * (float)cmin[a] / 128.f);
                hi[a] = std::max(hi[a], 0.5f - 0.5f * size + size * (float)(cmax[a] + 1) / 128.f);
            }
        }
        for (int a = 0; a < 3; ++a) {
            const float pad = 1e-3f;   // conservative against float rounding in the box test
            d.occ_min[a] = lo[a] - pad;
            d.occ_max[a] = hi[a] + pad;
        }
        // bounding sphere of the occupied cells around the box centre (second pass: farthest cell corner); blob-shaped
        // objects leave the corners of their box empty, and a ray is done once it is past box AND sphere
        double r2 = 0.0;
        const double ctr[3] = {0.5 * ((double)lo[0] + hi[0]), 0.5 * ((double)lo[1] + hi[1]), 0.5 * ((double)lo[2] + hi[2])};
        for (int c = 0; c <= cfg->max_cascade; ++c) {
            const uint8_t* b = bits.data() + (size_t)c * (n / 8);
            const double size = ldexp(1.0, c), cell = size / 128.0;
            for (uint32_t byte = 0; byte < n / 8; ++byte) {
                if (!b[byte]) continue;
                for (int j = 0; j < 8; ++j) {
                    if (!(b[byte] & (1u << j))) continue;
                    const uint32_t idx = byte * 8 + j;
                    const uint32_t xyz[3] = {morton3D_invert(idx), morton3D_invert(idx >> 1), morton3D_invert(idx >> 2)};
                    double dd = 0.0;
                    for (int a = 0; a < 3; ++a) {
                        const double c0 = 0.5 - 0.5 * size + cell * xyz[a], c1 = c0 + cell;
                        const double far = std::max(std::fabs(c0 - ctr[a]), std::fabs(c1 - ctr[a]));
                        dd += far * far;
                    }
                    r2 = std::max(r2, dd);
                }
            }
        }
        const double rad = std::sqrt(r2) * (1.0 + 1e-3) + 1e-3;     // conservative, like the box pad
        for (int a = 0; a < 3; ++a) d.occ_ctr[a] = (float)ctr[a];
        d.occ_r2 = lo[0] <= hi[0] ? (float)(rad * rad) : -1.0f;     // nothing occupied: no sphere
    }
    D2R_CUDA(cudaDeviceSynchronize());
    return D2R_OK;
}

extern "C" int d2r_model_load(const void* params_f16_host, size_t n_params, const float* density_grid_f32_host,
                              size_t n_grid_cells, const d2r_model_cfg* cfg, int device, d2r_model** out) {
    D2R_REQUIRE(params_f16_host && density_grid_f32_host && cfg && out, "d2r_model_load: null argument");
    DeviceGuard dg(device);
    d2r_model* m = new d2r_model();
    memset(m, 0, sizeof(*m));
    m->device = device;
    const int rc = model_load_impl(params_f16_host, n_params, density_grid_f32_host, n_grid_cells, cfg, device, m);
    if (rc != D2R_OK) { d2r_model_free(m); return rc; }
    *out = m;
    return D2R_OK;
}

extern "C" void d2r_model_free(d2r_model* m) {
    if (!m) return;
    DeviceGuard dg(m->device);
    cudaFree(m->params_dev);
    cudaFree(m->bitfield_dev);
    cudaFree(m->bitfield_lin_dev);
    cudaFree(m->w_umma_dev);
    delete m;
}

extern "C" int d2r_model_set_min_transmittance(d2r_model* m, float v) {
    D2R_REQUIRE(m, "d2r_model_set_min_transmittance: null model");
    D2R_REQUIRE(v >= 0.f && v < 1.f, "d2r_model_set_min_transmittance: value must be in [0,1)");
    m->dev.min_transmittance = v;
    m->cfg.min_transmittance = v;
    return D2R_OK;
}

extern "C" int d2r_model_get_bitfield(const d2r_model* m, uint8_t* out, size_t n_bytes) {
    D2R_REQUIRE(m && out, "d2r_model_get_bitfield: null argument");
    D2R_REQUIRE(n_bytes == (size_t)NERF_GRID_N_CELLS / 8 * NERF_CASCADES, "d2r_model_get_bitfield: size must be 8*128^3/8");
    DeviceGuard dg(m->device);
    D2R_CUDA(cudaMemcpy(out, m->bitfield_dev, n_bytes, cudaMemcpyDeviceToHost));
    return D2R_OK;
}

extern "C" int d2r_model_get_occupied_aabb(const d2r_model* m, float* out6) {
    D2R_REQUIRE(m && out6, "d2r_model_get_occupied_aabb: null argument");
    for (int a = 0; a < 3; ++a) { out6[a] = m->dev.occ_min[a]; out6[3 + a] = m->dev.occ_max[a]; }
    return D2R_OK;
}

extern "C" int d2r_view_prepare(const d2r_camera* cam, int device, d2r_view** out) {
    D2R_REQUIRE(cam && out, "d2r_view_prepare: null argument");
    D2R_REQUIRE(cam->width > 0 && cam->height > 0 && cam->width <= 16384 && cam->height <= 16384, "d2r_view_prepare: bad resolution");
    D2R_REQUIRE(cam->lens_mode == 0 || cam->lens_mode == 1, "d2r_view_prepare: only perspective and OpenCV lenses are on this path");
    DeviceGuard dg(device);
    d2r_view* v = new d2r_view();
    memset(v, 0, sizeof(*v));
    v->device = device;
    v->W = cam->width;
    v->H = cam->height;
    v->cam = *cam;
    if (cudaMalloc(&v->dirs_dev, (size_t)v->W * v->H * sizeof(float2)) != cudaSuccess) { delete v; set_error("d2r_view_prepare: cudaMalloc failed"); return D2R_ERR_NOMEM; }
    dim3 threads(16, 8), blocks((v->W + 15) / 16, (v->H + 7) / 8);
    k_view_dirs<<<blocks, threads>>>(v->W, v->H, cam->focal[0], cam->focal[1], cam->screen_center[0], cam->screen_center[1],
                                     cam->lens_mode, cam->lens_params[0], cam->lens_params[1], cam->lens_params[2],
                                     cam->lens_params[3], v->dirs_dev);
    count_launch();
    {
        const int W = v->W, H = v->H;
        DevBuf rb;
        if (cudaMalloc(&rb.p, (size_t)2 * (W + H) * sizeof(float)) != cudaSuccess) { d2r_view_free(v); set_error("d2r_view_prepare: cudaMalloc failed"); return D2R_ERR_NOMEM; }
        float* r = (float*)rb.p;
        k_view_ranges<<<(std::max(W, H) + 127) / 128, 128>>>(W, H, v->dirs_dev, r, r + W, r + 2 * W, r + 2 * W + H);
        count_launch();
        v->ranges_host = new float[(size_t)2 * (W + H)];
        const cudaError_t e = cudaMemcpy(v->ranges_host, r, (size_t)2 * (W + H) * sizeof(float), cudaMemcpyDeviceToHost);   // also the sync of this set-up call
        if (e != cudaSuccess) { d2r_view_free(v); set_error(std::string("d2r_view_prepare: ") + cudaGetErrorString(e)); return D2R_ERR_CUDA; }
    }
    *out = v;
    return D2R_OK;
}

extern "C" void d2r_view_free(d2r_view* v) {
    if (!v) return;
    DeviceGuard dg(v->device);
    cudaFree(v->dirs_dev);
    delete[] v->ranges_host;
    delete v;
}

extern "C" int d2r_view_get_dirs(const d2r_view* v, float* out) {
    D2R_REQUIRE(v && out, "d2r_view_get_dirs: null argument");
    DeviceGuard dg(v->device);
    D2R_CUDA(cudaMemcpy(out, v->dirs_dev, (size_t)v->W * v->H * sizeof(float2), cudaMemcpyDeviceToHost));
    return D2R_OK;
}
