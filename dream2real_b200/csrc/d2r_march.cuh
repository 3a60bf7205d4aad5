// Device maths of the Instant-NGP ray march, restated for sm_100a.
// Every function cites the reference code whose arithmetic (operation order, fp32/fp16 rounding
// points) it follows; "NGP" = /root/reference/reconstruction/instant-ngp, "TCNN" = its
// dependencies/tiny-cuda-nn.  The reference is built with --use_fast_math; this file is built with
// -ftz=true -prec-div=false -prec-sqrt=false and spells the fast intrinsics (__logf/__expf/__powf)
// out, so the maths lowers to the same approximate instructions (MUFU.LG2/EX2/RCP/RSQ).
#pragma once
#include "d2r_common.cuh"

namespace d2r {

__device__ __forceinline__ float SQRT3() { return 1.73205080757f; }
__device__ __forceinline__ float STEPSIZE() { return SQRT3() / 1024.0f; }                         // NGP nerf_device.cuh:30-31
__device__ __forceinline__ float MIN_CONE_STEPSIZE() { return STEPSIZE(); }
__device__ __forceinline__ float MAX_CONE_STEPSIZE() { return STEPSIZE() * (1 << 7) * 1024.0f / 128.0f; }  // :35
__device__ __forceinline__ float MAX_DEPTH() { return 16384.0f; }                                  // common_device.cuh:32

// ---- Sobol start jitter: NGP random_val.cuh:215-291 (bit-exact integer code) ---------------------
__device__ __forceinline__ uint32_t laine_karras_permutation(uint32_t x, uint32_t seed) {
    x += seed;
    x ^= x * 0x6c50b47cu;
    x ^= x * 0xb82f1e52u;
    x ^= x * 0xc7afe638u;
    x ^= x * 0x8d22f6e6u;
    return x;
}
__device__ __forceinline__ uint32_t nested_uniform_scramble_base2(uint32_t x, uint32_t seed) {
    x = __brev(x);
    x = laine_karras_permutation(x, seed);
    return __brev(x);
}
__device__ __forceinline__ float ld_random_val0(uint32_t seed) {
    // ld_random_val(index = 0, seed, dim = 0); sobol(.,0) is a bit reversal
    const float S = float(1.0 / (1ull << 32));
    const uint32_t index = nested_uniform_scramble_base2(0u, seed);
    const uint32_t hc = seed ^ (0u + (seed << 6) + (seed >> 2));     // hash_combine(seed, 0)
    return (float)nested_uniform_scramble_base2(__brev(index), hc) * S;
}

// ---- stepping space: NGP nerf_device.cuh:378-428 -------------------------------------------------
// The reference recomputes log1p_c, a, b, at, bt inside every call; they depend on the cone angle
// only, so they are evaluated once per thread here -- with the same instruction sequence, hence the
// same values -- and carried in StepC.
struct StepC {
    float cone, log1p_c, a, b, at, bt;
};
__device__ __forceinline__ StepC make_stepc(float cone_angle) {
    StepC c;
    c.cone = cone_angle;
    c.log1p_c = __logf(1.0f + cone_angle);
    c.a = (__logf(MIN_CONE_STEPSIZE()) - __logf(c.log1p_c)) / c.log1p_c;
    c.b = (__logf(MAX_CONE_STEPSIZE()) - __logf(c.log1p_c)) / c.log1p_c;
    c.at = __expf(c.a * c.log1p_c);
    c.bt = __expf(c.b * c.log1p_c);
    return c;
}
__device__ __forceinline__ float to_stepping_space(float t, const StepC& c) {
    if (c.cone <= 1e-5f) return t / MIN_CONE_STEPSIZE();
    if (t <= c.at) return (t - c.at) / MIN_CONE_STEPSIZE() + c.a;
    else if (t <= c.bt) return __logf(t) / c.log1p_c;
    else return (t - c.bt) / MAX_CONE_STEPSIZE() + c.b;
}
__device__ __forceinline__ float from_stepping_space(float n, const StepC& c) {
    if (c.cone <= 1e-5f) return n * MIN_CONE_STEPSIZE();
    if (n <= c.a) return (n - c.a) * MIN_CONE_STEPSIZE() + c.at;
    else if (n <= c.b) return __expf(n * c.log1p_c);
    else return (n - c.b) * MAX_CONE_STEPSIZE() + c.bt;
}
__device__ __forceinline__ float advance_n_steps(float t, const StepC& c, float n) {
    return from_stepping_space(to_stepping_space(t, c) + n, c);
}
__device__ __forceinline__ float calc_dt(float t, const StepC& c) { return advance_n_steps(t, c, 1.0f) - t; }

__device__ __forceinline__ float warp_dt(float dt) {      // nerf_device.cuh:306-309
    float max_stepsize = MIN_CONE_STEPSIZE() * (1 << 7);
    return (dt - MIN_CONE_STEPSIZE()) / (max_stepsize - MIN_CONE_STEPSIZE());
}
__device__ __forceinline__ float unwarp_dt(float dt) {    // nerf_device.cuh:311-314
    float max_stepsize = MIN_CONE_STEPSIZE() * (1 << 7);
    return dt * (max_stepsize - MIN_CONE_STEPSIZE()) + MIN_CONE_STEPSIZE();
}

// ---- occupancy grid: NGP nerf_device.cuh:316-340, 359-367, 430-447, 462-494 ----------------------
__device__ __forceinline__ uint32_t expand_bits_d(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton3D_d(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits_d(x) | (expand_bits_d(y) << 1) | (expand_bits_d(z) << 2);
}
// 2^-mip / 128 * 2^-mip as exact powers of two (scalbnf(1.0f, -mip), scalbnf(128.0f, -mip) without the library call)
__device__ __forceinline__ float pow2_neg(uint32_t mip) { return __int_as_float((int)((127u - mip) << 23)); }
__device__ __forceinline__ float res_of_mip(uint32_t mip) { return __int_as_float((int)((134u - mip) << 23)); }
// cascaded_grid_idx_at (nerf_device.cuh:430-447): same cell arithmetic; the index is x + 128*y + 128^2*z into the
// linearised copy of the bitfield instead of the Morton code (the bit it selects is the same cell's)
__device__ __forceinline__ uint32_t cascaded_grid_idx_at(float px, float py, float pz, uint32_t mip) {
    float mip_scale = pow2_neg(mip);
    px -= 0.5f; py -= 0.5f; pz -= 0.5f;
    px *= mip_scale; py *= mip_scale; pz *= mip_scale;
    px += 0.5f; py += 0.5f; pz += 0.5f;
    int ix = (int)(px * 128.0f), iy = (int)(py * 128.0f), iz = (int)(pz * 128.0f);
    if ((uint32_t)(ix | iy | iz) >= 128u) return 0xFFFFFFFFu;      // any coordinate negative or >= 128
    return (uint32_t)ix | ((uint32_t)iy << 7) | ((uint32_t)iz << 14);
}
__device__ __forceinline__ bool density_grid_occupied_at(float px, float py, float pz, const uint8_t* __restrict__ bits_lin, uint32_t mip) {
    uint32_t idx = cascaded_grid_idx_at(px, py, pz, mip);
    if (idx == 0xFFFFFFFFu) return false;
    return __ldg(bits_lin + idx / 8 + (NERF_GRID_N_CELLS * mip) / 8) & (1 << (idx % 8));
}
__device__ __forceinline__ uint32_t mip_from_pos(float px, float py, float pz) {
    // frexpf(maxval, &exponent): exponent field - 126 for normal numbers, 0 for zero (denormals are flushed: -ftz)
    float maxval = fmaxf(fmaxf(fabsf(px - 0.5f), fabsf(py - 0.5f)), fabsf(pz - 0.5f));
    const int e = (int)((__float_as_uint(maxval) >> 23) & 0xffu);
    const int exponent = e ? e - 126 : 0;
    return (uint32_t)min(max(exponent + 1, 0), 7);
}
__device__ __forceinline__ float distance_to_next_voxel(float px, float py, float pz, float dx, float dy, float dz,
                                                        float ix, float iy, float iz, float res) {
    float qx = res * (px - 0.5f), qy = res * (py - 0.5f), qz = res * (pz - 0.5f);
    float tx = (floorf(qx + 0.5f + 0.5f * copysignf(1.0f, dx)) - qx) * ix;
    float ty = (floorf(qy + 0.5f + 0.5f * copysignf(1.0f, dy)) - qy) * iy;
    float tz = (floorf(qz + 0.5f + 0.5f * copysignf(1.0f, dz)) - qz) * iz;
    float t = fminf(fminf(tx, ty), tz);
    return fmaxf(t / res, 0.0f);
}
__device__ __forceinline__ float advance_to_next_voxel(float t, const StepC& cone, float px, float py, float pz, float dx, float dy,
                                                       float dz, float ix, float iy, float iz, uint32_t mip) {
    float res = res_of_mip(mip);
    float t_target = t + distance_to_next_voxel(px, py, pz, dx, dy, dz, ix, iy, iz, res);
    t = to_stepping_space(t, cone);
    t_target = to_stepping_space(t_target, cone);
    return from_stepping_space(t + ceilf(fmaxf(t_target - t, 0.5f)), cone);
}

struct RayGeom {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz;
    float t_exit;   // where the ray leaves the box around all occupied cells: nothing to find beyond it
};

__device__ __forceinline__ bool raabb_contains(const ModelDev& M, float px, float py, float pz) {
    if (!M.r2l_identity) {
        const float lx = M.r2l[0] * px + M.r2l[1] * py + M.r2l[2] * pz;
        const float ly = M.r2l[3] * px + M.r2l[4] * py + M.r2l[5] * pz;
        const float lz = M.r2l[6] * px + M.r2l[7] * py + M.r2l[8] * pz;
        px = lx; py = ly; pz = lz;
    }
    return px >= M.raabb_min[0] && px <= M.raabb_max[0] && py >= M.raabb_min[1] && py <= M.raabb_max[1] &&
           pz >= M.raabb_min[2] && pz <= M.raabb_max[2];
}

// if_unoccupied_advance_to_next_occupied_voxel<false>, min_mip = 0
__device__ __forceinline__ float skip_to_occupied(float t, const StepC& cone, const RayGeom& r, const ModelDev& M) {
    const uint32_t max_mip = (uint32_t)M.max_cascade;
    while (true) {
        const float px = r.ox + t * r.dx, py = r.oy + t * r.dy, pz = r.oz + t * r.dz;
        // (t > t_exit: result-preserving early out -- the reference keeps stepping through empty voxels until it
        // leaves the render aabb, which yields MAX_DEPTH as well)
        if (t >= MAX_DEPTH() || t > r.t_exit || !raabb_contains(M, px, py, pz)) return MAX_DEPTH();
        uint32_t mip = min(max(mip_from_pos(px, py, pz), 0u), max_mip);
        if (density_grid_occupied_at(px, py, pz, M.bitfield_lin, mip)) return t;
        while (mip < max_mip && !density_grid_occupied_at(px, py, pz, M.bitfield_lin, mip + 1)) ++mip;
        t = advance_to_next_voxel(t, cone, px, py, pz, r.dx, r.dy, r.dz, r.ix, r.iy, r.iz, mip);
    }
}

// Jump over the empty space in front of the box that bounds every occupied cell.  Every t the
// reference visits along a ray is from_stepping_space(s0 + n), n integer, s0 = to_stepping_space of
// the jittered start (sample steps add 1, advance_to_next_voxel adds ceil(.) >= 1), and its DDA only
// ever jumps across empty voxels, so the set of lattice points that fall into occupied cells -- the
// samples -- does not depend on where along the empty prefix the walk starts.  We restart it one
// lattice step before the box instead of walking ~50 voxels from the camera (values agree with the
// reference's walk up to the rounding of the to/from_stepping_space round trip).
__device__ __forceinline__ float fast_forward_to_box(float t, const StepC& cone, float t_box_entry) {
    if (t_box_entry > t) {
        const float s0 = to_stepping_space(t, cone), s1 = to_stepping_space(t_box_entry, cone);
        const float n = floorf(s1 - s0) - 1.0f;
        if (n >= 1.0f) t = from_stepping_space(s0 + n, cone);
    }
    return t;
}

// BoundingBox::ray_intersect (NGP bounding_box.cuh:163-213); returns tmin, FLT_MAX on a miss
__device__ __forceinline__ float2 box_ray_intersect(const float* mn, const float* mx, float ox, float oy, float oz,
                                                    float dx, float dy, float dz) {
    const float FMAX = 3.402823466e+38f;
    float tmin = (mn[0] - ox) / dx, tmax = (mx[0] - ox) / dx;
    if (tmin > tmax) { float s = tmin; tmin = tmax; tmax = s; }
    float tymin = (mn[1] - oy) / dy, tymax = (mx[1] - oy) / dy;
    if (tymin > tymax) { float s = tymin; tymin = tymax; tymax = s; }
    if (tmin > tymax || tymin > tmax) return make_float2(FMAX, FMAX);
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (mn[2] - oz) / dz, tzmax = (mx[2] - oz) / dz;
    if (tzmin > tzmax) { float s = tzmin; tzmin = tzmax; tzmax = s; }
    if (tmin > tzmax || tzmin > tmax) return make_float2(FMAX, FMAX);
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return make_float2(tmin, tmax);
}

// ---- hash-grid encoding: TCNN grid.h:47-165, common_device.h:697-713, 826-838 --------------------
// grid_index<3, CoherentPrime> restated without the runtime modulo.  Per level the host classifies it once
// (d2r_model_load, same stride loop as the reference):
//   hashed  <=> hashmap_size < res^3; then hashmap_size == 2^log2_hashmap_size, so `% size` is `& (size-1)`;
//   dense   <=> index = x + y*res + z*res^2 with coordinates in [0, res] (the +0.5 offset lets a corner reach
//               `res`: TCNN's "wraparound indexing in dense grids"), so index < 2*size and `% size` is one
//               conditional subtract.
// Same indices as the reference, an order of magnitude fewer instructions.
//
// one level -> 4 fp16 features (two half2).  The trilinear blend is an fp16 fma chain with the fp32
// weight rounded to fp16 first: `result = fma((T)weight, grid_val(...), result)` (grid.h:144-165)
__device__ __forceinline__ void encode_level(const ModelDev& M, int level, float x, float y, float z, __half2& f01, __half2& f23) {
    const uint2* __restrict__ table = M.level_table[level];    // 64-bit base once per level, 32-bit entry index per corner
    const uint32_t hashmap_size = M.level_size[level];
    const float scale = M.level_scale[level];
    float pos[3];
    uint32_t pg[3];
    const float in[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {    // pos_fract
        pos[d] = fmaf(scale, in[d], 0.5f);
        float tmp = floorf(pos[d]);
        pg[d] = (uint32_t)(int)tmp;
        pos[d] -= tmp;
    }
    uint2 v[8];
    if (M.level_hashed[level]) {
        const uint32_t mask = hashmap_size - 1;
        const uint32_t hx[2] = {pg[0], pg[0] + 1u};
        const uint32_t hy[2] = {pg[1] * 2654435761u, (pg[1] + 1u) * 2654435761u};
        const uint32_t hz[2] = {pg[2] * 805459861u, (pg[2] + 1u) * 805459861u};
#pragma unroll
        for (int idx = 0; idx < 8; ++idx)
            v[idx] = __ldg(table + ((hx[idx & 1] ^ hy[(idx >> 1) & 1] ^ hz[(idx >> 2) & 1]) & mask));
    } else {
        const uint32_t res = M.level_res[level];
        const uint32_t r2 = res * res;
        const uint32_t base = pg[0] + pg[1] * res + pg[2] * r2;
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) {
            uint32_t i = base + (idx & 1) + ((idx >> 1) & 1) * res + ((idx >> 2) & 1) * r2;
            if (i >= hashmap_size) i -= hashmap_size;
            v[idx] = __ldg(table + i);
        }
    }
    __half2 r01 = __float2half2_rn(0.f), r23 = __float2half2_rn(0.f);
#pragma unroll
    for (int idx = 0; idx < 8; ++idx) {
        float weight = 1;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if ((idx & (1 << d)) == 0) weight *= 1 - pos[d];
            else weight *= pos[d];
        }
        const __half2 w2 = __float2half2_rn(weight);
        r01 = __hfma2(w2, *reinterpret_cast<const __half2*>(&v[idx].x), r01);
        r23 = __hfma2(w2, *reinterpret_cast<const __half2*>(&v[idx].y), r23);
    }
    f01 = r01;
    f23 = r23;
}

// NL consecutive levels at once: all 8*NL gathers are issued before the first blend, so a thread keeps 8*NL loads
// in flight instead of 8 (the per-level hashed/dense branch otherwise fences each level's loads off from the next's).
// Same arithmetic per level as encode_level.
template <int NL>
__device__ __forceinline__ void encode_levels(const ModelDev& M, int level0, float x, float y, float z, __half2* f /* [2 * NL] */) {
    uint32_t ei[NL][8];
    float pos[NL][3];
    const uint2* table[NL];
    const float in[3] = {x, y, z};
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int level = level0 + l;
        table[l] = M.level_table[level];
        const uint32_t hashmap_size = M.level_size[level];
        const float scale = M.level_scale[level];
        uint32_t pg[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {    // pos_fract
            pos[l][d] = fmaf(scale, in[d], 0.5f);
            float tmp = floorf(pos[l][d]);
            pg[d] = (uint32_t)(int)tmp;
            pos[l][d] -= tmp;
        }
        if (M.level_hashed[level]) {
            const uint32_t mask = hashmap_size - 1;
            const uint32_t hx[2] = {pg[0], pg[0] + 1u};
            const uint32_t hy[2] = {pg[1] * 2654435761u, (pg[1] + 1u) * 2654435761u};
            const uint32_t hz[2] = {pg[2] * 805459861u, (pg[2] + 1u) * 805459861u};
#pragma unroll
            for (int idx = 0; idx < 8; ++idx) ei[l][idx] = (hx[idx & 1] ^ hy[(idx >> 1) & 1] ^ hz[(idx >> 2) & 1]) & mask;
        } else {
            const uint32_t res = M.level_res[level];
            const uint32_t r2 = res * res;
            const uint32_t base = pg[0] + pg[1] * res + pg[2] * r2;
#pragma unroll
            for (int idx = 0; idx < 8; ++idx) {
                uint32_t i = base + (idx & 1) + ((idx >> 1) & 1) * res + ((idx >> 2) & 1) * r2;
                if (i >= hashmap_size) i -= hashmap_size;
                ei[l][idx] = i;
            }
        }
    }
    uint2 v[NL][8];
#pragma unroll
    for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) v[l][idx] = __ldg(table[l] + ei[l][idx]);
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        __half2 r01 = __float2half2_rn(0.f), r23 = __float2half2_rn(0.f);
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) {
            float weight = 1;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if ((idx & (1 << d)) == 0) weight *= 1 - pos[l][d];
                else weight *= pos[l][d];
            }
            const __half2 w2 = __float2half2_rn(weight);
            r01 = __hfma2(w2, *reinterpret_cast<const __half2*>(&v[l][idx].x), r01);
            r23 = __hfma2(w2, *reinterpret_cast<const __half2*>(&v[l][idx].y), r23);
        }
        f[2 * l] = r01;
        f[2 * l + 1] = r23;
    }
}

// ---- spherical harmonics degree 4: TCNN common_device.h:340-365 ----------------------------------
__device__ __forceinline__ void sh_enc4(float x, float y, float z, float* o) {
    float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// ---- colour: NGP common_device.cuh:34-40, TCNN common_device.h:42-44 -----------------------------
__device__ __forceinline__ float srgb_to_linear_d(float srgb) {
    if (srgb <= 0.04045f) return srgb / 12.92f;
    return __powf((srgb + 0.055f) / 1.055f, 2.4f);
}
__device__ __forceinline__ float logistic_d(float x) { return 1.0f / (1.0f + __expf(-x)); }

}  // namespace d2r
