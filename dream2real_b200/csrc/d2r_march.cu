// NeRF ray march for K candidate cameras (sm_100a): ray generation -> occupancy DDA -> hash-grid encode ->
// density + colour MLPs -> transmittance compositing (colour AND depth in one pass) -> shade /
// tonemap epilogue -> optional depth-test composite against a cached background + sRGB/u8 pack.
//
// Replaces, per candidate pose, two complete Testbed::render calls of the reference
// (NGP src/testbed_nerf.cu:1549-1980: init_rays_with_payload_kernel_nerf, advance_pos_nerf_kernel,
//  compact_kernel_nerf, generate_next_nerf_network_inputs, kernel_grid, kernel_mlp_fused x2,
//  kernel_sh, extract_density, composite_kernel_nerf, shade_kernel_nerf; src/render_buffer.cu:
//  228-262, 529-561) plus the NumPy compositing of reconstruction/combined_rendering.py:133-155.
//
// Work decomposition (launch_march below): every candidate gets a conservative screen rectangle that contains every
// pixel whose ray can touch an occupied density-grid cell; rectangles are cut into 16x8 tiles; k_classify walks every
// tile pixel to its first occupied sample and builds the hit list; then either
//   * rounds of k_gather_round + k_mlp_round (d2r_march_split.cuh; the default for launches of >= 2^20 rays), or
//   * one persistent fused kernel, k_march_tc2 / k_march_tc (d2r_march_tc2.cuh, d2r_march_tc.cuh), that refills ray slots
//     from the hit list (small launches; D2R_MARCH=fused|solo|lpi4|tc1 force a variant), or
//   * k_march, the round-1 CUDA-core bring-up kernel in this file (D2R_MARCH=simt),
// take the rays to the end, and k_finish turns the per-ray accumulators into pixels.
#include <algorithm>
#include <vector>

#include <stdlib.h>
#include <string.h>

#include "d2r_gemm.cuh"
#include "d2r_march.cuh"

namespace d2r {

constexpr int TILE_W = 16, TILE_H = 8, CTA = TILE_W * TILE_H;   // 128 threads = 128 rays
constexpr int MARCH_ITER = 10000;                                // NGP src/testbed_nerf.cu:59

struct RayEntry {      // one primary ray that found an occupied sample (k_classify -> k_march_tc)
    uint32_t k;        // candidate
    uint32_t idx;      // pixel index x + W*y
    float t;           // ray parameter of the first occupied sample
    float t_exit;      // exit of the occupied box
};

struct MarchParams {
    ModelDev M;
    const float2* dirs;
    int W, H;
    const Mat3x4* cams;
    int K;
    const int4* bbox;              // per candidate: x0, y0, x1, y1 (inclusive), x1 < x0 = empty
    const uint32_t* tile_prefix;   // [K+1]
    const uint16_t* tile_cand;     // [total tiles]: candidate of every tile (saves the binary search over tile_prefix)
    uint32_t* counter;             // work-queue head
    float bg[4];                   // Testbed.background_color of the rendered model (sRGB + alpha)
    float4* rgba_out;              // [K,H,W] or null   (Shade)
    float4* depth_out;             // [K,H,W] or null   (Depth)
    const float4* bg_rgba;         // composite mode: cached background render [H,W]
    const float* bg_depth;         //                 cached background depth  [H,W]
    uint8_t* u8_out;               //                 [K,H,W,3]
    unsigned long long* n_samples;
    unsigned long long* prof;      // profiling counters {samples, primary rays owned, work items} or null
    RayEntry* entries;             // hit list (tensor-core path)
    uint32_t* n_entries;           //   number of entries (device)
    uint32_t* entry_cursor;        //   consumption cursor (device)
    float4* res_rgbd;              //   per-entry accumulated (r, g, b, depth), written by k_march_tc
    float* res_a;                  //   per-entry accumulated alpha
};

// shared-memory plan (floats): fp32 copies of the fp16 MLP weights, row-major [out][in]
constexpr int SW_D0 = 0, SW_D1 = SW_D0 + 64 * 32, SW_C0 = SW_D1 + 16 * 64, SW_C1 = SW_C0 + 64 * 32,
              SW_C2 = SW_C1 + 64 * 64, SW_END = SW_C2 + 4 * 64;
constexpr size_t SMEM_BYTES = SW_END * sizeof(float) + 64 * CTA * sizeof(__half) + 16;

__device__ __forceinline__ float h2f_round(float v) { return __half2float(__float2half_rn(v)); }

// One FullyFusedMLP layer for this thread's sample: out[j] = act(sum_k W[j][k] * in[k]), fp32
// accumulate, result rounded to fp16 (TCNN keeps activations in fp16, fully_fused_mlp.cu:47-129).
template <int NIN, int NOUT, bool RELU, bool TO_SMEM>
__device__ __forceinline__ void mlp_layer(const float* __restrict__ Wsm, const float (&in)[NIN], __half* __restrict__ act_col,
                                          float* out_regs) {
#pragma unroll 1
    for (int j = 0; j < NOUT; j += 4) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float4* w0 = reinterpret_cast<const float4*>(Wsm + (j + 0) * NIN);
        const float4* w1 = reinterpret_cast<const float4*>(Wsm + (j + 1) * NIN);
        const float4* w2 = reinterpret_cast<const float4*>(Wsm + (j + 2) * NIN);
        const float4* w3 = reinterpret_cast<const float4*>(Wsm + (j + 3) * NIN);
#pragma unroll
        for (int k = 0; k < NIN / 4; ++k) {
            const float4 q0 = w0[k], q1 = w1[k], q2 = w2[k], q3 = w3[k];
            a0 = fmaf(q0.x, in[4 * k], a0); a0 = fmaf(q0.y, in[4 * k + 1], a0); a0 = fmaf(q0.z, in[4 * k + 2], a0); a0 = fmaf(q0.w, in[4 * k + 3], a0);
            a1 = fmaf(q1.x, in[4 * k], a1); a1 = fmaf(q1.y, in[4 * k + 1], a1); a1 = fmaf(q1.z, in[4 * k + 2], a1); a1 = fmaf(q1.w, in[4 * k + 3], a1);
            a2 = fmaf(q2.x, in[4 * k], a2); a2 = fmaf(q2.y, in[4 * k + 1], a2); a2 = fmaf(q2.z, in[4 * k + 2], a2); a2 = fmaf(q2.w, in[4 * k + 3], a2);
            a3 = fmaf(q3.x, in[4 * k], a3); a3 = fmaf(q3.y, in[4 * k + 1], a3); a3 = fmaf(q3.z, in[4 * k + 2], a3); a3 = fmaf(q3.w, in[4 * k + 3], a3);
        }
        if (RELU) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
        if (TO_SMEM) {
            act_col[(j + 0) * CTA] = __float2half_rn(a0);
            act_col[(j + 1) * CTA] = __float2half_rn(a1);
            act_col[(j + 2) * CTA] = __float2half_rn(a2);
            act_col[(j + 3) * CTA] = __float2half_rn(a3);
        } else {
            out_regs[j + 0] = h2f_round(a0); out_regs[j + 1] = h2f_round(a1);
            out_regs[j + 2] = h2f_round(a2); out_regs[j + 3] = h2f_round(a3);
        }
    }
}

// python side of the path: reconstruction/combined_rendering.py:133-155 + NGP scripts/common.py:142-144
__device__ __forceinline__ float linear_to_srgb_py(float x) {
    // numpy evaluates every operator separately in float32: no fma contraction here
    return x > 0.0031308f ? __fsub_rn(__fmul_rn(1.055f, powf(x, 1.0f / 2.4f)), 0.055f) : __fmul_rn(12.92f, x);
}
__device__ __forceinline__ uint8_t to_u8(float v) {
    return (uint8_t)__fadd_rn(__fmul_rn(fminf(fmaxf(v, 0.0f), 1.0f), 255.0f), 0.5f);
}

__device__ __forceinline__ void composite_pixel(float4 fg, float fg_d, float4 bgc, float bg_d, uint8_t* out3) {
    if (fg_d < 0.05f) fg_d = 100.f;
    if (bg_d < 0.05f) bg_d = 100.f;
    const float4 c = (fg_d < bg_d) ? fg : bgc;
    float r = 0.f, g = 0.f, b = 0.f;
    if (c.w != 0.f) { r = __fdiv_rn(c.x, c.w); g = __fdiv_rn(c.y, c.w); b = __fdiv_rn(c.z, c.w); }
    const uint8_t a8 = to_u8(c.w);
    uint8_t r8 = to_u8(linear_to_srgb_py(r)), g8 = to_u8(linear_to_srgb_py(g)), b8 = to_u8(linear_to_srgb_py(b));
    if (a8 < 130) { r8 = 0; g8 = 0; b8 = 0; }
    out3[0] = r8; out3[1] = g8; out3[2] = b8;
}

__global__ void __launch_bounds__(CTA, 3) k_march(const __grid_constant__ MarchParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Wsm = reinterpret_cast<float*>(smem_raw);
    __half* act = reinterpret_cast<__half*>(smem_raw + SW_END * sizeof(float));
    uint32_t* s_tile = reinterpret_cast<uint32_t*>(smem_raw + SW_END * sizeof(float) + 64 * CTA * sizeof(__half));
    const ModelDev& M = P.M;
    const int tid = threadIdx.x;

    // stage MLP weights (fp16 -> fp32, exact) once per CTA
    for (int i = tid; i < 64 * 32; i += CTA) Wsm[SW_D0 + i] = __half2float(M.w_d0[i]);
    for (int i = tid; i < 16 * 64; i += CTA) Wsm[SW_D1 + i] = __half2float(M.w_d1[i]);
    for (int i = tid; i < 64 * 32; i += CTA) Wsm[SW_C0 + i] = __half2float(M.w_c0[i]);
    for (int i = tid; i < 64 * 64; i += CTA) Wsm[SW_C1 + i] = __half2float(M.w_c1[i]);
    for (int i = tid; i < 4 * 64; i += CTA) Wsm[SW_C2 + i] = __half2float(M.w_c2[i]);
    __half* act_col = act + tid;
    const uint32_t total_tiles = P.tile_prefix[P.K];
    unsigned long long my_samples = 0;

    while (true) {
        __syncthreads();
        if (tid == 0) *s_tile = atomicAdd(P.counter, 1u);
        __syncthreads();
        const uint32_t tile = *s_tile;
        if (tile >= total_tiles) break;
        // candidate = last k with prefix[k] <= tile
        int lo = 0, hi = P.K;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (P.tile_prefix[mid] <= tile) lo = mid; else hi = mid;
        }
        const int k = lo;
        const int4 bb = P.bbox[k];
        const uint32_t local = tile - P.tile_prefix[k];
        const int tiles_x = (bb.z - bb.x + TILE_W) / TILE_W;
        const int x = bb.x + (int)(local % tiles_x) * TILE_W + (tid % TILE_W);
        const int y = bb.y + (int)(local / tiles_x) * TILE_H + (tid / TILE_W);
        if (x > bb.z || y > bb.w) continue;
        const uint32_t idx = (uint32_t)x + (uint32_t)P.W * (uint32_t)y;

        // ---- init_rays_with_payload_kernel_nerf (NGP testbed_nerf.cu:1394-1482) ----
        const Mat3x4 C = P.cams[k];
        const float2 dc = __ldg(P.dirs + idx);
        RayGeom r;
        {   // dir = mat3(camera) * (dc.x, dc.y, 1); tcnn mat*vec accumulates column by column
            float vx = 0.f, vy = 0.f, vz = 0.f;
            vx += C.c[0][0] * dc.x; vy += C.c[0][1] * dc.x; vz += C.c[0][2] * dc.x;
            vx += C.c[1][0] * dc.y; vy += C.c[1][1] * dc.y; vz += C.c[1][2] * dc.y;
            vx += C.c[2][0] * 1.0f; vy += C.c[2][1] * 1.0f; vz += C.c[2][2] * 1.0f;
            float len2 = 0.f;
            len2 += vx * vx; len2 += vy * vy; len2 += vz * vz;
            const float len = sqrtf(len2);
            r.dx = vx / len; r.dy = vy / len; r.dz = vz / len;
            r.ox = C.c[3][0]; r.oy = C.c[3][1]; r.oz = C.c[3][2];
            r.ix = 1.0f / r.dx; r.iy = 1.0f / r.dy; r.iz = 1.0f / r.dz;
        }
        bool alive;
        float t, t_box = 0.f;
        {
            float lox = r.ox, loy = r.oy, loz = r.oz, ldx = r.dx, ldy = r.dy, ldz = r.dz;
            if (!M.r2l_identity) {
                lox = M.r2l[0] * r.ox + M.r2l[1] * r.oy + M.r2l[2] * r.oz;
                loy = M.r2l[3] * r.ox + M.r2l[4] * r.oy + M.r2l[5] * r.oz;
                loz = M.r2l[6] * r.ox + M.r2l[7] * r.oy + M.r2l[8] * r.oz;
                ldx = M.r2l[0] * r.dx + M.r2l[1] * r.dy + M.r2l[2] * r.dz;
                ldy = M.r2l[3] * r.dx + M.r2l[4] * r.dy + M.r2l[5] * r.dz;
                ldz = M.r2l[6] * r.dx + M.r2l[7] * r.dy + M.r2l[8] * r.dz;
            }
            t = fmaxf(box_ray_intersect(M.raabb_min, M.raabb_max, lox, loy, loz, ldx, ldy, ldz).x, 0.0f) + 1e-6f;
            alive = raabb_contains(M, r.ox + t * r.dx, r.oy + t * r.dy, r.oz + t * r.dz);
            // rays that miss the box around all occupied cells can never take a sample
            if (alive) {
                const float2 oc = box_ray_intersect(M.occ_min, M.occ_max, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz);
                if (oc.x > 1e37f || oc.y < 0.f) alive = false;
                t_box = oc.x;
                r.t_exit = oc.y;
            }
        }
        const StepC cone = make_stepc(M.cone);   // calc_cone_angle returns the constant (nerf_device.cuh:369-376)
        // ---- advance_pos_nerf (testbed_nerf.cu:333-362) ----
        if (alive) {
            t = advance_n_steps(t, cone, ld_random_val0(idx * 786433u));
            t = fast_forward_to_box(t, cone, t_box);
            t = skip_to_occupied(t, cone, r, M);
            if (t >= MAX_DEPTH()) alive = false;
        }
        float cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f, ca = 0.f;
        if (alive) {
            float sh[16];
            {   // kernel_sh input is warp_direction(dir) = (dir+1)*0.5, mapped back by *2-1 (spherical_harmonics.h:62-70)
                const float wx = (r.dx + 1.0f) * 0.5f, wy = (r.dy + 1.0f) * 0.5f, wz = (r.dz + 1.0f) * 0.5f;
                sh_enc4(wx * 2.f - 1.f, wy * 2.f - 1.f, wz * 2.f - 1.f, sh);
#pragma unroll
                for (int i = 0; i < 16; ++i) sh[i] = h2f_round(sh[i]);
            }
            const float fwx = C.c[2][0], fwy = C.c[2][1], fwz = C.c[2][2];
            int n_steps = 0;
            while (true) {
                // ---- generate_next_nerf_network_inputs (testbed_nerf.cu:454-467) ----
                t = skip_to_occupied(t, cone, r, M);
                if (t >= MAX_DEPTH()) break;
                const float dt = calc_dt(t, cone);
                const float px = r.ox + r.dx * t, py = r.oy + r.dy * t, pz = r.oz + r.dz * t;
                const float wpx = (px - M.aabb_min[0]) / M.aabb_diag[0];        // warp_position = aabb.relative_pos
                const float wpy = (py - M.aabb_min[1]) / M.aabb_diag[1];
                const float wpz = (pz - M.aabb_min[2]) / M.aabb_diag[2];
                const float wdt = warp_dt(dt);
                t += dt;
                ++my_samples;
                // ---- NerfNetwork::inference_mixed_precision (nerf_network.h:105-140) ----
                float raw[4];
                {
                    float in[32];
#pragma unroll
                    for (int l = 0; l < 8; ++l) {
                        __half2 f01, f23;
                        encode_level(M, l, wpx, wpy, wpz, f01, f23);
                        const float2 a = __half22float2(f01), b = __half22float2(f23);
                        in[4 * l] = a.x; in[4 * l + 1] = a.y; in[4 * l + 2] = b.x; in[4 * l + 3] = b.y;
                    }
                    mlp_layer<32, 64, true, true>(Wsm + SW_D0, in, act_col, nullptr);
                }
                {
                    float in[64];
#pragma unroll
                    for (int i = 0; i < 64; ++i) in[i] = __half2float(act_col[i * CTA]);
                    float in2[32];
                    mlp_layer<64, 16, false, false>(Wsm + SW_D1, in, nullptr, in2);
                    raw[3] = in2[0];                                              // extract_density: row 0
#pragma unroll
                    for (int i = 0; i < 16; ++i) in2[16 + i] = sh[i];
                    mlp_layer<32, 64, true, true>(Wsm + SW_C0, in2, act_col, nullptr);
                }
                {
                    float in[64];
#pragma unroll
                    for (int i = 0; i < 64; ++i) in[i] = __half2float(act_col[i * CTA]);
                    mlp_layer<64, 64, true, true>(Wsm + SW_C1, in, act_col, nullptr);
                }
                {
                    float in[64];
#pragma unroll
                    for (int i = 0; i < 64; ++i) in[i] = __half2float(act_col[i * CTA]);
                    float o4[4];
                    mlp_layer<64, 4, false, false>(Wsm + SW_C2, in, nullptr, o4);
                    raw[0] = o4[0]; raw[1] = o4[1]; raw[2] = o4[2];
                }
                // ---- composite_kernel_nerf (testbed_nerf.cu:511-667) ----
                const float ux = M.aabb_min[0] + wpx * M.aabb_diag[0];           // unwarp_position
                const float uy = M.aabb_min[1] + wpy * M.aabb_diag[1];
                const float uz = M.aabb_min[2] + wpz * M.aabb_diag[2];
                const float T = 1.f - ca;
                const float dtu = unwarp_dt(wdt);
                const float alpha = 1.f - __expf(-__expf(raw[3]) * dtu);
                const float weight = alpha * T;
                const float rr = logistic_d(raw[0]), gg = logistic_d(raw[1]), bb_ = logistic_d(raw[2]);
                float dep = 0.f;
                dep += fwx * (ux - r.ox); dep += fwy * (uy - r.oy); dep += fwz * (uz - r.oz);
                dep *= M.depth_scale;
                cr += rr * weight; cg += gg * weight; cb += bb_ * weight; cd += dep * weight; ca += weight;
                if (ca > (1.0f - M.min_transmittance)) {
                    cr /= ca; cg /= ca; cb /= ca; cd /= ca; ca /= ca;
                    break;
                }
                if (++n_steps >= MARCH_ITER - 1) { cr = cg = cb = cd = ca = 0.f; break; }   // never reaches the hit buffer
            }
        }
        // ---- compact (keep a > 0.001) + shade_kernel_nerf + accumulate + tonemap background ----
        if (!(ca > 0.001f)) { cr = cg = cb = cd = ca = 0.f; }
        float4 shade = make_float4(srgb_to_linear_d(cr), srgb_to_linear_d(cg), srgb_to_linear_d(cb), ca);
        float4 depth = make_float4(cd, cd, cd, ca);
        {
            const float w = (1.f - ca) * P.bg[3];
            const float blr = srgb_to_linear_d(P.bg[0]), blg = srgb_to_linear_d(P.bg[1]), blb = srgb_to_linear_d(P.bg[2]);
            shade.x += blr * w; shade.y += blg * w; shade.z += blb * w; shade.w += w;
            depth.x += blr * w; depth.y += blg * w; depth.z += blb * w; depth.w += w;
        }
        const size_t o = (size_t)k * P.W * P.H + idx;
        if (P.rgba_out) P.rgba_out[o] = shade;
        if (P.depth_out) P.depth_out[o] = depth;
        if (P.u8_out) composite_pixel(shade, depth.x, __ldg(P.bg_rgba + idx), __ldg(P.bg_depth + idx), P.u8_out + o * 3);
    }
    if (P.n_samples || P.prof) {
        for (int o = 16; o > 0; o >>= 1) my_samples += __shfl_xor_sync(0xffffffffu, my_samples, o);
        if ((tid & 31) == 0 && my_samples) {
            if (P.n_samples) atomicAdd(P.n_samples, my_samples);
            if (P.prof) atomicAdd(P.prof, my_samples);
        }
        if (P.prof && blockIdx.x == 0 && tid == 0) atomicAdd(P.prof + 1, (unsigned long long)total_tiles * CTA);
    }
}

}  // namespace d2r
#include "d2r_march_tc.cuh"
#include "d2r_march_tc2.cuh"
#include "d2r_march_split.cuh"
namespace d2r {

// ---- per-candidate screen rectangle + tile prefix ---------------------------------------------------
// Conservative: contains every pixel whose (undistorted) camera-plane direction lies inside the
// perspective projection of the box around all occupied cells.  col_lo/col_hi (row_lo/row_hi) are the
// per-column (per-row) min/max of the direction table, so lens distortion is handled exactly.
__global__ void k_candidate_bbox(int K, int W, int H, const Mat3x4* __restrict__ cams, const float* __restrict__ col_lo,
                                 const float* __restrict__ col_hi, const float* __restrict__ row_lo,
                                 const float* __restrict__ row_hi, const float occ_min_x, const float occ_min_y,
                                 const float occ_min_z, const float occ_max_x, const float occ_max_y, const float occ_max_z,
                                 int /*unused*/, int4* __restrict__ bbox, uint32_t* __restrict__ tiles) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;     // one warp per candidate
    if (k >= K) return;
    int x0 = 0, y0 = 0, x1 = W - 1, y1 = H - 1;
    {
        const Mat3x4 C = cams[k];
        float u0 = 1e30f, u1 = -1e30f, v0 = 1e30f, v1 = -1e30f;
        bool behind = false;
        for (int c = 0; c < 8; ++c) {
            const float wx = ((c & 1) ? occ_max_x : occ_min_x) - C.c[3][0];
            const float wy = ((c & 2) ? occ_max_y : occ_min_y) - C.c[3][1];
            const float wz = ((c & 4) ? occ_max_z : occ_min_z) - C.c[3][2];
            // camera-space = R^T * (p - o)   (columns of C are the camera axes)
            const float cx = C.c[0][0] * wx + C.c[0][1] * wy + C.c[0][2] * wz;
            const float cy = C.c[1][0] * wx + C.c[1][1] * wy + C.c[1][2] * wz;
            const float cz = C.c[2][0] * wx + C.c[2][1] * wy + C.c[2][2] * wz;
            if (cz < 1e-3f) { behind = true; break; }
            const float u = __fdiv_rn(cx, cz), v = __fdiv_rn(cy, cz);
            u0 = fminf(u0, u); u1 = fmaxf(u1, u); v0 = fminf(v0, v); v1 = fmaxf(v1, v);
        }
        if (!behind) {      // (uniform across the warp: every lane did the same arithmetic)
            const float eu = 1e-4f * (1.f + fmaxf(fabsf(u0), fabsf(u1))), ev = 1e-4f * (1.f + fmaxf(fabsf(v0), fabsf(v1)));
            u0 -= eu; u1 += eu; v0 -= ev; v1 += ev;
            x0 = W; x1 = -1; y0 = H; y1 = -1;
            for (int x = lane; x < W; x += 32) if (col_hi[x] >= u0 && col_lo[x] <= u1) { x0 = min(x0, x); x1 = max(x1, x); }
            for (int y = lane; y < H; y += 32) if (row_hi[y] >= v0 && row_lo[y] <= v1) { y0 = min(y0, y); y1 = max(y1, y); }
            for (int o = 16; o > 0; o >>= 1) {
                x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
                y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
            }
            if (x1 >= x0 && y1 >= y0) {
                x0 = max(x0 - 1, 0); y0 = max(y0 - 1, 0); x1 = min(x1 + 1, W - 1); y1 = min(y1 + 1, H - 1);
            }
        }
    }
    if (lane != 0) return;
    uint32_t n = 0;
    if (x1 >= x0 && y1 >= y0) {
        n = (uint32_t)((x1 - x0 + TILE_W) / TILE_W) * (uint32_t)((y1 - y0 + TILE_H) / TILE_H);
    }
    else { x0 = 0; y0 = 0; x1 = -1; y1 = -1; }
    bbox[k] = make_int4(x0, y0, x1, y1);
    tiles[k] = n;
}

// exclusive scan of tiles[K] into prefix[K+1] (single CTA; K is at most a few hundred thousand)
__global__ void k_tile_prefix(int K, const uint32_t* __restrict__ tiles, uint32_t* __restrict__ prefix, uint32_t* __restrict__ counter) {
    __shared__ uint32_t s[1024];
    const int tid = threadIdx.x;
    const int chunk = (K + 1023) / 1024;
    const int b = tid * chunk, e = min(b + chunk, K);
    uint32_t sum = 0;
    for (int i = b; i < e; ++i) sum += tiles[i];
    s[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        uint32_t v = tid >= o ? s[tid - o] : 0;
        __syncthreads();
        s[tid] += v;
        __syncthreads();
    }
    uint32_t run = s[tid] - sum;
    for (int i = b; i < e; ++i) { prefix[i] = run; run += tiles[i]; }
    if (tid == 1023) prefix[K] = s[1023];
    if (tid == 0) *counter = 0;
}

__global__ void k_tile_map(int K, const uint32_t* __restrict__ prefix, uint16_t* __restrict__ map) {
    const int k = blockIdx.x;
    for (uint32_t t = prefix[k] + threadIdx.x; t < prefix[k + 1]; t += blockDim.x) map[t] = (uint16_t)k;
}

// Every frame starts as a copy of the composited background (u8) / the constant background blend (float);
// the march kernel then overwrites the pixels of the candidate's rectangle.  Pure streaming stores:
// 128-bit vectors whenever a frame is a whole number of 16-byte words.
__global__ void k_fill_frames(int K, int W, int H, float4 shade_bg, float4 depth_bg, float4* __restrict__ rgba_out,
                              float4* __restrict__ depth_out, const uint8_t* __restrict__ bg_u8, uint8_t* __restrict__ u8_out, int vec_ok) {
    const int k = blockIdx.x;
    const size_t npx = (size_t)W * H;
    const size_t stride = (size_t)gridDim.y * blockDim.x, first = (size_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (rgba_out) for (size_t p = first; p < npx; p += stride) rgba_out[(size_t)k * npx + p] = shade_bg;
    if (depth_out) for (size_t p = first; p < npx; p += stride) depth_out[(size_t)k * npx + p] = depth_bg;
    if (u8_out) {
        const size_t bytes = npx * 3;
        if (vec_ok) {
            const uint4* __restrict__ src = reinterpret_cast<const uint4*>(bg_u8);
            uint4* __restrict__ dst = reinterpret_cast<uint4*>(u8_out + (size_t)k * bytes);
            for (size_t i = first; i < bytes / 16; i += stride) dst[i] = __ldg(src + i);
        } else {
            for (size_t i = first; i < bytes; i += stride) u8_out[(size_t)k * bytes + i] = bg_u8[i];
        }
    }
}

__global__ void k_bg_u8(int P_, float4 fg_empty, const float4* __restrict__ bg_rgba, const float* __restrict__ bg_depth, uint8_t* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P_) return;
    composite_pixel(fg_empty, 0.f, bg_rgba[p], bg_depth[p], out + (size_t)p * 3);
}

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

// ---- profiling: CUDA events around the march kernel only (bench.py roofline) ----------------------
struct Prof {
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    size_t used = 0;
    unsigned long long* counters = nullptr;   // device {samples, rays, items}
};
static Prof g_prof[16];

// ---- host launcher -------------------------------------------------------------------------------
struct Scratch {   // per-device scratch reused across calls (grown on demand)
    Mat3x4* cams = nullptr; int4* bbox = nullptr; uint32_t* tiles = nullptr; uint32_t* prefix = nullptr;
    uint32_t* counter = nullptr; int capK = 0;
    float* ranges = nullptr; int capWH = 0; const void* ranges_view = nullptr;
    uint8_t* bg_u8 = nullptr; size_t cap_bg = 0;
    RayEntry* entries = nullptr; size_t cap_entries = 0; uint32_t* entry_counters = nullptr;
    uint16_t* tile_cand = nullptr; size_t cap_tile_cand = 0;
    float4* res_rgbd = nullptr; float* res_a = nullptr;
    int n_sm = 0;
    // round-based split path (k_gather_round / k_mlp_round)
    size_t cap_split = 0;
    unsigned char* sp_feat = nullptr; float2* sp_aux = nullptr; uint4* sp_shb = nullptr; uint8_t* sp_nsb = nullptr;
    float* sp_t = nullptr; uint32_t* sp_live[2] = {nullptr, nullptr}; uint32_t* sp_cnt = nullptr;
};
constexpr int SPLIT_MAX_ROUNDS = MARCH_ITER / 2 + 2;
static Scratch g_scratch[16];

static int ensure_scratch(int device, int K, int W, int H) {
    Scratch& s = g_scratch[device];
    if (K > s.capK) {
        cudaFree(s.cams); cudaFree(s.bbox); cudaFree(s.tiles); cudaFree(s.prefix);
        s.cams = nullptr; s.bbox = nullptr; s.tiles = nullptr; s.prefix = nullptr; s.capK = 0;
        D2R_CUDA(cudaMalloc(&s.cams, (size_t)K * sizeof(Mat3x4)));
        D2R_CUDA(cudaMalloc(&s.bbox, (size_t)K * sizeof(int4)));
        D2R_CUDA(cudaMalloc(&s.tiles, (size_t)K * sizeof(uint32_t)));
        D2R_CUDA(cudaMalloc(&s.prefix, (size_t)(K + 1) * sizeof(uint32_t)));
        s.capK = K;
    }
    if (!s.counter) D2R_CUDA(cudaMalloc(&s.counter, sizeof(uint32_t)));
    if (!s.entry_counters) D2R_CUDA(cudaMalloc(&s.entry_counters, 2 * sizeof(uint32_t)));
    if (W + H > s.capWH) {
        cudaFree(s.ranges);
        s.ranges = nullptr; s.capWH = 0;
        D2R_CUDA(cudaMalloc(&s.ranges, (size_t)2 * (W + H) * sizeof(float)));
        s.capWH = W + H; s.ranges_view = nullptr;
    }
    if ((size_t)W * H * 3 > s.cap_bg) {
        cudaFree(s.bg_u8);
        s.bg_u8 = nullptr; s.cap_bg = 0;
        D2R_CUDA(cudaMalloc(&s.bg_u8, (size_t)W * H * 3));
        s.cap_bg = (size_t)W * H * 3;
    }
    if (!s.n_sm) D2R_CUDA(cudaDeviceGetAttribute(&s.n_sm, cudaDevAttrMultiProcessorCount, device));
    return D2R_OK;
}

int launch_march(const d2r_model* m, const d2r_view* v, const float* cams_ngp_host, int K, const float bg[4],
                 float* rgba_out, float* depth_out, const float* bg_rgba, const float* bg_depth, uint8_t* u8_out,
                 unsigned long long* n_samples, cudaStream_t stream, int* rects_out = nullptr, uint8_t* bg_u8_out = nullptr) {
    D2R_REQUIRE(m && v && cams_ngp_host && bg, "render: null argument");
    D2R_REQUIRE(K > 0, "render: K must be positive");
    D2R_REQUIRE(K <= 65535, "render: at most 65535 candidates per launch");
    D2R_REQUIRE(m->device == v->device && m->device < 16, "render: model and view live on different devices");
    D2R_REQUIRE(rgba_out || depth_out || u8_out, "render: no output requested");
    D2R_REQUIRE(!u8_out || (bg_rgba && bg_depth), "render_composite: background buffers missing");
    D2R_CUDA(cudaSetDevice(m->device));
    const int W = v->W, H = v->H;
    int rc = ensure_scratch(m->device, K, W, H);
    if (rc) return rc;
    Scratch& s = g_scratch[m->device];

    // cameras: host [K,3,4] row-major (rows = xyz, cols = 3 axes + origin) -> column structs
    {
        std::vector<Mat3x4> tmp(K);
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < 4; ++c)
                for (int rr = 0; rr < 3; ++rr) tmp[k].c[c][rr] = cams_ngp_host[(size_t)k * 12 + rr * 4 + c];
        D2R_CUDA(cudaMemcpyAsync(s.cams, tmp.data(), (size_t)K * sizeof(Mat3x4), cudaMemcpyHostToDevice, stream));
        D2R_CUDA(cudaStreamSynchronize(stream));   // tmp is pageable and dies here
    }
    float* col_lo = s.ranges, *col_hi = s.ranges + W, *row_lo = s.ranges + 2 * W, *row_hi = s.ranges + 2 * W + H;
    if (s.ranges_view != (const void*)v->dirs_dev) {
        k_view_ranges<<<(std::max(W, H) + 127) / 128, 128, 0, stream>>>(W, H, v->dirs_dev, col_lo, col_hi, row_lo, row_hi);
        count_launch();
        s.ranges_view = (const void*)v->dirs_dev;
    }
    const ModelDev& M = m->dev;
    // D2R_MARCH=simt selects the round-1 CUDA-core kernel (kept for A/B measurements); default: tensor-core kernel
    static const bool use_tc = []() { const char* e = getenv("D2R_MARCH"); return !(e && strcmp(e, "simt") == 0); }();
    static const int ctas_per_sm = []() { const char* e = getenv("D2R_MARCH_CTAS"); const int v = e ? atoi(e) : 4; return v >= 1 && v <= 4 ? v : 4; }();
    static const bool use_solo = []() { const char* e = getenv("D2R_MARCH"); return e && strcmp(e, "solo") == 0; }();    // per-thread gathers (no lane pairing)
    // default: round-based gather / MLP kernels (d2r_march_split.cuh); D2R_MARCH=fused|solo|lpi4|tc1|simt select the older kernels
    static const bool use_split = []() { const char* e = getenv("D2R_MARCH"); return !e || !*e || strcmp(e, "split") == 0; }();
    // launches with few rays (a single background frame, small test renders) would spend their time on ~40 pairs of tiny
    // round kernels: below this many screen-rectangle rays the fused kernel (one launch, same results) takes them
    static const bool split_forced = []() { const char* e = getenv("D2R_MARCH"); return e && strcmp(e, "split") == 0; }();
    constexpr size_t SPLIT_MIN_RAYS = 1u << 20;
    static const bool split_mlp_old = []() { const char* e = getenv("D2R_SPLIT_MLP"); return e && strcmp(e, "old") == 0; }();
    static const bool split_coop = []() { const char* e = getenv("D2R_SPLIT_COOP"); return e && atoi(e) != 0; }();
    static const int split_gctas = []() { const char* e = getenv("D2R_SPLIT_GCTAS"); const int v = e ? atoi(e) : 7; return v >= 1 && v <= 16 ? v : 7; }();
    static const int abl = []() { const char* e = getenv("D2R_MARCH_ABL"); return e ? atoi(e) : 0; }();   // timing ablations, wrong-free results
    static const bool use_solo4 = []() { const char* e = getenv("D2R_MARCH"); return e && strcmp(e, "solo4") == 0; }();
    static const bool use_lpi4 = []() { const char* e = getenv("D2R_MARCH"); return e && strcmp(e, "lpi4") == 0; }();    // 4 levels per gather batch
    static const bool use_tc1 = []() { const char* e = getenv("D2R_MARCH"); return e && strcmp(e, "tc1") == 0; }();   // one sample per round
    k_candidate_bbox<<<(K * 32 + 127) / 128, 128, 0, stream>>>(K, W, H, s.cams, col_lo, col_hi, row_lo, row_hi, M.occ_min[0], M.occ_min[1],
                                                          M.occ_min[2], M.occ_max[0], M.occ_max[1], M.occ_max[2], 0, s.bbox, s.tiles);
    k_tile_prefix<<<1, 1024, 0, stream>>>(K, s.tiles, s.prefix, s.counter);
    count_launch(2);
    if (rects_out) D2R_CUDA(cudaMemcpyAsync(rects_out, s.bbox, (size_t)K * sizeof(int4), cudaMemcpyDeviceToDevice, stream));

    // what a pixel no ray reaches looks like: accumulate 0, then the tonemap background blend
    const float w0 = bg[3];
    auto s2l = [](float x) { return x <= 0.04045f ? x / 12.92f : powf((x + 0.055f) / 1.055f, 2.4f); };
    const float4 empty = make_float4(s2l(bg[0]) * w0, s2l(bg[1]) * w0, s2l(bg[2]) * w0, w0);
    if (u8_out) {
        k_bg_u8<<<(W * H + 255) / 256, 256, 0, stream>>>(W * H, empty, (const float4*)bg_rgba, bg_depth, s.bg_u8);
        count_launch();
        if (bg_u8_out) D2R_CUDA(cudaMemcpyAsync(bg_u8_out, s.bg_u8, (size_t)W * H * 3, cudaMemcpyDeviceToDevice, stream));
    }
    {
        dim3 grid(K, std::min((W * H * 3 / 16 + 255) / 256 + 1, 32));
        const int vec_ok = ((size_t)W * H * 3) % 16 == 0 && ((uintptr_t)u8_out % 16) == 0 && ((uintptr_t)s.bg_u8 % 16) == 0;
        k_fill_frames<<<grid, 256, 0, stream>>>(K, W, H, empty, empty, (float4*)rgba_out, (float4*)depth_out, s.bg_u8, u8_out, vec_ok);
        count_launch();
    }
    MarchParams P;
    P.M = M; P.dirs = v->dirs_dev; P.W = W; P.H = H; P.cams = s.cams; P.K = K; P.bbox = s.bbox; P.tile_prefix = s.prefix;
    P.counter = s.counter;
    for (int i = 0; i < 4; ++i) P.bg[i] = bg[i];
    P.rgba_out = (float4*)rgba_out; P.depth_out = (float4*)depth_out; P.bg_rgba = (const float4*)bg_rgba; P.bg_depth = bg_depth;
    P.u8_out = u8_out; P.n_samples = n_samples;
    P.tile_cand = nullptr;
    P.entries = nullptr; P.n_entries = nullptr; P.entry_cursor = nullptr; P.res_rgbd = nullptr; P.res_a = nullptr;
    static bool attr_set[16] = {false};
    if (!attr_set[m->device]) {
        D2R_CUDA(cudaFuncSetAttribute(k_march, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        D2R_CUDA(cudaFuncSetAttribute(k_march_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_TOTAL));
        D2R_CUDA(cudaFuncSetAttribute(k_march_tc2<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_TOTAL));
        D2R_CUDA(cudaFuncSetAttribute(k_march_tc2<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_TOTAL));
        D2R_CUDA(cudaFuncSetAttribute(k_march_tc2<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_TOTAL));
        D2R_CUDA(cudaFuncSetAttribute(k_march_tc2<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_TOTAL));
        D2R_CUDA(cudaFuncSetAttribute((k_march_tc2<2, true, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_TOTAL));
        D2R_CUDA(cudaFuncSetAttribute((k_march_tc2<2, true, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_TOTAL));
        attr_set[m->device] = true;
    }
    Prof& pf = g_prof[m->device];
    P.prof = pf.on ? pf.counters : nullptr;
    std::pair<cudaEvent_t, cudaEvent_t>* evp = nullptr;
    if (pf.on) {
        if (pf.used == pf.ev.size()) {
            cudaEvent_t a, b;
            D2R_CUDA(cudaEventCreate(&a));
            D2R_CUDA(cudaEventCreate(&b));
            pf.ev.emplace_back(a, b);
        }
        evp = &pf.ev[pf.used++];
        D2R_CUDA(cudaEventRecord(evp->first, stream));
    }
    if (use_tc) {
        // pass 1 needs the tile count on the host (grid size, hit-list capacity): one 4-byte read-back
        uint32_t total_tiles = 0;
        D2R_CUDA(cudaMemcpyAsync(&total_tiles, s.prefix + K, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        D2R_CUDA(cudaStreamSynchronize(stream));
        const size_t need = (size_t)total_tiles * CTA;
        if (need > s.cap_entries) {
            if (s.entries) { cudaFree(s.entries); cudaFree(s.res_rgbd); cudaFree(s.res_a); }
            s.entries = nullptr; s.res_rgbd = nullptr; s.res_a = nullptr; s.cap_entries = 0;
            D2R_CUDA(cudaMalloc(&s.entries, need * sizeof(RayEntry)));
            D2R_CUDA(cudaMalloc(&s.res_rgbd, need * sizeof(float4)));
            D2R_CUDA(cudaMalloc(&s.res_a, need * sizeof(float)));
            s.cap_entries = need;
        }
        if (total_tiles > s.cap_tile_cand) {
            if (s.tile_cand) cudaFree(s.tile_cand);
            s.tile_cand = nullptr; s.cap_tile_cand = 0;
            D2R_CUDA(cudaMalloc(&s.tile_cand, (size_t)total_tiles * sizeof(uint16_t)));
            s.cap_tile_cand = total_tiles;
        }
        D2R_CUDA(cudaMemsetAsync(s.entry_counters, 0, 2 * sizeof(uint32_t), stream));
        P.entries = s.entries; P.n_entries = s.entry_counters; P.entry_cursor = s.entry_counters + 1;
        P.res_rgbd = s.res_rgbd; P.res_a = s.res_a;
        P.tile_cand = s.tile_cand;
        if (total_tiles) {
            k_tile_map<<<K, 64, 0, stream>>>(K, s.prefix, s.tile_cand);
            k_classify<<<total_tiles, CTA, 0, stream>>>(P);
            count_launch(2);
            if (evp) D2R_CUDA(cudaEventRecord(evp->first, stream));   // time the march kernel alone
            if (use_split && (split_forced || need >= SPLIT_MIN_RAYS)) {
                if (need > s.cap_split) {
                    cudaFree(s.sp_feat); cudaFree(s.sp_aux); cudaFree(s.sp_shb); cudaFree(s.sp_nsb); cudaFree(s.sp_t);
                    cudaFree(s.sp_live[0]); cudaFree(s.sp_live[1]);
                    s.sp_feat = nullptr; s.sp_aux = nullptr; s.sp_shb = nullptr; s.sp_nsb = nullptr; s.sp_t = nullptr;
                    s.sp_live[0] = s.sp_live[1] = nullptr;
                    s.cap_split = 0;       // a failed allocation below must not leave a stale capacity behind
                    const size_t blocks = (need + 127) / 128;
                    D2R_CUDA(cudaMalloc(&s.sp_feat, blocks * 2 * SPLIT_TILE_BYTES));
                    D2R_CUDA(cudaMalloc(&s.sp_aux, blocks * 2 * 128 * sizeof(float2)));
                    D2R_CUDA(cudaMalloc(&s.sp_shb, need * 2 * sizeof(uint4)));
                    D2R_CUDA(cudaMalloc(&s.sp_nsb, blocks * 128));
                    D2R_CUDA(cudaMalloc(&s.sp_t, need * sizeof(float)));
                    D2R_CUDA(cudaMalloc(&s.sp_live[0], need * sizeof(uint32_t)));
                    D2R_CUDA(cudaMalloc(&s.sp_live[1], need * sizeof(uint32_t)));
                    s.cap_split = need;
                }
                if (!s.sp_cnt) D2R_CUDA(cudaMalloc(&s.sp_cnt, (SPLIT_MAX_ROUNDS + 2) * sizeof(uint32_t)));
                D2R_CUDA(cudaMemsetAsync(s.sp_cnt, 0, (SPLIT_MAX_ROUNDS + 2) * sizeof(uint32_t), stream));
                static bool split_attr[16] = {false};
                if (!split_attr[m->device]) {
                    D2R_CUDA(cudaFuncSetAttribute(k_mlp_round<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_TOTAL));
                    D2R_CUDA(cudaFuncSetAttribute(k_mlp_round<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_TOTAL));
                    split_attr[m->device] = true;
                }
                SplitParams Q;
                Q.feat = s.sp_feat; Q.aux = s.sp_aux; Q.shb = s.sp_shb; Q.nsb = s.sp_nsb; Q.t_cur = s.sp_t;
                for (int r = 0; r < SPLIT_MAX_ROUNDS; ++r) {
                    Q.round = r;
                    Q.cnt_in = s.sp_cnt + r; Q.cnt_out = s.sp_cnt + r + 1;
                    Q.live_in = s.sp_live[r & 1]; Q.live_out = s.sp_live[(r + 1) & 1];
                    if (split_coop) k_gather_round<true, 6><<<s.n_sm * std::min(split_gctas, 6), 128, 0, stream>>>(P, Q);
                    else if (split_gctas >= 8) k_gather_round<false, 8><<<s.n_sm * 8, 128, 0, stream>>>(P, Q);
                    else k_gather_round<false, 7><<<s.n_sm * split_gctas, 128, 0, stream>>>(P, Q);
                    if (split_mlp_old) k_mlp_round<false><<<s.n_sm * 4, TC_THREADS, T2_TOTAL, stream>>>(P, Q);
                    else k_mlp_round<true><<<s.n_sm * 4, TC_THREADS, T2_TOTAL, stream>>>(P, Q);
                    count_launch(2);
                    if (r >= 7 && (r & 3) == 3) {       // every 4th round from round 7 on: is anything left?  (one 4-byte read-back)
                        uint32_t left = 0;
                        D2R_CUDA(cudaMemcpyAsync(&left, s.sp_cnt + r + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
                        D2R_CUDA(cudaStreamSynchronize(stream));
                        if (!left) break;
                    }
                }
            }
            else if (use_tc1) k_march_tc<<<s.n_sm * 4, TC_THREADS, TS_TOTAL, stream>>>(P);
            else if (use_lpi4) k_march_tc2<4, true><<<s.n_sm * ctas_per_sm, TC_THREADS, T2_TOTAL, stream>>>(P);
            else if (abl == 1) k_march_tc2<2, true, 1><<<s.n_sm * ctas_per_sm, TC_THREADS, T2_TOTAL, stream>>>(P);
            else if (abl == 2) k_march_tc2<2, true, 2><<<s.n_sm * ctas_per_sm, TC_THREADS, T2_TOTAL, stream>>>(P);
            else if (use_solo4) k_march_tc2<4, false><<<s.n_sm * ctas_per_sm, TC_THREADS, T2_TOTAL, stream>>>(P);
            else if (use_solo) k_march_tc2<2, false><<<s.n_sm * ctas_per_sm, TC_THREADS, T2_TOTAL, stream>>>(P);
            else k_march_tc2<2, true><<<s.n_sm * ctas_per_sm, TC_THREADS, T2_TOTAL, stream>>>(P);
            if (evp) { D2R_CUDA(cudaEventRecord(evp->second, stream)); evp = nullptr; }
            k_finish<<<s.n_sm * 8, 256, 0, stream>>>(P);
            count_launch(2);
        }
    } else {
        k_march<<<s.n_sm * 3, CTA, SMEM_BYTES, stream>>>(P);
        count_launch();
    }
    if (evp) D2R_CUDA(cudaEventRecord(evp->second, stream));
    D2R_CUDA(cudaGetLastError());
    return D2R_OK;
}

}  // namespace d2r

extern "C" int d2r_profile_enable(int device, int on) {
    D2R_REQUIRE(device >= 0 && device < 16, "d2r_profile_enable: bad device");
    d2r::Prof& pf = d2r::g_prof[device];
    D2R_CUDA(cudaSetDevice(device));
    if (on && !pf.counters) D2R_CUDA(cudaMalloc(&pf.counters, 4 * sizeof(unsigned long long)));
    if (pf.counters) D2R_CUDA(cudaMemset(pf.counters, 0, 4 * sizeof(unsigned long long)));
    pf.used = 0;
    pf.on = on != 0;
    return D2R_OK;
}

extern "C" int d2r_profile_read(int device, float* march_ms_total, int* n_launches, unsigned long long* n_samples,
                                unsigned long long* n_tiles) {
    D2R_REQUIRE(device >= 0 && device < 16 && march_ms_total && n_launches && n_samples && n_tiles, "d2r_profile_read: bad argument");
    d2r::Prof& pf = d2r::g_prof[device];
    D2R_CUDA(cudaSetDevice(device));
    float total = 0.f;
    for (size_t i = 0; i < pf.used; ++i) {
        D2R_CUDA(cudaEventSynchronize(pf.ev[i].second));
        float ms = 0.f;
        D2R_CUDA(cudaEventElapsedTime(&ms, pf.ev[i].first, pf.ev[i].second));
        total += ms;
    }
    unsigned long long c[2] = {0, 0};
    if (pf.counters) D2R_CUDA(cudaMemcpy(c, pf.counters, sizeof(c), cudaMemcpyDeviceToHost));
    *march_ms_total = total;
    *n_launches = (int)pf.used;
    *n_samples = c[0];
    *n_tiles = c[1];
    return D2R_OK;
}

extern "C" int d2r_render(const d2r_model* m, const d2r_view* v, const float* cams_ngp_host, int K, const float background_rgba[4],
                          float* rgba_out_dev, float* depth_out_dev, unsigned long long* n_samples_out_dev, void* stream) {
    return d2r::launch_march(m, v, cams_ngp_host, K, background_rgba, rgba_out_dev, depth_out_dev, nullptr, nullptr, nullptr,
                             n_samples_out_dev, (cudaStream_t)stream);
}

extern "C" int d2r_render_composite_ex(const d2r_model* fg, const d2r_view* v, const float* cams_ngp_host, int K,
                                       const float fg_background_rgba[4], const float* bg_rgba_dev, const float* bg_depth_dev,
                                       uint8_t* rgb_u8_out_dev, int* rects_out_dev, uint8_t* bg_u8_out_dev,
                                       unsigned long long* n_samples_out_dev, void* stream) {
    if (!rgb_u8_out_dev) { d2r::set_error("d2r_render_composite_ex: rgb_u8_out_dev is null"); return D2R_ERR_INVALID; }
    return d2r::launch_march(fg, v, cams_ngp_host, K, fg_background_rgba, nullptr, nullptr, bg_rgba_dev, bg_depth_dev, rgb_u8_out_dev,
                             n_samples_out_dev, (cudaStream_t)stream, rects_out_dev, bg_u8_out_dev);
}

extern "C" int d2r_render_composite(const d2r_model* fg, const d2r_view* v, const float* cams_ngp_host, int K,
                                    const float fg_background_rgba[4], const float* bg_rgba_dev, const float* bg_depth_dev,
                                    uint8_t* rgb_u8_out_dev, unsigned long long* n_samples_out_dev, void* stream) {
    if (!rgb_u8_out_dev) { d2r::set_error("d2r_render_composite: rgb_u8_out_dev is null"); return D2R_ERR_INVALID; }
    return d2r::launch_march(fg, v, cams_ngp_host, K, fg_background_rgba, nullptr, nullptr, bg_rgba_dev, bg_depth_dev, rgb_u8_out_dev,
                             n_samples_out_dev, (cudaStream_t)stream);
}
