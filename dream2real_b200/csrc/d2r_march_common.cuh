// Shared pieces of the ray-march kernels (sm_100a): launch parameters, the hit list, primary-ray set-up, pass 1
// (k_classify) and pass 3 (k_finish), the depth-test composite, UMMA operand helpers.
// Included by d2r_march.cu; the march kernels proper live in d2r_march_ws.cuh (default) and d2r_march_split.cuh.
#pragma once
#include "d2r_gemm.cuh"
#include "d2r_march.cuh"

namespace d2r {

constexpr int TILE_W = 16, TILE_H = 8, CTA = TILE_W * TILE_H;   // 128 threads = 128 rays
constexpr int MARCH_ITER = 10000;                                // NGP src/testbed_nerf.cu:59

struct RayEntry {      // one primary ray that found an occupied sample (k_classify -> k_march_ws)
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
    float bg[4];                   // Testbed.background_color of the rendered model (sRGB + alpha)
    float4* rgba_out;              // [K,H,W] or null   (Shade)
    float4* depth_out;             // [K,H,W] or null   (Depth)
    const float4* bg_rgba;         // composite mode: cached background render [H,W]
    const float* bg_depth;         //                 cached background depth  [H,W]
    uint8_t* u8_out;               //                 [K,H,W,3]
    unsigned long long* n_samples;
    unsigned long long* prof;      // profiling counters {samples, primary rays owned, work items} or null
    RayEntry* entries;             // hit list
    uint32_t* n_entries;           //   number of entries (device)
    uint32_t* entry_cursor;        //   consumption cursor (device)
    float4* res_rgbd;              //   per-entry accumulated (r, g, b, depth), written by the march kernel
    float* res_a;                  //   per-entry accumulated alpha
    // k_march_ws work: entry ids to march (null = every hit-list entry) and how many; resume_steps > 0: the rays come out of
    // `resume_steps / 2` rounds of the split kernels -- ray parameter from t_cur (by entry id), accumulators from resume_acc4 /
    // resume_acca (by position in work_list)
    const uint32_t* work_list;
    const uint32_t* work_count;
    const float* t_cur;
    const float4* resume_acc4;
    const float* resume_acca;
    int resume_steps;
    uint32_t split_cap;            //   the split rounds only ran if the launch has at most this many hits
    uint32_t* feedback;            // host-mapped {hits, hit-list slots asked for} of this launch, written by k_finish, or null
    uint32_t feedback_slots;
    float* res_n;                  //   per-entry step count as the reference's payload.n_steps ends up (Cost render mode), or null
    float* cost_out;               // [K,H,W] or null   (Cost: n_steps of the rays the reference keeps, 0 elsewhere)
};

__device__ __forceinline__ float h2f_round(float v) { return __half2float(__float2half_rn(v)); }

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

constexpr int TC_THREADS = 128;
constexpr int TC_CHUNK = 256;             // hit-list entries a gather group claims at a time

// MLP weights as UMMA B operands (K-major, no swizzle), one contiguous 20 KB blob: built once per model on the host
// (d2r_model_load) and copied into shared memory with ONE cp.async.bulk per CTA
constexpr int W_D0 = 0;                    // [64 x 32]  fp16, SBO 512
constexpr int W_D1 = W_D0 + 4096;          // [16 x 64]        SBO 1024
constexpr int W_C0 = W_D1 + 2048;          // [64 x 32]
constexpr int W_C1 = W_C0 + 4096;          // [64 x 64]
constexpr int W_C2 = W_C1 + 8192;          // [16 x 64]
constexpr int W_BYTES = W_C2 + 2048;       // 20480

// byte offset of the 16-byte chunk (row r, columns 8*kc .. 8*kc+7) of a K-major no-swizzle operand with K columns
__device__ __forceinline__ uint32_t umma_chunk_off(int r, int kc, int K) {
    return (uint32_t)((r >> 3) * (K / 8) * 128 + kc * 128 + (r & 7) * 16);
}

__device__ __forceinline__ uint32_t pack_relu_h2(uint32_t a, uint32_t b, bool relu) {
    uint32_t d;
    // round-then-ReLU == ReLU-then-round; cvt.rn.relu.f16x2.f32 does both in one instruction (upper half <- first source)
    if (relu) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
    return d;
}

// origin, normalised direction and its reciprocal only: the first part of setup_ray, same arithmetic (the hit list
// already holds what the box / sphere tests produce)
__device__ __forceinline__ void ray_geom_only(const Mat3x4& C, float2 dc, RayGeom& r) {
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

// primary-ray set-up shared by k_classify and the slot refill: init_rays_with_payload_kernel_nerf
// (NGP testbed_nerf.cu:1394-1482).  Returns false when the ray can never take a sample.
__device__ __forceinline__ bool setup_ray(const ModelDev& M, const Mat3x4& C, float2 dc, RayGeom& r, float& t, float& t_box) {
    float vx = 0.f, vy = 0.f, vz = 0.f;   // mat3(camera) * (dc.x, dc.y, 1): tcnn accumulates column by column
    vx += C.c[0][0] * dc.x; vy += C.c[0][1] * dc.x; vz += C.c[0][2] * dc.x;
    vx += C.c[1][0] * dc.y; vy += C.c[1][1] * dc.y; vz += C.c[1][2] * dc.y;
    vx += C.c[2][0] * 1.0f; vy += C.c[2][1] * 1.0f; vz += C.c[2][2] * 1.0f;
    float len2 = 0.f;
    len2 += vx * vx; len2 += vy * vy; len2 += vz * vz;
    const float len = sqrtf(len2);
    r.dx = vx / len; r.dy = vy / len; r.dz = vz / len;
    r.ox = C.c[3][0]; r.oy = C.c[3][1]; r.oz = C.c[3][2];
    r.ix = 1.0f / r.dx; r.iy = 1.0f / r.dy; r.iz = 1.0f / r.dz;
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
    if (!raabb_contains(M, r.ox + t * r.dx, r.oy + t * r.dy, r.oz + t * r.dz)) return false;
    // rays that miss the box around all occupied cells can never take a sample
    const float2 oc = box_ray_intersect(M.occ_min, M.occ_max, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz);
    if (oc.x > 1e37f || oc.y < 0.f) return false;
    t_box = oc.x;
    r.t_exit = oc.y;
    // ... and rays that miss the bounding sphere of the occupied cells; the chord through box AND sphere bounds the walk
    {
        const float cx = r.ox - M.occ_ctr[0], cy = r.oy - M.occ_ctr[1], cz = r.oz - M.occ_ctr[2];
        const float b = cx * r.dx + cy * r.dy + cz * r.dz;            // |d| = 1
        const float disc = b * b - (cx * cx + cy * cy + cz * cz - M.occ_r2);
        if (disc < 0.f) return false;
        const float sq = sqrtf(disc);
        if (-b + sq < 0.f) return false;
        t_box = fmaxf(t_box, -b - sq);
        r.t_exit = fminf(r.t_exit, -b + sq);
    }
    return true;
}

// Pass 1: one thread per pixel of every candidate's screen rectangle (16x8 tiles).  Generates the ray,
// applies the Sobol start jitter (advance_pos_nerf, testbed_nerf.cu:333-362) and walks it to its first
// occupied sample.  Rays that find one are appended (warp-aggregated) to the hit list; all others leave
// their pixel as the fill kernel wrote it (the background).
__global__ void __launch_bounds__(128) k_classify(const __grid_constant__ MarchParams P) {
    const ModelDev& M = P.M;
    const int tid = threadIdx.x;
    const uint32_t tile = blockIdx.x;
    const int k = (int)P.tile_cand[tile];
    const int4 bb = P.bbox[k];
    const uint32_t local = tile - P.tile_prefix[k];
    const int tiles_x = (bb.z - bb.x + TILE_W) / TILE_W;
    const int x = bb.x + (int)(local % tiles_x) * TILE_W + (tid % TILE_W);
    const int y = bb.y + (int)(local / tiles_x) * TILE_H + (tid / TILE_W);
    bool hit = false;
    RayEntry e;
    if (x <= bb.z && y <= bb.w) {
        const uint32_t idx = (uint32_t)x + (uint32_t)P.W * (uint32_t)y;
        const Mat3x4 C = P.cams[k];
        RayGeom r;
        float t, t_box;
        if (setup_ray(M, C, __ldg(P.dirs + idx), r, t, t_box)) {
            const StepC cone = make_stepc(M.cone);
            t = advance_n_steps(t, cone, ld_random_val0(idx * 786433u));
            t = fast_forward_to_box(t, cone, t_box);
            t = skip_to_occupied(t, cone, r, M);
            if (t < MAX_DEPTH()) { hit = true; e.k = (uint32_t)k; e.idx = idx; e.t = t; e.t_exit = r.t_exit; }
        }
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, hit);
    if (mask) {
        const int lane = tid & 31, leader = __ffs(mask) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(P.n_entries, (uint32_t)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (hit) P.entries[base + __popc(mask & ((1u << lane) - 1))] = e;
    }
}

// finish a ray: keep rule (a > 0.001), shade / tonemap background blend, outputs -- compact_kernel_nerf +
// shade_kernel_nerf + tonemap_kernel (NGP testbed_nerf.cu:1302-1367, render_buffer.cu:529-561), then the
// depth-test composite when a u8 frame is requested.
__device__ __forceinline__ void finish_ray(const MarchParams& P, float cr, float cg, float cb, float cd, float ca, float n_steps, uint32_t idx, uint32_t k) {
    if (!(ca > 0.001f)) { cr = cg = cb = cd = ca = 0.f; n_steps = 0.f; }
    float4 shade = make_float4(srgb_to_linear_d(cr), srgb_to_linear_d(cg), srgb_to_linear_d(cb), ca);
    float4 depth = make_float4(cd, cd, cd, ca);
    const float w = (1.f - ca) * P.bg[3];
    const float blr = srgb_to_linear_d(P.bg[0]), blg = srgb_to_linear_d(P.bg[1]), blb = srgb_to_linear_d(P.bg[2]);
    shade.x += blr * w; shade.y += blg * w; shade.z += blb * w; shade.w += w;
    depth.x += blr * w; depth.y += blg * w; depth.z += blb * w; depth.w += w;
    const size_t o = (size_t)k * ((size_t)P.W * P.H) + idx;
    if (P.rgba_out) P.rgba_out[o] = shade;
    if (P.depth_out) P.depth_out[o] = depth;
    if (P.cost_out) P.cost_out[o] = n_steps;
    if (P.u8_out) composite_pixel(shade, depth.x, __ldg(P.bg_rgba + idx), __ldg(P.bg_depth + idx), P.u8_out + o * 3);
}

// Pass 3: one thread per hit-list entry, every lane busy.
__global__ void __launch_bounds__(256) k_finish(const __grid_constant__ MarchParams P) {
    const uint32_t n = *P.n_entries;
    if (P.feedback && blockIdx.x == 0 && threadIdx.x == 0) { P.feedback[1] = P.feedback_slots; P.feedback[0] = n; __threadfence_system(); }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const RayEntry e = P.entries[i];
        const float4 c = P.res_rgbd[i];
        finish_ray(P, c.x, c.y, c.z, c.w, P.res_a[i], P.res_n ? P.res_n[i] : 0.f, e.idx, e.k);
    }
}

// plain (non-tensor) bulk copy global -> shared memory of this CTA, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// D[tmem] (+)= A . B^T with the two smem descriptors passed as 32-bit halves (no 64-bit arithmetic on the issuing thread's
// critical path): lo = (address >> 4) | LBO field, hi = SBO field | version
template <bool ACC>
__device__ __forceinline__ void umma_f16_ss_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
// one MLP layer for ONE 128-row tile: K/16 MMAs, fully unrolled
template <int K, int N>
__device__ __forceinline__ void issue_tile_ws(uint32_t a_lo, uint32_t b_lo, uint32_t tmem_d) {
    constexpr uint32_t hi = (uint32_t)((K / 8) * 128 >> 4) | (1u << 14);      // SBO | descriptor version 1 (bit 46)
    constexpr uint32_t idesc = umma_idesc_f16(128, N, 0);
    umma_f16_ss_lohi<false>(tmem_d, a_lo, b_lo, hi, idesc);
#pragma unroll
    for (int kk = 1; kk < K / 16; ++kk) umma_f16_ss_lohi<true>(tmem_d, a_lo + kk * 16, b_lo + kk * 16, hi, idesc);
}
// one MLP layer for both sample tiles of a group: 2 x K/16 MMAs, fully unrolled
template <int K, int N>
__device__ __forceinline__ void issue_layer_ws(uint32_t a_lo, uint32_t b_lo, uint32_t tmem_d) {
    constexpr uint32_t hi = (uint32_t)((K / 8) * 128 >> 4) | (1u << 14);      // SBO | descriptor version 1 (bit 46)
    constexpr uint32_t idesc = umma_idesc_f16(128, N, 0);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        umma_f16_ss_lohi<false>(tmem_d + s * 64, a_lo + s * 1024, b_lo, hi, idesc);
#pragma unroll
        for (int kk = 1; kk < K / 16; ++kk) umma_f16_ss_lohi<true>(tmem_d + s * 64, a_lo + s * 1024 + kk * 16, b_lo + kk * 16, hi, idesc);
    }
}

}  // namespace d2r
