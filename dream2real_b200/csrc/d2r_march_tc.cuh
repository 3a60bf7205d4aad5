// k_march_tc: the ray-march kernel with the two NeRF MLPs on the 5th-generation tensor cores.
// Included by d2r_march.cu (shares MarchParams, composite_pixel and the device maths of d2r_march.cuh).
//
// One persistent CTA = 128 ray slots = the 128 rows (TMEM lanes) of an M=128 UMMA tile.  Every round
// each live slot advances its ray to the next occupied sample, gathers the 8-level hash-grid features
// (fp16 fma chain like the reference) and stores them as one row of the A operand in shared memory
// (canonical K-major no-swizzle UMMA layout).  One elected thread then issues the five layers
//     32->64 (ReLU) -> 16 | [16 density-out, 16 SH] -> 64 (ReLU) -> 64 (ReLU) -> 16(3)
// as tcgen05.mma.cta_group::1.kind::f16 (M=128, N=64|16, K=16 per instruction) with the weights as
// the K-major B operand, fp32 accumulators in TMEM; every slot reads its own row back with tcgen05.ld
// (lane == thread), applies ReLU, rounds to fp16 (the reference keeps fp16 activations, TCNN
// fully_fused_mlp.cu:47-129) and writes the next layer's A row.  Compositing stays in registers.
// Slots whose ray finished pull the next entry of the global hit list (written by k_classify, which
// generates the rays, applies the Sobol jitter and walks each one to its first occupied sample), so the
// tile stays dense without per-round global-memory compaction or host round trips (the reference
// compacts through global memory and syncs with the host every round, NGP testbed_nerf.cu:1664-1748).
#pragma once

namespace d2r {

constexpr int TC_THREADS = 128;
constexpr int TC_CHUNK = 256;             // hit-list entries a CTA claims at a time

// shared memory plan (bytes); UMMA operand tiles want 16-byte alignment without swizzle, we give them 128
constexpr int TS_WD0 = 0;                  // [64 x 32]  fp16, SBO 512
constexpr int TS_WD1 = TS_WD0 + 4096;      // [16 x 64]        SBO 1024
constexpr int TS_WC0 = TS_WD1 + 2048;      // [64 x 32]
constexpr int TS_WC1 = TS_WC0 + 4096;      // [64 x 64]
constexpr int TS_WC2 = TS_WC1 + 8192;      // [16 x 64]
constexpr int TS_A32 = TS_WC2 + 2048;      // [128 x 32] activations, K = 32
constexpr int TS_A64 = TS_A32 + 8192;      // [128 x 64] activations, K = 64
constexpr int TS_MISC = TS_A64 + 16384;    // mbarrier, tmem slot, cursor, item, camera
constexpr int TS_TOTAL = TS_MISC + 256;

// byte offset of the 16-byte chunk (row r, columns 8*kc .. 8*kc+7) of a K-major no-swizzle operand with K columns
__device__ __forceinline__ uint32_t umma_chunk_off(int r, int kc, int K) {
    return (uint32_t)((r >> 3) * (K / 8) * 128 + kc * 128 + (r & 7) * 16);
}

// stage W[N][K] (row-major fp16, global) as a UMMA B operand
__device__ __forceinline__ void stage_weights(unsigned char* dst, const __half* __restrict__ W, int N, int K, int tid) {
    const int chunks = N * (K / 8);
    for (int i = tid; i < chunks; i += TC_THREADS) {
        const int n = i / (K / 8), kc = i % (K / 8);
        const uint4 v = *reinterpret_cast<const uint4*>(W + (size_t)n * K + kc * 8);
        *reinterpret_cast<uint4*>(dst + umma_chunk_off(n, kc, K)) = v;
    }
}

// D[tmem_col .. +N) = A[128 x K] . B[N x K]^T, then arrive on `bar` when the MMAs have retired
__device__ __forceinline__ void issue_layer(uint32_t a_addr, uint32_t b_addr, int K, int N, uint32_t tmem_d, uint64_t* bar) {
    const uint32_t idesc = umma_idesc_f16(128, N, /*fp16*/ 0);
    const uint32_t sbo = (uint32_t)(K / 8) * 128;
    for (int kk = 0; kk < K / 16; ++kk) {
        const uint64_t da = umma_desc_noswz(a_addr + kk * 256, 128, sbo);
        const uint64_t db = umma_desc_noswz(b_addr + kk * 256, 128, sbo);
        umma_f16_ss(tmem_d, da, db, idesc, kk > 0);
    }
    tc_commit(bar);
}

__device__ __forceinline__ uint32_t pack_relu_h2(uint32_t a, uint32_t b, bool relu) {
    uint32_t d;
    // round-then-ReLU == ReLU-then-round; cvt.rn.relu.f16x2.f32 does both in one instruction (upper half <- first source)
    if (relu) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
    return d;
}

struct TcRay {
    RayGeom g;
    float t;
    float cr, cg, cb, cd, ca;
    float fwx, fwy, fwz;  // camera forward axis of this slot's candidate (depth compositing)
    uint32_t ei;          // hit-list index of the ray in this slot
    int n_steps;
    uint32_t sh[8];      // 16 fp16 SH coefficients
};

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
__device__ __forceinline__ void finish_ray(const MarchParams& P, float cr, float cg, float cb, float cd, float ca, uint32_t idx, uint32_t k) {
    if (!(ca > 0.001f)) { cr = cg = cb = cd = ca = 0.f; }
    float4 shade = make_float4(srgb_to_linear_d(cr), srgb_to_linear_d(cg), srgb_to_linear_d(cb), ca);
    float4 depth = make_float4(cd, cd, cd, ca);
    const float w = (1.f - ca) * P.bg[3];
    const float blr = srgb_to_linear_d(P.bg[0]), blg = srgb_to_linear_d(P.bg[1]), blb = srgb_to_linear_d(P.bg[2]);
    shade.x += blr * w; shade.y += blg * w; shade.z += blb * w; shade.w += w;
    depth.x += blr * w; depth.y += blg * w; depth.z += blb * w; depth.w += w;
    const size_t o = (size_t)k * ((size_t)P.W * P.H) + idx;
    if (P.rgba_out) P.rgba_out[o] = shade;
    if (P.depth_out) P.depth_out[o] = depth;
    if (P.u8_out) composite_pixel(shade, depth.x, __ldg(P.bg_rgba + idx), __ldg(P.bg_depth + idx), P.u8_out + o * 3);
}

// Pass 3: one thread per hit-list entry, every lane busy.
__global__ void __launch_bounds__(256) k_finish(const __grid_constant__ MarchParams P) {
    const uint32_t n = *P.n_entries;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const RayEntry e = P.entries[i];
        const float4 c = P.res_rgbd[i];
        finish_ray(P, c.x, c.y, c.z, c.w, P.res_a[i], e.idx, e.k);
    }
}

__global__ void __launch_bounds__(TC_THREADS, 4) k_march_tc(const __grid_constant__ MarchParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const ModelDev& M = P.M;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + TS_MISC);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TS_MISC + 8);
    uint32_t* s_cursor = reinterpret_cast<uint32_t*>(smem + TS_MISC + 12);
    uint32_t* s_end = reinterpret_cast<uint32_t*>(smem + TS_MISC + 16);
    uint32_t* s_done = reinterpret_cast<uint32_t*>(smem + TS_MISC + 20);

    stage_weights(smem + TS_WD0, M.w_d0, 64, 32, tid);
    stage_weights(smem + TS_WD1, M.w_d1, 16, 64, tid);
    stage_weights(smem + TS_WC0, M.w_c0, 64, 32, tid);
    stage_weights(smem + TS_WC1, M.w_c1, 64, 64, tid);
    stage_weights(smem + TS_WC2, M.w_c2, 16, 64, tid);
    if (tid == 0) { mbar_init(mbar, 1); fence_barrier_init(); *s_cursor = 0; *s_end = 0; *s_done = 0; }
    if (warp == 0) tmem_alloc<128>(tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);    // this warp's 32 lanes
    const uint32_t a32 = smem_u32(smem + TS_A32), a64 = smem_u32(smem + TS_A64);
    const uint32_t wd0 = smem_u32(smem + TS_WD0), wd1 = smem_u32(smem + TS_WD1), wc0 = smem_u32(smem + TS_WC0),
                   wc1 = smem_u32(smem + TS_WC1), wc2 = smem_u32(smem + TS_WC2);
    unsigned char* rowA32 = smem + TS_A32 + umma_chunk_off(tid, 0, 32);
    unsigned char* rowA64 = smem + TS_A64 + umma_chunk_off(tid, 0, 64);
    uint32_t phase = 0;
    const uint32_t total_entries = *P.n_entries;
    const StepC cone = make_stepc(M.cone);
    unsigned long long my_samples = 0, my_rays = 0;

    bool alive = false;
    TcRay R;
    // a finished ray only parks its accumulators (20 B, slot = its hit-list index); k_finish turns them into
    // pixels afterwards with every lane busy (inline, this epilogue ran with ~2 of 32 lanes active and cost
    // 13 % of the kernel's issue slots)
    auto finish = [&](float cr, float cg, float cb, float cd, float ca, uint32_t ei) {
        P.res_rgbd[ei] = make_float4(cr, cg, cb, cd);
        P.res_a[ei] = ca;
    };

    {
        while (true) {
            // ---- A. claim hit-list entries: thread 0 fetches a new chunk when the current one is used up ----
            if (tid == 0 && *s_cursor >= *s_end && !*s_done) {
                const uint32_t base = atomicAdd(P.entry_cursor, (uint32_t)TC_CHUNK);
                if (base >= total_entries) { *s_done = 1; }
                else { *s_cursor = base; *s_end = min(base + (uint32_t)TC_CHUNK, total_entries); }
            }
            __syncthreads();
            const bool done = *s_done != 0;   // stable until every thread has passed this round's second barrier
            if (!alive) {
                const uint32_t i = atomicAdd(s_cursor, 1u);
                if (i < *s_end) {
                    const RayEntry e = P.entries[i];
                    const Mat3x4 C = P.cams[e.k];
                    float t0, t_box;
                    setup_ray(M, C, __ldg(P.dirs + e.idx), R.g, t0, t_box);   // same arithmetic as pass 1
                    R.g.t_exit = e.t_exit;
                    R.t = e.t; R.ei = i; R.n_steps = 0;
                    R.fwx = C.c[2][0]; R.fwy = C.c[2][1]; R.fwz = C.c[2][2];
                    R.cr = R.cg = R.cb = R.cd = R.ca = 0.f;
                    float sh[16];
                    const float wx = (R.g.dx + 1.0f) * 0.5f, wy = (R.g.dy + 1.0f) * 0.5f, wz = (R.g.dz + 1.0f) * 0.5f;
                    sh_enc4(wx * 2.f - 1.f, wy * 2.f - 1.f, wz * 2.f - 1.f, sh);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        __half2 h = __floats2half2_rn(sh[2 * j], sh[2 * j + 1]);
                        R.sh[j] = *reinterpret_cast<uint32_t*>(&h);
                    }
                    alive = true;
                    ++my_rays;
                }
            }
            // ---- B. next sample position + hash-grid features -> A row ----
            bool has_sample = false;
            float wpx = 0.f, wpy = 0.f, wpz = 0.f, wdt = 0.f;
            if (alive) {
                const float t = skip_to_occupied(R.t, cone, R.g, M);     // generate_next_nerf_network_inputs (:454-467)
                if (t >= MAX_DEPTH()) {
                    finish(R.cr, R.cg, R.cb, R.cd, R.ca, R.ei);
                    alive = false;
                } else {
                    const float dt = calc_dt(t, cone);
                    const float px = R.g.ox + R.g.dx * t, py = R.g.oy + R.g.dy * t, pz = R.g.oz + R.g.dz * t;
                    wpx = (px - M.aabb_min[0]) / M.aabb_diag[0];
                    wpy = (py - M.aabb_min[1]) / M.aabb_diag[1];
                    wpz = (pz - M.aabb_min[2]) / M.aabb_diag[2];
                    wdt = warp_dt(dt);
                    R.t = t + dt;
                    has_sample = true;
                    ++my_samples;
#pragma unroll 1   // keep the kernel inside the instruction cache: 4 CTAs per SM sit at different PCs
                    for (int c = 0; c < 4; ++c) {
                        __half2 f0, f1, f2, f3;
                        encode_level(M, 2 * c, wpx, wpy, wpz, f0, f1);
                        encode_level(M, 2 * c + 1, wpx, wpy, wpz, f2, f3);
                        uint4 v;
                        v.x = *reinterpret_cast<uint32_t*>(&f0); v.y = *reinterpret_cast<uint32_t*>(&f1);
                        v.z = *reinterpret_cast<uint32_t*>(&f2); v.w = *reinterpret_cast<uint32_t*>(&f3);
                        *reinterpret_cast<uint4*>(rowA32 + c * 128) = v;
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            if (!__syncthreads_or(has_sample ? 1 : 0)) {
                if (done) break;
                continue;
            }
            // ---- C. density layer 0: 32 -> 64, ReLU ----
            if (tid == 0) { tc_fence_after(); issue_layer(a32, wd0, 32, 64, tmem_base + 0, mbar); }
            mbar_wait(mbar, phase); phase ^= 1;
            tc_fence_after();
            {
                uint32_t r[64];
                tmem_ld_32x32_x64(tmem_lane + 0, r);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 v;
                    v.x = pack_relu_h2(r[8 * c + 0], r[8 * c + 1], true); v.y = pack_relu_h2(r[8 * c + 2], r[8 * c + 3], true);
                    v.z = pack_relu_h2(r[8 * c + 4], r[8 * c + 5], true); v.w = pack_relu_h2(r[8 * c + 6], r[8 * c + 7], true);
                    *reinterpret_cast<uint4*>(rowA64 + c * 128) = v;
                }
            }
            fence_proxy_async(); tc_fence_before(); __syncthreads();
            // ---- D. density layer 1: 64 -> 16 (row 0 = raw density), then rgb input = [16 density-out | 16 SH] ----
            if (tid == 0) { tc_fence_after(); issue_layer(a64, wd1, 64, 16, tmem_base + 64, mbar); }
            mbar_wait(mbar, phase); phase ^= 1;
            tc_fence_after();
            float sigma;
            {
                uint32_t r[16];
                tmem_ld_32x32_x16(tmem_lane + 64, r);
                tmem_ld_wait();
                sigma = h2f_round(__uint_as_float(r[0]));
                uint4 v0, v1, v2, v3;
                v0.x = pack_relu_h2(r[0], r[1], false); v0.y = pack_relu_h2(r[2], r[3], false);
                v0.z = pack_relu_h2(r[4], r[5], false); v0.w = pack_relu_h2(r[6], r[7], false);
                v1.x = pack_relu_h2(r[8], r[9], false); v1.y = pack_relu_h2(r[10], r[11], false);
                v1.z = pack_relu_h2(r[12], r[13], false); v1.w = pack_relu_h2(r[14], r[15], false);
                v2 = make_uint4(R.sh[0], R.sh[1], R.sh[2], R.sh[3]);
                v3 = make_uint4(R.sh[4], R.sh[5], R.sh[6], R.sh[7]);
                *reinterpret_cast<uint4*>(rowA32 + 0) = v0;
                *reinterpret_cast<uint4*>(rowA32 + 128) = v1;
                *reinterpret_cast<uint4*>(rowA32 + 256) = v2;
                *reinterpret_cast<uint4*>(rowA32 + 384) = v3;
            }
            fence_proxy_async(); tc_fence_before(); __syncthreads();
            // ---- E. rgb layer 0: 32 -> 64, ReLU ----
            if (tid == 0) { tc_fence_after(); issue_layer(a32, wc0, 32, 64, tmem_base + 0, mbar); }
            mbar_wait(mbar, phase); phase ^= 1;
            tc_fence_after();
            {
                uint32_t r[64];
                tmem_ld_32x32_x64(tmem_lane + 0, r);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 v;
                    v.x = pack_relu_h2(r[8 * c + 0], r[8 * c + 1], true); v.y = pack_relu_h2(r[8 * c + 2], r[8 * c + 3], true);
                    v.z = pack_relu_h2(r[8 * c + 4], r[8 * c + 5], true); v.w = pack_relu_h2(r[8 * c + 6], r[8 * c + 7], true);
                    *reinterpret_cast<uint4*>(rowA64 + c * 128) = v;
                }
            }
            fence_proxy_async(); tc_fence_before(); __syncthreads();
            // ---- F. rgb layer 1: 64 -> 64, ReLU ----
            if (tid == 0) { tc_fence_after(); issue_layer(a64, wc1, 64, 64, tmem_base + 64, mbar); }
            mbar_wait(mbar, phase); phase ^= 1;
            tc_fence_after();
            {
                uint32_t r[64];
                tmem_ld_32x32_x64(tmem_lane + 64, r);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 v;
                    v.x = pack_relu_h2(r[8 * c + 0], r[8 * c + 1], true); v.y = pack_relu_h2(r[8 * c + 2], r[8 * c + 3], true);
                    v.z = pack_relu_h2(r[8 * c + 4], r[8 * c + 5], true); v.w = pack_relu_h2(r[8 * c + 6], r[8 * c + 7], true);
                    *reinterpret_cast<uint4*>(rowA64 + c * 128) = v;
                }
            }
            fence_proxy_async(); tc_fence_before(); __syncthreads();
            // ---- G. rgb output layer: 64 -> 16 (3 used) ----
            if (tid == 0) { tc_fence_after(); issue_layer(a64, wc2, 64, 16, tmem_base + 0, mbar); }
            mbar_wait(mbar, phase); phase ^= 1;
            tc_fence_after();
            float raw0, raw1, raw2;
            {
                uint32_t r[16];
                tmem_ld_32x32_x16(tmem_lane + 0, r);
                tmem_ld_wait();
                raw0 = h2f_round(__uint_as_float(r[0])); raw1 = h2f_round(__uint_as_float(r[1])); raw2 = h2f_round(__uint_as_float(r[2]));
            }
            // ---- H. composite_kernel_nerf (testbed_nerf.cu:511-667) ----
            if (has_sample) {
                const float ux = M.aabb_min[0] + wpx * M.aabb_diag[0];
                const float uy = M.aabb_min[1] + wpy * M.aabb_diag[1];
                const float uz = M.aabb_min[2] + wpz * M.aabb_diag[2];
                const float T = 1.f - R.ca;
                const float dtu = unwarp_dt(wdt);
                const float alpha = 1.f - __expf(-__expf(sigma) * dtu);
                const float weight = alpha * T;
                const float rr = logistic_d(raw0), gg = logistic_d(raw1), bb_ = logistic_d(raw2);
                float dep = 0.f;
                dep += R.fwx * (ux - R.g.ox); dep += R.fwy * (uy - R.g.oy); dep += R.fwz * (uz - R.g.oz);
                dep *= M.depth_scale;
                R.cr += rr * weight; R.cg += gg * weight; R.cb += bb_ * weight; R.cd += dep * weight; R.ca += weight;
                if (R.ca > (1.0f - M.min_transmittance)) {
                    R.cr /= R.ca; R.cg /= R.ca; R.cb /= R.ca; R.cd /= R.ca; R.ca /= R.ca;
                    finish(R.cr, R.cg, R.cb, R.cd, R.ca, R.ei);
                    alive = false;
                } else if (++R.n_steps >= MARCH_ITER - 1) {
                    finish(0.f, 0.f, 0.f, 0.f, 0.f, R.ei);        // never reaches the hit buffer in the reference
                    alive = false;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<128>(tmem_base); }
    if (P.n_samples || P.prof) {
        for (int o = 16; o > 0; o >>= 1) my_samples += __shfl_xor_sync(0xffffffffu, my_samples, o);
        for (int o = 16; o > 0; o >>= 1) my_rays += __shfl_xor_sync(0xffffffffu, my_rays, o);
        if ((tid & 31) == 0) {
            if (P.n_samples && my_samples) atomicAdd(P.n_samples, my_samples);
            if (P.prof && my_samples) atomicAdd(P.prof, my_samples);
            if (P.prof && my_rays) atomicAdd(P.prof + 1, my_rays);
        }

    }
}

}  // namespace d2r
