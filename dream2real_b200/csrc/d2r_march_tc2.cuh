// k_march_tc2: k_march_tc with TWO consecutive samples of every live ray per round.
// Included by d2r_march.cu after d2r_march_tc.cuh (shares its helpers).
//
// The next sample position of a ray depends on the occupancy grid only, never on the network output,
// so a slot can prepare samples n and n+1 together (the reference prepares up to 8 per compaction round,
// NGP testbed_nerf.cu:1692-1694).  Each layer is then issued for two M=128 tiles (sample 0 rows / sample 1
// rows, separate TMEM columns) behind ONE commit + barrier, which halves the per-sample cost of the five
// serialised MMA round trips and doubles the gathers in flight per thread.  Compositing stays sequential:
// sample 1 is dropped when sample 0 already saturated the ray -- exactly what the reference does with the
// unused tail of its n_steps.
#pragma once

namespace d2r {

// shared memory plan (bytes)
constexpr int T2_WD0 = 0;
constexpr int T2_WD1 = T2_WD0 + 4096;
constexpr int T2_WC0 = T2_WD1 + 2048;
constexpr int T2_WC1 = T2_WC0 + 4096;
constexpr int T2_WC2 = T2_WC1 + 8192;
// The K=32 operands (hash features; density-out | SH) and the K=64 operands (hidden activations) alternate
// strictly -- each is dead once the MMA that reads it has retired, which is before the next one is written --
// so both live in the same 16 KB per tile.  52 KB per CTA -> 4 CTAs per SM.
constexpr int T2_A64 = T2_WC2 + 2048;          // 2 tiles x [128 x 64]
constexpr int T2_A32 = T2_A64;                 // 2 tiles x [128 x 32], aliased, same 16 KB tile pitch
constexpr int T2_TILE = 16384;
constexpr int T2_MISC = T2_A64 + 2 * T2_TILE;
constexpr int T2_TOTAL = T2_MISC + 256;

// LPI = hash-grid levels gathered per loop iteration (8 loads each in flight together): 2 or 4
// COOP = lane-pair cooperative gathers (encode_levels_pair)
template <int LPI, bool COOP, int ABL = 0>   // ABL: timing ablations (1: encode twice, 2: walk twice, 3: every MMA round trip twice)
__global__ void __launch_bounds__(TC_THREADS, 4) k_march_tc2(const __grid_constant__ MarchParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const ModelDev& M = P.M;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + T2_MISC);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T2_MISC + 8);
    uint32_t* s_cursor = reinterpret_cast<uint32_t*>(smem + T2_MISC + 12);
    uint32_t* s_end = reinterpret_cast<uint32_t*>(smem + T2_MISC + 16);
    uint32_t* s_done = reinterpret_cast<uint32_t*>(smem + T2_MISC + 20);

    stage_weights(smem + T2_WD0, M.w_d0, 64, 32, tid);
    stage_weights(smem + T2_WD1, M.w_d1, 16, 64, tid);
    stage_weights(smem + T2_WC0, M.w_c0, 64, 32, tid);
    stage_weights(smem + T2_WC1, M.w_c1, 64, 64, tid);
    stage_weights(smem + T2_WC2, M.w_c2, 16, 64, tid);
    if (tid == 0) { mbar_init(mbar, 1); fence_barrier_init(); *s_cursor = 0; *s_end = 0; *s_done = 0; }
    if (warp == 0) tmem_alloc<128>(tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t a32 = smem_u32(smem + T2_A32), a64 = smem_u32(smem + T2_A64);
    const uint32_t wd0 = smem_u32(smem + T2_WD0), wd1 = smem_u32(smem + T2_WD1), wc0 = smem_u32(smem + T2_WC0),
                   wc1 = smem_u32(smem + T2_WC1), wc2 = smem_u32(smem + T2_WC2);
    // tile s of this thread's rows: + s * T2_TILE; TMEM columns + s * 64
    unsigned char* rowA32 = smem + T2_A32 + umma_chunk_off(tid, 0, 32);
    unsigned char* rowA64 = smem + T2_A64 + umma_chunk_off(tid, 0, 64);
    uint32_t phase = 0;
    const uint32_t total_entries = *P.n_entries;
    const StepC cone = make_stepc(M.cone);
    unsigned long long my_samples = 0, my_rays = 0;

    bool alive = false;
    TcRay R;
    auto finish = [&](float cr, float cg, float cb, float cd, float ca, uint32_t ei) {
        P.res_rgbd[ei] = make_float4(cr, cg, cb, cd);
        P.res_a[ei] = ca;
    };
    // both tiles of one layer behind one commit
    auto issue2 = [&](uint32_t a_addr, uint32_t a_tile_bytes, uint32_t b_addr, int K, int N) {
        const uint32_t idesc = umma_idesc_f16(128, N, 0);
        const uint32_t sbo = (uint32_t)(K / 8) * 128;
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            for (int kk = 0; kk < K / 16; ++kk) {
                const uint64_t da = umma_desc_noswz(a_addr + s * a_tile_bytes + kk * 256, 128, sbo);
                const uint64_t db = umma_desc_noswz(b_addr + kk * 256, 128, sbo);
                umma_f16_ss(tmem_base + s * 64, da, db, idesc, kk > 0);
            }
        }
        tc_commit(mbar);
    };

    while (true) {
        // ---- A. claim hit-list entries ----
        if (tid == 0 && *s_cursor >= *s_end && !*s_done) {
            const uint32_t base = atomicAdd(P.entry_cursor, (uint32_t)TC_CHUNK);
            if (base >= total_entries) { *s_done = 1; }
            else { *s_cursor = base; *s_end = min(base + (uint32_t)TC_CHUNK, total_entries); }
        }
        __syncthreads();
        const bool done = *s_done != 0;
        if (!alive) {
            const uint32_t i = atomicAdd(s_cursor, 1u);
            if (i < *s_end) {
                const RayEntry e = P.entries[i];
                const Mat3x4 C = P.cams[e.k];
                float t0, t_box;
                setup_ray(M, C, __ldg(P.dirs + e.idx), R.g, t0, t_box);
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
        // ---- B. up to two sample positions + hash-grid features -> A rows of tile 0 / tile 1 ----
        // One walk loop for both samples (generate_next_nerf_network_inputs, testbed_nerf.cu:454-467, twice): the same
        // sequence of operations on t as two calls of if_unoccupied_advance_to_next_occupied_voxel, but lanes that need an
        // empty-voxel step for sample 0 and lanes that need one for sample 1 take it in the same warp iteration.
        int n_s = 0;                               // samples prepared this round (0, 1 or 2)
        bool exits = false;                        // the ray leaves the occupied region after its last prepared sample
        float wp0x = 0.f, wp0y = 0.f, wp0z = 0.f, wp1x = 0.f, wp1y = 0.f, wp1z = 0.f;
        float dep0 = 0.f, dep1 = 0.f, dtu0 = 0.f, dtu1 = 0.f;     // what compositing needs of a sample: depth and unwarped dt
#pragma unroll 1
        for (int rep = (ABL == 2 ? 0 : 1); rep < 2; ++rep)
        if (alive) {
            if (ABL == 2) { asm volatile("" :: "r"(n_s), "f"(dep0), "f"(dep1), "f"(wp0x), "f"(wp1x)); n_s = 0; exits = false; }
            float t = R.t;
            const uint32_t max_mip = (uint32_t)M.max_cascade;
            while (true) {
                const float px = R.g.ox + t * R.g.dx, py = R.g.oy + t * R.g.dy, pz = R.g.oz + t * R.g.dz;
                if (t >= MAX_DEPTH() || t > R.g.t_exit || !raabb_contains(M, px, py, pz)) { exits = true; break; }
                uint32_t mip = min(max(mip_from_pos(px, py, pz), 0u), max_mip);
                if (density_grid_occupied_at(px, py, pz, M.bitfield_lin, mip)) {
                    const float dt = calc_dt(t, cone);
                    const float wx = (px - M.aabb_min[0]) / M.aabb_diag[0];
                    const float wy = (py - M.aabb_min[1]) / M.aabb_diag[1];
                    const float wz = (pz - M.aabb_min[2]) / M.aabb_diag[2];
                    // composite_kernel_nerf reads the position back from the network input (unwarp_position) and
                    // the step from warp_dt/unwarp_dt: same arithmetic here, evaluated before the MLP instead of after
                    const float ux = M.aabb_min[0] + wx * M.aabb_diag[0];
                    const float uy = M.aabb_min[1] + wy * M.aabb_diag[1];
                    const float uz = M.aabb_min[2] + wz * M.aabb_diag[2];
                    float dep = 0.f;
                    dep += R.fwx * (ux - R.g.ox); dep += R.fwy * (uy - R.g.oy); dep += R.fwz * (uz - R.g.oz);
                    dep *= M.depth_scale;
                    const float dtu = unwarp_dt(warp_dt(dt));
                    if (n_s == 0) { wp0x = wx; wp0y = wy; wp0z = wz; dep0 = dep; dtu0 = dtu; }
                    else { wp1x = wx; wp1y = wy; wp1z = wz; dep1 = dep; dtu1 = dtu; }
                    t += dt;
                    if (++n_s == 2) break;
                    continue;
                }
                while (mip < max_mip && !density_grid_occupied_at(px, py, pz, M.bitfield_lin, mip + 1)) ++mip;
                t = advance_to_next_voxel(t, cone, px, py, pz, R.g.dx, R.g.dy, R.g.dz, R.g.ix, R.g.iy, R.g.iz, mip);
            }
            if (ABL == 2 && rep == 0) { asm volatile("" :: "f"(t)); continue; }
            R.t = t;
            if (n_s == 0) {                         // nothing left to sample: the ray is finished as it stands
                finish(R.cr, R.cg, R.cb, R.cd, R.ca, R.ei);
                alive = false;
            }
        }
        // hash-grid features: warp-converged (the cooperative gather shuffles between lanes); lanes without a sample
        // compute on the zero position and skip the store
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (__any_sync(0xffffffffu, s < n_s)) {
                const float sx = s ? wp1x : wp0x, sy = s ? wp1y : wp0y, sz = s ? wp1z : wp0z;
                float pe[3], po[3];            // the pair's two sample positions (lanes 2j / 2j+1), for the cooperative gather
                if (COOP || ABL == 1) {
                    const int lane_e = (tid & 31) & ~1, lane_o = (tid & 31) | 1;
                    pe[0] = __shfl_sync(0xffffffffu, sx, lane_e); pe[1] = __shfl_sync(0xffffffffu, sy, lane_e); pe[2] = __shfl_sync(0xffffffffu, sz, lane_e);
                    po[0] = __shfl_sync(0xffffffffu, sx, lane_o); po[1] = __shfl_sync(0xffffffffu, sy, lane_o); po[2] = __shfl_sync(0xffffffffu, sz, lane_o);
                }
#pragma unroll 1
                for (int c = 0; c < 8 / LPI; ++c) {
                    __half2 f[2 * LPI];
                    __half2 fd[2 * LPI];
                    if (ABL == 1) {
                        const float qe[3] = {pe[0] * 0.97f + 0.011f, pe[1] * 0.97f + 0.013f, pe[2] * 0.97f + 0.017f};
                        const float qo[3] = {po[0] * 0.97f + 0.011f, po[1] * 0.97f + 0.013f, po[2] * 0.97f + 0.017f};
                        encode_levels_pair<LPI>(M, LPI * c, sx * 0.97f + 0.011f, sy * 0.97f + 0.013f, sz * 0.97f + 0.017f, qe, qo, fd);
                    }
                    if (COOP) encode_levels_pair<LPI>(M, LPI * c, sx, sy, sz, pe, po, f);
                    else encode_levels<LPI>(M, LPI * c, sx, sy, sz, f);
                    if (ABL == 1) {
                        const __half2 z = __float2half2_rn(M.depth_scale * 0.0f);      // a zero the compiler cannot see
#pragma unroll
                        for (int q = 0; q < 2 * LPI; ++q) f[q] = __hfma2(fd[q], z, f[q]);
                    }
                    if (s < n_s) {
#pragma unroll
                        for (int q = 0; q < LPI / 2; ++q) {
                            uint4 v;
                            v.x = *reinterpret_cast<uint32_t*>(&f[4 * q + 0]); v.y = *reinterpret_cast<uint32_t*>(&f[4 * q + 1]);
                            v.z = *reinterpret_cast<uint32_t*>(&f[4 * q + 2]); v.w = *reinterpret_cast<uint32_t*>(&f[4 * q + 3]);
                            *reinterpret_cast<uint4*>(rowA32 + s * T2_TILE + (c * (LPI / 2) + q) * 128) = v;
                        }
                    }
                }
            }
        }
        fence_proxy_async();
        tc_fence_before();
        if (!__syncthreads_or(n_s > 0 ? 1 : 0)) {
            if (done) break;
            continue;
        }
        // ---- C. density layer 0: 32 -> 64, ReLU ----
        if (tid == 0) { tc_fence_after(); issue2(a32, T2_TILE, wd0, 32, 64); }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            uint32_t r[64];
            tmem_ld_32x32_x64(tmem_lane + s * 64, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint4 v;
                v.x = pack_relu_h2(r[8 * c + 0], r[8 * c + 1], true); v.y = pack_relu_h2(r[8 * c + 2], r[8 * c + 3], true);
                v.z = pack_relu_h2(r[8 * c + 4], r[8 * c + 5], true); v.w = pack_relu_h2(r[8 * c + 6], r[8 * c + 7], true);
                *reinterpret_cast<uint4*>(rowA64 + s * T2_TILE + c * 128) = v;
            }
        }
        fence_proxy_async(); tc_fence_before(); __syncthreads();
        // ---- D. density layer 1: 64 -> 16 (row 0 = raw density); rgb input = [16 density-out | 16 SH] ----
        if (tid == 0) { tc_fence_after(); issue2(a64, T2_TILE, wd1, 64, 16); }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        float sigma0 = 0.f, sigma1 = 0.f;
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            uint32_t r[16];
            tmem_ld_32x32_x16(tmem_lane + s * 64, r);
            tmem_ld_wait();
            const float sg = h2f_round(__uint_as_float(r[0]));
            if (s == 0) sigma0 = sg; else sigma1 = sg;
            uint4 v0, v1;
            v0.x = pack_relu_h2(r[0], r[1], false); v0.y = pack_relu_h2(r[2], r[3], false);
            v0.z = pack_relu_h2(r[4], r[5], false); v0.w = pack_relu_h2(r[6], r[7], false);
            v1.x = pack_relu_h2(r[8], r[9], false); v1.y = pack_relu_h2(r[10], r[11], false);
            v1.z = pack_relu_h2(r[12], r[13], false); v1.w = pack_relu_h2(r[14], r[15], false);
            unsigned char* row = rowA32 + s * T2_TILE;
            *reinterpret_cast<uint4*>(row + 0) = v0;
            *reinterpret_cast<uint4*>(row + 128) = v1;
            *reinterpret_cast<uint4*>(row + 256) = make_uint4(R.sh[0], R.sh[1], R.sh[2], R.sh[3]);
            *reinterpret_cast<uint4*>(row + 384) = make_uint4(R.sh[4], R.sh[5], R.sh[6], R.sh[7]);
        }
        fence_proxy_async(); tc_fence_before(); __syncthreads();
        // ---- E. rgb layer 0: 32 -> 64, ReLU ----
        if (tid == 0) { tc_fence_after(); issue2(a32, T2_TILE, wc0, 32, 64); }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            uint32_t r[64];
            tmem_ld_32x32_x64(tmem_lane + s * 64, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint4 v;
                v.x = pack_relu_h2(r[8 * c + 0], r[8 * c + 1], true); v.y = pack_relu_h2(r[8 * c + 2], r[8 * c + 3], true);
                v.z = pack_relu_h2(r[8 * c + 4], r[8 * c + 5], true); v.w = pack_relu_h2(r[8 * c + 6], r[8 * c + 7], true);
                *reinterpret_cast<uint4*>(rowA64 + s * T2_TILE + c * 128) = v;
            }
        }
        fence_proxy_async(); tc_fence_before(); __syncthreads();
        // ---- F. rgb layer 1: 64 -> 64, ReLU (in place: the MMA has retired before the rows are overwritten) ----
        if (tid == 0) { tc_fence_after(); issue2(a64, T2_TILE, wc1, 64, 64); }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            uint32_t r[64];
            tmem_ld_32x32_x64(tmem_lane + s * 64, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint4 v;
                v.x = pack_relu_h2(r[8 * c + 0], r[8 * c + 1], true); v.y = pack_relu_h2(r[8 * c + 2], r[8 * c + 3], true);
                v.z = pack_relu_h2(r[8 * c + 4], r[8 * c + 5], true); v.w = pack_relu_h2(r[8 * c + 6], r[8 * c + 7], true);
                *reinterpret_cast<uint4*>(rowA64 + s * T2_TILE + c * 128) = v;
            }
        }
        fence_proxy_async(); tc_fence_before(); __syncthreads();
        // ---- G. rgb output layer: 64 -> 16 (3 used) ----
        if (tid == 0) { tc_fence_after(); issue2(a64, T2_TILE, wc2, 64, 16); }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        float raw[2][3];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            uint32_t r[16];
            tmem_ld_32x32_x16(tmem_lane + s * 64, r);
            tmem_ld_wait();
            raw[s][0] = h2f_round(__uint_as_float(r[0])); raw[s][1] = h2f_round(__uint_as_float(r[1])); raw[s][2] = h2f_round(__uint_as_float(r[2]));
        }
        // ---- H. composite_kernel_nerf (testbed_nerf.cu:511-667), sample 0 then sample 1 ----
        if (alive) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (alive && s < n_s) {
                    ++my_samples;                  // composited samples only (a dropped sample 1 is not counted)
                    const float T = 1.f - R.ca;
                    const float alpha = 1.f - __expf(-__expf(s == 0 ? sigma0 : sigma1) * (s == 0 ? dtu0 : dtu1));
                    const float weight = alpha * T;
                    const float rr = logistic_d(raw[s][0]), gg = logistic_d(raw[s][1]), bb_ = logistic_d(raw[s][2]);
                    const float dep = s == 0 ? dep0 : dep1;
                    R.cr += rr * weight; R.cg += gg * weight; R.cb += bb_ * weight; R.cd += dep * weight; R.ca += weight;
                    if (R.ca > (1.0f - M.min_transmittance)) {
                        R.cr /= R.ca; R.cg /= R.ca; R.cb /= R.ca; R.cd /= R.ca; R.ca /= R.ca;
                        finish(R.cr, R.cg, R.cb, R.cd, R.ca, R.ei);
                        alive = false;
                    } else if (++R.n_steps >= MARCH_ITER - 1) {
                        finish(0.f, 0.f, 0.f, 0.f, 0.f, R.ei);       // never reaches the hit buffer in the reference
                        alive = false;
                    }
                }
            }
            if (alive && exits) {                   // ran out of occupied cells after the prepared samples
                finish(R.cr, R.cg, R.cb, R.cd, R.ca, R.ei);
                alive = false;
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
