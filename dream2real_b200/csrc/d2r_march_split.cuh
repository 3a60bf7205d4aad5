// Round-based split of the ray march (D2R_MARCH=split; kept as the A/B partner of k_march_ws): k_gather_round (walk + hash-grid gather, no barriers, no tensor-core state ->
// few registers, many warps per SM) and k_mlp_round (the five MLP layers on tcgen05 + compositing), alternating until no
// ray is left.  Included by d2r_march.cu after d2r_march_common.cuh.
//
// Why (round 1): a fused kernel whose threads do everything runs at 16 warps per SM -- 128 registers per thread, all TMEM columns, 208 KB of
// shared memory -- and every component's latency (walk, gather, five MMA round trips) is exposed in full: the kernel's time
// does not depend on the table size (2^14 vs 2^19 entries: same time per sample), doubling the gather adds 55 %, doubling
// the walk 17 % (profiles/).  The gather needs neither TMEM nor shared memory, so on its own it runs at 3x the occupancy.
// Price: 64 B of fp16 features per sample go through HBM once (written in the UMMA canonical operand layout, so the MLP
// kernel stages them with plain 16-byte copies), plus 24 B of per-ray accumulators per round.
//
// Per round r every live ray takes its next (up to) two samples -- the same walk, features and compositing arithmetic as
// k_march_ws, so the results are identical; a ray that finishes keeps its accumulators in res_rgbd / res_a (k_finish turns
// them into pixels), a ray that continues is appended to the next round's live list.  A live ray of round r has taken
// exactly 2 r samples, so the reference's step counter needs no storage.
#pragma once

namespace d2r {

struct SplitParams {
    int round;
    const uint32_t* cnt_in;     // live rays of this round (round 0: *n_entries)
    uint32_t* cnt_out;          // appended to by k_mlp_round
    const uint32_t* live_in;    // entry ids (round 0: identity)
    uint32_t* live_out;
    unsigned char* feat;        // [block][2][8 KB]: sample tile = 128 rows x 32 fp16 in UMMA canonical K-major layout
    float2* aux;                // [block][2][128]: (depth of the sample, unwarped dt)
    uint4* shb;                 // [block*128][2]: 16 fp16 SH coefficients of the ray direction, by position in this round's list
    const float4* acc4_in;      // [block*128]: accumulated (r, g, b, depth) of the ray at this position (rounds > 0) ...
    const float* acca_in;       //              ... and its accumulated alpha
    float4* acc4_out;           // the same for the next round's list, written where k_mlp_round appends the ray
    float* acca_out;
    uint8_t* nsb;               // [block*128]: samples prepared (0..2) | 16 if the ray leaves the occupied region after them
    float* t_cur;               // [entries]: ray parameter after the samples taken so far
    uint32_t cap;               // positions the per-round buffers hold; a launch with more hits than that skips the split rounds
};                              // on the device (k_march_ws then takes every ray from the start)

constexpr int SPLIT_TILE_BYTES = 128 * 32 * 2;

__global__ void __launch_bounds__(128, 7) k_gather_round(const __grid_constant__ MarchParams P, const __grid_constant__ SplitParams Q) {
    const ModelDev& M = P.M;
    const int tid = threadIdx.x;
    if (*P.n_entries > Q.cap) return;
    const uint32_t n_live = Q.round == 0 ? *P.n_entries : *Q.cnt_in;
    const StepC cone = make_stepc(M.cone);
    const uint32_t max_mip = (uint32_t)M.max_cascade;
    unsigned long long my_samples = 0;
    for (uint32_t blk = blockIdx.x; (size_t)blk * 128 < n_live; blk += gridDim.x) {
        const uint32_t i = blk * 128 + tid;
        const bool valid = i < n_live;
        if (!valid) continue;
        const uint32_t e = !valid ? 0u : (Q.round == 0 ? i : Q.live_in[i]);
        const RayEntry en = P.entries[e];
        const Mat3x4 C = P.cams[en.k];
        RayGeom g;
        ray_geom_only(C, __ldg(P.dirs + en.idx), g);                 // same arithmetic as pass 1 (setup_ray)
        g.t_exit = en.t_exit;
        float t = Q.round == 0 ? en.t : Q.t_cur[e];
        const float fwx = C.c[2][0], fwy = C.c[2][1], fwz = C.c[2][2];
        {   // SH coefficients of the ray direction, by position like everything else k_mlp_round reads: its inputs are contiguous
            float sh[16];
            const float wx = (g.dx + 1.0f) * 0.5f, wy = (g.dy + 1.0f) * 0.5f, wz = (g.dz + 1.0f) * 0.5f;
            sh_enc4(wx * 2.f - 1.f, wy * 2.f - 1.f, wz * 2.f - 1.f, sh);
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                __half2 h = __floats2half2_rn(sh[2 * j], sh[2 * j + 1]);
                pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            Q.shb[(size_t)i * 2] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            Q.shb[(size_t)i * 2 + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        // the walk of k_march_ws (phase B), verbatim
        int n_s = 0;
        bool exits = false;
        float wp0x = 0.f, wp0y = 0.f, wp0z = 0.f, wp1x = 0.f, wp1y = 0.f, wp1z = 0.f;
        float dep0 = 0.f, dep1 = 0.f, dtu0 = 0.f, dtu1 = 0.f;
        while (valid) {
            const float px = g.ox + t * g.dx, py = g.oy + t * g.dy, pz = g.oz + t * g.dz;
            if (t >= MAX_DEPTH() || t > g.t_exit || !raabb_contains(M, px, py, pz)) { exits = true; break; }
            uint32_t mip = min(max(mip_from_pos(px, py, pz), 0u), max_mip);
            if (density_grid_occupied_at(px, py, pz, M.bitfield_lin, mip)) {
                const float dt = calc_dt(t, cone);
                const float wx = (px - M.aabb_min[0]) / M.aabb_diag[0];
                const float wy = (py - M.aabb_min[1]) / M.aabb_diag[1];
                const float wz = (pz - M.aabb_min[2]) / M.aabb_diag[2];
                const float ux = M.aabb_min[0] + wx * M.aabb_diag[0];
                const float uy = M.aabb_min[1] + wy * M.aabb_diag[1];
                const float uz = M.aabb_min[2] + wz * M.aabb_diag[2];
                float dep = 0.f;
                dep += fwx * (ux - g.ox); dep += fwy * (uy - g.oy); dep += fwz * (uz - g.oz);
                dep *= M.depth_scale;
                const float dtu = unwarp_dt(warp_dt(dt));
                if (n_s == 0) { wp0x = wx; wp0y = wy; wp0z = wz; dep0 = dep; dtu0 = dtu; }
                else { wp1x = wx; wp1y = wy; wp1z = wz; dep1 = dep; dtu1 = dtu; }
                t += dt;
                if (++n_s == 2) break;
                continue;
            }
            while (mip < max_mip && !density_grid_occupied_at(px, py, pz, M.bitfield_lin, mip + 1)) ++mip;
            t = advance_to_next_voxel(t, cone, px, py, pz, g.dx, g.dy, g.dz, g.ix, g.iy, g.iz, mip);
        }
        if (valid) {
            Q.t_cur[e] = t;
            Q.nsb[i] = (uint8_t)(n_s | (exits ? 16 : 0));
            Q.aux[((size_t)blk * 2 + 0) * 128 + tid] = make_float2(dep0, dtu0);
            Q.aux[((size_t)blk * 2 + 1) * 128 + tid] = make_float2(dep1, dtu1);
        }
        my_samples += (unsigned)n_s;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (s < n_s) {
                const float sx = s ? wp1x : wp0x, sy = s ? wp1y : wp0y, sz = s ? wp1z : wp0z;
                unsigned char* row = Q.feat + ((size_t)blk * 2 + s) * SPLIT_TILE_BYTES + umma_chunk_off(tid, 0, 32);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    __half2 f[4];
                    encode_levels<2>(M, 2 * c, sx, sy, sz, f);
                    uint4 v;
                    v.x = *reinterpret_cast<uint32_t*>(&f[0]); v.y = *reinterpret_cast<uint32_t*>(&f[1]);
                    v.z = *reinterpret_cast<uint32_t*>(&f[2]); v.w = *reinterpret_cast<uint32_t*>(&f[3]);
                    *reinterpret_cast<uint4*>(row + c * 128) = v;
                }
            }
            if (s >= n_s) {      // no such sample: a zero row, so that k_mlp_round can fetch rows before it knows n_s
                unsigned char* row = Q.feat + ((size_t)blk * 2 + s) * SPLIT_TILE_BYTES + umma_chunk_off(tid, 0, 32);
#pragma unroll
                for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(row + c * 128) = make_uint4(0, 0, 0, 0);
            }
        }
    }
    (void)my_samples;
}

// shared memory plan of k_mlp_round (bytes): weights, 2 tiles whose K=32 and K=64 operands alias, and the feature staging
// buffer the bulk copies land in (68 KB -> 3 CTAs per SM)
constexpr int T2_A64 = W_BYTES;
constexpr int T2_A32 = T2_A64;
constexpr int T2_TILE = 16384;
constexpr int T2_FEAT = T2_A64 + 2 * T2_TILE;          // [2][8 KB]: a block's two sample tiles exactly as k_gather_round wrote them
constexpr int T2_MISC = T2_FEAT + 2 * SPLIT_TILE_BYTES;
constexpr int T2_TOTAL = T2_MISC + 256;

// A block's two feature tiles (16 KB, contiguous, already in UMMA operand layout) arrive with ONE cp.async.bulk straight into the
// buffer the first layer's MMAs read: no thread touches them.  The copy for the next block is issued as soon as those MMAs
// have retired, so it flies during the other four layers; a block's live-list append is finished during the next block's
// first MMA wait (the atomic's round trip is off the critical path).
__global__ void __launch_bounds__(TC_THREADS, 3) k_mlp_round(const __grid_constant__ MarchParams P, const __grid_constant__ SplitParams Q) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (*P.n_entries > Q.cap) return;
    const ModelDev& M = P.M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + T2_MISC);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T2_MISC + 16);      // (mbarriers at +0, +8, +24)

    for (int i = tid; i < W_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(M.w_umma)[i];
    uint64_t* bar_feat = mbar + 3;
    if (tid == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); mbar_init(bar_feat, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<128>(tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t a64 = smem_u32(smem + T2_A64);
    unsigned char* rowA32 = smem + T2_A32 + umma_chunk_off(tid, 0, 32);
    unsigned char* rowA64 = smem + T2_A64 + umma_chunk_off(tid, 0, 64);
    uint32_t phase0 = 0, phase1 = 0, phase_feat = 0;
    const uint32_t n_live = Q.round == 0 ? *P.n_entries : *Q.cnt_in;
    const uint32_t f_lo = (smem_u32(smem + T2_FEAT) >> 4) + (8u << 16);
    auto fetch_features = [&](uint32_t b) {      // tid 0: the two sample tiles of block b -> staging buffer
        mbar_arrive_expect_tx(bar_feat, 2 * SPLIT_TILE_BYTES);
        bulk_g2s(smem + T2_FEAT, Q.feat + (size_t)b * 2 * SPLIT_TILE_BYTES, 2 * SPLIT_TILE_BYTES, bar_feat);
    };
    if (tid == 0 && (size_t)blockIdx.x * 128 < n_live) fetch_features(blockIdx.x);
    unsigned long long my_samples = 0, my_rays = 0;
    // descriptor low words ((address >> 4) | LBO 128 B): the issuing thread's path is a handful of 32-bit adds per MMA
    const uint32_t a_lo = (a64 >> 4) + (8u << 16), w_lo = (smem_u32(smem) >> 4) + (8u << 16);
    // the previous block's live-list append (its atomic is already in flight): entry id and accumulators of the rays that go on
    uint32_t pend_go = 0, pend_base = 0, pend_e = 0;
    bool pend_alive = false;
    float4 pend_a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float pend_aa = 0.f;
    auto flush_append = [&]() {
        if (pend_go) {
            const uint32_t base = __shfl_sync(0xffffffffu, pend_base, __ffs(pend_go) - 1);
            if (pend_alive) {
                const uint32_t pos = base + __popc(pend_go & ((1u << lane) - 1));
                Q.live_out[pos] = pend_e;
                Q.acc4_out[pos] = pend_a4;
                Q.acca_out[pos] = pend_aa;
            }
            pend_go = 0;
        }
    };
    // Everything a block reads per ray is indexed by the ray's POSITION in this round's list (coalesced), and is fetched one
    // block ahead: the loads are issued after the last proxy fence of the previous block, so no fence ever waits for them.
    struct In { uint32_t ns_raw, e; uint4 sh0, sh1; float2 ax0, ax1; float4 a4; float aa; };
    auto load_inputs = [&](uint32_t b) {
        In in;
        const uint32_t i = b * 128 + tid;
        in.ns_raw = 0u; in.e = 0u; in.sh0 = in.sh1 = make_uint4(0, 0, 0, 0); in.ax0 = in.ax1 = make_float2(0.f, 0.f);
        in.a4 = make_float4(0.f, 0.f, 0.f, 0.f); in.aa = 0.f;
        if (i < n_live) {
            in.ns_raw = Q.nsb[i] | 0x100u;                                   // bit 8: a ray sits at this position
            in.e = Q.round == 0 ? i : Q.live_in[i];
            in.sh0 = Q.shb[(size_t)i * 2]; in.sh1 = Q.shb[(size_t)i * 2 + 1];
            in.ax0 = Q.aux[((size_t)b * 2 + 0) * 128 + tid];
            in.ax1 = Q.aux[((size_t)b * 2 + 1) * 128 + tid];
            if (Q.round > 0) { in.a4 = Q.acc4_in[i]; in.aa = Q.acca_in[i]; }
        }
        return in;
    };
    In nxt = load_inputs(blockIdx.x);
    for (uint32_t blk = blockIdx.x; (size_t)blk * 128 < n_live; blk += gridDim.x) {
        const In cur = nxt;
        const bool valid = (cur.ns_raw & 0x100u) != 0;
        const int n_s = (int)(cur.ns_raw & 15u);
        const bool exits = (cur.ns_raw & 16u) != 0;
        const uint32_t e = cur.e;
        float cr = cur.a4.x, cg = cur.a4.y, cb = cur.a4.z, cd = cur.a4.w, ca = cur.aa;
        const uint4 sh0 = cur.sh0, sh1 = cur.sh1;
        const float2 ax0 = cur.ax0, ax1 = cur.ax1;
        if (valid && Q.round == 0) ++my_rays;
        // The five layers, the two sample tiles skewed against each other: each tile has its own mbarrier and its own TMEM
        // columns, and the MMAs of one tile's next layer are issued the moment its rows are written -- they run while every
        // thread is busy with the other tile's tcgen05.ld / ReLU / fp16 pack / store.  (Issuing is a handful of 32-bit adds
        // per MMA: issue_tile_ws.)
        auto wait_tile = [&](int s) {
            if (s == 0) { mbar_wait(mbar, phase0); phase0 ^= 1; } else { mbar_wait(mbar + 1, phase1); phase1 ^= 1; }
            tc_fence_after();
        };
        auto hidden_rows = [&](int s) {      // 64 outputs, ReLU, fp16 -> the tile's K=64 operand rows
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
        };
        auto rows_done = [&]() { fence_proxy_async(); tc_fence_before(); __syncthreads(); };
        // ---- density layer 0: 32 -> 64, ReLU; A = the staged feature tiles ----
        if (tid == 0) {
            mbar_wait(bar_feat, phase_feat);
            tc_fence_after();
            issue_tile_ws<32, 64>(f_lo, w_lo + (W_D0 >> 4), tmem_base); tc_commit(mbar);
            issue_tile_ws<32, 64>(f_lo + (SPLIT_TILE_BYTES >> 4), w_lo + (W_D0 >> 4), tmem_base + 64); tc_commit(mbar + 1);
        }
        phase_feat ^= 1;
        flush_append();          // the previous block's append: its atomic returned long ago
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            wait_tile(s);
            if (s == 1 && tid == 0) {      // both tiles' first-layer MMAs have retired: the staging buffer is free for the next block
                const uint32_t nb = blk + gridDim.x;
                if ((size_t)nb * 128 < n_live) fetch_features(nb);
            }
            hidden_rows(s);
            rows_done();
            // ---- density layer 1: 64 -> 16 (row 0 = raw density) ----
            if (tid == 0) { tc_fence_after(); issue_tile_ws<64, 16>(a_lo + s * 1024, w_lo + (W_D1 >> 4), tmem_base + s * 64); tc_commit(mbar + s); }
        }
        float sigma0 = 0.f, sigma1 = 0.f;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            wait_tile(s);
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
            unsigned char* row = rowA32 + s * T2_TILE;      // rgb input = [16 density-out | 16 SH]
            *reinterpret_cast<uint4*>(row + 0) = v0;
            *reinterpret_cast<uint4*>(row + 128) = v1;
            *reinterpret_cast<uint4*>(row + 256) = sh0;
            *reinterpret_cast<uint4*>(row + 384) = sh1;
            rows_done();
            // ---- rgb layer 0: 32 -> 64, ReLU ----
            if (tid == 0) { tc_fence_after(); issue_tile_ws<32, 64>(a_lo + s * 1024, w_lo + (W_C0 >> 4), tmem_base + s * 64); tc_commit(mbar + s); }
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            wait_tile(s);
            hidden_rows(s);
            rows_done();
            // ---- rgb layer 1: 64 -> 64, ReLU ----
            if (tid == 0) { tc_fence_after(); issue_tile_ws<64, 64>(a_lo + s * 1024, w_lo + (W_C1 >> 4), tmem_base + s * 64); tc_commit(mbar + s); }
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            wait_tile(s);
            hidden_rows(s);      // in place: the MMA that read these rows has retired
            rows_done();
            // ---- rgb output layer: 64 -> 16 (3 used) ----
            if (tid == 0) { tc_fence_after(); issue_tile_ws<64, 16>(a_lo + s * 1024, w_lo + (W_C2 >> 4), tmem_base + s * 64); tc_commit(mbar + s); }
        }
        // that was this block's last proxy fence: the next block's per-ray inputs start their way now and land during the
        // output layer, the compositing and the next block's first layer
        nxt = load_inputs(blk + gridDim.x);
        float raw[2][3];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            wait_tile(s);
            uint32_t r[16];
            tmem_ld_32x32_x16(tmem_lane + s * 64, r);
            tmem_ld_wait();
            raw[s][0] = h2f_round(__uint_as_float(r[0])); raw[s][1] = h2f_round(__uint_as_float(r[1])); raw[s][2] = h2f_round(__uint_as_float(r[2]));
        }
        // ---- composite_kernel_nerf (testbed_nerf.cu:511-667), sample 0 then sample 1; same arithmetic as k_march_ws ----
        bool alive = valid;
        if (alive) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (alive && s < n_s) {
                    ++my_samples;
                    const float T = 1.f - ca;
                    const float alpha = 1.f - __expf(-__expf(s == 0 ? sigma0 : sigma1) * (s == 0 ? ax0.y : ax1.y));
                    const float weight = alpha * T;
                    const float rr = logistic_d(raw[s][0]), gg = logistic_d(raw[s][1]), bb_ = logistic_d(raw[s][2]);
                    const float dep = s == 0 ? ax0.x : ax1.x;
                    cr += rr * weight; cg += gg * weight; cb += bb_ * weight; cd += dep * weight; ca += weight;
                    if (ca > (1.0f - M.min_transmittance)) {
                        cr /= ca; cg /= ca; cb /= ca; cd /= ca; ca /= ca;
                        alive = false;
                    } else if (2 * Q.round + s + 1 >= MARCH_ITER - 1) {      // a live ray of round r has taken 2 r samples
                        cr = cg = cb = cd = ca = 0.f;                          // never reaches the hit buffer in the reference
                        alive = false;
                    }
                }
            }
            if (alive && (exits || n_s < 2)) { alive = false; }   // ran out of occupied cells
        }
        if (valid && !alive) {      // finished: what k_finish turns into a pixel
            P.res_rgbd[e] = make_float4(cr, cg, cb, cd);
            P.res_a[e] = ca;
        }
        // rays that go on: append to the next round's live list (one atomic per warp)
        const uint32_t go = __ballot_sync(0xffffffffu, alive);
        if (go) {
            const int leader = __ffs(go) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(Q.cnt_out, (uint32_t)__popc(go));
            pend_go = go; pend_base = base; pend_alive = alive; pend_e = e;      // finished by flush_append()
            pend_a4 = make_float4(cr, cg, cb, cd); pend_aa = ca;
        }
        tc_fence_before();
        __syncthreads();      // every thread has read its output rows: the next block's first MMAs may overwrite the TMEM columns
    }
    flush_append();
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
