// Round-based split of the ray march: k_gather_round (walk + hash-grid gather, no barriers, no tensor-core state -> few
// registers, many warps per SM) and k_mlp_round (the five MLP layers on tcgen05 + compositing), alternating for a fixed number
// of rounds; whatever is still alive then resumes in k_march_ws.  Included by d2r_march.cu after d2r_march_common.cuh.
//
// Why: a fused kernel whose threads do everything runs at 16 gather warps per SM next to its tensor-core state, and every
// component's latency (walk, gather, five MMA round trips) is exposed; the gather needs neither TMEM nor shared memory, so on
// its own it runs at 28 warps per SM.  Price: 64 B of fp16 features per sample go through HBM once (written in the UMMA
// canonical operand layout, so k_mlp_round stages a block's two tiles with one bulk copy), plus the per-ray accumulators.
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

// ---------------------------------------------------------------------------------------------------------------------------
// k_mlp_round: the five MLP layers of a round's samples on tcgen05, the activations kept in TENSOR MEMORY between layers
// ("TS" operand mode), then the compositing and the live-list append.
//
// tools/tmem_port_bench.cu (profiles/r2_tmem_port_bench.json): one layer of one 128-sample tile is a ~1000-cycle dependent
// chain (MMA round trip ~780 cycles), SS-mode MMAs spend most of their time fetching the 128-row A operand from shared memory,
// and throughput = tiles in flight / chain length.  Here a thread's ReLU'd fp16 row goes back with tcgen05.st into 32 columns
// next to the accumulators and the next layer's MMAs read A from there: no shared-memory operand rows, no proxy fences, MMAs of
// 32 / 8 cycles instead of 48 / 39.  A tile needs 96 columns, so an SM holds five: five independent 128-thread groups per CTA,
// each taking its block's two sample tiles through the layers one after the other (the other four groups fill its round
// trips); the MMAs of a layer leave back to back from an elected lane of the group's converged first warp with warp-uniform
// descriptors.  Only the first layer is SS: its A operand is the block's feature tiles (16 KB, contiguous, already in UMMA
// operand layout), which ONE cp.async.bulk lands in the group's staging buffer -- no thread touches them -- issued for the
// next block as soon as this block's first layers have retired.  A block's live-list append is finished during the next
// block's first MMA wait (the atomic's round trip is off the critical path).
// Same arithmetic per sample as k_march_ws (same MMA shapes and K order, same rounding points): identical frames.
// The shared-memory-operand form this replaces (3 CTAs x 2 skewed tiles per SM) took 43 ms of a 191 ms step, this one 30.
constexpr int TS_GROUPS = 5;
constexpr int TS_COLS = 96;                                   // per group: accumulators [0, 64) + fp16 operand rows [64, 96)
constexpr int TS_FEAT_BYTES = 2 * SPLIT_TILE_BYTES;           // per group: a block's two staged sample tiles
constexpr int TS_MISC = W_BYTES + TS_GROUPS * TS_FEAT_BYTES;
constexpr int TS_TOTAL = TS_MISC + 32 * TS_GROUPS + 64;

template <bool ACC>
__device__ __forceinline__ void umma_f16_ts_lohi(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 db;\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(hi), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
// D[128 x N] = A[128 x K] (TMEM: row = lane, two fp16 per 32-bit column) * W[N x K]^T (shared memory, K-major, no swizzle)
template <int K, int N>
__device__ __forceinline__ void issue_tile_ts(uint32_t tmem_a, uint32_t b_lo, uint32_t tmem_d) {
    constexpr uint32_t hi = (uint32_t)((K / 8) * 128 >> 4) | (1u << 14);
    constexpr uint32_t idesc = umma_idesc_f16(128, N, 0);
    umma_f16_ts_lohi<false>(tmem_d, tmem_a, b_lo, hi, idesc);
#pragma unroll
    for (int kk = 1; kk < K / 16; ++kk) umma_f16_ts_lohi<true>(tmem_d, tmem_a + kk * 8, b_lo + kk * 16, hi, idesc);
}
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(128 * TS_GROUPS, 1) k_mlp_round(const __grid_constant__ MarchParams P, const __grid_constant__ SplitParams Q) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (*P.n_entries > Q.cap) return;
    const ModelDev& M = P.M;
    // the warp index as a value the compiler knows to be warp-uniform: descriptors, TMEM and mbarrier addresses of the issue
    // path then live in uniform registers (no per-lane R2UR waterfall in front of every tcgen05.mma)
    const int warp_cta = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int grp = warp_cta >> 2, warp = warp_cta & 3, tid = threadIdx.x & 127, lane = tid & 31;      // warp, tid: within the group
    unsigned char* feat = smem + W_BYTES + grp * TS_FEAT_BYTES;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + TS_MISC + 32 * grp);      // [0] the group's MMAs, [1] its staged features
    uint64_t* bar_feat = mbar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TS_MISC + 32 * TS_GROUPS);
    const uint32_t vblk0 = blockIdx.x * TS_GROUPS + grp, vstride = gridDim.x * TS_GROUPS;
    auto group_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); };

    for (int i = threadIdx.x; i < W_BYTES / 16; i += 128 * TS_GROUPS) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(M.w_umma)[i];
    if (tid == 0) { mbar_init(mbar, 1); mbar_init(bar_feat, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_all = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t tmem_d = tmem_all + TS_COLS * grp, tmem_a = tmem_d + 64;
    const uint32_t lane_d = tmem_d + ((uint32_t)(warp * 32) << 16), lane_a = lane_d + 64;
    const uint32_t n_live = Q.round == 0 ? *P.n_entries : *Q.cnt_in;
    const uint32_t f_lo = (smem_u32(feat) >> 4) + (8u << 16), w_lo = (smem_u32(smem) >> 4) + (8u << 16);
    uint32_t phase = 0, phase_feat = 0;
    auto fetch_features = [&](uint32_t b) {      // tid 0: the two sample tiles of block b -> the group's staging buffer
        mbar_arrive_expect_tx(bar_feat, TS_FEAT_BYTES);
        bulk_g2s(feat, Q.feat + (size_t)b * TS_FEAT_BYTES, TS_FEAT_BYTES, bar_feat);
    };
    if (tid == 0 && (size_t)vblk0 * 128 < n_live) fetch_features(vblk0);
    unsigned long long my_samples = 0, my_rays = 0;
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
    auto wait_mma = [&]() { mbar_wait(mbar, phase); phase ^= 1; tc_fence_after(); };
    auto rows_done = [&]() { tmem_st_wait(); tc_fence_before(); group_barrier(); };
    auto hidden_rows = [&]() {      // 64 outputs, ReLU, fp16 -> this thread's row of the next layer's A operand (32 columns)
#pragma unroll
        for (int h = 0; h < 2; ++h) {      // in halves, 32 accumulators -> 16 packed columns: five groups share the register file
            uint32_t r[32];                //   (both loads in flight at once spills and measured 4 % slower per step)
            tmem_ld_32x32(lane_d + 32 * h, r);
            tmem_ld_wait();
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = pack_relu_h2(r[2 * j], r[2 * j + 1], true);
            tmem_st_32x32_x16(lane_a + 16 * h, v);
        }
    };
    for (uint32_t blk = vblk0; (size_t)blk * 128 < n_live; blk += vstride) {
        // Everything a block reads per ray is indexed by the ray's POSITION in this round's list (coalesced); the loads leave
        // now and are first needed a layer (SH), or all ten layers (the rest), later.
        const uint32_t i = blk * 128 + tid;
        const bool valid = i < n_live;
        uint32_t ns_raw = 0, e = 0;
        uint4 sh0 = make_uint4(0, 0, 0, 0), sh1 = sh0;
        float2 ax0 = make_float2(0.f, 0.f), ax1 = ax0;
        float cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f, ca = 0.f;
        if (valid) {
            ns_raw = Q.nsb[i];
            e = Q.round == 0 ? i : Q.live_in[i];
            sh0 = Q.shb[(size_t)i * 2]; sh1 = Q.shb[(size_t)i * 2 + 1];
            ax0 = Q.aux[((size_t)blk * 2 + 0) * 128 + tid];
            ax1 = Q.aux[((size_t)blk * 2 + 1) * 128 + tid];
            if (Q.round > 0) { const float4 a4 = Q.acc4_in[i]; cr = a4.x; cg = a4.y; cb = a4.z; cd = a4.w; ca = Q.acca_in[i]; }
            if (Q.round == 0) ++my_rays;
        }
        float sigma[2], raw[2][3];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            // ---- density layer 0: 32 -> 64, ReLU; A = the staged feature tile (shared memory) ----
            if (warp == 0) {
                if (s == 0) mbar_wait(bar_feat, phase_feat);
                tc_fence_after();
                if (elect_one()) { issue_tile_ws<32, 64>(f_lo + s * (SPLIT_TILE_BYTES >> 4), w_lo + (W_D0 >> 4), tmem_d); tc_commit(mbar); }
                __syncwarp();
            }
            if (s == 0) flush_append();          // the previous block's append: its atomic returned long ago
            wait_mma();
            if (s == 1 && tid == 0) {            // both tiles' first layers have retired: the staging buffer is free for the next block
                const uint32_t nb = blk + vstride;
                if ((size_t)nb * 128 < n_live) fetch_features(nb);
            }
            hidden_rows();
            rows_done();
            // ---- density layer 1: 64 -> 16 (row 0 = raw density) ----
            if (warp == 0) { tc_fence_after(); if (elect_one()) { issue_tile_ts<64, 16>(tmem_a, w_lo + (W_D1 >> 4), tmem_d); tc_commit(mbar); } __syncwarp(); }
            wait_mma();
            {
                uint32_t r[16];
                tmem_ld_32x32_x16(lane_d, r);
                tmem_ld_wait();
                sigma[s] = h2f_round(__uint_as_float(r[0]));
                uint32_t v[16];      // rgb input = [16 density-out | 16 SH]
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = pack_relu_h2(r[2 * j], r[2 * j + 1], false);
                v[8] = sh0.x; v[9] = sh0.y; v[10] = sh0.z; v[11] = sh0.w; v[12] = sh1.x; v[13] = sh1.y; v[14] = sh1.z; v[15] = sh1.w;
                tmem_st_32x32_x16(lane_a, v);
            }
            rows_done();
            // ---- rgb layer 0: 32 -> 64, ReLU ----
            if (warp == 0) { tc_fence_after(); if (elect_one()) { issue_tile_ts<32, 64>(tmem_a, w_lo + (W_C0 >> 4), tmem_d); tc_commit(mbar); } __syncwarp(); }
            wait_mma();
            hidden_rows();
            rows_done();
            // ---- rgb layer 1: 64 -> 64, ReLU ----
            if (warp == 0) { tc_fence_after(); if (elect_one()) { issue_tile_ts<64, 64>(tmem_a, w_lo + (W_C1 >> 4), tmem_d); tc_commit(mbar); } __syncwarp(); }
            wait_mma();
            hidden_rows();
            rows_done();
            // ---- rgb output layer: 64 -> 16 (3 used) ----
            if (warp == 0) { tc_fence_after(); if (elect_one()) { issue_tile_ts<64, 16>(tmem_a, w_lo + (W_C2 >> 4), tmem_d); tc_commit(mbar); } __syncwarp(); }
            wait_mma();
            {
                uint32_t r[16];
                tmem_ld_32x32_x16(lane_d, r);
                tmem_ld_wait();
                raw[s][0] = h2f_round(__uint_as_float(r[0])); raw[s][1] = h2f_round(__uint_as_float(r[1])); raw[s][2] = h2f_round(__uint_as_float(r[2]));
            }
            tc_fence_before();
            group_barrier();      // every thread has read its output row: the next tile's first MMAs may overwrite the accumulators
        }
        phase_feat ^= 1;
        const int n_s = (int)(ns_raw & 15u);
        const bool exits = (ns_raw & 16u) != 0;
        // ---- composite_kernel_nerf (testbed_nerf.cu:511-667), sample 0 then sample 1; same arithmetic as k_march_ws ----
        bool alive = valid;
        if (alive) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (alive && s < n_s) {
                    ++my_samples;
                    const float T = 1.f - ca;
                    const float alpha = 1.f - __expf(-__expf(sigma[s]) * (s == 0 ? ax0.y : ax1.y));
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
        // rays that go on: append to the next round's live list (one atomic per warp), finished by flush_append()
        const uint32_t go = __ballot_sync(0xffffffffu, alive);
        if (go) {
            const int leader = __ffs(go) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(Q.cnt_out, (uint32_t)__popc(go));
            pend_go = go; pend_base = base; pend_alive = alive; pend_e = e;
            pend_a4 = make_float4(cr, cg, cb, cd); pend_aa = ca;
        }
    }
    flush_append();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem_all); }
    if (P.n_samples || P.prof) {
        for (int o = 16; o > 0; o >>= 1) my_samples += __shfl_xor_sync(0xffffffffu, my_samples, o);
        for (int o = 16; o > 0; o >>= 1) my_rays += __shfl_xor_sync(0xffffffffu, my_rays, o);
        if (lane == 0) {
            if (P.n_samples && my_samples) atomicAdd(P.n_samples, my_samples);
            if (P.prof && my_samples) atomicAdd(P.prof, my_samples);
            if (P.prof && my_rays) atomicAdd(P.prof + 1, my_rays);
        }
    }
}

}  // namespace d2r
