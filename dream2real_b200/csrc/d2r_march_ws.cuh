// k_march_ws: the ray march as ONE persistent, warp-specialised kernel (one CTA per SM).
// Included by d2r_march.cu after d2r_march_common.cuh.
//
// Replaces the reference's per-round kernel chain with a host synchronisation after every round
// (NGP src/testbed_nerf.cu:1632-1748: generate_next_nerf_network_inputs, kernel_grid, kernel_mlp_fused x2, kernel_sh,
// composite_kernel_nerf, compact_kernel_nerf): rays never leave the SM between their first and their last sample.
//
// Roles inside a CTA (NG = 4 gather groups):
//   * warps [0, 4 NG)      GATHER: thread = ray slot.  Refills its slot from the global hit list (warp-aggregated claims, so
//                          lanes that die together get neighbouring pixels), walks the occupancy grid to the ray's next two
//                          samples, gathers the 8-level hash-grid features (fp16 fma chain like the reference) and writes
//                          them as its row of the group's two A tiles (UMMA canonical K-major layout) in shared memory.
//   * warps [4 NG, 4 NG+4) EPILOGUE: thread = TMEM lane = tile row.  After every layer: tcgen05.ld of its row, ReLU + fp16
//                          rounding (the reference keeps fp16 activations, TCNN fully_fused_mlp.cu:47-129), next layer's A
//                          row back to shared memory; after the last layer the transmittance compositing of colour AND depth
//                          (composite_kernel_nerf, testbed_nerf.cu:511-667) on accumulators that live in shared memory.
//   * warp 4 NG + 4        MMA: one thread polls the groups' "A operand ready" mbarriers and issues the five layers
//                          32->64->16 | [16 | 16 SH]->64->64->16(3) as tcgen05.mma.cta_group::1.kind::f16 (M=128, N=64|16, K=16)
//                          against the weights (one cp.async.bulk of the model's pre-arranged operand blob), accumulators in TMEM
//                          (NG x 2 tiles x 64 columns = all 512).  Every issue is followed by a tcgen05.commit onto the next
//                          slot of a small ring of mbarriers; tcgen05 operations retire in issue order, so the epilogue
//                          warps consume the ring in order and never poll.
// While a group waits for its MLP round, the other groups' gather warps keep the load/store units busy and vice versa:
// the gather (L1-wavefront / issue bound) and the MLP round trips (latency bound) overlap inside one SM, and the 64 B of
// features per sample never touch HBM.  Same arithmetic per sample as k_gather_round / k_mlp_round: identical frames.
#pragma once

namespace d2r {

constexpr int WS_NG = 4;                                   // gather groups per CTA
constexpr int WS_THREADS = WS_NG * 128 + 128 + 32;          // 672
constexpr int WS_RING = 8;                                  // commit ring slots (>= WS_NG, see the flow-control note below)
constexpr uint32_t WS_EXIT = 0xffu;

// shared memory plan (bytes)
constexpr int WS_W = 0;                                     // 20480: MLP weights as UMMA B operands (W_* sub-offsets)
constexpr int WS_A = 20480;                                 // per group 2 tiles x 16 KB (K=32 and K=64 operands alias)
constexpr int WS_SH = WS_A + WS_NG * 32768;                 // per group: sh_lo[128], sh_hi[128] uint4 (16 fp16 SH coefficients)
constexpr int WS_AUX = WS_SH + WS_NG * 4096;                // per group: float4 (depth0, dt0, depth1, dt1) per slot
constexpr int WS_META = WS_AUX + WS_NG * 2048;              // per group: entry id[128], flags[128]
constexpr int WS_ACC = WS_META + WS_NG * 1024;              // per group: cr, cg, cb, cd, ca, n_steps [6][128]
constexpr int WS_SIG = WS_ACC + WS_NG * 3072;               // per group: raw density of the two samples [2][128]
constexpr int WS_STAT = WS_SIG + WS_NG * 1024;              // per group: alive after this round [128]
constexpr int WS_MISC = WS_STAT + WS_NG * 512;              // barriers, ring, cursors
constexpr int WS_TOTAL = WS_MISC + 512;

struct WsMisc {
    uint64_t bar_w;                 // weights landed
    uint64_t bar_a[WS_NG];          // group's next A operand is complete (128 arrivals: gather threads or epilogue threads)
    uint64_t bar_done[WS_NG];       // group's round is composited (128 epilogue arrivals)
    uint64_t bar_seq[WS_RING];      // commit ring
    uint32_t ring[WS_RING];         // (sequence number << 16) | (layer << 8) | group
    uint32_t cur[WS_NG], end[WS_NG], done[WS_NG], exit_[WS_NG];
    uint32_t tmem_slot;
};
static_assert(sizeof(WsMisc) <= 512, "WsMisc must fit its slot");

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// wait with a watchdog: a protocol bug must end in a launch failure, never in a hung GPU
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)      // suspend-time hint (ns): sleep in hardware, wake on completion
            : "memory");
        if (ok) return;
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ bool group_or(int id, bool v) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 q, %1, 0;\n"
        "bar.red.or.pred p, %2, 128, q;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(r)
        : "r"((uint32_t)v), "r"(id)
        : "memory");
    return r != 0;
}
__device__ __forceinline__ void tmem_ld_32x32_x32(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }

// STATS: accumulate the per-role cycle statistics (tools/ws_stats.py); the production instantiation carries none of it
template <bool STATS>
__global__ void __launch_bounds__(WS_THREADS, 1) k_march_ws(const __grid_constant__ MarchParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const ModelDev& M = P.M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    WsMisc* mi = reinterpret_cast<WsMisc*>(smem + WS_MISC);

    if (tid == 0) {
        mbar_init(&mi->bar_w, 1);
        for (int g = 0; g < WS_NG; ++g) {
            mbar_init(&mi->bar_a[g], 128);
            mbar_init(&mi->bar_done[g], 128);
            mi->cur[g] = 0; mi->end[g] = 0; mi->done[g] = 0; mi->exit_[g] = 0;
        }
        for (int s = 0; s < WS_RING; ++s) { mbar_init(&mi->bar_seq[s], 1); mi->ring[s] = 0xffffffffu; }
        fence_barrier_init();
    }
    if (warp == WS_NG * 4 + 4) tmem_alloc<512>(&mi->tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = mi->tmem_slot;
    // rays resumed from the split rounds -- unless this launch had more hits than their buffers hold, in which case those
    // kernels returned at once and every ray starts here
    const bool resumed = P.resume_steps > 0 && *P.n_entries <= P.split_cap;
    const uint32_t total_entries = resumed ? *P.work_count : *P.n_entries;
    unsigned long long my_samples = 0, my_rays = 0;

    if (warp < WS_NG * 4) {
        // ================================================= GATHER =================================================
        const int g = warp >> 2, r = tid & 127, bar_id = 1 + g;
        const StepC cone = make_stepc(M.cone);
        const uint32_t max_mip = (uint32_t)M.max_cascade;
        unsigned char* tileA = smem + WS_A + g * 32768;
        unsigned char* rowA32 = tileA + umma_chunk_off(r, 0, 32);
        uint4* sh_lo = reinterpret_cast<uint4*>(smem + WS_SH + g * 4096);
        uint4* sh_hi = sh_lo + 128;
        float4* aux = reinterpret_cast<float4*>(smem + WS_AUX + g * 2048);
        uint32_t* eis = reinterpret_cast<uint32_t*>(smem + WS_META + g * 1024);
        uint32_t* flags = eis + 128;
        float* acc = reinterpret_cast<float*>(smem + WS_ACC + g * 3072);
        const uint32_t* stat = reinterpret_cast<const uint32_t*>(smem + WS_STAT + g * 512);
        bool alive = false;
        RayGeom G;
        float t = 0.f, fwx = 0.f, fwy = 0.f, fwz = 0.f;
        uint32_t ei = 0, pd = 0;
        // role statistics (profiling runs only): cycles spent per phase, measured by lane 0 of every warp
        const bool st = STATS && lane == 0;
        long long c_wait = 0, c_walk = 0, c_enc = 0, c_refill = 0, c_rounds = 0;
        const long long c_begin = st ? clock64() : 0;
        // the ray's next (up to) two samples: position in the network's input space, depth and step for compositing
        int n_s = 0;
        bool exits = false, walked = false;
        float wp0x = 0.f, wp0y = 0.f, wp0z = 0.f, wp1x = 0.f, wp1y = 0.f, wp1z = 0.f;
        float dep0 = 0.f, dep1 = 0.f, dtu0 = 0.f, dtu1 = 0.f;
        // One walk loop for both samples (generate_next_nerf_network_inputs, testbed_nerf.cu:454-467, twice): the same sequence
        // of operations on t as two calls of if_unoccupied_advance_to_next_occupied_voxel.  The next sample positions depend on
        // the occupancy grid only, never on the network output, so a ray that may go on walks to them while its MLP round is
        // still in flight; if the round turns out to saturate it, the walk was for nothing (like the unused tail of the
        // reference's n_steps).
        auto do_walk = [&]() {
            n_s = 0;
            exits = false;
            while (true) {
                const float px = G.ox + t * G.dx, py = G.oy + t * G.dy, pz = G.oz + t * G.dz;
                if (t >= MAX_DEPTH() || t > G.t_exit || !raabb_contains(M, px, py, pz)) { exits = true; break; }
                uint32_t mip = min(max(mip_from_pos(px, py, pz), 0u), max_mip);
                if (density_grid_occupied_at(px, py, pz, M.bitfield_lin, mip)) {
                    const float dt = calc_dt(t, cone);
                    const float wx = (px - M.aabb_min[0]) / M.aabb_diag[0];
                    const float wy = (py - M.aabb_min[1]) / M.aabb_diag[1];
                    const float wz = (pz - M.aabb_min[2]) / M.aabb_diag[2];
                    // composite_kernel_nerf reads the position back from the network input (unwarp_position) and the
                    // step from warp_dt/unwarp_dt: same arithmetic here, evaluated before the MLP instead of after
                    const float ux = M.aabb_min[0] + wx * M.aabb_diag[0];
                    const float uy = M.aabb_min[1] + wy * M.aabb_diag[1];
                    const float uz = M.aabb_min[2] + wz * M.aabb_diag[2];
                    float dep = 0.f;
                    dep += fwx * (ux - G.ox); dep += fwy * (uy - G.oy); dep += fwz * (uz - G.oz);
                    dep *= M.depth_scale;
                    const float dtu = unwarp_dt(warp_dt(dt));
                    if (n_s == 0) { wp0x = wx; wp0y = wy; wp0z = wz; dep0 = dep; dtu0 = dtu; }
                    else { wp1x = wx; wp1y = wy; wp1z = wz; dep1 = dep; dtu1 = dtu; }
                    t += dt;
                    if (++n_s == 2) break;
                    continue;
                }
                while (mip < max_mip && !density_grid_occupied_at(px, py, pz, M.bitfield_lin, mip + 1)) ++mip;
                t = advance_to_next_voxel(t, cone, px, py, pz, G.dx, G.dy, G.dz, G.ix, G.iy, G.iz, mip);
            }
        };
        while (true) {
            long long c0 = st ? clock64() : 0;
            // ---- A. claim hit-list entries: one thread fetches a new chunk when the group's current one is used up ----
            if (r == 0 && mi->cur[g] >= mi->end[g] && !mi->done[g]) {
                const uint32_t base = atomicAdd(P.entry_cursor, (uint32_t)TC_CHUNK);
                if (base >= total_entries) mi->done[g] = 1;
                else { mi->cur[g] = base; mi->end[g] = min(base + (uint32_t)TC_CHUNK, total_entries); }
            }
            group_sync(bar_id);
            const bool done = mi->done[g] != 0;
            {   // warp-aggregated refill: the dead lanes of a warp take consecutive entries
                const uint32_t dead = __ballot_sync(0xffffffffu, !alive);
                if (dead) {
                    uint32_t base = 0;
                    const int leader = __ffs(dead) - 1;
                    if (lane == leader) base = atomicAdd(&mi->cur[g], (uint32_t)__popc(dead));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    const uint32_t i = base + __popc(dead & ((1u << lane) - 1));
                    if (!alive && i < mi->end[g]) {
                        const uint32_t id = resumed ? P.work_list[i] : i;
                        const RayEntry e = P.entries[id];
                        const Mat3x4 C = P.cams[e.k];
                        ray_geom_only(C, __ldg(P.dirs + e.idx), G);          // same arithmetic as pass 1 (setup_ray)
                        G.t_exit = e.t_exit;
                        ei = id;
                        t = resumed ? P.t_cur[id] : e.t;
                        fwx = C.c[2][0]; fwy = C.c[2][1]; fwz = C.c[2][2];
                        float sh[16];
                        const float wx = (G.dx + 1.0f) * 0.5f, wy = (G.dy + 1.0f) * 0.5f, wz = (G.dz + 1.0f) * 0.5f;
                        sh_enc4(wx * 2.f - 1.f, wy * 2.f - 1.f, wz * 2.f - 1.f, sh);
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            __half2 h = __floats2half2_rn(sh[2 * j], sh[2 * j + 1]);
                            pk[j] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        sh_lo[r] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        sh_hi[r] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        if (resumed) {      // a ray the split kernels brought this far: it has taken exactly resume_steps samples
                            const float4 a = P.resume_acc4[i];
                            acc[0 * 128 + r] = a.x; acc[1 * 128 + r] = a.y; acc[2 * 128 + r] = a.z; acc[3 * 128 + r] = a.w;
                            acc[4 * 128 + r] = P.resume_acca[i];
                            acc[5 * 128 + r] = __int_as_float(P.resume_steps);
                        } else {
#pragma unroll
                            for (int j = 0; j < 6; ++j) acc[j * 128 + r] = 0.f;
                            ++my_rays;
                        }
                        alive = true;
                        walked = false;
                    }
                }
            }
            // ---- B. up to two sample positions (rays that continue did this walk while their last MLP round was in flight) ----
            if (st) { const long long c1 = clock64(); c_refill += c1 - c0; c0 = c1; }
            if (alive && !walked) do_walk();
            if (!alive) { n_s = 0; exits = false; }
            if (alive) {
                if (n_s == 0) {      // nothing left to sample: the ray is finished as it stands (accumulators of the last round)
                    P.res_rgbd[ei] = make_float4(acc[0 * 128 + r], acc[1 * 128 + r], acc[2 * 128 + r], acc[3 * 128 + r]);
                    P.res_a[ei] = acc[4 * 128 + r];
                    if (P.res_n) P.res_n[ei] = (float)(__float_as_int(acc[5 * 128 + r]) + 1);      // an exhausted ray: see the epilogue
                    alive = false;
                }
            }
            // ---- C. hash-grid features -> this slot's rows of the group's two A tiles ----
            __syncwarp();
            if (st) { const long long c1 = clock64(); c_walk += c1 - c0; c0 = c1; }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s < n_s) {
                    const float sx = s ? wp1x : wp0x, sy = s ? wp1y : wp0y, sz = s ? wp1z : wp0z;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        __half2 f[4];
                        encode_levels<2>(M, 2 * c, sx, sy, sz, f);
                        uint4 v;
                        v.x = *reinterpret_cast<uint32_t*>(&f[0]); v.y = *reinterpret_cast<uint32_t*>(&f[1]);
                        v.z = *reinterpret_cast<uint32_t*>(&f[2]); v.w = *reinterpret_cast<uint32_t*>(&f[3]);
                        *reinterpret_cast<uint4*>(rowA32 + s * 16384 + c * 128) = v;
                    }
                }
            }
            aux[r] = make_float4(dep0, dtu0, dep1, dtu1);
            eis[r] = ei;
            flags[r] = (uint32_t)n_s | (exits ? 16u : 0u) | (alive ? 64u : 0u);
            fence_proxy_async();
            __syncwarp();
            if (st) { const long long c1 = clock64(); c_enc += c1 - c0; c0 = c1; ++c_rounds; }
            if (!group_or(bar_id, n_s > 0)) {
                if (done) break;
                continue;
            }
            mbar_arrive(&mi->bar_a[g]);
            // the published samples are with the MMA / epilogue warps now: walk to the next two in the meantime
            walked = false;
            if (alive && n_s == 2 && !exits) { do_walk(); walked = true; }
            __syncwarp();
            if (st) { const long long c1 = clock64(); c_walk += c1 - c0; c0 = c1; }
            mbar_wait_wd(&mi->bar_done[g], pd); pd ^= 1;
            if (alive && !stat[r]) alive = false;
            if (st) c_wait += clock64() - c0;
        }
        if (st) {
            atomicAdd(P.prof + 2, (unsigned long long)(clock64() - c_begin));
            atomicAdd(P.prof + 3, (unsigned long long)c_wait);
            atomicAdd(P.prof + 4, (unsigned long long)c_walk);
            atomicAdd(P.prof + 5, (unsigned long long)c_enc);
            atomicAdd(P.prof + 6, (unsigned long long)c_refill);
            atomicAdd(P.prof + 12, (unsigned long long)c_rounds);
        }
        // tell the MMA thread that this group is gone
        if (r == 0) mi->exit_[g] = 1;
        group_sync(bar_id);
        mbar_arrive(&mi->bar_a[g]);
    } else if (warp < WS_NG * 4 + 4) {
        // ================================================ EPILOGUE ================================================
        const int e = warp - WS_NG * 4, r = e * 32 + lane;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(e * 32) << 16);
        uint32_t seq = 0;
        const bool st = STATS && lane == 0;
        long long c_wait = 0;
        const long long c_begin = st ? clock64() : 0;
        while (true) {
            const uint32_t slot = seq % WS_RING;
            const long long c0 = st ? clock64() : 0;
            mbar_wait_wd(&mi->bar_seq[slot], (seq / WS_RING) & 1u);
            if (st) c_wait += clock64() - c0;
            uint32_t item = *reinterpret_cast<const volatile uint32_t*>(&mi->ring[slot]);
            if ((item >> 16) != (seq & 0xffffu)) {
                // the entry was written before the commit was issued; the tag makes reading it independent of how the commit's arrive is ordered
                const volatile uint32_t* rp = &mi->ring[slot];
                long long t0 = 0;
                while (((item = *rp) >> 16) != (seq & 0xffffu)) {
                    const long long now = clock64();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > 4000000000ll) __trap();
                }
            }
            ++seq;
            const uint32_t g = item & 0xffu, layer = (item >> 8) & 0xffu;
            if (g == WS_EXIT) {
                if (st) {
                    atomicAdd(P.prof + 7, (unsigned long long)(clock64() - c_begin));
                    atomicAdd(P.prof + 8, (unsigned long long)c_wait);
                    atomicAdd(P.prof + 9, (unsigned long long)(seq - 1));
                }
                break;
            }
            tc_fence_after();
            unsigned char* tileA = smem + WS_A + g * 32768;
            unsigned char* rowA32 = tileA + umma_chunk_off(r, 0, 32);
            unsigned char* rowA64 = tileA + umma_chunk_off(r, 0, 64);
            const uint32_t tm = tmem_lane + g * 128;
            float* sig = reinterpret_cast<float*>(smem + WS_SIG + g * 1024);
            if (layer == 0 || layer == 2 || layer == 3) {
                // 64 outputs, ReLU, fp16 -> K=64 operand rows (in place for layer 3: its MMA has retired).  Four chunks of 32
                // columns, software-pipelined: chunk q+1 is on its way from TMEM while chunk q is packed and stored
                uint32_t va[32], vb[32];
                auto pack_store = [&](const uint32_t (&v)[32], int q) {
                    unsigned char* dst = rowA64 + (q >> 1) * 16384 + (q & 1) * 512;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 o;
                        o.x = pack_relu_h2(v[8 * c + 0], v[8 * c + 1], true); o.y = pack_relu_h2(v[8 * c + 2], v[8 * c + 3], true);
                        o.z = pack_relu_h2(v[8 * c + 4], v[8 * c + 5], true); o.w = pack_relu_h2(v[8 * c + 6], v[8 * c + 7], true);
                        *reinterpret_cast<uint4*>(dst + c * 128) = o;
                    }
                };
                tmem_ld_32x32_x32(tm, va);
                tmem_ld_wait();
                tmem_ld_32x32_x32(tm + 32, vb);
                pack_store(va, 0);
                tmem_ld_wait();
                tmem_ld_32x32_x32(tm + 64, va);
                pack_store(vb, 1);
                tmem_ld_wait();
                tmem_ld_32x32_x32(tm + 96, vb);
                pack_store(va, 2);
                tmem_ld_wait();
                pack_store(vb, 3);
            } else if (layer == 1) {
                // density layer 1: 16 outputs (row 0 = raw density); rgb input = [16 density-out | 16 SH]
                const uint4 s0 = reinterpret_cast<const uint4*>(smem + WS_SH + g * 4096)[r];
                const uint4 s1 = reinterpret_cast<const uint4*>(smem + WS_SH + g * 4096)[128 + r];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    uint32_t v[16];
                    tmem_ld_32x32_x16(tm + s * 64, v);
                    tmem_ld_wait();
                    sig[s * 128 + r] = h2f_round(__uint_as_float(v[0]));
                    uint4 v0, v1;
                    v0.x = pack_relu_h2(v[0], v[1], false); v0.y = pack_relu_h2(v[2], v[3], false);
                    v0.z = pack_relu_h2(v[4], v[5], false); v0.w = pack_relu_h2(v[6], v[7], false);
                    v1.x = pack_relu_h2(v[8], v[9], false); v1.y = pack_relu_h2(v[10], v[11], false);
                    v1.z = pack_relu_h2(v[12], v[13], false); v1.w = pack_relu_h2(v[14], v[15], false);
                    unsigned char* row = rowA32 + s * 16384;
                    *reinterpret_cast<uint4*>(row + 0) = v0;
                    *reinterpret_cast<uint4*>(row + 128) = v1;
                    *reinterpret_cast<uint4*>(row + 256) = s0;
                    *reinterpret_cast<uint4*>(row + 384) = s1;
                }
            }
            if (layer < 4) {
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&mi->bar_a[g]);
                continue;
            }
            // ---- rgb output layer (3 of 16 used) + composite_kernel_nerf (testbed_nerf.cu:511-667), sample 0 then sample 1 ----
            float raw[2][3];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint32_t v[16];
                tmem_ld_32x32_x16(tm + s * 64, v);
                tmem_ld_wait();
                raw[s][0] = h2f_round(__uint_as_float(v[0])); raw[s][1] = h2f_round(__uint_as_float(v[1])); raw[s][2] = h2f_round(__uint_as_float(v[2]));
            }
            const uint32_t fl = reinterpret_cast<const uint32_t*>(smem + WS_META + g * 1024)[128 + r];
            const int n_s = (int)(fl & 15u);
            bool alive = (fl & 64u) != 0;
            if (alive) {
                const uint32_t ei = reinterpret_cast<const uint32_t*>(smem + WS_META + g * 1024)[r];
                const float4 ax = reinterpret_cast<const float4*>(smem + WS_AUX + g * 2048)[r];
                float* acc = reinterpret_cast<float*>(smem + WS_ACC + g * 3072);
                float cr = acc[0 * 128 + r], cg = acc[1 * 128 + r], cb = acc[2 * 128 + r], cd = acc[3 * 128 + r], ca = acc[4 * 128 + r];
                int n_steps = __float_as_int(acc[5 * 128 + r]);
                bool fin = false, saturated = false;
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (alive && s < n_s) {
                        ++my_samples;                  // composited samples only (a dropped sample 1 is not counted)
                        ++n_steps;
                        const float T = 1.f - ca;
                        const float alpha = 1.f - __expf(-__expf(sig[s * 128 + r]) * (s == 0 ? ax.y : ax.w));
                        const float weight = alpha * T;
                        const float rr = logistic_d(raw[s][0]), gg = logistic_d(raw[s][1]), bb_ = logistic_d(raw[s][2]);
                        const float dep = s == 0 ? ax.x : ax.z;
                        cr += rr * weight; cg += gg * weight; cb += bb_ * weight; cd += dep * weight; ca += weight;
                        if (ca > (1.0f - M.min_transmittance)) {
                            cr /= ca; cg /= ca; cb /= ca; cd /= ca; ca /= ca;
                            alive = false; fin = true; saturated = true;
                        } else if (n_steps >= MARCH_ITER - 1) {
                            cr = cg = cb = cd = ca = 0.f;      // never reaches the hit buffer in the reference
                            alive = false; fin = true;
                        }
                    }
                }
                if (alive && (fl & 16u)) { alive = false; fin = true; }      // ran out of occupied cells after the prepared samples
                if (fin) {
                    P.res_rgbd[ei] = make_float4(cr, cg, cb, cd);
                    P.res_a[ei] = ca;
                    // payload.n_steps as the reference leaves it (testbed_nerf.cu:669-672 with current_step starting at 1, :1674):
                    // the samples composited for a ray that saturates, one more for a ray that runs out of occupied cells
                    if (P.res_n) P.res_n[ei] = (float)(n_steps + (saturated ? 0 : 1));
                } else {
                    acc[0 * 128 + r] = cr; acc[1 * 128 + r] = cg; acc[2 * 128 + r] = cb; acc[3 * 128 + r] = cd; acc[4 * 128 + r] = ca;
                    acc[5 * 128 + r] = __int_as_float(n_steps);
                }
            }
            reinterpret_cast<uint32_t*>(smem + WS_STAT + g * 512)[r] = alive ? 1u : 0u;
            tc_fence_before();
            mbar_arrive(&mi->bar_done[g]);
        }
    } else {
      // =================================================== MMA ==================================================
      if (lane == 0) {
        mbar_arrive_expect_tx(&mi->bar_w, 20480u);
        bulk_g2s(smem + WS_W, M.w_umma, 20480u, &mi->bar_w);
        mbar_wait_wd(&mi->bar_w, 0);
        const uint32_t wbase = smem_u32(smem + WS_W), abase = smem_u32(smem + WS_A);
        // per-group state packed into registers (dynamic indexing of local arrays would go through local memory):
        // layers: 4 bits per group = next layer to issue; pa: parity bit per group of bar_a; gone: bit per group
        uint32_t layers = 0, pa = 0, gone = 0;
        int n_active = WS_NG;
        uint32_t seq = 0;
        long long t_idle = 0;
        const long long c_begin = STATS ? clock64() : 0;
        long long c_issue = 0;
        while (n_active) {
            bool any = false;
#pragma unroll 1
            for (int g = 0; g < WS_NG; ++g) {
                if ((gone >> g) & 1u) continue;
                if (!mbar_test(&mi->bar_a[g], (pa >> g) & 1u)) continue;
                any = true;
                pa ^= 1u << g;
                const uint32_t L = (layers >> (4 * g)) & 15u;
                if (L == 0 && *reinterpret_cast<volatile uint32_t*>(&mi->exit_[g])) { gone |= 1u << g; --n_active; continue; }
                tc_fence_after();
                const long long ci = STATS ? clock64() : 0;
                const uint32_t slot = seq % WS_RING;
                *reinterpret_cast<volatile uint32_t*>(&mi->ring[slot]) = ((seq & 0xffffu) << 16) | (L << 8) | (uint32_t)g;
                // descriptor low words: (address >> 4) | LBO (128 B >> 4) << 16; all operand addresses are below 256 KB
                const uint32_t a_lo = (abase >> 4) + (uint32_t)g * 2048u + (8u << 16), w_lo = (wbase >> 4) + (8u << 16);
                const uint32_t td = tmem_base + (uint32_t)g * 128u;
                switch (L) {
                    case 0: issue_layer_ws<32, 64>(a_lo, w_lo + (W_D0 >> 4), td); break;
                    case 1: issue_layer_ws<64, 16>(a_lo, w_lo + (W_D1 >> 4), td); break;
                    case 2: issue_layer_ws<32, 64>(a_lo, w_lo + (W_C0 >> 4), td); break;
                    case 3: issue_layer_ws<64, 64>(a_lo, w_lo + (W_C1 >> 4), td); break;
                    default: issue_layer_ws<64, 16>(a_lo, w_lo + (W_C2 >> 4), td); break;
                }
                tc_commit(&mi->bar_seq[slot]);
                if (STATS) c_issue += clock64() - ci;
                ++seq;
                layers = (layers & ~(15u << (4 * g))) | ((L == 4 ? 0u : L + 1u) << (4 * g));
            }
            if (any) t_idle = 0;
            else {
                const long long now = clock64();
                if (t_idle == 0) t_idle = now;
                else if (now - t_idle > 4000000000ll) __trap();
            }
        }
        if (STATS) {
            atomicAdd(P.prof + 10, (unsigned long long)(clock64() - c_begin));
            atomicAdd(P.prof + 11, (unsigned long long)c_issue);
        }
        // every group is gone, every earlier item has been consumed: hand the epilogue warps their exit item
        const uint32_t slot = seq % WS_RING;
        *reinterpret_cast<volatile uint32_t*>(&mi->ring[slot]) = ((seq & 0xffffu) << 16) | WS_EXIT;
        __threadfence_block();
        mbar_arrive(&mi->bar_seq[slot]);
      }
      __syncwarp();      // the other 31 lanes park here: the whole warp reaches the CTA barrier below together
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WS_NG * 4 + 4) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
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
