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
// Work decomposition (launch_march below): the HOST plans every candidate's conservative screen rectangle (the pixels whose
// ray can touch an occupied density-grid cell) from the camera matrices it is handed anyway and the view's direction ranges,
// cuts the rectangles into 16x8 tiles and uploads cameras + rectangles + tile prefix with one pinned asynchronous copy -- so
// the tile count is known without a read-back and a launch never blocks the host.  On the device:
//   k_fill_frames  every frame starts as the composited background (pure streaming stores)
//   k_classify     one thread per tile pixel: ray, Sobol start jitter, walk to the first occupied sample -> hit list
//   k_march_ws     ONE persistent warp-specialised kernel takes every hit-list ray to its end (d2r_march_ws.cuh);
//                  D2R_MARCH=split selects the round-1 pair of per-round kernels instead (d2r_march_split.cuh, A/B only:
//                  that path reads a live count back every 4th round)
//   k_finish       per-ray accumulators -> pixels (keep rule, shade / tonemap blend, depth-test composite, sRGB, u8)
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "d2r_march_common.cuh"
#include "d2r_march_ws.cuh"
#include "d2r_march_split.cuh"

namespace d2r {

__global__ void k_tile_map(int K, const uint32_t* __restrict__ prefix, uint16_t* __restrict__ map) {
    const int k = blockIdx.x;
    for (uint32_t t = prefix[k] + threadIdx.x; t < prefix[k + 1]; t += blockDim.x) map[t] = (uint16_t)k;
}

// Every frame starts as a copy of the composited background (u8) / the constant background blend (float);
// the march kernel then overwrites the pixels of the candidate's rectangle.  Pure streaming stores:
// 128-bit vectors whenever a frame is a whole number of 16-byte words.
__global__ void k_fill_frames(int K, int W, int H, float4 shade_bg, float4 depth_bg, float4* __restrict__ rgba_out,
                              float4* __restrict__ depth_out, float* __restrict__ cost_out, const uint8_t* __restrict__ bg_u8,
                              uint8_t* __restrict__ u8_out, int vec_ok) {
    const int k = blockIdx.x;
    const size_t npx = (size_t)W * H;
    const size_t stride = (size_t)gridDim.y * blockDim.x, first = (size_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (rgba_out) for (size_t p = first; p < npx; p += stride) rgba_out[(size_t)k * npx + p] = shade_bg;
    if (depth_out) for (size_t p = first; p < npx; p += stride) depth_out[(size_t)k * npx + p] = depth_bg;
    if (cost_out) for (size_t p = first; p < npx; p += stride) cost_out[(size_t)k * npx + p] = 0.f;
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
    // a pixel without a foreground ray: the fg Depth render holds the background blend there too (tonemap_kernel), so its
    // depth channel is fg_empty.x, exactly what k_finish composites for a ray inside the rectangle that hits nothing
    composite_pixel(fg_empty, fg_empty.x, bg_rgba[p], bg_depth[p], out + (size_t)p * 3);
}

// ---- profiling: CUDA events around the march kernel only (bench.py roofline) ----------------------
struct Prof {
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    size_t used = 0;
    unsigned long long* counters = nullptr;   // device {samples, rays, items}
};
static Prof g_prof[16];

// ---- per (device, stream) launch context -----------------------------------------------------------
// Everything a launch needs besides the caller's buffers.  Keyed by stream, so two host threads driving two streams of one GPU
// never share scratch (kernels of one stream run in order, so one set per stream is enough); the map itself is guarded.
constexpr int RING = 8;              // pinned upload slots: the host may run this many launches ahead of the device
struct Ctx {
    int device = 0;
    int n_sm = 0;
    // upload: [cams K x Mat3x4 | bbox K x int4 | prefix (K+1) x u32], one pinned slot per launch in flight, one device copy
    void* up_host[RING] = {nullptr};
    cudaEvent_t up_done[RING] = {nullptr};
    bool up_used[RING] = {false};
    size_t up_cap = 0;
    unsigned up_next = 0;
    unsigned char* up_dev = nullptr;
    uint8_t* bg_u8 = nullptr; size_t cap_bg = 0;
    RayEntry* entries = nullptr; size_t cap_entries = 0; uint32_t* entry_counters = nullptr;
    uint16_t* tile_cand = nullptr; size_t cap_tile_cand = 0;
    float4* res_rgbd = nullptr; float* res_a = nullptr; float* res_n = nullptr;
    // round-based split path (D2R_MARCH=split)
    size_t cap_split = 0;
    unsigned char* sp_feat = nullptr; float2* sp_aux = nullptr; uint4* sp_shb = nullptr; uint8_t* sp_nsb = nullptr;
    float* sp_t = nullptr; uint32_t* sp_live[2] = {nullptr, nullptr}; uint32_t* sp_cnt = nullptr;
    float4* sp_acc4[2] = {nullptr, nullptr}; float* sp_acca[2] = {nullptr, nullptr};
    // {hits, hit-list slots} of the latest finished launch, written by k_finish into host-mapped memory: sizes the per-round
    // buffers of the split kernels by what launches actually hit instead of by their screen rectangles (no read-back: the
    // host just looks at what has landed; a wrong guess costs speed, never correctness -- SplitParams::cap)
    volatile uint32_t* fb_host = nullptr; uint32_t* fb_dev = nullptr;
    double hit_ratio = 0.6;
    size_t cap_t = 0;
};
static std::mutex g_ctx_mutex;
static std::map<std::pair<int, cudaStream_t>, Ctx*> g_ctx;

static Ctx* get_ctx(int device, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    Ctx*& c = g_ctx[std::make_pair(device, stream)];
    if (!c) { c = new Ctx(); c->device = device; }
    return c;
}

static size_t upload_bytes(int K) { return (size_t)K * (sizeof(Mat3x4) + sizeof(int4)) + (size_t)(K + 1) * sizeof(uint32_t); }

static int ensure_ctx(Ctx& s, int K, int W, int H) {
    if (!s.n_sm) D2R_CUDA(cudaDeviceGetAttribute(&s.n_sm, cudaDevAttrMultiProcessorCount, s.device));
    if (upload_bytes(K) > s.up_cap) {
        for (int i = 0; i < RING; ++i) {
            if (s.up_used[i]) D2R_CUDA(cudaEventSynchronize(s.up_done[i]));      // growth only: earlier uploads must have left their slot
            if (s.up_host[i]) cudaFreeHost(s.up_host[i]);
            s.up_host[i] = nullptr; s.up_used[i] = false;
        }
        if (s.up_dev) cudaFree(s.up_dev);
        s.up_dev = nullptr; s.up_cap = 0;
        const size_t cap = upload_bytes(std::max(K, 1024));
        for (int i = 0; i < RING; ++i) {
            D2R_CUDA(cudaMallocHost(&s.up_host[i], cap));
            if (!s.up_done[i]) D2R_CUDA(cudaEventCreateWithFlags(&s.up_done[i], cudaEventDisableTiming));
        }
        D2R_CUDA(cudaMalloc(&s.up_dev, cap));
        s.up_cap = cap;
    }
    if (!s.entry_counters) D2R_CUDA(cudaMalloc(&s.entry_counters, 2 * sizeof(uint32_t)));
    if (!s.fb_host) {
        void* h = nullptr;
        D2R_CUDA(cudaHostAlloc(&h, 2 * sizeof(uint32_t), cudaHostAllocMapped));
        memset(h, 0, 2 * sizeof(uint32_t));
        s.fb_host = (volatile uint32_t*)h;
        D2R_CUDA(cudaHostGetDevicePointer((void**)&s.fb_dev, h, 0));
    }
    if ((size_t)W * H * 3 > s.cap_bg) {
        cudaFree(s.bg_u8);
        s.bg_u8 = nullptr; s.cap_bg = 0;
        D2R_CUDA(cudaMalloc(&s.bg_u8, (size_t)W * H * 3));
        s.cap_bg = (size_t)W * H * 3;
    }
    return D2R_OK;
}

// Conservative screen rectangle of one candidate, planned on the host: contains every pixel whose (undistorted) camera-plane
// direction lies inside the perspective projection of the box around all occupied cells.  col_lo/col_hi (row_lo/row_hi) are
// the per-column (per-row) min/max of the direction table, so lens distortion is handled exactly.
static void plan_rect(const float* cam /* [3][4] row-major, NGP convention */, int W, int H, const float* col_lo, const float* col_hi,
                      const float* row_lo, const float* row_hi, const float* occ_min, const float* occ_max, int4& bb) {
    int x0 = 0, y0 = 0, x1 = W - 1, y1 = H - 1;
    float u0 = 1e30f, u1 = -1e30f, v0 = 1e30f, v1 = -1e30f;
    bool behind = false;
    for (int c = 0; c < 8 && !behind; ++c) {
        const float wx = ((c & 1) ? occ_max[0] : occ_min[0]) - cam[3];
        const float wy = ((c & 2) ? occ_max[1] : occ_min[1]) - cam[7];
        const float wz = ((c & 4) ? occ_max[2] : occ_min[2]) - cam[11];
        // camera-space = R^T * (p - o)   (columns of the matrix are the camera axes)
        const float cx = cam[0] * wx + cam[4] * wy + cam[8] * wz;
        const float cy = cam[1] * wx + cam[5] * wy + cam[9] * wz;
        const float cz = cam[2] * wx + cam[6] * wy + cam[10] * wz;
        if (cz < 1e-3f) { behind = true; break; }
        const float u = cx / cz, v = cy / cz;
        u0 = fminf(u0, u); u1 = fmaxf(u1, u); v0 = fminf(v0, v); v1 = fmaxf(v1, v);
    }
    if (!behind) {      // a corner behind the camera: keep the whole frame
        const float eu = 1e-4f * (1.f + fmaxf(fabsf(u0), fabsf(u1))), ev = 1e-4f * (1.f + fmaxf(fabsf(v0), fabsf(v1)));
        u0 -= eu; u1 += eu; v0 -= ev; v1 += ev;
        x0 = W; x1 = -1; y0 = H; y1 = -1;
        for (int x = 0; x < W; ++x) if (col_hi[x] >= u0 && col_lo[x] <= u1) { x0 = std::min(x0, x); x1 = std::max(x1, x); }
        for (int y = 0; y < H; ++y) if (row_hi[y] >= v0 && row_lo[y] <= v1) { y0 = std::min(y0, y); y1 = std::max(y1, y); }
        if (x1 >= x0 && y1 >= y0) {
            x0 = std::max(x0 - 1, 0); y0 = std::max(y0 - 1, 0); x1 = std::min(x1 + 1, W - 1); y1 = std::min(y1 + 1, H - 1);
        }
    }
    if (x1 >= x0 && y1 >= y0) bb = make_int4(x0, y0, x1, y1);
    else bb = make_int4(0, 0, -1, -1);
}

int launch_march(const d2r_model* m, const d2r_view* v, const float* cams_ngp_host, int K, const float bg[4],
                 float* rgba_out, float* depth_out, const float* bg_rgba, const float* bg_depth, uint8_t* u8_out,
                 unsigned long long* n_samples, cudaStream_t stream, int* rects_out = nullptr, uint8_t* bg_u8_out = nullptr,
                 float* cost_out = nullptr) {
    D2R_REQUIRE(m && v && cams_ngp_host && bg, "render: null argument");
    D2R_REQUIRE(K > 0, "render: K must be positive");
    D2R_REQUIRE(K <= 65535, "render: at most 65535 candidates per launch");
    D2R_REQUIRE(m->device == v->device && m->device < 16, "render: model and view live on different devices");
    D2R_REQUIRE(rgba_out || depth_out || u8_out || cost_out, "render: no output requested");
    D2R_REQUIRE(!u8_out || (bg_rgba && bg_depth), "render_composite: background buffers missing");
    DeviceGuard dg(m->device);
    const int W = v->W, H = v->H;
    Ctx& s = *get_ctx(m->device, stream);
    int rc = ensure_ctx(s, K, W, H);
    if (rc) return rc;
    const ModelDev& M = m->dev;

    // ---- host plan -> one pinned asynchronous upload (no read-back anywhere in this function) ----
    const unsigned slot = s.up_next++ % RING;
    if (s.up_used[slot]) D2R_CUDA(cudaEventSynchronize(s.up_done[slot]));      // back-pressure only: RING launches ahead of the device
    unsigned char* up = (unsigned char*)s.up_host[slot];
    Mat3x4* h_cams = (Mat3x4*)up;
    int4* h_bbox = (int4*)(up + (size_t)K * sizeof(Mat3x4));
    uint32_t* h_prefix = (uint32_t*)(up + (size_t)K * (sizeof(Mat3x4) + sizeof(int4)));
    uint32_t total_tiles = 0;
    {
        const float* col_lo = v->ranges_host, *col_hi = col_lo + W, *row_lo = col_lo + 2 * W, *row_hi = col_lo + 2 * W + H;
        for (int k = 0; k < K; ++k) {
            const float* c = cams_ngp_host + (size_t)k * 12;      // [3,4] row-major: rows = xyz, columns = 3 axes + origin
            for (int col = 0; col < 4; ++col)
                for (int rr = 0; rr < 3; ++rr) h_cams[k].c[col][rr] = c[rr * 4 + col];
            plan_rect(c, W, H, col_lo, col_hi, row_lo, row_hi, M.occ_min, M.occ_max, h_bbox[k]);
            h_prefix[k] = total_tiles;
            const int4 bb = h_bbox[k];
            if (bb.z >= bb.x && bb.w >= bb.y) total_tiles += (uint32_t)((bb.z - bb.x + TILE_W) / TILE_W) * (uint32_t)((bb.w - bb.y + TILE_H) / TILE_H);
        }
        h_prefix[K] = total_tiles;
    }
    D2R_CUDA(cudaMemcpyAsync(s.up_dev, up, upload_bytes(K), cudaMemcpyHostToDevice, stream));
    D2R_CUDA(cudaEventRecord(s.up_done[slot], stream));
    s.up_used[slot] = true;
    const Mat3x4* d_cams = (const Mat3x4*)s.up_dev;
    const int4* d_bbox = (const int4*)(s.up_dev + (size_t)K * sizeof(Mat3x4));
    const uint32_t* d_prefix = (const uint32_t*)(s.up_dev + (size_t)K * (sizeof(Mat3x4) + sizeof(int4)));
    if (rects_out) D2R_CUDA(cudaMemcpyAsync(rects_out, d_bbox, (size_t)K * sizeof(int4), cudaMemcpyDeviceToDevice, stream));

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
        k_fill_frames<<<grid, 256, 0, stream>>>(K, W, H, empty, empty, (float4*)rgba_out, (float4*)depth_out, cost_out, s.bg_u8, u8_out, vec_ok);
        count_launch();
    }
    Prof& pf = g_prof[m->device];
    std::pair<cudaEvent_t, cudaEvent_t>* evp = nullptr;
    if (pf.on) {
        if (pf.used == pf.ev.size()) {
            cudaEvent_t a, b;
            D2R_CUDA(cudaEventCreate(&a));
            D2R_CUDA(cudaEventCreate(&b));
            pf.ev.emplace_back(a, b);
        }
        evp = &pf.ev[pf.used++];
    }
    if (!total_tiles) {      // nothing in view: the frames are the background
        if (evp) { D2R_CUDA(cudaEventRecord(evp->first, stream)); D2R_CUDA(cudaEventRecord(evp->second, stream)); }
        D2R_CUDA(cudaGetLastError());
        return D2R_OK;
    }

    // hit list: at most one entry per tile pixel (grown with head-room; growth is the only time a launch touches the allocator)
    const size_t need = (size_t)total_tiles * CTA;
    if (need > s.cap_entries) {
        if (s.entries) { cudaFree(s.entries); cudaFree(s.res_rgbd); cudaFree(s.res_a); cudaFree(s.res_n); }
        s.entries = nullptr; s.res_rgbd = nullptr; s.res_a = nullptr; s.res_n = nullptr; s.cap_entries = 0;
        const size_t cap = need + need / 4;
        D2R_CUDA(cudaMalloc(&s.entries, cap * sizeof(RayEntry)));
        D2R_CUDA(cudaMalloc(&s.res_rgbd, cap * sizeof(float4)));
        D2R_CUDA(cudaMalloc(&s.res_a, cap * sizeof(float)));
        D2R_CUDA(cudaMalloc(&s.res_n, cap * sizeof(float)));
        s.cap_entries = cap;
    }
    if (total_tiles > s.cap_tile_cand) {
        if (s.tile_cand) cudaFree(s.tile_cand);
        s.tile_cand = nullptr; s.cap_tile_cand = 0;
        const size_t cap = (size_t)total_tiles + total_tiles / 4;
        D2R_CUDA(cudaMalloc(&s.tile_cand, cap * sizeof(uint16_t)));
        s.cap_tile_cand = cap;
    }
    MarchParams P;
    P.M = M; P.dirs = v->dirs_dev; P.W = W; P.H = H; P.cams = d_cams; P.K = K; P.bbox = d_bbox; P.tile_prefix = d_prefix;
    for (int i = 0; i < 4; ++i) P.bg[i] = bg[i];
    P.rgba_out = (float4*)rgba_out; P.depth_out = (float4*)depth_out; P.bg_rgba = (const float4*)bg_rgba; P.bg_depth = bg_depth;
    P.u8_out = u8_out; P.n_samples = n_samples;
    P.prof = pf.on ? pf.counters : nullptr;
    P.entries = s.entries; P.n_entries = s.entry_counters; P.entry_cursor = s.entry_counters + 1;
    P.res_rgbd = s.res_rgbd; P.res_a = s.res_a;
    P.res_n = cost_out ? s.res_n : nullptr; P.cost_out = cost_out;
    P.tile_cand = s.tile_cand;
    D2R_CUDA(cudaMemsetAsync(s.entry_counters, 0, 2 * sizeof(uint32_t), stream));
    k_tile_map<<<K, 64, 0, stream>>>(K, d_prefix, s.tile_cand);
    k_classify<<<total_tiles, CTA, 0, stream>>>(P);
    count_launch(2);
    if (evp) D2R_CUDA(cudaEventRecord(evp->first, stream));   // time the march kernel alone

    // Large launches: the first rounds -- where nearly every ray is still alive -- run as k_gather_round / k_mlp_round pairs, each
    // kernel with the whole SM to itself (28 gather warps per SM; d2r_march_split.cuh); a FIXED number of them, so nothing is
    // read back.  Whatever is still alive after that, and every small launch as a whole, goes through the persistent
    // warp-specialised k_march_ws, which runs its rays to the end on the device.  Same arithmetic per sample in both: which
    // kernel takes a sample never shows in the frames.
    // D2R_MARCH=ws | split force one kernel for every launch (split: its tail still goes through k_march_ws).
    static const int mode = []() { const char* e = getenv("D2R_MARCH"); return !e || !*e ? 0 : strcmp(e, "ws") == 0 ? 1 : strcmp(e, "split") == 0 ? 2 : 0; }();
    static const int split_rounds = []() { const char* e = getenv("D2R_SPLIT_ROUNDS"); const int v = e ? atoi(e) : 16; return v >= 1 && v <= 256 ? v : 16; }();
    constexpr size_t SPLIT_MIN_RAYS = 1u << 20;
    const bool use_split = !cost_out && (mode == 2 || (mode == 0 && need >= SPLIT_MIN_RAYS));
    P.work_list = nullptr; P.work_count = nullptr; P.t_cur = nullptr; P.resume_acc4 = nullptr; P.resume_acca = nullptr; P.resume_steps = 0;
    P.split_cap = 0; P.feedback = s.fb_dev;
    if (use_split) {
        // how many of the rectangle pixels hit something: from the launches that have finished so far (hits / slots they asked for)
        {
            const uint32_t hits = s.fb_host[0], slots = s.fb_host[1];
            if (slots) s.hit_ratio = std::min(1.0, std::max(0.05, 1.25 * (double)hits / (double)slots));
        }
        const size_t want = std::min(need, (size_t)(s.hit_ratio * (double)need) + 4096);
        if (want > s.cap_split) {
            cudaFree(s.sp_feat); cudaFree(s.sp_aux); cudaFree(s.sp_shb); cudaFree(s.sp_nsb);
            cudaFree(s.sp_live[0]); cudaFree(s.sp_live[1]);
            cudaFree(s.sp_acc4[0]); cudaFree(s.sp_acc4[1]); cudaFree(s.sp_acca[0]); cudaFree(s.sp_acca[1]);
            s.sp_acc4[0] = s.sp_acc4[1] = nullptr; s.sp_acca[0] = s.sp_acca[1] = nullptr;
            s.sp_feat = nullptr; s.sp_aux = nullptr; s.sp_shb = nullptr; s.sp_nsb = nullptr;
            s.sp_live[0] = s.sp_live[1] = nullptr;
            s.cap_split = 0;       // a failed allocation below must not leave a stale capacity behind
            const size_t cap = want + want / 8, blocks = (cap + 127) / 128;
            D2R_CUDA(cudaMalloc(&s.sp_feat, blocks * 2 * SPLIT_TILE_BYTES));
            D2R_CUDA(cudaMalloc(&s.sp_aux, blocks * 2 * 128 * sizeof(float2)));
            D2R_CUDA(cudaMalloc(&s.sp_shb, blocks * 128 * 2 * sizeof(uint4)));
            D2R_CUDA(cudaMalloc(&s.sp_nsb, blocks * 128));
            D2R_CUDA(cudaMalloc(&s.sp_live[0], cap * sizeof(uint32_t)));
            D2R_CUDA(cudaMalloc(&s.sp_live[1], cap * sizeof(uint32_t)));
            for (int i = 0; i < 2; ++i) {
                D2R_CUDA(cudaMalloc(&s.sp_acc4[i], cap * sizeof(float4)));
                D2R_CUDA(cudaMalloc(&s.sp_acca[i], cap * sizeof(float)));
            }
            s.cap_split = blocks * 128;
        }
        if (s.cap_entries > s.cap_t) {      // the ray parameter is kept by entry id
            cudaFree(s.sp_t);
            s.sp_t = nullptr; s.cap_t = 0;
            D2R_CUDA(cudaMalloc(&s.sp_t, s.cap_entries * sizeof(float)));
            s.cap_t = s.cap_entries;
        }
        if (!s.sp_cnt) D2R_CUDA(cudaMalloc(&s.sp_cnt, (256 + 2) * sizeof(uint32_t)));
        D2R_CUDA(cudaMemsetAsync(s.sp_cnt, 0, (256 + 2) * sizeof(uint32_t), stream));
        static bool split_attr[16] = {false};
        if (!split_attr[m->device]) {
            D2R_CUDA(cudaFuncSetAttribute(k_mlp_round, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_TOTAL));
            split_attr[m->device] = true;
        }
        SplitParams Q;
        Q.feat = s.sp_feat; Q.aux = s.sp_aux; Q.shb = s.sp_shb; Q.nsb = s.sp_nsb; Q.t_cur = s.sp_t;
        Q.cap = (uint32_t)std::min<size_t>(s.cap_split, 0xffffffffu);
        P.split_cap = Q.cap;
        for (int r = 0; r < split_rounds; ++r) {
            Q.round = r;
            Q.cnt_in = s.sp_cnt + r; Q.cnt_out = s.sp_cnt + r + 1;
            Q.live_in = s.sp_live[r & 1]; Q.live_out = s.sp_live[(r + 1) & 1];
            Q.acc4_in = s.sp_acc4[r & 1]; Q.acca_in = s.sp_acca[r & 1]; Q.acc4_out = s.sp_acc4[(r + 1) & 1]; Q.acca_out = s.sp_acca[(r + 1) & 1];
            k_gather_round<<<s.n_sm * 7, 128, 0, stream>>>(P, Q);
            k_mlp_round<<<s.n_sm, 128 * TS_GROUPS, TS_TOTAL, stream>>>(P, Q);
            count_launch(2);
        }
        // the survivors: a live ray of round r has taken exactly 2 r samples
        P.work_list = s.sp_live[split_rounds & 1]; P.work_count = s.sp_cnt + split_rounds; P.t_cur = s.sp_t;
        P.resume_acc4 = s.sp_acc4[split_rounds & 1]; P.resume_acca = s.sp_acca[split_rounds & 1];
        P.resume_steps = 2 * split_rounds;
    }
    {
        static bool attr_set[16] = {false};
        if (!attr_set[m->device]) {
            D2R_CUDA(cudaFuncSetAttribute(k_march_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_TOTAL));
            D2R_CUDA(cudaFuncSetAttribute(k_march_ws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_TOTAL));
            attr_set[m->device] = true;
        }
        // D2R_WS_STATS=1 + profiling enabled: the instrumented instantiation (tools/ws_stats.py)
        static const bool want_stats = []() { const char* e = getenv("D2R_WS_STATS"); return e && atoi(e) != 0; }();
        if (want_stats && P.prof) k_march_ws<true><<<s.n_sm, WS_THREADS, WS_TOTAL, stream>>>(P);
        else k_march_ws<false><<<s.n_sm, WS_THREADS, WS_TOTAL, stream>>>(P);
        count_launch();
    }
    if (evp) D2R_CUDA(cudaEventRecord(evp->second, stream));
    P.feedback_slots = (uint32_t)std::min<size_t>(need, 0xffffffffu);
    k_finish<<<s.n_sm * 8, 256, 0, stream>>>(P);
    count_launch();
    D2R_CUDA(cudaGetLastError());
    return D2R_OK;
}

}  // namespace d2r

extern "C" int d2r_profile_enable(int device, int on) {
    D2R_REQUIRE(device >= 0 && device < 16, "d2r_profile_enable: bad device");
    d2r::Prof& pf = d2r::g_prof[device];
    d2r::DeviceGuard dg(device);
    if (on && !pf.counters) D2R_CUDA(cudaMalloc(&pf.counters, 16 * sizeof(unsigned long long)));
    if (pf.counters) D2R_CUDA(cudaMemset(pf.counters, 0, 16 * sizeof(unsigned long long)));
    pf.used = 0;
    pf.on = on != 0;
    return D2R_OK;
}

extern "C" int d2r_profile_read(int device, float* march_ms_total, int* n_launches, unsigned long long* n_samples,
                                unsigned long long* n_tiles) {
    D2R_REQUIRE(device >= 0 && device < 16 && march_ms_total && n_launches && n_samples && n_tiles, "d2r_profile_read: bad argument");
    d2r::Prof& pf = d2r::g_prof[device];
    d2r::DeviceGuard dg(device);
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

extern "C" int d2r_profile_read_stats(int device, unsigned long long* out16) {
    D2R_REQUIRE(device >= 0 && device < 16 && out16, "d2r_profile_read_stats: bad argument");
    d2r::Prof& pf = d2r::g_prof[device];
    d2r::DeviceGuard dg(device);
    D2R_REQUIRE(pf.counters, "d2r_profile_read_stats: profiling was never enabled on this device");
    D2R_CUDA(cudaDeviceSynchronize());
    D2R_CUDA(cudaMemcpy(out16, pf.counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return D2R_OK;
}

extern "C" int d2r_render(const d2r_model* m, const d2r_view* v, const float* cams_ngp_host, int K, const float background_rgba[4],
                          float* rgba_out_dev, float* depth_out_dev, unsigned long long* n_samples_out_dev, void* stream) {
    return d2r::launch_march(m, v, cams_ngp_host, K, background_rgba, rgba_out_dev, depth_out_dev, nullptr, nullptr, nullptr,
                             n_samples_out_dev, (cudaStream_t)stream);
}

extern "C" int d2r_render_ex(const d2r_model* m, const d2r_view* v, const float* cams_ngp_host, int K, const float background_rgba[4],
                             float* rgba_out_dev, float* depth_out_dev, float* cost_out_dev, unsigned long long* n_samples_out_dev, void* stream) {
    return d2r::launch_march(m, v, cams_ngp_host, K, background_rgba, rgba_out_dev, depth_out_dev, nullptr, nullptr, nullptr,
                             n_samples_out_dev, (cudaStream_t)stream, nullptr, nullptr, cost_out_dev);
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
