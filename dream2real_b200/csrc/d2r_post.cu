// CLIP-side pre/post-processing kernels: rot90 + PIL-exact antialiased bicubic resize + normalise
// (transformers CLIPImageProcessor, PIL backend, as called at reference clip_scoring.py:145-147,177)
// and the logit / score arithmetic of clip_scoring.py:180-203.
//
// The resize restates Pillow's ImagingResample for 8-bit images (src/libImaging/Resample.c:
// precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc):
// double-precision bicubic (a = -0.5) weights, normalised, quantised to 22 fractional bits,
// integer accumulate with +0.5 rounding, clamp to u8 -- after the horizontal AND after the vertical
// pass.  Integer arithmetic => bit-exact with PIL.  Pillow is a third-party dependency of the
// reference (requirements.txt: Pillow==9.4.0), not vendored under /root/reference.
#include <math.h>

#include <map>
#include <mutex>
#include <vector>

#include "d2r_common.cuh"

namespace d2r {

constexpr int PRECISION_BITS = 32 - 8 - 2;

static double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for the full-image box
static int precompute_coeffs(int inSize, int outSize, std::vector<int>& bounds, std::vector<int>& kk) {
    const double in0 = 0.0, in1 = (double)inSize;
    double scale = (in1 - in0) / outSize, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)outSize * 2, 0);
    kk.assign((size_t)outSize * ksize, 0);
    std::vector<double> k(ksize);
    for (int xx = 0; xx < outSize; ++xx) {
        const double center = in0 + (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > inSize) xmax = inSize;
        xmax -= xmin;
        int x;
        for (x = 0; x < xmax; ++x) {
            const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        for (; x < ksize; ++x) k[x] = 0;
        for (x = 0; x < ksize; ++x) {
            const double v = k[x];
            kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << PRECISION_BITS)) : (int)(0.5 + v * (1 << PRECISION_BITS));
        }
        bounds[xx * 2 + 0] = xmin;
        bounds[xx * 2 + 1] = xmax;
    }
    return ksize;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PRECISION_BITS;
    return (uint8_t)min(max(v, 0), 255);
}

// horizontal pass over the (optionally rot90'd) image.  The rotated image is rot[i][j] = src[j][W-1-i]
// (np.rot90 k=1 on the image axes); tmp layout is [K][R][Hr][3] with i fastest (coalesced both ways).
__global__ void k_resize_h(const uint8_t* __restrict__ src, int H, int W, int rot90, int Hr, int Wr, int R, int ksize,
                           const int* __restrict__ bounds, const int* __restrict__ kk, uint8_t* __restrict__ tmp) {
    const int k = blockIdx.z;
    const int xx = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Hr) return;
    const uint8_t* img = src + (size_t)k * H * W * 3;
    const int xmin = bounds[xx * 2], xmax = bounds[xx * 2 + 1];
    const int* kr = kk + (size_t)xx * ksize;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < xmax; ++x) {
        const int j = xmin + x;
        const uint8_t* p = rot90 ? img + ((size_t)j * W + (W - 1 - i)) * 3 : img + ((size_t)i * W + j) * 3;
        const int w = kr[x];
        s0 += p[0] * w; s1 += p[1] * w; s2 += p[2] * w;
    }
    uint8_t* o = tmp + (((size_t)k * R + xx) * Hr + i) * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// Fast horizontal pass for the production case (rot90, W % 4 == 0): the rotated image's rows are source
// columns, so four neighbouring rotated rows are 12 contiguous, 4-byte aligned source bytes.  One thread
// owns such a quad: 3 word loads per tap instead of 12 byte loads, 3 word stores instead of 12 byte stores.
// Same integer arithmetic as k_resize_h (bit-exact).
__global__ void k_resize_h_rot4(const uint8_t* __restrict__ src, int H, int W, int R, int ksize, const int* __restrict__ bounds,
                                const int* __restrict__ kk, uint8_t* __restrict__ tmp) {
    const int k = blockIdx.z, xx = blockIdx.y;
    const int quad = blockIdx.x * blockDim.x + threadIdx.x;      // source columns 4*quad .. 4*quad+3
    if (quad * 4 >= W) return;
    const int c0 = quad * 4;
    const uint32_t* img = reinterpret_cast<const uint32_t*>(src + (size_t)k * H * W * 3);
    const int xmin = bounds[xx * 2], xmax = bounds[xx * 2 + 1];
    const int* kr = kk + (size_t)xx * ksize;
    int acc[12];
#pragma unroll
    for (int b = 0; b < 12; ++b) acc[b] = 1 << (PRECISION_BITS - 1);
    for (int x = 0; x < xmax; ++x) {
        const int j = xmin + x;                                   // source row
        const uint32_t* p = img + ((size_t)j * W + c0) * 3 / 4;
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        const int w = kr[x];
        acc[0] += (int)(w0 & 0xff) * w; acc[1] += (int)((w0 >> 8) & 0xff) * w; acc[2] += (int)((w0 >> 16) & 0xff) * w;
        acc[3] += (int)(w0 >> 24) * w; acc[4] += (int)(w1 & 0xff) * w; acc[5] += (int)((w1 >> 8) & 0xff) * w;
        acc[6] += (int)((w1 >> 16) & 0xff) * w; acc[7] += (int)(w1 >> 24) * w; acc[8] += (int)(w2 & 0xff) * w;
        acc[9] += (int)((w2 >> 8) & 0xff) * w; acc[10] += (int)((w2 >> 16) & 0xff) * w; acc[11] += (int)(w2 >> 24) * w;
    }
    // source column c -> rotated row i = W-1-c; the quad is rotated rows W-4-c0 .. W-1-c0, i.e. pixels in reverse order
    uint8_t o[12];
#pragma unroll
    for (int px = 0; px < 4; ++px) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) o[(3 - px) * 3 + ch] = clip8(acc[px * 3 + ch]);
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(tmp + (((size_t)k * R + xx) * W + (W - 4 - c0)) * 3);
    dst[0] = o[0] | (o[1] << 8) | (o[2] << 16) | ((uint32_t)o[3] << 24);
    dst[1] = o[4] | (o[5] << 8) | (o[6] << 16) | ((uint32_t)o[7] << 24);
    dst[2] = o[8] | (o[9] << 8) | (o[10] << 16) | ((uint32_t)o[11] << 24);
}

// vertical pass + /255 + normalise, written patch-major fp16 (K padded to Kp columns) and optionally
// as float32 pixel_values [K,3,R,R].
__global__ void k_resize_v_norm(const uint8_t* __restrict__ tmp, int Hr, int R, int ksize, const int* __restrict__ bounds,
                                const int* __restrict__ kk, int P, int Kp, float m0, float m1, float m2, float i0, float i1,
                                float i2, __half* __restrict__ patches, float* __restrict__ pixels) {
    const int k = blockIdx.z;
    const int yy = blockIdx.y;
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= R) return;
    const int ymin = bounds[yy * 2], ymax = bounds[yy * 2 + 1];
    const int* kr = kk + (size_t)yy * ksize;
    // this thread's taps are ymax*3 consecutive bytes of tmp: fetch them as aligned 32-bit words
    const size_t start = (((size_t)k * R + xx) * Hr + ymin) * 3;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(tmp + (start & ~(size_t)3));
    int b = (int)(start & 3);
    uint32_t wv = __ldg(wp);
    int widx = 0;
    auto next_byte = [&]() {
        const int wi = b >> 2;
        if (wi != widx) { wv = __ldg(wp + wi); widx = wi; }
        const int v = (int)((wv >> (8 * (b & 3))) & 0xff);
        ++b;
        return v;
    };
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < ymax; ++y) {
        const int w = kr[y];
        s0 += next_byte() * w; s1 += next_byte() * w; s2 += next_byte() * w;
    }
    const uint8_t u[3] = {clip8(s0), clip8(s1), clip8(s2)};
    const float mean[3] = {m0, m1, m2}, stdv[3] = {i0, i1, i2};
    const int np_side = R / P;
    const size_t prow = (size_t)k * np_side * np_side + (size_t)(yy / P) * np_side + xx / P;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // transformers 4.27 image_transforms: rescale = (u8 * (1/255) in float64).astype(float32);
        // normalize = (image - mean) / std in float32
        const float v = (float)((double)u[c] * (1.0 / 255.0));
        const float nv = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
        if (patches) patches[prow * Kp + (size_t)c * P * P + (yy % P) * P + (xx % P)] = __float2half_rn(nv);
        if (pixels) pixels[(((size_t)k * 3 + c) * R + yy) * R + xx] = nv;
    }
}

__global__ void k_zero_pad_cols(__half* patches, size_t rows, int Kp, int K0) {
    const size_t r = (size_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (r >= rows) return;
    for (int c = K0 + threadIdx.x; c < Kp; c += blockDim.x) patches[r * Kp + c] = __float2half_rn(0.f);
}

// logits_per_image = exp(logit_scale) * img @ txt^T; score = mean(goal) / mean(norm)  (clip_scoring.py:180-203)
__global__ void k_score(const float* __restrict__ img, const float* __restrict__ txt, int K, int Cn, int D, float scale, int n_goal,
                        float* __restrict__ scores, float* __restrict__ logits) {
    const int k = blockIdx.x;
    extern __shared__ float s_logit[];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
    for (int c = warp; c < Cn; c += nw) {
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) acc += img[(size_t)k * D + d] * txt[(size_t)c * D + d];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) { s_logit[c] = acc * scale; if (logits) logits[(size_t)k * Cn + c] = acc * scale; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float g = 0.f, n = 0.f;
        for (int c = 0; c < n_goal; ++c) g += s_logit[c];
        g /= (float)n_goal;
        if (Cn > n_goal) {
            for (int c = n_goal; c < Cn; ++c) n += s_logit[c];
            n /= (float)(Cn - n_goal);
            scores[k] = g / n;
        } else {
            scores[k] = g;
        }
    }
}

// ---- delta preprocessing ---------------------------------------------------------------------------------------
// Every candidate frame equals the composited background outside the candidate's screen rectangle, so its resized
// image equals the resized background except where a filter window touches the rectangle.  The background goes
// through the two passes once; per candidate only the affected outputs are recomputed -- with the same integer
// arithmetic on the same source bytes, hence bit-identical to resizing the whole frame -- and written over a copy
// of the background's patch rows.
struct DeltaRange { int i0, i1, xx0, xx1, yy0, yy1; };   // inclusive; xx1 < xx0: nothing to do

__global__ void k_delta_ranges(int K, int W, int rot90, int R, const int4* __restrict__ rects, const int* __restrict__ hb,
                               const int* __restrict__ vb, DeltaRange* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int4 r = rects[k];                       // x0, y0, x1, y1 (inclusive) in frame coordinates
    DeltaRange d;
    d.i0 = 0; d.i1 = -1; d.xx0 = 0; d.xx1 = -1; d.yy0 = 0; d.yy1 = -1;
    if (r.z >= r.x && r.w >= r.y) {
        // rotated image rot[i][j] = src[j][W-1-i]: rows <- frame columns (reversed), columns <- frame rows
        const int i0 = rot90 ? W - 1 - r.z : r.y, i1 = rot90 ? W - 1 - r.x : r.w;
        const int j0 = rot90 ? r.y : r.x, j1 = rot90 ? r.w : r.z;
        int a = R, b = -1;
        for (int xx = 0; xx < R; ++xx) {
            const int lo = hb[2 * xx], hi = lo + hb[2 * xx + 1] - 1;
            if (hi >= j0 && lo <= j1) { a = min(a, xx); b = max(b, xx); }
        }
        int c = R, e = -1;
        for (int yy = 0; yy < R; ++yy) {
            const int lo = vb[2 * yy], hi = lo + vb[2 * yy + 1] - 1;
            if (hi >= i0 && lo <= i1) { c = min(c, yy); e = max(e, yy); }
        }
        d.i0 = i0; d.i1 = i1; d.xx0 = a; d.xx1 = b; d.yy0 = c; d.yy1 = e;
    }
    out[k] = d;
}

__global__ void k_copy_patches(const uint4* __restrict__ bg, size_t n16, uint4* __restrict__ dst) {
    uint4* o = dst + (size_t)blockIdx.y * n16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) o[i] = __ldg(bg + i);
}

__global__ void k_resize_h_delta(const uint8_t* __restrict__ src, int H, int W, int rot90, int Hr, int R, int ksize,
                                 const int* __restrict__ bounds, const int* __restrict__ kk, const DeltaRange* __restrict__ ranges,
                                 uint8_t* __restrict__ tmp) {
    const int k = blockIdx.z;
    const DeltaRange d = ranges[k];
    const uint8_t* img = src + (size_t)k * H * W * 3;
    for (int xx = d.xx0 + blockIdx.y; xx <= d.xx1; xx += gridDim.y) {
        const int xmin = bounds[xx * 2], xmax = bounds[xx * 2 + 1];
        const int* kr = kk + (size_t)xx * ksize;
        for (int i = d.i0 + blockIdx.x * blockDim.x + threadIdx.x; i <= d.i1; i += gridDim.x * blockDim.x) {
            int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
            for (int x = 0; x < xmax; ++x) {
                const int j = xmin + x;
                const uint8_t* p = rot90 ? img + ((size_t)j * W + (W - 1 - i)) * 3 : img + ((size_t)i * W + j) * 3;
                const int w = kr[x];
                s0 += p[0] * w; s1 += p[1] * w; s2 += p[2] * w;
            }
            uint8_t* o = tmp + (((size_t)k * R + xx) * Hr + i) * 3;
            o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
        }
    }
}

__global__ void k_resize_v_delta(const uint8_t* __restrict__ tmp, const uint8_t* __restrict__ bg_tmp, int Hr, int R, int ksize,
                                 const int* __restrict__ bounds, const int* __restrict__ kk, const DeltaRange* __restrict__ ranges, int P,
                                 int Kp, float m0, float m1, float m2, float i0_, float i1_, float i2_, __half* __restrict__ patches) {
    const int k = blockIdx.z;
    const DeltaRange d = ranges[k];
    const float mean[3] = {m0, m1, m2}, stdv[3] = {i0_, i1_, i2_};
    const int np_side = R / P;
    for (int yy = d.yy0 + blockIdx.y; yy <= d.yy1; yy += gridDim.y) {
        const int ymin = bounds[yy * 2], ymax = bounds[yy * 2 + 1];
        const int* kr = kk + (size_t)yy * ksize;
        for (int xx = d.xx0 + blockIdx.x * blockDim.x + threadIdx.x; xx <= d.xx1; xx += gridDim.x * blockDim.x) {
            const uint8_t* own = tmp + (((size_t)k * R + xx) * Hr) * 3;
            const uint8_t* bgr = bg_tmp + ((size_t)xx * Hr) * 3;
            int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
            for (int y = 0; y < ymax; ++y) {
                const int i = ymin + y;
                const uint8_t* p = (i >= d.i0 && i <= d.i1 ? own : bgr) + (size_t)i * 3;
                const int w = kr[y];
                s0 += p[0] * w; s1 += p[1] * w; s2 += p[2] * w;
            }
            const uint8_t u[3] = {clip8(s0), clip8(s1), clip8(s2)};
            const size_t prow = (size_t)k * np_side * np_side + (size_t)(yy / P) * np_side + xx / P;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = (float)((double)u[c] * (1.0 / 255.0));
                const float nv = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
                patches[prow * Kp + (size_t)c * P * P + (yy % P) * P + (xx % P)] = __float2half_rn(nv);
            }
        }
    }
}

struct ResizePlan {
    int in_size = 0, out_size = 0, ksize = 0;
    int* bounds_dev = nullptr;
    int* kk_dev = nullptr;
};
// resize coefficient tables: read-only once built, shared per device (guarded); scratch: per (device, stream), so two host threads
// driving two streams of one GPU never share it (kernels of one stream run in order, so one set per stream is enough)
static std::mutex g_post_mutex;
static std::vector<ResizePlan> g_plans[16];
struct DeltaScratch { uint8_t* bg_tmp = nullptr; size_t bg_tmp_cap = 0; __half* bg_patches = nullptr; size_t bg_patches_cap = 0;
                      DeltaRange* ranges = nullptr; int ranges_cap = 0; };
struct PostCtx { uint8_t* tmp = nullptr; size_t tmp_cap = 0; DeltaScratch delta; };
static std::map<std::pair<int, cudaStream_t>, PostCtx*> g_post_ctx;

static PostCtx* get_post_ctx(int device, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_post_mutex);
    PostCtx*& c = g_post_ctx[std::make_pair(device, stream)];
    if (!c) c = new PostCtx();
    return c;
}

static int get_plan(int device, int in_size, int out_size, ResizePlan* out) {
    std::lock_guard<std::mutex> lock(g_post_mutex);
    for (const ResizePlan& p : g_plans[device])
        if (p.in_size == in_size && p.out_size == out_size) { *out = p; return D2R_OK; }
    ResizePlan p;
    std::vector<int> bounds, kk;
    p.ksize = precompute_coeffs(in_size, out_size, bounds, kk);
    p.in_size = in_size; p.out_size = out_size;
    D2R_CUDA(cudaMalloc(&p.bounds_dev, bounds.size() * sizeof(int)));
    D2R_CUDA(cudaMalloc(&p.kk_dev, kk.size() * sizeof(int)));
    D2R_CUDA(cudaMemcpy(p.bounds_dev, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
    D2R_CUDA(cudaMemcpy(p.kk_dev, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice));
    g_plans[device].push_back(p);
    *out = p;
    return D2R_OK;
}

}  // namespace d2r

using namespace d2r;

static int run_full_resize(const uint8_t* rgb_u8_dev, int K, int H, int W, int rot90, int R, int P, const float mean[3], const float std_[3],
                           const ResizePlan* ph, const ResizePlan* pv, uint8_t* tmp, __half* patches, float* pixels, cudaStream_t stream);

extern "C" int d2r_clip_preprocess(const uint8_t* rgb_u8_dev, int K, int H, int W, int rot90, int R, int P, const float mean[3],
                                   const float std_[3], void* patches_out_dev, float* pixels_f32_out_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D2R_REQUIRE(rgb_u8_dev && mean && std_ && (patches_out_dev || pixels_f32_out_dev), "d2r_clip_preprocess: null argument");
    D2R_REQUIRE(K > 0 && H > 0 && W > 0 && R > 0 && P > 0 && R % P == 0, "d2r_clip_preprocess: bad sizes");
    D2R_REQUIRE(H == W, "d2r_clip_preprocess: only square renders are on this path (resize shortest edge == both edges)");
    D2R_REQUIRE(K <= 65535, "d2r_clip_preprocess: K must be <= 65535 per call");
    int device;
    D2R_CUDA(cudaGetDevice(&device));
    D2R_REQUIRE(device < 16, "d2r_clip_preprocess: device index too large");
    const int Hr = rot90 ? W : H, Wr = rot90 ? H : W;
    ResizePlan ph_, pv_;
    int rc = get_plan(device, Wr, R, &ph_);
    if (rc) return rc;
    rc = get_plan(device, Hr, R, &pv_);
    if (rc) return rc;
    const ResizePlan *ph = &ph_, *pv = &pv_;
    const size_t tmp_bytes = (size_t)K * R * Hr * 3 + 16;   // + slack: the vertical pass reads whole aligned words
    PostCtx& ctx = *get_post_ctx(device, stream);
    if (tmp_bytes > ctx.tmp_cap) {
        if (ctx.tmp) D2R_CUDA(cudaFree(ctx.tmp));
        ctx.tmp = nullptr; ctx.tmp_cap = 0;
        D2R_CUDA(cudaMalloc(&ctx.tmp, tmp_bytes));
        ctx.tmp_cap = tmp_bytes;
    }
    rc = run_full_resize(rgb_u8_dev, K, H, W, rot90, R, P, mean, std_, ph, pv, ctx.tmp, (__half*)patches_out_dev, pixels_f32_out_dev, stream);
    if (rc) return rc;
    D2R_CUDA(cudaGetLastError());
    return D2R_OK;
}

static int run_full_resize(const uint8_t* rgb_u8_dev, int K, int H, int W, int rot90, int R, int P, const float mean[3], const float std_[3],
                           const ResizePlan* ph, const ResizePlan* pv, uint8_t* tmp, __half* patches, float* pixels, cudaStream_t stream) {
    const int Hr = rot90 ? W : H, Wr = rot90 ? H : W;
    if (rot90 && W % 4 == 0 && ((uintptr_t)rgb_u8_dev % 4) == 0) {
        dim3 grid((W / 4 + 127) / 128, R, K);
        k_resize_h_rot4<<<grid, 128, 0, stream>>>(rgb_u8_dev, H, W, R, ph->ksize, ph->bounds_dev, ph->kk_dev, tmp);
    } else {
        dim3 grid((Hr + 127) / 128, R, K);
        k_resize_h<<<grid, 128, 0, stream>>>(rgb_u8_dev, H, W, rot90, Hr, Wr, R, ph->ksize, ph->bounds_dev, ph->kk_dev, tmp);
    }
    const int K0 = 3 * P * P, Kp = (K0 + 63) / 64 * 64;
    dim3 grid((R + 127) / 128, R, K);
    k_resize_v_norm<<<grid, 128, 0, stream>>>(tmp, Hr, R, pv->ksize, pv->bounds_dev, pv->kk_dev, P, Kp, mean[0], mean[1], mean[2], std_[0],
                                              std_[1], std_[2], patches, pixels);
    count_launch(2);
    if (patches && Kp != K0) {
        const size_t rows = (size_t)K * (R / P) * (R / P);
        dim3 block(32, 8);
        k_zero_pad_cols<<<(unsigned)((rows + 7) / 8), block, 0, stream>>>(patches, rows, Kp, K0);
        count_launch();
    }
    return D2R_OK;
}

extern "C" int d2r_clip_preprocess_delta(const uint8_t* rgb_u8_dev, int K, int H, int W, int rot90, int R, int P, const float mean[3],
                                         const float std_[3], const uint8_t* bg_u8_dev, const int* rects_dev, void* patches_out_dev,
                                         void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D2R_REQUIRE(rgb_u8_dev && mean && std_ && bg_u8_dev && rects_dev && patches_out_dev, "d2r_clip_preprocess_delta: null argument");
    D2R_REQUIRE(K > 0 && H > 0 && W > 0 && R > 0 && P > 0 && R % P == 0, "d2r_clip_preprocess_delta: bad sizes");
    D2R_REQUIRE(H == W, "d2r_clip_preprocess_delta: only square renders are on this path");
    D2R_REQUIRE(K <= 65535, "d2r_clip_preprocess_delta: K must be <= 65535 per call");
    int device;
    D2R_CUDA(cudaGetDevice(&device));
    D2R_REQUIRE(device < 16, "d2r_clip_preprocess_delta: device index too large");
    const int Hr = rot90 ? W : H, Wr = rot90 ? H : W;
    ResizePlan ph, pv;
    int rc = get_plan(device, Wr, R, &ph);
    if (rc) return rc;
    rc = get_plan(device, Hr, R, &pv);
    if (rc) return rc;
    const size_t tmp_bytes = (size_t)K * R * Hr * 3 + 16;
    PostCtx& ctx = *get_post_ctx(device, stream);
    if (tmp_bytes > ctx.tmp_cap) {
        if (ctx.tmp) D2R_CUDA(cudaFree(ctx.tmp));
        ctx.tmp = nullptr; ctx.tmp_cap = 0;
        D2R_CUDA(cudaMalloc(&ctx.tmp, tmp_bytes));
        ctx.tmp_cap = tmp_bytes;
    }
    DeltaScratch& ds = ctx.delta;
    const int K0 = 3 * P * P, Kp = (K0 + 63) / 64 * 64, np = (R / P) * (R / P);
    const size_t bg_tmp_bytes = (size_t)R * Hr * 3 + 16, bg_patch_elems = (size_t)np * Kp;
    if (bg_tmp_bytes > ds.bg_tmp_cap) {
        if (ds.bg_tmp) D2R_CUDA(cudaFree(ds.bg_tmp));
        D2R_CUDA(cudaMalloc(&ds.bg_tmp, bg_tmp_bytes));
        ds.bg_tmp_cap = bg_tmp_bytes;
    }
    if (bg_patch_elems > ds.bg_patches_cap) {
        if (ds.bg_patches) D2R_CUDA(cudaFree(ds.bg_patches));
        D2R_CUDA(cudaMalloc(&ds.bg_patches, bg_patch_elems * sizeof(__half)));
        ds.bg_patches_cap = bg_patch_elems;
    }
    if (K > ds.ranges_cap) {
        if (ds.ranges) D2R_CUDA(cudaFree(ds.ranges));
        D2R_CUDA(cudaMalloc(&ds.ranges, (size_t)K * sizeof(DeltaRange)));
        ds.ranges_cap = K;
    }
    // 1. the background frame through both passes (one image)
    rc = run_full_resize(bg_u8_dev, 1, H, W, rot90, R, P, mean, std_, &ph, &pv, ds.bg_tmp, ds.bg_patches, nullptr, stream);
    if (rc) return rc;
    // 2. per candidate: affected output ranges, background patch rows, then the affected outputs only
    k_delta_ranges<<<(K + 127) / 128, 128, 0, stream>>>(K, W, rot90, R, (const int4*)rects_dev, ph.bounds_dev, pv.bounds_dev, ds.ranges);
    const size_t n16 = bg_patch_elems * sizeof(__half) / 16;       // Kp is a multiple of 64 halves
    k_copy_patches<<<dim3(8, K), 256, 0, stream>>>((const uint4*)ds.bg_patches, n16, (uint4*)patches_out_dev);
    k_resize_h_delta<<<dim3(2, 64, K), 128, 0, stream>>>(rgb_u8_dev, H, W, rot90, Hr, R, ph.ksize, ph.bounds_dev, ph.kk_dev, ds.ranges,
                                                         ctx.tmp);
    k_resize_v_delta<<<dim3(1, 64, K), 64, 0, stream>>>(ctx.tmp, ds.bg_tmp, Hr, R, pv.ksize, pv.bounds_dev, pv.kk_dev, ds.ranges, P, Kp,
                                                        mean[0], mean[1], mean[2], std_[0], std_[1], std_[2], (__half*)patches_out_dev);
    count_launch(4);
    D2R_CUDA(cudaGetLastError());
    return D2R_OK;
}

extern "C" int d2r_score(const float* img_embeds_dev, const float* txt_embeds_dev, int K, int C, int D, float logit_scale_exp, int n_goal,
                         float* scores_out_dev, float* logits_out_dev, void* stream) {
    D2R_REQUIRE(img_embeds_dev && txt_embeds_dev && scores_out_dev, "d2r_score: null argument");
    D2R_REQUIRE(K > 0 && C > 0 && D > 0 && n_goal > 0 && n_goal <= C, "d2r_score: bad sizes");
    k_score<<<K, 128, C * sizeof(float), (cudaStream_t)stream>>>(img_embeds_dev, txt_embeds_dev, K, C, D, logit_scale_exp, n_goal,
                                                                scores_out_dev, logits_out_dev);
    count_launch();
    D2R_CUDA(cudaGetLastError());
    return D2R_OK;
}
