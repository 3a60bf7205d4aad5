// CLIP vision tower forward (ViT-B/32-224, ViT-L/14-336, ... config-driven) for the image side of
// reference clip_scoring.py:180-181 (transformers CLIPModel.forward: vision_model ->
// visual_projection -> L2 normalise).  Third-party algorithm: transformers==4.27.3
// models/clip/modeling_clip.py (CLIPVisionEmbeddings, CLIPEncoderLayer, CLIPAttention, CLIPMLP with
// quick_gelu, pre_layrnorm / post_layernorm); not vendored under /root/reference.
//
// Every dense contraction runs on the tcgen05 GEMM (d2r_gemm.cu) with fp16 operands and fp32
// accumulation; the residual stream, LayerNorm statistics, softmax and the final normalisation
// are fp32.  q is pre-scaled by head_dim^-0.5 = 1/8 (exact in fp16) at load time.
#include <math.h>

#include <vector>

#include "d2r_common.cuh"
#include "d2r_gemm_api.h"

struct d2r_clip {
    int device;
    d2r_clip_cfg cfg;
    int T, NP, Kp;            // tokens, patches, padded patch length
    // parameters
    __half *w_patch, *w_proj;
    float *cls, *pos, *pre_w, *pre_b, *post_w, *post_b;
    struct Layer {
        float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *b_qkv, *b_o, *b_1, *b_2;
        __half *w_qkv, *w_o, *w_1, *w_2;
    };
    std::vector<Layer> layers;
    std::vector<void*> allocs;
    // workspaces for max_batch images
    float *pe, *x, *emb;
    __half *h, *qkv, *o, *m, *pooled;
};

namespace d2r {

// ---- small kernels ---------------------------------------------------------------------------------
// one warp per row LayerNorm (biased variance, eps inside the sqrt: torch.nn.LayerNorm)
template <typename OUT>
__device__ __forceinline__ void warp_layernorm_row(const float* __restrict__ in, int d, const float* __restrict__ w,
                                                   const float* __restrict__ b, float eps, OUT* __restrict__ out, int lane) {
    float v[32];
    const int n = d / 32;   // d <= 1024, multiple of 32
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n) { v[i] = in[i * 32 + lane]; s += v[i]; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n) { const float c = v[i] - mean; q += c * c; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)d + eps);
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n) {
            const int c = i * 32 + lane;
            const float y = (v[i] - mean) * rstd * w[c] + b[c];
            if constexpr (sizeof(OUT) == 2) out[c] = __float2half_rn(y);
            else out[c] = y;
        }
}

// embeddings (class token | patch embeddings) + position embeddings, then pre_layrnorm -> residual stream x
__global__ void k_embed_preln(const float* __restrict__ pe, const float* __restrict__ cls, const float* __restrict__ pos,
                              const float* __restrict__ w, const float* __restrict__ b, float eps, int B, int T, int d,
                              float* __restrict__ x) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= B * T) return;
    const int img = row / T, t = row % T;
    float* xr = x + (size_t)row * d;
    const float* src = t == 0 ? cls : pe + ((size_t)img * (T - 1) + (t - 1)) * d;
    for (int c = lane; c < d; c += 32) xr[c] = src[c] + pos[(size_t)t * d + c];
    __syncwarp();
    warp_layernorm_row<float>(xr, d, w, b, eps, xr, lane);
}

__global__ void k_layernorm_f16(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float eps,
                                int rows, int d, size_t in_row_stride, __half* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (row >= rows) return;
    warp_layernorm_row<__half>(x + (size_t)row * in_row_stride, d, w, b, eps, out + (size_t)row * d, threadIdx.x % 32);
}

// softmax(q k^T) v for one (image, head); q already carries the 1/sqrt(head_dim) scale.
// K/V of the head live in shared memory (fp16), scores / probabilities in fp32 registers.
template <int NJ>   // ceil(T / 32)
__global__ void __launch_bounds__(128) k_attention(const __half* __restrict__ qkv, int T, int d, __half* __restrict__ out) {
    constexpr int HD = 64, KS = 66;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* Ks = reinterpret_cast<__half*>(smem_raw);
    __half* Vs = Ks + (size_t)T * KS;
    const int img = blockIdx.x, head = blockIdx.y;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const __half* base = qkv + (size_t)img * T * 3 * d + head * HD;
    for (int i = threadIdx.x; i < T * (HD / 2); i += blockDim.x) {
        const int t = i / (HD / 2), c2 = i % (HD / 2);
        const __half2 kv = *reinterpret_cast<const __half2*>(base + (size_t)t * 3 * d + d + c2 * 2);
        const __half2 vv = *reinterpret_cast<const __half2*>(base + (size_t)t * 3 * d + 2 * d + c2 * 2);
        *reinterpret_cast<__half2*>(Ks + (size_t)t * KS + c2 * 2) = kv;
        *reinterpret_cast<__half2*>(Vs + (size_t)t * HD + c2 * 2) = vv;
    }
    __syncthreads();
    for (int r = warp; r < T; r += 4) {
        float2 q[HD / 2];
        const __half2* qp = reinterpret_cast<const __half2*>(base + (size_t)r * 3 * d);
#pragma unroll
        for (int c = 0; c < HD / 2; ++c) q[c] = __half22float2(qp[c]);
        float s[NJ];
        float mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const int j = jj * 32 + lane;
            float acc = -INFINITY;
            if (j < T) {
                acc = 0.f;
                const __half2* kp = reinterpret_cast<const __half2*>(Ks + (size_t)j * KS);
#pragma unroll
                for (int c = 0; c < HD / 2; ++c) {
                    const float2 kk = __half22float2(kp[c]);
                    acc = fmaf(q[c].x, kk.x, acc);
                    acc = fmaf(q[c].y, kk.y, acc);
                }
            }
            s[jj] = acc;
            mx = fmaxf(mx, acc);
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const float p = (jj * 32 + lane < T) ? expf(s[jj] - mx) : 0.f;
            s[jj] = p;
            sum += p;
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const int jmax = min(32, T - jj * 32);
            for (int src = 0; src < jmax; ++src) {
                const float p = __shfl_sync(0xffffffffu, s[jj], src);
                const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(Vs + (size_t)(jj * 32 + src) * HD + lane * 2));
                a0 = fmaf(p, vv.x, a0);
                a1 = fmaf(p, vv.y, a1);
            }
        }
        *reinterpret_cast<__half2*>(out + ((size_t)img * T + r) * d + head * HD + lane * 2) = __floats2half2_rn(a0 * inv, a1 * inv);
    }
}

__global__ void k_l2norm(const float* __restrict__ in, int rows, int D, float* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
    if (row >= rows) return;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) { const float v = in[(size_t)row * D + c]; s += v * v; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float n = sqrtf(s);
    for (int c = lane; c < D; c += 32) out[(size_t)row * D + c] = in[(size_t)row * D + c] / n;
}

template <typename T>
static int dev_alloc(d2r_clip* c, T** p, size_t n) {
    D2R_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
    c->allocs.push_back((void*)*p);
    return D2R_OK;
}
static int upload_f32(d2r_clip* c, float** dst, const float* src, size_t n, float scale = 1.f) {
    int rc = dev_alloc(c, dst, n);
    if (rc) return rc;
    if (scale == 1.f) {
        D2R_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    } else {
        std::vector<float> t(src, src + n);
        for (auto& v : t) v *= scale;
        D2R_CUDA(cudaMemcpy(*dst, t.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    }
    return D2R_OK;
}
// fp32 [rows, cols] host -> fp16 [rows, cols_padded] device at `dst` (row offset already applied)
static int upload_f16_rows(__half* dst, const float* src, size_t rows, size_t cols, size_t cols_padded, float scale = 1.f) {
    std::vector<__half> t(rows * cols_padded, __float2half_rn(0.f));
    for (size_t r = 0; r < rows; ++r)
        for (size_t cc = 0; cc < cols; ++cc) t[r * cols_padded + cc] = __float2half_rn(src[r * cols + cc] * scale);
    D2R_CUDA(cudaMemcpy(dst, t.data(), t.size() * sizeof(__half), cudaMemcpyHostToDevice));
    return D2R_OK;
}

}  // namespace d2r

using namespace d2r;

extern "C" void d2r_clip_free(d2r_clip* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (void* p : c->allocs) cudaFree(p);
    delete c;
}

#define CLIP_TRY(expr) do { int _rc = (expr); if (_rc) { d2r_clip_free(c); return _rc; } } while (0)

extern "C" int d2r_clip_load(const d2r_clip_cfg* cfg, const float* const* W, int n_weights, int device, d2r_clip** out) {
    D2R_REQUIRE(cfg && W && out, "d2r_clip_load: null argument");
    const int d = cfg->hidden, P = cfg->patch_size, L = cfg->layers;
    D2R_REQUIRE(cfg->image_size % P == 0, "d2r_clip_load: image_size must be a multiple of patch_size");
    D2R_REQUIRE(d % 64 == 0 && d <= 1024 && cfg->mlp % 64 == 0 && cfg->proj % 64 == 0, "d2r_clip_load: hidden/mlp/proj must be multiples of 64, hidden <= 1024");
    D2R_REQUIRE(cfg->heads > 0 && d / cfg->heads == 64 && d % cfg->heads == 0, "d2r_clip_load: head_dim must be 64");
    D2R_REQUIRE(cfg->max_batch > 0, "d2r_clip_load: max_batch must be positive");
    D2R_REQUIRE(n_weights == 5 + 16 * L + 3, "d2r_clip_load: expected 5 + 16*layers + 3 weight tensors");
    for (int i = 0; i < n_weights; ++i) D2R_REQUIRE(W[i] != nullptr, "d2r_clip_load: null weight pointer");
    D2R_CUDA(cudaSetDevice(device));
    d2r_clip* c = new d2r_clip();
    c->device = device;
    c->cfg = *cfg;
    const int side = cfg->image_size / P;
    c->NP = side * side;
    c->T = c->NP + 1;
    D2R_REQUIRE(c->T <= 32 * 19, "d2r_clip_load: at most 608 tokens are supported");
    const int K0 = 3 * P * P;
    c->Kp = (K0 + 63) / 64 * 64;
    const int mlp = cfg->mlp;

    CLIP_TRY(dev_alloc(c, &c->w_patch, (size_t)d * c->Kp));
    CLIP_TRY(upload_f16_rows(c->w_patch, W[0], d, K0, c->Kp));
    CLIP_TRY(upload_f32(c, &c->cls, W[1], d));
    CLIP_TRY(upload_f32(c, &c->pos, W[2], (size_t)c->T * d));
    CLIP_TRY(upload_f32(c, &c->pre_w, W[3], d));
    CLIP_TRY(upload_f32(c, &c->pre_b, W[4], d));
    c->layers.resize(L);
    const float qscale = 0.125f;   // head_dim^-0.5, modeling_clip.py CLIPAttention.scale
    for (int l = 0; l < L; ++l) {
        const float* const* w = W + 5 + 16 * l;
        d2r_clip::Layer& y = c->layers[l];
        CLIP_TRY(upload_f32(c, &y.ln1_w, w[0], d));
        CLIP_TRY(upload_f32(c, &y.ln1_b, w[1], d));
        CLIP_TRY(dev_alloc(c, &y.w_qkv, (size_t)3 * d * d));
        CLIP_TRY(upload_f16_rows(y.w_qkv, w[2], d, d, d, qscale));
        CLIP_TRY(upload_f16_rows(y.w_qkv + (size_t)d * d, w[4], d, d, d));
        CLIP_TRY(upload_f16_rows(y.w_qkv + (size_t)2 * d * d, w[6], d, d, d));
        CLIP_TRY(dev_alloc(c, &y.b_qkv, (size_t)3 * d));
        {
            std::vector<float> b(3 * d);
            for (int i = 0; i < d; ++i) { b[i] = w[3][i] * qscale; b[d + i] = w[5][i]; b[2 * d + i] = w[7][i]; }
            D2R_CUDA(cudaMemcpy(y.b_qkv, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
        CLIP_TRY(dev_alloc(c, &y.w_o, (size_t)d * d));
        CLIP_TRY(upload_f16_rows(y.w_o, w[8], d, d, d));
        CLIP_TRY(upload_f32(c, &y.b_o, w[9], d));
        CLIP_TRY(upload_f32(c, &y.ln2_w, w[10], d));
        CLIP_TRY(upload_f32(c, &y.ln2_b, w[11], d));
        CLIP_TRY(dev_alloc(c, &y.w_1, (size_t)mlp * d));
        CLIP_TRY(upload_f16_rows(y.w_1, w[12], mlp, d, d));
        CLIP_TRY(upload_f32(c, &y.b_1, w[13], mlp));
        CLIP_TRY(dev_alloc(c, &y.w_2, (size_t)d * mlp));
        CLIP_TRY(upload_f16_rows(y.w_2, w[14], d, mlp, mlp));
        CLIP_TRY(upload_f32(c, &y.b_2, w[15], d));
    }
    const float* const* wt = W + 5 + 16 * L;
    CLIP_TRY(upload_f32(c, &c->post_w, wt[0], d));
    CLIP_TRY(upload_f32(c, &c->post_b, wt[1], d));
    CLIP_TRY(dev_alloc(c, &c->w_proj, (size_t)cfg->proj * d));
    CLIP_TRY(upload_f16_rows(c->w_proj, wt[2], cfg->proj, d, d));

    const size_t B = cfg->max_batch, M = B * c->T;
    CLIP_TRY(dev_alloc(c, &c->pe, B * c->NP * d));
    CLIP_TRY(dev_alloc(c, &c->x, M * d));
    CLIP_TRY(dev_alloc(c, &c->h, M * d));
    CLIP_TRY(dev_alloc(c, &c->qkv, M * 3 * d));
    CLIP_TRY(dev_alloc(c, &c->o, M * d));
    CLIP_TRY(dev_alloc(c, &c->m, M * mlp));
    CLIP_TRY(dev_alloc(c, &c->pooled, B * d));
    CLIP_TRY(dev_alloc(c, &c->emb, B * cfg->proj));
    *out = c;
    return D2R_OK;
}

template <int NJ>
static int launch_attention(const d2r_clip* c, int B, cudaStream_t stream) {
    const int T = c->T, d = c->cfg.hidden;
    const size_t smem = (size_t)T * 66 * 2 + (size_t)T * 64 * 2;
    static size_t granted[16] = {0};   // the attribute is a limit: only ever raise it
    if (smem > granted[c->device & 15]) {
        D2R_CUDA(cudaFuncSetAttribute(k_attention<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        granted[c->device & 15] = smem;
    }
    dim3 grid(B, c->cfg.heads);
    k_attention<NJ><<<grid, 128, smem, stream>>>(c->qkv, T, d, c->o);
    count_launch();
    return D2R_OK;
}

extern "C" int d2r_clip_encode(d2r_clip* c, const void* patches_dev, int B, float* embeds_out_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D2R_REQUIRE(c && patches_dev && embeds_out_dev, "d2r_clip_encode: null argument");
    D2R_REQUIRE(B > 0 && B <= c->cfg.max_batch, "d2r_clip_encode: batch exceeds max_batch");
    D2R_CUDA(cudaSetDevice(c->device));
    const int d = c->cfg.hidden, T = c->T, mlp = c->cfg.mlp, M = B * T;
    const float eps = c->cfg.ln_eps;
    int rc;
    // patch embedding: conv(stride = kernel = P, no bias) == GEMM over patch-major pixels
    rc = gemm_f16((const __half*)patches_dev, c->Kp, c->w_patch, c->Kp, B * c->NP, d, c->Kp, nullptr, GEMM_OUT_F32, c->pe, d, stream);
    if (rc) return rc;
    const int rows_per_block = 8;
    k_embed_preln<<<(M + rows_per_block - 1) / rows_per_block, rows_per_block * 32, 0, stream>>>(c->pe, c->cls, c->pos, c->pre_w, c->pre_b,
                                                                                               eps, B, T, d, c->x);
    count_launch();
    for (const d2r_clip::Layer& y : c->layers) {
        k_layernorm_f16<<<(M + 7) / 8, 256, 0, stream>>>(c->x, y.ln1_w, y.ln1_b, eps, M, d, (size_t)d, c->h);
        count_launch();
        rc = gemm_f16(c->h, d, y.w_qkv, d, M, 3 * d, d, y.b_qkv, GEMM_OUT_F16, c->qkv, 3 * d, stream);
        if (rc) return rc;
        const int nj = (T + 31) / 32;
        if (nj <= 2) rc = launch_attention<2>(c, B, stream);
        else if (nj <= 7) rc = launch_attention<7>(c, B, stream);
        else if (nj <= 9) rc = launch_attention<9>(c, B, stream);
        else rc = launch_attention<19>(c, B, stream);
        if (rc) return rc;
        rc = gemm_f16(c->o, d, y.w_o, d, M, d, d, y.b_o, GEMM_RESIDUAL_F32, c->x, d, stream);
        if (rc) return rc;
        k_layernorm_f16<<<(M + 7) / 8, 256, 0, stream>>>(c->x, y.ln2_w, y.ln2_b, eps, M, d, (size_t)d, c->h);
        count_launch();
        rc = gemm_f16(c->h, d, y.w_1, d, M, mlp, d, y.b_1, GEMM_OUT_F16_QUICKGELU, c->m, mlp, stream);
        if (rc) return rc;
        rc = gemm_f16(c->m, mlp, y.w_2, mlp, M, d, mlp, y.b_2, GEMM_RESIDUAL_F32, c->x, d, stream);
        if (rc) return rc;
    }
    // pooled = post_layernorm(x[:, 0]); image_embeds = visual_projection(pooled); L2 normalise
    k_layernorm_f16<<<(B + 7) / 8, 256, 0, stream>>>(c->x, c->post_w, c->post_b, eps, B, d, (size_t)T * d, c->pooled);
    count_launch();
    rc = gemm_f16(c->pooled, d, c->w_proj, d, B, c->cfg.proj, d, nullptr, GEMM_OUT_F32, c->emb, c->cfg.proj, stream);
    if (rc) return rc;
    k_l2norm<<<(B + 7) / 8, 256, 0, stream>>>(c->emb, B, c->cfg.proj, embeds_out_dev);
    count_launch();
    D2R_CUDA(cudaGetLastError());
    return D2R_OK;
}
