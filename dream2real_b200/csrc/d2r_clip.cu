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

#include <stdlib.h>
#include <string.h>

#include <algorithm>

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
    if ((d & 127) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        // vector path: a lane owns 4 consecutive columns of every 128-column group (float4 loads, 8/16-byte stores)
        float4 v[8];
        const int n = d / 128;   // d <= 1024
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < n) { v[i] = *reinterpret_cast<const float4*>(in + i * 128 + lane * 4); s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < n) {
                const float a = v[i].x - mean, b2 = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
                q += (a * a + b2 * b2) + (c * c + e * e);
            }
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q / (float)d + eps);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < n) {
                const int c = i * 128 + lane * 4;
                const float4 ww = *reinterpret_cast<const float4*>(w + c), bb = *reinterpret_cast<const float4*>(b + c);
                const float y0 = (v[i].x - mean) * rstd * ww.x + bb.x, y1 = (v[i].y - mean) * rstd * ww.y + bb.y;
                const float y2 = (v[i].z - mean) * rstd * ww.z + bb.z, y3 = (v[i].w - mean) * rstd * ww.w + bb.w;
                if constexpr (sizeof(OUT) == 2) {
                    __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    *reinterpret_cast<uint2*>(out + c) = pk;
                } else {
                    *reinterpret_cast<float4*>(out + c) = make_float4(y0, y1, y2, y3);
                }
            }
        return;
    }
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

// softmax(q k^T) v, flash-attention style, for head_dim 64; q already carries the 1/sqrt(head_dim)
// scale.  One CTA = 64 query rows of one (image, head); each of the 4 warps owns 16 rows and walks the
// keys in blocks of 64: S = Q.K^T and O += P.V on the warp-level tensor-core path (mma.sync m16n8k16,
// fp16 operands, fp32 accumulate), online softmax in fp32 registers.  The ViT contractions proper (98 %
// of the FLOPs at T = 50) run on tcgen05 (d2r_gemm.cu); a 50 x 50 x 64 problem per head does not fill
// an M = 128 UMMA tile, which is why the attention core uses warp MMAs.
constexpr int ATT_HD = 64, ATT_BQ = 64, ATT_BK = 64, ATT_LD = 72;   // smem row pitch (halves): 144 B, conflict-free ldmatrix

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const __half* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const __half* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) k_attention(const __half* __restrict__ qkv, int T, int d, __half* __restrict__ out) {
    __shared__ __align__(16) __half Qs[ATT_BQ * ATT_LD];
    __shared__ __align__(16) __half Ks[ATT_BK * ATT_LD];
    __shared__ __align__(16) __half Vs[ATT_BK * ATT_LD];
    const int img = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * ATT_BQ;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const __half* base = qkv + (size_t)img * T * 3 * d + head * ATT_HD;
    // stage this CTA's query rows (zero beyond T)
    for (int i = tid; i < ATT_BQ * (ATT_HD / 8); i += 128) {
        const int r = i / (ATT_HD / 8), c8 = i % (ATT_HD / 8);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q0 + r < T) v = *reinterpret_cast<const uint4*>(base + (size_t)(q0 + r) * 3 * d + c8 * 8);
        *reinterpret_cast<uint4*>(Qs + r * ATT_LD + c8 * 8) = v;
    }
    __syncthreads();
    uint32_t qf[4][4];   // A fragments of this warp's 16 rows, 4 k-steps of 16 dims
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(qf[kk], Qs + (warp * 16 + (lane & 15)) * ATT_LD + kk * 16 + (lane >> 4) * 8);
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;    // rows g and g+8

    for (int k0 = 0; k0 < T; k0 += ATT_BK) {
        __syncthreads();
        for (int i = tid; i < ATT_BK * (ATT_HD / 8); i += 128) {
            const int r = i / (ATT_HD / 8), c8 = i % (ATT_HD / 8);
            uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
            if (k0 + r < T) {
                kv = *reinterpret_cast<const uint4*>(base + (size_t)(k0 + r) * 3 * d + d + c8 * 8);
                vv = *reinterpret_cast<const uint4*>(base + (size_t)(k0 + r) * 3 * d + 2 * d + c8 * 8);
            }
            *reinterpret_cast<uint4*>(Ks + r * ATT_LD + c8 * 8) = kv;
            *reinterpret_cast<uint4*>(Vs + r * ATT_LD + c8 * 8) = vv;
        }
        __syncthreads();
        // S = Q . K^T : 8 n-tiles of 8 keys
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {     // two n-tiles per ldmatrix.x4
                uint32_t bfr[4];
                // matrices: (keys 16jp..+7, dims 16kk..+7), (same keys, dims +8), (keys +8, dims ..+7), (keys +8, dims +8)
                ldmatrix_x4(bfr, Ks + (jp * 16 + (lane & 7) + ((lane >> 4) << 3)) * ATT_LD + kk * 16 + ((lane >> 3) & 1) * 8);
                mma_16816(s[2 * jp], qf[kk], bfr[0], bfr[1]);
                mma_16816(s[2 * jp + 1], qf[kk], bfr[2], bfr[3]);
            }
        }
        // mask keys beyond T, online softmax
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int key = k0 + j * 8 + 2 * t4;
            if (key >= T) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
            if (key + 1 >= T) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = expf(m0 - mx0), c1 = expf(m1 - mx1);     // 0 on the first block (m = -inf)
        m0 = mx0; m1 = mx1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[8][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = expf(s[j][0] - mx0), p1 = expf(s[j][1] - mx0), p2 = expf(s[j][2] - mx1), p3 = expf(s[j][3] - mx1);
            rs0 += p0 + p1; rs1 += p2 + p3;
            __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
            pf[j][0] = *reinterpret_cast<uint32_t*>(&h01);
            pf[j][1] = *reinterpret_cast<uint32_t*>(&h23);
            o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
        }
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
        // O += P . V : k-steps of 16 keys, 8 n-tiles of 8 dims
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t af[4] = {pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1]};
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t bfr[4];
                // transposed 8x8 loads of V[keys][dims]: (keys 16kk..+7, dims 16jp..+7), (keys +8, same dims), (keys ..+7, dims +8), (keys +8, dims +8)
                ldmatrix_x4_trans(bfr, Vs + (kk * 16 + (lane & 15)) * ATT_LD + jp * 16 + (lane >> 4) * 8);
                mma_16816(o[2 * jp], af, bfr[0], bfr[1]);
                mma_16816(o[2 * jp + 1], af, bfr[2], bfr[3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = head * ATT_HD + j * 8 + 2 * t4;
        if (r0 < T) *reinterpret_cast<__half2*>(out + ((size_t)img * T + r0) * d + col) = __floats2half2_rn(o[j][0] * i0, o[j][1] * i0);
        if (r1 < T) *reinterpret_cast<__half2*>(out + ((size_t)img * T + r1) * d + col) = __floats2half2_rn(o[j][2] * i1, o[j][3] * i1);
    }
}

// Persistent, double-buffered version of k_attention: a CTA walks work items (image, head, 64-query block) and, inside
// an item, the key blocks; while it computes one {Q, K, V} stage the next one is already in flight through cp.async
// (no registers held), so the global-load latency that dominated the one-item-per-CTA kernel at T = 50 is hidden.
// Arithmetic (mma.sync tiles, online softmax order) is identical to k_attention.
constexpr int ATT_STAGE_HALVES = 3 * 64 * ATT_LD;                  // Q, K, V tiles of one stage
constexpr int ATT2_SMEM = 2 * ATT_STAGE_HALVES * (int)sizeof(__half);

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 16 : 0;                                  // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(128, 3) k_attention2(const __half* __restrict__ qkv, int T, int d, int heads, int B, __half* __restrict__ out) {
    extern __shared__ __align__(16) __half att_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int qblocks = (T + ATT_BQ - 1) / ATT_BQ, kblocks = (T + ATT_BK - 1) / ATT_BK;
    const long n_items = (long)B * heads * qblocks;
    const long n_steps_total = n_items * kblocks;        // step = (item, key block), items strided over the grid
    // steps of this CTA: items blockIdx.x, blockIdx.x + gridDim.x, ...; each with kblocks steps
    const long my_items = n_items > blockIdx.x ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long my_steps = my_items * kblocks;
    (void)n_steps_total;

    auto issue = [&](long step) {                        // cp.async the tiles of local step `step` into stage step & 1
        const long item = blockIdx.x + (step / kblocks) * (long)gridDim.x;
        const int kb = (int)(step % kblocks);
        const int qb = (int)(item % qblocks), head = (int)((item / qblocks) % heads), img = (int)(item / ((long)qblocks * heads));
        const __half* base = qkv + (size_t)img * T * 3 * d + head * ATT_HD;
        __half* st = att_smem + (step & 1) * ATT_STAGE_HALVES;
        if (kb == 0) {
            for (int i = tid; i < ATT_BQ * (ATT_HD / 8); i += 128) {
                const int r = i / (ATT_HD / 8), c8 = i % (ATT_HD / 8);
                const bool ok = qb * ATT_BQ + r < T;
                cp_async_16(st + r * ATT_LD + c8 * 8, base + (size_t)(ok ? qb * ATT_BQ + r : 0) * 3 * d + c8 * 8, ok);
            }
        }
        for (int i = tid; i < ATT_BK * (ATT_HD / 8); i += 128) {
            const int r = i / (ATT_HD / 8), c8 = i % (ATT_HD / 8);
            const bool ok = kb * ATT_BK + r < T;
            const __half* row = base + (size_t)(ok ? kb * ATT_BK + r : 0) * 3 * d + c8 * 8;
            cp_async_16(st + (64 + r) * ATT_LD + c8 * 8, row + d, ok);
            cp_async_16(st + (128 + r) * ATT_LD + c8 * 8, row + 2 * d, ok);
        }
        cp_async_commit();
    };

    if (my_steps > 0) issue(0);
    uint32_t qf[4][4];
    float o[8][4];
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (long step = 0; step < my_steps; ++step) {
        if (step + 1 < my_steps) { issue(step + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const long item = blockIdx.x + (step / kblocks) * (long)gridDim.x;
        const int kb = (int)(step % kblocks), k0 = kb * ATT_BK;
        const int qb = (int)(item % qblocks), head = (int)((item / qblocks) % heads), img = (int)(item / ((long)qblocks * heads));
        const int q0 = qb * ATT_BQ;
        const __half* Qs = att_smem + (step & 1) * ATT_STAGE_HALVES;
        const __half* Ks = Qs + 64 * ATT_LD;
        const __half* Vs = Qs + 128 * ATT_LD;
        if (kb == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(qf[kk], Qs + (warp * 16 + (lane & 15)) * ATT_LD + kk * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
            m0 = -INFINITY; m1 = -INFINITY; l0 = 0.f; l1 = 0.f;
        }
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t bfr[4];
                ldmatrix_x4(bfr, Ks + (jp * 16 + (lane & 7) + ((lane >> 4) << 3)) * ATT_LD + kk * 16 + ((lane >> 3) & 1) * 8);
                mma_16816(s[2 * jp], qf[kk], bfr[0], bfr[1]);
                mma_16816(s[2 * jp + 1], qf[kk], bfr[2], bfr[3]);
            }
        }
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int key = k0 + j * 8 + 2 * t4;
            if (key >= T) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
            if (key + 1 >= T) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = expf(m0 - mx0), c1 = expf(m1 - mx1);
        m0 = mx0; m1 = mx1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[8][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = expf(s[j][0] - mx0), p1 = expf(s[j][1] - mx0), p2 = expf(s[j][2] - mx1), p3 = expf(s[j][3] - mx1);
            rs0 += p0 + p1; rs1 += p2 + p3;
            __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
            pf[j][0] = *reinterpret_cast<uint32_t*>(&h01);
            pf[j][1] = *reinterpret_cast<uint32_t*>(&h23);
            o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
        }
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t af[4] = {pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1]};
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t bfr[4];
                ldmatrix_x4_trans(bfr, Vs + (kk * 16 + (lane & 15)) * ATT_LD + jp * 16 + (lane >> 4) * 8);
                mma_16816(o[2 * jp], af, bfr[0], bfr[1]);
                mma_16816(o[2 * jp + 1], af, bfr[2], bfr[3]);
            }
        }
        if (kb == kblocks - 1) {
            float t0 = l0, t1 = l1;
            t0 += __shfl_xor_sync(0xffffffffu, t0, 1); t0 += __shfl_xor_sync(0xffffffffu, t0, 2);
            t1 += __shfl_xor_sync(0xffffffffu, t1, 1); t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
            const float i0 = 1.0f / t0, i1 = 1.0f / t1;
            const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = head * ATT_HD + j * 8 + 2 * t4;
                if (r0 < T) *reinterpret_cast<__half2*>(out + ((size_t)img * T + r0) * d + col) = __floats2half2_rn(o[j][0] * i0, o[j][1] * i0);
                if (r1 < T) *reinterpret_cast<__half2*>(out + ((size_t)img * T + r1) * d + col) = __floats2half2_rn(o[j][2] * i1, o[j][3] * i1);
            }
        }
        __syncthreads();      // every warp is done with this stage before the next iteration's prefetch overwrites it
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
    DeviceGuard dg(c->device);
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
    DeviceGuard dg(device);
    d2r_clip* c = new d2r_clip();
    c->device = device;
    c->cfg = *cfg;
    const int side = cfg->image_size / P;
    c->NP = side * side;
    c->T = c->NP + 1;
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

static int launch_attention(const d2r_clip* c, int B, cudaStream_t stream) {
    static const bool old_kernel = []() { const char* e = getenv("D2R_ATT"); return e && strcmp(e, "old") == 0; }();
    if (old_kernel) {
        dim3 grid((c->T + ATT_BQ - 1) / ATT_BQ, c->cfg.heads, B);
        k_attention<<<grid, 128, 0, stream>>>(c->qkv, c->T, c->cfg.hidden, c->o);
    } else {
        static bool attr_done[16] = {false};
        static int n_sm[16] = {0};
        const int dev = c->device & 15;
        if (!attr_done[dev]) {
            D2R_CUDA(cudaFuncSetAttribute(k_attention2, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
            D2R_CUDA(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, c->device));
            attr_done[dev] = true;
        }
        const long items = (long)B * c->cfg.heads * ((c->T + ATT_BQ - 1) / ATT_BQ);
        const int grid = (int)std::min<long>(items, (long)n_sm[dev] * 3);      // 54 KB of shared memory and ~144 registers per thread: 3 CTAs per SM
        k_attention2<<<grid, 128, ATT2_SMEM, stream>>>(c->qkv, c->T, c->cfg.hidden, c->cfg.heads, B, c->o);
    }
    count_launch();
    return D2R_OK;
}

extern "C" int d2r_clip_encode(d2r_clip* c, const void* patches_dev, int B, float* embeds_out_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D2R_REQUIRE(c && patches_dev && embeds_out_dev, "d2r_clip_encode: null argument");
    D2R_REQUIRE(B > 0 && B <= c->cfg.max_batch, "d2r_clip_encode: batch exceeds max_batch");
    DeviceGuard dg(c->device);
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
        rc = launch_attention(c, B, stream);
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
