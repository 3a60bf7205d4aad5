// fp16 x fp16 -> fp32 GEMM on the 5th-generation tensor cores: C[M,N] = A[M,K] . B[N,K]^T (+ epilogue).
// Both operands K-major ("TN"), i.e. activations [rows, features] times nn.Linear weights [out, in].
//
// Structure (persistent: one CTA per SM walks 128 x BN output tiles; 192 threads; two TMEM accumulators
// so the epilogue of one tile overlaps the main loop of the next):
//   warp 0      TMA producer : cp.async.bulk.tensor 2-D loads (128B swizzle) into a 6-stage smem ring,
//                              completion via mbarrier complete_tx
//   warp 1      MMA issuer   : one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//                              (M=128, N=BN, K=16) x4 per 64-wide k block; accumulator in TMEM;
//                              tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2..9  epilogue     : two warps per TMEM lane quadrant (each half of the columns): tcgen05.ld
//                              (32 lanes x 32 columns) -> bias / quick-GELU -> 32-row x 128-byte staging tile in
//                              shared memory (128B swizzle, conflict-free) -> ONE TMA store per chunk
//                              (cp.async.bulk.tensor, full 128-byte rows); the fp32 residual mode uses the TMA
//                              reduce-add (cp.reduce.async.bulk.tensor .add): x += A.W^T + b happens in L2, the
//                              old value never travels to the SM
// Used for every dense contraction of the CLIP vision tower (reference clip_scoring.py:180-181 ->
// transformers CLIPModel.forward -> nn.Linear / patch-embedding conv).
#include <dlfcn.h>

#include <algorithm>
#include <mutex>

#include "d2r_common.cuh"
#include "d2r_gemm.cuh"
#include "d2r_gemm_api.h"

namespace d2r {

constexpr int BM = 128, BK = 64, GEMM_THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps

template <int BN>
struct GemmSmem {
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);   // 4 x 48 KB / 6 x 32 KB / 8 x 24 KB
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;           // epilogue staging: 8 warps x [32 rows x 128 B]
    static constexpr int STG_BYTES = 8 * 4096;
    static constexpr int BAR_OFFSET = STG_OFFSET + STG_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // barriers + tmem slot + alignment slack
};

struct GemmArgs {
    int M, N, K;
    const float* bias;     // [N] or null
    __half* out_f16;       // modes 0,1
    float* out_f32;        // modes 2 (in-place residual), 3
    int ldo;               // leading dimension of the output (elements)
    int mode;
};

__device__ __forceinline__ float quick_gelu(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }   // x * sigmoid(1.702 x)

// Persistent: grid = #SMs, every CTA walks output tiles (n fastest, so CTAs running together share the
// same rows of A through L2); TMEM holds TWO accumulators so the epilogue of tile i overlaps the main
// loop of tile i+1.
template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_f16(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
           const __grid_constant__ CUtensorMap tma_o, const GemmArgs g) {
    using S = GemmSmem<BN>;
    constexpr int STAGES = S::STAGES;
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B needs 1024-byte aligned tiles
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int num_kb = g.K / BK;
    const int tiles_n = g.N / BN, tiles_m = (g.M + BM - 1) / BM;
    const int num_tiles = tiles_n * tiles_m;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        tma_prefetch_desc(&tma_o);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], BN == 64 ? 4 : 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;   // running k-block counter across tiles -> stage / phase
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    unsigned char* sa = smem + s * S::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[s], S::STAGE_BYTES);
                    tma_load_2d(sa, &tma_a, &full_bar[s], kb * BK, m0);
                    tma_load_2d(sa + S::A_BYTES, &tma_b, &full_bar[s], kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(BM, BN, /*fp16*/ 0);
            uint32_t it = 0, tile_it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
                const uint32_t acc = tile_it & 1;
                mbar_wait(&tmem_empty_bar[acc], ((tile_it >> 1) & 1) ^ 1);     // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
                    const uint64_t da = umma_desc_sw128(a_addr), db = umma_desc_sw128(a_addr + S::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)   // advance 16 elements = 32 bytes = 2 descriptor units along K
                        umma_f16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    tc_commit(&empty_bar[s]);           // smem stage free once these MMAs retire
                }
                tc_commit(&tmem_full_bar[acc]);         // accumulator complete
            }
        }
    } else {
        // epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
        const int q = warp & 3;
        const int ew = warp - 2;
        const int chalf = ew >> 2;             // 0: columns [0, BN/2), 1: [BN/2, BN); BN == 64: warps 6..9 idle
        constexpr int COLS_PER_WARP = BN == 64 ? 64 : BN / 2;
        const bool f16out = g.mode == GEMM_OUT_F16 || g.mode == GEMM_OUT_F16_QUICKGELU;
        // staging tile of this warp: 32 rows x 128 B, SWIZZLE_128B (16-byte chunk j of row r sits at chunk j ^ (r & 7))
        unsigned char* stg = smem + S::STG_OFFSET + ew * 4096;
        const uint32_t stg_u32 = smem_u32(stg);
        unsigned char* my_row = stg + lane * 128;
        const int sw = lane & 7;
        if (BN != 64 || chalf == 0) {
            uint32_t tile_it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
                const uint32_t acc = tile_it & 1;
                const int row0 = m0 + q * 32;
                mbar_wait(&tmem_full_bar[acc], (tile_it >> 1) & 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
                const int cstep = f16out ? 64 : 32;
#pragma unroll 1
                for (int c0 = chalf * COLS_PER_WARP; c0 < (chalf + 1) * COLS_PER_WARP; c0 += cstep) {
                    // the previous TMA store of this warp must have finished reading the staging tile
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll 1
                    for (int h = 0; h < (f16out ? 2 : 1); ++h) {
                        uint32_t r[32];
                        tmem_ld_32x32(tacc + (uint32_t)(c0 + h * 32), r);
                        tmem_ld_wait();
                        const int col0 = n0 + c0 + h * 32;
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        if (g.bias) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b = *reinterpret_cast<const float4*>(g.bias + col0 + j);
                                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                            }
                        }
                        if (f16out) {
                            if (g.mode == GEMM_OUT_F16_QUICKGELU) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {      // 4 x 16 B = 32 fp16 columns
                                uint4 pk;
                                __half2 h0 = __floats2half2_rn(v[8 * j], v[8 * j + 1]), h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
                                __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]), h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
                                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                                *reinterpret_cast<uint4*>(my_row + (((h * 4 + j) ^ sw) << 4)) = pk;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)        // 8 x 16 B = 32 fp32 columns
                                *reinterpret_cast<float4*>(my_row + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                    }
                    fence_proxy_async();          // generic-proxy writes -> visible to the TMA engine
                    __syncwarp();
                    if (lane == 0 && row0 < g.M) {   // rows beyond M are clipped by the tensor map
                        if (g.mode == GEMM_RESIDUAL_F32) tma_reduce_add_2d(&tma_o, stg_u32, n0 + c0, row0);
                        else tma_store_2d(&tma_o, stg_u32, n0 + c0, row0);
                        tma_store_commit();
                    }
                }
                // this warp is done reading the accumulator: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            }
            if (lane == 0) tma_store_wait_read<0>();   // the staging tile must outlive the last store
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * BN>(tmem_base);
    }
}

// ---- host side: tensor maps through the driver entry point (no link-time libcuda dependency) ------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

static int get_encode() {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = (PFN_encodeTiled)fn;
    });
    if (!g_encode) { set_error("cuTensorMapEncodeTiled is not available from the CUDA driver"); return D2R_ERR_CUDA; }
    return D2R_OK;
}

// 2-D row-major [rows, cols] tensor (fp16, or fp32 for epilogue outputs), box = [box_rows, 128 bytes of columns],
// 128-byte swizzle, OOB -> zeros on loads / clipped on stores
int make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows, bool f32) {
    int rc = get_encode();
    if (rc) return rc;
    const size_t es = f32 ? sizeof(float) : sizeof(__half);
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld_elems * es};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / es), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (base must be 16-byte aligned, "
                  "row pitch a multiple of 16 bytes)");
        return D2R_ERR_CUDA;
    }
    return D2R_OK;
}

template <int BN>
static int launch_gemm_bn(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmArgs& g, cudaStream_t stream) {
    static bool attr_done[16] = {false};
    int dev;
    D2R_CUDA(cudaGetDevice(&dev));
    if (!attr_done[dev & 15]) {
        D2R_CUDA(cudaFuncSetAttribute(k_gemm_f16<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::TOTAL));
        attr_done[dev & 15] = true;
    }
    static int n_sm[16] = {0};
    if (!n_sm[dev & 15]) D2R_CUDA(cudaDeviceGetAttribute(&n_sm[dev & 15], cudaDevAttrMultiProcessorCount, dev));
    const int num_tiles = ((g.M + BM - 1) / BM) * (g.N / BN);
    k_gemm_f16<BN><<<std::min(num_tiles, n_sm[dev & 15]), GEMM_THREADS, GemmSmem<BN>::TOTAL, stream>>>(ta, tb, to, g);
    count_launch();
    D2R_CUDA(cudaGetLastError());
    return D2R_OK;
}

int gemm_f16(const __half* A, int lda, const __half* B, int ldb, int M, int N, int K, const float* bias, int mode, void* out, int ldo,
             cudaStream_t stream) {
    D2R_REQUIRE(A && B && out, "gemm_f16: null argument");
    D2R_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_f16: empty problem");
    D2R_REQUIRE(K % BK == 0, "gemm_f16: K must be a multiple of 64 (pad the operands)");
    D2R_REQUIRE(N % 64 == 0, "gemm_f16: N must be a multiple of 64");
    D2R_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0, "gemm_f16: leading dimensions must be multiples of 8 elements");
    D2R_REQUIRE(mode >= 0 && mode <= 3, "gemm_f16: bad epilogue mode");
    // 256-wide tiles move 25 % fewer operand bytes per MAC through L2 than 128-wide ones; use them when they fill the SMs
    const long tiles256 = (long)((M + BM - 1) / BM) * (N / 256);
    const int BN = (N % 256 == 0 && tiles256 >= 296) ? 256 : ((N % 128 == 0) ? 128 : 64);
    const bool f32out = mode == GEMM_RESIDUAL_F32 || mode == GEMM_OUT_F32;
    D2R_REQUIRE(((uintptr_t)out & 15) == 0, "gemm_f16: the output must be 16-byte aligned");
    CUtensorMap ta, tb, to;
    int rc = make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, false);
    if (rc) return rc;
    rc = make_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)BN, false);
    if (rc) return rc;
    rc = make_tmap(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 32, f32out);   // epilogue chunks: 32 rows x 128 B
    if (rc) return rc;
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.bias = bias; g.mode = mode; g.ldo = ldo;
    g.out_f16 = (mode == GEMM_OUT_F16 || mode == GEMM_OUT_F16_QUICKGELU) ? (__half*)out : nullptr;
    g.out_f32 = (mode == GEMM_RESIDUAL_F32 || mode == GEMM_OUT_F32) ? (float*)out : nullptr;
    if (BN == 256) return launch_gemm_bn<256>(ta, tb, to, g, stream);
    return BN == 128 ? launch_gemm_bn<128>(ta, tb, to, g, stream) : launch_gemm_bn<64>(ta, tb, to, g, stream);
}

}  // namespace d2r

extern "C" int d2r_gemm_f16(const void* a_dev, int lda, const void* b_dev, int ldb, int M, int N, int K, const float* bias_dev, int mode,
                            void* out_dev, int ldo, void* stream) {
    return d2r::gemm_f16((const __half*)a_dev, lda, (const __half*)b_dev, ldb, M, N, K, bias_dev, mode, out_dev, ldo, (cudaStream_t)stream);
}
