// internal interface of the tcgen05 GEMM (d2r_gemm.cu)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

enum {
    GEMM_OUT_F16 = 0,            // out_f16 = A.B^T + bias
    GEMM_OUT_F16_QUICKGELU = 1,  // out_f16 = quick_gelu(A.B^T + bias)
    GEMM_RESIDUAL_F32 = 2,       // out_f32 += A.B^T + bias   (residual stream, in place)
    GEMM_OUT_F32 = 3             // out_f32 = A.B^T + bias
};

namespace d2r {
int gemm_f16(const __half* A, int lda, const __half* B, int ldb, int M, int N, int K, const float* bias, int mode, void* out, int ldo,
             cudaStream_t stream);
}
