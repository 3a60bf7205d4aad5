// Micro-benchmark behind DESIGN.md's bound for k_mlp_round: do tcgen05.mma (accumulators in TMEM), the epilogue's tcgen05.ld
// and its shared-memory stores run side by side on one SM, or do they take turns?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dream2real_b200/csrc tools/tmem_port_bench.cu -o gpurun_out/tmem_port_bench
//   gpurun_out/tmem_port_bench
//
// One CTA per SM, G groups of (4 epilogue warps + 1 MMA-issuing warp).  Per iteration a group's unit of work is what one hidden
// layer of one 128-sample tile costs in k_mlp_round: 4 MMAs (M128 N64 K16, fp16 -> fp32), one 32x32b.x64 tcgen05.ld per
// epilogue warp, 8 uint4 shared-memory stores per epilogue thread.  bit 0 of the mode = MMAs on, bit 1 = loads, bit 2 = stores.
// Prints SM cycles per unit for every mode and G = 1..4 ("ports").
//
// Part 2 ("chain") is k_mlp_round's dependency chain without its global-memory side: per tile and layer
//   wait for the MMAs -> tcgen05.ld x64 -> ReLU + fp16 pack -> operand rows for the next layer -> fence -> group barrier -> 4 MMAs
// with the operand rows either in shared memory (SS: st.shared + fence.proxy.async, what k_mlp_round does) or in tensor memory
// (TS: tcgen05.st, the MMA reads A from TMEM: 96 columns per tile instead of 64).  G groups of 128 threads, 1 or 2 tiles per group.
// Part 0 checks the TS operand layout this relies on (row = lane, two fp16 per 32-bit column) against exact integer arithmetic.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "d2r_gemm.cuh"

using namespace d2r;

constexpr int ITERS = 4000;

__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, int acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}

template <int N>
__global__ void __launch_bounds__(640, 1) k_port(int mode, int G, unsigned long long* cycles, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    // per group: A tile 16 KB (K=64), B 8 KB, store target 16 KB
    constexpr int GROUP = 40960;
    __shared__ uint64_t bars[4][2];
    __shared__ uint32_t tmem_slot;
    __shared__ unsigned long long t_end;      // the latest finish of any thread (a clock read after the barrier can be scheduled before it)
    if (threadIdx.x == 0) t_end = 0;
    const int grp = threadIdx.x / 160, t = threadIdx.x % 160, warp = t >> 5, lane = t & 31;
    for (int i = threadIdx.x; i < G * GROUP / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int g = 0; g < 4; ++g) { mbar_init(&bars[g][0], 1); mbar_init(&bars[g][1], 1); }
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot + grp * 128;
    unsigned char* gs = smem + grp * GROUP;
    const unsigned long long t0 = clock64();
    float acc = 0.f;
    if (grp < G) {
        if (warp == 4) {
            if (lane == 0 && (mode & 1)) {
                const uint32_t a_lo = (smem_u32(gs) >> 4) + (8u << 16), b_lo = (smem_u32(gs + 16384) >> 4) + (8u << 16);
                constexpr uint32_t hi = (uint32_t)((64 / 8) * 128 >> 4) | (1u << 14);
                constexpr uint32_t idesc = umma_idesc_f16(128, N, 0);
                uint32_t ph[2] = {0, 0};
                for (int i = 0; i < ITERS; ++i) {
                    const int b = i & 1;
                    if (i >= 2) { mbar_wait(&bars[grp][b], ph[b]); ph[b] ^= 1; }
                    tc_fence_after();
                    for (int kk = 0; kk < 4; ++kk) mma_ss(tmem + b * 64, a_lo + kk * 16, b_lo + kk * 16, hi, idesc, kk > 0);
                    tc_commit(&bars[grp][b]);
                }
                mbar_wait(&bars[grp][0], ph[0]);
                mbar_wait(&bars[grp][1], ph[1]);
            }
            __syncwarp();      // the whole warp reaches the CTA barrier together
        } else if (mode & 6) {
            const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
            unsigned char* row = gs + 24576 + (t >> 3) * 1024 + (t & 7) * 16;
            for (int i = 0; i < ITERS; ++i) {
                uint32_t r[64];
                if (mode & 2) {
                    tmem_ld_32x32_x64(tl + (i & 1) * 64, r);      // in k_mlp_round the columns an MMA finished earlier; here: whatever is there
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 64; ++j) r[j] = i;
                }
                if (mode & 4) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4*>(row + c * 128) = make_uint4(r[8 * c] ^ r[8 * c + 1], r[8 * c + 2] ^ r[8 * c + 3], r[8 * c + 4] ^ r[8 * c + 5], r[8 * c + 6] ^ r[8 * c + 7]);
                } else {
                    acc += __uint_as_float(r[0] ^ r[63]);      // the loads are volatile: nothing else has to consume them
                }
            }
        }
    }
    atomicMax(&t_end, (unsigned long long)clock64());
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t_end - t0;
    if (acc == 123.456f) sink[0] = acc;
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem_slot); }
}


// ---- TS mode: A operand in tensor memory ----------------------------------------------------------
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t hi, uint32_t idesc, int acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 db;\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2_relu(uint32_t a, uint32_t b) {
    const __half2 h = __floats2half2_rn(fmaxf(__uint_as_float(a), 0.f), fmaxf(__uint_as_float(b), 0.f));
    return *reinterpret_cast<const uint32_t*>(&h);
}
// K-major, no swizzle: 8x8 core matrices of 128 contiguous bytes, K-adjacent ones 128 B apart, 8-row groups K/8*128 B apart
__device__ __forceinline__ int kmajor_off(int row, int k, int K) { return (row >> 3) * (K / 8 * 128) + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2; }

// Part 0: D[128 x 64] = A[128 x 64] * W[64 x 64]^T with A written by tcgen05.st; small integers / 8, so fp32 sums are exact.
__global__ void __launch_bounds__(128, 1) k_ts_check(unsigned int* mismatches) {
    __shared__ __align__(128) unsigned char wsm[64 * 64 * 2];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    auto a_val = [](int r, int k) { return (float)((r + 3 * k) % 11 - 5) * 0.125f; };
    auto w_val = [](int n, int k) { return (float)((2 * n + k) % 7 - 3) * 0.25f; };
    for (int i = t; i < 64 * 64; i += 128) {
        const int n = i >> 6, k = i & 63;
        *reinterpret_cast<__half*>(wsm + kmajor_off(n, k, 64)) = __float2half(w_val(n, k));
    }
    if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<128>(&tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot, tl = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const __half2 h = __floats2half2_rn(a_val(t, 2 * j), a_val(t, 2 * j + 1));      // k even in the low half
        a[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tmem_st_32x32_x32(tl + 64, a);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
        tc_fence_after();
        const uint32_t b_lo = (smem_u32(wsm) >> 4) + (8u << 16);
        constexpr uint32_t hi = (uint32_t)((64 / 8) * 128 >> 4) | (1u << 14);
        constexpr uint32_t idesc = umma_idesc_f16(128, 64, 0);
        for (int kk = 0; kk < 4; ++kk) mma_ts(tmem, tmem + 64 + kk * 8, b_lo + kk * 16, hi, idesc, kk > 0);
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t d[64];
    tmem_ld_32x32_x64(tl, d);
    tmem_ld_wait();
    unsigned int bad = 0;
    for (int n = 0; n < 64; ++n) {
        float ref = 0.f;
        for (int k = 0; k < 64; ++k) ref += a_val(t, k) * w_val(n, k);
        bad += __uint_as_float(d[n]) != ref;
    }
    if (bad) atomicAdd(mismatches, bad);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<128>(tmem); }
}

// Part 2: the per-tile dependency chain.  flags: 1 = TS (operand rows in TMEM), 2 = leave out fence.proxy.async (timing only),
// 4 = leave out the operand-row stores (timing only), 8 = N = 16 MMAs
__global__ void __launch_bounds__(640, 1) k_chain(int G, int TPG, int flags, int iters, unsigned long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];      // [8 KB weights][tiles x 16 KB operand rows]
    __shared__ uint64_t bars[5][2];
    __shared__ uint32_t tmem_slot;
    __shared__ unsigned long long t_end;
    // warp index as a value the compiler knows to be warp-uniform: the descriptors built from it live in uniform registers
    const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int grp = (flags & 16) ? warp_u >> 2 : threadIdx.x >> 7, t = threadIdx.x & 127, warp = (flags & 16) ? warp_u & 3 : t >> 5;
    const bool ts = flags & 1;
    for (int i = threadIdx.x; i < (8192 + 8 * 16384) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        t_end = 0;
        for (int g = 0; g < 5; ++g) { mbar_init(&bars[g][0], 1); mbar_init(&bars[g][1], 1); }
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned long long t0 = clock64();
    if (grp < G) {
        const uint32_t b_lo = (smem_u32(smem) >> 4) + (8u << 16);
        constexpr uint32_t hi = (uint32_t)((64 / 8) * 128 >> 4) | (1u << 14);
        const uint32_t idesc = (flags & 8) ? umma_idesc_f16(128, 16, 0) : umma_idesc_f16(128, 64, 0);
        const int cols = ts ? 96 : 64;
        uint32_t ph[2] = {0, 0};
        auto issue = [&](int s) {
            const int tile = grp * TPG + s;
            const uint32_t d = tmem_slot + tile * cols;
            if (ts) {
                for (int kk = 0; kk < 4; ++kk) mma_ts(d, d + 64 + kk * 8, b_lo + kk * 16, hi, idesc, kk > 0);
            } else {
                const uint32_t a_lo = (smem_u32(smem + 8192 + tile * 16384) >> 4) + (8u << 16);
                for (int kk = 0; kk < 4; ++kk) mma_ss(d, a_lo + kk * 16, b_lo + kk * 16, hi, idesc, kk > 0);
            }
            tc_commit(&bars[grp][s]);
        };
        if (flags & 16) {
            if (warp == 0) for (int s = 0; s < TPG; ++s) { if (elect_one()) issue(s); __syncwarp(); }
        } else if (t == 0) for (int s = 0; s < TPG; ++s) issue(s);
        for (int i = 0; i < iters; ++i) {
            for (int s = 0; s < TPG; ++s) {
                const int tile = grp * TPG + s;
                mbar_wait(&bars[grp][s], ph[s]);
                ph[s] ^= 1;
                tc_fence_after();
                const uint32_t tl = tmem_slot + tile * cols + ((uint32_t)(warp * 32) << 16);
                uint32_t r[64];
                tmem_ld_32x32_x64(tl, r);
                tmem_ld_wait();
                uint32_t v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = pack2_relu(r[2 * j], r[2 * j + 1]);
                if (ts) {
                    if (!(flags & 4)) { tmem_st_32x32_x32(tl + 64, v); tmem_st_wait(); }
                } else {
                    unsigned char* row = smem + 8192 + tile * 16384 + kmajor_off(t, 0, 64);
                    if (!(flags & 4)) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(row + c * 128) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    }
                    if (!(flags & 2)) fence_proxy_async();
                }
                if (flags & 4) { uint32_t x = 0; for (int j = 0; j < 32; ++j) x ^= v[j]; if (x == 0x12345u) cycles[200] = x; }
                tc_fence_before();
                asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
                if (flags & 16) {      // the whole first warp of the group walks the issue path; one elected lane issues
                    if (warp == 0) { tc_fence_after(); if (elect_one()) issue(s); __syncwarp(); }
                } else if (t == 0) { tc_fence_after(); issue(s); }
            }
        }
        for (int s = 0; s < TPG; ++s) mbar_wait(&bars[grp][s], ph[s]);
    }
    atomicMax(&t_end, (unsigned long long)clock64());
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t_end - t0;
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem_slot); }
}

static double run_chain(int G, int TPG, int flags, unsigned long long* d_cyc, int n_sm) {
    const int iters = 2000, smem = 8192 + 8 * 16384;
    cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) k_chain<<<n_sm, 640, smem>>>(G, TPG, flags, iters, d_cyc);
    if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("k_chain failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    unsigned long long h[256];
    cudaMemcpy(h, d_cyc, n_sm * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < n_sm; ++i) s += (double)h[i];
    return s / n_sm / iters / (G * TPG);      // SM cycles per tile-layer
}

template <int N>
static double run(int mode, int G, unsigned long long* d_cyc, float* d_sink, int n_sm) {
    cudaFuncSetAttribute(k_port<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 40960);
    k_port<N><<<n_sm, 640, 4 * 40960>>>(mode, G, d_cyc, d_sink);
    k_port<N><<<n_sm, 640, 4 * 40960>>>(mode, G, d_cyc, d_sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    unsigned long long h[256];
    cudaMemcpy(h, d_cyc, n_sm * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < n_sm; ++i) s += (double)h[i];
    return s / n_sm / ITERS;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int n_sm = p.multiProcessorCount;
    unsigned long long* d_cyc;
    float* d_sink;
    cudaMalloc(&d_cyc, 256 * sizeof(unsigned long long));
    cudaMalloc(&d_sink, 4);
    unsigned int* d_bad;
    cudaMalloc(&d_bad, 4);
    cudaMemset(d_bad, 0, 4);
    k_ts_check<<<1, 128>>>(d_bad);
    unsigned int bad = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("k_ts_check failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
    const char* names[8] = {"nothing", "mma", "ld", "mma+ld", "st", "mma+st", "ld+st", "mma+ld+st"};
    printf("{\"device\": \"%s\", \"sms\": %d, \"ts_operand_layout_mismatches_of_8192\": %u, \"unit\": \"SM cycles per (4 MMAs M128 N* K16 | x64 tcgen05.ld per warp | 8 uint4 st.shared per thread) per group\",\n", p.name, n_sm, bad);
    printf(" \"chain_unit\": \"SM cycles per tile-layer (SM time / layers / tiles in flight); k_mlp_round = SS, 3 groups x 2 tiles\",\n \"chain\": {\n");
    {
        struct Cfg { const char* name; int G, TPG, flags; };
        const Cfg cfgs[] = {
            {"SS g1x2", 1, 2, 0}, {"SS g2x2", 2, 2, 0}, {"SS g3x2", 3, 2, 0}, {"SS g4x2", 4, 2, 0}, {"SS g5x1", 5, 1, 0},
            {"SS g3x2 no proxy fence", 3, 2, 2}, {"SS g3x2 no stores", 3, 2, 4 | 2}, {"SS g3x2 N16", 3, 2, 8}, {"SS g4x2 N16", 4, 2, 8},
            {"TS g1x2", 1, 2, 1}, {"TS g2x2", 2, 2, 1}, {"TS g3x1", 3, 1, 1}, {"TS g4x1", 4, 1, 1}, {"TS g5x1", 5, 1, 1},
            {"SS g3x2 uniform issue", 3, 2, 16}, {"SS g4x2 uniform issue", 4, 2, 16}, {"TS g5x1 uniform issue", 5, 1, 1 | 16}, {"TS g2x2 uniform issue", 2, 2, 1 | 16},
            {"TS g2x2 N16", 2, 2, 1 | 8}, {"TS g5x1 N16", 5, 1, 1 | 8}, {"TS g2x2 no stores", 2, 2, 1 | 4},
        };
        const int n = sizeof(cfgs) / sizeof(cfgs[0]);
        for (int i = 0; i < n; ++i) printf("  \"%s\": %.1f%s\n", cfgs[i].name, run_chain(cfgs[i].G, cfgs[i].TPG, cfgs[i].flags, d_cyc, n_sm), i + 1 < n ? "," : "");
    }
    printf(" },\n");
    for (int n = 0; n < 2; ++n) {
        printf(" \"N=%d\": {\n", n == 0 ? 64 : 16);
        for (int G = 1; G <= 4; ++G) {
            printf("  \"groups=%d\": {", G);
            for (int mode = 1; mode < 8; ++mode) {
                const double c = n == 0 ? run<64>(mode, G, d_cyc, d_sink, n_sm) : run<16>(mode, G, d_cyc, d_sink, n_sm);
                printf("\"%s\": %.1f%s", names[mode], c, mode < 7 ? ", " : "");
            }
            printf("}%s\n", G < 4 ? "," : "");
        }
        printf(" }%s\n", n == 0 ? "," : "");
    }
    printf("}\n");
    return 0;
}
