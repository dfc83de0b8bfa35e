// Downwash MLP on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// The two wide layers (128->64 and 64->128, 93 % of the flops) run as tcgen05.mma kind::f16 with
// fp32 accumulators in tensor memory; the K=6 input layer and the N=3 output layer stay on the
// CUDA cores in fp32 and are fused around the MMAs, as are the relative-feature construction
// (other - ego)[0:6] and the 1 m gate (downwash_nn.py:21-29, ndp_nmpc_leader_node.py:65-76).
//
// Precision: the reference runs this net in fp32 (cuBLAS SGEMM).  Plain bf16/fp16 operands would
// miss the u0 parity target, so every operand is split into two fp16 halves
//     a = a_hi + 2^-11 a_lo,   a_hi = fp16(a),  a_lo = fp16((a - a_hi) * 2^11)
// and the product is accumulated as  D_main += a_hi b_hi,  D_corr += a_hi b_lo + a_lo b_hi,
// result = D_main + 2^-11 D_corr  (dropped term a_lo b_lo ~ 2^-22): three MMAs per k-step,
// ~fp32 accuracy with 16-bit operands.  Valid for |feature| < ~1e3 (fp16 range of the activations).
//
// CTA = 128 threads = 128 rows (thread t <-> row t <-> TMEM lane t), persistent over row tiles.
// Operands live in shared memory in the canonical no-swizzle K-major UMMA layout: 8x8 fp16 core
// matrices of 128 contiguous bytes, row-group stride SBO = 128 B, k-chunk stride LBO = (rows/8)*128 B.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mlp_kernel.cuh"

namespace ndp {

constexpr long long MLPT_MIN_ROWS = 2048;  // below this the fp32 CUDA-core kernel is used (latency path)
constexpr int MLPT_ROWS = 128;
constexpr int MLPT_THREADS = 128;
constexpr float MLPT_LO_SCALE = 2048.f;       // 2^11
constexpr float MLPT_LO_INV = 1.f / 2048.f;

// fp16 operand images (elements): W2 hi, W2 lo [64 x 128], W3 hi, W3 lo [128 x 64]
constexpr int MLPT_W2_ELEMS = MLP_H2 * MLP_H1;
constexpr int MLPT_W3_ELEMS = MLP_H3 * MLP_H2;
constexpr int MLPT_WIMG_ELEMS = 2 * MLPT_W2_ELEMS + 2 * MLPT_W3_ELEMS;
// fp32 side parameters staged in smem: W1[128][6] b1[128] b2[64] b3[128] W4[3][128] b4[4]
constexpr int MLPT_PW1 = 0, MLPT_PB1 = 768, MLPT_PB2 = 896, MLPT_PB3 = 960, MLPT_PW4 = 1088, MLPT_PB4 = 1472, MLPT_PN = 1476;

// shared memory map (bytes)
constexpr int MLPT_S_W2H = 0;
constexpr int MLPT_S_W2L = MLPT_S_W2H + MLPT_W2_ELEMS * 2;
constexpr int MLPT_S_W3H = MLPT_S_W2L + MLPT_W2_ELEMS * 2;
constexpr int MLPT_S_W3L = MLPT_S_W3H + MLPT_W3_ELEMS * 2;
constexpr int MLPT_S_A1H = MLPT_S_W3L + MLPT_W3_ELEMS * 2;       // h1 hi [128 x 128]
constexpr int MLPT_S_A1L = MLPT_S_A1H + MLPT_ROWS * MLP_H1 * 2;
constexpr int MLPT_S_A2H = MLPT_S_A1L + MLPT_ROWS * MLP_H1 * 2;  // h2 hi [128 x 64]
constexpr int MLPT_S_A2L = MLPT_S_A2H + MLPT_ROWS * MLP_H2 * 2;
constexpr int MLPT_S_PAR = MLPT_S_A2L + MLPT_ROWS * MLP_H2 * 2;
constexpr int MLPT_S_BAR = MLPT_S_PAR + MLPT_PN * 4;             // mbarrier (8 B) + tmem base (4 B)
constexpr int MLPT_SMEM = MLPT_S_BAR + 16;

// element offset of (row r, col k) inside a [R x K] K-major no-swizzle operand
__host__ __device__ __forceinline__ int umma_off(int r, int k, int R) { return ((k >> 3) * (R >> 3) + (r >> 3)) * 64 + (r & 7) * 8 + (k & 7); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // SM100 shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
    // layout_type [61,64) = 0 (no swizzle)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// split 8 fp32 values into hi / scaled-lo fp16 and store both 16-byte chunks
__device__ __forceinline__ void split_store8(const float (&h)[8], __half* dst_hi, __half* dst_lo) {
    __half2 hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const __half a = __float2half_rn(h[2 * i]), b = __float2half_rn(h[2 * i + 1]);
        hi[i] = __halves2half2(a, b);
        lo[i] = __halves2half2(__float2half_rn((h[2 * i] - __half2float(a)) * MLPT_LO_SCALE),
                               __float2half_rn((h[2 * i + 1] - __half2float(b)) * MLPT_LO_SCALE));
    }
    *reinterpret_cast<uint4*>(dst_hi) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(dst_lo) = *reinterpret_cast<uint4*>(lo);
}

__global__ void __launch_bounds__(MLPT_THREADS, 1) mlp_tc_kernel(const float* __restrict__ params, const __half* __restrict__ wimg, const MlpIo io) {
    extern __shared__ __align__(1024) unsigned char smt[];
    __half* sW2h = reinterpret_cast<__half*>(smt + MLPT_S_W2H);
    __half* sA1h = reinterpret_cast<__half*>(smt + MLPT_S_A1H);
    __half* sA1l = reinterpret_cast<__half*>(smt + MLPT_S_A1L);
    __half* sA2h = reinterpret_cast<__half*>(smt + MLPT_S_A2H);
    __half* sA2l = reinterpret_cast<__half*>(smt + MLPT_S_A2L);
    float* sPar = reinterpret_cast<float*>(smt + MLPT_S_PAR);
    uint64_t* sBar = reinterpret_cast<uint64_t*>(smt + MLPT_S_BAR);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(smt + MLPT_S_BAR + 8);
    const int t = threadIdx.x, warp = t >> 5;

    // ---- one-time setup: weights -> smem, mbarrier, TMEM ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(wimg);
        uint4* dst = reinterpret_cast<uint4*>(sW2h);
        for (int i = t; i < MLPT_WIMG_ELEMS * 2 / 16; i += MLPT_THREADS) dst[i] = src[i];
        for (int i = t; i < MLP_H1 * MLP_IN; i += MLPT_THREADS) sPar[MLPT_PW1 + i] = params[MLP_OW1 + i];
        for (int i = t; i < MLP_H1; i += MLPT_THREADS) sPar[MLPT_PB1 + i] = params[MLP_OB1 + i];
        for (int i = t; i < MLP_H2; i += MLPT_THREADS) sPar[MLPT_PB2 + i] = params[MLP_OB2 + i];
        for (int i = t; i < MLP_H3; i += MLPT_THREADS) sPar[MLPT_PB3 + i] = params[MLP_OB3 + i];
        for (int i = t; i < MLP_OUT * MLP_H3; i += MLPT_THREADS) sPar[MLPT_PW4 + i] = params[MLP_OW4 + i];
        if (t < 4) sPar[MLPT_PB4 + t] = params[MLP_OB4 + t];
    }
    const uint32_t bar = smem_u32(sBar);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *sTmem;
    const uint32_t tD1m = tmem + 0, tD1c = tmem + 64, tD2m = tmem + 128, tD2c = tmem + 256;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;  // this warp's TMEM lane quarter
    const uint32_t aW2h = smem_u32(smt + MLPT_S_W2H), aW2l = smem_u32(smt + MLPT_S_W2L);
    const uint32_t aW3h = smem_u32(smt + MLPT_S_W3H), aW3l = smem_u32(smt + MLPT_S_W3L);
    const uint32_t aA1h = smem_u32(sA1h), aA1l = smem_u32(sA1l), aA2h = smem_u32(sA2h), aA2l = smem_u32(sA2l);
    constexpr uint32_t ID1 = umma_idesc(128, MLP_H2), ID2 = umma_idesc(128, MLP_H3);
    uint32_t phase = 0;

    const long long n_tiles = (io.M + MLPT_ROWS - 1) / MLPT_ROWS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row = tile * MLPT_ROWS + t;
        float x[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        bool on = false;
        if (row < io.M) on = mlp_fetch_row(io, row, x);
        // ---- layer 1 (CUDA cores, fp32) -> h1 hi/lo operand tiles ----
#pragma unroll 2
        for (int kb = 0; kb < MLP_H1 / 8; kb++) {
            float h[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int n = kb * 8 + i;
                float acc = sPar[MLPT_PB1 + n];
#pragma unroll
                for (int q = 0; q < 6; q++) acc = fmaf(sPar[MLPT_PW1 + n * 6 + q], x[q], acc);
                h[i] = fmaxf(acc, 0.f);
            }
            const int off = umma_off(t, kb * 8, MLPT_ROWS);
            split_store8(h, sA1h + off, sA1l + off);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        // ---- layer 2: D1[128 x 64] = h1[128 x 128] . W2^T  (tcgen05, K = 128 -> 8 k-steps x 3 MMAs) ----
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < MLP_H1 / 16; ks++) {
                const uint32_t ao = ks * 2 * (MLPT_ROWS / 8) * 128;  // two k-chunks per step
                const uint32_t bo = ks * 2 * (MLP_H2 / 8) * 128;
                const uint64_t dAh = umma_desc(aA1h + ao, (MLPT_ROWS / 8) * 128, 128), dAl = umma_desc(aA1l + ao, (MLPT_ROWS / 8) * 128, 128);
                const uint64_t dBh = umma_desc(aW2h + bo, (MLP_H2 / 8) * 128, 128), dBl = umma_desc(aW2l + bo, (MLP_H2 / 8) * 128, 128);
                umma_f16(tD1m, dAh, dBh, ID1, ks > 0);
                umma_f16(tD1c, dAh, dBl, ID1, ks > 0);
                umma_f16(tD1c, dAl, dBh, ID1, 1);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 1: h2 = relu(D1 + b2) -> hi/lo operand tiles ----
#pragma unroll
        for (int c0 = 0; c0 < MLP_H2; c0 += 32) {
            float m[32], cr[32];
            tmem_ld32(tD1m + lane_sel + c0, m);
            tmem_ld32(tD1c + lane_sel + c0, cr);
#pragma unroll
            for (int kb = 0; kb < 4; kb++) {
                float h[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int n = c0 + kb * 8 + i;
                    h[i] = fmaxf(m[kb * 8 + i] + cr[kb * 8 + i] * MLPT_LO_INV + sPar[MLPT_PB2 + n], 0.f);
                }
                const int off = umma_off(t, c0 + kb * 8, MLPT_ROWS);
                split_store8(h, sA2h + off, sA2l + off);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        // ---- layer 3: D2[128 x 128] = h2[128 x 64] . W3^T  (K = 64 -> 4 k-steps x 3 MMAs) ----
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < MLP_H2 / 16; ks++) {
                const uint32_t ao = ks * 2 * (MLPT_ROWS / 8) * 128;
                const uint32_t bo = ks * 2 * (MLP_H3 / 8) * 128;
                const uint64_t dAh = umma_desc(aA2h + ao, (MLPT_ROWS / 8) * 128, 128), dAl = umma_desc(aA2l + ao, (MLPT_ROWS / 8) * 128, 128);
                const uint64_t dBh = umma_desc(aW3h + bo, (MLP_H3 / 8) * 128, 128), dBl = umma_desc(aW3l + bo, (MLP_H3 / 8) * 128, 128);
                umma_f16(tD2m, dAh, dBh, ID2, ks > 0);
                umma_f16(tD2c, dAh, dBl, ID2, ks > 0);
                umma_f16(tD2c, dAl, dBh, ID2, 1);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 2: h3 = relu(D2 + b3); layer 4 (CUDA cores, fp32): out = W4 h3 + b4 ----
        float o0 = sPar[MLPT_PB4 + 0], o1 = sPar[MLPT_PB4 + 1], o2 = sPar[MLPT_PB4 + 2];
#pragma unroll
        for (int c0 = 0; c0 < MLP_H3; c0 += 32) {
            float m[32], cr[32];
            tmem_ld32(tD2m + lane_sel + c0, m);
            tmem_ld32(tD2c + lane_sel + c0, cr);
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const int n = c0 + i;
                const float h = fmaxf(m[i] + cr[i] * MLPT_LO_INV + sPar[MLPT_PB3 + n], 0.f);
                o0 = fmaf(sPar[MLPT_PW4 + n], h, o0);
                o1 = fmaf(sPar[MLPT_PW4 + MLP_H3 + n], h, o1);
                o2 = fmaf(sPar[MLPT_PW4 + 2 * MLP_H3 + n], h, o2);
            }
        }
        if (row < io.M) mlp_store_row(io, row, on, o0, o1, o2);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// host: build the fp16 hi/lo operand images of W2, W3 in UMMA layout and upload them
inline int mlp_tc_prepare(const float* host_params, void** out) {
    __half* img = new __half[MLPT_WIMG_ELEMS];
    auto put = [&](const float* W, int R, int K, __half* hi, __half* lo) {
        for (int r = 0; r < R; r++)
            for (int k = 0; k < K; k++) {
                const float w = W[r * K + k];
                const __half h = __float2half_rn(w);
                hi[umma_off(r, k, R)] = h;
                lo[umma_off(r, k, R)] = __float2half_rn((w - __half2float(h)) * MLPT_LO_SCALE);
            }
    };
    put(host_params + MLP_OW2, MLP_H2, MLP_H1, img, img + MLPT_W2_ELEMS);
    put(host_params + MLP_OW3, MLP_H3, MLP_H2, img + 2 * MLPT_W2_ELEMS, img + 2 * MLPT_W2_ELEMS + MLPT_W3_ELEMS);
    cudaError_t e = cudaMalloc(out, sizeof(__half) * MLPT_WIMG_ELEMS);
    if (e == cudaSuccess) e = cudaMemcpy(*out, img, sizeof(__half) * MLPT_WIMG_ELEMS, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MLPT_SMEM);
    delete[] img;
    return (int)e;
}

inline int mlp_tc_launch(const float* params, const void* wimg, const MlpIo& io, int n_sm, cudaStream_t st) {
    const long long tiles = (io.M + MLPT_ROWS - 1) / MLPT_ROWS;
    const int grd = (int)(tiles < n_sm ? tiles : n_sm);
    mlp_tc_kernel<<<grd, MLPT_THREADS, MLPT_SMEM, st>>>(params, reinterpret_cast<const __half*>(wimg), io);
    return (int)cudaGetLastError();
}

}  // namespace ndp
