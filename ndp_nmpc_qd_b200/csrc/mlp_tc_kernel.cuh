// Downwash MLP on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// The two wide layers (128->64 and 64->128, 93 % of the flops) run as tcgen05.mma kind::f16 with
// fp32 accumulators in tensor memory; the K=6 input layer and the N=3 output layer stay on the
// CUDA cores in fp32 and are fused around the MMAs, as are the relative-feature construction
// (other - ego)[0:6] and the 1 m gate (downwash_nn.py:21-29, ndp_nmpc_leader_node.py:65-76).
//
// Precision: the reference runs this net in fp32 (cuBLAS SGEMM).  Plain bf16/fp16 operands would
// miss the u0 parity target, so every operand is split into two fp16 halves
//     a = a_hi + a_lo,   a_hi = fp16(a),  a_lo = fp16(a - a_hi)   (a_lo ~ 2^-11 |a|: an fp16 subnormal below |a| ~ 0.12,
//                                                                  absolute error <= 3e-8 there)
// and the product is accumulated into ONE fp32 accumulator,  D += a_hi b_hi + a_hi b_lo + a_lo b_hi  (dropped term
// a_lo b_lo ~ 2^-22): three MMAs per k-step, ~fp32 accuracy with 16-bit operands (max error 7.7e-6 N against fp64 on
// the shipped weights; plain fp32 is at 5.9e-6).  Valid for |feature| < ~1e3 (fp16 range of the activations).
// An earlier version scaled the lo halves by 2^11 and kept main / correction accumulators apart: the epilogues then
// read twice the tensor memory, and TMEM read bandwidth (~100 B/clk/SM measured) is what bounds this kernel --
// cutting 25 % of the CUDA-core instructions changed nothing, halving the TMEM reads did.
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
constexpr int MLPT_THREADS = 256;  // per 128-row tile: two threads per row (each owns half of the columns)

// fp16 operand images (elements): [W2 hi; W2 lo] [128 x 128], [W3 hi; W3 lo] [256 x 64]
constexpr int MLPT_W2_ELEMS = MLP_H2 * MLP_H1;
constexpr int MLPT_W3_ELEMS = MLP_H3 * MLP_H2;
constexpr int MLPT_K1 = 16;                        // layer-1 K padded to one MMA k-step: 6 features, a constant 1 (bias column), zeros
constexpr int MLPT_W1_ELEMS = MLP_H1 * MLPT_K1;    // [W1 | b1 | 0] per half: [128 x 16]
constexpr int MLPT_WIMG_ELEMS = 2 * MLPT_W2_ELEMS + 2 * MLPT_W3_ELEMS + 2 * MLPT_W1_ELEMS;
// fp32 side parameters (layers 1 and 4, biases): a small device buffer, staged once per CTA in shared memory and
// read with broadcast 128-bit loads (indexed constant-bank loads measured 4x slower)
struct alignas(16) MlpSmall {
    float W1[MLP_H1 * MLP_IN], b1[MLP_H1], b2[MLP_H2], b3[MLP_H3], W4[MLP_OUT * MLP_H3], b4[4];
};

// shared memory map (bytes): weights once per CTA, one activation region per 128-thread group
// (h2 hi/lo alias the front of the h1 region: h1 is dead once the layer-2 MMAs have committed)
constexpr int MLPT_GROUPS = 2;
constexpr int MLPT_TMEM_COLS = 512;  // per group: a 128-column fp32 accumulator (reused by the three layers) + the fp16 activation
                                     // operand of the next layer, hi | lo, 64 columns each (two K elements per 32-bit column)
constexpr int MLPT_CTA_THREADS = MLPT_GROUPS * MLPT_THREADS;
constexpr int MLPT_S_W2H = 0;
constexpr int MLPT_S_W2L = MLPT_S_W2H + MLPT_W2_ELEMS * 2;
constexpr int MLPT_S_W3H = MLPT_S_W2L + MLPT_W2_ELEMS * 2;
constexpr int MLPT_S_W3L = MLPT_S_W3H + MLPT_W3_ELEMS * 2;
constexpr int MLPT_S_W1H = MLPT_S_W3L + MLPT_W3_ELEMS * 2;   // [W1_hi; W1_lo]: 256 rows x 16
constexpr int MLPT_S_ACT = MLPT_S_W1H + 2 * MLPT_W1_ELEMS * 2;
constexpr int MLPT_ACT_BYTES = 2 * MLPT_ROWS * MLPT_K1 * 2;      // per group: the feature tiles [128 x 16] fp16, hi + lo
constexpr int MLPT_A0H = 0, MLPT_A0L = MLPT_ROWS * MLPT_K1 * 2;
constexpr int MLPT_S_PAR = MLPT_S_ACT + MLPT_GROUPS * MLPT_ACT_BYTES;  // fp32 side parameters (MlpSmall image)
constexpr int MLPT_S_OUT = MLPT_S_PAR + (int)sizeof(MlpSmall);         // layer-4 partial sums of the upper column half [groups][128][4] fp32
constexpr int MLPT_S_BAR = MLPT_S_OUT + MLPT_GROUPS * MLPT_ROWS * 16;  // 3 mbarriers (24 B) + tmem base (4 B)
constexpr int MLPT_SMEM = MLPT_S_BAR + 48;

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
// same with the A operand in tensor memory (row m of A in lane m, two K elements per 32-bit column): the activations never
// pass through shared memory, and the tensor core reads only the weights from it
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 columns of this thread's TMEM lane <- 16 registers (32 fp16 values)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// packed fp32 pairs (sm_100 add / sub / fma .f32x2): one issue slot for two lanes of arithmetic -- the epilogues of this
// kernel are bound by issue slots once the activations no longer travel through shared memory
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void sub2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// split 32 fp32 values (after ReLU and an optional bias) into fp16 hi / lo pairs, packed two per register
__device__ __forceinline__ void split32(const float (&m)[32], int first, const float* bias, uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float a = m[2 * i], b = m[2 * i + 1];
        if (bias) {
            const float2 bb = *reinterpret_cast<const float2*>(bias + first + 2 * i);
            add2(a, b, a, b, bb.x, bb.y);
        }
        a = fmaxf(a, 0.f); b = fmaxf(b, 0.f);
        const __half2 h = __floats2half2_rn(a, b);
        const float2 f = __half22float2(h);
        float la, lb;
        sub2(la, lb, a, b, f.x, f.y);
        const __half2 l = __floats2half2_rn(la, lb);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
}
// one elected lane of a converged warp (lets ptxas keep the MMA operands on the uniform datapath)
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}\n"
        : "+r"(laneid), "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred;
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

// split 8 fp32 values into hi / lo fp16 (lo = the rounding residual, unscaled) and store both 16-byte chunks
__device__ __forceinline__ void split_store8(const float (&h)[8], __half* dst_hi, __half* dst_lo) {
    __half2 hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        hi[i] = __floats2half2_rn(h[2 * i], h[2 * i + 1]);
        const float2 f = __half22float2(hi[i]);
        lo[i] = __floats2half2_rn(h[2 * i] - f.x, h[2 * i + 1] - f.y);
    }
    *reinterpret_cast<uint4*>(dst_hi) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(dst_lo) = *reinterpret_cast<uint4*>(lo);
}

__device__ long long g_mlpt_prof[128];
#define MLPT_STAMP(i) do { const int i_ = (i); if (io.prof && blockIdx.x == 0 && threadIdx.x == 0 && i_ < 128) g_mlpt_prof[i_] = clock64(); } while (0)

__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(MLPT_THREADS) : "memory"); }

// CTA = 2 groups of 256 threads; each group owns a 128-row tile, its own activation tiles, mbarrier and
// 256 TMEM columns, so one group's CUDA-core phases overlap the other's MMAs and memory latency.  Inside
// a group, threads t and t + 128 share row t (= TMEM lane t: warps w and w + 4 address the same lane
// quarter) and each owns one half of the columns of every layer.  Features of the next tile are
// prefetched while the current one computes.
__global__ void __launch_bounds__(MLPT_CTA_THREADS, 1) mlp_tc_kernel(const MlpSmall* __restrict__ gsp, const __half* __restrict__ wimg, const MlpIo io) {
    extern __shared__ __align__(1024) unsigned char smt[];
    uint64_t* sBar = reinterpret_cast<uint64_t*>(smt + MLPT_S_BAR);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(smt + MLPT_S_BAR + 32);
    const int grp = threadIdx.x / MLPT_THREADS, tg = threadIdx.x % MLPT_THREADS, t = tg & 127, hf = tg >> 7, warp = tg >> 5;
    float4* sOut = reinterpret_cast<float4*>(smt + MLPT_S_OUT) + grp * MLPT_ROWS;
    unsigned char* act = smt + MLPT_S_ACT + grp * MLPT_ACT_BYTES;
    __half* sA0h = reinterpret_cast<__half*>(act + MLPT_A0H);
    __half* sA0l = reinterpret_cast<__half*>(act + MLPT_A0L);
    const MlpSmall& sq = *reinterpret_cast<const MlpSmall*>(smt + MLPT_S_PAR);
    int stamp = 0;
    MLPT_STAMP(stamp++);  // kernel entry
    {
        // side parameters: one coalesced 128-bit load per thread from the device copy (staging them from the kernel's
        // constant-bank argument took 9 % of the kernel: every lane of a warp read a different constant address)
        static_assert(sizeof(MlpSmall) % 16 == 0, "MlpSmall is copied as float4");
        float4* dstp = reinterpret_cast<float4*>(smt + MLPT_S_PAR);
        const float4* srcp = reinterpret_cast<const float4*>(gsp);
        for (int i = threadIdx.x; i < (int)(sizeof(MlpSmall) / 16); i += MLPT_CTA_THREADS) dstp[i] = srcp[i];
    }

    // ---- one-time setup: mbarriers, weights -> smem by TMA bulk copy (lands under the first tile's
    //      layer-1 compute; only the tensor core reads them), TMEM ----
    const uint32_t bar = smem_u32(sBar + grp), wbar = smem_u32(sBar + 2);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(sBar)), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(sBar + 1)), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wbar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        constexpr uint32_t kBytes = MLPT_WIMG_ELEMS * 2, kChunk = 8192;
        static_assert(kBytes % kChunk == 0, "weight image is copied in whole chunks");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wbar), "r"(kBytes) : "memory");
#pragma unroll
        for (uint32_t o = 0; o < kBytes; o += kChunk)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(smt + MLPT_S_W2H) + o),
                         "l"(reinterpret_cast<const unsigned char*>(wimg) + o), "r"(kChunk), "r"(wbar)
                         : "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(MLPT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tD = *sTmem + grp * (MLPT_TMEM_COLS / MLPT_GROUPS);  // the accumulator of every layer (each is dead before the next is issued)
    const uint32_t tAh = tD + 128, tAl = tD + 192;  // the next layer's A operand in tensor memory: fp16 hi | lo, element k in column k / 2
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;  // this warp's TMEM lane quarter
    const uint32_t aW2h = smem_u32(smt + MLPT_S_W2H), aW3h = smem_u32(smt + MLPT_S_W3H), aW1h = smem_u32(smt + MLPT_S_W1H);
    const uint32_t aA0h = smem_u32(sA0h), aA0l = smem_u32(sA0l);
    constexpr uint32_t ID0 = umma_idesc(128, MLP_H1), ID1 = umma_idesc(128, MLP_H2), ID2 = umma_idesc(128, MLP_H3);
    // weight images are [W_hi; W_lo] (2R rows) per k-chunk: the lo rows start R / 8 row groups into each chunk
    constexpr uint32_t LO1 = (MLP_H1 / 8) * 128, LO2 = (MLP_H2 / 8) * 128, LO3 = (MLP_H3 / 8) * 128;
    uint32_t phase = 0;
    // This kernel is launched as a programmatic dependent of whatever precedes it in the stream: everything above (barriers,
    // weight copy, tensor-memory allocation -- nothing another kernel writes) has run under that kernel's tail; its
    // results (features, pair lists, the row count) are read from here on.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const long long M = mlp_rows(io);
    // Rows are dealt to the (CTA, group) pairs in units of 32 (one warp's rows), as evenly as they go: every group gets
    // floor or ceil of (row-warps / groups) of them as one contiguous range, walked in 128-row tiles.  The last tile of a
    // range is then partial everywhere (its idle warps skip the CUDA-core phases) instead of a full extra round on a
    // fraction of the SMs: 86 016 rows on 296 groups are 128 + 128 + 32 (or 64) rows each, not three full tiles on 80
    // groups and two on the rest.  Group g of CTA c has index g * gridDim.x + c, so the longer ranges land on different SMs.
    const long long n_w = (M + 31) / 32, n_grp = (long long)gridDim.x * MLPT_GROUPS;
    const long long gidx = (long long)grp * gridDim.x + blockIdx.x;
    const long long w_q = n_w / n_grp, w_r = n_w % n_grp;
    const long long row_begin = (gidx * w_q + (gidx < w_r ? gidx : w_r)) * 32;
    const long long row_end_raw = row_begin + (w_q + (gidx < w_r ? 1 : 0)) * 32;
    const long long row_end = row_end_raw < M ? row_end_raw : M;
    // (device-sized launches: a CTA may find no rows for either of its groups; it still waits for its weight copy and
    // frees its tensor memory below)
    const long long n_tiles = row_end > row_begin ? (row_end - row_begin + MLPT_ROWS - 1) / MLPT_ROWS : 0;
    MlpRaw raw;  // loaded, not yet used: the features are formed one tile later (mlp_fetch_finish)
#pragma unroll
    for (int i = 0; i < 6; i++) { raw.o[i] = 0.f; raw.e[i] = 0.f; }
    raw.gated = false; raw.on = false; raw.ox = raw.oy = raw.gx = raw.gy = 0.f;
    if (row_begin + t < row_end) mlp_fetch_issue(io, row_begin + t, raw);
    // a solve launched as a programmatic dependent (NDP_UPDATE_F_FROM_PREVIOUS_KERNEL) may be scheduled as CTAs of this
    // grid retire; it synchronises on this grid's completion itself before it reads the forces.  Issued only now, after
    // this grid's own dependency wait: the dependent's prologue reads data (the iterate) that the kernels before this one
    // may still have been writing.
    asm volatile("griddepcontrol.launch_dependents;");

    MLPT_STAMP(stamp++);  // setup done
    for (long long tile = 0; tile < n_tiles; tile++) {
        const long long row = row_begin + tile * MLPT_ROWS + t;
        // this warp's 32 rows lie inside the range (uniform per warp: ranges are multiples of 32 rows but for the very last)
        const bool live = row_begin + tile * MLPT_ROWS + (t & ~31) < row_end;
        float x[6];
        const bool on = mlp_fetch_finish(io, raw, x);
        {   // prefetch the next tile's raw rows; the loads complete under this tile's compute
            const long long nrow = row + MLPT_ROWS;
#pragma unroll
            for (int i = 0; i < 6; i++) { raw.o[i] = 0.f; raw.e[i] = 0.f; }
            raw.gated = false; raw.on = false;
            if (nrow < row_end) mlp_fetch_issue(io, nrow, raw);
        }
        MLPT_STAMP(stamp++);
        // ---- layer 1 on the tensor cores too: D0[128 x 128] = [x | 1 | 0][128 x 16] . [W1 | b1 | 0]^T (one k-step, the same
        //      hi / lo split; the bias rides on the constant-1 column).  On the CUDA cores this layer was 45 % of the
        //      kernel's instructions (768 FFMA + 112 LDS.128 per row). ----
        if (live && !hf) {
            float v[8] = {x[0], x[1], x[2], x[3], x[4], x[5], 1.f, 0.f};
            const int off = umma_off(t, 0, MLPT_ROWS);
            split_store8(v, sA0h + off, sA0l + off);
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sA0h + umma_off(t, 8, MLPT_ROWS)) = z;   // k = 8..15: the MMA's k-step is 16 wide
            *reinterpret_cast<uint4*>(sA0l + umma_off(t, 8, MLPT_ROWS)) = z;
        }
        MLPT_STAMP(stamp++);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        group_sync(grp);
        if (warp == 0) {
          if (elect_one_sync()) {
            mbar_wait(wbar, 0);  // weights have landed (returns immediately after the first tile)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t dAh = umma_desc(aA0h, (MLPT_ROWS / 8) * 128, 128), dAl = umma_desc(aA0l, (MLPT_ROWS / 8) * 128, 128);
            const uint64_t dBh = umma_desc(aW1h, (2 * MLP_H1 / 8) * 128, 128), dBl = umma_desc(aW1h + LO1, (2 * MLP_H1 / 8) * 128, 128);
            // the two small correction products first, the main product last: the tensor core aligns every addend to
            // the accumulator's exponent, so the corrections are summed while the accumulator is still small
            umma_f16(tD, dAh, dBl, ID0, 0);   // x_hi . W1_lo
            umma_f16(tD, dAl, dBh, ID0, 1);   // + x_lo . W1_hi
            umma_f16(tD, dAh, dBh, ID0, 1);   // + x_hi . W1_hi
            umma_commit(bar);
          }
          __syncwarp();
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 0: h1 = relu(D0) -> fp16 hi / lo, straight back into tensor memory as layer 2's A operand ----
#pragma unroll 1
        for (int c0 = hf * 64; live && c0 < hf * 64 + 64; c0 += 32) {
            float m[32];
            tmem_ld32(tD + lane_sel + c0, m);
            uint32_t hi[16], lo[16];
            split32(m, 0, nullptr, hi, lo);
            tmem_st16(tAh + lane_sel + c0 / 2, hi);
            tmem_st16(tAl + lane_sel + c0 / 2, lo);
        }
        tmem_st_wait();
        MLPT_STAMP(stamp++);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        group_sync(grp);
        MLPT_STAMP(stamp++);
        // ---- layer 2: D1[128 x 64] = h1[128 x 128] . W2^T  (tcgen05, K = 128 -> 8 k-steps x 3 MMAs) ----
        if (warp == 0) {
          if (elect_one_sync()) {
            mbar_wait(wbar, 0);  // weights have landed (returns immediately after the first tile)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < MLP_H1 / 16; ks++) {
                const uint32_t bo = ks * 2 * (2 * MLP_H2 / 8) * 128;   // two k-chunks of the weight image per step; 8 columns of A
                const uint64_t dBh = umma_desc(aW2h + bo, (2 * MLP_H2 / 8) * 128, 128), dBl = umma_desc(aW2h + bo + LO2, (2 * MLP_H2 / 8) * 128, 128);
                umma_f16_ts(tD, tAh + ks * 8, dBl, ID1, ks > 0);   // corrections of every k-step first (see layer 1)
                umma_f16_ts(tD, tAl + ks * 8, dBh, ID1, 1);
            }
#pragma unroll
            for (int ks = 0; ks < MLP_H1 / 16; ks++) {
                const uint32_t bo = ks * 2 * (2 * MLP_H2 / 8) * 128;
                umma_f16_ts(tD, tAh + ks * 8, umma_desc(aW2h + bo, (2 * MLP_H2 / 8) * 128, 128), ID1, 1);
            }
            umma_commit(bar);
          }
          __syncwarp();
        }
        MLPT_STAMP(stamp++);
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        MLPT_STAMP(stamp++);
        // ---- epilogue 1: h2 = relu(D1 + b2) -> fp16 hi / lo into tensor memory (over the dead h1 operand) ----
        if (live) {
            const int c0 = hf * 32;
            float m[32];
            tmem_ld32(tD + lane_sel + c0, m);
            uint32_t hi[16], lo[16];
            split32(m, c0, sq.b2, hi, lo);
            tmem_st16(tAh + lane_sel + c0 / 2, hi);
            tmem_st16(tAl + lane_sel + c0 / 2, lo);
        }
        tmem_st_wait();
        MLPT_STAMP(stamp++);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        group_sync(grp);
        MLPT_STAMP(stamp++);
        // ---- layer 3: D2[128 x 128] = h2[128 x 64] . W3^T  (K = 64 -> 4 k-steps x 3 MMAs) ----
        if (warp == 0) {
          if (elect_one_sync()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < MLP_H2 / 16; ks++) {
                const uint32_t bo = ks * 2 * (2 * MLP_H3 / 8) * 128;
                const uint64_t dBh = umma_desc(aW3h + bo, (2 * MLP_H3 / 8) * 128, 128), dBl = umma_desc(aW3h + bo + LO3, (2 * MLP_H3 / 8) * 128, 128);
                umma_f16_ts(tD, tAh + ks * 8, dBl, ID2, ks > 0);
                umma_f16_ts(tD, tAl + ks * 8, dBh, ID2, 1);
            }
#pragma unroll
            for (int ks = 0; ks < MLP_H2 / 16; ks++) {
                const uint32_t bo = ks * 2 * (2 * MLP_H3 / 8) * 128;
                umma_f16_ts(tD, tAh + ks * 8, umma_desc(aW3h + bo, (2 * MLP_H3 / 8) * 128, 128), ID2, 1);
            }
            umma_commit(bar);
          }
          __syncwarp();
        }
        MLPT_STAMP(stamp++);
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        MLPT_STAMP(stamp++);
        // ---- epilogue 2: h3 = relu(D2 + b3); layer 4 (CUDA cores, fp32): out = W4 h3 + b4 ----
        // two partial sums per output (even / odd elements): six independent FMA chains
        float o0 = hf ? 0.f : sq.b4[0], o1 = hf ? 0.f : sq.b4[1], o2 = hf ? 0.f : sq.b4[2];
        float q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll 1
        for (int c0 = hf * 64; live && c0 < hf * 64 + 64; c0 += 32) {
            float m[32];
            tmem_ld32(tD + lane_sel + c0, m);
#pragma unroll
            for (int i4 = 0; i4 < 8; i4++) {
                const float4 b = *reinterpret_cast<const float4*>(sq.b3 + c0 + 4 * i4);
                const float4 w0 = *reinterpret_cast<const float4*>(sq.W4 + c0 + 4 * i4);
                const float4 w1 = *reinterpret_cast<const float4*>(sq.W4 + MLP_H3 + c0 + 4 * i4);
                const float4 w2 = *reinterpret_cast<const float4*>(sq.W4 + 2 * MLP_H3 + c0 + 4 * i4);
                float h0, h1, h2, h3;
                add2(h0, h1, m[4 * i4 + 0], m[4 * i4 + 1], b.x, b.y);
                add2(h2, h3, m[4 * i4 + 2], m[4 * i4 + 3], b.z, b.w);
                h0 = fmaxf(h0, 0.f); h1 = fmaxf(h1, 0.f); h2 = fmaxf(h2, 0.f); h3 = fmaxf(h3, 0.f);
                // (o, q) = the partial sums over the even / odd columns: one packed FMA per output and column pair
                ffma2(o0, q0, w0.x, w0.y, h0, h1); ffma2(o1, q1, w1.x, w1.y, h0, h1); ffma2(o2, q2, w2.x, w2.y, h0, h1);
                ffma2(o0, q0, w0.z, w0.w, h2, h3); ffma2(o1, q1, w1.z, w1.w, h2, h3); ffma2(o2, q2, w2.z, w2.w, h2, h3);
            }
        }
        o0 += q0; o1 += q1; o2 += q2;
        MLPT_STAMP(stamp++);
        if (hf) sOut[t] = make_float4(o0, o1, o2, 0.f);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync(grp);
        if (!hf && row < row_end) {
            const float4 p = sOut[t];
            mlp_store_row(io, row, on, o0 + p.x, o1 + p.y, o2 + p.z);
        }
        MLPT_STAMP(stamp++);
    }
    MLPT_STAMP(stamp++);  // this group's tiles done
    if (threadIdx.x == 0) mbar_wait(wbar, 0);  // the bulk copy into this CTA's shared memory must not outlive it
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    MLPT_STAMP(stamp++);  // both groups done
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*sTmem), "r"(MLPT_TMEM_COLS) : "memory");
}

// host: build the fp16 hi/lo operand images of W2, W3 in UMMA layout and upload them
inline int mlp_tc_prepare(const float* host_params, void** out) {
    __half* img = new __half[MLPT_WIMG_ELEMS];
    // each image is the concatenation [W_hi; W_lo] (2R rows) in UMMA K-major layout
    auto put = [&](const float* W, int R, int K, __half* dst) {
        for (int r = 0; r < R; r++)
            for (int k = 0; k < K; k++) {
                const float w = W[r * K + k];
                const __half h = __float2half_rn(w);
                dst[umma_off(r, k, 2 * R)] = h;
                dst[umma_off(R + r, k, 2 * R)] = __float2half_rn(w - __half2float(h));
            }
    };
    put(host_params + MLP_OW2, MLP_H2, MLP_H1, img);
    put(host_params + MLP_OW3, MLP_H3, MLP_H2, img + 2 * MLPT_W2_ELEMS);
    {   // layer 1: [W1 | b1 | 0] with K padded to 16 (the kernel feeds [x | 1 | 0])
        float* w1 = new float[MLP_H1 * MLPT_K1]();
        for (int r = 0; r < MLP_H1; r++) {
            for (int k = 0; k < MLP_IN; k++) w1[r * MLPT_K1 + k] = host_params[MLP_OW1 + r * MLP_IN + k];
            w1[r * MLPT_K1 + MLP_IN] = host_params[MLP_OB1 + r];
        }
        put(w1, MLP_H1, MLPT_K1, img + 2 * MLPT_W2_ELEMS + 2 * MLPT_W3_ELEMS);
        delete[] w1;
    }
    cudaError_t e = cudaMalloc(out, sizeof(__half) * MLPT_WIMG_ELEMS);
    if (e == cudaSuccess) e = cudaMemcpy(*out, img, sizeof(__half) * MLPT_WIMG_ELEMS, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MLPT_SMEM);
    delete[] img;
    return (int)e;
}

inline void mlp_tc_small(const float* host_params, MlpSmall* sp) {
    for (int i = 0; i < MLP_H1 * MLP_IN; i++) sp->W1[i] = host_params[MLP_OW1 + i];
    for (int i = 0; i < MLP_H1; i++) sp->b1[i] = host_params[MLP_OB1 + i];
    for (int i = 0; i < MLP_H2; i++) sp->b2[i] = host_params[MLP_OB2 + i];
    for (int i = 0; i < MLP_H3; i++) sp->b3[i] = host_params[MLP_OB3 + i];
    for (int i = 0; i < MLP_OUT * MLP_H3; i++) sp->W4[i] = host_params[MLP_OW4 + i];
    for (int i = 0; i < 4; i++) sp->b4[i] = i < MLP_OUT ? host_params[MLP_OB4 + i] : 0.f;
}

inline int mlp_tc_launch(const MlpSmall* sp_dev, const void* wimg, const MlpIo& io, int n_sm, cudaStream_t st) {
    const long long tiles = (io.M + MLPT_ROWS - 1) / MLPT_ROWS;  // io.M is the row capacity when the count lives on the device
    const long long ctas = (tiles + MLPT_GROUPS - 1) / MLPT_GROUPS;
    const int grd = (int)(ctas < n_sm ? ctas : n_sm);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grd);
    cfg.blockDim = dim3(MLPT_CTA_THREADS);
    cfg.dynamicSmemBytes = MLPT_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see griddepcontrol.wait in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, mlp_tc_kernel, sp_dev, reinterpret_cast<const __half*>(wimg), io);
    return (int)cudaGetLastError();
}

}  // namespace ndp
