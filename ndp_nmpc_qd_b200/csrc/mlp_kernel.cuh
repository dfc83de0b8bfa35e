// Downwash MLP 6-128-64-128-3 (ReLU), fused with relative-feature construction and gating.
// Reference: dnwash_nn_est/nn_net.py:7-18 (architecture), downwash_nn.py:21-29 (features =
// (other - ego)[:, 0:6] cast to fp32), ndp_nmpc_leader_node.py:60-76 (1 m horizontal gate).
//
// This file holds the CUDA-core fp32 path (exact fp32 FMA arithmetic; used for small row counts
// such as the batch-1 drop-in call and as the numerics anchor for the tensor-core path in
// mlp_tc_kernel.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ndp {

constexpr int MLP_IN = 6, MLP_H1 = 128, MLP_H2 = 64, MLP_H3 = 128, MLP_OUT = 3;
// packed fp32 parameter block (floats): W1[128][6] b1[128] W2[64][128] b2[64] W3[128][64] b3[128] W4[3][128] b4[3](+1 pad)
constexpr int MLP_OW1 = 0;
constexpr int MLP_OB1 = MLP_OW1 + MLP_H1 * MLP_IN;
constexpr int MLP_OW2 = MLP_OB1 + MLP_H1;
constexpr int MLP_OB2 = MLP_OW2 + MLP_H2 * MLP_H1;
constexpr int MLP_OW3 = MLP_OB2 + MLP_H2;
constexpr int MLP_OB3 = MLP_OW3 + MLP_H3 * MLP_H2;
constexpr int MLP_OW4 = MLP_OB3 + MLP_H3;
constexpr int MLP_OB4 = MLP_OW4 + MLP_OUT * MLP_H3;
constexpr int MLP_NPARAM = MLP_OB4 + 4;  // 17 860 floats

// The swarm's trajectory tensor traj[n_all][n_nodes][6] fp32, possibly split in equal contiguous parts that
// live on different GPUs: p[r] is the base of rows [r * part_rows, (r + 1) * part_rows) -- local memory or a
// peer GPU's buffer mapped over NVLink (symmetric memory), read directly by the kernels below.
constexpr int MLP_MAX_PARTS = 16;
struct TrajParts {
    const float* p[MLP_MAX_PARTS];
    int n_parts, part_rows;
    __device__ __forceinline__ const float* row(int j, int n_nodes) const {
        const int r = j / part_rows;
        return p[r] + (long long)(j - r * part_rows) * n_nodes * 6;
    }
};

// How rows are produced / consumed.
struct MlpIo {
    // mode 0: rows given directly, in[M][6] fp32
    // mode 1: pairs, ego/other [P][n_nodes][10] (float or double), optional gate
    // mode 2: pair list over a shared trajectory tensor traj[n_all][n_nodes][6] fp32
    int mode;
    int precision;  // element type of ego/other/out for mode 1, of out for mode 2 (0 f32, 1 f64)
    int n_nodes;
    int accumulate;
    int other_ld;  // mode 1: row stride of `other` (10 = full state rows, 6 = position+velocity only)
    long long M;  // total rows (host-known), or the row capacity when m_dev is set
    const int* m_dev;  // optional: device-side count of n_nodes-row groups (pairs); rows = min(*m_dev * n_nodes, M)
    const float* in;
    const void* ego;
    const void* other;
    const void* gate_xy;
    double r2;
    TrajParts tp;
    const int2* pairs;  // (ego_global, other_global)
    int prof;           // debug: record phase timestamps of CTA 0 (mlp_tc_kernel)
    void* out;          // mode 0: float [M][3]; mode 1: precision [P][n_nodes][3]; mode 2: float [n_pairs][n_nodes][3]
};

// number of rows of this launch
__device__ __forceinline__ long long mlp_rows(const MlpIo& io) {
    if (!io.m_dev) return io.M;
    const long long r = (long long)(*io.m_dev) * io.n_nodes;
    return r < io.M ? r : io.M;
}

// feature row + gate for row index `row`
__device__ __forceinline__ bool mlp_fetch_row(const MlpIo& io, long long row, float (&x)[6]) {
    if (io.mode == 0) {
#pragma unroll
        for (int i = 0; i < 6; i++) x[i] = io.in[row * 6 + i];
        return true;
    } else if (io.mode == 1) {
        const long long p = row / io.n_nodes;
        bool on = true;
        if (io.precision == 1) {
            const double* e = (const double*)io.ego + row * 10;
            const double* o = (const double*)io.other + row * io.other_ld;
#pragma unroll
            for (int i = 0; i < 6; i++) x[i] = (float)(o[i] - e[i]);
            if (io.gate_xy) {
                const double* o0 = (const double*)io.other + p * io.n_nodes * io.other_ld;
                const double* g = (const double*)io.gate_xy + p * 2;
                const double dx = o0[0] - g[0], dy = o0[1] - g[1];
                on = dx * dx + dy * dy < io.r2;
            }
        } else {
            const float* e = (const float*)io.ego + row * 10;
            const float* o = (const float*)io.other + row * io.other_ld;
#pragma unroll
            for (int i = 0; i < 6; i++) x[i] = o[i] - e[i];
            if (io.gate_xy) {
                const float* o0 = (const float*)io.other + p * io.n_nodes * io.other_ld;
                const float* g = (const float*)io.gate_xy + p * 2;
                const float dx = o0[0] - g[0], dy = o0[1] - g[1];
                on = dx * dx + dy * dy < (float)io.r2;
            }
        }
        return on;
    } else {
        const long long p = row / io.n_nodes;
        const int k = (int)(row - p * io.n_nodes);
        const int2 pr = io.pairs[p];
        const float* e = io.tp.row(pr.x, io.n_nodes) + k * 6;
        const float* o = io.tp.row(pr.y, io.n_nodes) + k * 6;
#pragma unroll
        for (int i = 0; i < 6; i++) x[i] = o[i] - e[i];
        return true;
    }
}

// Two-phase form of mlp_fetch_row for software prefetching: issue() only LOADS (nothing depends on the values, so an
// in-order warp does not wait for them), finish() forms the features / gate one tile later.
struct MlpRaw {
    float o[6], e[6], ox, oy, gx, gy;
    bool gated, on;
};
__device__ __forceinline__ void mlp_fetch_issue(const MlpIo& io, long long row, MlpRaw& r) {
    r.gated = false; r.on = true;
    r.ox = r.oy = r.gx = r.gy = 0.f;
    if (io.mode == 1 && io.precision == 0) {
        const long long p = row / io.n_nodes;
        const float* e = (const float*)io.ego + row * 10;
        const float* o = (const float*)io.other + row * io.other_ld;
#pragma unroll
        for (int i = 0; i < 6; i++) { r.o[i] = o[i]; r.e[i] = e[i]; }
        if (io.gate_xy) {
            const float* o0 = (const float*)io.other + p * io.n_nodes * io.other_ld;
            const float* g = (const float*)io.gate_xy + p * 2;
            r.ox = o0[0]; r.oy = o0[1]; r.gx = g[0]; r.gy = g[1];
            r.gated = true;
        }
    } else if (io.mode == 2) {
        const long long p = row / io.n_nodes;
        const int k = (int)(row - p * io.n_nodes);
        const int2 pr = io.pairs[p];
        const float* e = io.tp.row(pr.x, io.n_nodes) + k * 6;
        const float* o = io.tp.row(pr.y, io.n_nodes) + k * 6;
#pragma unroll
        for (int i = 0; i < 6; i++) { r.o[i] = o[i]; r.e[i] = e[i]; }
    } else {
        // plain rows / fp64 pairs: formed at once (not the throughput paths)
        float x[6];
        r.on = mlp_fetch_row(io, row, x);
#pragma unroll
        for (int i = 0; i < 6; i++) { r.o[i] = x[i]; r.e[i] = 0.f; }
    }
}
__device__ __forceinline__ bool mlp_fetch_finish(const MlpIo& io, const MlpRaw& r, float (&x)[6]) {
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = r.o[i] - r.e[i];
    bool on = r.on;
    if (r.gated) {
        const float dx = r.ox - r.gx, dy = r.oy - r.gy;
        on = dx * dx + dy * dy < (float)io.r2;
    }
    return on;
}

__device__ __forceinline__ void mlp_store_row(const MlpIo& io, long long row, bool on, float f0, float f1, float f2) {
    if (!on) { f0 = 0.f; f1 = 0.f; f2 = 0.f; }
    if (io.mode == 1 && io.precision == 1) {
        double* o = (double*)io.out + row * 3;
        if (io.accumulate) { o[0] += f0; o[1] += f1; o[2] += f2; }
        else { o[0] = f0; o[1] = f1; o[2] = f2; }
    } else {
        float* o = (float*)io.out + row * 3;
        if (io.accumulate) { o[0] += f0; o[1] += f1; o[2] += f2; }
        else { o[0] = f0; o[1] = f1; o[2] = f2; }
    }
}

// CUDA-core fp32 kernel: CTA = 256 threads, tile = 64 rows, weights + activations in shared memory.
constexpr int MLPF_ROWS = 64;
constexpr int MLPF_THREADS = 256;
constexpr int MLPF_LD1 = MLP_H1 + 4;  // padded activation strides (floats)
constexpr int MLPF_LD2 = MLP_H2 + 4;
constexpr size_t MLPF_SMEM = (size_t)(MLP_NPARAM + MLPF_ROWS * MLPF_LD1 + MLPF_ROWS * MLPF_LD2 + MLPF_ROWS * 8) * sizeof(float);

__global__ void __launch_bounds__(MLPF_THREADS, 1) mlp_fp32_kernel(const float* __restrict__ params, const MlpIo io) {
    extern __shared__ __align__(16) float smf[];
    float* sw = smf;                             // parameters
    float* sA = sw + MLP_NPARAM;                 // h1 / h3  [64][132]
    float* sB = sA + MLPF_ROWS * MLPF_LD1;       // h2       [64][68]
    float* sF = sB + MLPF_ROWS * MLPF_LD2;       // features [64][8] (6 + gate + pad)
    const int t = threadIdx.x;
    for (int i = t; i < MLP_NPARAM; i += MLPF_THREADS) sw[i] = params[i];
    const long long n_tiles = (io.M + MLPF_ROWS - 1) / MLPF_ROWS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * MLPF_ROWS;
        __syncthreads();
        if (t < MLPF_ROWS) {
            float x[6] = {0, 0, 0, 0, 0, 0};
            bool on = false;
            if (row0 + t < io.M) on = mlp_fetch_row(io, row0 + t, x);
#pragma unroll
            for (int i = 0; i < 6; i++) sF[t * 8 + i] = x[i];
            sF[t * 8 + 6] = on ? 1.f : 0.f;
        }
        __syncthreads();
        // layer 1: 64 x 128, thread -> (row r = t/4, 32 neurons n = (t%4) + 4*i)
        {
            const int r = t >> 2, nb = t & 3;
            float x[6];
#pragma unroll
            for (int i = 0; i < 6; i++) x[i] = sF[r * 8 + i];
#pragma unroll 8
            for (int i = 0; i < 32; i++) {
                const int n = nb + 4 * i;
                float acc = sw[MLP_OB1 + n];
#pragma unroll
                for (int q = 0; q < 6; q++) acc = fmaf(sw[MLP_OW1 + n * 6 + q], x[q], acc);
                sA[r * MLPF_LD1 + n] = fmaxf(acc, 0.f);
            }
        }
        __syncthreads();
        // layer 2: 64 x 64, K = 128.  thread -> rows r0..r0+3 (r0 = 4*(t/16)), neurons n = (t%16) + 16*c
        {
            const int r0 = (t >> 4) * 4, nb = t & 15;
            float acc[4][4];
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int r = 0; r < 4; r++) acc[r][c] = sw[MLP_OB2 + nb + 16 * c];
            for (int k = 0; k < MLP_H1; k += 4) {
                float4 a[4], w[4];
#pragma unroll
                for (int r = 0; r < 4; r++) a[r] = *reinterpret_cast<const float4*>(sA + (r0 + r) * MLPF_LD1 + k);
#pragma unroll
                for (int c = 0; c < 4; c++) w[c] = *reinterpret_cast<const float4*>(sw + MLP_OW2 + (nb + 16 * c) * MLP_H1 + k);
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        acc[r][c] = fmaf(a[r].x, w[c].x, acc[r][c]);
                        acc[r][c] = fmaf(a[r].y, w[c].y, acc[r][c]);
                        acc[r][c] = fmaf(a[r].z, w[c].z, acc[r][c]);
                        acc[r][c] = fmaf(a[r].w, w[c].w, acc[r][c]);
                    }
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) sB[(r0 + r) * MLPF_LD2 + nb + 16 * c] = fmaxf(acc[r][c], 0.f);
        }
        __syncthreads();
        // layer 3: 64 x 128, K = 64.  thread -> rows r0..r0+3, neurons n = (t%16) + 16*c, c < 8
        {
            const int r0 = (t >> 4) * 4, nb = t & 15;
            float acc[4][8];
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 4; r++) acc[r][c] = sw[MLP_OB3 + nb + 16 * c];
            for (int k = 0; k < MLP_H2; k += 4) {
                float4 a[4];
#pragma unroll
                for (int r = 0; r < 4; r++) a[r] = *reinterpret_cast<const float4*>(sB + (r0 + r) * MLPF_LD2 + k);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const float4 w = *reinterpret_cast<const float4*>(sw + MLP_OW3 + (nb + 16 * c) * MLP_H2 + k);
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        acc[r][c] = fmaf(a[r].x, w.x, acc[r][c]);
                        acc[r][c] = fmaf(a[r].y, w.y, acc[r][c]);
                        acc[r][c] = fmaf(a[r].z, w.z, acc[r][c]);
                        acc[r][c] = fmaf(a[r].w, w.w, acc[r][c]);
                    }
                }
            }
            __syncthreads();  // everyone is done reading h1 (layer 2 finished before the previous barrier)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 8; c++) sA[(r0 + r) * MLPF_LD1 + nb + 16 * c] = fmaxf(acc[r][c], 0.f);
        }
        __syncthreads();
        // layer 4: 64 x 3, K = 128.  thread t < 192 -> (row t/3, output t%3)
        if (t < MLPF_ROWS * 3) {
            const int r = t / 3, o = t - 3 * r;
            float acc = sw[MLP_OB4 + o];
            for (int k = 0; k < MLP_H3; k += 4) {
                const float4 a = *reinterpret_cast<const float4*>(sA + r * MLPF_LD1 + k);
                const float4 w = *reinterpret_cast<const float4*>(sw + MLP_OW4 + o * MLP_H3 + k);
                acc = fmaf(a.x, w.x, acc);
                acc = fmaf(a.y, w.y, acc);
                acc = fmaf(a.z, w.z, acc);
                acc = fmaf(a.w, w.w, acc);
            }
            sB[r * 4 + o] = acc;  // h2 is dead; reuse as the output tile
        }
        __syncthreads();
        if (t < MLPF_ROWS && row0 + t < io.M)
            mlp_store_row(io, row0 + t, sF[t * 8 + 6] != 0.f, sB[t * 4 + 0], sB[t * 4 + 1], sB[t * 4 + 2]);
    }
}

// Latency kernel for a handful of rows (the reference's own call: 21 rows per DownwashNN.update): one CTA per row,
// 256 threads split every layer's dot products (partial sums combined with shuffles); the weights are read straight
// from L2 (71 KB per row) instead of being staged per CTA, so 21 rows take 21 SMs for a few microseconds instead of
// one SM for ~50.  fp32 FMA arithmetic like mlp_fp32_kernel (summation order differs: ~1e-7 relative).
constexpr int MLPR_THREADS = 256;
constexpr long long MLPR_MAX_ROWS = 512;
__global__ void __launch_bounds__(MLPR_THREADS) mlp_row_kernel(const float* __restrict__ params, const MlpIo io) {
    __shared__ __align__(16) float sx[8], h1[MLP_H1], h2[MLP_H2], h3[MLP_H3], so[4];
    const long long row = blockIdx.x;
    const int t = threadIdx.x;
    if (t == 0) {
        float x[6];
        const bool on = mlp_fetch_row(io, row, x);
#pragma unroll
        for (int i = 0; i < 6; i++) sx[i] = x[i];
        sx[6] = on ? 1.f : 0.f;
    }
    __syncthreads();
    if (t < MLP_H1) {   // layer 1: 128 outputs, K = 6
        const float* w = params + MLP_OW1 + t * MLP_IN;
        float acc = params[MLP_OB1 + t];
#pragma unroll
        for (int q = 0; q < MLP_IN; q++) acc = fmaf(w[q], sx[q], acc);
        h1[t] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    {   // layer 2: 64 outputs x K = 128, four threads per output (32 inputs each)
        const int o = t >> 2, part = t & 3;
        const float4* w = reinterpret_cast<const float4*>(params + MLP_OW2 + o * MLP_H1 + part * 32);
        const float4* a = reinterpret_cast<const float4*>(h1 + part * 32);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 wv = w[i], av = a[i];
            acc = fmaf(wv.x, av.x, acc); acc = fmaf(wv.y, av.y, acc); acc = fmaf(wv.z, av.z, acc); acc = fmaf(wv.w, av.w, acc);
        }
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
        if (part == 0) h2[o] = fmaxf(acc + params[MLP_OB2 + o], 0.f);
    }
    __syncthreads();
    {   // layer 3: 128 outputs x K = 64, two threads per output
        const int o = t >> 1, part = t & 1;
        const float4* w = reinterpret_cast<const float4*>(params + MLP_OW3 + o * MLP_H2 + part * 32);
        const float4* a = reinterpret_cast<const float4*>(h2 + part * 32);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 wv = w[i], av = a[i];
            acc = fmaf(wv.x, av.x, acc); acc = fmaf(wv.y, av.y, acc); acc = fmaf(wv.z, av.z, acc); acc = fmaf(wv.w, av.w, acc);
        }
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
        if (part == 0) h3[o] = fmaxf(acc + params[MLP_OB3 + o], 0.f);
    }
    __syncthreads();
    if (t < 32 * MLP_OUT) {   // layer 4: 3 outputs x K = 128, one warp per output
        const int o = t >> 5, lane = t & 31;
        const float4 wv = *reinterpret_cast<const float4*>(params + MLP_OW4 + o * MLP_H3 + lane * 4);
        const float4 av = *reinterpret_cast<const float4*>(h3 + lane * 4);
        float acc = fmaf(wv.x, av.x, fmaf(wv.y, av.y, fmaf(wv.z, av.z, wv.w * av.w)));
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, s);
        if (lane == 0) so[o] = acc + params[MLP_OB4 + o];
    }
    __syncthreads();
    if (t == 0) mlp_store_row(io, row, sx[6] != 0.f, so[0], so[1], so[2]);
}

// ---- swarm support: neighbour lists and ordered reduction ----
// One warp per ego builds its gated neighbour list in ascending j (deterministic order inside the ego's
// segment): ballot-compacted count pass, one atomicAdd on *total for the segment offset, fill pass.  The
// node-0 positions of all quads stream through shared-memory tiles, so every (possibly remote) position is
// read once per CTA.  Segments of different egos land in arrival order; the per-ego sum (swarm_reduce_kernel)
// walks only the ego's own segment, so the result does not depend on it.  Pairs beyond `cap` are dropped
// (*total still counts them: the host re-runs with a larger buffer when it cannot rule that out up front).
constexpr int SWARM_TILE = 2048;
constexpr int SWARM_CTA = 1024;  // 32 egos per CTA
__global__ void __launch_bounds__(SWARM_CTA) swarm_pairs_kernel(const TrajParts tp, const float* __restrict__ odom_xy, int n_all, int ego_begin,
                                                                 int n_ego, int n_nodes, float r2, int* __restrict__ total,
                                                                 int2* __restrict__ seg, int2* __restrict__ pairs, int cap, int group) {
    __shared__ float2 sxy[SWARM_TILE];
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (SWARM_CTA / 32) + (threadIdx.x >> 5);
    const bool act = i < n_ego;
    const int gi = ego_begin + i;
    float ex = 0.f, ey = 0.f;
    if (act) {
        const float* me = tp.row(gi, n_nodes);
        ex = odom_xy ? odom_xy[i * 2] : me[0];
        ey = odom_xy ? odom_xy[i * 2 + 1] : me[1];
    }
    int cnt = 0, off = 0;
    for (int pass = 0; pass < 2; pass++) {
        for (int j0 = 0; j0 < n_all; j0 += SWARM_TILE) {
            const int m = min(SWARM_TILE, n_all - j0);
            if (pass == 0 || n_all > SWARM_TILE) {   // a single tile stays resident for the fill pass
                __syncthreads();
                for (int e = threadIdx.x; e < m; e += SWARM_CTA) {
                    const float* q = tp.row(j0 + e, n_nodes);
                    sxy[e] = make_float2(q[0], q[1]);
                }
                __syncthreads();
            }
            if (act) {
                for (int e0 = 0; e0 < m; e0 += 32) {
                    const int e = e0 + lane, j = j0 + e;
                    bool hit = false;
                    // group > 0: only quads of the same contiguous block of `group` interact (independent scenarios)
                    if (e < m && j != gi && (group <= 0 || j / group == gi / group)) {
                        const float dx = sxy[e].x - ex, dy = sxy[e].y - ey;
                        hit = dx * dx + dy * dy < r2;
                    }
                    const unsigned b = __ballot_sync(0xFFFFFFFFu, hit);
                    if (pass == 0) cnt += __popc(b);
                    else {
                        const int o = off + __popc(b & ((1u << lane) - 1u));
                        if (hit && o < cap) pairs[o] = make_int2(gi, j);
                        off += __popc(b);
                    }
                }
            }
        }
        if (pass == 0 && act) {
            if (lane == 0) {
                off = cnt ? atomicAdd(total, cnt) : 0;
                seg[i] = make_int2(off, cnt);
            }
            off = __shfl_sync(0xFFFFFFFFu, off, 0);
        }
    }
}
// f[i][k][:] = sum over the ego's pair segment, in list order
template <typename TO>
__global__ void swarm_reduce_kernel(const float* __restrict__ fpair, const int2* __restrict__ seg, int n_ego, int n_nodes, int cap,
                                    TO* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n_ego * n_nodes * 3;
    if (idx >= total) return;
    const int i = (int)(idx / (n_nodes * 3));
    const int rem = (int)(idx - (long long)i * n_nodes * 3);
    const int2 sg = seg[i];
    const int end = min(sg.x + sg.y, cap);
    float acc = 0.f;
    for (int p = sg.x; p < end; p++) acc += fpair[(long long)p * n_nodes * 3 + rem];
    out[idx] = (TO)acc;
}

}  // namespace ndp
