// Shared-memory wavefront microbenchmark: what does a 128-bit broadcast load cost when the two problems of a warp read
// different addresses?  Patterns (per warp-wide LDS.128):
//   0  every lane the same 16-byte chunk
//   1  lanes 0-15 chunk A, lanes 16-31 chunk B (the kernel's layout: one problem per half warp), A / B on disjoint bank halves
//   2  even lanes chunk A, odd lanes chunk B (problems interleaved lane by lane)
//   3  every lane its own chunk (32 x 16 B)
//   4  pattern 1 as LDS.64      5  pattern 1 as LDS.32
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bcast lds_bcast.cu ; run: ./lds_bcast
#include <cstdio>
#include <cuda_runtime.h>

template <int kMode>
__global__ void bench(float* out, long long* cyc, int iters) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 1e-6f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int off;  // in floats
    if (kMode == 0) off = 0;
    else if (kMode == 1 || kMode == 4 || kMode == 5) off = (lane >> 4) * (16 + 32 * 9);
    else if (kMode == 2) off = (lane & 1) * (16 + 32 * 9);
    else off = lane * 4;
    off += warp * 36;  // different warps, different rows
    float4 acc = make_float4(0, 0, 0, 0);
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + off * 4;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const unsigned a = base + ((it * 16 + u) & 63) * 64;
            if (kMode == 4) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
                acc.x += v.x; acc.y += v.y;
            } else if (kMode == 5) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
                acc.x += v;
            } else {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int kMode>
void run(const char* name) {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2048, threads = 512;
    bench<kMode><<<148, threads>>>(out, cyc, iters);
    bench<kMode><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %.2f cycles per warp-level load per SM (16 warps)\n", name, (double)h / (iters * 16.0 * (threads / 32)));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("LDS.128 all lanes one chunk");
    run<1>("LDS.128 half warps A | B");
    run<2>("LDS.128 even / odd lanes A | B");
    run<3>("LDS.128 32 distinct chunks");
    run<4>("LDS.64  half warps A | B");
    run<5>("LDS.32  half warps A | B");
    return 0;
}
