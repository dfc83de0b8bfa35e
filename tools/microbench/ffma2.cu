// Issue-rate microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2, sm_100a) -- does the packed form halve the issue
// slots per FMA?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run: ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

template <int kMode, int kIlp>
__global__ void bench(float* out, long long* cyc, int iters, float s) {
    float2 acc[kIlp];
#pragma unroll
    for (int i = 0; i < kIlp; i++) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 a = make_float2(s, s * 0.5f), b = make_float2(0.999f, 1.001f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < kIlp; i++) {
            if (kMode == 0) {  // two scalar FFMAs
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(a.x), "f"(b.x));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].y) : "f"(a.y), "f"(b.y));
            } else {
                acc[i] = ffma2(a, b, acc[i]);
            }
        }
    }
    const long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < kIlp; i++) r += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int kMode, int kIlp>
void run(const char* name, int threads) {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    bench<kMode, kIlp><<<148, threads>>>(out, cyc, iters, 1e-3f);
    bench<kMode, kIlp><<<148, threads>>>(out, cyc, iters, 1e-3f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double fma_per_thread = 2.0 * kIlp * iters;
    const int warps_per_smsp = threads / 128;
    printf("%-8s ilp=%d warps/SMSP=%d : %.3f cycles per warp-level FMA pair-issue per SMSP  (%.1f FMA lanes/clk/SM)\n", name, kIlp, warps_per_smsp,
           (double)h / (kIlp * (double)iters * warps_per_smsp), fma_per_thread * threads / (double)h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512, 1024}) {
        if (threads == 128) { run<0, 8>("FFMAx2", 128); run<1, 8>("FFMA2", 128); }
        if (threads == 256) { run<0, 8>("FFMAx2", 256); run<1, 8>("FFMA2", 256); }
        if (threads == 512) { run<0, 8>("FFMAx2", 512); run<1, 8>("FFMA2", 512); }
        if (threads == 1024) { run<0, 4>("FFMAx2", 1024); run<1, 4>("FFMA2", 1024); }
    }
    return 0;
}
