// One instantiation of the SQP-RTI kernels per translation unit (precision x horizon x latency build): the nominal
// kernel and, for the non-latency builds, the constrained kernel that takes the problems it hands over.  One
// instantiation per module keeps ptxas to seconds, and the instantiations build in parallel (ndp_nmpc_qd_b200/build.py).
//   nvcc -c rti_inst.cu -DNDP_INST_T=float -DNDP_INST_N=20 -DNDP_INST_LAT=false -DNDP_INST_TAG=f32_20_0
#include "rti_kernel.cuh"

// a bare `nvcc -c rti_inst.cu` builds the headline instantiation
#ifndef NDP_INST_T
#define NDP_INST_T float
#define NDP_INST_N 20
#define NDP_INST_LAT false
#define NDP_INST_TAG f32_20_0
#endif

#ifndef NDP_INST_LAT_IS_TRUE
#define NDP_INST_LAT_IS_TRUE 0
#endif
#define NDP_CAT2(a, b) a##b
#define NDP_CAT(a, b) NDP_CAT2(a, b)

namespace ndp {

// pdl: launch as a programmatic dependent of the previous kernel in the stream (the kernel then starts while that
// kernel drains and synchronises on it with griddepcontrol.wait before it reads the forces)
void NDP_CAT(rti_launch_, NDP_INST_TAG)(int grid, int threads, size_t smem, cudaStream_t st, const RtiCfg<NDP_INST_T>& c,
                                        const RtiArgs<NDP_INST_T>& a, bool pdl) {
    if (!pdl) {
        rti_step_kernel<NDP_INST_T, NDP_INST_N, NDP_INST_LAT><<<grid, threads, smem, st>>>(c, a);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, rti_step_kernel<NDP_INST_T, NDP_INST_N, NDP_INST_LAT>, c, a);
}

#ifdef NDP_RTI_PROF
int NDP_CAT(rti_prof_, NDP_INST_TAG)(unsigned long long* host) { return (int)cudaMemcpyFromSymbol(host, g_rti_prof, sizeof(unsigned long long) * 2048 * 8); }
int NDP_CAT(rti_cprof_, NDP_INST_TAG)(unsigned long long* host) { return (int)cudaMemcpyFromSymbol(host, g_con_prof, sizeof(unsigned long long) * (2 * 8192 + 2)); }
#endif

const void* NDP_CAT(rti_kernel_, NDP_INST_TAG)() { return (const void*)rti_step_kernel<NDP_INST_T, NDP_INST_N, NDP_INST_LAT>; }

#if !NDP_INST_LAT_IS_TRUE
// the constrained kernel of this (precision, horizon): always a programmatic dependent of the nominal launch before it
void NDP_CAT(rti_claunch_, NDP_INST_TAG)(int grid, int threads, size_t smem, cudaStream_t st, const RtiCfg<NDP_INST_T>& c,
                                         const RtiArgs<NDP_INST_T>& a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, rti_constrained_kernel<NDP_INST_T, NDP_INST_N>, c, a);
}

const void* NDP_CAT(rti_ckernel_, NDP_INST_TAG)() { return (const void*)rti_constrained_kernel<NDP_INST_T, NDP_INST_N>; }
#endif

}  // namespace ndp
