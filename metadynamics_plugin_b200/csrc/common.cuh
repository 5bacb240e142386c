// common.cuh -- shared helpers for the sm_100a kernels behind include/metad_b200.h
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/metad_b200.h"

namespace metad {

// ---- error plumbing (no exceptions cross the C ABI) ---------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define METAD_CUDA(call)                                                              \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) return ::metad::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define METAD_LAUNCH_CHECK() METAD_CUDA(cudaGetLastError())

#ifdef __CUDACC__
// Programmatic dependent launch (sm_90+).  Kernels of the per-step sequence are launched with the
// programmatic-stream-serialization attribute (launch_pdl below): the next kernel's CTAs may be scheduled as soon as
// every CTA of this one has passed pdl_trigger(), and they block in pdl_wait() until this grid has completed and its
// memory is visible.  Everything before pdl_wait() must therefore touch only shared memory and data that no kernel of
// the sequence writes (twiddle tables).  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... Exp, class... Act>
inline cudaError_t launch_pdl(bool pdl, void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Exp>(args)...);
}
#endif

#ifdef __CUDACC__
// the same with a thread-block cluster of `cluster_x` CTAs along x (distributed shared memory between them)
template <class... Exp, class... Act>
inline cudaError_t launch_cluster_pdl(bool pdl, unsigned cluster_x, void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Exp>(args)...);
}
#endif

#define METAD_REQUIRE(cond, msg)                 \
    do {                                         \
        if (!(cond)) {                           \
            ::metad::set_error(msg);             \
            return METAD_ERR_INVALID;            \
        }                                        \
    } while (0)

int device_sm_count();

// ---- single-precision box, derived exactly like a SINGLE_PRECISION HOOMD BoxDim -----------------
struct BoxF {
    float lo[3], hi[3], L[3];
};
inline BoxF make_boxf(const metad_box* b) {
    BoxF r;
    for (int i = 0; i < 3; ++i) {
        r.L[i] = (float)b->L[i];
        r.hi[i] = r.L[i] / 2.0f;
        r.lo[i] = -r.hi[i];
    }
    return r;
}

// ---- device reductions ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of a double; result valid in thread 0.  smem must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = 0.0;
    if (w == 0) {
        r = lane < nw ? smem[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// streaming 128-bit load that does not allocate in L1 (data touched once per pass)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

}  // namespace metad
