// Micro test: which forms of cp.reduce.async.bulk.tensor (UTMAREDG) run on sm_100a?  Each variant adds one shared-memory box
// to a small 3-D tensor and checks the result.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_reduce_test tma_reduce_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/ptx>
#include <cstdio>
#include <vector>
#include <cstring>

template <class T> __global__ void k(const __grid_constant__ CUtensorMap tm, int bx, int by, int bz, int cx, int cy, int cz) {
    extern __shared__ __align__(128) unsigned char raw[];
    T* box = reinterpret_cast<T*>(raw);
    const int n = bx * by * bz;
    for (int i = threadIdx.x; i < n; i += blockDim.x) box[i] = (T)(i % 7 + 1);
    __syncthreads();
    cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int32_t crd[3] = {cx, cy, cz};
        cuda::ptx::cp_reduce_async_bulk_tensor(cuda::ptx::space_global, cuda::ptx::space_shared, cuda::ptx::op_add, &tm, crd, box);
        cuda::ptx::cp_async_bulk_commit_group();
        cuda::ptx::cp_async_bulk_wait_group_read(cuda::ptx::n32_t<0>());
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class T> int run(EncodeFn enc, CUtensorMapDataType dt, const char* name, int nx, int ny, int nz, int bx, int by, int bz, int cx, int cy, int cz) {
    T* d;
    const size_t M = (size_t)nx * ny * nz;
    cudaMalloc(&d, M * sizeof(T));
    cudaMemset(d, 0, M * sizeof(T));
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz}, strides[2] = {(cuuint64_t)nx * sizeof(T), (cuuint64_t)nx * ny * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, dt, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-28s encode failed %d\n", name, (int)r); return 1; }
    const size_t smem = (size_t)bx * by * bz * sizeof(T);
    cudaFuncSetAttribute(k<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<T><<<1, 128, smem>>>(tm, bx, by, bz, cx, cy, cz);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s box %dx%dx%d at (%d,%d,%d): CUDA error %s\n", name, bx, by, bz, cx, cy, cz, cudaGetErrorString(e)); return 2; }
    std::vector<T> h(M);
    cudaMemcpy(h.data(), d, M * sizeof(T), cudaMemcpyDeviceToHost);
    long bad = 0; double sum = 0, want = 0;
    for (int z = 0; z < nz; ++z) for (int y = 0; y < ny; ++y) for (int x = 0; x < nx; ++x) {
        const int lx = x - cx, ly = y - cy, lz = z - cz;
        double w = 0;
        if (lx >= 0 && lx < bx && ly >= 0 && ly < by && lz >= 0 && lz < bz) w = ((lz * by + ly) * bx + lx) % 7 + 1;
        const double v = (double)h[((size_t)z * ny + y) * nx + x];
        if (v != w) ++bad;
        sum += v; want += w;
    }
    printf("%-28s box %dx%dx%d at (%d,%d,%d): ok, mismatches %ld (sum %.0f, expected %.0f)\n", name, bx, by, bz, cx, cy, cz, bad, sum, want);
    cudaFree(d);
    return 0;
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    int rc = 0;
    rc |= run<float>(enc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, "f32 interior", 64, 64, 64, 24, 20, 20, 8, 8, 8);
    if (rc == 2) { cudaDeviceReset(); }
    rc |= run<int>(enc, CU_TENSOR_MAP_DATA_TYPE_INT32, "s32 interior", 64, 64, 64, 24, 20, 20, 8, 8, 8);
    if (rc & 2) { cudaDeviceReset(); }
    rc |= run<unsigned>(enc, CU_TENSOR_MAP_DATA_TYPE_UINT32, "u32 interior", 64, 64, 64, 24, 20, 20, 8, 8, 8);
    if (rc & 2) { cudaDeviceReset(); }
    rc |= run<int>(enc, CU_TENSOR_MAP_DATA_TYPE_INT32, "s32 negative coords", 64, 64, 64, 24, 20, 20, -4, -2, -2);
    if (rc & 2) { cudaDeviceReset(); }
    rc |= run<int>(enc, CU_TENSOR_MAP_DATA_TYPE_INT32, "s32 upper overhang", 64, 64, 64, 24, 20, 20, 60, 62, 62);
    if (rc & 2) { cudaDeviceReset(); }
    rc |= run<int>(enc, CU_TENSOR_MAP_DATA_TYPE_INT32, "s32 small box", 32, 32, 32, 16, 12, 12, 4, 6, 6);
    return rc;
}
