// Bisect: 2-D TMA load, tensor map as __grid_constant__ param vs in global memory.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ void body(const CUtensorMap* tm, float* out, int n, int c0, int c1) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    float* dst = reinterpret_cast<float*>(smem + 1024);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(su32(bar)), "r"(n * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(su32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(su32(bar)), "r"(c0), "r"(c1) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(su32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}
__global__ void k_param(const __grid_constant__ CUtensorMap tm, float* out, int n, int c0, int c1) { body(&tm, out, n, c0, c1); }
__global__ void k_global(const CUtensorMap* tm, float* out, int n, int c0, int c1) { body(tm, out, n, c0, c1); }
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int H = 64, W = 64;
    std::vector<float> h((size_t)H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* q = nullptr; cudaDriverEntryPointQueryResult r;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r);
    Enc enc = (Enc)q;
    CUtensorMap tm;
    cuuint64_t dims[2] = {W, H}, str[1] = {W * 4ull};
    cuuint32_t box[2] = {32, 8}, es[2] = {1, 1};
    CUresult e = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)e);
    const int n = 32 * 8;
    cudaMalloc(&o, n * 4);
    CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    std::vector<float> ho(n);
    for (int variant = 0; variant < 3; ++variant) {
        cudaMemset(o, 0, n * 4);
        if (variant == 0) k_global<<<1, 128, 1024 + n * 4>>>(dtm, o, n, 0, 0);
        if (variant == 1) k_param<<<1, 128, 1024 + n * 4>>>(tm, o, n, 0, 0);
        if (variant == 2) k_param<<<1, 128, 1024 + n * 4>>>(tm, o, n, -1, -1);
        cudaError_t ce = cudaDeviceSynchronize();
        cudaMemcpy(ho.data(), o, n * 4, cudaMemcpyDeviceToHost);
        printf("variant %d: %s | %g %g %g ... row1: %g %g\n", variant, cudaGetErrorString(ce), ho[0], ho[1], ho[2], ho[32], ho[33]);
        if (ce != cudaSuccess) break;
    }
    return 0;
}
