// Stand-alone probe: cp.async.bulk.tensor.{3d,5d} of an fp32 NCDHW tensor with out-of-bounds boxes.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int n, int c0, int c1, int c2, int c3, int c4) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    float* dst = reinterpret_cast<float*>(smem + 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(su32(bar)), "r"(n * 4) : "memory");
        if (RANK == 5)
            asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                         ::"r"(su32(dst)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(su32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(su32(dst)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(su32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(su32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int B = 1, C = 16, D = 2, H = 20, W = 40;
    std::vector<float> h((size_t)B * C * D * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* q = nullptr; cudaDriverEntryPointQueryResult r;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r);
    Enc enc = (Enc)q;
    printf("entry %p res %d\n", q, (int)r);
    for (int rank : {3, 5}) {
        const int SR = 10;
        CUtensorMap tm;
        cuuint64_t dims5[5] = {W, H, D, C, B}, str5[4] = {W * 4ull, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4, (cuuint64_t)C * D * H * W * 4};
        cuuint32_t box5[5] = {32, SR, 1, 8, 1}, es[5] = {1, 1, 1, 1, 1};
        cuuint64_t dims3[3] = {W, H, (cuuint64_t)D * C}, str3[2] = {W * 4ull, (cuuint64_t)H * W * 4};
        cuuint32_t box3[3] = {32, SR, 8};
        CUresult e = rank == 5 ? enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, dims5, str5, box5, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                               : enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims3, str3, box3, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int n = 32 * SR * 8;
        cudaMalloc(&o, n * 4);
        cudaFuncSetAttribute(k<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (rank == 5) k<5><<<1, 128, 128 + n * 4>>>(tm, o, n, -1, -1, 1, 8, 0);
        else k<3><<<1, 128, 128 + n * 4>>>(tm, o, n, -1, -1, 8, 0, 0);
        cudaError_t ce = cudaDeviceSynchronize();
        std::vector<float> ho(n);
        cudaMemcpy(ho.data(), o, n * 4, cudaMemcpyDeviceToHost);
        printf("rank %d encode %d sync %s | first row: %g %g %g | second row: %g %g ... expect 0 0 0 | 0 %g\n", rank, (int)e, cudaGetErrorString(ce),
               ho[0], ho[1], ho[2], ho[32], ho[33], rank == 5 ? h[(size_t)(8 * D + 1) * H * W] : h[(size_t)8 * H * W]);
    }
    return 0;
}
