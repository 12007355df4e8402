// Shared helpers for libtstereo.so kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/tstereo.h"

namespace tstereo {

void set_error(const char* fmt, ...);
void count_launch();

inline int check_launch(const char* what) {
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return TSTEREO_E_CUDA;
    }
    return TSTEREO_OK;
}

// ---- programmatic dependent launch (PDL): every hot-path kernel starts with pdl_sync() — it lets the NEXT kernel of the
// stream be scheduled right away (its CTAs run their prologue: barrier init, TMEM allocation, index math) and then waits
// until the PREVIOUS kernel has completed and flushed its writes, before touching any dependent memory.  launch_k() sets
// the matching launch attribute (TSTEREO_PDL=0: plain stream order).  Captured into a CUDA graph the edges become
// programmatic dependencies: the ~1-2 us launch gap between the ~160 dependent kernels of a frame overlaps their tails.
bool pdl_enabled();
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define TS_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            tstereo::set_error(__VA_ARGS__);  \
            return TSTEREO_E_ARG;             \
        }                                     \
    } while (0)

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

// x * sigmoid(x); full-precision expf so the result tracks the fp32 reference to ~1 ulp.
__device__ __forceinline__ float silu_f(float x) { return __fdiv_rn(x, 1.0f + expf(-x)); }

// SiLU with ex2.approx / rcp.approx (~1e-6 relative, the form the tensor-core epilogues use): ~6 instructions instead of
// the ~50 of the full-precision expf + IEEE division
__device__ __forceinline__ float silu_fast(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
__device__ __forceinline__ float apply_act_fast(float x, int act) {
    if (act == TSTEREO_ACT_SILU) return silu_fast(x);
    if (act == TSTEREO_ACT_RELU) return fmaxf(x, 0.0f);
    return x;
}

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == TSTEREO_ACT_SILU) return silu_f(x);
    if (act == TSTEREO_ACT_RELU) return fmaxf(x, 0.0f);
    return x;
}

// 4-byte cp.async with zero fill when !valid (src must still be a legal address).
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- S-format (fp16 hi / lo split, include/tstereo.h `tstereo_split`) helpers
// two floats -> packed fp16x2 (low half = a).  satfinite: a value beyond the fp16 range clamps to +-65504 (and its
// lo part likewise) instead of turning the whole accumulation into inf - inf = NaN
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t h) {
    float2 f;
    asm("{\n\t.reg .b16 l, u;\n\tmov.b32 {l, u}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, u;\n\t}" : "=f"(f.x), "=f"(f.y) : "r"(h));
    return f;
}

__device__ __forceinline__ void stg128(void* ptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}


// ATen's align_corners source index: scale = (in-1)/(out-1) in float, src = scale*dst
// (aten/src/ATen/native/UpSample.h area_pixel_compute_scale / guard_index_and_lambda).
struct LerpIdx {
    int i0, i1;
    float w0, w1;
};
__device__ __forceinline__ float ac_scale(int in_size, int out_size) {
    return out_size > 1 ? __fdiv_rn((float)(in_size - 1), (float)(out_size - 1)) : 0.0f;
}
__device__ __forceinline__ LerpIdx ac_index(float scale, int dst, int in_size) {
    float src = __fmul_rn(scale, (float)dst);
    int i0 = min((int)floorf(src), in_size - 1);
    float l = fminf(fmaxf(src - (float)i0, 0.0f), 1.0f);
    LerpIdx r;
    r.i0 = i0;
    r.i1 = min(i0 + 1, in_size - 1);
    r.w1 = l;
    r.w0 = 1.0f - l;
    return r;
}

}  // namespace tstereo
