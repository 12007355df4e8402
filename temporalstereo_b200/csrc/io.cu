// The data formats either side of the hot path: the wire format of the input images and the
// on-device evaluation of the output disparity.
//
// ref: architecture/data/datasets/base.py:120-127 (ToTensor: uint8 HWC -> float CHW / 255; Normalize: (v - mean) / std),
//      :183-185 (eval-time bilinear resize, align_corners=True: tstereo_bilinear_resize);
//      architecture/data/evaluation/pixel_error.py:6-71 (calc_error: masked EPE and 1/2/3/5-px error rates).
#include "common.cuh"

namespace tstereo {

// thread = one pixel: 3 interleaved bytes in, 3 planar floats out (coalesced per plane).
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, long long osB, long long osC, int HW,
                    float m0, float m1, float m2, float s0, float s1, float s2, long long total) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const long long b = i / HW;
    const int p = (int)(i - b * HW);
    const unsigned char* q = in + i * 3;
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    float* o = out + b * osB + p;
#pragma unroll
    for (int c = 0; c < 3; ++c)   // same op order as ToTensor().div(255) then Normalize.sub_(mean).div_(std)
        o[c * osC] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)q[c], 255.0f), mean[c]), sd[c]);
}

// acc[0] = sum |gt - est| over the mask, acc[1] = mask count, acc[2..5] = counts of |err| > 1, 2, 3, 5 px (doubles;
// integer-valued counts and a double sum: the result does not depend on the atomics' order beyond double rounding).
__global__ void __launch_bounds__(256)
disp_error_kernel(const float* __restrict__ est, const float* __restrict__ gt, float lb, float ub, int use_lb, int use_ub,
                  long long total, double* __restrict__ acc) {
    double s = 0.0;
    unsigned n = 0, c1 = 0, c2 = 0, c3 = 0, c5 = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const float g = __ldg(gt + i);
        if ((use_lb && !(g > lb)) || (use_ub && !(g < ub))) continue;
        const float e = fabsf(__fsub_rn(g, __ldg(est + i)));
        s += (double)e;
        ++n;
        c1 += e > 1.0f;
        c2 += e > 2.0f;
        c3 += e > 3.0f;
        c5 += e > 5.0f;
    }
    __shared__ double sm[6][8];
    double v[6] = {s, (double)n, (double)c1, (double)c2, (double)c3, (double)c5};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[threadIdx.x][w];
        atomicAdd(acc + threadIdx.x, t);
    }
}

}  // namespace tstereo

using namespace tstereo;

extern "C" {

int tstereo_normalize_u8(const unsigned char* in, float* out, long long osB, long long osC, int B, int H, int W,
                         const float* mean3, const float* std3, void* stream) {
    TS_REQUIRE(in && out && mean3 && std3, "normalize_u8: null pointer");
    TS_REQUIRE(B > 0 && H > 0 && W > 0 && osC >= (long long)H * W, "normalize_u8: bad sizes");
    const long long total = (long long)B * H * W;
    normalize_u8_kernel<<<(unsigned)cdivll(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, osB, osC, H * W, mean3[0], mean3[1],
                                                                                      mean3[2], std3[0], std3[1], std3[2], total);
    return check_launch("normalize_u8");
}

int tstereo_disp_error(const float* est, const float* gt, float lb, float ub, int use_lb, int use_ub, long long n,
                       double* acc6, void* stream) {
    TS_REQUIRE(est && gt && acc6 && n > 0, "disp_error: bad arguments");
    TS_REQUIRE((((size_t)acc6) & 7) == 0, "disp_error: accumulator must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(acc6, 0, 6 * sizeof(double), st);
    if (e != cudaSuccess) {
        set_error("disp_error: cudaMemsetAsync: %s", cudaGetErrorString(e));
        return TSTEREO_E_CUDA;
    }
    const unsigned grid = (unsigned)(cdivll(n, 256) < 148 * 8 ? cdivll(n, 256) : 148 * 8);
    disp_error_kernel<<<grid, 256, 0, st>>>(est, gt, lb, ub, use_lb, use_ub, n, acc6);
    return check_launch("disp_error");
}

}  // extern "C"
