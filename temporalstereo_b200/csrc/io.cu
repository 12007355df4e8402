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

// ---------------------------------------------------------------------------------------------------------------------
// Training losses, forward only (the engine has no backward): on-device evaluation of the two loss
// terms the reference logs per level, one launch each, accumulators in doubles.
//
// ref: architecture/modeling/losses/smooth_l1_loss.py:49-74 (loss_per_level), warsserstein_distance_loss.py:53-81.
// Ground truth at a coarser level: gt / scale, then adaptive average (dense) or max (sparse) pooling onto the level's
// grid (ATen adaptive pooling bins: [floor(i*in/out), ceil((i+1)*in/out))), scale = Wg / W; valid where
// start_disp < gt_s < max_disp / scale.
__device__ __forceinline__ float pooled_gt(const float* __restrict__ g, int Hg, int Wg, int H, int W, int y, int x, float scale,
                                           int sparse) {
    if (Hg == H && Wg == W) return __ldg(g + (size_t)y * Wg + x);
    const int y0 = (int)(((long long)y * Hg) / H), y1 = (int)((((long long)y + 1) * Hg + H - 1) / H);
    const int x0 = (int)(((long long)x * Wg) / W), x1 = (int)((((long long)x + 1) * Wg + W - 1) / W);
    float acc = sparse ? -INFINITY : 0.f;
    for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) {
            const float v = __fdiv_rn(__ldg(g + (size_t)yy * Wg + xx), scale);
            acc = sparse ? fmaxf(acc, v) : __fadd_rn(acc, v);
        }
    return sparse ? acc : __fdiv_rn(acc, (float)((y1 - y0) * (x1 - x0)));
}

__device__ __forceinline__ void block_sum2(double a, double b, double* acc) {
    __shared__ double sa[256], sb[256];
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) {
            sa[threadIdx.x] += sa[threadIdx.x + k];
            sb[threadIdx.x] += sb[threadIdx.x + k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd(acc, sa[0]);
        atomicAdd(acc + 1, sb[0]);
    }
}

// acc[0] = sum of smooth-L1(est - gt_s) over the valid pixels, acc[1] = their count
__global__ void __launch_bounds__(256)
loss_smooth_l1_kernel(const float* __restrict__ est, const float* __restrict__ gt, int H, int W, int Hg, int Wg, float scale,
                      float lo, float hi, int sparse, long long n, double* __restrict__ acc) {
    double s = 0.0, c = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const long long b = i / ((long long)W * H);
        const float g = pooled_gt(gt + b * (size_t)Hg * Wg, Hg, Wg, H, W, y, x, scale, sparse);
        if (g > lo && g < hi) {
            const float d = fabsf(__ldg(est + i) - g);
            s += (double)(d < 1.0f ? 0.5f * d * d : d - 0.5f);
            c += 1.0;
        }
    }
    block_sum2(s, c, acc);
}

// acc[0] = sum over ALL pixels of  sum_d (softmax_d(cost) + 0.25) * |off_d + sample_d - gt_s| * valid,  acc[1] = valid count
__global__ void __launch_bounds__(256)
loss_wasserstein_kernel(const float* __restrict__ cost, const float* __restrict__ off, const float* __restrict__ smp,
                        const float* __restrict__ gt, int D, int H, int W, int Hg, int Wg, float scale, float lo, float hi,
                        int sparse, long long n, double* __restrict__ acc) {
    double s = 0.0, c = 0.0;
    const long long HW = (long long)H * W;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const long long b = i / HW, p = i - b * HW;
        const int x = (int)(p % W), y = (int)(p / W);
        const float g = pooled_gt(gt + b * (size_t)Hg * Wg, Hg, Wg, H, W, y, x, scale, sparse);
        if (g > lo && g < hi) {
            const float* cp = cost + b * D * HW + p;
            float m = -INFINITY;
            for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(cp + d * HW));
            float den = 0.f;
            for (int d = 0; d < D; ++d) den += expf(__ldg(cp + d * HW) - m);
            float l = 0.f;
            for (int d = 0; d < D; ++d) {
                const float pr = __fdiv_rn(expf(__ldg(cp + d * HW) - m), den);
                const float e = fabsf(__ldg(off + b * D * HW + d * HW + p) + __ldg(smp + b * D * HW + d * HW + p) - g);
                l += (pr + 0.25f) * e;
            }
            s += (double)l;
            c += 1.0;
        }
    }
    block_sum2(s, c, acc);
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

static int loss_common(const float* gt, double* acc2, int B, int H, int W, int Hg, int Wg, cudaStream_t st, const char* what) {
    TS_REQUIRE(gt && acc2, "%s: null pointer", what);
    TS_REQUIRE(B > 0 && H > 0 && W > 0 && Hg >= H && Wg >= W, "%s: bad sizes (the ground truth must be at least as large as the level)", what);
    TS_REQUIRE((((size_t)acc2) & 7) == 0, "%s: accumulator must be 8-byte aligned", what);
    cudaError_t e = cudaMemsetAsync(acc2, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) {
        set_error("%s: cudaMemsetAsync: %s", what, cudaGetErrorString(e));
        return TSTEREO_E_CUDA;
    }
    return TSTEREO_OK;
}

int tstereo_loss_smooth_l1(const float* est, const float* gt, int B, int H, int W, int Hg, int Wg, float max_disp, float start_disp,
                           int sparse, double* acc2, void* stream) {
    TS_REQUIRE(est, "loss_smooth_l1: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = loss_common(gt, acc2, B, H, W, Hg, Wg, st, "loss_smooth_l1");
    if (rc) return rc;
    const float scale = (float)((double)Wg / (double)W);
    const long long n = (long long)B * H * W;
    const unsigned grid = (unsigned)(cdivll(n, 256) < 148 * 8 ? cdivll(n, 256) : 148 * 8);
    loss_smooth_l1_kernel<<<grid, 256, 0, st>>>(est, gt, H, W, Hg, Wg, scale, start_disp, (float)((double)max_disp / ((double)Wg / (double)W)),
                                                sparse, n, acc2);
    return check_launch("loss_smooth_l1");
}

int tstereo_loss_wasserstein(const float* cost, const float* off, const float* samples, const float* gt, int B, int D, int H, int W,
                             int Hg, int Wg, float max_disp, float start_disp, int sparse, double* acc2, void* stream) {
    TS_REQUIRE(cost && off && samples && D > 0, "loss_wasserstein: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = loss_common(gt, acc2, B, H, W, Hg, Wg, st, "loss_wasserstein");
    if (rc) return rc;
    const float scale = (float)((double)Wg / (double)W);
    const long long n = (long long)B * H * W;
    const unsigned grid = (unsigned)(cdivll(n, 256) < 148 * 8 ? cdivll(n, 256) : 148 * 8);
    loss_wasserstein_kernel<<<grid, 256, 0, st>>>(cost, off, samples, gt, D, H, W, Hg, Wg, scale, start_disp,
                                                  (float)((double)max_disp / ((double)Wg / (double)W)), sparse, n, acc2);
    return check_launch("loss_wasserstein");
}

}  // extern "C"
