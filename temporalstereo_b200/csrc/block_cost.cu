// Cost-volume construction.
//
// ref: architecture/modeling/aggregation/utils/block_cost.py:16-83 — int branch (:34-45),
//      tensor branch (:47-58) via layers/inverse_warp_3d.py:4-58, pooled group-wise terms
//      (:6-13, :64-78).
//
// Design (HBM-bound: the output is 6x the input bytes):
//  * main kernel: one warp = 32 consecutive x of a 4-row strip for one (b, d, group of 8
//    channels).  Every L / R / output access is a fully coalesced 128 B row segment.  The
//    per-pixel warp coordinate is computed once and reused for the 8 channels.  Because
//    avg_pool(L) - avg_pool(R_d) == avg_pool(L - R_d), the 2x2 and 4x4 pooled differences are
//    warp-shuffle reductions of the per-pixel difference (xor 1, xor 2) — no second pass over
//    the features and no shared memory.  The thread writes L, R_d (or -(L-R_d)^2), the
//    full-resolution group term g0 and the tiny pooled terms G1, G2 (scratch, L2 resident).
//  * resize kernel: bilinear align_corners up-sampling of G1, G2 into the last 2*C/8 planes.
#include "common.cuh"

namespace tstereo {

template <bool WARP>
__global__ void __launch_bounds__(128)
block_cost_main_kernel(const float* __restrict__ L, const float* __restrict__ R,
                       const float* __restrict__ smp, float* __restrict__ out,
                       float* __restrict__ g1, float* __restrict__ g2,
                       int C, int H, int W, int D) {
    const int G = C >> 3;
    int z = blockIdx.z;
    const int g = z % G;
    z /= G;
    const int d = z % D;
    const int b = z / D;
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int rq = blockIdx.y * 4 + threadIdx.y;  // 4-row strip index
    const int y0 = rq * 4;
    const size_t HW = (size_t)H * W;
    const int outC = (WARP ? 2 * C : C) + 3 * G;
    const float Wm1 = (float)(W - 1);

    int x0[4], ylo[4];
    float w0[4], w1[4], wy0[4], wy1[4];
    bool v0[4], v1[4], pin[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int y = y0 + r;
        pin[r] = (x < W) && (y < H);
        x0[r] = 0;
        w0[r] = 0.f;
        w1[r] = 0.f;
        v0[r] = false;
        v1[r] = false;
        ylo[r] = y;
        wy0[r] = 1.f;
        wy1[r] = 0.f;
        if (pin[r]) {
            if (WARP) {
                // The y coordinate goes through the same normalise / un-normalise round trip
                // (inverse_warp_3d.py:46, grid_sampler_unnormalize); for some (H, y) it lands a few
                // 1e-6 px off the integer, which blends two rows.  Warp-uniform (one y per warp).
                const float Hm1 = (float)(H - 1);
                const float gyn = __fsub_rn(__fmul_rn(__fdiv_rn((float)y, Hm1), 2.0f), 1.0f);
                const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gyn, 1.0f), 2.0f), Hm1);
                const float fy = floorf(iy);
                ylo[r] = (int)fy;
                wy0[r] = __fsub_rn(fy + 1.0f, iy);
                wy1[r] = __fsub_rn(iy, fy);
                // same op sequence as inverse_warp_3d.py:40-47 + ATen grid_sampler_unnormalize
                const float dsp = smp[((size_t)(b * D + d) * H + y) * W + x];
                const float gx = __fadd_rn((float)x, -dsp);
                const float gn = __fsub_rn(__fmul_rn(__fdiv_rn(gx, Wm1), 2.0f), 1.0f);
                const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gn, 1.0f), 2.0f), Wm1);
                const float fx = floorf(ix);
                if (fx >= -1.0f && fx <= Wm1) {
                    const int xi = (int)fx;
                    x0[r] = xi;
                    w0[r] = __fsub_rn(fx + 1.0f, ix);
                    w1[r] = __fsub_rn(ix, fx);
                    v0[r] = (xi >= 0);
                    v1[r] = (xi + 1 < W);
                }
            } else {
                x0[r] = x - d;
                w0[r] = 1.f;
                v0[r] = (x - d >= 0);
            }
        }
    }

    float a0[4] = {0.f, 0.f, 0.f, 0.f};
    float a1[2] = {0.f, 0.f};
    float a2 = 0.f;
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
        const int ch = g * 8 + c;
        const float* Lp = L + ((size_t)b * C + ch) * HW;
        const float* Rp = R + ((size_t)b * C + ch) * HW;
        float* o_first = out + (((size_t)b * outC + ch) * D + d) * HW;
        float* o_second = out + (((size_t)b * outC + C + ch) * D + d) * HW;  // WARP only
        float e[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float l = 0.f, rv = 0.f;
            if (pin[r]) {
                const size_t row = (size_t)(y0 + r) * W;
                l = __ldg(Lp + row + x);
                if (WARP) {
                    if (ylo[r] >= 0 && wy0[r] != 0.f) {
                        const size_t rr = (size_t)ylo[r] * W;
                        const float ra = v0[r] ? __ldg(Rp + rr + x0[r]) : 0.f;
                        const float rb = v1[r] ? __ldg(Rp + rr + x0[r] + 1) : 0.f;
                        rv = wy0[r] * fmaf(rb, w1[r], __fmul_rn(ra, w0[r]));
                    }
                    if (ylo[r] + 1 < H && wy1[r] != 0.f) {
                        const size_t rr = (size_t)(ylo[r] + 1) * W;
                        const float ra = v0[r] ? __ldg(Rp + rr + x0[r]) : 0.f;
                        const float rb = v1[r] ? __ldg(Rp + rr + x0[r] + 1) : 0.f;
                        rv = fmaf(wy1[r], fmaf(rb, w1[r], __fmul_rn(ra, w0[r])), rv);
                    }
                    o_first[row + x] = l;
                    o_second[row + x] = rv;
                } else {
                    const float ra = v0[r] ? __ldg(Rp + row + x0[r]) : 0.f;
                    rv = ra;
                    const float df = l - rv;
                    o_first[row + x] = -(df * df);
                }
            }
            e[r] = l - rv;
            a0[r] = fmaf(e[r], e[r], a0[r]);
        }
        float s01 = e[0] + e[1], s23 = e[2] + e[3];
        s01 += __shfl_xor_sync(0xffffffffu, s01, 1);
        s23 += __shfl_xor_sync(0xffffffffu, s23, 1);
        const float m1a = s01 * 0.25f, m1b = s23 * 0.25f;
        a1[0] = fmaf(m1a, m1a, a1[0]);
        a1[1] = fmaf(m1b, m1b, a1[1]);
        float s4 = s01 + s23;
        s4 += __shfl_xor_sync(0xffffffffu, s4, 2);
        const float m2 = s4 * 0.0625f;
        a2 = fmaf(m2, m2, a2);
    }

    const int base = WARP ? 2 * C : C;
    float* og0 = out + (((size_t)b * outC + base + g) * D + d) * HW;
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (pin[r]) og0[(size_t)(y0 + r) * W + x] = -a0[r];

    const int H1 = H >> 1, W1 = W >> 1, H2 = H >> 2, W2 = W >> 2;
    if ((x & 1) == 0 && x + 1 < W) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i1 = 2 * rq + h;
            if (2 * i1 + 1 < H)
                g1[((((size_t)b * G + g) * D + d) * H1 + i1) * W1 + (x >> 1)] = -a1[h];
        }
    }
    if ((x & 3) == 0 && x + 3 < W && y0 + 3 < H)
        g2[((((size_t)b * G + g) * D + d) * H2 + rq) * W2 + (x >> 2)] = -a2;
}

__device__ __forceinline__ float bilerp_plane(const float* __restrict__ p, int Wp, const LerpIdx& iy,
                                              const LerpIdx& ix) {
    const float t0 = ix.w0 * __ldg(p + (size_t)iy.i0 * Wp + ix.i0) + ix.w1 * __ldg(p + (size_t)iy.i0 * Wp + ix.i1);
    const float t1 = ix.w0 * __ldg(p + (size_t)iy.i1 * Wp + ix.i0) + ix.w1 * __ldg(p + (size_t)iy.i1 * Wp + ix.i1);
    return iy.w0 * t0 + iy.w1 * t1;
}

// out planes [base+G+g] and [base+2G+g] <- bilinear(G1), bilinear(G2)   (block_cost.py:74)
__global__ void __launch_bounds__(256)
block_cost_resize_kernel(const float* __restrict__ g1, const float* __restrict__ g2, float* __restrict__ out,
                         int G, int D, int H, int W, int outC, int base, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % W);
    long long t = idx / W;
    const int y = (int)(t % H);
    t /= H;
    const int d = (int)(t % D);
    t /= D;
    const int g = (int)(t % G);
    const int b = (int)(t / G);
    const int H1 = H >> 1, W1 = W >> 1, H2 = H >> 2, W2 = W >> 2;
    const size_t HW = (size_t)H * W;
    {
        const LerpIdx iy = ac_index(ac_scale(H1, H), y, H1), ix = ac_index(ac_scale(W1, W), x, W1);
        const float* p = g1 + (((size_t)b * G + g) * D + d) * (size_t)H1 * W1;
        out[(((size_t)b * outC + base + G + g) * D + d) * HW + (size_t)y * W + x] = bilerp_plane(p, W1, iy, ix);
    }
    {
        const LerpIdx iy = ac_index(ac_scale(H2, H), y, H2), ix = ac_index(ac_scale(W2, W), x, W2);
        const float* p = g2 + (((size_t)b * G + g) * D + d) * (size_t)H2 * W2;
        out[(((size_t)b * outC + base + 2 * G + g) * D + d) * HW + (size_t)y * W + x] = bilerp_plane(p, W2, iy, ix);
    }
}

static int block_cost_launch(bool warp, const float* L, const float* R, const float* smp, float* out,
                             float* scratch, int B, int C, int H, int W, int D, cudaStream_t st) {
    TS_REQUIRE(L && R && out && scratch, "block_cost: null pointer");
    TS_REQUIRE(!warp || smp, "block_cost_warp: null samples");
    TS_REQUIRE(B > 0 && D > 0 && C > 0 && C % 8 == 0, "block_cost: C=%d must be a positive multiple of 8 (B=%d D=%d)", C, B, D);
    TS_REQUIRE(H >= 4 && W >= 4, "block_cost: H=%d W=%d must be >= 4 for the three pooled scales", H, W);
    const int G = C / 8;
    TS_REQUIRE((long long)B * D * G <= 65535, "block_cost: B*D*C/8 = %lld exceeds grid.z", (long long)B * D * G);
    const int H1 = H / 2, W1 = W / 2;
    float* g1 = scratch;
    float* g2 = scratch + (size_t)B * G * D * H1 * W1;
    dim3 grid(cdiv(W, 32), cdiv(H, 16), B * D * G), block(32, 4);
    if (warp)
        block_cost_main_kernel<true><<<grid, block, 0, st>>>(L, R, smp, out, g1, g2, C, H, W, D);
    else
        block_cost_main_kernel<false><<<grid, block, 0, st>>>(L, R, nullptr, out, g1, g2, C, H, W, D);
    int rc = check_launch("block_cost_main");
    if (rc) return rc;
    const int outC = (warp ? 2 * C : C) + 3 * G;
    const long long total = (long long)B * G * D * H * W;
    block_cost_resize_kernel<<<(unsigned)cdivll(total, 256), 256, 0, st>>>(g1, g2, out, G, D, H, W, outC,
                                                                          warp ? 2 * C : C, total);
    return check_launch("block_cost_resize");
}

}  // namespace tstereo

extern "C" {

long long tstereo_block_cost_scratch_floats(int B, int C, int H, int W, int D) {
    const long long G = C / 8;
    return (long long)B * G * D * ((long long)(H / 2) * (W / 2) + (long long)(H / 4) * (W / 4));
}

int tstereo_block_cost_shift(const float* left, const float* right, float* out, float* scratch,
                             int B, int C, int H, int W, int D, void* stream) {
    return tstereo::block_cost_launch(false, left, right, nullptr, out, scratch, B, C, H, W, D, (cudaStream_t)stream);
}

int tstereo_block_cost_warp(const float* left, const float* right, const float* samples, float* out,
                            float* scratch, int B, int C, int H, int W, int S, void* stream) {
    return tstereo::block_cost_launch(true, left, right, samples, out, scratch, B, C, H, W, S, (cudaStream_t)stream);
}

}  // extern "C"
