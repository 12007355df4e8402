// Cost-volume construction.
//
// ref: architecture/modeling/aggregation/utils/block_cost.py:16-83 — int branch (:34-45),
//      tensor branch (:47-58) via layers/inverse_warp_3d.py:4-58, pooled group-wise terms
//      (:6-13, :64-78).
//
// HBM-bound streaming kernel: the output is ~6x the input bytes (C2 precise: 34 MB in, 198 MB out).
//
//  * main kernel.  thread = ONE pixel, one group of 8 channels, SC disparity candidates.
//    lane layout inside a warp is 8 columns x 4 rows, a 256-thread CTA covers 32 columns x 8 rows,
//    so every global access of a warp is four fully used 32 B sectors and a CTA writes whole 128 B
//    lines.  Per pixel the warp coordinate / tap weights are computed once per candidate and reused
//    for the 8 channels; the left feature is loaded once and reused for the SC candidates.
//    avg_pool(L) - avg_pool(R_d) == avg_pool(L - R_d), so the 2x2 and 4x4 pooled differences are
//    warp-shuffle reductions of the per-pixel difference (xor 1, 8 | xor 2, 16) — no second pass over
//    the features, no shared memory.  The thread writes L, R_d (or -(L-R_d)^2), the full-resolution
//    group term g0 and the small pooled terms G1, G2 (scratch, L2 resident).
//  * resize kernel: bilinear align_corners up-sampling of G1, G2 into the last 2*C/8 planes,
//    four output columns per thread.
#include "common.cuh"

namespace tstereo {

template <bool WARP, int SC>
__global__ void __launch_bounds__(256)
block_cost_main_kernel(const float* __restrict__ L, const float* __restrict__ R,
                       const float* __restrict__ smp, float* __restrict__ out,
                       float* __restrict__ g1, float* __restrict__ g2,
                       int C, int H, int W, int D, int nchunk) {
    const int G = C >> 3;
    int z = blockIdx.z;
    const int chunk = z % nchunk;
    z /= nchunk;
    const int g = z % G;
    const int b = z / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    const size_t HW = (size_t)H * W;
    const int outC = (WARP ? 2 * C : C) + 3 * G;
    const bool pin = (x < W) && (y < H);
    const int d0 = chunk * SC;

    // per-candidate horizontal taps (x0 = -2: nothing to sample)
    int x0[SC];
    float w0[SC], w1[SC];
    // vertical taps (warp branch only; same for every candidate)
    int ylo = y;
    float wy0 = 1.f, wy1 = 0.f;
#pragma unroll
    for (int s = 0; s < SC; ++s) {
        x0[s] = -2;
        w0[s] = 0.f;
        w1[s] = 0.f;
    }
    if (pin) {
        if (WARP) {
            // The y coordinate goes through the same normalise / un-normalise round trip
            // (inverse_warp_3d.py:46, grid_sampler_unnormalize); for some (H, y) it lands a few
            // 1e-6 px off the integer, which blends two rows.
            const float Hm1 = (float)(H - 1), Wm1 = (float)(W - 1);
            const float gyn = __fsub_rn(__fmul_rn(__fdiv_rn((float)y, Hm1), 2.0f), 1.0f);
            const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gyn, 1.0f), 2.0f), Hm1);
            const float fy = floorf(iy);
            ylo = (int)fy;
            wy0 = __fsub_rn(fy + 1.0f, iy);
            wy1 = __fsub_rn(iy, fy);
#pragma unroll
            for (int s = 0; s < SC; ++s) {
                if (d0 + s < D) {
                    // same op sequence as inverse_warp_3d.py:40-47 + ATen grid_sampler_unnormalize
                    const float dsp = __ldg(smp + ((size_t)(b * D + d0 + s) * H + y) * W + x);
                    const float gx = __fadd_rn((float)x, -dsp);
                    const float gn = __fsub_rn(__fmul_rn(__fdiv_rn(gx, Wm1), 2.0f), 1.0f);
                    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gn, 1.0f), 2.0f), Wm1);
                    const float fx = floorf(ix);
                    if (fx >= -1.0f && fx <= Wm1) {
                        x0[s] = (int)fx;
                        w0[s] = __fsub_rn(fx + 1.0f, ix);
                        w1[s] = __fsub_rn(ix, fx);
                    }
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < SC; ++s)
                if (d0 + s < D && x - (d0 + s) >= 0) x0[s] = x - (d0 + s);
        }
    }
    // rows to blend (warp branch): r = 0 -> (ylo, wy0), r = 1 -> (ylo + 1, wy1); zero-weight or
    // out-of-image rows are skipped exactly like grid_sample's zeros padding
    int rbeg = 0, rend = 1;
    if (WARP) {
        rbeg = (ylo >= 0 && wy0 != 0.f) ? 0 : 1;
        rend = (ylo + 1 < H && wy1 != 0.f) ? 2 : 1;
    }

    float a0[SC], a1[SC], a2[SC];
#pragma unroll
    for (int s = 0; s < SC; ++s) a0[s] = a1[s] = a2[s] = 0.f;

    const size_t pix = (size_t)y * W + x;
    const float* Lp = L + ((size_t)b * C + g * 8) * HW + pix;
    const float* Rb = R + ((size_t)b * C + g * 8) * HW;
    float* o1 = out + (((size_t)b * outC + g * 8) * D + d0) * HW + pix;   // first half, plane (ch, d0)
    const size_t second = (size_t)C * D * HW;                             // offset of the R half (WARP)
    const size_t chs = (size_t)D * HW;                                     // channel stride in `out`

#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
        const float l = pin ? __ldg(Lp + (size_t)c * HW) : 0.f;
        const float* Rp = Rb + (size_t)c * HW;
        float rv[SC];
#pragma unroll
        for (int s = 0; s < SC; ++s) rv[s] = 0.f;
        if (WARP) {
#pragma unroll 1
            for (int r = rbeg; r < rend; ++r) {
                const float* rr = Rp + (ylo + r) * W;
                const float wy = r ? wy1 : wy0;
#pragma unroll
                for (int s = 0; s < SC; ++s) {
                    const bool v0 = (unsigned)x0[s] < (unsigned)W, v1 = (unsigned)(x0[s] + 1) < (unsigned)W;
                    const float ra = v0 ? __ldg(rr + x0[s]) : 0.f;
                    const float rb = v1 ? __ldg(rr + x0[s] + 1) : 0.f;
                    rv[s] = fmaf(wy, fmaf(rb, w1[s], __fmul_rn(ra, w0[s])), rv[s]);
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < SC; ++s)
                if (x0[s] >= 0) rv[s] = __ldg(Rp + y * W + x0[s]);
        }
#pragma unroll
        for (int s = 0; s < SC; ++s) {
            const bool live = pin && (d0 + s < D);
            float* op = o1 + (size_t)c * chs + (size_t)s * HW;
            const float e = live ? (l - rv[s]) : 0.f;
            if (live) {
                if (WARP) {
                    op[0] = l;
                    op[second] = rv[s];
                } else {
                    op[0] = -(e * e);
                }
            }
            a0[s] = fmaf(e, e, a0[s]);
            float s2 = e + __shfl_xor_sync(0xffffffffu, e, 1);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
            const float m1 = s2 * 0.25f;
            a1[s] = fmaf(m1, m1, a1[s]);
            float s4 = s2 + __shfl_xor_sync(0xffffffffu, s2, 2);
            s4 += __shfl_xor_sync(0xffffffffu, s4, 16);
            const float m2 = s4 * 0.0625f;
            a2[s] = fmaf(m2, m2, a2[s]);
        }
    }

    const int base = WARP ? 2 * C : C;
    const int H1 = H >> 1, W1 = W >> 1, H2 = H >> 2, W2 = W >> 2;
    const bool cell1 = ((x | y) & 1) == 0 && x + 1 < W && y + 1 < H;
    const bool cell2 = ((x | y) & 3) == 0 && x + 3 < W && y + 3 < H;
#pragma unroll
    for (int s = 0; s < SC; ++s) {
        const int d = d0 + s;
        if (d >= D) break;
        if (pin) out[(((size_t)b * outC + base + g) * D + d) * HW + pix] = -a0[s];
        const size_t pl = ((size_t)b * G + g) * D + d;
        if (cell1) g1[(pl * H1 + (y >> 1)) * W1 + (x >> 1)] = -a1[s];
        if (cell2) g2[(pl * H2 + (y >> 2)) * W2 + (x >> 2)] = -a2[s];
    }
}

// out planes [base+G+g] and [base+2G+g] <- bilinear_align_corners(G1), (G2)   (block_cost.py:74)
// thread = 4 consecutive output columns of one row of one (b, g, d) plane.
__global__ void __launch_bounds__(128)
block_cost_resize_kernel(const float* __restrict__ g1, const float* __restrict__ g2, float* __restrict__ out,
                         int G, int D, int H, int W, int outC, int base,
                         float sy1, float sx1, float sy2, float sx2) {
    const int x4 = (blockIdx.x * 128 + threadIdx.x) * 4;
    if (x4 >= W) return;
    const int y = blockIdx.y;
    int z = blockIdx.z;
    const int d = z % D;
    z /= D;
    const int g = z % G;
    const int b = z / G;
    const int H1 = H >> 1, W1 = W >> 1, H2 = H >> 2, W2 = W >> 2;
    const size_t HW = (size_t)H * W;
    const size_t pl = ((size_t)b * G + g) * D + d;
    const float* p1 = g1 + pl * (size_t)H1 * W1;
    const float* p2 = g2 + pl * (size_t)H2 * W2;
    float* o1 = out + (((size_t)b * outC + base + G + g) * D + d) * HW + (size_t)y * W + x4;
    float* o2 = out + (((size_t)b * outC + base + 2 * G + g) * D + d) * HW + (size_t)y * W + x4;
    const LerpIdx iy1 = ac_index(sy1, y, H1), iy2 = ac_index(sy2, y, H2);
    const float* r10 = p1 + (size_t)iy1.i0 * W1;
    const float* r11 = p1 + (size_t)iy1.i1 * W1;
    const float* r20 = p2 + (size_t)iy2.i0 * W2;
    const float* r21 = p2 + (size_t)iy2.i1 * W2;
    float v1[4], v2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = min(x4 + k, W - 1);
        const LerpIdx ix1 = ac_index(sx1, x, W1), ix2 = ac_index(sx2, x, W2);
        {
            const float t0 = ix1.w0 * __ldg(r10 + ix1.i0) + ix1.w1 * __ldg(r10 + ix1.i1);
            const float t1 = ix1.w0 * __ldg(r11 + ix1.i0) + ix1.w1 * __ldg(r11 + ix1.i1);
            v1[k] = iy1.w0 * t0 + iy1.w1 * t1;
        }
        {
            const float t0 = ix2.w0 * __ldg(r20 + ix2.i0) + ix2.w1 * __ldg(r20 + ix2.i1);
            const float t1 = ix2.w0 * __ldg(r21 + ix2.i0) + ix2.w1 * __ldg(r21 + ix2.i1);
            v2[k] = iy2.w0 * t0 + iy2.w1 * t1;
        }
    }
    if ((W & 3) == 0) {
        *reinterpret_cast<float4*>(o1) = make_float4(v1[0], v1[1], v1[2], v1[3]);
        *reinterpret_cast<float4*>(o2) = make_float4(v2[0], v2[1], v2[2], v2[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (x4 + k < W) {
                o1[k] = v1[k];
                o2[k] = v2[k];
            }
    }
}

static inline float host_ac_scale(int in_size, int out_size) {
    return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.0f;   // IEEE fp32 divide == __fdiv_rn
}

template <bool WARP, int SC>
static void launch_main(const float* L, const float* R, const float* smp, float* out, float* g1, float* g2, int B,
                        int C, int H, int W, int D, cudaStream_t st) {
    const int nchunk = cdiv(D, SC);
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B * (C / 8) * nchunk);
    block_cost_main_kernel<WARP, SC><<<grid, 256, 0, st>>>(L, R, smp, out, g1, g2, C, H, W, D, nchunk);
}

static int block_cost_launch(bool warp, const float* L, const float* R, const float* smp, float* out,
                             float* scratch, int B, int C, int H, int W, int D, cudaStream_t st) {
    TS_REQUIRE(L && R && out && scratch, "block_cost: null pointer");
    TS_REQUIRE(!warp || smp, "block_cost_warp: null samples");
    TS_REQUIRE(B > 0 && D > 0 && C > 0 && C % 8 == 0, "block_cost: C=%d must be a positive multiple of 8 (B=%d D=%d)", C, B, D);
    TS_REQUIRE(H >= 4 && W >= 4, "block_cost: H=%d W=%d must be >= 4 for the three pooled scales", H, W);
    const int G = C / 8;
    // candidates per thread: the chunk size that wastes the fewest slots (5 -> 5, 8 -> 4+4, 12 -> 6+6, 16 -> 4x4, 20 -> 4x5)
    int SC = 4, waste = cdiv(D, 4) * 4 - D;
    for (int c = 5; c <= 6; ++c) {
        const int w = cdiv(D, c) * c - D;
        if (w < waste || (w == waste && cdiv(D, c) < cdiv(D, SC))) {
            SC = c;
            waste = w;
        }
    }
    TS_REQUIRE((long long)B * G * cdiv(D, SC) <= 65535, "block_cost: B*C/8*chunks = %lld exceeds grid.z",
               (long long)B * G * cdiv(D, SC));
    TS_REQUIRE((long long)B * G * D <= 65535 && H <= 65535, "block_cost: B*C/8*D = %lld exceeds grid.z", (long long)B * G * D);
    const int H1 = H / 2, W1 = W / 2;
    float* g1 = scratch;
    float* g2 = scratch + (size_t)B * G * D * H1 * W1;
#define TS_BC(WP, S_) launch_main<WP, S_>(L, R, smp, out, g1, g2, B, C, H, W, D, st)
    if (warp) {
        if (SC == 4) TS_BC(true, 4); else if (SC == 5) TS_BC(true, 5); else TS_BC(true, 6);
    } else {
        if (SC == 4) TS_BC(false, 4); else if (SC == 5) TS_BC(false, 5); else TS_BC(false, 6);
    }
#undef TS_BC
    int rc = check_launch("block_cost_main");
    if (rc) return rc;
    const int outC = (warp ? 2 * C : C) + 3 * G;
    dim3 rgrid(cdiv(cdiv(W, 4), 128), H, B * G * D);
    block_cost_resize_kernel<<<rgrid, 128, 0, st>>>(g1, g2, out, G, D, H, W, outC, warp ? 2 * C : C,
                                                    host_ac_scale(H / 2, H), host_ac_scale(W / 2, W),
                                                    host_ac_scale(H / 4, H), host_ac_scale(W / 4, W));
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
