// Cost-volume construction.
//
// ref: architecture/modeling/aggregation/utils/block_cost.py:16-83 — int branch (:34-45),
//      tensor branch (:47-58) via layers/inverse_warp_3d.py:4-58, pooled group-wise terms
//      (:6-13, :64-78).
//
// HBM-bound streaming kernel: the output is ~6x the input bytes (C2 precise: 34 MB in, 198 MB out).
//
//  * main kernel.  thread = FOUR consecutive pixels of a row, one group of 8 channels, one
//    disparity candidate.  A warp is 8 threads (32 px = 128 B) x 4 rows, a 128-thread CTA covers
//    64 px x 8 rows.  The left features and every output plane move as 16 B vectors (one 128 B line
//    per warp row); only the right-feature taps are scalar gathers.  Per pixel the warp coordinate /
//    tap weights are computed once and reused for the 8 channels.
//    avg_pool(L) - avg_pool(R_d) == avg_pool(L - R_d), so the 2x2 and 4x4 pooled differences are
//    in-thread sums along x plus warp shuffles along y (xor 8, xor 16) — no second pass over the
//    features, no shared memory.  The thread writes L, R_d (or -(L-R_d)^2), the full-resolution
//    group term g0 and the small pooled terms G1, G2 (scratch, L2 resident).
//  * resize kernel: bilinear align_corners up-sampling of G1, G2 into the last 2*C/8 planes,
//    four output columns per thread.
#include "common.cuh"
#include "cost_coord.cuh"

namespace tstereo {

// Predicated read-only loads / stores as single instructions (no branch, so ptxas can hoist the next
// channel's loads above the current channel's arithmetic and stores).
__device__ __forceinline__ float ldg_if(const float* p, bool pred) {
    float v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
        : "=f"(v) : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ float4 ldg4_if(const float* p, bool pred) {
    float4 v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
        "mov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t@q ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ void stg4_if(float* p, float a, float b, float c, float d, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q st.global.v4.f32 [%0], {%1,%2,%3,%4};\n\t}"
                 :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "r"((int)pred));
}

// VEC: W % 4 == 0 and 16 B aligned bases -> float4 loads / stores; otherwise scalar with tail guards.
//
// Taps outside the image are read from a clamped in-image address with weight 0 (no predicate per
// load), so the arithmetic on finite features is unchanged.  grid_sample's y coordinate
// (inverse_warp_3d.py:46 + grid_sampler_unnormalize) lands up to ~1e-6 px off the integer row on
// ~20 % of the rows, which would blend two rows with weights (1-eps, eps); the kernel samples the
// nearer row only (deviation <= 2e-6 * |R|, far below the fp32 noise of the following contraction).
// GONLY: only the three group-wise terms are produced, into a compact [B, 3G, D, H, W] volume (`out`): the form the
// fused cost -> first-conv path uses (the L / R_d / -(L-R_d)^2 planes are rebuilt by that conv's producer).
// SOUT (with GONLY, shift form): the C cost planes are ALSO produced, as the S-format (fp16 hi / lo split, chunk = this
// thread's group of 8 channels: the thread holds all 8 channels of its 4 pixels) that the first conv stages by TMA — the
// materialised coarse volume in the layout its only consumer reads, same bytes as fp32.
template <bool WARP, bool VEC, bool GONLY, bool SOUT = false>
__global__ void __launch_bounds__(128)
block_cost_main_kernel(const float* __restrict__ L, const float* __restrict__ R,
                       const float* __restrict__ smp, float* __restrict__ out,
                       float* __restrict__ g1, float* __restrict__ g2,
                       int C, int H, int W, int D, unsigned short* __restrict__ so = nullptr, long long ssB = 0,
                       long long ssD = 0, long long ssP = 0, long long ssC8 = 0) {
    pdl_sync();
    const int G = C >> 3;
    int z = blockIdx.z;
    const int d = z % D;
    z /= D;
    const int g = z % G;
    const int b = z / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = (blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7)) * 4;       // first of the 4 pixels
    const int y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const int HW = H * W;
    const int outC = GONLY ? 3 * G : (WARP ? 2 * C : C) + 3 * G;
    const bool rowin = y < H;
    bool pin[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pin[k] = rowin && (x + k < W);
    const int yc = min(y, H - 1);                 // threads outside the image read a valid row, store nothing

    // per pixel: offset of tap 0 inside the channel plane and the two tap weights
    int off[4];
    float wa[4], wb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        off[k] = yc * W;
        wa[k] = 0.f;
        wb[k] = 0.f;
    }
    if (WARP) {
        // same op sequence as inverse_warp_3d.py:40-47 + ATen grid_sampler_unnormalize (cost_coord.cuh)
        const int yn = warp_row(yc, H);
        const float* sp = smp + ((size_t)(b * D + d) * H + yc) * W;
        float dsp[4];
        if (VEC) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(sp + min(x, W - 4)));
            dsp[0] = t.x; dsp[1] = t.y; dsp[2] = t.z; dsp[3] = t.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) dsp[k] = __ldg(sp + min(x + k, W - 1));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int xa;
            warp_col(x + k, dsp[k], W, pin[k], xa, wa[k], wb[k]);
            off[k] = yn * W + xa;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xs = x + k - d;
            if (pin[k] && xs >= 0) wa[k] = 1.f;
            off[k] = yc * W + min(max(xs, 0), W - 1);
        }
    }

    float a0[4] = {0.f, 0.f, 0.f, 0.f};
    float a1[2] = {0.f, 0.f};
    float a2 = 0.f;

    const int pix = y * W + x;                                                   // store position
    const int lpix = yc * W + (VEC ? min(x, W - 4) : 0);                          // always-valid load position
    const float* Lp = L + ((size_t)b * C + g * 8) * HW + lpix;
    const float* Rp = R + ((size_t)b * C + g * 8) * HW;
    float* o1 = out + (((size_t)b * outC + g * 8) * D + d) * HW + pix;           // first half, plane (ch, d)
    const size_t second = (size_t)C * D * HW;                                    // offset of the R half (WARP)
    const size_t chs = (size_t)D * HW;                                            // channel stride in `out`
    const float lmask = pin[0] ? 1.f : 0.f;

#ifndef TS_BC_UNROLL
#define TS_BC_UNROLL 8
#endif
    constexpr int kUnroll = TS_BC_UNROLL;     // channels whose loads are in flight together
    [[maybe_unused]] float sv[SOUT ? 8 : 1][4];
#pragma unroll kUnroll
    for (int c = 0; c < 8; ++c) {
        float l[4], rv[4];
        if (VEC) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(Lp));
            l[0] = t.x; l[1] = t.y; l[2] = t.z; l[3] = t.w;
            if (!pin[0]) l[0] = l[1] = l[2] = l[3] = 0.f;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) l[k] = pin[k] ? __ldg(Lp + min(x + k, W - 1)) : 0.f;
        }
        if (WARP) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float ra = __ldg(Rp + off[k]);
                const float rb = __ldg(Rp + off[k] + 1);
                rv[k] = fmaf(rb, wb[k], __fmul_rn(ra, wa[k]));
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) rv[k] = __fmul_rn(__ldg(Rp + off[k]), wa[k]);
        }
        float e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            e[k] = l[k] - rv[k];                 // 0 outside the image (l = 0, weights = 0)
            a0[k] = fmaf(e[k], e[k], a0[k]);
        }
        if constexpr (SOUT) {
#pragma unroll
            for (int k = 0; k < 4; ++k) sv[c][k] = WARP ? rv[k] : -(e[k] * e[k]);
        }
        if (GONLY) {
        } else if (VEC) {
            if (pin[0]) {
                if (WARP) {
                    *reinterpret_cast<float4*>(o1) = make_float4(l[0], l[1], l[2], l[3]);
                    *reinterpret_cast<float4*>(o1 + second) = make_float4(rv[0], rv[1], rv[2], rv[3]);
                } else {
                    *reinterpret_cast<float4*>(o1) = make_float4(-(e[0] * e[0]), -(e[1] * e[1]), -(e[2] * e[2]), -(e[3] * e[3]));
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (pin[k]) {
                    if (WARP) {
                        o1[k] = l[k];
                        o1[second + k] = rv[k];
                    } else {
                        o1[k] = -(e[k] * e[k]);
                    }
                }
        }
        // 2x2 pooled difference: x pairs in-thread, y pairs across lanes (xor 8); 4x4: + xor 16
        float pa = e[0] + e[1], pb = e[2] + e[3];
        pa += __shfl_xor_sync(0xffffffffu, pa, 8);
        pb += __shfl_xor_sync(0xffffffffu, pb, 8);
        const float ma = pa * 0.25f, mb = pb * 0.25f;
        a1[0] = fmaf(ma, ma, a1[0]);
        a1[1] = fmaf(mb, mb, a1[1]);
        float q = pa + pb;
        q += __shfl_xor_sync(0xffffffffu, q, 16);
        const float mq = q * 0.0625f;
        a2 = fmaf(mq, mq, a2);
        Lp += HW;
        Rp += HW;
        o1 += chs;
    }
    (void)lmask;
    if constexpr (SOUT) {
        unsigned short* sp = so + b * ssB + d * ssD + g * ssC8 + (long long)pix * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hi[j] = pack_h2(sv[2 * j][k], sv[2 * j + 1][k]);
                const float2 hf = unpack_h2(hi[j]);
                lo[j] = pack_h2(sv[2 * j][k] - hf.x, sv[2 * j + 1][k] - hf.y);
            }
            if (pin[k]) {
                stg128(sp + k * 8, hi[0], hi[1], hi[2], hi[3]);
                stg128(sp + k * 8 + ssP, lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }

    const int base = GONLY ? 0 : (WARP ? 2 * C : C);
    const int H1 = H >> 1, W1 = W >> 1, H2 = H >> 2, W2 = W >> 2;
    float* og = out + (((size_t)b * outC + base + g) * D + d) * HW + pix;
    if (VEC) {
        if (pin[0]) *reinterpret_cast<float4*>(og) = make_float4(-a0[0], -a0[1], -a0[2], -a0[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (pin[k]) og[k] = -a0[k];
    }
    const size_t pl = ((size_t)b * G + g) * D + d;
    if ((y & 1) == 0 && y + 1 < H) {
        float* p1 = g1 + (pl * H1 + (y >> 1)) * W1 + (x >> 1);
        if (x + 1 < W) p1[0] = -a1[0];
        if (x + 3 < W) p1[1] = -a1[1];
    }
    if ((y & 3) == 0 && y + 3 < H && x + 3 < W) g2[(pl * H2 + (y >> 2)) * W2 + (x >> 2)] = -a2;
}

// out planes [base+G+g] and [base+2G+g] <- bilinear_align_corners(G1), (G2)   (block_cost.py:74)
// thread = 4 consecutive output columns of one row; the align_corners index / weight math is done
// once per thread and reused for all D candidate planes of the (b, g) group.
__global__ void __launch_bounds__(128)
block_cost_resize_kernel(const float* __restrict__ g1, const float* __restrict__ g2, float* __restrict__ out,
                         int G, int D, int H, int W, int outC, int base,
                         float sy1, float sx1, float sy2, float sx2) {
    pdl_sync();
    // blockDim = (TX, 128 / TX): TX = 32 | 64 | 128 threads along x (4 columns each), the rest along y
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x4 >= W || y >= H) return;
    const int g = blockIdx.z % G;
    const int b = blockIdx.z / G;
    const int H1 = H >> 1, W1 = W >> 1, H2 = H >> 2, W2 = W >> 2;
    const int HW = H * W, P1 = H1 * W1, P2 = H2 * W2;
    const LerpIdx iy1 = ac_index(sy1, y, H1), iy2 = ac_index(sy2, y, H2);
    // per column: offsets of the 4 source texels (relative to the plane) and the x weight
    int o1a[4], o1b[4], o2a[4], o2b[4];       // row i0: (a, a + d1), row i1: (b, b + d1)
    int d1[4], d2[4];
    float wx1[4], wx2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = min(x4 + k, W - 1);
        const LerpIdx i1 = ac_index(sx1, x, W1), i2 = ac_index(sx2, x, W2);
        o1a[k] = iy1.i0 * W1 + i1.i0;
        o1b[k] = iy1.i1 * W1 + i1.i0;
        d1[k] = i1.i1 - i1.i0;
        wx1[k] = i1.w1;
        o2a[k] = iy2.i0 * W2 + i2.i0;
        o2b[k] = iy2.i1 * W2 + i2.i0;
        d2[k] = i2.i1 - i2.i0;
        wx2[k] = i2.w1;
    }
    const size_t pl = ((size_t)b * G + g) * D;
    const float* p1 = g1 + pl * P1;
    const float* p2 = g2 + pl * P2;
    float* q1 = out + (((size_t)b * outC + base + G + g) * D) * HW + (size_t)y * W + x4;
    float* q2 = out + (((size_t)b * outC + base + 2 * G + g) * D) * HW + (size_t)y * W + x4;
    const bool vec = (W & 3) == 0;
    for (int d = 0; d < D; ++d) {
        float v1[4], v2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            {
                const float a0 = __ldg(p1 + o1a[k]), a1 = __ldg(p1 + o1a[k] + d1[k]);
                const float b0 = __ldg(p1 + o1b[k]), b1 = __ldg(p1 + o1b[k] + d1[k]);
                const float t0 = (1.0f - wx1[k]) * a0 + wx1[k] * a1;
                const float t1 = (1.0f - wx1[k]) * b0 + wx1[k] * b1;
                v1[k] = iy1.w0 * t0 + iy1.w1 * t1;
            }
            {
                const float a0 = __ldg(p2 + o2a[k]), a1 = __ldg(p2 + o2a[k] + d2[k]);
                const float b0 = __ldg(p2 + o2b[k]), b1 = __ldg(p2 + o2b[k] + d2[k]);
                const float t0 = (1.0f - wx2[k]) * a0 + wx2[k] * a1;
                const float t1 = (1.0f - wx2[k]) * b0 + wx2[k] * b1;
                v2[k] = iy2.w0 * t0 + iy2.w1 * t1;
            }
        }
        if (vec) {
            *reinterpret_cast<float4*>(q1) = make_float4(v1[0], v1[1], v1[2], v1[3]);
            *reinterpret_cast<float4*>(q2) = make_float4(v2[0], v2[1], v2[2], v2[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (x4 + k < W) {
                    q1[k] = v1[k];
                    q2[k] = v2[k];
                }
        }
        p1 += P1;
        p2 += P2;
        q1 += HW;
        q2 += HW;
    }
}

static inline float host_ac_scale(int in_size, int out_size) {
    return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.0f;   // IEEE fp32 divide == __fdiv_rn
}

static int block_cost_launch(bool warp, bool gonly, const float* L, const float* R, const float* smp, float* out,
                             float* scratch, int B, int C, int H, int W, int D, cudaStream_t st, const tstereo_split* sout = nullptr) {
    TS_REQUIRE(L && R && out && scratch, "block_cost: null pointer");
    TS_REQUIRE(!warp || smp, "block_cost_warp: null samples");
    TS_REQUIRE(B > 0 && D > 0 && C > 0 && C % 8 == 0, "block_cost: C=%d must be a positive multiple of 8 (B=%d D=%d)", C, B, D);
    TS_REQUIRE(H >= 4 && W >= 4, "block_cost: H=%d W=%d must be >= 4 for the three pooled scales", H, W);
    const int G = C / 8;
    TS_REQUIRE((long long)B * G * D <= 65535 && H <= 65535 * 8, "block_cost: B*C/8*D = %lld exceeds grid.z", (long long)B * G * D);
    TS_REQUIRE((long long)H * W < (1ll << 30), "block_cost: plane too large");
    const int H1 = H / 2, W1 = W / 2;
    float* g1 = scratch;
    float* g2 = scratch + (size_t)B * G * D * H1 * W1;
    const bool vec = (W % 4 == 0) && (((size_t)L | (size_t)out | (size_t)(smp ? smp : L)) & 15) == 0;
    dim3 grid(cdiv(W, 64), cdiv(H, 8), B * G * D);
#define TS_BC(WP, VC) (gonly ? launch_k(block_cost_main_kernel<WP, VC, true>, dim3(grid), dim3(128), 0, st, L, R, smp, out, g1, g2, C, H, W, D, \
                                        (unsigned short*)nullptr, 0ll, 0ll, 0ll, 0ll)                                                        \
                             : launch_k(block_cost_main_kernel<WP, VC, false>, dim3(grid), dim3(128), 0, st, L, R, smp, out, g1, g2, C, H, W, D, \
                                        (unsigned short*)nullptr, 0ll, 0ll, 0ll, 0ll))
    if (sout) {
        TS_REQUIRE(!warp && gonly, "block_cost: the S-format cost planes belong to the shift form with compact group terms");
        TS_REQUIRE(sout->ptr && sout->parts == 2 && sout->C8 >= G, "block_cost: bad S-format output");
        TS_REQUIRE((((size_t)sout->ptr) & 15) == 0 && (sout->sB & 7) == 0 && (sout->sD & 7) == 0 && (sout->sP & 7) == 0 && (sout->sC8 & 7) == 0,
                   "block_cost: S-format output must be 16-byte aligned");
        unsigned short* sp = (unsigned short*)sout->ptr;
        if (vec) launch_k(block_cost_main_kernel<false, true, true, true>, dim3(grid), dim3(128), 0, st, L, R, smp, out, g1, g2, C, H, W, D, sp,
                          sout->sB, sout->sD, sout->sP, sout->sC8);
        else launch_k(block_cost_main_kernel<false, false, true, true>, dim3(grid), dim3(128), 0, st, L, R, smp, out, g1, g2, C, H, W, D, sp,
                      sout->sB, sout->sD, sout->sP, sout->sC8);
    } else if (warp) {
        if (vec) TS_BC(true, true); else TS_BC(true, false);
    } else {
        if (vec) TS_BC(false, true); else TS_BC(false, false);
    }
#undef TS_BC
    int rc = check_launch("block_cost_main");
    if (rc) return rc;
    const int base = gonly ? 0 : (warp ? 2 * C : C);
    const int outC = base + 3 * G;
    const int tx = cdiv(W, 4) <= 32 ? 32 : (cdiv(W, 4) <= 64 ? 64 : 128);
    dim3 rblock(tx, 128 / tx);
    dim3 rgrid(cdiv(cdiv(W, 4), tx), cdiv(H, (int)rblock.y), B * G);
    launch_k(block_cost_resize_kernel, dim3(rgrid), dim3(rblock), 0, st, g1, g2, out, G, D, H, W, outC, base,
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
    return tstereo::block_cost_launch(false, false, left, right, nullptr, out, scratch, B, C, H, W, D, (cudaStream_t)stream);
}

int tstereo_block_cost_warp(const float* left, const float* right, const float* samples, float* out,
                            float* scratch, int B, int C, int H, int W, int S, void* stream) {
    return tstereo::block_cost_launch(true, false, left, right, samples, out, scratch, B, C, H, W, S, (cudaStream_t)stream);
}

/* group-wise terms only (compact [B, 3C/8, D, H, W]): the side input of the fused cost -> first-conv path */
int tstereo_group_cost_shift(const float* left, const float* right, float* gvol, float* scratch,
                             int B, int C, int H, int W, int D, void* stream) {
    return tstereo::block_cost_launch(false, true, left, right, nullptr, gvol, scratch, B, C, H, W, D, (cudaStream_t)stream);
}

/* shift volume for a TMA-fed first conv: the C cost planes -(L - R_d)^2 as S-format chunks [0, C/8) of `sout` + the compact
 * group terms `gvol` (to be appended as chunks [C/8, C/8 + 3C/64) with tstereo_split_pack) */
int tstereo_block_cost_shift_s(const float* left, const float* right, const tstereo_split* sout, float* gvol, float* scratch,
                               int B, int C, int H, int W, int D, void* stream) {
    TS_REQUIRE(sout, "block_cost_shift_s: null S-format output");
    return tstereo::block_cost_launch(false, true, left, right, nullptr, gvol, scratch, B, C, H, W, D, (cudaStream_t)stream, sout);
}

int tstereo_group_cost_warp(const float* left, const float* right, const float* samples, float* gvol,
                            float* scratch, int B, int C, int H, int W, int S, void* stream) {
    return tstereo::block_cost_launch(true, true, left, right, samples, gvol, scratch, B, C, H, W, S, (cudaStream_t)stream);
}

}  // extern "C"
