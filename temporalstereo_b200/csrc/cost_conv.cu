// Fused cost volume -> first (1,3,3) convolution of a level (SURVEY.md §8d "fused path", §7 step 4).
//
// ref: aggregation/utils/block_cost.py:34-58, 64-81 (the cost volume) feeding the first half of
//      aggregation/TemporalStereo/module.py:111-147 (DepthwiseConv3D: (1,3,3) conv -> BN -> SiLU) through
//      coarse.py:82-83, fine.py:102-103, precise.py:88-90.
//
// The raw cost volume [B, 2C + 3C/8, S, H, W] (198 MB per frame at the 1/4 level of a 544x960 pair) is never written:
// the tensor-core convolution's producer (conv_tc2_kernel.cuh, FUSE) rebuilds its K-chunks from the feature maps —
//   warp form : R sampled at x - candidate (bilinear along x), 8 channels per chunk;  the left-feature half of the
//               volume does not depend on the candidate, so its contribution  conv(L, W[:, :C])  is computed ONCE per
//               frame as a plain 2-D convolution and enters here as the epilogue addend `addL`;
//   shift form: -(L - R[x - d])^2 per channel;
// and reads only the 3C/8 group-wise channels (tstereo_group_cost_*, 6.3x smaller than the volume) from memory.
#include "conv_tc2_kernel.cuh"

using namespace tstereo;
using namespace tstereo::tc2;

namespace {

int cost_conv(int fuse, const float* left, const float* right, const float* samples, const float* gvol, const float* addL,
              float* out, long long osB, long long osC, long long osD, const float* wpack, const float* bias, const float* oscale,
              int B, int C, int Cout, int D, int H, int W, int act, int half, cudaStream_t st, const char* what) {
    TS_REQUIRE(right && gvol && out && wpack, "%s: null pointer", what);
    TS_REQUIRE(fuse == 1 ? samples != nullptr : left != nullptr, "%s: missing %s", what, fuse == 1 ? "samples" : "left features");
    TS_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && Cout > 0 && D > 0 && H >= 4 && W >= 4, "%s: bad sizes (C=%d must be a multiple of 8)", what, C);
    TS_REQUIRE((long long)B * D <= 65535, "%s: B*D exceeds grid.y", what);
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "%s: packed weights must be 16-byte aligned", what);
    TS_REQUIRE((long long)C * H * W < (1ll << 31) && (long long)3 * (C / 8) * D * H * W < (1ll << 40), "%s: volume exceeds the offset range", what);
    TS_REQUIRE((long long)D * H * W * 8 < (1ll << 31) && osC >= 0 && osC * 32 < (1ll << 31), "%s: channel strides exceed 32 bits", what);
    const int G3 = 3 * (C / 8);
    tc2::Params p = {};
    p.in = right; p.isB = (long long)C * H * W; p.isC = H * W; p.isD = 0;
    p.left = left; p.smp = samples;
    p.in2 = gvol; p.i2sB = (long long)G3 * D * H * W; p.i2sC = D * H * W; p.i2sD = (long long)H * W; p.C2 = G3;
    p.add = addL; p.asB = (long long)Cout * H * W; p.asC = H * W;
    p.out = out; p.osB = osB; p.osC = (int)osC; p.osD = osD;
    p.wpack = wpack; p.bias = bias; p.oscale = oscale;
    p.H = H; p.W = W; p.D = D; p.Hin = H; p.Win = W;
    p.isY = W; p.isX = 1; p.osY = W; p.osX = 1;
    p.dil = 1; p.act = act; p.nky = 3; p.half = half != 0; p.terms = 3;
    p.wchunks = C / 8;
    p.Cin = C + G3;
    p.cpp = p.wchunks + (G3 + 7) / 8;
    p.nchunk = p.cpp;
    p.G = 8;
    return fuse == 1 ? run_groups<1>(p, Cout, B * D, st, what) : run_groups<2>(p, Cout, B * D, st, what);
}

int mma_chunks(int units, int half) { return half ? (units + 1) / 2 : units; }

// ---------------------------------------------------------------------------------------------------------------------
// "Tap projection" form of the warp levels' first conv (round 2).  The right half of the warp volume is
//     Rw[c, d, y, x] = wa * R[c, y, xa] + wb * R[c, y, xa + 1],   (xa, wa, wb) = taps of (x - candidate[d, y, x])
// — a per-position lerp of two columns, the SAME for every channel — so the channel contraction commutes with the warp:
//     sum_c W[co, c, t] * Rw[c, d, p]  =  wa * T[t, co, y, xa] + wb * T[t, co, y, xa + 1],   T[t, co] = sum_c W[co, c, t] * R[c]
// T (9 taps x Cout channels) is ONE 1x1 convolution of the right features per frame (tensor cores, tstereo_conv_d_tc2 with
// k = 1), independent of the candidates; per candidate only the 9 x 2 x Cout gathers + lerps below remain — 16x fewer
// gathers than re-building the 128-channel warped volume in the conv's producer, and no contraction per candidate at all.
// The kernel also applies what the conv's epilogue did: + the left-half conv (once per frame), + the conv over the
// group-wise channels, + bias, SiLU; and writes fp32 and / or the S-format the following (3,1,1) conv stages by TMA.
// ref: aggregation/utils/block_cost.py:47-58 (warp volume) -> aggregation/TemporalStereo/module.py:111-147 ((1,3,3) conv,
//      zero padding of the VOLUME: a neighbour outside the image contributes nothing), inverse_warp_3d.py:40-47 (taps).
constexpr int TAP_TW = 32, TAP_TH = 4;
template <int COUT>
__global__ void __launch_bounds__(TAP_TW * TAP_TH)
cost_taps_kernel(const float* __restrict__ T, const float* __restrict__ smp, const float* __restrict__ gconv,
                 const float* __restrict__ addL, const float* __restrict__ bias, float* __restrict__ out, long long osB,
                 long long osC, long long osD, unsigned short* __restrict__ so, long long ssB, long long ssD, long long ssP,
                 long long ssC8, int parts, int H, int W, int D, int tiles_x, int act) {
    pdl_sync();
    __shared__ int s_off[TAP_TH + 2][TAP_TW + 2];       // row * W + xa of the neighbour's two taps, or -1 (outside the image)
    __shared__ float s_wa[TAP_TH + 2][TAP_TW + 2], s_wb[TAP_TH + 2][TAP_TW + 2];
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int d = blockIdx.y, b = blockIdx.z;
    const int x0 = tx * TAP_TW, y0 = ty * TAP_TH;
    const int tid = threadIdx.x;
    const float* sp = smp + ((long long)b * D + d) * H * W;
    for (int i = tid; i < (TAP_TH + 2) * (TAP_TW + 2); i += TAP_TW * TAP_TH) {
        const int r = i / (TAP_TW + 2), c = i - r * (TAP_TW + 2);
        const int gy = y0 - 1 + r, gx = x0 - 1 + c;
        int o = -1;
        float wa = 0.f, wb = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            int xa;
            warp_col(gx, __ldg(sp + (long long)gy * W + gx), W, true, xa, wa, wb);
            o = warp_row(gy, H) * W + xa;
        }
        s_off[r][c] = o;
        s_wa[r][c] = wa;
        s_wb[r][c] = wb;
    }
    __syncthreads();
    const int lx = tid % TAP_TW, ly = tid / TAP_TW;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= W || y >= H) return;
    const long long HW = (long long)H * W;
    const long long pix = (long long)y * W + x;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        float a = __ldg(bias + co);
        if (addL) a += __ldg(addL + ((long long)b * COUT + co) * HW + pix);
        if (gconv) a += __ldg(gconv + (((long long)b * COUT + co) * D + d) * HW + pix);
        acc[co] = a;
    }
    const float* Tb = T + (long long)b * 9 * COUT * HW;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int o = s_off[ly + ky][lx + kx];
            if (o < 0) continue;
            const float wa = s_wa[ly + ky][lx + kx], wb = s_wb[ly + ky][lx + kx];
            const float* tp = Tb + (long long)((ky * 3 + kx) * COUT) * HW + o;
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                const float ra = __ldg(tp + co * HW), rb = __ldg(tp + co * HW + 1);
                acc[co] += fmaf(rb, wb, __fmul_rn(ra, wa));
            }
        }
    }
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = apply_act_fast(acc[co], act);   // the conv epilogue's SiLU (ex2 / rcp approx)
    if (out) {
        float* o = out + (long long)b * osB + (long long)d * osD + pix;
#pragma unroll
        for (int co = 0; co < COUT; ++co) o[co * osC] = acc[co];
    }
    if (so) {
        unsigned short* q = so + (long long)b * ssB + (long long)d * ssD + pix * 8;
#pragma unroll
        for (int c8 = 0; c8 < COUT / 8; ++c8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hi[j] = pack_h2(acc[c8 * 8 + 2 * j], acc[c8 * 8 + 2 * j + 1]);
                const float2 hf = unpack_h2(hi[j]);
                lo[j] = pack_h2(acc[c8 * 8 + 2 * j] - hf.x, acc[c8 * 8 + 2 * j + 1] - hf.y);
            }
            stg128(q + c8 * ssC8, hi[0], hi[1], hi[2], hi[3]);
            if (parts == 2) stg128(q + c8 * ssC8 + ssP, lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

}  // namespace

extern "C" {

/* operand image of the fused conv: tstereo_conv_hw3_tc2's layout over the virtual channels [C feature | 3C/8 group] */
long long tstereo_cost_conv_wpack_floats(int C, int Cout, int half) {
    return wpack_floats(mma_chunks(C / 8 + (3 * (C / 8) + 7) / 8, half), Cout);
}

int tstereo_cost_conv_warp(const float* right, const float* samples, const float* gvol, const float* addL,
                           float* out, long long osB, long long osC, long long osD,
                           const float* wpack, const float* bias, const float* oscale,
                           int B, int C, int Cout, int S, int H, int W, int act, int half, void* stream) {
    return cost_conv(1, nullptr, right, samples, gvol, addL, out, osB, osC, osD, wpack, bias, oscale, B, C, Cout, S, H, W, act, half,
                     (cudaStream_t)stream, "cost_conv_warp");
}

int tstereo_cost_conv_shift(const float* left, const float* right, const float* gvol,
                            float* out, long long osB, long long osC, long long osD,
                            const float* wpack, const float* bias, const float* oscale,
                            int B, int C, int Cout, int D, int H, int W, int act, int half, void* stream) {
    return cost_conv(2, left, right, nullptr, gvol, nullptr, out, osB, osC, osD, wpack, bias, oscale, B, C, Cout, D, H, W, act, half,
                     (cudaStream_t)stream, "cost_conv_shift");
}

int tstereo_cost_taps(const float* T, const float* samples, const float* gconv, const float* addL, const float* bias,
                      float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                      int B, int Cout, int S, int H, int W, int act, void* stream) {
    TS_REQUIRE(T && samples && bias, "cost_taps: null pointer");
    TS_REQUIRE(out || sout, "cost_taps: no output (fp32 or S-format)");
    TS_REQUIRE(B > 0 && S > 0 && H > 0 && W >= 2, "cost_taps: bad sizes");
    TS_REQUIRE(Cout == 8 || Cout == 16 || Cout == 32, "cost_taps: Cout=%d unsupported (8, 16, 32)", Cout);
    TS_REQUIRE(S <= 65535 && B <= 65535 && (long long)9 * Cout * H * W < (1ll << 31), "cost_taps: grid / plane too large");
    unsigned short* so = nullptr;
    long long ssB = 0, ssD = 0, ssP = 0, ssC8 = 0;
    int parts = 0;
    if (sout) {
        TS_REQUIRE(sout->ptr && (sout->parts == 1 || sout->parts == 2) && sout->C8 >= Cout / 8, "cost_taps: bad S-format output");
        TS_REQUIRE((((size_t)sout->ptr) & 15) == 0 && (sout->sB & 7) == 0 && (sout->sD & 7) == 0 && (sout->sP & 7) == 0 && (sout->sC8 & 7) == 0,
                   "cost_taps: S-format output must be 16-byte aligned");
        so = (unsigned short*)sout->ptr; ssB = sout->sB; ssD = sout->sD; ssP = sout->sP; ssC8 = sout->sC8; parts = sout->parts;
    }
    const int tiles_x = cdiv(W, TAP_TW), tiles_y = cdiv(H, TAP_TH);
    dim3 grid(tiles_x * tiles_y, S, B);
    cudaStream_t st = (cudaStream_t)stream;
#define TS_TAPS(CC)                                                                                                         \
    launch_k(cost_taps_kernel<CC>, grid, dim3(TAP_TW * TAP_TH), 0, st, T, samples, gconv, addL, bias, out, osB, osC, osD, so, ssB, \
             ssD, ssP, ssC8, parts, H, W, S, tiles_x, act)
    if (Cout == 8) TS_TAPS(8);
    else if (Cout == 16) TS_TAPS(16);
    else TS_TAPS(32);
#undef TS_TAPS
    return check_launch("cost_taps");
}

}  // extern "C"
