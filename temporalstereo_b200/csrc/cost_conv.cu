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

}  // extern "C"
