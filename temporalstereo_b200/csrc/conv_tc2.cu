// C ABI of the tensor-core convolutions (kernel template: conv_tc2_kernel.cuh).
//
// ref: architecture/modeling/layers/basic_layers.py:194-235, 340-388 as used by aggregation/TemporalStereo/module.py:111-184, 424-492.
#include "conv_tc2_kernel.cuh"

using namespace tstereo;
using namespace tstereo::tc2;

extern "C" {

static int mma_chunks(int units, int half) { return half ? (units + 1) / 2 : units; }

long long tstereo_conv_hw3_tc2_wpack_floats(int Cin, int Cout, int half) { return wpack_floats(mma_chunks((Cin + 7) / 8, half), Cout); }

int tstereo_conv_hw3_tc2(const float* in, long long isB, long long isC, long long isD,
                         float* out, long long osB, long long osC, long long osD,
                         const float* wpack, const float* bias, const float* oscale,
                         int B, int Cin, int Cout, int D, int H, int W,
                         int dilation, int act, int half, void* stream) {
    TS_REQUIRE(in && out && wpack, "conv_hw3_tc2: null pointer");
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "conv_hw3_tc2: bad sizes");
    TS_REQUIRE(dilation == 1 || dilation == 2, "conv_hw3_tc2: dilation %d unsupported", dilation);
    TS_REQUIRE((long long)B * D <= 65535, "conv_hw3_tc2: B*D exceeds grid.y");
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "conv_hw3_tc2: packed weights must be 16-byte aligned");
    TS_REQUIRE((long long)H * W < (1ll << 31), "conv_hw3_tc2: plane exceeds 32-bit offsets");
    TS_REQUIRE(isC >= 0 && osC >= 0 && isC * 8 < (1ll << 31) && osC * 32 < (1ll << 31), "conv_hw3_tc2: channel strides exceed 32 bits");
    tc2::Params p = {};
    p.in = in; p.isB = isB; p.isC = (int)isC; p.isD = isD;
    p.out = out; p.osB = osB; p.osC = (int)osC; p.osD = osD;
    p.wpack = wpack; p.bias = bias; p.oscale = oscale;
    p.Cin = Cin; p.H = H; p.W = W; p.D = D; p.Hin = H; p.Win = W;
    p.isY = W; p.isX = 1; p.osY = W; p.osX = 1;
    p.dil = dilation; p.act = act; p.nky = 3; p.half = half != 0; p.terms = half == 2 ? 1 : 3;
    p.cpp = (Cin + 7) / 8;
    p.nchunk = p.cpp;
    p.G = 8;
    return run_groups(p, Cout, B * D, (cudaStream_t)stream, "conv_hw3_tc2");
}

long long tstereo_conv_hw3s2_tc2_wpack_floats(int Cin, int Cout, int half) { return wpack_floats(mma_chunks(4 * ((Cin + 7) / 8), half), Cout); }

int tstereo_conv_hw3s2_tc2(const float* in, long long isB, long long isC, long long isD,
                           float* out, long long osB, long long osC, long long osD,
                           const float* wpack, const float* bias, const float* oscale,
                           int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream) {
    TS_REQUIRE(in && out && wpack, "conv_hw3s2_tc2: null pointer");
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && Hin > 0 && Win > 0, "conv_hw3s2_tc2: bad sizes");
    TS_REQUIRE((long long)B * D <= 65535, "conv_hw3s2_tc2: B*D exceeds grid.y");
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "conv_hw3s2_tc2: packed weights must be 16-byte aligned");
    TS_REQUIRE((long long)Hin * Win < (1ll << 30), "conv_hw3s2_tc2: plane exceeds 32-bit offsets");
    TS_REQUIRE(isC >= 0 && osC >= 0 && isC * 8 < (1ll << 31) && osC * 32 < (1ll << 31), "conv_hw3s2_tc2: channel strides exceed 32 bits");
    const int H = (Hin - 1) / 2 + 1, W = (Win - 1) / 2 + 1;
    tc2::Params p = {};
    p.in = in; p.isB = isB; p.isC = (int)isC; p.isD = isD;
    p.out = out; p.osB = osB; p.osC = (int)osC; p.osD = osD;
    p.wpack = wpack; p.bias = bias; p.oscale = oscale;
    p.Cin = Cin; p.H = H; p.W = W; p.D = D; p.Hin = Hin; p.Win = Win;
    p.isY = 2 * Win; p.isX = 2; p.osY = W; p.osX = 1;
    p.dil = 1; p.act = act; p.nky = 3; p.half = half != 0; p.terms = half == 2 ? 1 : 3;
    p.cpp = (Cin + 7) / 8;
    p.nchunk = 4 * p.cpp;
    p.G = 32;     // of a chunk's 9 taps only the 1-4 that exist for its phase are non-zero: same products per group as G = 8
    return run_groups(p, Cout, B * D, (cudaStream_t)stream, "conv_hw3s2_tc2");
}

long long tstereo_deconv_hw_tc2_wpack_floats(int Cin, int Cout, int half) { return 4 * wpack_floats(mma_chunks((Cin + 7) / 8, half), Cout); }

int tstereo_deconv_hw_tc2(const float* in, long long isB, long long isC, long long isD,
                          float* out, long long osB, long long osC, long long osD,
                          const float* wpack, const float* bias, const float* oscale,
                          int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream) {
    TS_REQUIRE(in && out && wpack, "deconv_hw_tc2: null pointer");
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && Hin > 0 && Win > 0, "deconv_hw_tc2: bad sizes");
    TS_REQUIRE((long long)B * D <= 65535, "deconv_hw_tc2: B*D exceeds grid.y");
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "deconv_hw_tc2: packed weights must be 16-byte aligned");
    TS_REQUIRE((long long)Hin * Win < (1ll << 29), "deconv_hw_tc2: plane exceeds 32-bit offsets");
    TS_REQUIRE(isC >= 0 && osC >= 0 && isC * 8 < (1ll << 31) && osC * 32 < (1ll << 31), "deconv_hw_tc2: channel strides exceed 32 bits");
    tc2::Params p = {};
    p.in = in; p.isB = isB; p.isC = (int)isC; p.isD = isD;
    p.osB = osB; p.osC = (int)osC; p.osD = osD;
    p.bias = bias; p.oscale = oscale;
    p.Cin = Cin; p.H = Hin; p.W = Win; p.D = D; p.Hin = Hin; p.Win = Win;
    p.isY = Win; p.isX = 1; p.osY = 4 * Win; p.osX = 2;      // output plane is (2*Hin) x (2*Win)
    p.dil = 1; p.act = act; p.nky = 3; p.half = half != 0; p.terms = half == 2 ? 1 : 3;
    p.cpp = (Cin + 7) / 8;
    p.nchunk = p.cpp;
    p.G = 8;
    const long long per_phase = wpack_floats(mma_chunks(p.nchunk, p.half), Cout);
    for (int ph = 0; ph < 4; ++ph) {                           // output parity phase (py, px)
        p.out = out + (long long)(ph >> 1) * 2 * Win + (ph & 1);
        p.wpack = wpack + ph * per_phase;
        const int rc = run_groups(p, Cout, B * D, (cudaStream_t)stream, "deconv_hw_tc2");
        if (rc != TSTEREO_OK) return rc;
    }
    return TSTEREO_OK;
}

// the fp16 form of the (k,1,1) conv uses the single-column accumulator (N = CP); the tf32 form keeps N = 3*CP
long long tstereo_conv_d_tc2_wpack_floats(int Cin, int Cout, int k, int half) {
    return wpack_floats(mma_chunks(k * ((Cin + 7) / 8), half), Cout, 1, half ? 1 : 3);
}

int tstereo_conv_d_tc2(const float* in, long long isB, long long isC, long long isD,
                       float* out, long long osB, long long osC, long long osD,
                       const float* wpack, const float* bias, const float* oscale,
                       int B, int Cin, int Cout, int Din, int Dout, int H, int W,
                       int k, int stride, int dilation, int transposed, int act, int half, void* stream) {
    TS_REQUIRE(in && out && wpack, "conv_d_tc2: null pointer");
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Din > 0 && Dout > 0 && H > 0 && W > 0, "conv_d_tc2: bad sizes");
    TS_REQUIRE(k == 3 || k == 5, "conv_d_tc2: k=%d unsupported", k);
    if (transposed) {
        TS_REQUIRE(k == 3 && Dout == 2 * Din, "conv_d_tc2: transposed needs k=3, Dout=2*Din (got k=%d Din=%d Dout=%d)", k, Din, Dout);
    } else {
        TS_REQUIRE((stride == 1 || stride == 2) && (dilation == 1 || dilation == 2), "conv_d_tc2: bad stride/dilation");
        TS_REQUIRE(Dout == (Din - 1) / stride + 1, "conv_d_tc2: Dout=%d inconsistent with Din=%d stride=%d", Dout, Din, stride);
    }
    TS_REQUIRE((long long)B * Dout <= 65535, "conv_d_tc2: B*Dout exceeds grid.y");
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "conv_d_tc2: packed weights must be 16-byte aligned");
    TS_REQUIRE((long long)H * W < (1ll << 31), "conv_d_tc2: plane exceeds 32-bit offsets");
    TS_REQUIRE(isC >= 0 && osC >= 0 && isC * 8 < (1ll << 31) && osC * 32 < (1ll << 31), "conv_d_tc2: channel strides exceed 32 bits");
    tc2::Params p = {};
    p.in = in; p.isB = isB; p.isC = (int)isC; p.isD = isD;
    p.out = out; p.osB = osB; p.osC = (int)osC; p.osD = osD;
    p.wpack = wpack; p.bias = bias; p.oscale = oscale;
    p.Cin = Cin; p.H = H; p.W = W; p.D = Dout; p.Hin = H; p.Win = W;
    p.isY = W; p.isX = 1; p.osY = W; p.osX = 1;
    p.dil = 0; p.act = act; p.nky = 1; p.half = half != 0; p.terms = half == 2 ? 1 : 3; p.fold = half ? 1 : 3;
    p.kd = k; p.dstride = stride; p.ddil = dilation; p.Din = Din; p.dtrans = transposed;
    p.cpp = (Cin + 7) / 8;
    p.nchunk = k * p.cpp;
    p.G = 24;     // one tap per chunk: 8 products per term and chunk, the same group size in products as G = 8 of a 3x3
    return run_groups(p, Cout, B * Dout, (cudaStream_t)stream, "conv_d_tc2");
}

}  // extern "C"
