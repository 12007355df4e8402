// C ABI of the tensor-core convolutions (kernel template: conv_tc2_kernel.cuh).
//
// ref: architecture/modeling/layers/basic_layers.py:194-235, 340-388 as used by aggregation/TemporalStereo/module.py:111-184, 424-492.
//
// Every operator exists in two spellings over ONE implementation: `_tc2` (fp32 tensors in and out) and `_s` (either side
// may be an S-format tensor, include/tstereo.h `tstereo_split`: an S-format input is staged by TMA, an S-format output is
// written by the epilogue).
#include "conv_tc2_kernel.cuh"

using namespace tstereo;
using namespace tstereo::tc2;

namespace {

int mma_chunks(int units, int half) { return half ? (units + 1) / 2 : units; }

// S-format arguments -> kernel parameters.  `sin`: chunks the layer reads; `sout`: where the epilogue writes.
int bind_split(tc2::Params& p, SIn& si, const tstereo_split* sin, const tstereo_split* sout, const float* in, const float* out,
               int Cout, int half, const char* what) {
    TS_REQUIRE(sin || in, "%s: no input (fp32 or S-format)", what);
    TS_REQUIRE(sout || out, "%s: no output (fp32 or S-format)", what);
    if (sin) {
        TS_REQUIRE(half, "%s: an S-format input needs the fp16 split (half = 1 | 2)", what);
        TS_REQUIRE(sin->ptr && (sin->parts == 1 || sin->parts == 2) && sin->C8 > 0, "%s: bad S-format input", what);
        si.ptr = (const unsigned short*)sin->ptr;
        si.sB = sin->sB; si.sD = sin->sD; si.sP = sin->sP; si.sC8 = sin->sC8;
        si.C8 = sin->C8; si.parts = sin->parts;
    }
    if (sout) {
        TS_REQUIRE(sout->ptr && (sout->parts == 1 || sout->parts == 2), "%s: bad S-format output", what);
        TS_REQUIRE(sout->C8 >= (Cout + 7) / 8, "%s: S-format output has %d chunks, the layer writes %d", what, sout->C8, (Cout + 7) / 8);
        TS_REQUIRE((((size_t)sout->ptr) & 15) == 0 && (sout->sB & 7) == 0 && (sout->sD & 7) == 0 && (sout->sP & 7) == 0 && (sout->sC8 & 7) == 0,
                   "%s: S-format output must be 16-byte aligned", what);
        p.outs = (unsigned short*)sout->ptr;
        p.ossB = sout->sB; p.ossD = sout->sD; p.ossP = sout->sP; p.ossC8 = sout->sC8;
        p.s_parts = sout->parts;
        p.s_nb = sout->nb > 0 ? sout->nb : 0x7fffffff;
    }
    p.s_sx = 1;
    return TSTEREO_OK;
}

int conv_hw3_impl(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                  float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                  const float* wpack, const float* bias, const float* oscale,
                  int B, int Cin, int Cout, int D, int H, int W, int dilation, int act, int half, void* stream, const char* what) {
    TS_REQUIRE(wpack, "%s: null pointer", what);
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "%s: bad sizes", what);
    TS_REQUIRE(dilation == 1 || dilation == 2, "%s: dilation %d unsupported", what, dilation);
    TS_REQUIRE((long long)B * D <= 65535, "%s: B*D exceeds grid.y", what);
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "%s: packed weights must be 16-byte aligned", what);
    TS_REQUIRE((long long)H * W < (1ll << 28), "%s: plane exceeds 32-bit offsets", what);
    TS_REQUIRE(isC >= 0 && osC >= 0 && isC * 8 < (1ll << 31) && osC * 32 < (1ll << 31), "%s: channel strides exceed 32 bits", what);
    tc2::Params p = {};
    SIn si = {};
    const int rc = bind_split(p, si, sin, sout, in, out, Cout, half, what);
    if (rc != TSTEREO_OK) return rc;
    p.in = in; p.isB = isB; p.isC = (int)isC; p.isD = isD;
    p.out = out; p.osB = osB; p.osC = (int)osC; p.osD = osD;
    p.wpack = wpack; p.bias = bias; p.oscale = oscale;
    p.Cin = Cin; p.H = H; p.W = W; p.D = D; p.Hin = H; p.Win = W;
    p.isY = W; p.isX = 1; p.osY = W; p.osX = 1;
    p.dil = dilation; p.act = act; p.nky = 3; p.half = half != 0; p.terms = half == 2 ? 1 : 3;
    p.cpp = (Cin + 7) / 8;
    p.nchunk = p.cpp;
    p.G = 8;
    return run_groups(p, Cout, B * D, (cudaStream_t)stream, what, sin ? &si : nullptr);
}

int conv_hw3s2_impl(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                    float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                    const float* wpack, const float* bias, const float* oscale,
                    int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream, const char* what) {
    TS_REQUIRE(wpack, "%s: null pointer", what);
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && Hin > 0 && Win > 0, "%s: bad sizes", what);
    TS_REQUIRE((long long)B * D <= 65535, "%s: B*D exceeds grid.y", what);
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "%s: packed weights must be 16-byte aligned", what);
    TS_REQUIRE((long long)Hin * Win < (1ll << 28), "%s: plane exceeds 32-bit offsets", what);
    TS_REQUIRE(isC >= 0 && osC >= 0 && isC * 8 < (1ll << 31) && osC * 32 < (1ll << 31), "%s: channel strides exceed 32 bits", what);
    const int H = (Hin - 1) / 2 + 1, W = (Win - 1) / 2 + 1;
    tc2::Params p = {};
    SIn si = {};
    const int rc = bind_split(p, si, sin, sout, in, out, Cout, half, what);
    if (rc != TSTEREO_OK) return rc;
    p.s_sx = 2;
    p.in = in; p.isB = isB; p.isC = (int)isC; p.isD = isD;
    p.out = out; p.osB = osB; p.osC = (int)osC; p.osD = osD;
    p.wpack = wpack; p.bias = bias; p.oscale = oscale;
    p.Cin = Cin; p.H = H; p.W = W; p.D = D; p.Hin = Hin; p.Win = Win;
    p.isY = 2 * Win; p.isX = 2; p.osY = W; p.osX = 1;
    p.dil = 1; p.act = act; p.nky = 3; p.half = half != 0; p.terms = half == 2 ? 1 : 3;
    p.cpp = (Cin + 7) / 8;
    p.nchunk = 4 * p.cpp;
    p.G = 32;     // of a chunk's 9 taps only the 1-4 that exist for its phase are non-zero: same products per group as G = 8
    return run_groups(p, Cout, B * D, (cudaStream_t)stream, what, sin ? &si : nullptr);
}

int deconv_hw_impl(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                   float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                   const float* wpack, const float* bias, const float* oscale,
                   int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream, const char* what) {
    TS_REQUIRE(wpack, "%s: null pointer", what);
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && Hin > 0 && Win > 0, "%s: bad sizes", what);
    TS_REQUIRE((long long)B * D <= 65535, "%s: B*D exceeds grid.y", what);
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "%s: packed weights must be 16-byte aligned", what);
    TS_REQUIRE((long long)Hin * Win < (1ll << 26), "%s: plane exceeds 32-bit offsets", what);
    TS_REQUIRE(isC >= 0 && osC >= 0 && isC * 8 < (1ll << 31) && osC * 32 < (1ll << 31), "%s: channel strides exceed 32 bits", what);
    tc2::Params p = {};
    SIn si = {};
    const int rc0 = bind_split(p, si, sin, sout, in, out, Cout, half, what);
    if (rc0 != TSTEREO_OK) return rc0;
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
    unsigned short* outs = p.outs;
    for (int ph = 0; ph < 4; ++ph) {                           // output parity phase (py, px)
        const long long pos = (long long)(ph >> 1) * 2 * Win + (ph & 1);
        p.out = out ? out + pos : nullptr;
        p.outs = outs ? outs + pos * 8 : nullptr;
        p.wpack = wpack + ph * per_phase;
        const int rc = run_groups(p, Cout, B * D, (cudaStream_t)stream, what, sin ? &si : nullptr);
        if (rc != TSTEREO_OK) return rc;
    }
    return TSTEREO_OK;
}

int conv_d_impl(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                const float* wpack, const float* bias, const float* oscale,
                int B, int Cin, int Cout, int Din, int Dout, int H, int W,
                int k, int stride, int dilation, int transposed, int act, int half, void* stream, const char* what) {
    TS_REQUIRE(wpack, "%s: null pointer", what);
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Din > 0 && Dout > 0 && H > 0 && W > 0, "%s: bad sizes", what);
    TS_REQUIRE(k == 1 || k == 3 || k == 5, "%s: k=%d unsupported", what, k);   // k = 1: a 1x1x1 channel contraction
    if (transposed) {
        TS_REQUIRE(k == 3 && Dout == 2 * Din, "%s: transposed needs k=3, Dout=2*Din (got k=%d Din=%d Dout=%d)", what, k, Din, Dout);
    } else {
        TS_REQUIRE((stride == 1 || stride == 2) && (dilation == 1 || dilation == 2), "%s: bad stride/dilation", what);
        TS_REQUIRE(Dout == (Din - 1) / stride + 1, "%s: Dout=%d inconsistent with Din=%d stride=%d", what, Dout, Din, stride);
    }
    TS_REQUIRE((long long)B * Dout <= 65535, "%s: B*Dout exceeds grid.y", what);
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "%s: packed weights must be 16-byte aligned", what);
    TS_REQUIRE((long long)H * W < (1ll << 28), "%s: plane exceeds 32-bit offsets", what);
    TS_REQUIRE(isC >= 0 && osC >= 0 && isC * 8 < (1ll << 31) && osC * 32 < (1ll << 31), "%s: channel strides exceed 32 bits", what);
    tc2::Params p = {};
    SIn si = {};
    const int rc = bind_split(p, si, sin, sout, in, out, Cout, half, what);
    if (rc != TSTEREO_OK) return rc;
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
    return run_groups(p, Cout, B * Dout, (cudaStream_t)stream, what, sin ? &si : nullptr);
}

// fp32 NC(D)HW view -> S-format: one thread per (position, 8-channel chunk); lanes walk x, so the eight strided loads and
// the two 16-byte stores are each coalesced across the warp
__global__ void split_pack_kernel(const float* __restrict__ in, long long isB, long long isC, long long isD, unsigned short* __restrict__ so,
                                  long long sB, long long sD, long long sP, long long sC8, int parts, int C, int D, int HW, long long total) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int C8 = (C + 7) / 8;
    const int pos = (int)(i % HW);
    long long r = i / HW;
    const int c8 = (int)(r % C8);
    r /= C8;
    const int d = (int)(r % D);
    const int b = (int)(r / D);
    const float* src = in + b * isB + d * isD + (long long)(c8 * 8) * isC + pos;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = (c8 * 8 + c < C) ? __ldg(src + c * isC) : 0.f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hi[j] = pack_h2(v[2 * j], v[2 * j + 1]);
        const float2 hf = unpack_h2(hi[j]);
        lo[j] = pack_h2(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
    }
    unsigned short* dst = so + b * sB + d * sD + c8 * sC8 + (long long)pos * 8;
    stg128(dst, hi[0], hi[1], hi[2], hi[3]);
    if (parts == 2) stg128(dst + sP, lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace

extern "C" {

long long tstereo_conv_hw3_tc2_wpack_floats(int Cin, int Cout, int half) { return wpack_floats(mma_chunks((Cin + 7) / 8, half), Cout); }

int tstereo_conv_hw3_tc2(const float* in, long long isB, long long isC, long long isD,
                         float* out, long long osB, long long osC, long long osD,
                         const float* wpack, const float* bias, const float* oscale,
                         int B, int Cin, int Cout, int D, int H, int W,
                         int dilation, int act, int half, void* stream) {
    TS_REQUIRE(in && out, "conv_hw3_tc2: null pointer");
    return conv_hw3_impl(in, isB, isC, isD, nullptr, out, osB, osC, osD, nullptr, wpack, bias, oscale, B, Cin, Cout, D, H, W, dilation, act,
                         half, stream, "conv_hw3_tc2");
}

int tstereo_conv_hw3_s(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                       float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                       const float* wpack, const float* bias, const float* oscale,
                       int B, int Cin, int Cout, int D, int H, int W,
                       int dilation, int act, int half, void* stream) {
    return conv_hw3_impl(in, isB, isC, isD, sin, out, osB, osC, osD, sout, wpack, bias, oscale, B, Cin, Cout, D, H, W, dilation, act, half,
                         stream, "conv_hw3_s");
}

long long tstereo_conv_hw3s2_tc2_wpack_floats(int Cin, int Cout, int half) { return wpack_floats(mma_chunks(4 * ((Cin + 7) / 8), half), Cout); }

int tstereo_conv_hw3s2_tc2(const float* in, long long isB, long long isC, long long isD,
                           float* out, long long osB, long long osC, long long osD,
                           const float* wpack, const float* bias, const float* oscale,
                           int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream) {
    TS_REQUIRE(in && out, "conv_hw3s2_tc2: null pointer");
    return conv_hw3s2_impl(in, isB, isC, isD, nullptr, out, osB, osC, osD, nullptr, wpack, bias, oscale, B, Cin, Cout, D, Hin, Win, act, half,
                           stream, "conv_hw3s2_tc2");
}

int tstereo_conv_hw3s2_s(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                         float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                         const float* wpack, const float* bias, const float* oscale,
                         int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream) {
    return conv_hw3s2_impl(in, isB, isC, isD, sin, out, osB, osC, osD, sout, wpack, bias, oscale, B, Cin, Cout, D, Hin, Win, act, half,
                           stream, "conv_hw3s2_s");
}

long long tstereo_deconv_hw_tc2_wpack_floats(int Cin, int Cout, int half) { return 4 * wpack_floats(mma_chunks((Cin + 7) / 8, half), Cout); }

int tstereo_deconv_hw_tc2(const float* in, long long isB, long long isC, long long isD,
                          float* out, long long osB, long long osC, long long osD,
                          const float* wpack, const float* bias, const float* oscale,
                          int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream) {
    TS_REQUIRE(in && out, "deconv_hw_tc2: null pointer");
    return deconv_hw_impl(in, isB, isC, isD, nullptr, out, osB, osC, osD, nullptr, wpack, bias, oscale, B, Cin, Cout, D, Hin, Win, act, half,
                          stream, "deconv_hw_tc2");
}

int tstereo_deconv_hw_s(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                        float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                        const float* wpack, const float* bias, const float* oscale,
                        int B, int Cin, int Cout, int D, int Hin, int Win, int act, int half, void* stream) {
    return deconv_hw_impl(in, isB, isC, isD, sin, out, osB, osC, osD, sout, wpack, bias, oscale, B, Cin, Cout, D, Hin, Win, act, half,
                          stream, "deconv_hw_s");
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
    TS_REQUIRE(in && out, "conv_d_tc2: null pointer");
    return conv_d_impl(in, isB, isC, isD, nullptr, out, osB, osC, osD, nullptr, wpack, bias, oscale, B, Cin, Cout, Din, Dout, H, W, k, stride,
                       dilation, transposed, act, half, stream, "conv_d_tc2");
}

int tstereo_conv_d_s(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sin,
                     float* out, long long osB, long long osC, long long osD, const tstereo_split* sout,
                     const float* wpack, const float* bias, const float* oscale,
                     int B, int Cin, int Cout, int Din, int Dout, int H, int W,
                     int k, int stride, int dilation, int transposed, int act, int half, void* stream) {
    return conv_d_impl(in, isB, isC, isD, sin, out, osB, osC, osD, sout, wpack, bias, oscale, B, Cin, Cout, Din, Dout, H, W, k, stride,
                       dilation, transposed, act, half, stream, "conv_d_s");
}

int tstereo_split_pack(const float* in, long long isB, long long isC, long long isD, const tstereo_split* sout,
                       int B, int C, int D, int H, int W, void* stream) {
    TS_REQUIRE(in && sout && sout->ptr, "split_pack: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "split_pack: bad sizes");
    TS_REQUIRE(sout->parts == 1 || sout->parts == 2, "split_pack: parts must be 1 or 2");
    TS_REQUIRE(sout->C8 >= (C + 7) / 8, "split_pack: S-format output has %d chunks, the tensor needs %d", sout->C8, (C + 7) / 8);
    TS_REQUIRE((((size_t)sout->ptr) & 15) == 0 && (sout->sB & 7) == 0 && (sout->sD & 7) == 0 && (sout->sP & 7) == 0 && (sout->sC8 & 7) == 0,
               "split_pack: S-format output must be 16-byte aligned");
    TS_REQUIRE((long long)H * W < (1ll << 28), "split_pack: plane too large");
    const long long total = (long long)B * D * ((C + 7) / 8) * H * W;
    const int threads = 256;
    launch_k(split_pack_kernel, dim3((unsigned)cdivll(total, threads)), dim3(threads), 0, (cudaStream_t)stream, 
        in, isB, isC, isD, (unsigned short*)sout->ptr, sout->sB, sout->sD, sout->sP, sout->sC8, sout->parts, C, D, H * W, total);
    return check_launch("split_pack");
}

}  // extern "C"
