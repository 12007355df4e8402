// fp32 convolution family of the aggregation.
//
// ref: architecture/modeling/layers/basic_layers.py:194-235 (Conv3d: conv -> BN -> act),
//      :340-388 (ConvTranspose3d); aggregation/TemporalStereo/module.py:111-184 ("DepthwiseConv3D"
//      = a full channel-mixing (1,k,k) conv followed by a (k,1,1) conv).
//
// Every conv of the model is either a 3x3 conv over (H,W) batched over B*D planes, a k-tap conv
// along D, or a stride-2 transposed version of those.  Eval-mode BatchNorm is folded into the
// packed weights/bias by the host, the activation runs in the epilogue.
//
// Precision: plain TF32 operand rounding moves the regressed disparity by 0.5 px on the
// conditioned oracle (top-2 selection cascades across levels), so the
// contractions run on the fp32 FMA pipe with fp32 accumulation.
//
//  conv_hw3 : direct conv, activations stay NCHW (W contiguous).  A 256-thread CTA computes a
//             32 x TH output tile for COB output channels.  Input halo tile + weight slab of 8
//             input channels are staged with cp.async (zero-fill = padding) into a double
//             buffered shared-memory ring, overlapped with the FMAs of the previous slab.
//             lane = x  -> conflict-free LDS of the tile and coalesced stores;
//             thread  = P rows x Q couts register tile -> weights are warp-broadcast LDS.128,
//             amortised over P rows; input rows are reused across the 3 ky taps in registers.
#include "common.cuh"

namespace tstereo {

constexpr int CK = 8;  // input channels per shared-memory slab

template <int Q, int NCG, int P, int S, int DIL>
struct HW3Cfg {
    static constexpr int COB = Q * NCG;             // couts per CTA
    static constexpr int RG = 8 / NCG;              // row groups (warps along y)
    static constexpr int TH = RG * P;               // output rows per CTA
    static constexpr int IH = (TH - 1) * S + 2 * DIL + 1;
    static constexpr int IW = 31 * S + 2 * DIL + 1;
    static constexpr int HALF = (IW + 1) / 2;
    static constexpr int IWP = (S == 2) ? 2 * HALF : IW;   // S==2: even | odd columns de-interleaved
    static constexpr int IN_FLOATS = CK * IH * IWP;
    static constexpr int W_FLOATS = CK * 9 * COB;
    static constexpr int SMEM_BYTES = 2 * (IN_FLOATS + W_FLOATS) * 4;
    static constexpr int NR = (P - 1) * S + 2 * DIL + 1;  // input rows a thread touches
};

template <int Q, int NCG, int P, int S, int DIL>
__global__ void __launch_bounds__(256)
conv_hw3_kernel(const float* __restrict__ in, long long isB, long long isC, long long isD,
                float* __restrict__ out, long long osB, long long osC, long long osD,
                const float* __restrict__ w, const float* __restrict__ bias,
                int Cin, int Cout, int CoutP, int D, int Hin, int Win, int Hout, int Wout, int act) {
    pdl_sync();
    using Cfg = HW3Cfg<Q, NCG, P, S, DIL>;
    extern __shared__ __align__(16) float smem[];
    float* in_s = smem;                           // [2][CK][IH][IWP]
    float* w_s = smem + 2 * Cfg::IN_FLOATS;       // [2][CK][9][COB]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int cg = warp % NCG, rg = warp / NCG;

    const int CB = (Cout + Cfg::COB - 1) / Cfg::COB;
    const int cb = blockIdx.z % CB;
    const int n = blockIdx.z / CB;
    const int b = n / D, d = n % D;
    const int co0 = cb * Cfg::COB;
    const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * Cfg::TH;
    const int gx0 = ox0 * S - DIL, gy0 = oy0 * S - DIL;
    const float* in_img = in + (long long)b * isB + (long long)d * isD;

    auto load_slab = [&](int buf, int ci0) {
        float* dst = in_s + buf * Cfg::IN_FLOATS;
        for (int i = tid; i < CK * Cfg::IH * Cfg::IW; i += 256) {
            const int c = i % Cfg::IW;
            const int t = i / Cfg::IW;
            const int r = t % Cfg::IH;
            const int ci = t / Cfg::IH;
            const int gy = gy0 + r, gx = gx0 + c;
            const bool ok = (ci0 + ci < Cin) && gy >= 0 && gy < Hin && gx >= 0 && gx < Win;
            const float* src = ok ? in_img + (long long)(ci0 + ci) * isC + (long long)gy * Win + gx : in;
            const int cp = (S == 2) ? ((c & 1) * Cfg::HALF + (c >> 1)) : c;
            cp_async4(dst + (ci * Cfg::IH + r) * Cfg::IWP + cp, src, ok);
        }
        float* wd = w_s + buf * Cfg::W_FLOATS;
        constexpr int NV = CK * 9 * (Cfg::COB / 4);
        for (int i = tid; i < NV; i += 256) {
            const int q4 = i % (Cfg::COB / 4);
            const int t = i / (Cfg::COB / 4);      // ci*9 + tap
            const int ci = t / 9;
            const int co = co0 + q4 * 4;
            float* dp = wd + t * Cfg::COB + q4 * 4;
            if (ci0 + ci < Cin && co < CoutP)
                cp_async16(dp, w + ((long long)(ci0) * 9 + t) * CoutP + co);
            else
                *reinterpret_cast<float4*>(dp) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    float acc[P][Q];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int q = 0; q < Q; ++q) acc[p][q] = 0.f;

    const int nslab = (Cin + CK - 1) / CK;
    load_slab(0, 0);
    cp_async_commit();
    for (int k = 0; k < nslab; ++k) {
        if (k + 1 < nslab) {
            load_slab((k + 1) & 1, (k + 1) * CK);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* is = in_s + (k & 1) * Cfg::IN_FLOATS + (rg * P * S) * Cfg::IWP;
        const float* ws = w_s + (k & 1) * Cfg::W_FLOATS + cg * Q;
#pragma unroll 2
        for (int ci = 0; ci < CK; ++ci) {
            float v[Cfg::NR][3];
#pragma unroll
            for (int r = 0; r < Cfg::NR; ++r)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int c = lane * S + kx * DIL;
                    const int cp = (S == 2) ? ((c & 1) * Cfg::HALF + (c >> 1)) : c;
                    v[r][kx] = is[(ci * Cfg::IH + r) * Cfg::IWP + cp];
                }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    float wv[Q];
                    const float4* wp = reinterpret_cast<const float4*>(ws + (ci * 9 + ky * 3 + kx) * Cfg::COB);
#pragma unroll
                    for (int q4 = 0; q4 < Q / 4; ++q4) {
                        const float4 t = wp[q4];
                        wv[q4 * 4 + 0] = t.x;
                        wv[q4 * 4 + 1] = t.y;
                        wv[q4 * 4 + 2] = t.z;
                        wv[q4 * 4 + 3] = t.w;
                    }
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const float iv = v[p * S + ky * DIL][kx];
#pragma unroll
                        for (int q = 0; q < Q; ++q) acc[p][q] = fmaf(iv, wv[q], acc[p][q]);
                    }
                }
        }
        __syncthreads();
    }

    const int ox = ox0 + lane;
    if (ox >= Wout) return;
    float* out_img = out + (long long)b * osB + (long long)d * osD;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int co = co0 + cg * Q + q;
        if (co >= Cout) continue;
        const float bv = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int oy = oy0 + rg * P + p;
            if (oy < Hout) out_img[(long long)co * osC + (long long)oy * Wout + ox] = apply_act(acc[p][q] + bv, act);
        }
    }
}

template <int Q, int NCG, int P, int S, int DIL>
static int launch_hw3(const float* in, long long isB, long long isC, long long isD, float* out, long long osB,
                      long long osC, long long osD, const float* w, const float* bias, int B, int Cin, int Cout,
                      int D, int Hin, int Win, int Hout, int Wout, int act, cudaStream_t st) {
    using Cfg = HW3Cfg<Q, NCG, P, S, DIL>;
    auto kern = conv_hw3_kernel<Q, NCG, P, S, DIL>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) {
            set_error("conv_hw3: cudaFuncSetAttribute(%d B): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
            return TSTEREO_E_CUDA;
        }
        attr_done = true;
    }
    const int CB = cdiv(Cout, Cfg::COB);
    const long long gz = (long long)B * D * CB;
    TS_REQUIRE(gz <= 65535, "conv_hw3: B*D*cout_blocks = %lld exceeds grid.z", gz);
    dim3 grid(cdiv(Wout, 32), cdiv(Hout, Cfg::TH), (unsigned)gz);
    const int CoutP = (Cout + 3) & ~3;
    launch_k(kern, dim3(grid), dim3(256), Cfg::SMEM_BYTES, st, in, isB, isC, isD, out, osB, osC, osD, w, bias, Cin, Cout, CoutP, D, Hin,
                                             Win, Hout, Wout, act);
    return check_launch("conv_hw3");
}

// --------------------------------------------------------------------------- conv along D
// thread = PPT pixels x COB couts for one (b, d_out); weights [Cin][K][COB] in shared memory
// (warp-broadcast LDS.128), activations straight from global (coalesced along H*W).  The loads of CIB
// input channels x K taps are issued together before their FMAs: the layers that run here are the small
// hourglass volumes (a few hundred CTAs at most), where one dependent load per FMA group was a pure
// latency chain.  Small volumes use the (PPT = 1, 128 threads) form so the grid still covers the SMs.
template <int K, int COB, int PPT>
__global__ void __launch_bounds__(256)
conv_d_kernel(const float* __restrict__ in, long long isB, long long isC, long long isD,
              float* __restrict__ out, long long osB, long long osC, long long osD,
              const float* __restrict__ w, const float* __restrict__ bias,
              int Cin, int Cout, int CoutP, int Din, int HW, int stride, int dil, int transposed, int act) {
    pdl_sync();
    extern __shared__ __align__(16) float ws[];   // [Cin][K][COB]
    constexpr int CIB = (COB >= 32 || PPT == 2) ? 4 : 8;   // input channels whose loads are in flight together
    const int NT = blockDim.x;
    const int CB = (Cout + COB - 1) / COB;
    const int cb = blockIdx.z % CB, b = blockIdx.z / CB;
    const int co0 = cb * COB;
    const int dout = blockIdx.y;
    for (int i = threadIdx.x; i < Cin * K * (COB / 4); i += NT) {
        const int q4 = i % (COB / 4);
        const int t = i / (COB / 4);
        const int co = co0 + q4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (co < CoutP) v = *reinterpret_cast<const float4*>(w + (long long)t * CoutP + co);
        *reinterpret_cast<float4*>(ws + t * COB + q4 * 4) = v;
    }

    int px[PPT];
    bool ok[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        px[q] = blockIdx.x * (NT * PPT) + q * NT + threadIdx.x;
        ok[q] = px[q] < HW;
        px[q] = min(px[q], HW - 1);              // always a legal address; the store is guarded
    }
    float acc[PPT][COB];
#pragma unroll
    for (int q = 0; q < PPT; ++q)
#pragma unroll
        for (int c = 0; c < COB; ++c) acc[q][c] = 0.f;

    // input plane of tap k for this output plane (or -1); uniform across the CTA
    long long doff[K];
    bool dok[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        int v;
        if (transposed) {                       // dout = di*2 - 1 + k
            const int t = dout + 1 - k;
            v = (t >= 0 && (t & 1) == 0) ? (t >> 1) : -1;
        } else {
            v = dout * stride - dil * (K / 2) + k * dil;
        }
        dok[k] = v >= 0 && v < Din;
        doff[k] = dok[k] ? (long long)v * isD : 0ll;
    }
    const float* ib = in + (long long)b * isB;
    __syncthreads();
    for (int ci0 = 0; ci0 < Cin; ci0 += CIB) {
        float v[CIB][K][PPT];
#pragma unroll
        for (int u = 0; u < CIB; ++u) {
            const bool cok = ci0 + u < Cin;
            const float* src = ib + (long long)(cok ? ci0 + u : ci0) * isC;
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int q = 0; q < PPT; ++q) v[u][k][q] = (cok && dok[k]) ? __ldg(src + doff[k] + px[q]) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < CIB; ++u) {
            if (ci0 + u >= Cin) break;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (!dok[k]) continue;           // uniform across the CTA
                const float4* wp = reinterpret_cast<const float4*>(ws + ((ci0 + u) * K + k) * COB);
#pragma unroll
                for (int q4 = 0; q4 < COB / 4; ++q4) {
                    const float4 t = wp[q4];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        acc[q][q4 * 4 + 0] = fmaf(v[u][k][q], t.x, acc[q][q4 * 4 + 0]);
                        acc[q][q4 * 4 + 1] = fmaf(v[u][k][q], t.y, acc[q][q4 * 4 + 1]);
                        acc[q][q4 * 4 + 2] = fmaf(v[u][k][q], t.z, acc[q][q4 * 4 + 2]);
                        acc[q][q4 * 4 + 3] = fmaf(v[u][k][q], t.w, acc[q][q4 * 4 + 3]);
                    }
                }
            }
        }
    }
    float* ob = out + (long long)b * osB + (long long)dout * osD;
#pragma unroll
    for (int c = 0; c < COB; ++c) {
        const int co = co0 + c;
        if (co >= Cout) continue;
        const float bv = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
        for (int q = 0; q < PPT; ++q)
            if (ok[q]) ob[(long long)co * osC + px[q]] = apply_act(acc[q][c] + bv, act);
    }
}

template <int K, int COB>
static int launch_d(const float* in, long long isB, long long isC, long long isD, float* out, long long osB,
                    long long osC, long long osD, const float* w, const float* bias, int B, int Cin, int Cout,
                    int Din, int Dout, int HW, int stride, int dil, int transposed, int act, cudaStream_t st) {
    const int smem = Cin * K * COB * 4;
    TS_REQUIRE(smem <= 200 * 1024, "conv_d: weight slab of %d B does not fit shared memory", smem);
    const int CB = cdiv(Cout, COB);
    TS_REQUIRE(Dout <= 65535 && (long long)B * CB <= 65535, "conv_d: grid too large");
    // 512-pixel CTAs (2 px / thread) when they already give every SM two CTAs; otherwise 128-pixel CTAs
    const bool small = (long long)cdiv(HW, 512) * Dout * B * CB < 2 * 148;
    auto kern = small ? conv_d_kernel<K, COB, 1> : conv_d_kernel<K, COB, 2>;
    static int attr_max[2] = {48 * 1024, 48 * 1024};
    if (smem > attr_max[small]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            set_error("conv_d: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return TSTEREO_E_CUDA;
        }
        attr_max[small] = 200 * 1024;
    }
    const int nt = small ? 128 : 256, ppt = small ? 1 : 2;
    dim3 grid(cdiv(HW, nt * ppt), Dout, B * CB);
    const int CoutP = (Cout + 3) & ~3;
    launch_k(kern, dim3(grid), dim3(nt), smem, st, in, isB, isC, isD, out, osB, osC, osD, w, bias, Cin, Cout, CoutP, Din, HW, stride,
                                 dil, transposed, act);
    return check_launch("conv_d");
}

// --------------------------------------------------------------------------- transposed conv over (H,W)
// stride 2, padding 1; KS=3 (output_padding 1) or KS=4.  Gather form: thread = one output row y,
// the output column pair (2j, 2j+1), COB couts.  out[y] = sum_ky in[(y + 1 - ky) / 2] * w[ky] over the
// ky of the parity of y + 1 (at most two input rows, warp-uniform); along x the pair reads the input
// columns j-1, j, j+1 with compile-time tap indices.  The loads of CIB input channels are issued
// together before their FMAs (these layers are small: latency, not throughput, is what they wait on).
template <int KS, int COB>
__global__ void __launch_bounds__(256)
deconv_hw_kernel(const float* __restrict__ in, long long isB, long long isC, long long isD,
                 float* __restrict__ out, long long osB, long long osC, long long osD,
                 const float* __restrict__ w, const float* __restrict__ bias,
                 int Cin, int Cout, int CoutP, int D, int Hin, int Win, int act) {
    pdl_sync();
    extern __shared__ __align__(16) float ws[];   // [Cin][KS*KS][COB]
    constexpr int CIB = 4;
    const int CB = (Cout + COB - 1) / COB;
    const int cb = blockIdx.z % CB, n = blockIdx.z / CB;
    const int b = n / D, d = n % D;
    const int co0 = cb * COB;
    const int tid = threadIdx.y * 64 + threadIdx.x;
    for (int i = tid; i < Cin * KS * KS * (COB / 4); i += 256) {
        const int q4 = i % (COB / 4);
        const int t = i / (COB / 4);
        const int co = co0 + q4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (co < CoutP) v = *reinterpret_cast<const float4*>(w + (long long)t * CoutP + co);
        *reinterpret_cast<float4*>(ws + t * COB + q4 * 4) = v;
    }
    __syncthreads();

    const int Hout = 2 * Hin, Wout = 2 * Win;
    const int j = blockIdx.x * 64 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (y >= Hout || j >= Win) return;
    float acc[2][COB];
#pragma unroll
    for (int q = 0; q < COB; ++q) acc[0][q] = acc[1][q] = 0.f;

    // row slots s = 0, 1: ky = kpar + 2 s, input row (y + 1 - ky) / 2
    const int kpar = (y + 1) & 1;
    int roff[2], kyv[2];
    bool rok[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int ky = kpar + 2 * s;
        const int iy = (y + 1 - ky) >> 1;        // y + 1 - ky is even and >= -2
        rok[s] = ky < KS && (y + 1 - ky) >= 0 && iy < Hin;
        roff[s] = rok[s] ? iy * Win : 0;
        kyv[s] = ky;
    }
    // columns j-1, j, j+1 (clamped address, zero value outside)
    const bool cok[3] = {j >= 1, true, j + 1 < Win};
    const int coff[3] = {max(j - 1, 0), j, min(j + 1, Win - 1)};

    const float* ib = in + (long long)b * isB + (long long)d * isD;
    for (int ci0 = 0; ci0 < Cin; ci0 += CIB) {
        float v[CIB][2][3];
#pragma unroll
        for (int u = 0; u < CIB; ++u) {
            const bool chok = ci0 + u < Cin;
            const float* ip = ib + (long long)(chok ? ci0 + u : ci0) * isC;
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (KS == 3 && c == 0) {     // a 3-tap kernel never reads column j-1
                        v[u][s][c] = 0.f;
                        continue;
                    }
                    v[u][s][c] = (chok && rok[s] && cok[c]) ? __ldg(ip + roff[s] + coff[c]) : 0.f;
                }
        }
#pragma unroll
        for (int u = 0; u < CIB; ++u) {
            if (ci0 + u >= Cin) break;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (!rok[s]) continue;           // warp-uniform
                const float* wrow = ws + ((ci0 + u) * KS * KS + kyv[s] * KS) * COB;
#pragma unroll
                for (int px = 0; px < 2; ++px)
#pragma unroll
                    for (int kx = 0; kx < KS; ++kx) {
                        if ((px + 1 - kx) & 1) continue;         // compile-time after unrolling
                        // input column j + (px + 1 - kx) / 2 -> slot {j-1, j, j+1}
                        constexpr int dummy = 0;
                        (void)dummy;
                        const int col = 1 + (px + 1 - kx >= 0 ? (px + 1 - kx) / 2 : -((kx - px - 1) / 2));
                        const float val = v[u][s][col];
                        const float4* wp = reinterpret_cast<const float4*>(wrow + kx * COB);
#pragma unroll
                        for (int q4 = 0; q4 < COB / 4; ++q4) {
                            const float4 t = wp[q4];
                            acc[px][q4 * 4 + 0] = fmaf(val, t.x, acc[px][q4 * 4 + 0]);
                            acc[px][q4 * 4 + 1] = fmaf(val, t.y, acc[px][q4 * 4 + 1]);
                            acc[px][q4 * 4 + 2] = fmaf(val, t.z, acc[px][q4 * 4 + 2]);
                            acc[px][q4 * 4 + 3] = fmaf(val, t.w, acc[px][q4 * 4 + 3]);
                        }
                    }
            }
        }
    }
    float* ob = out + (long long)b * osB + (long long)d * osD + (long long)y * Wout + 2 * j;
#pragma unroll
    for (int q = 0; q < COB; ++q) {
        const int co = co0 + q;
        if (co >= Cout) continue;
        const float bv = bias ? __ldg(bias + co) : 0.f;
        float2 o;
        o.x = apply_act(acc[0][q] + bv, act);
        o.y = apply_act(acc[1][q] + bv, act);
        *reinterpret_cast<float2*>(ob + (long long)co * osC) = o;
    }
}

template <int KS, int COB>
static int launch_deconv(const float* in, long long isB, long long isC, long long isD, float* out, long long osB,
                         long long osC, long long osD, const float* w, const float* bias, int B, int Cin, int Cout,
                         int D, int Hin, int Win, int act, cudaStream_t st) {
    auto kern = deconv_hw_kernel<KS, COB>;
    const int smem = Cin * KS * KS * COB * 4;
    TS_REQUIRE(smem <= 200 * 1024, "deconv_hw: weight slab of %d B does not fit shared memory", smem);
    static int attr_max = 48 * 1024;
    if (smem > attr_max) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            set_error("deconv_hw: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return TSTEREO_E_CUDA;
        }
        attr_max = 200 * 1024;
    }
    const int CB = cdiv(Cout, COB);
    const long long gz = (long long)B * D * CB;
    TS_REQUIRE(gz <= 65535, "deconv_hw: grid.z too large");
    dim3 grid(cdiv(Win, 64), cdiv(2 * Hin, 4), (unsigned)gz), block(64, 4);
    const int CoutP = (Cout + 3) & ~3;
    launch_k(kern, dim3(grid), dim3(block), smem, st, in, isB, isC, isD, out, osB, osC, osD, w, bias, Cin, Cout, CoutP, D, Hin, Win,
                                    act);
    return check_launch("deconv_hw");
}

}  // namespace tstereo

using namespace tstereo;

extern "C" {

int tstereo_conv_hw3(const float* in, long long isB, long long isC, long long isD,
                     float* out, long long osB, long long osC, long long osD,
                     const float* w, const float* bias,
                     int B, int Cin, int Cout, int D, int Hin, int Win, int Hout, int Wout,
                     int stride, int dilation, int act, void* stream) {
    TS_REQUIRE(in && out && w, "conv_hw3: null pointer");
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && Hin > 0 && Win > 0, "conv_hw3: bad sizes");
    TS_REQUIRE((stride == 1 || stride == 2) && (dilation == 1 || dilation == 2) && !(stride == 2 && dilation == 2),
               "conv_hw3: unsupported stride=%d dilation=%d", stride, dilation);
    TS_REQUIRE(Hout == (Hin - 1) / stride + 1 && Wout == (Win - 1) / stride + 1,
               "conv_hw3: output %dx%d inconsistent with input %dx%d stride %d (padding = dilation)", Hout, Wout, Hin, Win, stride);
    cudaStream_t st = (cudaStream_t)stream;
#define TS_HW3(Q, NCG, P, S, DL) \
    return launch_hw3<Q, NCG, P, S, DL>(in, isB, isC, isD, out, osB, osC, osD, w, bias, B, Cin, Cout, D, Hin, Win, Hout, Wout, act, st)
    if (stride == 1 && dilation == 1) {
        if (Cout <= 8) { if (Hout >= 64) TS_HW3(8, 1, 4, 1, 1); else TS_HW3(8, 1, 2, 1, 1); }
        if (Cout <= 16) TS_HW3(8, 2, 4, 1, 1);
        if (Cout <= 32) TS_HW3(16, 2, 4, 1, 1);
        TS_HW3(16, 4, 4, 1, 1);
    } else if (stride == 1) {
        if (Cout <= 8) TS_HW3(8, 1, 2, 1, 2);
        if (Cout <= 16) TS_HW3(8, 2, 4, 1, 2);
        if (Cout <= 32) TS_HW3(16, 2, 4, 1, 2);
        TS_HW3(16, 4, 4, 1, 2);
    } else {
        if (Cout <= 16) TS_HW3(8, 2, 2, 2, 1);
        if (Cout <= 32) TS_HW3(16, 2, 2, 2, 1);
        TS_HW3(16, 4, 2, 2, 1);
    }
#undef TS_HW3
}

int tstereo_conv_d(const float* in, long long isB, long long isC, long long isD,
                   float* out, long long osB, long long osC, long long osD,
                   const float* w, const float* bias,
                   int B, int Cin, int Cout, int Din, int Dout, int HW,
                   int k, int stride, int dilation, int transposed, int act, void* stream) {
    TS_REQUIRE(in && out && w, "conv_d: null pointer");
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Din > 0 && Dout > 0 && HW > 0, "conv_d: bad sizes");
    TS_REQUIRE(k == 3 || k == 5, "conv_d: k=%d unsupported", k);
    if (transposed) {
        TS_REQUIRE(k == 3 && Dout == 2 * Din, "conv_d: transposed needs k=3, Dout=2*Din (got k=%d Din=%d Dout=%d)", k, Din, Dout);
    } else {
        TS_REQUIRE((stride == 1 || stride == 2) && (dilation == 1 || dilation == 2), "conv_d: bad stride/dilation");
        TS_REQUIRE(Dout == (Din - 1) / stride + 1, "conv_d: Dout=%d inconsistent with Din=%d stride=%d", Dout, Din, stride);
    }
    cudaStream_t st = (cudaStream_t)stream;
#define TS_D(K, COB) \
    return launch_d<K, COB>(in, isB, isC, isD, out, osB, osC, osD, w, bias, B, Cin, Cout, Din, Dout, HW, stride, dilation, transposed, act, st)
    if (k == 3) {
        if (Cout <= 8) TS_D(3, 8);
        if (Cout <= 16) TS_D(3, 16);
        TS_D(3, 32);
    } else {
        if (Cout <= 8) TS_D(5, 8);
        if (Cout <= 16) TS_D(5, 16);
        TS_D(5, 32);
    }
#undef TS_D
}

int tstereo_deconv_hw(const float* in, long long isB, long long isC, long long isD,
                      float* out, long long osB, long long osC, long long osD,
                      const float* w, const float* bias,
                      int B, int Cin, int Cout, int D, int Hin, int Win,
                      int k, int act, void* stream) {
    TS_REQUIRE(in && out && w, "deconv_hw: null pointer");
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && Hin > 0 && Win > 0, "deconv_hw: bad sizes");
    TS_REQUIRE(k == 3 || k == 4, "deconv_hw: k=%d unsupported", k);
    cudaStream_t st = (cudaStream_t)stream;
#define TS_DC(KS, COB) \
    return launch_deconv<KS, COB>(in, isB, isC, isD, out, osB, osC, osD, w, bias, B, Cin, Cout, D, Hin, Win, act, st)
    if (k == 3) {
        if (Cout <= 8) TS_DC(3, 8);
        if (Cout <= 16) TS_DC(3, 16);
        TS_DC(3, 32);
    } else {
        if (Cout <= 8) TS_DC(4, 8);
        if (Cout <= 12) TS_DC(4, 12);
        if (Cout <= 16) TS_DC(4, 16);
        TS_DC(4, 32);
    }
#undef TS_DC
}

}  // extern "C"
