// HBM/latency-bound glue of the aggregation:
// resize+add+act, 5^3 pooling, temporal-memory merge (stable sort + plane gather), prediction
// heads, top-2 soft-argmin, candidate generation, convex / UNet up-sampling, bilinear resize.
//
// ref: architecture/modeling/aggregation/TemporalStereo/module.py:285-295 (resize+add),
//      :300-353 (ConvexUpsample), :356-398 (PredictionHeads), :401-421 (PyramidFusion),
//      :468-483 (UNet.upsample); coarse.py:69-75, 84-105; fine.py:78-95, 104-122; precise.py:98-103.
//
// All kernels map threadIdx.x to the contiguous W (or H*W) axis so every global access is a
// coalesced row segment; per-pixel work along D (<= 32 candidates) lives in registers.
#include "common.cuh"
#include <cfloat>

namespace tstereo {

// --------------------------------------------------------------------------- strided plane copy
// grid-stride over 16-byte vectors (or scalars when the plane size / alignment forbids it)
template <typename T>
__global__ void __launch_bounds__(256)
copy_planes_kernel(const T* __restrict__ in, T* __restrict__ out, long long osB, long long osC, int C, int HWv, long long total) {
    pdl_sync();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int p = (int)(i % HWv);
        const long long pl = i / HWv;
        const int c = (int)(pl % C);
        const long long b = pl / C;
        out[b * osB + c * osC + p] = in[i];
    }
}

// --------------------------------------------------------------------------- resize + add + act
// out = act(trilinear_ac(a) + skip).  thread = one output (y, x) of one (b, c); it walks all D output planes, so the
// (y, x) source indices / weights are computed once and the per-plane work is 8 loads + 1 store.  The align_corners
// scales are IEEE fp32 divisions done by the host (== __fdiv_rn): the first version spent ~300 instructions per output
// on three in-kernel divisions and the index math (ncu: issue slots 80 %, DRAM 10 %).
__global__ void __launch_bounds__(128)
resize_add_act_kernel(const float* __restrict__ a, const float* __restrict__ skip, float* __restrict__ out,
                      int Da, int Ha, int Wa, int D, int H, int W, float sd, float sy, float sx, int act) {
    pdl_sync();
    const int pix = blockIdx.x * 128 + threadIdx.x;     // linear over H*W: full CTAs whatever W is
    const int z = blockIdx.y;                           // b*C + c
    if (pix >= H * W) return;
    const int y = pix / W, x = pix - y * W;
    const LerpIdx iy = ac_index(sy, y, Ha);
    const LerpIdx ix = ac_index(sx, x, Wa);
    const size_t pl = (size_t)Ha * Wa;
    const float* p = a + (size_t)z * Da * pl;
    const int o00 = iy.i0 * Wa + ix.i0, o01 = iy.i0 * Wa + ix.i1, o10 = iy.i1 * Wa + ix.i0, o11 = iy.i1 * Wa + ix.i1;
    auto plane = [&](int dd) {
        const float* q = p + dd * pl;
        const float t0 = ix.w0 * __ldg(q + o00) + ix.w1 * __ldg(q + o01);
        const float t1 = ix.w0 * __ldg(q + o10) + ix.w1 * __ldg(q + o11);
        return iy.w0 * t0 + iy.w1 * t1;
    };
    size_t o = ((size_t)z * D * H + y) * W + x;
    const size_t HW = (size_t)H * W;
#pragma unroll 2
    for (int d = 0; d < D; ++d) {
        const LerpIdx id = ac_index(sd, d, Da);         // warp-uniform
        float v = id.w0 * plane(id.i0);
        if (id.w1 != 0.f) v += id.w1 * plane(id.i1);
        if (skip) v += __ldg(skip + o);
        out[o] = apply_act(v, act);
        o += HW;
    }
}

// The same operator writing the S-format (fp16 hi / lo split) its consumer convolution stages by TMA: thread = one output
// (d, y, x) of one (b, 8-channel chunk); it walks the chunk's 8 channels (each channel's loads are coalesced across the
// warp) and stores the two 16-byte vectors of the position.  (One thread per plane: walking D inside the thread as the
// fp32 kernel does left 8 x fewer threads with 8 x the serial work — 368 us per step against 224 for the fp32 form.)
__global__ void __launch_bounds__(128)
resize_add_act_s_kernel(const float* __restrict__ a, const float* __restrict__ skip, unsigned short* __restrict__ so,
                        long long sB, long long sD, long long sP, long long sC8, int parts, int C,
                        int Da, int Ha, int Wa, int D, int H, int W, float sd, float sy, float sx, int act) {
    pdl_sync();
    const int pix = blockIdx.x * 128 + threadIdx.x;
    const int C8 = (C + 7) >> 3;
    const int b = blockIdx.y / C8, c8 = blockIdx.y - b * C8;
    if (pix >= H * W) return;
    const int y = pix / W, x = pix - y * W;
    const LerpIdx iy = ac_index(sy, y, Ha);
    const LerpIdx ix = ac_index(sx, x, Wa);
    const size_t pl = (size_t)Ha * Wa, HW = (size_t)H * W;
    const int o00 = iy.i0 * Wa + ix.i0, o01 = iy.i0 * Wa + ix.i1, o10 = iy.i1 * Wa + ix.i0, o11 = iy.i1 * Wa + ix.i1;
    const int nc = min(8, C - c8 * 8);
    const float* pa = a + ((size_t)b * C + (size_t)c8 * 8) * Da * pl;
    const float* ps = skip ? skip + ((size_t)b * C + (size_t)c8 * 8) * D * HW + pix : nullptr;
    unsigned short* dst = so + b * sB + c8 * sC8 + (long long)pix * 8;
    {
        const int d = blockIdx.z;
        const LerpIdx id = ac_index(sd, d, Da);         // warp-uniform
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            v[c] = 0.f;
            if (c < nc) {
                const float* p = pa + (size_t)c * Da * pl;
                auto plane = [&](int dd) {
                    const float* q = p + dd * pl;
                    const float t0 = ix.w0 * __ldg(q + o00) + ix.w1 * __ldg(q + o01);
                    const float t1 = ix.w0 * __ldg(q + o10) + ix.w1 * __ldg(q + o11);
                    return iy.w0 * t0 + iy.w1 * t1;
                };
                float r = id.w0 * plane(id.i0);
                if (id.w1 != 0.f) r += id.w1 * plane(id.i1);
                if (ps) r += __ldg(ps + ((size_t)c * D + d) * HW);
                v[c] = apply_act_fast(r, act);      // the S-format consumers are tensor-core convs: same SiLU as their epilogues
            }
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = pack_h2(v[2 * j], v[2 * j + 1]);
            const float2 hf = unpack_h2(hi[j]);
            lo[j] = pack_h2(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        }
        stg128(dst + d * sD, hi[0], hi[1], hi[2], hi[3]);
        if (parts == 2) stg128(dst + d * sD + sP, lo[0], lo[1], lo[2], lo[3]);
    }
}

// --------------------------------------------------------------------------- 5x5x5 avg + max pooling
// CTA = one (b, c) and an 8 x 32 tile of (H, W); it walks all D planes once.  Per plane: halo tile
// -> shared, horizontal 5-tap pass -> shared, vertical 5-tap pass -> registers; a 5-deep register
// ring along D produces avg (count_include_pad: always /125) and max (-inf padding).
__global__ void __launch_bounds__(256)
pool5_kernel(const float* __restrict__ x, long long xsB, long long xsC, float* __restrict__ avg,
             float* __restrict__ mx, long long osB, long long osC, int C, int D, int H, int W) {
    pdl_sync();
    __shared__ float tile[12][36];
    __shared__ float hs[12][32], hm[12][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    const int c = blockIdx.z % C, b = blockIdx.z / C;
    const float* xp = x + (long long)b * xsB + (long long)c * xsC;
    float* ap = avg + (long long)b * osB + (long long)c * osC;
    float* mp = mx + (long long)b * osB + (long long)c * osC;
    const size_t HW = (size_t)H * W;
    const int ox = x0 + tx, oy = y0 + ty;
    const bool live = ox < W && oy < H;

    float rs[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    float rm[5] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int p = 0; p < D + 2; ++p) {
        float ps = 0.f, pm = -FLT_MAX;
        if (p < D) {
            const float* pl = xp + (size_t)p * HW;
            for (int i = threadIdx.x; i < 12 * 36; i += 256) {
                const int r = i / 36, cc = i % 36;
                const int gy = y0 - 2 + r, gx = x0 - 2 + cc;
                tile[r][cc] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(pl + (size_t)gy * W + gx) : NAN;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < 12 * 32; i += 256) {
                const int r = i >> 5, cc = i & 31;
                float s = 0.f, m = -FLT_MAX;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float v = tile[r][cc + k];
                    if (v == v) {  // NaN marks padding: contributes 0 to the sum, nothing to the max
                        s += v;
                        m = fmaxf(m, v);
                    }
                }
                hs[r][cc] = s;
                hm[r][cc] = m;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                ps += hs[ty + k][tx];
                pm = fmaxf(pm, hm[ty + k][tx]);
            }
        }
        // ring: slot j holds plane p-4+j after the shift
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rs[j] = rs[j + 1];
            rm[j] = rm[j + 1];
        }
        rs[4] = ps;
        rm[4] = pm;
        const int d = p - 2;
        if (d >= 0 && live) {
            const float s = rs[0] + rs[1] + rs[2] + rs[3] + rs[4];
            const float m = fmaxf(fmaxf(fmaxf(rm[0], rm[1]), fmaxf(rm[2], rm[3])), rm[4]);
            const size_t o = (size_t)d * HW + (size_t)oy * W + ox;
            ap[o] = s * (1.0f / 125.0f);
            mp[o] = m;
        }
    }
}

// --------------------------------------------------------------------------- temporal memory merge
// thread = one pixel.  Candidates = D volume samples + M memory samples; stable ascending sort
// (insertion by (value, original index)); plane j of the output volume is either the input plane
// order[j] (< D) or act(past_w[c] * mem_cost + past_b[c]) of memory slot order[j]-D.
constexpr int MAXD = 32;

__global__ void __launch_bounds__(128)
merge_memory_kernel(const float* __restrict__ vol, const float* __restrict__ samples,
                    const float* __restrict__ mem_sample, const float* __restrict__ mem_cost,
                    const float* __restrict__ past_w, const float* __restrict__ past_b,
                    float* __restrict__ out_vol, long long osB, long long osC, float* __restrict__ out_samples,
                    int C, int D, int M, int HW) {
    pdl_sync();
    const int p = blockIdx.x * 128 + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= HW) return;
    const int N = D + M;
    float key[MAXD];
    int ord[MAXD];
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        float v;
        if (i < D)
            v = __ldg(samples + ((size_t)b * D + i) * HW + p);
        else
            v = mem_sample ? __ldg(mem_sample + ((size_t)b * M + (i - D)) * HW + p) : 0.f;
        int j = i;
        while (j > 0 && key[j - 1] > v) {  // strict > keeps equal keys in original order
            key[j] = key[j - 1];
            ord[j] = ord[j - 1];
            --j;
        }
        key[j] = v;
        ord[j] = i;
    }
    // blockIdx.z splits the channels: every slice re-derives the (cheap) order, slice 0 writes the sorted candidates
    constexpr int CG = 4;
    const int c_lo = blockIdx.z * CG, c_hi = min(C, c_lo + CG);
    if (blockIdx.z == 0)
        for (int j = 0; j < N; ++j) out_samples[((size_t)b * N + j) * HW + p] = key[j];
    float mc[4];
    for (int m = 0; m < M && m < 4; ++m) mc[m] = mem_cost ? __ldg(mem_cost + ((size_t)b * M + m) * HW + p) : 0.f;
    for (int c = c_lo; c < c_hi; ++c) {
        const float* vp = vol + ((size_t)b * C + c) * D * HW + p;
        float* op = out_vol + (long long)b * osB + (long long)c * osC + p;
        const float w = __ldg(past_w + c), bb = __ldg(past_b + c);
        for (int j = 0; j < N; ++j) {
            const int o = ord[j];
            float v;
            if (o < D) {
                v = __ldg(vp + (size_t)o * HW);
            } else {
                float m = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (o - D == q) m = mc[q];
                v = silu_f(fmaf(w, m, bb));
            }
            op[(size_t)j * HW] = v;
        }
    }
}

// --------------------------------------------------------------------------- prediction heads
// cost / offset = (1,3,3) conv, C -> 1, no bias, of the two head feature stacks.
// thread = FOUR consecutive pixels of a row of one (b, d) plane: per channel and row it loads 6 values (one 16-byte vector
// + the two neighbours) for 12 FMAs per head tap row, and every weight (shared memory, warp-broadcast) serves 4 pixels —
// 2.7x fewer instructions than one pixel per thread with nine scalar loads per channel (ncu r02: that form sat at 6-10 % of
// HBM with issue slots 30-40 %: a load-latency chain).  Outside taps are zero-filled at load time (no mask multiply).
template <bool VEC>
__global__ void __launch_bounds__(128)
heads_kernel(const float* __restrict__ feat, const float* __restrict__ w, float* __restrict__ cost,
             float* __restrict__ off, int C, int D, int H, int W, float delta) {
    pdl_sync();
    extern __shared__ float ws[];  // [2][C][9]
    for (int i = threadIdx.x; i < 2 * C * 9; i += 128) ws[i] = w[i];
    __syncthreads();
    const int W4 = (W + 3) >> 2;
    const int q = blockIdx.x * 128 + threadIdx.x;       // quad index, linear over H * W4
    const int d = blockIdx.y % D, b = blockIdx.y / D;
    if (q >= H * W4) return;
    const int y = q / W4, x = (q - y * W4) * 4;
    const size_t HW = (size_t)H * W;
    // rows y-1, y, y+1: offset or -1
    int roff[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int gy = y + r - 1;
        roff[r] = (gy >= 0 && gy < H) ? gy * W + x : -1;
    }
    const bool lok = x >= 1, rok = x + 4 < W;
    float acc[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[h][k] = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float* fp = feat + (((size_t)b * 2 * C + h * C) * D + d) * HW;
        const float* wp = ws + h * C * 9;
#pragma unroll 2
        for (int c = 0; c < C; ++c) {
            const float* fq = fp + (size_t)c * D * HW;
            float v[3][6];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                if (roff[r] < 0) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) v[r][k] = 0.f;
                    continue;
                }
                const float* row = fq + roff[r];
                v[r][0] = lok ? __ldg(row - 1) : 0.f;
                if (VEC) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(row));
                    v[r][1] = t.x; v[r][2] = t.y; v[r][3] = t.z; v[r][4] = t.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[r][1 + k] = (x + k < W) ? __ldg(row + k) : 0.f;
                }
                v[r][5] = rok ? __ldg(row + 4) : 0.f;
            }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const float wv = wp[c * 9 + r * 3 + t];
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[h][k] = fmaf(v[r][k + t], wv, acc[h][k]);
                }
        }
    }
    const size_t o = ((size_t)b * D + d) * HW + (size_t)y * W + x;
    float offv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) offv[k] = fminf(fmaxf(tanhf(__fdiv_rn(acc[1][k], 100.0f)), -1.0f), 1.0f) * delta;
    if (VEC) {
        *reinterpret_cast<float4*>(cost + o) = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
        *reinterpret_cast<float4*>(off + o) = make_float4(offv[0], offv[1], offv[2], offv[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (x + k < W) {
                cost[o + k] = acc[0][k];
                off[o + k] = offv[k];
            }
    }
}

// --------------------------------------------------------------------------- top-2 soft-argmin
__global__ void __launch_bounds__(128)
predict_disp_kernel(const float* __restrict__ cost, const float* __restrict__ samples, const float* __restrict__ off,
                    float* __restrict__ disp, float* __restrict__ top_disp, float* __restrict__ top_cost,
                    int D, int HW) {
    pdl_sync();
    const int p = blockIdx.x * 128 + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= HW) return;
    float c0 = -INFINITY, c1 = -INFINITY;
    int i0 = 0, i1 = 0;
    for (int d = 0; d < D; ++d) {
        const float v = __ldg(cost + ((size_t)b * D + d) * HW + p);
        if (v > c0) {
            c1 = c0;
            i1 = i0;
            c0 = v;
            i0 = d;
        } else if (v > c1 || d == 1) {
            c1 = v;
            i1 = d;
        }
    }
    const size_t s0 = ((size_t)b * D + i0) * HW + p, s1 = ((size_t)b * D + i1) * HW + p;
    const float d0 = __fadd_rn(__ldg(samples + s0), __ldg(off + s0));
    const float d1 = __fadd_rn(__ldg(samples + s1), __ldg(off + s1));
    // softmax over the two (c0 >= c1)
    const float e1 = expf(c1 - c0);
    const float den = 1.0f + e1;
    const float p0 = __fdiv_rn(1.0f, den), p1 = __fdiv_rn(e1, den);
    disp[(size_t)b * HW + p] = __fadd_rn(__fmul_rn(p0, d0), __fmul_rn(p1, d1));
    if (top_disp) {
        top_disp[((size_t)b * 2 + 0) * HW + p] = d0;
        top_disp[((size_t)b * 2 + 1) * HW + p] = d1;
    }
    if (top_cost) {
        top_cost[((size_t)b * 2 + 0) * HW + p] = c0;
        top_cost[((size_t)b * 2 + 1) * HW + p] = c1;
    }
}

// --------------------------------------------------------------------------- candidate generation
__global__ void __launch_bounds__(256)
range_samples_kernel(const float* __restrict__ disp, float radius, float* __restrict__ low, float* __restrict__ high,
                     float* __restrict__ samples, int S_total, int c_off, int HW, long long total) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int p = (int)(i % HW);
    const int b = (int)(i / HW);
    const float dv = __ldg(disp + i);
    const float lo = __fsub_rn(dv, radius), hi = __fadd_rn(dv, radius);
    if (low) low[i] = lo;
    if (high) high[i] = hi;
    const float span = fabsf(__fsub_rn(hi, lo)), base = fminf(lo, hi);
    const float frac[5] = {0.0f, 0.375f, 0.5f, 0.625f, 1.0f};
#pragma unroll
    for (int k = 0; k < 5; ++k)
        samples[((size_t)b * S_total + c_off + k) * HW + p] = __fadd_rn(__fmul_rn(span, frac[k]), base);
}

// --------------------------------------------------------------------------- convex up-sampling x2
// thread = one coarse pixel: 36 logits = w[36][64] . m[64] + b, softmax over the 9 window taps for
// each of the 4 sub-pixels, combined with the 3x3 neighbourhood of 2*disp (zero padded).
__global__ void __launch_bounds__(128)
convex_upsample_kernel(const float* __restrict__ m, const float* __restrict__ w, const float* __restrict__ bias,
                       const float* __restrict__ disp, float* __restrict__ out, int H, int W) {
    pdl_sync();
    __shared__ __align__(16) float ws[64 * 36];  // [ci][36]
    __shared__ float bs[36];
    for (int i = threadIdx.x; i < 36 * 64; i += 128) {
        const int co = i / 64, ci = i % 64;
        ws[ci * 36 + co] = w[i];
    }
    if (threadIdx.x < 36) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const size_t HW = (size_t)H * W;
    float lg[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) lg[k] = bs[k];
    const float* mp = m + (size_t)b * 64 * HW + (size_t)y * W + x;
    for (int ci = 0; ci < 64; ++ci) {
        const float v = __ldg(mp + (size_t)ci * HW);
        const float4* wp = reinterpret_cast<const float4*>(ws + ci * 36);
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const float4 t = wp[q];
            lg[q * 4 + 0] = fmaf(v, t.x, lg[q * 4 + 0]);
            lg[q * 4 + 1] = fmaf(v, t.y, lg[q * 4 + 1]);
            lg[q * 4 + 2] = fmaf(v, t.z, lg[q * 4 + 2]);
            lg[q * 4 + 3] = fmaf(v, t.w, lg[q * 4 + 3]);
        }
    }
    float nb[9];
    const float* dp = disp + (size_t)b * HW;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int gy = y + ky - 1, gx = x + kx - 1;
            nb[ky * 3 + kx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __fmul_rn(__ldg(dp + (size_t)gy * W + gx), 2.0f) : 0.f;
        }
    float res[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {  // s = i*2 + j ; channel = k*4 + s
        float mxv = lg[s];
#pragma unroll
        for (int k = 1; k < 9; ++k) mxv = fmaxf(mxv, lg[k * 4 + s]);
        float den = 0.f, num = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float e = expf(lg[k * 4 + s] - mxv);
            den += e;
            num = fmaf(e, nb[k], num);
        }
        res[s] = __fdiv_rn(num, den);
    }
    float* op = out + (size_t)b * 4 * HW;
    const int W2 = 2 * W;
    *reinterpret_cast<float2*>(op + (size_t)(2 * y) * W2 + 2 * x) = make_float2(res[0], res[1]);
    *reinterpret_cast<float2*>(op + (size_t)(2 * y + 1) * W2 + 2 * x) = make_float2(res[2], res[3]);
}

// --------------------------------------------------------------------------- UNet up-sampling
// full[y,x] = sum_k softmax_k(logits) * bilinear_ac(unfold3x3(disp)[k] * W / w)
// CTA = 128 consecutive x of one output row.  The pre-scaled low-resolution neighbourhood the CTA needs (4 rows x <= 64
// columns) is staged once in shared memory — one IEEE division per staged value instead of 16 per thread — and the
// align_corners scales come from the host.  (ncu r02: the per-thread form was instruction-bound, issue slots 83 %.)
constexpr int UU_TW = 64;
__global__ void __launch_bounds__(128)
unet_upsample_kernel(const float* __restrict__ logits, const float* __restrict__ disp, float* __restrict__ full,
                     int H, int W, int h, int w, float sy, float sx) {
    pdl_sync();
    __shared__ float tile[4][UU_TW];
    const int x0 = blockIdx.x * 128;
    const int x = x0 + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const LerpIdx iy = ac_index(sy, y, h);
    const int xc = min(x, W - 1);
    const LerpIdx ix = ac_index(sx, xc, w);
    const float* dp = disp + (size_t)b * h * w;
    const float mulv = (float)W, divv = (float)w;
    // staged columns [cmin, cmin + ncol): from one left of the first pixel's i0 to two right of the last pixel's
    const int cmin = ac_index(sx, x0, w).i0 - 1;
    const int ncol = ac_index(sx, min(x0 + 127, W - 1), w).i0 + 2 - cmin + 1;
    const bool staged = ncol <= UU_TW;                   // always for the model's 4x ratio; other ratios read directly
    auto scaled = [&](int gy, int gx) {
        return (gy >= 0 && gy < h && gx >= 0 && gx < w) ? __fdiv_rn(__fmul_rn(__ldg(dp + (size_t)gy * w + gx), mulv), divv) : 0.f;
    };
    if (staged) {
        for (int i = threadIdx.x; i < 4 * ncol; i += 128) {
            const int r = i / ncol, c = i - r * ncol;
            tile[r][c] = scaled(iy.i0 - 1 + r, cmin + c);
        }
        __syncthreads();
    }
    if (x >= W) return;
    // 4x4 neighbourhood of the low-res disparity around (iy.i0, ix.i0), zero outside, pre-scaled
    float nb[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            nb[r][c] = staged ? tile[r][ix.i0 - 1 - cmin + c] : scaled(iy.i0 - 1 + r, ix.i0 - 1 + c);
    const int dy1 = iy.i1 - iy.i0, dx1 = ix.i1 - ix.i0;  // 0 at the last row / column
    const float* lp = logits + (size_t)b * 9 * HW + (size_t)y * W + x;
    float lg[9], mxv = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        lg[k] = __ldg(lp + (size_t)k * HW);
        mxv = fmaxf(mxv, lg[k]);
    }
    float den = 0.f, num = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            // unfolded plane k at low-res (r, c) = disp[r+ky-1, c+kx-1]
            // dx1, dy1 are 0 or 1: selects instead of dynamic register indexing (which would spill nb to local memory)
            const float v00 = nb[ky][kx], v01 = dx1 ? nb[ky][kx + 1] : v00;
            const float v10 = dy1 ? nb[ky + 1][kx] : v00;
            const float v11 = dy1 ? (dx1 ? nb[ky + 1][kx + 1] : nb[ky + 1][kx]) : v01;
            const float t0 = ix.w0 * v00 + ix.w1 * v01;
            const float t1 = ix.w0 * v10 + ix.w1 * v11;
            const float v = iy.w0 * t0 + iy.w1 * t1;
            const float e = expf(lg[ky * 3 + kx] - mxv);
            den += e;
            num = fmaf(e, v, num);
        }
    full[(size_t)b * HW + (size_t)y * W + x] = __fdiv_rn(num, den);
}

// --------------------------------------------------------------------------- bilinear resize
__global__ void __launch_bounds__(128)
bilinear_resize_kernel(const float* __restrict__ in, float* __restrict__ out, float mul, float div, int C, int Hi,
                       int Wi, int Ho, int Wo, int C_total, int c_off) {
    pdl_sync();
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    const int c = blockIdx.z % C, b = blockIdx.z / C;
    if (x >= Wo) return;
    const LerpIdx iy = ac_index(ac_scale(Hi, Ho), y, Hi);
    const LerpIdx ix = ac_index(ac_scale(Wi, Wo), x, Wi);
    const float* p = in + ((size_t)b * C + c) * Hi * Wi;
    auto ld = [&](int yy, int xx) { return __fdiv_rn(__fmul_rn(__ldg(p + (size_t)yy * Wi + xx), mul), div); };
    const float t0 = ix.w0 * ld(iy.i0, ix.i0) + ix.w1 * ld(iy.i0, ix.i1);
    const float t1 = ix.w0 * ld(iy.i1, ix.i0) + ix.w1 * ld(iy.i1, ix.i1);
    out[(((size_t)b * C_total + c_off + c) * Ho + y) * Wo + x] = iy.w0 * t0 + iy.w1 * t1;
}

}  // namespace tstereo

using namespace tstereo;

extern "C" {

int tstereo_copy_planes(const float* in, float* out, long long osB, long long osC, int B, int C, int HW, void* stream) {
    TS_REQUIRE(in && out, "copy_planes: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && HW > 0 && osB >= 0 && osC >= HW, "copy_planes: bad sizes");
    const bool vec = (HW % 4 == 0) && (osB % 4 == 0) && (osC % 4 == 0) && ((((size_t)in) | ((size_t)out)) & 15) == 0;
    const long long total = (long long)B * C * (vec ? HW / 4 : HW);
    const unsigned grid = (unsigned)(cdivll(total, 256) < 148 * 16 ? cdivll(total, 256) : 148 * 16);
    if (vec)
        launch_k(copy_planes_kernel<float4>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out),
                                                                           osB / 4, osC / 4, C, HW / 4, total);
    else
        launch_k(copy_planes_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, in, out, osB, osC, C, HW, total);
    return check_launch("copy_planes");
}

int tstereo_resize_add_act(const float* a, const float* skip, float* out, int B, int C, int Da, int Ha, int Wa,
                           int D, int H, int W, int act, void* stream) {
    TS_REQUIRE(a && out, "resize_add_act: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && Da > 0 && Ha > 0 && Wa > 0 && D > 0 && H > 0 && W > 0, "resize_add_act: bad sizes");
    TS_REQUIRE((long long)B * C <= 65535 && (long long)H * W < (1ll << 31), "resize_add_act: grid too large");
    auto scale = [](int in_size, int out_size) { return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.0f; };
    dim3 grid(cdiv(H * W, 128), B * C);
    launch_k(resize_add_act_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, a, skip, out, Da, Ha, Wa, D, H, W, scale(Da, D), scale(Ha, H),
                                                                  scale(Wa, W), act);
    return check_launch("resize_add_act");
}

int tstereo_resize_add_act_s(const float* a, const float* skip, const tstereo_split* sout, int B, int C, int Da, int Ha, int Wa,
                             int D, int H, int W, int act, void* stream) {
    TS_REQUIRE(a && sout && sout->ptr, "resize_add_act_s: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && Da > 0 && Ha > 0 && Wa > 0 && D > 0 && H > 0 && W > 0, "resize_add_act_s: bad sizes");
    TS_REQUIRE((long long)B * ((C + 7) / 8) <= 65535 && D <= 65535 && (long long)H * W < (1ll << 28), "resize_add_act_s: grid too large");
    TS_REQUIRE((sout->parts == 1 || sout->parts == 2) && sout->C8 >= (C + 7) / 8, "resize_add_act_s: bad S-format output");
    TS_REQUIRE((((size_t)sout->ptr) & 15) == 0 && (sout->sB & 7) == 0 && (sout->sD & 7) == 0 && (sout->sP & 7) == 0 && (sout->sC8 & 7) == 0,
               "resize_add_act_s: S-format output must be 16-byte aligned");
    auto scale = [](int in_size, int out_size) { return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.0f; };
    dim3 grid(cdiv(H * W, 128), B * ((C + 7) / 8), D);
    launch_k(resize_add_act_s_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, a, skip, (unsigned short*)sout->ptr, sout->sB, sout->sD, sout->sP,
                                                                    sout->sC8, sout->parts, C, Da, Ha, Wa, D, H, W, scale(Da, D),
                                                                    scale(Ha, H), scale(Wa, W), act);
    return check_launch("resize_add_act_s");
}

int tstereo_pool5(const float* x, long long xsB, long long xsC, float* avg, float* mx, long long osB, long long osC,
                  int B, int C, int D, int H, int W, void* stream) {
    TS_REQUIRE(x && avg && mx, "pool5: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "pool5: bad sizes");
    TS_REQUIRE((long long)B * C <= 65535 && cdiv(H, 8) <= 65535, "pool5: grid too large");
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B * C);
    launch_k(pool5_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, xsB, xsC, avg, mx, osB, osC, C, D, H, W);
    return check_launch("pool5");
}

int tstereo_merge_memory(const float* vol, const float* samples, const float* mem_sample, const float* mem_cost,
                         const float* past_w, const float* past_b, float* out_vol, long long osB, long long osC,
                         float* out_samples, int B, int C, int D, int M, int H, int W, void* stream) {
    TS_REQUIRE(vol && samples && past_w && past_b && out_vol && out_samples, "merge_memory: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && D > 0 && M >= 0 && M <= 4 && D + M <= MAXD, "merge_memory: need M <= 4 and D+M <= %d (D=%d M=%d)", MAXD, D, M);
    TS_REQUIRE((mem_sample == nullptr) == (mem_cost == nullptr), "merge_memory: mem_sample and mem_cost must both be set or both NULL");
    TS_REQUIRE(B <= 65535, "merge_memory: B too large");
    const int HW = H * W;
    dim3 grid(cdiv(HW, 128), B, cdiv(C, 4));
    launch_k(merge_memory_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, vol, samples, mem_sample, mem_cost, past_w, past_b,
                                                               out_vol, osB, osC, out_samples, C, D, M, HW);
    return check_launch("merge_memory");
}

int tstereo_heads(const float* feat, const float* w, float* cost, float* off, int B, int C, int D, int H, int W,
                  float delta, void* stream) {
    TS_REQUIRE(feat && w && cost && off, "heads: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && C <= 256 && D > 0 && H > 0 && W > 0, "heads: bad sizes");
    TS_REQUIRE((long long)B * D <= 65535 && (long long)H * W < (1ll << 31), "heads: grid too large");
    dim3 grid(cdiv(H * ((W + 3) / 4), 128), B * D);
    const bool vec = (W % 4 == 0) && ((((size_t)feat) | ((size_t)cost) | ((size_t)off)) & 15) == 0;
    if (vec) launch_k(heads_kernel<true>, dim3(grid), dim3(128), 2 * C * 9 * sizeof(float), (cudaStream_t)stream, feat, w, cost, off, C, D, H, W, delta);
    else launch_k(heads_kernel<false>, dim3(grid), dim3(128), 2 * C * 9 * sizeof(float), (cudaStream_t)stream, feat, w, cost, off, C, D, H, W, delta);
    return check_launch("heads");
}

int tstereo_predict_disp(const float* cost, const float* samples, const float* off, float* disp, float* top_disp,
                         float* top_cost, int B, int D, int H, int W, void* stream) {
    TS_REQUIRE(cost && samples && off && disp, "predict_disp: null pointer");
    TS_REQUIRE(B > 0 && B <= 65535 && D >= 2 && H > 0 && W > 0, "predict_disp: need D >= 2 (D=%d)", D);
    const int HW = H * W;
    dim3 grid(cdiv(HW, 128), B);
    launch_k(predict_disp_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, cost, samples, off, disp, top_disp, top_cost, D, HW);
    return check_launch("predict_disp");
}

int tstereo_range_samples(const float* disp, float radius, float* low, float* high, float* samples, int S_total,
                          int c_off, int B, int H, int W, void* stream) {
    TS_REQUIRE(disp && samples, "range_samples: null pointer");
    TS_REQUIRE(B > 0 && H > 0 && W > 0 && c_off >= 0 && c_off + 5 <= S_total, "range_samples: bad sizes");
    const long long total = (long long)B * H * W;
    launch_k(range_samples_kernel, dim3((unsigned)cdivll(total, 256)), dim3(256), 0, (cudaStream_t)stream, disp, radius, low, high, samples,
                                                                                       S_total, c_off, H * W, total);
    return check_launch("range_samples");
}

int tstereo_convex_upsample(const float* m, const float* w, const float* b, const float* disp, float* out, int B,
                            int H, int W, void* stream) {
    TS_REQUIRE(m && w && b && disp && out, "convex_upsample: null pointer");
    TS_REQUIRE(B > 0 && B <= 65535 && H > 0 && H <= 65535 && W > 0, "convex_upsample: bad sizes");
    dim3 grid(cdiv(W, 128), H, B);
    launch_k(convex_upsample_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, m, w, b, disp, out, H, W);
    return check_launch("convex_upsample");
}

int tstereo_unet_upsample(const float* logits, const float* disp, float* full, int B, int H, int W, int h, int w,
                          void* stream) {
    TS_REQUIRE(logits && disp && full, "unet_upsample: null pointer");
    TS_REQUIRE(B > 0 && B <= 65535 && H > 0 && H <= 65535 && W > 0 && h > 0 && w > 0, "unet_upsample: bad sizes");
    dim3 grid(cdiv(W, 128), H, B);
    auto scale = [](int in_size, int out_size) { return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.0f; };
    launch_k(unet_upsample_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, logits, disp, full, H, W, h, w, scale(h, H), scale(w, W));
    return check_launch("unet_upsample");
}

int tstereo_bilinear_resize(const float* in, float* out, float mul, float div, int B, int C, int Hi, int Wi, int Ho,
                            int Wo, int C_total, int c_off, void* stream) {
    TS_REQUIRE(in && out, "bilinear_resize: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && c_off >= 0 && c_off + C <= C_total,
               "bilinear_resize: bad sizes");
    TS_REQUIRE(Ho <= 65535 && (long long)B * C <= 65535, "bilinear_resize: grid too large");
    dim3 grid(cdiv(Wo, 128), Ho, B * C);
    launch_k(bilinear_resize_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, in, out, mul, div, C, Hi, Wi, Ho, Wo, C_total, c_off);
    return check_launch("bilinear_resize");
}

}  // extern "C"
