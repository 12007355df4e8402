// Pose-conditioned temporal warp of the recurrent state.
//
// ref: projects/TemporalStereo/TemporalStereo.py:326-461 (update_map),
//      architecture/modeling/layers/inverse_warp.py:92-178 (project_to_3d),
//      architecture/modeling/layers/softsplat.py:8-53 (forward splat), :334-360 (softmax packing).
//
// Everything here runs at 1/8 resolution (a few 10^4 pixels): the kernels are latency-bound, so
// each stage is a single small launch on the caller's stream (the reference's CuPy launch ignores
// torch's current stream; this one does not).
#include "common.cuh"

namespace tstereo {

// params per batch item: invK (row-major 3x3) [0..8] | P = (down_K @ T)[:3,:] (3x4) [9..20] |
// focal [21] | baseline [22] | pad [23]
constexpr int NPARAM = 24;

// pose / intrinsics of batch item b -> the NPARAM floats at `o` (global or shared memory)
__device__ void pose_prep_one(const float* __restrict__ K, const float* __restrict__ T_now,
                              const float* __restrict__ inv_T_prev, const float* __restrict__ baseline,
                              float factor, float* o, int b) {
    const float* Kb = K + b * 16;
    const float* A = T_now + b * 16;
    const float* Bm = inv_T_prev + b * 16;
    float T[16], dK[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s = fmaf(A[i * 4 + k], Bm[k * 4 + j], s);
            T[i * 4 + j] = s;
        }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) dK[i * 4 + j] = (i < 2) ? __fdiv_rn(Kb[i * 4 + j], factor) : Kb[i * 4 + j];
    // 4x4 inverse: Gauss-Jordan with partial pivoting in double, rounded to fp32
    double M[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            M[i][j] = (double)dK[i * 4 + j];
            M[i][4 + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        double best = fabs(M[c][c]);
        for (int r = c + 1; r < 4; ++r)
            if (fabs(M[r][c]) > best) {
                best = fabs(M[r][c]);
                piv = r;
            }
        if (piv != c)
            for (int j = 0; j < 8; ++j) {
                const double t = M[c][j];
                M[c][j] = M[piv][j];
                M[piv][j] = t;
            }
        const double inv = 1.0 / M[c][c];
        for (int j = 0; j < 8; ++j) M[c][j] *= inv;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double f = M[r][c];
            for (int j = 0; j < 8; ++j) M[r][j] -= f * M[c][j];
        }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) o[i * 3 + j] = (float)M[i][4 + j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s = fmaf(dK[i * 4 + k], T[k * 4 + j], s);
            o[9 + i * 4 + j] = s;
        }
    o[21] = dK[0];
    o[22] = baseline[b];
    o[23] = 0.f;
}

__global__ void pose_prep_kernel(const float* __restrict__ K, const float* __restrict__ T_now,
                                 const float* __restrict__ inv_T_prev, const float* __restrict__ baseline,
                                 float factor, float* __restrict__ params, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    pose_prep_one(K, T_now, inv_T_prev, baseline, factor, params + b * NPARAM, b);
}

// disparity at pixel (x, y) -> depth -> camera point -> moved + re-projected: (sx, sy, sz) of inverse_warp.py:138-160
struct Proj {
    float sx, sy, sz, bf;
};
__device__ __forceinline__ Proj reproject_point(const float* pr, float disp, float fx, float fy) {
    Proj r;
    r.bf = __fmul_rn(pr[22], pr[21]);
    const float depth = __fdiv_rn(r.bf, __fadd_rn(disp, 1e-5f));
    const float X = __fmul_rn(fmaf(pr[2], 1.0f, fmaf(pr[1], fy, __fmul_rn(pr[0], fx))), depth);
    const float Y = __fmul_rn(fmaf(pr[5], 1.0f, fmaf(pr[4], fy, __fmul_rn(pr[3], fx))), depth);
    const float Z = __fmul_rn(fmaf(pr[8], 1.0f, fmaf(pr[7], fy, __fmul_rn(pr[6], fx))), depth);
    const float* P = pr + 9;
    r.sx = fmaf(P[3], 1.0f, fmaf(P[2], Z, fmaf(P[1], Y, __fmul_rn(P[0], X))));
    r.sy = fmaf(P[7], 1.0f, fmaf(P[6], Z, fmaf(P[5], Y, __fmul_rn(P[4], X))));
    r.sz = fmaf(P[11], 1.0f, fmaf(P[10], Z, fmaf(P[9], Y, __fmul_rn(P[8], X))));
    return r;
}

// the four bilinear corners a source pixel splats to (softsplat.py:20-50)
struct Corners {
    int nwx, nwy;
    float wnw, wne, wsw, wse;
    bool x0ok, x1ok, y0ok, y1ok, ok;
};
__device__ __forceinline__ Corners splat_corners(float ox, float oy, int h, int w) {
    Corners c;
    c.ok = (fabsf(ox) < 1e9f) && (fabsf(oy) < 1e9f);      // NaN / inf targets land nowhere
    c.nwx = c.ok ? (int)floorf(ox) : 0;
    c.nwy = c.ok ? (int)floorf(oy) : 0;
    c.wnw = __fmul_rn(__fsub_rn((float)(c.nwx + 1), ox), __fsub_rn((float)(c.nwy + 1), oy));
    c.wne = __fmul_rn(__fsub_rn(ox, (float)c.nwx), __fsub_rn((float)(c.nwy + 1), oy));
    c.wsw = __fmul_rn(__fsub_rn((float)(c.nwx + 1), ox), __fsub_rn(oy, (float)c.nwy));
    c.wse = __fmul_rn(__fsub_rn(ox, (float)c.nwx), __fsub_rn(oy, (float)c.nwy));
    c.x0ok = c.ok && c.nwx >= 0 && c.nwx < w;
    c.x1ok = c.ok && c.nwx + 1 >= 0 && c.nwx + 1 < w;
    c.y0ok = c.nwy >= 0 && c.nwy < h;
    c.y1ok = c.nwy + 1 >= 0 && c.nwy + 1 < h;
    return c;
}
__device__ __forceinline__ void splat_add(float* a, const Corners& c, float v, int w) {
    if (c.x0ok && c.y0ok) atomicAdd(a + (size_t)c.nwy * w + c.nwx, __fmul_rn(v, c.wnw));
    if (c.x1ok && c.y0ok) atomicAdd(a + (size_t)c.nwy * w + c.nwx + 1, __fmul_rn(v, c.wne));
    if (c.x0ok && c.y1ok) atomicAdd(a + (size_t)(c.nwy + 1) * w + c.nwx, __fmul_rn(v, c.wsw));
    if (c.x1ok && c.y1ok) atomicAdd(a + (size_t)(c.nwy + 1) * w + c.nwx + 1, __fmul_rn(v, c.wse));
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused update_map: three launches for the whole temporal warp (projects/TemporalStereo/TemporalStereo.py:326-461).
//   1. prep  : prev_disp -> 1/8 scale (bilinear align_corners, x w / W), per-CTA partial sums of it (the batch-global
//              mean of the splat metric), and the splat accumulators zeroed
//   2. splat : mean from the partials (fixed order), pose parameters of the CTA's batch item (thread 0, shared memory),
//              flow + metric from the down-sampled disparity, re-projection of the stored top-2 samples and of the
//              local-map stack, softmax splat of both groups with the SAME flow / metric (update_past_cost :386-426 and
//              update_local_map :340-384 compute them twice, identically)
//   3. norm  : x / (norm + 1e-22) of both groups
// Same arithmetic as the per-stage kernels above (they stay the operator-level drop-ins of project_to_3d / FunctionSoftsplat).
constexpr int UM_MAXLM = 4;

__global__ void __launch_bounds__(256)
um_prep_kernel(const float* __restrict__ prev, int Hf, int Wf, float sy, float sx, float* __restrict__ pd,
               double* __restrict__ partial, float* __restrict__ acc_mem, int cm, float* __restrict__ acc_lm, int cl, int h, int w) {
    __shared__ double sm[256];
    const int hw = h * w;
    const int p = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    double v = 0.0;
    if (p < hw) {
        const int y = p / w, x = p - y * w;
        const LerpIdx iy = ac_index(sy, y, Hf), ix = ac_index(sx, x, Wf);
        const float* q = prev + (size_t)b * Hf * Wf;
        const float mul = (float)w, div = (float)Wf;
        auto ld = [&](int yy, int xx) { return __fdiv_rn(__fmul_rn(__ldg(q + (size_t)yy * Wf + xx), mul), div); };
        const float t0 = ix.w0 * ld(iy.i0, ix.i0) + ix.w1 * ld(iy.i0, ix.i1);
        const float t1 = ix.w0 * ld(iy.i1, ix.i0) + ix.w1 * ld(iy.i1, ix.i1);
        const float r = iy.w0 * t0 + iy.w1 * t1;
        pd[(size_t)b * hw + p] = r;
        v = (double)r;
        for (int c = 0; c < cm; ++c) acc_mem[((size_t)b * cm + c) * hw + p] = 0.f;
        for (int c = 0; c < cl; ++c) acc_lm[((size_t)b * cl + c) * hw + p] = 0.f;
    }
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sm[threadIdx.x] += sm[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(256)
um_splat_kernel(const float* __restrict__ pd, const double* __restrict__ partial, int nparts, long long total,
                const float* __restrict__ K, const float* __restrict__ T_now, const float* __restrict__ inv_T_prev,
                const float* __restrict__ baseline, float factor,
                const float* __restrict__ mem_sample, const float* __restrict__ mem_cost, int M, float* __restrict__ acc_mem,
                const float* __restrict__ lm, int n_in, int n_out, float* __restrict__ acc_lm, int h, int w) {
    __shared__ float pr[NPARAM];
    __shared__ double red[256];
    __shared__ float mean_s;
    const int b = blockIdx.y;
    // batch-global mean of the down-sampled disparity: every CTA sums the same partials in the same order
    double s = 0.0;
    for (int k = threadIdx.x; k < nparts; k += 256) s += partial[k];
    red[threadIdx.x] = s;
    if (threadIdx.x == 0) pose_prep_one(K, T_now, inv_T_prev, baseline, factor, pr, b);
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) mean_s = (float)(red[0] / (double)total);
    __syncthreads();
    const int hw = h * w;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= hw) return;
    const int y = p / w, x = p - y * w;
    const float fx = (float)x, fy = (float)y;
    const float d0 = __ldg(pd + (size_t)b * hw + p);
    const float metric = fminf(fmaxf(__fsub_rn(d0, mean_s), -50.0f), 50.0f);
    const float e = expf(metric);
    const Proj p0 = reproject_point(pr, d0, fx, fy);
    const float den = __fadd_rn(p0.sz, 1e-7f);
    // target = pixel + flow, flow = reprojected - pixel (computed in that order by the reference: keep both roundings)
    const float ox = __fadd_rn(fx, __fsub_rn(__fdiv_rn(p0.sx, den), fx));
    const float oy = __fadd_rn(fy, __fsub_rn(__fdiv_rn(p0.sy, den), fy));
    const Corners c = splat_corners(ox, oy, h, w);
    if (!c.ok) return;
    if (acc_mem) {                                    // [reprojected samples (M) | costs (M) | norm]
        float* a = acc_mem + (size_t)b * (2 * M + 1) * hw;
        for (int m = 0; m < M; ++m) {
            const Proj q = reproject_point(pr, __ldg(mem_sample + ((size_t)b * M + m) * hw + p), fx, fy);
            splat_add(a + (size_t)m * hw, c, __fmul_rn(__fdiv_rn(q.bf, __fadd_rn(q.sz, 1e-5f)), e), w);
        }
        for (int m = 0; m < M; ++m)
            splat_add(a + (size_t)(M + m) * hw, c, __fmul_rn(__ldg(mem_cost + ((size_t)b * M + m) * hw + p), e), w);
        splat_add(a + (size_t)(2 * M) * hw, c, e, w);
    }
    if (acc_lm) {                                     // stack = [pd, local_map][:n_out], each re-projected
        float* a = acc_lm + (size_t)b * (n_out + 1) * hw;
        for (int k = 0; k < n_out; ++k) {
            const float dv = k == 0 ? d0 : __ldg(lm + ((size_t)b * n_in + (k - 1)) * hw + p);
            const Proj q = k == 0 ? p0 : reproject_point(pr, dv, fx, fy);
            splat_add(a + (size_t)k * hw, c, __fmul_rn(__fdiv_rn(q.bf, __fadd_rn(q.sz, 1e-5f)), e), w);
        }
        splat_add(a + (size_t)n_out * hw, c, e, w);
    }
}

__global__ void __launch_bounds__(256)
um_norm_kernel(const float* __restrict__ acc_mem, int cm, float* __restrict__ out_sample, float* __restrict__ out_cost, int M,
               const float* __restrict__ acc_lm, int cl, float* __restrict__ out_lm, int hw) {
    const int p = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= hw) return;
    if (acc_mem) {
        const float* a = acc_mem + (size_t)b * cm * hw + p;
        const float n = __fadd_rn(__ldg(a + (size_t)(cm - 1) * hw), 1e-22f);
        for (int m = 0; m < M; ++m) {
            out_sample[((size_t)b * M + m) * hw + p] = __fdiv_rn(__ldg(a + (size_t)m * hw), n);
            out_cost[((size_t)b * M + m) * hw + p] = __fdiv_rn(__ldg(a + (size_t)(M + m) * hw), n);
        }
    }
    if (acc_lm) {
        const float* a = acc_lm + (size_t)b * cl * hw + p;
        const float n = __fadd_rn(__ldg(a + (size_t)(cl - 1) * hw), 1e-22f);
        for (int k = 0; k < cl - 1; ++k) out_lm[((size_t)b * (cl - 1) + k) * hw + p] = __fdiv_rn(__ldg(a + (size_t)k * hw), n);
    }
}

// disparity -> depth -> camera point -> moved + re-projected.  One thread per (b, c, y, x).
__global__ void __launch_bounds__(256)
reproject_kernel(const float* __restrict__ disp, const float* __restrict__ params, float* __restrict__ flow,
                 float* __restrict__ new_disp, int C_total, int c_off, int C, int h, int w, long long total) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % w);
    long long t = i / w;
    const int y = (int)(t % h);
    t /= h;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const float* pr = params + b * NPARAM;
    const float bf = __fmul_rn(pr[22], pr[21]);
    const float depth = __fdiv_rn(bf, __fadd_rn(__ldg(disp + i), 1e-5f));
    const float fx = (float)x, fy = (float)y;
    // cam = invK[:3,:3] @ [x, y, 1] * depth            (inverse_warp.py:138)
    const float X = __fmul_rn(fmaf(pr[2], 1.0f, fmaf(pr[1], fy, __fmul_rn(pr[0], fx))), depth);
    const float Y = __fmul_rn(fmaf(pr[5], 1.0f, fmaf(pr[4], fy, __fmul_rn(pr[3], fx))), depth);
    const float Z = __fmul_rn(fmaf(pr[8], 1.0f, fmaf(pr[7], fy, __fmul_rn(pr[6], fx))), depth);
    // src = P @ [X, Y, Z, 1]                              (inverse_warp.py:152)
    const float* P = pr + 9;
    const float sx = fmaf(P[3], 1.0f, fmaf(P[2], Z, fmaf(P[1], Y, __fmul_rn(P[0], X))));
    const float sy = fmaf(P[7], 1.0f, fmaf(P[6], Z, fmaf(P[5], Y, __fmul_rn(P[4], X))));
    const float sz = fmaf(P[11], 1.0f, fmaf(P[10], Z, fmaf(P[9], Y, __fmul_rn(P[8], X))));
    if (flow && c == 0) {
        const float den = __fadd_rn(sz, 1e-7f);
        const size_t o = ((size_t)b * 2) * h * w + (size_t)y * w + x;
        flow[o] = __fsub_rn(__fdiv_rn(sx, den), fx);
        flow[o + (size_t)h * w] = __fsub_rn(__fdiv_rn(sy, den), fy);
    }
    if (new_disp)
        new_disp[(((size_t)b * C_total + c_off + c) * h + y) * w + x] = __fdiv_rn(bf, __fadd_rn(sz, 1e-5f));
}

// project_to_3d drop-in: depth in, flow of every channel ([B,2C,h,w], channel pairs) + triangular depth out.
__global__ void __launch_bounds__(256)
project_depth_kernel(const float* __restrict__ depth_in, const float* __restrict__ params, float* __restrict__ flow,
                     float* __restrict__ tri, int C, int h, int w, long long total) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % w);
    long long t = i / w;
    const int y = (int)(t % h);
    t /= h;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const float* pr = params + b * NPARAM;
    const float depth = __ldg(depth_in + i);
    const float fx = (float)x, fy = (float)y;
    const float X = __fmul_rn(fmaf(pr[2], 1.0f, fmaf(pr[1], fy, __fmul_rn(pr[0], fx))), depth);
    const float Y = __fmul_rn(fmaf(pr[5], 1.0f, fmaf(pr[4], fy, __fmul_rn(pr[3], fx))), depth);
    const float Z = __fmul_rn(fmaf(pr[8], 1.0f, fmaf(pr[7], fy, __fmul_rn(pr[6], fx))), depth);
    const float* P = pr + 9;
    const float sx = fmaf(P[3], 1.0f, fmaf(P[2], Z, fmaf(P[1], Y, __fmul_rn(P[0], X))));
    const float sy = fmaf(P[7], 1.0f, fmaf(P[6], Z, fmaf(P[5], Y, __fmul_rn(P[4], X))));
    const float sz = fmaf(P[11], 1.0f, fmaf(P[10], Z, fmaf(P[9], Y, __fmul_rn(P[8], X))));
    if (flow) {
        const float den = __fadd_rn(sz, 1e-7f);
        const size_t o = (((size_t)b * C + c) * 2) * h * w + (size_t)y * w + x;
        flow[o] = __fsub_rn(__fdiv_rn(sx, den), fx);
        flow[o + (size_t)h * w] = __fsub_rn(__fdiv_rn(sy, den), fy);
    }
    if (tri) tri[i] = sz;
}

// stage 1 of the batch-global mean: fixed-shape partial sums (deterministic order).
__global__ void __launch_bounds__(256)
metric_partial_kernel(const float* __restrict__ pd, int C, int hw, long long total, double* __restrict__ partial) {
    __shared__ double sm[256];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long b = i / hw, p = i % hw;
        s += (double)__ldg(pd + (b * C) * hw + p);
    }
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sm[threadIdx.x] += sm[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(256)
metric_apply_kernel(const float* __restrict__ pd, float* __restrict__ metric, const double* __restrict__ partial,
                    int nparts, int C, int hw, long long total) {
    __shared__ float mean_s;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < nparts; ++k) s += partial[k];
        mean_s = (float)(s / (double)total);
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const long long b = i / hw, p = i % hw;
    const float v = __fsub_rn(__ldg(pd + (b * C) * hw + p), mean_s);
    metric[i] = fminf(fmaxf(v, -50.0f), 50.0f);
}

// forward softmax splat: one thread per source pixel, all channels + the normaliser.
__global__ void __launch_bounds__(256)
splat_accumulate_kernel(const float* __restrict__ x, const float* __restrict__ flow, const float* __restrict__ metric,
                        float* __restrict__ acc, int C, int h, int w, long long total) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int px = (int)(i % w);
    long long t = i / w;
    const int py = (int)(t % h);
    const int b = (int)(t / h);
    const size_t hw = (size_t)h * w;
    const float ox = __fadd_rn((float)px, __ldg(flow + ((size_t)b * 2) * hw + (size_t)py * w + px));
    const float oy = __fadd_rn((float)py, __ldg(flow + ((size_t)b * 2 + 1) * hw + (size_t)py * w + px));
    if (!(fabsf(ox) < 1e9f) || !(fabsf(oy) < 1e9f)) return;  // NaN / inf targets land nowhere
    const int nwx = (int)floorf(ox), nwy = (int)floorf(oy);
    const float wnw = __fmul_rn(__fsub_rn((float)(nwx + 1), ox), __fsub_rn((float)(nwy + 1), oy));
    const float wne = __fmul_rn(__fsub_rn(ox, (float)nwx), __fsub_rn((float)(nwy + 1), oy));
    const float wsw = __fmul_rn(__fsub_rn((float)(nwx + 1), ox), __fsub_rn(oy, (float)nwy));
    const float wse = __fmul_rn(__fsub_rn(ox, (float)nwx), __fsub_rn(oy, (float)nwy));
    const float e = expf(__ldg(metric + (size_t)b * hw + (size_t)py * w + px));
    const bool x0ok = nwx >= 0 && nwx < w, x1ok = nwx + 1 >= 0 && nwx + 1 < w;
    const bool y0ok = nwy >= 0 && nwy < h, y1ok = nwy + 1 >= 0 && nwy + 1 < h;
    for (int c = 0; c <= C; ++c) {
        const float v = (c < C) ? __fmul_rn(__ldg(x + ((size_t)b * C + c) * hw + (size_t)py * w + px), e) : e;
        float* a = acc + ((size_t)b * (C + 1) + c) * hw;
        if (x0ok && y0ok) atomicAdd(a + (size_t)nwy * w + nwx, __fmul_rn(v, wnw));
        if (x1ok && y0ok) atomicAdd(a + (size_t)nwy * w + nwx + 1, __fmul_rn(v, wne));
        if (x0ok && y1ok) atomicAdd(a + (size_t)(nwy + 1) * w + nwx, __fmul_rn(v, wsw));
        if (x1ok && y1ok) atomicAdd(a + (size_t)(nwy + 1) * w + nwx + 1, __fmul_rn(v, wse));
    }
}

__global__ void __launch_bounds__(256)
splat_normalise_kernel(const float* __restrict__ acc, float* __restrict__ out, int C, int hw, long long total) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int p = (int)(i % hw);
    long long t = i / hw;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const float n = __ldg(acc + ((size_t)b * (C + 1) + C) * hw + p);
    out[i] = __fdiv_rn(__ldg(acc + ((size_t)b * (C + 1) + c) * hw + p), __fadd_rn(n, 1e-22f));
}

}  // namespace tstereo

using namespace tstereo;

extern "C" {

int tstereo_pose_prep(const float* K, const float* T_now, const float* inv_T_prev, const float* baseline,
                      float factor, float* params, int B, void* stream) {
    TS_REQUIRE(K && T_now && inv_T_prev && baseline && params, "pose_prep: null pointer");
    TS_REQUIRE(B > 0 && factor > 0.f, "pose_prep: bad sizes");
    pose_prep_kernel<<<cdiv(B, 32), 32, 0, (cudaStream_t)stream>>>(K, T_now, inv_T_prev, baseline, factor, params, B);
    return check_launch("pose_prep");
}

int tstereo_reproject_disp(const float* disp, const float* params, float* flow, float* new_disp, int C_total,
                           int c_off, int B, int C, int h, int w, void* stream) {
    TS_REQUIRE(disp && params, "reproject_disp: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0, "reproject_disp: bad sizes");
    TS_REQUIRE(!new_disp || (c_off >= 0 && c_off + C <= C_total), "reproject_disp: channel slice out of range");
    const long long total = (long long)B * C * h * w;
    reproject_kernel<<<(unsigned)cdivll(total, 256), 256, 0, (cudaStream_t)stream>>>(disp, params, flow, new_disp,
                                                                                   C_total, c_off, C, h, w, total);
    return check_launch("reproject_disp");
}

int tstereo_project_to_3d(const float* depth, const float* params, float* flow, float* tri, int B, int C, int h,
                          int w, void* stream) {
    TS_REQUIRE(depth && params, "project_to_3d: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0, "project_to_3d: bad sizes");
    const long long total = (long long)B * C * h * w;
    project_depth_kernel<<<(unsigned)cdivll(total, 256), 256, 0, (cudaStream_t)stream>>>(depth, params, flow, tri, C, h,
                                                                                       w, total);
    return check_launch("project_to_3d");
}

int tstereo_splat_metric(const float* pd, float* metric, float* scratch, int B, int C, int h, int w, void* stream) {
    TS_REQUIRE(pd && metric && scratch, "splat_metric: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0, "splat_metric: bad sizes");
    TS_REQUIRE(((size_t)scratch & 7) == 0, "splat_metric: scratch must be 8-byte aligned");
    const long long total = (long long)B * h * w;
    const int nparts = (int)(cdivll(total, 256) < 256 ? cdivll(total, 256) : 256);  // 256 doubles <= 1024 floats
    double* partial = reinterpret_cast<double*>(scratch);
    metric_partial_kernel<<<nparts, 256, 0, (cudaStream_t)stream>>>(pd, C, h * w, total, partial);
    int rc = check_launch("splat_metric(partial)");
    if (rc) return rc;
    metric_apply_kernel<<<(unsigned)cdivll(total, 256), 256, 0, (cudaStream_t)stream>>>(pd, metric, partial, nparts, C,
                                                                                      h * w, total);
    return check_launch("splat_metric(apply)");
}

long long tstereo_update_map_scratch_floats(int B, int h, int w, int M, int n_lm_out) {
    const long long hw = (long long)h * w;
    const long long nparts = (long long)B * ((hw + 255) / 256);
    return B * hw + 2 * nparts + 2 + B * (2 * M + 1) * hw + B * (n_lm_out + 1) * hw;
}

int tstereo_update_map(const float* prev_disp, int Hf, int Wf, const float* K, const float* T_now, const float* inv_T_prev,
                       const float* baseline, const float* mem_sample, const float* mem_cost, int M,
                       const float* local_map, int n_lm_in, int n_lm_out,
                       float* out_sample, float* out_cost, float* out_lm, float* scratch,
                       int B, int h, int w, void* stream) {
    TS_REQUIRE(prev_disp && K && T_now && inv_T_prev && baseline && scratch, "update_map: null pointer");
    TS_REQUIRE(B > 0 && B <= 65535 && h > 0 && w > 0 && Hf > 0 && Wf > 0, "update_map: bad sizes");
    TS_REQUIRE((mem_sample == nullptr) == (mem_cost == nullptr) && (mem_sample == nullptr) == (out_sample == nullptr) &&
                   (out_sample == nullptr) == (out_cost == nullptr), "update_map: memory inputs / outputs must be all set or all NULL");
    TS_REQUIRE(M >= 0 && M <= 4 && (mem_sample == nullptr || M > 0), "update_map: M=%d out of range", M);
    TS_REQUIRE(n_lm_out >= 0 && n_lm_out <= UM_MAXLM && n_lm_in >= 0 && n_lm_out <= n_lm_in + 1 && (n_lm_out == 0) == (out_lm == nullptr),
               "update_map: local map sizes (in %d, out %d)", n_lm_in, n_lm_out);
    TS_REQUIRE(n_lm_out <= 1 || local_map, "update_map: local_map missing");
    TS_REQUIRE(((size_t)scratch & 7) == 0, "update_map: scratch must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = h * w;
    const int gx = cdiv(hw, 256);
    const int nparts = B * gx;
    // scratch: partial sums (doubles) | pd | acc_mem | acc_lm
    double* partial = reinterpret_cast<double*>(scratch);
    float* pd = scratch + 2 * (size_t)nparts + 2;
    float* acc_mem = pd + (size_t)B * hw;
    const int cm = out_sample ? 2 * M + 1 : 0, cl = out_lm ? n_lm_out + 1 : 0;
    float* acc_lm = acc_mem + (size_t)B * (2 * M + 1) * hw;
    auto scale = [](int in_size, int out_size) { return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.0f; };
    dim3 grid(gx, B);
    um_prep_kernel<<<grid, 256, 0, st>>>(prev_disp, Hf, Wf, scale(Hf, h), scale(Wf, w), pd, partial, cm ? acc_mem : nullptr, cm,
                                         cl ? acc_lm : nullptr, cl, h, w);
    int rc = check_launch("update_map(prep)");
    if (rc) return rc;
    um_splat_kernel<<<grid, 256, 0, st>>>(pd, partial, nparts, (long long)B * hw, K, T_now, inv_T_prev, baseline, (float)Wf / (float)w,
                                          mem_sample, mem_cost, M, cm ? acc_mem : nullptr, local_map, n_lm_in, n_lm_out,
                                          cl ? acc_lm : nullptr, h, w);
    rc = check_launch("update_map(splat)");
    if (rc) return rc;
    um_norm_kernel<<<grid, 256, 0, st>>>(cm ? acc_mem : nullptr, cm, out_sample, out_cost, M, cl ? acc_lm : nullptr, cl, out_lm, hw);
    return check_launch("update_map(normalise)");
}

int tstereo_softsplat(const float* x, const float* flow, const float* metric, float* acc, float* out, int B, int C,
                      int h, int w, void* stream) {
    TS_REQUIRE(x && flow && metric && acc && out, "softsplat: null pointer");
    TS_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0, "softsplat: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hw = (size_t)h * w;
    cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)B * (C + 1) * hw * sizeof(float), st);
    if (e != cudaSuccess) {
        set_error("softsplat: cudaMemsetAsync: %s", cudaGetErrorString(e));
        return TSTEREO_E_CUDA;
    }
    const long long npix = (long long)B * h * w;
    splat_accumulate_kernel<<<(unsigned)cdivll(npix, 256), 256, 0, st>>>(x, flow, metric, acc, C, h, w, npix);
    int rc = check_launch("softsplat(accumulate)");
    if (rc) return rc;
    const long long total = npix * C;
    splat_normalise_kernel<<<(unsigned)cdivll(total, 256), 256, 0, st>>>(acc, out, C, (int)hw, total);
    return check_launch("softsplat(normalise)");
}

}  // extern "C"
