// Tensor-core (tcgen05, sm_100a) implicit-GEMM convolutions — the weight contractions of the aggregation
// (SURVEY.md §8 rows a4-a7, a9, a10, a12, a13):
//   * tstereo_conv_hw3_tc : (1,3,3) / 3x3 conv over (H,W), stride 1, dilation 1|2
//   * tstereo_conv_d_tc   : (k,1,1) conv along D, k = 3|5, stride 1|2, dilation 1|2, or transposed stride 2
//
// ref: architecture/modeling/layers/basic_layers.py:194-235 (Conv3d: conv -> BN -> act), :340-388 (ConvTranspose3d)
//      as used by the separable pairs of aggregation/TemporalStereo/module.py:111-184 and the 2-D convs of
//      :300-353, 424-492.
//
// Precision: plain TF32 moves the regressed disparity by 0.7 px, so every product is
// error-compensated "3xTF32": a = a_hi + a_lo (both exactly representable in tf32), and
// D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi (~21 operand bits).  The kernel is bound by the shared-memory
// reads of the A operand (4 KB per M=128,K=8 MMA, ~44 cycles each, measured), so the two A_hi terms are ONE
// MMA of width 2N against [B_hi | B_lo] (two accumulator column blocks, summed when drained): 2 MMAs per
// (tap, chunk, M-tile).
//
// GEMM view:  M = 128 x MT pixel positions of one output plane, N = Cout (rounded up to 16),
// K = 8-channel chunks x taps.  A is staged K-major with 16 B per (position, 4 channels):
// smem A[part][khalf][position][4 ch] — UMMA canonical K-major SWIZZLE_NONE layout with SBO = 128 B
// (8 positions x 16 B, dense) and LBO = NPOS*16 B — so a tap is just a start-address offset into the staged
// positions:
//   hw3: positions are padded-linear q = y*PW + x with pitch PW = W + DIL (the DIL zero columns after each
//        row are the right padding of row y and the left padding of row y+1); tap (ky,kx) = offset
//        ky*DIL*PW + kx*DIL inside ONE staged tile with a halo of DIL*PW + DIL positions on both sides;
//   d  : every valid tap stages its own input plane (MT*128 positions each); tap t = offset t*MT*128.
//
//  warps 0-7 : producers — global NCDHW fp32 -> registers (4 channel loads per K-half) -> hi/lo split ->
//              one 16 B st.shared per (position, K-half, part); thread 0 also bulk-copies the chunk's
//              pre-split weights (cp.async.bulk + mbarrier complete_tx).  They also drain the per-chunk
//              TMEM accumulators into fp32 registers and run the epilogue (bias, activation, NCDHW stores).
//  warp 8    : TMEM alloc/dealloc; one elected lane issues the tcgen05.mma's of a chunk and
//              tcgen05.commit's the stage back to the producers and the accumulator buffer to the readers.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cstdint>

namespace tstereo {
namespace tc {

constexpr int NPROD = 256;             // producer threads (8 warps)
constexpr int NTHREADS = NPROD + 32;   // + MMA warp
constexpr int MAX_STAGES = 4;
constexpr int MAX_TAPS = 9;
constexpr size_t SMEM_MAX = 227 * 1024;

using namespace tcp;

enum Mode { MODE_HW3 = 0, MODE_D = 1 };

struct Params {
    const float* in;
    long long isB, isC, isD;
    float* out;
    long long osB, osC, osD;
    const float* wpack;     // [nchunk][tap T][khalf 2][part 2][N][4]  (rows n' = part*N + n of a 2N x 8 K-major matrix)
    const float* bias;      // [Cout] or null
    int mode, T;            // T = taps stored in wpack (9 | k)
    int Cin, Cout, H, W;
    int Din, Dout;          // planes of the input / output (hw3: equal)
    int dil, PW, NPOS, halo, nchunk, stages, act;
    int k_d, stride_d, transposed;   // MODE_D
    int tiles_per_plane;
};

// The tensor core accumulates into TMEM with truncation: measured on B200 the error of one long
// accumulation grows linearly with K with a bias towards zero (rms 3e-5 at K = 9*352, 20x the fp32 FMA
// chain; scripts/diag_tc.py).  So every 8-channel chunk accumulates from zero into one of two TMEM buffers
// and the producer warps add it into fp32 REGISTER accumulators (round-to-nearest) while the tensor core
// works on the next chunk in the other buffer.
template <int N, int MT>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_kernel(const Params p) {
    constexpr int TOTAL = 2 * N * MT;      // accumulator columns of one buffer: per M-tile [A*B_hi | A_hi*B_lo]
    constexpr int JT = (MT + 1) / 2;       // M-tiles per reader thread (tiles j = half + 2*jj)
    static_assert(JT * N <= 64, "register accumulators limited to 64 per thread");
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: [stages] x { A: part(2) x khalf(2) x NPOS x 16 B | B: tap(T) x khalf(2) x 2N x 16 B } | pos table | barriers | taps
    const uint32_t a_bytes = 4u * p.NPOS * 16u;
    const uint32_t b_bytes = (uint32_t)p.T * 2u * 2u * N * 16u;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    int* pos_tbl = reinterpret_cast<int*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes + (((size_t)p.NPOS * 4 + 15) & ~(size_t)15));
    uint64_t* full = bars;                       // [stages]   producers -> MMA
    uint64_t* empty = bars + MAX_STAGES;         // [stages]   MMA -> producers
    uint64_t* acc_full = bars + 2 * MAX_STAGES;  // [2]        MMA -> readers (chunk accumulated)
    uint64_t* acc_empty = acc_full + 2;          // [2]        readers -> MMA (buffer drained)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    int* tap_tbl = reinterpret_cast<int*>(tmem_slot + 2);   // [0] = ntap, [1] = staged positions, (a_off, w_tap) pairs, input planes

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int plane = blockIdx.y;
    const int b = plane / p.Dout, dout = plane % p.Dout;
    const int q0 = blockIdx.x * MT * 128;                     // first output position of this CTA
    const float* in_b = p.in + (long long)b * p.isB;
    constexpr uint32_t ncols = (2 * TOTAL <= 32) ? 32 : (2 * TOTAL <= 64) ? 64 : (2 * TOTAL <= 128) ? 128 : (2 * TOTAL <= 256) ? 256 : 512;

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], NPROD + 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], NPROD);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // tap list: (offset of the tap's first row inside the staged positions, weight tap index)
        int nt = 0;
        if (p.mode == MODE_HW3) {
            for (int t = 0; t < 9; ++t) {
                const int ky = t / 3 - 1, kx = t % 3 - 1;
                tap_tbl[2 + 2 * nt] = p.halo + ky * p.dil * p.PW + kx * p.dil;
                tap_tbl[3 + 2 * nt] = t;
                ++nt;
            }
            tap_tbl[1] = p.NPOS;
        } else {
            for (int kd = 0; kd < p.k_d; ++kd) {
                int din;
                if (p.transposed) {                 // dout = din*2 - 1 + kd
                    const int t2 = dout + 1 - kd;
                    din = (t2 >= 0 && (t2 & 1) == 0) ? (t2 >> 1) : -1;
                } else {
                    din = dout * p.stride_d - p.dil * (p.k_d / 2) + kd * p.dil;
                }
                if (din >= 0 && din < p.Din) {
                    tap_tbl[2 + 2 * nt] = nt * MT * 128;
                    tap_tbl[3 + 2 * nt] = kd;
                    tap_tbl[2 + 2 * MAX_TAPS + nt] = din;
                    ++nt;
                }
            }
            tap_tbl[1] = nt * MT * 128;
        }
        tap_tbl[0] = nt;
    }
    if (warp == NPROD / 32) {   // MMA warp allocates TMEM
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const int ntap = tap_tbl[0];
    const int npos = tap_tbl[1];               // staged positions actually used (<= p.NPOS)
    // position table: element offset (relative to in_b + c*isC) of staged position i, or -1 (zero)
    const int HW = p.H * p.W;
    for (int i = tid; i < npos; i += NTHREADS) {
        int off = -1;
        if (p.mode == MODE_HW3) {
            const int q = q0 - p.halo + i;
            if (q >= 0) {
                const int y = q / p.PW, x = q - y * p.PW;
                if (y < p.H && x < p.W) off = dout * (int)p.isD + y * p.W + x;
            }
        } else {
            const int t = i / (MT * 128), q = q0 + (i - t * MT * 128);
            if (q < HW) off = tap_tbl[2 + 2 * MAX_TAPS + t] * (int)p.isD + q;
        }
        pos_tbl[i] = off;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (ntap == 0) {
        // no input plane contributes (cannot happen for the supported k / stride / padding combinations)
    } else if (warp < NPROD / 32) {
        // ===================== producers / accumulator readers =====================
        const int quarter = warp & 3;                 // TMEM lanes this warp may read
        const int half = warp >> 2;                   // which M-tiles it drains
        float acc[JT][N];
#pragma unroll
        for (int jj = 0; jj < JT; ++jj)
#pragma unroll
            for (int n = 0; n < N; ++n) acc[jj][n] = 0.f;

        auto drain = [&](int kk) {
            const int bf = kk & 1;
            mbar_wait(&acc_full[bf], (uint32_t)(kk >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int jj = 0; jj < JT; ++jj) {
                const int j = half + 2 * jj;
                if (j < MT) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(bf * TOTAL + j * 2 * N);
#pragma unroll
                    for (int c0 = 0; c0 < 2 * N; c0 += 16) {
                        uint32_t r[16];
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                            : "r"(taddr + (uint32_t)c0));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int c = 0; c < 16; ++c) acc[jj][(c0 + c) % N] += __uint_as_float(r[c]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[bf]);
        };

        for (int k = 0; k < p.nchunk; ++k) {
            const int s = k % p.stages;
            const uint32_t ph = (uint32_t)(k / p.stages) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            uint8_t* st_base = smem + (size_t)s * stage_bytes;
            if (tid == 0) {
                mbar_arrive_expect_tx(&full[s], b_bytes);
                bulk_g2s(st_base + a_bytes, p.wpack + (size_t)k * (b_bytes / 4), b_bytes, &full[s]);
            }
            const uint32_t a_hi = smem_u32(st_base);                        // part 0 (hi): khalf 0, khalf 1
            const uint32_t a_lo = a_hi + 2u * p.NPOS * 16u;                // part 1 (lo)
            const uint32_t khalf = (uint32_t)p.NPOS * 16u;
            const int c0 = k * 8;
            const float* src = in_b + (long long)c0 * p.isC;
            // PB positions per thread per batch: all PB*8 global loads are issued before the first
            // conversion, so one memory latency covers the whole batch
            constexpr int PB = 4;
            for (int i0 = tid; i0 < npos; i0 += NPROD * PB) {
                float v[PB][8];
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const int i = i0 + u * NPROD;
                    const int off = (i < npos) ? pos_tbl[i] : -1;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        v[u][c] = (off >= 0 && c0 + c < p.Cin) ? __ldg(src + (long long)c * p.isC + off) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const int i = i0 + u * NPROD;
                    if (i < npos) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            hi[c] = f2tf32(v[u][c]);
                            lo[c] = f2tf32(v[u][c] - __uint_as_float(hi[c]));
                        }
                        const uint32_t o = (uint32_t)i * 16u;
                        sts128(a_hi + o, hi[0], hi[1], hi[2], hi[3]);
                        sts128(a_hi + khalf + o, hi[4], hi[5], hi[6], hi[7]);
                        sts128(a_lo + o, lo[0], lo[1], lo[2], lo[3]);
                        sts128(a_lo + khalf + o, lo[4], lo[5], lo[6], lo[7]);
                    }
                }
            }
            fence_proxy_async();          // generic-proxy st.shared -> visible to the tensor core (async proxy)
            mbar_arrive(&full[s]);
            if (k >= 1) drain(k - 1);     // overlaps with the MMAs of chunk k
        }
        drain(p.nchunk - 1);

        // ===================== epilogue: bias + activation from the register accumulators =====================
        float* out_pl = p.out + (long long)b * p.osB + (long long)dout * p.osD;
#pragma unroll
        for (int jj = 0; jj < JT; ++jj) {
            const int j = half + 2 * jj;
            if (j < MT) {
                const int q = q0 + j * 128 + quarter * 32 + lane;
                const int y = q / p.PW, x = q - y * p.PW;
                if (y < p.H && x < p.W) {
                    float* o = out_pl + (size_t)y * p.W + x;
#pragma unroll
                    for (int n = 0; n < N; ++n)
                        if (n < p.Cout) {
                            const float bv = p.bias ? __ldg(p.bias + n) : 0.f;
                            o[(long long)n * p.osC] = apply_act(acc[jj][n] + bv, p.act);
                        }
                }
            }
        }
    } else {
        // ===================== MMA issuer =====================
        // instruction descriptor: D fp32, A/B tf32, both K-major, N (or 2N), M = 128
        constexpr uint32_t idesc_n = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint32_t idesc_2n = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_lbo = (uint32_t)p.NPOS * 16u, b_lbo = 2u * N * 16u;
        for (int k = 0; k < p.nchunk; ++k) {
            const int s = k % p.stages;
            const uint32_t ph = (uint32_t)(k / p.stages) & 1u;
            const int bf = k & 1, u = k >> 1;
            mbar_wait(&full[s], ph);
            if (u >= 1) mbar_wait(&acc_empty[bf], (uint32_t)(u - 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t st_base = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t a_part[2] = {st_base, st_base + 2u * p.NPOS * 16u};
                const uint32_t b_base = st_base + a_bytes;
#pragma unroll 1
                for (int j = 0; j < MT; ++j) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(bf * TOTAL + j * 2 * N);
#pragma unroll 1
                    for (int t = 0; t < ntap; ++t) {
                        const uint32_t aoff = (uint32_t)(tap_tbl[2 + 2 * t] + j * 128) * 16u;
                        const uint64_t bd = make_desc(b_base + (uint32_t)(tap_tbl[3 + 2 * t] * 2 * 2 * N * 16), b_lbo, 128u);
                        // [D1 | D2] (+)= A_hi * [B_hi | B_lo]   (first MMA of the chunk overwrites both blocks)
                        tc_mma_tf32(d_tmem, make_desc(a_part[0] + aoff, a_lbo, 128u), bd, idesc_2n, t != 0 ? 1u : 0u);
                        // D1 += A_lo * B_hi
                        tc_mma_tf32(d_tmem, make_desc(a_part[1] + aoff, a_lbo, 128u), bd, idesc_n, 1u);
                    }
                }
                tc_commit(&empty[s]);        // stage free once these MMAs have read it
                tc_commit(&acc_full[bf]);    // chunk accumulated
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == NPROD / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// shared memory of a launch: stages x (A + B) + position table + barriers / tap table
static size_t smem_need(int stages, long long npos, int T, int N) {
    return (size_t)stages * ((size_t)npos * 64 + (size_t)T * 64 * N) + (size_t)((npos * 4 + 15) & ~15ll) + 384;
}

// picks MT in {1,2,4,8} (ceil(MT/2)*N <= 64 register accumulators per thread) and the stage count; launches.
static int launch(Params& p, int N, long long positions_per_plane, int planes, int halo, int taps_staged, cudaStream_t st,
                  const char* what) {
    int best_mt = 0, best_stages = 2;
    double best_cost = 1e30;
    for (int mt = 1; mt <= 8; mt *= 2) {
        if (((mt + 1) / 2) * N > 64) continue;
        const long long npos = (long long)taps_staged * mt * 128 + 2 * halo;
        if (smem_need(2, npos, p.T, N) > SMEM_MAX) continue;
        int stages = 2;
        while (stages < MAX_STAGES && smem_need(stages + 1, npos, p.T, N) <= SMEM_MAX) ++stages;
        const long long tiles = (positions_per_plane + mt * 128 - 1) / (mt * 128);
        const long long waves = (tiles * planes + 147) / 148;
        // cost model: waves x (outputs + halo re-staging + fixed per-CTA cost); fewer, fuller waves win
        const double cost = (double)waves * ((double)mt * 128 + 2.0 * halo * 0.6 + 96.0);
        if (cost < best_cost) {
            best_cost = cost;
            best_mt = mt;
            best_stages = stages;
        }
    }
    TS_REQUIRE(best_mt > 0, "%s: tile does not fit shared memory (row pitch %d)", what, p.PW);
    p.stages = best_stages;
    p.halo = halo;
    p.NPOS = taps_staged * best_mt * 128 + 2 * halo;
    p.tiles_per_plane = (int)((positions_per_plane + best_mt * 128 - 1) / (best_mt * 128));
    const size_t smem_bytes = smem_need(p.stages, p.NPOS, p.T, N);
    dim3 grid(p.tiles_per_plane, planes);
    bool launched = false;
#define TS_TC(NN, MM)                                                                                               \
    if (N == NN && best_mt == MM) {                                                                                 \
        auto kern = conv_tc_kernel<NN, MM>;                                                                         \
        static bool attr_done = false;                                                                              \
        if (!attr_done) {                                                                                           \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX); \
            if (e != cudaSuccess) {                                                                                 \
                set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));                             \
                return TSTEREO_E_CUDA;                                                                              \
            }                                                                                                       \
            attr_done = true;                                                                                       \
        }                                                                                                           \
        kern<<<grid, NTHREADS, smem_bytes, st>>>(p);                                                                \
        launched = true;                                                                                            \
    }
    TS_TC(16, 1) TS_TC(16, 2) TS_TC(16, 4) TS_TC(16, 8)
    TS_TC(32, 1) TS_TC(32, 2) TS_TC(32, 4)
    TS_TC(48, 1) TS_TC(48, 2)
    TS_TC(64, 1) TS_TC(64, 2)
#undef TS_TC
    TS_REQUIRE(launched, "%s: no kernel instance for N=%d MT=%d", what, N, best_mt);
    return check_launch(what);
}

}  // namespace tc
}  // namespace tstereo

using namespace tstereo;

extern "C" {

long long tstereo_conv_tc_wpack_floats(int Cin, int Cout, int taps) {
    const long long N = (Cout + 15) / 16 * 16, nchunk = (Cin + 7) / 8;
    return nchunk * taps * 2 * 2 * N * 4;
}

int tstereo_conv_hw3_tc(const float* in, long long isB, long long isC, long long isD,
                        float* out, long long osB, long long osC, long long osD,
                        const float* wpack, const float* bias,
                        int B, int Cin, int Cout, int D, int H, int W,
                        int dilation, int act, void* stream) {
    TS_REQUIRE(in && out && wpack, "conv_hw3_tc: null pointer");
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Cout <= 64 && D > 0 && H > 0 && W > 0, "conv_hw3_tc: bad sizes (Cout <= 64)");
    TS_REQUIRE(dilation == 1 || dilation == 2, "conv_hw3_tc: dilation %d unsupported", dilation);
    TS_REQUIRE((long long)B * D <= 65535, "conv_hw3_tc: B*D exceeds grid.y");
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "conv_hw3_tc: packed weights must be 16-byte aligned");
    TS_REQUIRE((long long)D * isD + (long long)H * W < (1ll << 31), "conv_hw3_tc: plane offsets exceed 32 bits");
    tc::Params p = {};
    p.in = in; p.isB = isB; p.isC = isC; p.isD = isD;
    p.out = out; p.osB = osB; p.osC = osC; p.osD = osD;
    p.wpack = wpack; p.bias = bias;
    p.mode = tc::MODE_HW3; p.T = 9;
    p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.Din = D; p.Dout = D;
    p.dil = dilation; p.act = act;
    p.PW = W + dilation;
    p.nchunk = (Cin + 7) / 8;
    return tc::launch(p, (Cout + 15) / 16 * 16, (long long)H * p.PW, B * D, dilation * p.PW + dilation, 1,
                      (cudaStream_t)stream, "conv_hw3_tc");
}

int tstereo_conv_d_tc(const float* in, long long isB, long long isC, long long isD,
                      float* out, long long osB, long long osC, long long osD,
                      const float* wpack, const float* bias,
                      int B, int Cin, int Cout, int Din, int Dout, int H, int W,
                      int k, int stride, int dilation, int transposed, int act, void* stream) {
    TS_REQUIRE(in && out && wpack, "conv_d_tc: null pointer");
    TS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Cout <= 64 && Din > 0 && Dout > 0 && H > 0 && W > 0, "conv_d_tc: bad sizes (Cout <= 64)");
    TS_REQUIRE(k == 3 || k == 5, "conv_d_tc: k=%d unsupported", k);
    if (transposed) {
        TS_REQUIRE(k == 3 && Dout == 2 * Din, "conv_d_tc: transposed needs k=3, Dout=2*Din (got k=%d Din=%d Dout=%d)", k, Din, Dout);
    } else {
        TS_REQUIRE((stride == 1 || stride == 2) && (dilation == 1 || dilation == 2), "conv_d_tc: bad stride/dilation");
        TS_REQUIRE(Dout == (Din - 1) / stride + 1, "conv_d_tc: Dout=%d inconsistent with Din=%d stride=%d", Dout, Din, stride);
    }
    TS_REQUIRE((long long)B * Dout <= 65535, "conv_d_tc: B*Dout exceeds grid.y");
    TS_REQUIRE((((size_t)wpack) & 15) == 0, "conv_d_tc: packed weights must be 16-byte aligned");
    TS_REQUIRE((long long)Din * isD + (long long)H * W < (1ll << 31), "conv_d_tc: plane offsets exceed 32 bits");
    tc::Params p = {};
    p.in = in; p.isB = isB; p.isC = isC; p.isD = isD;
    p.out = out; p.osB = osB; p.osC = osC; p.osD = osD;
    p.wpack = wpack; p.bias = bias;
    p.mode = tc::MODE_D; p.T = k;
    p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.Din = Din; p.Dout = Dout;
    p.dil = dilation; p.act = act;
    p.PW = W;                                   // no padding columns: positions are the plane's pixels
    p.k_d = k; p.stride_d = stride; p.transposed = transposed;
    p.nchunk = (Cin + 7) / 8;
    return tc::launch(p, (Cout + 15) / 16 * 16, (long long)H * W, B * Dout, 0, k, (cudaStream_t)stream, "conv_d_tc");
}

}  // extern "C"
