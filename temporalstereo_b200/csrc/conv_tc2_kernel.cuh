// Tensor-core (tcgen05, sm_100a) 3x3 convolution over (H,W), second generation: 2-D tiles with the kx taps
// folded into the MMA's N dimension  (SURVEY.md §8 rows a4-a7, a9, a12, a13).
//
// ref: architecture/modeling/layers/basic_layers.py:194-235 (Conv3d: conv -> BN -> act) as used by the (1,3,3)
//      halves of aggregation/TemporalStereo/module.py:111-147 and the 2-D convs of :424-492.
//
// Why a second kernel: the first one (conv_tc.cu) issues one MMA pair per tap, i.e. the A operand (128
// positions x 8 channels) is read from shared memory 18 times per chunk, and its linear position tiles stage a
// halo of a full image row on both sides.  Here
//   * a CTA owns a TR x (32 - 2*dil) output tile; the staged input is (TR + 2*dil) rows x 32 columns, position
//     index = row*32 + col, so one warp-wide TMEM lane quarter is exactly one tile row;
//   * the three kx taps are columns of the B operand: per ky ONE product
//         P[pos][kx][co] (+)= A[pos + ky*dil*32][ch] * W[ky][kx][ch][co]
//     (N = 3*CP), and the output is  out[x] = P[x][kx=0] + P[x+dil][kx=1] + P[x+2*dil][kx=2]  — two warp
//     shuffles per channel in the epilogue.  A is read 6 times per chunk instead of 18;
//   * error compensation as before (operands split hi + lo): D[0:2N) += A_hi*[B_hi|B_lo],  D[0:N) += A_lo*B_hi;
//   * the tensor core's accumulate truncates, so TMEM accumulates only G chunks (3*8*G products
//     per term) before the producer warps add the partial sums into fp32 registers; the MMA warp moves on to
//     the next M-tile meanwhile (per-M-tile full/empty barriers instead of a second TMEM buffer);
//   * producers prefetch the next chunk's 8 channel rows into registers before converting the current one, and
//     the small variants run two CTAs per SM, so global-load latency overlaps the MMAs of the other CTA.
//
//   * two epilogue modes: ACC (above; any Cin) and DIRECT (Cin <= 8*G, i.e. one accumulation group): all three
//     3xTF32 terms accumulate into the SAME N columns (3 MMAs per ky), TMEM holds MT*N columns, and the epilogue
//     streams TMEM -> shuffle -> bias/act -> store without register accumulators, so two CTAs fit one SM even
//     for Cout = 32.
//
//   * stride-2 convolutions and stride-2 transposed convolutions run through the SAME kernel as stride-1 3x3
//     convolutions over a virtual tensor: the input of a stride-2 conv is read as its four (row, column) parity
//     phases stacked on the channel axis (chunk k -> phase k / cpp: element offset pr*Win + pc, row pitch 2*Win,
//     x step 2), and a transposed conv writes one output parity phase per launch (output row pitch / x step 2);
//     the host repacks the weights accordingly (taps that do not exist for a phase are zero).
//   * (k,1,1) convolutions along D use the same machinery with the k input planes as "phases" (chunk k -> plane
//     dout*stride + (k/cpp - k_d/2)*dil, zero when outside), one ky tap (nky = 1) and no halo (dil = 0).
//   * producer modes (template RAW): 0 = the register path above (default: fastest on B200); 1 = TMA (opt-in
//     TSTEREO_TC2_TMA=1, tf32 split, inputs whose rows are 16-byte aligned and x-dense): one elected lane of the MMA
//     warp issues `cp.async.bulk.tensor.5d` per chunk — an (x 32|36, y TR+2*dil, plane 1, c 8) box of the NCDHW input
//     whose origin is rounded down to a multiple of 4 pixels, out-of-image rows / columns / planes / channels
//     zero-filled by the TMA unit — into a ring of raw fp32 stages which the 8 producer warps read (conflict-free
//     LDS), split and store as the K-major operand; 2 = the same ring filled by per-thread `cp.async` (opt-in
//     TSTEREO_TC2_CPA=1, fp16 split, any width).  DESIGN.md §5 has the measurements (both are slower than mode 0).
//     3 = SPLIT (round 2, the fastest form): the input is an "S-format" tensor — the fp16 hi / lo halves of the fp32
//     activation, laid out [.. part][8-channel chunk][H][W][8] by the epilogue of the layer that produced it — so a
//     K-chunk of the A operand is one TMA box `cp.async.bulk.tensor.5d` straight into the K-major stage (zero padding by
//     the TMA unit's out-of-bounds fill): no producer warps at all; warp 9 issues the boxes + the weight bulk copy,
//     warp 8 the MMAs, warps 0-7 only drain / run the epilogue.  The split is computed ONCE per element by the
//     producing epilogue instead of once per consumer tile (halo included) in front of every MMA.
//   * F16 variant (template flag, the engine's default): operands split as fp16 hi + fp16 lo (22 mantissa bits,
//     same three product terms) and multiplied with `kind::f16`: K = 16 channels per MMA instead of 8, i.e. half the
//     MMAs and half the shared-memory operand traffic per channel (the kernel is shared-memory-pipe bound) at twice
//     the tensor rate.  Measured EPE vs fp32 on the oracle: 8.0e-6 px (3xTF32: 8.6e-6).  fp16 range: activations
//     must stay below 65504 in magnitude (the model's are O(1..100)).
//
//   * FOLD = 1 (fp16 form of the (k,1,1) convs): a single accumulator column block, N = CP, no shuffles in the epilogue.
//
//  warps 0-7 : producers (global NCDHW fp32 -> hi/lo split -> K-major SWIZZLE_NONE smem) + accumulator readers
//              + epilogue;  warp 8 : TMEM alloc, MMA issue (elect.sync), commits (and the TMA issue in RAW = 1).
#pragma once
#include "common.cuh"
#include "cost_coord.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

namespace tstereo {
namespace tc2 {

using namespace tcp;

constexpr int NPROD = 256;
constexpr int NTHREADS = NPROD + 32;
constexpr int nthreads(int raw) { return NPROD + 32 + (raw == 3 ? 32 : 0); }   // SPLIT: one more warp issues the TMA boxes
constexpr int MAX_STAGES = 8;
constexpr int BAR_BYTES = 640;   // barriers + TMEM slot (<= 384 B) + the CTA's bias / output-scale vectors (2 x 32 floats)
constexpr int MAX_MT = 4;
constexpr size_t SMEM_MAX = 227 * 1024;

struct Params {
    const float* in;
    long long isB, isD;
    float* out;
    long long osB, osD;
    int isC, osC;         // channel strides (elements); < 2^31, checked by the host
    const float* wpack;   // [nchunk][ky nky][khalf 2][row 2N][4], row = part*N + kx*CP + co
    const float* bias;    // [Cout] or null
    const float* oscale;  // [Cout] or null: per-output-channel multiplier of the accumulator (undoes the fp16 weight pre-scale)
    int Cin, Cout, H, W, D;       // Cin = real channels per input phase; H, W = grid of the (virtual) stride-1 conv
    int Hin, Win;                 // real input plane (= H, W unless the input is phase-decomposed)
    int isY, isX;                 // input row pitch / x step in elements (W, 1 | 2*Win, 2)
    int osY, osX;                 // output row pitch / x step (W, 1 | 2*W.., 2 for one phase of a transposed conv)
    int cpp;                      // chunks per input phase = ceil(Cin / 8); nchunk = phases * cpp
    int dil, act, nchunk, G, stages, tiles_x;
    int tiles_y, nplanes;         // tiles along y; B * D planes: a launch covers tiles_x * tiles_y * nplanes tiles
    int persist;                  // MULTI instances: grid.x CTAs walk the tiles with stride grid.x (0: one tile per CTA)
    int rs;                       // TMA variant: raw fp32 stages in flight
    int bw;                       // TMA variant: box width in elements (32, or 36 when the halo shifts the 16-byte aligned origin)
    int nbatch;                   // batch size (extent of the TMA view's last dimension)
    int fold;                     // 3: kx taps folded into N (3x3 forms); 1: single column block ((k,1,1) form)
    int half;                     // operands as fp16 hi + lo (kind::f16, 16 channels per MMA) instead of tf32 hi + lo
    int terms;                    // 3: error-compensated hi+lo products; 1 (fp16, DIRECT only): A_hi * B_hi alone (11-bit operands)
    int nky;                      // ky taps of the virtual conv: 3, or 1 for the (k,1,1) convs along D
    int kd, dstride, ddil, Din, dtrans;   // kd > 0: phases are input planes of a conv along D (p.D = Dout)
    // ---- fused cost volume -> first conv (template FUSE != 0; SURVEY.md §8d "fused path").  The virtual input is the raw
    // cost volume, never materialised: chunks [0, wchunks) are built from the feature maps (`in` = right features
    // [B, C, H, W]; FUSE 1: R warped by the per-pixel candidates `smp` [B, D, H, W]; FUSE 2: -(L - R shifted by d)^2 with
    // `left` [B, C, H, W]), chunks [wchunks, nchunk) are read from the compact group-cost volume `in2` [B, C2, D, H, W].
    // FUSE 1 adds `add` [B, Cout, H, W] (the D-invariant left-feature half of the conv, computed once) before the bias.
    const float* smp;
    const float* left;
    const float* in2;
    long long i2sB, i2sD;
    int i2sC, C2, wchunks;
    const float* add;
    long long asB;
    int asC;
    // ---- S-format ("split") tensors: fp16 hi / lo halves of an fp32 activation, [B][D][part][C8][H][W][8 channels]
    // (strides in fp16 elements; the [H][W][8] block is dense).  Input (template RAW = 3): read through the two TMA maps
    // (hi, lo), s_sx = 2: the four parity phases of a stride-2 conv through the map's element strides.  Output (any
    // instance): `outs` != null -> the epilogue also (or only: out == null) writes the split result for batches < s_nb.
    int s_in, s_sx;
    int wres;                     // SPLIT, resident CTAs: the whole weight image (nmma chunks) is loaded into shared memory once
    unsigned short* outs;
    long long ossB, ossD, ossP, ossC8;
    int s_parts, s_nb;
};

// cvt.rna.tf32.f32 without the NaN/Inf handling ptxas wraps around it (operands here are finite activations)
__device__ __forceinline__ uint32_t tf32_rna(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16 operands, fp32 accumulate), M = 128, K = 16
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor: D fp32, A/B fp16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }
// Activation of the epilogue, compile-time selected (the epilogue is instruction-bound).  SiLU uses
// ex2.approx / rcp.approx (~1e-6 relative); the full-precision expf + IEEE division of silu_f cost ~50
// instructions per output.
template <int ACT>
__device__ __forceinline__ float act_t(float x) {
    if constexpr (ACT == TSTEREO_ACT_SILU) {
        float e, r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
        return x * r;
    } else if constexpr (ACT == TSTEREO_ACT_RELU) {
        return fmaxf(x, 0.0f);
    } else {
        return x;
    }
}

// Where the epilogue of one lane (= one output position) writes: fp32 NC(D)HW and / or the S-format halves.
struct EpiOut {
    float* o;                 // fp32: address of channel 0 at this position, or null
    int osC;
    unsigned short* so;       // S-format: the hi vector (8 channels) of chunk 0 at this position, or null
    long long so_lo;          // element offset hi -> lo vector (0: hi only)
    long long so_c8;          // element stride between 8-channel chunks
    bool sok;                 // this lane writes the S-format (position inside the image, batch < s_nb)
};
// kx shift-sum + bias + activation + store of 8 channels of one tile row (lane = tile column):
// out[x] = P0[x] + P1[x + dil] + P2[x + 2*dil]
// FOLD = 3: the kx taps are columns of the accumulator (3x3 convs); FOLD = 1: one column block (the 1x1 / (k,1,1) form)
template <int ACT, int FOLD>
__device__ __forceinline__ void epi_store8(const float (&a0)[8], const float (&a1)[8], const float (&a2)[8], int dil, const EpiOut& eo,
                                           int c0, const float* bias, int Cout, bool ok, const float* osc,
                                           const float* add = nullptr, int asC = 0) {
    // bias / osc: the CTA's shared-memory copies (zero / one padded to 32 channels)
    float bv[8], sv[8];
    *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(bias + c0);
    *reinterpret_cast<float4*>(bv + 4) = *reinterpret_cast<const float4*>(bias + c0 + 4);
    *reinterpret_cast<float4*>(sv) = *reinterpret_cast<const float4*>(osc + c0);
    *reinterpret_cast<float4*>(sv + 4) = *reinterpret_cast<const float4*>(osc + c0 + 4);
    float res[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float acc = a0[c];
        if constexpr (FOLD == 3) {
            acc += __shfl_down_sync(0xffffffffu, a1[c], dil);
            acc += __shfl_down_sync(0xffffffffu, a2[c], 2 * dil);
        }
        const bool live = ok && c0 + c < Cout;
        float pre = fmaf(acc, sv[c], bv[c]);
        if (add) pre += live ? __ldg(add + (long long)(c0 + c) * asC) : 0.f;
        res[c] = act_t<ACT>(pre);
    }
    if (eo.o) {                 // warp-uniform branch: an S-format-only layer skips the eight predicated stores and their address math
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (ok && c0 + c < Cout) eo.o[(c0 + c) * eo.osC] = res[c];
    }
    if (eo.so && c0 < Cout) {   // the S-format of this 8-channel chunk (chunks beyond Cout do not exist in the output): hi = fp16(x), lo = fp16(x - hi); padding channels are 0
        uint32_t hi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) hi[i] = pack_h2(res[2 * i], res[2 * i + 1]);
        unsigned short* sp = eo.so + (long long)(c0 >> 3) * eo.so_c8;
        if (eo.sok) stg128(sp, hi[0], hi[1], hi[2], hi[3]);
        if (eo.so_lo) {
            uint32_t lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 hf = unpack_h2(hi[i]);
                lo[i] = pack_h2(res[2 * i] - hf.x, res[2 * i + 1] - hf.y);
            }
            if (eo.sok) stg128(sp + eo.so_lo, lo[0], lo[1], lo[2], lo[3]);
        }
    }
}
template <int ACT, int CP, int FOLD>
__device__ __forceinline__ void epi_direct(uint32_t taddr, int dil, const EpiOut& eo, const float* bias, int Cout, bool ok,
                                           const float* osc) {
#pragma unroll
    for (int c0 = 0; c0 < CP; c0 += 8) {
        uint32_t r0[8], r1[8], r2[8];
        tmem_ld8(taddr + (uint32_t)c0, r0);
        if constexpr (FOLD == 3) {
            tmem_ld8(taddr + (uint32_t)(CP + c0), r1);
            tmem_ld8(taddr + (uint32_t)(2 * CP + c0), r2);
        }
        tmem_ld_wait();
        float a0[8], a1[8], a2[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            a0[c] = __uint_as_float(r0[c]);
            a1[c] = FOLD == 3 ? __uint_as_float(r1[c]) : 0.f;
            a2[c] = FOLD == 3 ? __uint_as_float(r2[c]) : 0.f;
        }
        epi_store8<ACT, FOLD>(a0, a1, a2, dil, eo, c0, bias, Cout, ok, osc);
    }
}
template <int ACT, int CP, int FOLD>
__device__ __forceinline__ void epi_acc(const float* acc, int dil, const EpiOut& eo, const float* bias, int Cout, bool ok,
                                        const float* osc, const float* add = nullptr, int asC = 0) {
#pragma unroll
    for (int c0 = 0; c0 < CP; c0 += 8) {
        float a0[8], a1[8], a2[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            a0[c] = acc[c0 + c];
            a1[c] = FOLD == 3 ? acc[(FOLD == 3 ? CP : 0) + c0 + c] : 0.f;
            a2[c] = FOLD == 3 ? acc[(FOLD == 3 ? 2 * CP : 0) + c0 + c] : 0.f;
        }
        epi_store8<ACT, FOLD>(a0, a1, a2, dil, eo, c0, bias, Cout, ok, osc, add, asC);
    }
}

template <int CP, int MT, bool DIRECT, int FOLD = 3>
struct Cfg {
    static constexpr int N = FOLD * CP;                // columns of one part block: [kx][co]
    static constexpr int N2 = (N + 15) / 16 * 16;      // width of an N-wide MMA (M = 128 needs N % 16 == 0)
    static constexpr int TS = DIRECT ? N2 : 2 * N;     // TMEM columns of one M-tile: ACC [A*B_hi (N) | A_hi*B_lo (N)]
    static constexpr int COLS = MT * TS;
    static constexpr int JT = (MT + 1) / 2;            // M-tiles drained per thread (tiles j = half + 2*jj)
    static constexpr int MINB = (DIRECT || JT * N <= 48) ? 2 : 1;
    static constexpr int RPW = (4 * MT + 4 + 7) / 8;   // staged rows per producer warp (dil <= 2)
    static constexpr uint32_t NCOLS = COLS <= 32 ? 32 : COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
    static constexpr uint32_t B_BYTES = 3u * 2u * 2u * N * 16u;
    static_assert(2 * N <= 256 && COLS <= 512, "tile does not fit one MMA / TMEM");
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

constexpr int MAX_RS = 8;
// The fused cost producer holds 48 gathers (3 staged rows x 8 channels x 2 taps) per thread in flight next to the 48
// accumulators of the ACC form: under the 96-register cap of two CTAs per SM ptxas serialises the gathers (and spills)
// (measured on B200, precise level B = 8: one CTA per SM with 168 registers, no spills and double-buffered gathers runs
// 654 us against 545 us for two spilling CTAs — the second CTA's eight producer warps matter more; default off)
#ifndef TS_FUSE_ONE_CTA
#define TS_FUSE_ONE_CTA 0
#endif
constexpr bool FUSE_ONE_CTA = TS_FUSE_ONE_CTA != 0;

// RAW: 0 = chunks are prefetched into registers; 1 = TMA boxes into a raw fp32 ring; 2 = per-thread cp.async (zero-fill)
// into the same ring, `rs` chunks deep — for the small, latency-bound layers (any width / alignment)
// FUSE: 0 = the input is a tensor; 1 / 2 = the input is the warp / shift cost volume built on the fly (see Params)
template <int CP, int MT, bool DIRECT, int RAW, bool F16, int FOLD, int FUSE = 0>
__global__ void __launch_bounds__(nthreads(RAW), (FUSE != 0 && FUSE_ONE_CTA) ? 1 : Cfg<CP, MT, DIRECT, FOLD>::MINB)
conv_tc2_kernel(const Params p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_lo) {
    static_assert(FUSE == 0 || (RAW == 0 && FOLD == 3 && !DIRECT), "the fused cost producer is the register path of the 3x3 ACC form");
    static_assert(RAW != 3 || F16, "the S-format input is the fp16 split");
    constexpr bool TMA = RAW == 1, CPA = RAW == 2, SPLIT = RAW == 3;
    using C = Cfg<CP, MT, DIRECT, FOLD>;
    constexpr int N = C::N, JT = C::JT, RPW = C::RPW;
    extern __shared__ __align__(128) uint8_t smem[];
    const int SR = 4 * MT + 2 * p.dil;                 // staged rows
    const uint32_t NPOS = (uint32_t)SR * 32u;
    // [part 2][khalf 2][NPOS][16 B]; the single-term SPLIT form stages the hi part only
    const uint32_t a_bytes = ((SPLIT && p.terms != 3) ? 2u : 4u) * NPOS * 16u;
    const uint32_t b_bytes = (uint32_t)p.nky * (C::B_BYTES / 3u);   // weight bytes of one chunk actually used
    const bool wres = SPLIT && p.wres;                              // weights resident: a stage holds the A operand only
    const uint32_t stage_bytes = a_bytes + (wres ? 0u : C::B_BYTES);
    const int nmma_w = (p.nchunk + 1) / 2;
    uint8_t* wbase = smem + (size_t)p.stages * stage_bytes;         // [nmma][b_bytes] when resident
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + (wres ? (size_t)nmma_w * b_bytes : 0));
    uint64_t* full = bars;                         // [stages]  producers -> MMA
    uint64_t* empty = bars + MAX_STAGES;           // [stages]  MMA -> producers
    uint64_t* acc_full = bars + 2 * MAX_STAGES;    // [MT]      MMA -> readers (a group of G chunks accumulated)
    uint64_t* acc_empty = acc_full + MAX_MT;       // [MT]      readers -> MMA (M-tile drained)
    uint64_t* raw_full = acc_empty + MAX_MT;       // [rs]      TMA -> producers (raw fp32 chunk landed)
    uint64_t* raw_empty = raw_full + MAX_RS;       // [rs]      producers -> TMA issuer (raw stage read)
    uint64_t* wfull = raw_empty + MAX_RS;          // resident weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 384);   // [32] bias, 0 beyond Cout
    float* s_osc = s_bias + 32;                                                         // [32] accumulator scale, 1 when absent
    uint8_t* raw_base = reinterpret_cast<uint8_t*>(bars) + BAR_BYTES;    // TMA variant: [rs][c 8][row SR][x bw] fp32
    const uint32_t raw_bytes = 32u * (uint32_t)p.bw * (uint32_t)SR;   // p.bw = 32 for the cp.async ring

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int VW = 32 - 2 * p.dil;                     // valid output columns of a tile
    // MULTI (DIRECT register-producer instances): a CTA walks tiles t0, t0 + grid.x, ... — barriers, TMEM and the
    // operand ring are set up once, the first loads of tile t+1 are in flight while tile t's epilogue runs.  Every other
    // instance keeps one tile per CTA (grid = tiles of a plane x planes).
    constexpr bool MULTI = DIRECT && (RAW == 0 || RAW == 3) && FUSE == 0;
    const int tiles_pp = p.tiles_x * p.tiles_y;
    const int T = tiles_pp * p.nplanes;                // < 2^31 (checked by the host)
    const int t_first = MULTI ? (int)blockIdx.x : (int)blockIdx.y * tiles_pp + (int)blockIdx.x;
    const int t_stride = MULTI ? (int)gridDim.x : T;
    int b, d, y0, x0;
    auto set_tile = [&](int t) {
        const int plane = t / tiles_pp;
        const int rem = t - plane * tiles_pp;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        b = plane / p.D;
        d = plane - b * p.D;
        y0 = ty * 4 * MT;
        x0 = tx * VW;
    };
    set_tile(t_first);

    // the epilogue's per-channel constants, staged once: 16 predicated scalar loads per 8 channels and output row (each with
    // its descriptor re-materialisation: ncu r02, ~20 of the epilogue's ~47 instructions per channel) become four LDS.128
    if (tid < 32) {
        const bool cok = tid < p.Cout;
        s_bias[tid] = cok ? __ldg(p.bias + tid) : 0.f;
        s_osc[tid] = (cok && p.oscale) ? __ldg(p.oscale + tid) : 1.f;
    }
    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], SPLIT ? 1 : NPROD + 1);
            mbar_init(&empty[s], 1);
        }
        for (int j = 0; j < MT; ++j) {
            mbar_init(&acc_full[j], 1);
            mbar_init(&acc_empty[j], NPROD / 2);
        }
        if constexpr (TMA) {
            for (int i = 0; i < p.rs; ++i) {
                mbar_init(&raw_full[i], 1);
                mbar_init(&raw_empty[i], NPROD / 32);
            }
        }
        if constexpr (CPA)
            for (int i = 0; i < p.rs; ++i) mbar_init(&raw_full[i], NPROD);    // one cp.async arrival per producer thread
        if constexpr (SPLIT) mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NPROD / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::NCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // everything above (barriers, TMEM, the layer's constants) overlapped the previous kernel's tail; its outputs are
    // visible from here on (common.cuh pdl_sync)
    pdl_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int nmma = F16 ? (p.nchunk + 1) / 2 : p.nchunk;     // MMA chunks (16 | 8 channels); p.nchunk counts 8-channel units
    const int ngroups = (nmma + p.G - 1) / p.G;

// TMA variant: the MMA warp keeps `rs` chunk loads in flight; chunk kk -> (phase, channel chunk) -> box coordinates
    int t_issue = 0;
    auto issue_tma = [&]() {        // MMA warp only, converged
        const int kk = t_issue++;
        const int rsl = kk % p.rs;
        mbar_wait(&raw_empty[rsl], (((uint32_t)(kk / p.rs)) & 1u) ^ 1u);
        if (elect_one()) {
            const int phase = kk / p.cpp, kc = kk - phase * p.cpp;
            int dd = d;
            if (p.kd) {
                if (p.dtrans) {         // transposed k3 s2 p1 op1: dout = 2*din - 1 + tap
                    const int t2 = d + 1 - phase;
                    dd = (t2 >= 0 && (t2 & 1) == 0) ? (t2 >> 1) : -1;
                } else {
                    dd = d * p.dstride + (phase - p.kd / 2) * p.ddil;
                }
                if (dd < 0 || dd >= p.Din) dd = -1;       // outside: the TMA unit fills zeros
            }
            mbar_arrive_expect_tx(&raw_full[rsl], raw_bytes);
            // the box origin must be 16-byte aligned in global memory: start at the multiple of 4 below x0 - dil
            tma_load_5d(raw_base + (size_t)rsl * raw_bytes, &tmap, &raw_full[rsl], (x0 - p.dil) & ~3, y0 - p.dil, dd, kc * 8, b);
        }
        __syncwarp();
    };
    if (warp < NPROD / 32) {
        // ===================== producers / accumulator readers =====================
        const int quarter = warp & 3;                 // TMEM lane quarter = tile row inside an M-tile
        const int half = warp >> 2;                   // which M-tiles this warp drains
        const float* in_pl;
        // the staged rows of this warp: r = warp + 8*u; lane = staged column
        int off[RPW];               // element offset of (row, lane) in phase (0,0), or -1
        int edge;                   // bit u: row r's odd-row phase lies below the input; bit 31: same for the column
        int l_phase, l_kc;          // load cursor: input phase and chunk inside the phase
        auto setup_tile = [&]() {   // per-tile producer state from (b, d, y0, x0)
            in_pl = p.in + (long long)b * p.isB + (p.kd ? 0ll : (long long)d * p.isD);
            edge = 0;
            l_phase = 0;
            l_kc = 0;
            const int gx = x0 - p.dil + lane;
            if (p.isX * gx + 1 >= p.Win) edge |= 1 << 31;
#pragma unroll
            for (int u = 0; u < RPW; ++u) {
                const int r = warp + 8 * u;
                const int gy = y0 - p.dil + r;
                off[u] = (r < SR && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) ? gy * p.isY + gx * p.isX : -1;
                if (p.isX * gy + 1 >= p.Hin) edge |= 1 << u;
            }
        };
        if constexpr (!SPLIT) setup_tile();
        // fused cost producer: per staged position the right-feature tap (offset inside a channel plane, two weights)
        int woff[FUSE ? RPW : 1];
        float wa[FUSE ? RPW : 1], wb[FUSE == 1 ? RPW : 1];
        if constexpr (FUSE != 0) {
            const int gx = x0 - p.dil + lane;
#pragma unroll
            for (int u = 0; u < RPW; ++u) {
                const int r = warp + 8 * u;
                const int gy = y0 - p.dil + r;
                const bool inside = off[u] >= 0;
                const int yc = min(max(gy, 0), p.H - 1), xc = min(max(gx, 0), p.W - 1);
                if constexpr (FUSE == 1) {
                    const float dsp = __ldg(p.smp + ((long long)(b * p.D + d) * p.H + yc) * p.W + xc);
                    int xa;
                    warp_col(xc, dsp, p.W, inside, xa, wa[u], wb[u]);
                    woff[u] = warp_row(yc, p.H) * p.W + xa;
                } else {
                    const int xs = xc - d;                      // candidate d of the shift volume: R[x - d], zero for x < d
                    wa[u] = (inside && xs >= 0) ? 1.f : 0.f;
                    woff[u] = yc * p.W + max(xs, 0);
                }
            }
        }
        // NB = 2 (DIRECT register-producer instances: 80-91 registers, room under the 112 of two CTAs per SM): the loads of
        // unit k+1 are issued BEFORE unit k is converted, into the other register buffer — with one buffer a unit's loads
        // could only be issued after the previous unit's conversion and each unit paid a full memory latency, hidden by
        // nothing but the other warps (ncu r02: 21 % of the stall samples on the first F2FP after the loads).
        constexpr int NB = (RAW == 0 && ((RPW <= 2 && DIRECT && FUSE == 0) || (FUSE != 0 && FUSE_ONE_CTA))) ? 2 : 1;   // RPW = 3 (MT = 4) at two CTAs per SM: no room for 24 more registers
        float v[NB][RPW][8];
        unsigned v_ok[NB];              // bit u: row u of the unit held in v[.] is real data (else: zero it when it is packed)
#pragma unroll
        for (int i = 0; i < NB; ++i) v_ok[i] = 0xffffffffu;
        auto load_chunk = [&](auto BUFC) {
            constexpr int BUF = decltype(BUFC)::value;
            if constexpr (FUSE != 0) {
                v_ok[BUF] = 0xffffffffu;
                // loads at 32-bit offsets from an opaque 64-bit base (one IMAD.WIDE.U32 per address; left to itself ptxas keeps
                // the base as a uniform element index and spends ~6 more integer instructions per load: ncu r02, 23 % of the
                // fused kernel's 3.0e8 warp instructions were address arithmetic); extents are checked by the host
                if (l_kc < p.wchunks) {          // 8 channels of the feature half, rebuilt from the feature maps
                    unsigned long long rsv;
                    asm("mov.u64 %0, %1;" : "=l"(rsv) : "l"(p.in + (long long)b * p.isB + (long long)(l_kc * 8) * p.isC));
                    const float* rs = reinterpret_cast<const float*>(rsv);
                    [[maybe_unused]] const float* ls = p.left + (long long)b * p.isB + (long long)(l_kc * 8) * p.isC;
#pragma unroll
                    for (int u = 0; u < RPW; ++u) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            if constexpr (FUSE == 1) {
                                const float* t = rs + ((unsigned)woff[u] + (unsigned)(c * p.isC));
                                const float ra = __ldg(t);
                                const float rb = __ldg(t + 1);
                                v[BUF][u][c] = fmaf(rb, wb[u], __fmul_rn(ra, wa[u]));
                            } else {
                                const float l = off[u] >= 0 ? __ldg(ls + (long long)c * p.isC + off[u]) : 0.f;
                                const float e = l - __fmul_rn(__ldg(rs + ((unsigned)woff[u] + (unsigned)(c * p.isC))), wa[u]);
                                v[BUF][u][c] = -(e * e);
                            }
                        }
                    }
                } else {                         // 8 channels of the group-cost volume
                    const int kc2 = l_kc - p.wchunks;
                    const float* src = p.in2 + (long long)b * p.i2sB + (long long)d * p.i2sD + (long long)(kc2 * 8) * p.i2sC;
                    const int nvalid = p.C2 - kc2 * 8;
                    if (nvalid >= 8) {           // full chunk: unconditional loads, padding positions zeroed when the unit is packed
                        unsigned long long sv64;
                        asm("mov.u64 %0, %1;" : "=l"(sv64) : "l"(src));
                        const float* sv = reinterpret_cast<const float*>(sv64);
                        v_ok[BUF] = 0u;
#pragma unroll
                        for (int u = 0; u < RPW; ++u) {
                            const bool ok = off[u] >= 0;
                            const unsigned o = ok ? (unsigned)off[u] : 0u;
                            v_ok[BUF] |= (ok ? 1u : 0u) << u;
#pragma unroll
                            for (int c = 0; c < 8; ++c) v[BUF][u][c] = __ldg(sv + (o + (unsigned)(c * p.i2sC)));
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < RPW; ++u) {
                            const float* su = src + off[u];
                            const bool ok = off[u] >= 0;
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                v[BUF][u][c] = (ok && c < nvalid) ? __ldg(su) : 0.f;
                                su += p.i2sC;
                            }
                        }
                    }
                }
                ++l_kc;
                return;
            }
            int pr = 0, pc = 0;
            bool col_ok = true;
            const float* src = in_pl + (long long)(l_kc * 8) * p.isC;
            if (p.kd) {             // phase = tap along D: input plane of this output plane
                int din;
                if (p.dtrans) {     // transposed k3 s2 p1 op1: dout = 2*din - 1 + tap
                    const int t2 = d + 1 - l_phase;
                    din = (t2 >= 0 && (t2 & 1) == 0) ? (t2 >> 1) : -1;
                } else {
                    din = d * p.dstride + (l_phase - p.kd / 2) * p.ddil;
                }
                col_ok = din >= 0 && din < p.Din;
                src += (long long)(col_ok ? din : 0) * p.isD;
            } else {                // phase = (row, column) parity of a stride-2 input
                pr = l_phase >> 1;
                pc = l_phase & 1;
                src += pr * p.Win + pc;
                col_ok = !(pc && edge < 0);
            }
            const int nvalid = p.Cin - l_kc * 8;                 // channels of this chunk that exist (>= 8: all)
            if (nvalid >= 8) {
                // Full chunk (warp-uniform branch): unconditional loads at 32-bit offsets from the chunk's base — a padding
                // position reads element 0 of the plane instead — and a per-row validity bit that zeroes the PACKED operand
                // at conversion time.  The predicated zero-fill form below costs ~9 instructions per load (64-bit index
                // increment, LEA pair, ISETP, two R2UR to re-materialise the descriptor under the predicate, a zero MOV:
                // ncu r02 source view, issue slots 57 %); this one ~3.  The zeroing must not touch v here: writing a load's
                // destination register would wait for the load and serialise the prefetch.
                v_ok[BUF] = 0u;
                // the chunk's base as an opaque 64-bit register: one IMAD.WIDE.U32 (base + 4 * offset) per load; left to
                // itself ptxas keeps the base as a uniform element index and spends IADD3 + IMAD.X + LEA + LEA.HI.X on each
                unsigned long long srcv;
                asm("mov.u64 %0, %1;" : "=l"(srcv) : "l"(src));
                const float* sv = reinterpret_cast<const float*>(srcv);
#pragma unroll
                for (int u = 0; u < RPW; ++u) {
                    const bool ok = off[u] >= 0 && col_ok && !(pr && ((edge >> u) & 1));
                    const unsigned o = ok ? (unsigned)off[u] : 0u;
                    v_ok[BUF] |= (ok ? 1u : 0u) << u;
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[BUF][u][c] = __ldg(sv + (o + (unsigned)(c * p.isC)));
                }
            } else {
                v_ok[BUF] = 0xffffffffu;
#pragma unroll
                for (int u = 0; u < RPW; ++u) {
                    const float* su = src + off[u];
                    const bool ok = off[u] >= 0 && col_ok && !(pr && ((edge >> u) & 1));
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        v[BUF][u][c] = (ok && c < nvalid) ? __ldg(su) : 0.f;
                        su += p.isC;
                    }
                }
            }
            if (++l_kc == p.cpp) {
                l_kc = 0;
                ++l_phase;
            }
        };
        // cp.async ring: every thread copies its own (row, lane) x 8 channels of the next unit into the raw stage and
        // posts one asynchronous arrival; a thread only ever reads back what it copied itself, so a stage needs no
        // "empty" barrier.  Uses the same cursor / predicates as load_chunk.
        int c_issue = 0;
        auto issue_cpa = [&]() {
            const int slot = c_issue % p.rs;
            ++c_issue;
            int pr = 0, pc = 0;
            bool col_ok = true;
            const float* src = in_pl + (long long)(l_kc * 8) * p.isC;
            if (p.kd) {
                int din;
                if (p.dtrans) {
                    const int t2 = d + 1 - l_phase;
                    din = (t2 >= 0 && (t2 & 1) == 0) ? (t2 >> 1) : -1;
                } else {
                    din = d * p.dstride + (l_phase - p.kd / 2) * p.ddil;
                }
                col_ok = din >= 0 && din < p.Din;
                src += (long long)(col_ok ? din : 0) * p.isD;
            } else {
                pr = l_phase >> 1;
                pc = l_phase & 1;
                src += pr * p.Win + pc;
                col_ok = !(pc && edge < 0);
            }
            const int nvalid = p.Cin - l_kc * 8;
            float* dst = reinterpret_cast<float*>(raw_base + (size_t)slot * raw_bytes) + lane;
#pragma unroll
            for (int u = 0; u < RPW; ++u) {
                const int r = warp + 8 * u;
                if (r < SR) {
                    const float* su = src + off[u];
                    const bool ok = off[u] >= 0 && col_ok && !(pr && ((edge >> u) & 1));
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const bool okc = ok && c < nvalid;
                        cp_async4(dst + (c * SR + r) * 32, okc ? su : p.in, okc);
                        su += p.isC;
                    }
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&raw_full[slot])) : "memory");
            if (++l_kc == p.cpp) {
                l_kc = 0;
                ++l_phase;
            }
        };
        constexpr int NACC = DIRECT ? 1 : N;
        float acc[JT][NACC];
#pragma unroll
        for (int jj = 0; jj < JT; ++jj)
#pragma unroll
            for (int n = 0; n < NACC; ++n) acc[jj][n] = 0.f;

        auto drain = [&](int g) {
            if constexpr (!DIRECT) {
#pragma unroll
            for (int jj = 0; jj < JT; ++jj) {
                const int j = half + 2 * jj;
                if (j < MT) {
                    mbar_wait(&acc_full[j], (uint32_t)g & 1u);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * 2 * N);
#pragma unroll
                    for (int c0 = 0; c0 < N; c0 += 8) {
                        uint32_t rh[8], rl[8];
                        tmem_ld8(taddr + (uint32_t)c0, rh);
                        tmem_ld8(taddr + (uint32_t)(N + c0), rl);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 8; ++c) acc[jj][c0 + c] += __uint_as_float(rh[c]) + __uint_as_float(rl[c]);
                    }
                    tc_fence_before();
                    mbar_arrive(&acc_empty[j]);
                }
            }
            }
        };

        if constexpr (CPA) {
            for (int i = 0; i < p.rs && i < p.nchunk; ++i) issue_cpa();
        } else if constexpr (!TMA && !SPLIT) {
            load_chunk(std::integral_constant<int, 0>{});
        }
        int s = 0, gk = 0, gdone = 0;      // stage of chunk k; chunks produced since the last group boundary; groups drained
        uint32_t ph = 0;
        auto step = [&](int k, auto BUFC) {
            constexpr int BUF = decltype(BUFC)::value;
            if constexpr (NB == 2) {        // unit k+1 goes in flight now, into the other buffer (whose unit k-1 is already packed)
                if (k + 1 < p.nchunk) load_chunk(std::integral_constant<int, 1 - BUF>{});
            }
            // F16: two 8-channel units (K halves) make one MMA chunk / shared-memory stage
            const bool first_half = !F16 || !(k & 1);
            const bool last_half = !F16 || (k & 1) || k == p.nchunk - 1;
            if (first_half) mbar_wait(&empty[s], ph ^ 1u);
            uint8_t* st_base = smem + (size_t)s * stage_bytes;
            if (first_half && tid == 0) {
                mbar_arrive_expect_tx(&full[s], b_bytes);
                bulk_g2s(st_base + a_bytes, p.wpack + (size_t)(F16 ? k >> 1 : k) * (b_bytes / 4), b_bytes, &full[s]);
            }
            const uint32_t a_hi = smem_u32(st_base);
            const uint32_t khalf = NPOS * 16u;
            const uint32_t a_lo = a_hi + 2u * khalf;
            // the unit's values: registers (v) or the raw fp32 stage (TMA / cp.async ring)
            const float* raw = nullptr;
            int rbw = 32;
            if constexpr (TMA || CPA) {
                const int rsl = k % p.rs;
                mbar_wait(&raw_full[rsl], ((uint32_t)(k / p.rs)) & 1u);
                raw = reinterpret_cast<const float*>(raw_base + (size_t)rsl * raw_bytes) + lane;
                if constexpr (TMA) {
                    raw += (x0 - p.dil) & 3;
                    rbw = p.bw;
                }
            }
            auto val = [&](int u, int r, int c) -> float {
                if constexpr (TMA || CPA) return raw[(c * SR + r) * rbw];
                else return v[BUF][u][c];
            };
            if constexpr (F16) {
                const uint32_t ko = (uint32_t)(k & 1) * khalf;         // K half of this unit inside the chunk
#pragma unroll
                for (int u = 0; u < RPW; ++u) {
                    const int r = warp + 8 * u;
                    if (r < SR) {
                        const bool live = (TMA || CPA) ? true : ((v_ok[BUF] >> u) & 1u) != 0u;
                        uint32_t hi[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) hi[c] = pack_h2(val(u, r, 2 * c), val(u, r, 2 * c + 1));
                        const uint32_t o = (uint32_t)(r * 32 + lane) * 16u;
                        if (live) sts128(a_hi + ko + o, hi[0], hi[1], hi[2], hi[3]);
                        else sts128(a_hi + ko + o, 0u, 0u, 0u, 0u);
                        if (p.terms == 3) {                             // uniform: the single-term form needs no lo half
                            uint32_t lo[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float2 hf = unpack_h2(hi[c]);
                                lo[c] = pack_h2(val(u, r, 2 * c) - hf.x, val(u, r, 2 * c + 1) - hf.y);
                            }
                            if (live) sts128(a_lo + ko + o, lo[0], lo[1], lo[2], lo[3]);
                            else sts128(a_lo + ko + o, 0u, 0u, 0u, 0u);
                        }
                        if (!(k & 1) && k == p.nchunk - 1) {            // odd number of units: the last K half is zero
                            sts128(a_hi + khalf + o, 0u, 0u, 0u, 0u);
                            sts128(a_lo + khalf + o, 0u, 0u, 0u, 0u);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < RPW; ++u) {
                    const int r = warp + 8 * u;
                    if (r < SR) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float x = val(u, r, c);
                            hi[c] = tf32_rna(x);
                            // exact remainder; the tensor core ignores the 13 low mantissa bits of a tf32 operand,
                            // so not rounding lo costs <= 2^-22 relative
                            lo[c] = __float_as_uint(x - __uint_as_float(hi[c]));
                        }
                        if (!(TMA || CPA) && !((v_ok[BUF] >> u) & 1u)) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) hi[c] = lo[c] = 0u;
                        }
                        const uint32_t o = (uint32_t)(r * 32 + lane) * 16u;
                        sts128(a_hi + o, hi[0], hi[1], hi[2], hi[3]);
                        sts128(a_hi + khalf + o, hi[4], hi[5], hi[6], hi[7]);
                        sts128(a_lo + o, lo[0], lo[1], lo[2], lo[3]);
                        sts128(a_lo + khalf + o, lo[4], lo[5], lo[6], lo[7]);
                    }
                }
            }
            if constexpr (TMA) {
                __syncwarp();
                if (elect_one()) mbar_arrive(&raw_empty[k % p.rs]);     // this warp has read the raw stage
                __syncwarp();
            } else if constexpr (CPA) {
                if (c_issue < p.nchunk) issue_cpa();    // refill the stage this thread just read
            } else if constexpr (NB == 1) {
                if (k + 1 < p.nchunk) load_chunk(std::integral_constant<int, 0>{});   // in flight across the barrier traffic and the drain below
            }
            if (last_half) {
                fence_proxy_async();                   // generic-proxy st.shared -> visible to the tensor core
                mbar_arrive(&full[s]);
                if (++s == p.stages) {
                    s = 0;
                    ph ^= 1u;
                }
                if constexpr (!DIRECT) {
                    // a group that finished one chunk ago is drained now: overlaps this chunk's MMAs
                    if (gk == p.G) {
                        drain(gdone++);
                        gk = 0;
                    }
                    ++gk;
                }
            }
        };
        uint32_t it = 0;                 // tiles this CTA has finished: parity of the per-tile accumulator barriers
        for (int t = t_first; t < T; t += t_stride, ++it) {
        if constexpr (SPLIT) {
            // nothing to produce: the TMA warp fills the operand stages; these warps only drain the accumulator groups
            set_tile(t);
            if constexpr (!DIRECT)
                for (int g = 0; g + 1 < ngroups; ++g) drain(g);
        } else if constexpr (NB == 2) {
            for (int k = 0; k < p.nchunk; k += 2) {
                step(k, std::integral_constant<int, 0>{});
                if (k + 1 < p.nchunk) step(k + 1, std::integral_constant<int, 1>{});
            }
        } else {
            for (int k = 0; k < p.nchunk; ++k) step(k, std::integral_constant<int, 0>{});
        }
        if constexpr (!DIRECT) drain(ngroups - 1);

        // ===================== epilogue: kx shift-sum, bias, activation, NCDHW stores =====================
        float* out_pl = p.out ? p.out + (long long)b * p.osB + (long long)d * p.osD : nullptr;
        unsigned short* outs_pl = p.outs ? p.outs + (long long)b * p.ossB + (long long)d * p.ossD : nullptr;
        const bool s_batch = b < p.s_nb;
        const int x = x0 + lane;
        const bool xok = lane < VW && x < p.W;
        const int xo = x * p.osX;
        const int ey0 = y0;
        if constexpr (MULTI && !SPLIT) {
            // the next tile's first unit goes in flight before this tile's epilogue (buffer 0 is free: every unit of this
            // tile has been packed); the epilogue below works from the coordinates saved above
            if (t_stride < T - t) {
                set_tile(t + t_stride);
                setup_tile();
                load_chunk(std::integral_constant<int, 0>{});
            }
        }
#pragma unroll
        for (int jj = 0; jj < JT; ++jj) {
            const int j = half + 2 * jj;
            if (j < MT) {                                  // warp-uniform
                const int y = ey0 + 4 * j + quarter;
                const bool ok = xok && y < p.H;
                const long long pos = (long long)y * p.osY + xo;
                EpiOut eo;
                eo.o = out_pl ? out_pl + pos : nullptr;
                eo.osC = p.osC;
                eo.so = outs_pl ? outs_pl + pos * 8 : nullptr;
                eo.so_lo = p.s_parts == 2 ? p.ossP : 0ll;
                eo.so_c8 = p.ossC8;
                eo.sok = ok && s_batch;
                if constexpr (DIRECT) {
                    // one accumulation group: stream TMEM -> registers 8 channels at a time
                    mbar_wait(&acc_full[j], it & 1u);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * C::TS);
                    if (p.act == TSTEREO_ACT_SILU) epi_direct<TSTEREO_ACT_SILU, CP, FOLD>(taddr, p.dil, eo, s_bias, p.Cout, ok, s_osc);
                    else if (p.act == TSTEREO_ACT_RELU) epi_direct<TSTEREO_ACT_RELU, CP, FOLD>(taddr, p.dil, eo, s_bias, p.Cout, ok, s_osc);
                    else epi_direct<TSTEREO_ACT_NONE, CP, FOLD>(taddr, p.dil, eo, s_bias, p.Cout, ok, s_osc);
                    if constexpr (MULTI) {                 // M-tile j of the accumulator is free for the next tile's MMAs
                        tc_fence_before();
                        mbar_arrive(&acc_empty[j]);
                    }
                } else {
                    const float* ad = nullptr;
                    if constexpr (FUSE == 1)
                        if (p.add) ad = p.add + (long long)b * p.asB + (long long)min(y, p.H - 1) * p.W + min(x, p.W - 1);
                    if (p.act == TSTEREO_ACT_SILU) epi_acc<TSTEREO_ACT_SILU, CP, FOLD>(acc[jj], p.dil, eo, s_bias, p.Cout, ok, s_osc, ad, p.asC);
                    else if (p.act == TSTEREO_ACT_RELU) epi_acc<TSTEREO_ACT_RELU, CP, FOLD>(acc[jj], p.dil, eo, s_bias, p.Cout, ok, s_osc, ad, p.asC);
                    else epi_acc<TSTEREO_ACT_NONE, CP, FOLD>(acc[jj], p.dil, eo, s_bias, p.Cout, ok, s_osc, ad, p.asC);
                }
            }
        }
        }   // tiles
        if constexpr (DIRECT) tc_fence_before();
    } else if (warp == NPROD / 32) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_2n = F16 ? idesc_f16(2 * N) : idesc_tf32(2 * N), idesc_n = F16 ? idesc_f16(C::N2) : idesc_tf32(C::N2);
        auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accf) {
            if constexpr (F16) tc_mma_f16(d, a, b, idesc, accf);
            else tc_mma_tf32(d, a, b, idesc, accf);
        };
        const uint32_t a_lbo = NPOS * 16u, b_lbo = 2u * N * 16u;
        int s = 0, g = 0, kg = 0;          // stage; accumulation group; chunk index inside the group
        uint32_t ph = 0;
        if constexpr (TMA)
            for (int i = 0; i < p.rs && i < p.nchunk; ++i) issue_tma();
        if (wres) mbar_wait(wfull, 0u);
        uint32_t it = 0;                 // tiles finished (MULTI): parity of acc_empty
        for (int t = t_first; t < T; t += t_stride, ++it)
        for (int k = 0; k < nmma; ++k) {
            const bool first = DIRECT ? k == 0 : kg == 0;
            const bool last = DIRECT ? k == nmma - 1 : (kg == p.G - 1 || k == nmma - 1);
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t st_base = smem_u32(smem + (size_t)s * stage_bytes);
            // descriptor start-address field is (addr >> 4): offsets add directly (no carry out of its 14 bits)
            const uint64_t a_hi = make_desc(st_base, a_lbo, 128u);
            const uint64_t a_lo = make_desc(st_base + 2u * NPOS * 16u, a_lbo, 128u);
            const uint64_t b_d = make_desc(wres ? smem_u32(wbase) + (uint32_t)k * b_bytes : st_base + a_bytes, b_lbo, 128u);
#pragma unroll 1
            for (int j = 0; j < MT; ++j) {
                if (!DIRECT && first && g >= 1) {
                    mbar_wait(&acc_empty[j], (uint32_t)(g - 1) & 1u);
                    tc_fence_after();
                }
                if (MULTI && first && it >= 1) {          // the previous tile's epilogue has read M-tile j out of TMEM
                    mbar_wait(&acc_empty[j], (it - 1u) & 1u);
                    tc_fence_after();
                }
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(j * C::TS);
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        if (ky >= p.nky) break;
                        const uint64_t ad = (uint64_t)((4 * j + ky * p.dil) * 32);
                        const uint64_t bd = b_d + (uint64_t)(ky * (2 * 2 * N));
                        const uint32_t acc0 = (first && ky == 0) ? 0u : 1u;
                        if constexpr (DIRECT) {
                            // all three terms into the same N columns (B rows [0,N) = hi, [N,2N) = lo)
                            mma(d_tmem, a_hi + ad, bd, idesc_n, acc0);
                            if (p.terms == 3) {
                                mma(d_tmem, a_hi + ad, bd + (uint64_t)N, idesc_n, 1u);
                                mma(d_tmem, a_lo + ad, bd, idesc_n, 1u);
                            }
                        } else {
                            // [A*B_hi | A_hi*B_lo] (+)= A_hi * [B_hi | B_lo]
                            mma(d_tmem, a_hi + ad, bd, idesc_2n, acc0);
                            // first block += A_lo * B_hi
                            mma(d_tmem, a_lo + ad, bd, idesc_n, 1u);
                        }
                    }
                    if (last) tc_commit(&acc_full[j]);
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit(&empty[s]);      // stage reusable once every MMA above has read it
            __syncwarp();
            if constexpr (TMA)
                if (t_issue < p.nchunk) issue_tma();   // raw stage of chunk k: the producers released it before full[s]
            if (++s == p.stages) {
                s = 0;
                ph ^= 1u;
            }
            if (++kg == p.G) {
                kg = 0;
                ++g;
            }
        }
        tc_fence_before();
    } else if constexpr (SPLIT) {
        // ===================== TMA issuer (SPLIT): one box per 8-channel unit and part + the weight image per stage =====================
        const uint32_t unit_bytes = NPOS * 16u;
        const int nparts = p.terms == 3 ? 2 : 1;
        int s = 0;
        uint32_t ph = 0;
        if (wres && t_first < T) {
            if (elect_one()) {
                mbar_arrive_expect_tx(wfull, (uint32_t)nmma * b_bytes);
                for (int k = 0; k < nmma; ++k) bulk_g2s(wbase + (size_t)k * b_bytes, p.wpack + (size_t)k * (b_bytes / 4), b_bytes, wfull);
            }
            __syncwarp();
        }
        for (int t = t_first; t < T; t += t_stride) {
            set_tile(t);
            for (int k = 0; k < nmma; ++k) {
                mbar_wait(&empty[s], ph ^ 1u);
                if (elect_one()) {
                    uint8_t* st_base = smem + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&full[s], 2u * (uint32_t)nparts * unit_bytes + (wres ? 0u : b_bytes));
                    if (!wres) bulk_g2s(st_base + a_bytes, p.wpack + (size_t)k * (b_bytes / 4), b_bytes, &full[s]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int u = 2 * k + h;            // 8-channel unit = K half h of MMA chunk k
                        int phase = u / p.cpp, kc = u - phase * p.cpp;
                        if (u >= p.nchunk) {                // odd number of units: the last K half is a box of zeros
                            phase = 0;
                            kc = p.cpp;
                        }
                        int c0, c1, c2, c3, c4;
                        if (p.s_sx == 2) {                  // map (ch 8, x, y, c8, plane), element strides (1, 2, 2, 1, 1)
                            c0 = 0;
                            c1 = 2 * (x0 - 1) + (phase & 1);
                            c2 = 2 * (y0 - 1) + (phase >> 1);
                            c3 = kc;
                            c4 = b * p.D + d;
                        } else {                            // map (x*8 + ch, y, c8, plane d, b)
                            int dd = d;
                            if (p.kd) {
                                if (p.dtrans) {             // transposed k3 s2 p1 op1: dout = 2*din - 1 + tap
                                    const int t2 = d + 1 - phase;
                                    dd = (t2 >= 0 && (t2 & 1) == 0) ? (t2 >> 1) : -1;
                                } else {
                                    dd = d * p.dstride + (phase - p.kd / 2) * p.ddil;
                                }
                                if (dd < 0 || dd >= p.Din) dd = -1;      // outside: the TMA unit fills zeros
                            }
                            c0 = (x0 - p.dil) * 8;
                            c1 = y0 - p.dil;
                            c2 = kc;
                            c3 = dd;
                            c4 = b;
                        }
                        tma_load_5d(st_base + (size_t)h * unit_bytes, &tmap, &full[s], c0, c1, c2, c3, c4);
                        if (nparts == 2) tma_load_5d(st_base + (size_t)(2 + h) * unit_bytes, &tmap_lo, &full[s], c0, c1, c2, c3, c4);
                    }
                }
                __syncwarp();
                if (++s == p.stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    }
    __syncthreads();
    if (warp == NPROD / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::NCOLS) : "memory");
    }
}

static size_t smem_need(int stages, int SR, int N, int rs = 0, int bw = 36, int a_units = 4, size_t wres_bytes = 0) {
    return (size_t)stages * ((size_t)SR * 512 * a_units + (wres_bytes ? 0 : (size_t)192 * N)) + wres_bytes + BAR_BYTES +
           (size_t)rs * 32 * bw * SR;
}

static int env_int(const char* name, int dflt);

// cuTensorMapEncodeTiled through the runtime's driver entry point (libtstereo does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* q = nullptr;
        cudaDriverEntryPointQueryResult res;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &res) == cudaSuccess &&
            res == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)q;
        cudaGetLastError();
    }
    return fn;
}

// (x, y, plane, c, b) view of the NCDHW input with an (bw, SR, 1, 8, 1) box; false when the view is not expressible
// (rows not 16-byte aligned, strided x, ...) -> the register path runs instead
static bool make_tmap(const Params& p, int SR, CUtensorMap* tm) {
    // opt-in (TSTEREO_TC2_TMA=1): measured 8-10 % slower than the register path on B200 — the kernel is bound by the
    // shared-memory pipe (tensor-core operand reads 41-55 % + producer stores), and the raw fp32 stage adds one more
    // shared-memory write + read per element
    if (!env_int("TSTEREO_TC2_TMA", 0) || p.isX != 1 || p.isY != p.W || p.Hin != p.H || p.Win != p.W) return false;
    if ((p.W & 3) || (p.isC & 3) || (p.isD & 3) || (p.isB & 3) || (((size_t)p.in) & 15)) return false;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const int planes = p.kd ? p.Din : p.D;
    const long long nb = p.nbatch;
    // dimensions in order of increasing stride (NCDHW: x, y, plane, channel, batch)
    const cuuint64_t dims[5] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)planes, (cuuint64_t)p.Cin, (cuuint64_t)nb};
    // a dimension of extent 1 may carry any stride: use a legal (monotone, multiple of 16 B) placeholder
    const cuuint64_t sY = (cuuint64_t)p.W * 4;
    const cuuint64_t sD = planes > 1 ? (cuuint64_t)p.isD * 4 : sY * (cuuint64_t)p.H;
    const cuuint64_t sC = (cuuint64_t)p.isC * 4;
    const cuuint64_t sB = nb > 1 ? (cuuint64_t)p.isB * 4 : sC * (cuuint64_t)p.Cin;
    if (sD < sY * (cuuint64_t)p.H || sC < sD * (cuuint64_t)planes || sB < sC * (cuuint64_t)p.Cin) return false;
    const cuuint64_t strides[4] = {sY, sD, sC, sB};
    const cuuint32_t box[5] = {(cuuint32_t)p.bw, (cuuint32_t)SR, 1, 8, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(p.in), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// S-format input [B][D][part][C8][H][W][8] fp16 (strides in elements) as a TMA view; one map per part (hi, lo).
//   stride 1: (x*8 + ch, y, c8, d, b), box (256, SR, 1, 1, 1): a box row is 32 positions x 8 channels = the K-major stage row
//   stride 2: (ch, x, y, c8, plane b*D + d), box (8, 64, 2*SR, 1, 1) walked with element strides (1, 2, 2, 1, 1)
struct SIn {
    const unsigned short* ptr;
    long long sB, sD, sP, sC8;
    int C8, parts;
};
static bool make_tmap_s(const SIn& si, int part, int B, int D, int H, int W, int SR, int sx, CUtensorMap* tm) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const unsigned short* base = si.ptr + (part ? si.sP : 0ll);
    if ((((size_t)base) & 15) || (si.sB & 7) || (si.sD & 7) || (si.sC8 & 7)) return false;
    const cuuint64_t sY = (cuuint64_t)W * 16;
    const cuuint64_t sC = (cuuint64_t)si.sC8 * 2;
    // a dimension of extent 1 may carry any stride: use a legal placeholder
    const cuuint64_t sDb = D > 1 ? (cuuint64_t)si.sD * 2 : sC * (cuuint64_t)si.C8;
    const cuuint64_t sBb = B > 1 ? (cuuint64_t)si.sB * 2 : sDb * (cuuint64_t)D;
    if (sx == 1) {
        const cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)si.C8, (cuuint64_t)D, (cuuint64_t)B};
        const cuuint64_t strides[4] = {sY, sC, sDb, sBb};
        const cuuint32_t box[5] = {256, (cuuint32_t)SR, 1, 1, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<unsigned short*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    // stride 2: batch and plane merge into one dimension, which needs sB == D * sD
    if (B > 1 && D > 1 && si.sB != (long long)D * si.sD) return false;
    const cuuint64_t sPl = (B * D > 1) ? (D > 1 ? (cuuint64_t)si.sD * 2 : (cuuint64_t)si.sB * 2) : sC * (cuuint64_t)si.C8;
    const cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)si.C8, (cuuint64_t)B * D};
    const cuuint64_t strides[4] = {16, sY, sC, sPl};
    const cuuint32_t box[5] = {8, 64, (cuuint32_t)(2 * SR), 1, 1};
    const cuuint32_t estr[5] = {1, 2, 2, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<unsigned short*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int CP, int MT, bool DIRECT, int RAW, bool F16, int FOLD = 3, int FUSE = 0>
static int launch_one(const Params& p, const CUtensorMap& tm, const CUtensorMap& tm2, dim3 grid, size_t smem_bytes, cudaStream_t st,
                      const char* what) {
    auto kern = conv_tc2_kernel<CP, MT, DIRECT, RAW, F16, FOLD, FUSE>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
        if (e != cudaSuccess) {
            set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return TSTEREO_E_CUDA;
        }
        attr_done = true;
    }
    launch_k(kern, grid, dim3(nthreads(RAW)), smem_bytes, st, p, tm, tm2);
    return check_launch(what);
}

// bias == null -> a device buffer of zeros, so the epilogue loads it unconditionally
__device__ float g_zero_bias[64];
static const float* zero_bias() {
    static const float* ptr = nullptr;
    if (!ptr) {
        void* q = nullptr;
        if (cudaGetSymbolAddress(&q, g_zero_bias) == cudaSuccess) ptr = (const float*)q;
    }
    return ptr;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

template <int FUSE = 0>
static int launch(Params& p, int CP, int planes, cudaStream_t st, const char* what, const SIn* sin = nullptr) {
    const int fold = p.fold == 1 ? 1 : 3;
    const int N = fold * CP, N2 = (N + 15) / 16 * 16;
    const int VW = 32 - 2 * p.dil;
    p.tiles_x = (p.W + VW - 1) / VW;
    const int g_env = env_int("TSTEREO_TC2_G", 0);      // experiments: chunks per accumulation group (per input phase)
    if (g_env > 0) p.G = g_env * (p.nchunk / p.cpp);
    if (p.G < 1) p.G = 1;
    const int nmma = p.half ? (p.nchunk + 1) / 2 : p.nchunk;
    if (p.half) p.G = (p.G + 1) / 2;                     // same products per accumulation group: 16 channels per chunk
    if (p.terms != 1) p.terms = 3;
    if (p.terms == 1) {
        TS_REQUIRE(p.half && FUSE == 0, "%s: the single-term form is an fp16 variant of the plain convolutions", what);
        p.G = nmma;                                      // 11-bit operands: nothing to gain from short TMEM accumulation groups
    }
    const bool direct = FUSE == 0 && nmma <= p.G && !env_int("TSTEREO_TC2_NODIRECT", 0);
    const bool split = FUSE == 0 && sin != nullptr;
    if (split) {
        TS_REQUIRE(p.half, "%s: the S-format input is the fp16 split", what);
        TS_REQUIRE(sin->parts == 2 || p.terms == 1, "%s: a hi-only S-format input serves the single-term form only", what);
        TS_REQUIRE(sin->C8 >= p.cpp, "%s: S-format input has %d chunks, the layer needs %d", what, sin->C8, p.cpp);
    }
    p.s_in = split;
    const int a_units = (split && p.terms != 3) ? 2 : 4;
    // M-tiles per CTA: 2 or 4 (TMEM: MT * columns-per-tile <= 512); cost = SM-time of all waves
    p.nbatch = planes / p.D;
    p.bw = p.dil ? 36 : 32;
    CUtensorMap tm = {};
    CUtensorMap tm2 = {};
    const bool tma_ok = FUSE == 0 && !split && !p.half && make_tmap(p, 4 * 2 + 2 * p.dil, &tm);   // eligibility (the box is re-encoded for the chosen tile)
    // cp.async ring (fp16 split only; opt-in TSTEREO_TC2_CPA=1): loads run `rs` chunks ahead without holding registers.
    // Measured on B200: no gain on the small hourglass layers (they sit at the fixed launch + prologue + epilogue cost,
    // ~10 us) and 5-20 % slower on the large ones (one more shared-memory round trip), so it is not the default.
    const int cpa_env = (FUSE == 0 && !split) ? env_int("TSTEREO_TC2_CPA", 0) : 0;
    int best_mt = 0, best_stages = 0, best_rs = 0, best_minb = 2;
    bool best_cpa = false, best_wres = false;
    // SPLIT + DIRECT (resident CTAs walking many tiles): keep the whole weight image in shared memory when that leaves
    // room for >= 3 operand stages — the per-tile weight reload is otherwise 30-65 % of the L2 -> SM traffic
    const size_t wres_bytes = (split && direct && env_int("TSTEREO_TC2_WRES", 1)) ? (size_t)nmma * p.nky * 64 * N : 0;
    double best_cost = 1e30;
    const int forced = env_int("TSTEREO_TC2_MT", 0);
    for (int mt = 2; mt <= 4; mt += 2) {
        const int cols = mt * (direct ? N2 : 2 * N);
        // no (32, 4) instance of the kx-folded form: one CTA per SM (384 TMEM columns) measured slower than two CTAs of (32, 2)
        // in every producer mode (SPLIT: 141 vs 129 us on the UNet 32 -> 32 conv) — the epilogue needs the second CTA's warps
        if (cols > 512 || (CP == 32 && mt == 4 && fold == 3)) continue;
        if (forced && mt != forced) continue;
        // fused cost producer: the (16, 4) instance needs 168 registers (one CTA per SM) and loses to two CTAs of (16, 2)
        // on the gather latency (fine level, B200: 295 vs 200 us)
        if (FUSE != 0 && !forced && CP == 16 && mt == 4) continue;
        const int jt = (mt + 1) / 2;
        int minb = (direct || jt * N <= 48) ? 2 : 1;
        if (FUSE != 0 && FUSE_ONE_CTA) minb = 1;
        if (cols > 256) minb = 1;                       // two CTAs need their TMEM columns side by side
        const int SR = 4 * mt + 2 * p.dil;
        const long long tiles = (long long)p.tiles_x * ((p.H + 4 * mt - 1) / (4 * mt)) * planes;
        const bool cpa = p.half && cpa_env == 1 && fold == 3;
        const size_t budget = (minb == 2 && !(cpa && tiles <= 148)) ? (size_t)112 * 1024 : SMEM_MAX;
        int stages = 0, rs = 0;
        if (cpa) {          // two operand stages + as many raw fp32 stages as fit
            for (int r = MAX_RS; r >= 2 && !stages; --r)
                if (smem_need(2, SR, N, r, 32) <= budget) {
                    stages = 2;
                    rs = r;
                }
        } else if (tma_ok) {       // operand stages + raw fp32 stages (1/2 the size): prefer depth on the raw side
            static const int combos[][2] = {{3, 4}, {3, 3}, {2, 4}, {2, 3}, {3, 2}, {2, 2}};
            for (const auto& c : combos)
                if (smem_need(c[0], SR, N, c[1], p.bw) <= budget) {
                    stages = c[0];
                    rs = c[1];
                    break;
                }
        }
        bool wres = false;
        if (!stages && wres_bytes) {
            for (int s = MAX_STAGES; s >= 3; --s)
                if (smem_need(s, SR, N, 0, 36, a_units, wres_bytes) <= budget) {
                    stages = s;
                    wres = true;
                    break;
                }
        }
        if (!stages) {
            rs = 0;
            for (int s = split ? MAX_STAGES : 4; s >= 2; --s)
                if (smem_need(s, SR, N, 0, 36, a_units) <= budget) {
                    stages = s;
                    break;
                }
        }
        if (!stages) continue;
        const long long waves = (tiles + 148 * minb - 1) / (148 * minb);
        const double cost = (double)waves * minb * (SR + 5.0) * ((stages >= 3 || rs) ? 1.0 : 1.15);
        if (cost < best_cost) {
            best_cost = cost;
            best_mt = mt;
            best_stages = stages;
            best_rs = rs;
            best_cpa = cpa && rs > 0;
            best_wres = wres;
            best_minb = minb;
        }
    }
    TS_REQUIRE(best_mt > 0, "%s: no tile configuration for Cout<=%d", what, CP);
    if (!p.bias) p.bias = zero_bias();
    TS_REQUIRE(p.bias, "%s: zero-bias buffer unavailable", what);
    p.stages = best_stages;
    p.rs = best_rs;
    p.wres = best_wres;
    const int SR = 4 * best_mt + 2 * p.dil;
    const bool tma = best_rs > 0 && !best_cpa;
    if (best_cpa) p.bw = 32;
    if (tma && !make_tmap(p, SR, &tm)) {
        set_error("%s: cuTensorMapEncodeTiled failed", what);
        return TSTEREO_E_CUDA;
    }
    if (split) {
        const int Dm = p.kd ? p.Din : p.D;
        TS_REQUIRE(make_tmap_s(*sin, 0, p.nbatch, Dm, p.Hin, p.Win, SR, p.s_sx, &tm) &&
                       (p.terms != 3 || make_tmap_s(*sin, 1, p.nbatch, Dm, p.Hin, p.Win, SR, p.s_sx, &tm2)),
                   "%s: the S-format input is not expressible as a TMA view (alignment / strides)", what);
    }
    const size_t smem_bytes = smem_need(p.stages, SR, N, p.rs, p.bw, a_units, best_wres ? wres_bytes : 0);   // p.bw = 32 for the cp.async ring
    p.tiles_y = (p.H + 4 * best_mt - 1) / (4 * best_mt);
    p.nplanes = planes;
    dim3 grid(p.tiles_x * p.tiles_y, planes);
    // MULTI instances (DIRECT, register producer): resident CTAs walk the tiles — 2 per SM (TSTEREO_TC2_PERSIST=0: one tile
    // per CTA, the same kernel with grid.x = tiles)
    const bool multi = FUSE == 0 && direct && !tma && !best_cpa;   // register producer or SPLIT
    if (multi) {
        const long long T = (long long)grid.x * planes;
        const long long want = (long long)148 * best_minb * (env_int("TSTEREO_TC2_PERSIST", 1) > 0 ? env_int("TSTEREO_TC2_PERSIST", 1) : 1);
        p.persist = env_int("TSTEREO_TC2_PERSIST", 1) != 0;
        TS_REQUIRE(T < (1ll << 31), "%s: too many tiles", what);
        grid = dim3((unsigned)((p.persist && T > want) ? want : T), 1);
    }
    if constexpr (FUSE != 0) {
        TS_REQUIRE(fold == 3 && !direct && !tma && !best_cpa, "%s: the fused cost producer runs the 3x3 ACC register form", what);
#define TS_TC2U(CC, MM)                                                                                        \
    if (CP == CC && best_mt == MM)                                                                             \
        return p.half ? launch_one<CC, MM, false, 0, true, 3, FUSE>(p, tm, tm2, grid, smem_bytes, st, what)    \
                      : launch_one<CC, MM, false, 0, false, 3, FUSE>(p, tm, tm2, grid, smem_bytes, st, what);
        TS_TC2U(8, 2) TS_TC2U(8, 4) TS_TC2U(16, 2) TS_TC2U(16, 4) TS_TC2U(32, 2)
#undef TS_TC2U
        TS_REQUIRE(false, "%s: no fused kernel instance for CP=%d MT=%d", what, CP, best_mt);
    } else {
#define TS_TC2S(CC, MM, FF)                                                                                    \
    if (split && fold == FF && CP == CC && best_mt == MM)                                                      \
        return direct ? launch_one<CC, MM, true, 3, true, FF>(p, tm, tm2, grid, smem_bytes, st, what)          \
                      : launch_one<CC, MM, false, 3, true, FF>(p, tm, tm2, grid, smem_bytes, st, what);
    TS_TC2S(8, 2, 3) TS_TC2S(8, 4, 3) TS_TC2S(16, 2, 3) TS_TC2S(16, 4, 3) TS_TC2S(32, 2, 3)
    TS_TC2S(8, 2, 1) TS_TC2S(8, 4, 1) TS_TC2S(16, 2, 1) TS_TC2S(16, 4, 1) TS_TC2S(32, 2, 1) TS_TC2S(32, 4, 1)
#undef TS_TC2S
    TS_REQUIRE(!split, "%s: no S-format kernel instance for CP=%d MT=%d fold=%d", what, CP, best_mt, fold);
#define TS_TC2F1(CC, MM)                                                                                       \
    if (fold == 1 && p.half && !best_cpa && CP == CC && best_mt == MM)                                         \
        return direct ? launch_one<CC, MM, true, 0, true, 1>(p, tm, tm2, grid, smem_bytes, st, what)           \
                      : launch_one<CC, MM, false, 0, true, 1>(p, tm, tm2, grid, smem_bytes, st, what);
    TS_TC2F1(8, 2) TS_TC2F1(8, 4) TS_TC2F1(16, 2) TS_TC2F1(16, 4) TS_TC2F1(32, 2) TS_TC2F1(32, 4)
#undef TS_TC2F1
    TS_REQUIRE(fold == 3, "%s: the single-column form needs the fp16 split and the register producer", what);
#define TS_TC2(CC, MM)                                                                                         \
    if (CP == CC && best_mt == MM) {                                                                           \
        if (p.half && best_cpa)                                                                                \
            return direct ? launch_one<CC, MM, true, 2, true>(p, tm, tm2, grid, smem_bytes, st, what)          \
                          : launch_one<CC, MM, false, 2, true>(p, tm, tm2, grid, smem_bytes, st, what);        \
        if (p.half)                                                                                            \
            return direct ? launch_one<CC, MM, true, 0, true>(p, tm, tm2, grid, smem_bytes, st, what)          \
                          : launch_one<CC, MM, false, 0, true>(p, tm, tm2, grid, smem_bytes, st, what);        \
        if (direct) return tma ? launch_one<CC, MM, true, 1, false>(p, tm, tm2, grid, smem_bytes, st, what)    \
                               : launch_one<CC, MM, true, 0, false>(p, tm, tm2, grid, smem_bytes, st, what);   \
        return tma ? launch_one<CC, MM, false, 1, false>(p, tm, tm2, grid, smem_bytes, st, what)               \
                   : launch_one<CC, MM, false, 0, false>(p, tm, tm2, grid, smem_bytes, st, what);              \
    }
    TS_TC2(8, 2) TS_TC2(8, 4) TS_TC2(16, 2) TS_TC2(16, 4) TS_TC2(32, 2)
#undef TS_TC2
    TS_REQUIRE(false, "%s: no kernel instance for CP=%d MT=%d", what, CP, best_mt);
    }
}


// ---- host helpers shared by the ABI translation units (conv_tc2.cu, cost_conv.cu)
inline int tc2_cp(int Cout) { return Cout <= 8 ? 8 : Cout <= 16 ? 16 : 32; }

// floats of the operand image of ONE output-channel group (<= 32 channels) over `nchunk` 8-channel chunks
inline long long group_floats(int nchunk, int cout_g, int nky = 3, int fold = 3) {
    return (long long)nchunk * nky * 2 * 2 * (fold * tc2_cp(cout_g)) * 4;
}

// total over the groups of 32 output channels
inline long long wpack_floats(int nchunk, int Cout, int nky = 3, int fold = 3) {
    long long n = 0;
    for (int c0 = 0; c0 < Cout; c0 += 32) n += group_floats(nchunk, Cout - c0 < 32 ? Cout - c0 : 32, nky, fold);
    return n;
}

// One (virtual) stride-1 3x3 convolution, output channels in groups of <= 32 (the producer work is repeated per
// group; Cout = 64 layers are rare and their inputs are L2 resident between the two launches).
template <int FUSE = 0>
inline int run_groups(tc2::Params p, int Cout, int planes, cudaStream_t st, const char* what, const SIn* sin = nullptr) {
    const float* wp = p.wpack;
    const float* bias = p.bias;
    const float* oscale = p.oscale;
    float* out = p.out;
    unsigned short* outs = p.outs;
    for (int c0 = 0; c0 < Cout; c0 += 32) {
        const int cg = Cout - c0 < 32 ? Cout - c0 : 32;
        // a fresh copy per group: launch() rewrites fields (G halves for the fp16 split, tile / stage choices) — reusing
        // one Params made the second group of a 64-channel layer fall out of the DIRECT form (G 8 -> 4 -> 2)
        tc2::Params q = p;
        q.Cout = cg;
        q.wpack = wp;
        q.bias = bias ? bias + c0 : nullptr;
        q.oscale = oscale ? oscale + c0 : nullptr;
        q.out = out ? out + (long long)c0 * p.osC : nullptr;
        q.outs = outs ? outs + (long long)(c0 / 8) * p.ossC8 : nullptr;
        if (p.add) q.add = p.add + (long long)c0 * p.asC;
        const int rc = tc2::launch<FUSE>(q, tc2_cp(cg), planes, st, what, sin);
        if (rc != TSTEREO_OK) return rc;
        wp += group_floats(p.half ? (p.nchunk + 1) / 2 : p.nchunk, cg, p.nky, p.fold == 1 ? 1 : 3);
    }
    return TSTEREO_OK;
}


}  // namespace tc2
}  // namespace tstereo
