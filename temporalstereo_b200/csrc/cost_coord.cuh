// Sampling coordinates of the cost volume's right-feature warp, shared by the materialising kernel
// (block_cost.cu) and the fused cost -> first-conv producer (conv_tc2_kernel.cuh, FUSE = 1).
//
// ref: architecture/modeling/layers/inverse_warp_3d.py:40-47 + ATen grid_sampler_unnormalize (align_corners=True):
//      the pixel coordinate goes through normalise ((x / (W-1)) * 2 - 1) and un-normalise (((g + 1) / 2) * (W-1)),
//      reproduced here op by op so floor() and the tap weights match the reference bit for bit.
#pragma once
#include <cuda_runtime.h>

namespace tstereo {

// Row the warp samples: grid_sample's y coordinate lands up to ~1e-6 px off the integer row and would blend two
// rows with weights (1-eps, eps); the kernels sample the nearer row only (deviation <= 2e-6 * |R|).
__device__ __forceinline__ int warp_row(int yc, int H) {
    const float Hm1 = (float)(H - 1);
    const float gyn = __fsub_rn(__fmul_rn(__fdiv_rn((float)yc, Hm1), 2.0f), 1.0f);
    const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gyn, 1.0f), 2.0f), Hm1);
    const float fy = floorf(iy);
    int yn = (int)fy;
    if (__fsub_rn(iy, fy) > 0.5f) yn += 1;
    return min(max(yn, 0), H - 1);
}

// Horizontal taps of pixel x with candidate disparity dsp: columns (xa, xa + 1) with weights (wa, wb).  Taps outside
// the image get weight 0 on a clamped in-image address (xa in [0, W-2]), so no load needs a predicate.
// `live` = false (pixel outside the image / padding) -> both weights 0.
__device__ __forceinline__ void warp_col(int x, float dsp, int W, bool live, int& xa, float& wa, float& wb) {
    const float Wm1 = (float)(W - 1);
    const float gx = __fadd_rn((float)x, -dsp);
    const float gn = __fsub_rn(__fmul_rn(__fdiv_rn(gx, Wm1), 2.0f), 1.0f);
    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gn, 1.0f), 2.0f), Wm1);
    const float fx = floorf(ix);
    xa = 0;
    wa = 0.f;
    wb = 0.f;
    if (live && fx >= -1.0f && fx <= Wm1) {
        xa = (int)fx;
        wa = __fsub_rn(fx + 1.0f, ix);
        wb = __fsub_rn(ix, fx);
        if (xa < 0) {                     // tap 0 left of the image: read columns 0,1 as (tap1, unused)
            xa = 0;
            wa = wb;
            wb = 0.f;
        } else if (xa + 1 >= W) {         // tap 1 right of the image: read columns W-2,W-1 as (unused, tap0)
            xa = W - 2;
            wb = wa;
            wa = 0.f;
        }
    }
}

}  // namespace tstereo
