// Probe: what bounds the block_cost write pattern?  nvcc -arch=sm_100a -O3 bc_probe.cu -o bc_probe
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// mode bit0: do loads (L and R gathers); bit1: do the big stores; bit2: linear instead of plane-scattered stores
template <int SC>
__global__ void __launch_bounds__(256) probe(const float* __restrict__ L, const float* __restrict__ R,
                                             float* __restrict__ out, int C, int H, int W, int D, int mode) {
    const int G = C >> 3;
    const int g = blockIdx.z % G, b = blockIdx.z / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)y * W + x;
    float acc = 0.f;
    for (int c = 0; c < 8; ++c) {
        const int ch = g * 8 + c;
        float l = 1.f, rv[SC];
#pragma unroll
        for (int s = 0; s < SC; ++s) rv[s] = 2.f;
        if (mode & 1) {
            l = __ldg(L + ((size_t)b * C + ch) * HW + pix);
#pragma unroll
            for (int s = 0; s < SC; ++s) {
                int xs = x - 3 * s - 1; if (xs < 0) xs = 0;
                const float* rr = R + ((size_t)b * C + ch) * HW + (size_t)y * W + xs;
                rv[s] = 0.3f * __ldg(rr) + 0.7f * __ldg(rr + 1);
            }
        }
        if (mode & 2) {
#pragma unroll
            for (int s = 0; s < SC; ++s) {
                if (mode & 4) {   // linear: each CTA-channel writes a contiguous chunk
                    size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
                    size_t o = ((cta * 8 + c) * SC + s) * 2 * 256 + threadIdx.x;
                    out[o] = l; out[o + 256] = rv[s];
                } else {
                    out[(((size_t)b * 2 * C + ch) * D + s) * HW + pix] = l;
                    out[(((size_t)b * 2 * C + C + ch) * D + s) * HW + pix] = rv[s];
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < SC; ++s) acc += l * rv[s];
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void copy4(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void fill4(float4* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = make_float4(1, 2, 3, 4);
}

int main() {
    const int B = 1, C = 128, H = 136, W = 240, D = 5;
    const size_t HW = (size_t)H * W, nin = (size_t)B * C * HW, nout = (size_t)B * 2 * C * D * HW + 4096;
    float *L, *R, *out, *big;
    CK(cudaMalloc(&L, nin * 4)); CK(cudaMalloc(&R, nin * 4)); CK(cudaMalloc(&out, nout * 4));
    const size_t nbig = 512u << 20; CK(cudaMalloc(&big, nbig));
    CK(cudaMemset(L, 0, nin * 4)); CK(cudaMemset(R, 0, nin * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid((W + 31) / 32, (H + 7) / 8, B * C / 8);
    const char* names[] = {"none", "loads only", "scatter stores only", "loads + scatter stores", "-", "-", "linear stores only", "loads + linear stores"};
    for (int mode : {1, 2, 3, 6, 7}) {
        float best = 1e9;
        for (int it = 0; it < 5; ++it) {
            cudaMemsetAsync(big, 0, nbig);   // flush L2 (leaves dirty lines, like a real pipeline would)
            cudaEventRecord(e0);
            probe<5><<<grid, 256>>>(L, R, out, C, H, W, D, mode);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double wr = (mode & 2) ? 2.0 * C * D * HW * 4 * B : 0, rd = (mode & 1) ? 2.0 * nin * 4 : 0;
        printf("%-26s %8.1f us   (%.0f MB written, %.0f MB unique read) -> %.0f GB/s\n", names[mode], best * 1e3, wr / 1e6, rd / 1e6, (wr + rd) / best / 1e6);
    }
    for (int k = 0; k < 2; ++k) {
        float best = 1e9;
        const size_t n4 = (size_t)2 * C * D * HW / 4;
        for (int it = 0; it < 5; ++it) {
            cudaMemsetAsync(big, 0, nbig);
            cudaEventRecord(e0);
            if (k == 0) fill4<<<148 * 8, 256>>>((float4*)out, n4); else copy4<<<148 * 8, 256>>>((const float4*)big, (float4*)out, n4);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-26s %8.1f us   -> %.0f GB/s (write bytes only)\n", k ? "float4 copy" : "float4 fill", best * 1e3, n4 * 16.0 / best / 1e6);
    }
    return 0;
}
