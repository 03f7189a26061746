// steps.cu -- step-compatible shims: one launch per GPAD step on caller-owned device buffers,
// replacing the five kernels the reference loop launches (main.cu:163-171,
// kernel_functions.cu:7-14, 16-64, 66-72, 142-200, 260-264).  Operators are read in the
// reference's flipped layout (M_G [m][n], G_L [n][m]) so its data files feed them unchanged.
//
// Both GEMVs are "A^T x" with the output index contiguous in memory: a CTA owns 32 output
// columns (one 128 B line per matrix row), its 32 warps stride over the rows, and the 32
// partial sums per column are combined through shared memory in a fixed order
// (deterministic, no atomics, no workspace).
#include "gpad_internal.h"

namespace gpad {

namespace {

constexpr int kGemvWarps = 32;

__global__ void step_one_kernel(const float* __restrict__ y, const float* __restrict__ y_prev,
                                float* __restrict__ w, float beta, int m) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const float yi = y[i];
        w[i] = __fadd_rn(yi, __fmul_rn(beta, __fsub_rn(yi, y_prev[i])));   // unfused, like the CPU build
    }
}

__global__ void step_three_kernel(float theta, const float* __restrict__ zhat, float* __restrict__ z, int n) {
    const float one_minus = 1.0f - theta;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        z[i] = __fadd_rn(__fmul_rn(one_minus, z[i]), __fmul_rn(theta, zhat[i]));
}

__global__ void copy_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// out[c] = epilogue(sum_r A[r*cols + c] * x[r]);  mode 0: - v1[c]            (step 2)
//                                                  mode 1: max(. + (v1[c] + v2[c]), 0) (step 4)
template <int MODE>
__global__ void __launch_bounds__(kGemvWarps * 32)
gemv_t_kernel(const float* __restrict__ A, const float* __restrict__ x, int rows, int cols,
              const float* __restrict__ v1, const float* __restrict__ v2, float* __restrict__ out) {
    __shared__ float part[kGemvWarps][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    if (c < cols) {
        int r = warp;
        for (; r + 3 * kGemvWarps < rows; r += 4 * kGemvWarps) {
            const float a0 = __ldg(A + (size_t)r * cols + c);
            const float a1 = __ldg(A + (size_t)(r + kGemvWarps) * cols + c);
            const float a2 = __ldg(A + (size_t)(r + 2 * kGemvWarps) * cols + c);
            const float a3 = __ldg(A + (size_t)(r + 3 * kGemvWarps) * cols + c);
            acc0 = fmaf(a0, __ldg(x + r), acc0);
            acc1 = fmaf(a1, __ldg(x + r + kGemvWarps), acc1);
            acc2 = fmaf(a2, __ldg(x + r + 2 * kGemvWarps), acc2);
            acc3 = fmaf(a3, __ldg(x + r + 3 * kGemvWarps), acc3);
        }
        for (; r < rows; r += kGemvWarps) acc0 = fmaf(__ldg(A + (size_t)r * cols + c), __ldg(x + r), acc0);
    }
    part[warp][lane] = (acc0 + acc1) + (acc2 + acc3);
    __syncthreads();
    if (warp == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kGemvWarps; ++k) s += part[k][lane];
        if (MODE == 0) {
            out[c] = s - v1[c];
        } else {
            s += v1[c] + v2[c];
            out[c] = 0.5f * (s + fabsf(s));
        }
    }
}

inline int ew_grid(int n) { return n <= 0 ? 1 : (n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256; }

}  // namespace

int launch_step_one(const float* y, const float* y_prev, float* w, float beta, int m, cudaStream_t s) {
    step_one_kernel<<<ew_grid(m), 256, 0, s>>>(y, y_prev, w, beta, m);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_step_three(float theta, const float* zhat, float* z, int n, cudaStream_t s) {
    step_three_kernel<<<ew_grid(n), 256, 0, s>>>(theta, zhat, z, n);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_copy(float* dst, const float* src, int n, cudaStream_t s) {
    copy_kernel<<<ew_grid(n), 256, 0, s>>>(dst, src, n);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_gemv_t(const float* A, const float* x, int rows, int cols, int mode, const float* v1,
                  const float* v2, float* out, cudaStream_t s) {
    const int grid = (cols + 31) / 32;
    if (mode == 0)
        gemv_t_kernel<0><<<grid, kGemvWarps * 32, 0, s>>>(A, x, rows, cols, v1, v2, out);
    else
        gemv_t_kernel<1><<<grid, kGemvWarps * 32, 0, s>>>(A, x, rows, cols, v1, v2, out);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

}  // namespace gpad
