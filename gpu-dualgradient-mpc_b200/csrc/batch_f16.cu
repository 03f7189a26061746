// batch_f16.cu -- row quantisation kernels of GPAD_PREC_FP16X3 (throughput mode, shared operators).
//
// The tcgen05 kind::f16 MMAs take fp16 operands: half the bytes and half the instructions of kind::tf32 for the same
// 11 significant bits per term.  fp16's narrow exponent is handled by scaling every operand ROW by a power of two
// chosen from the row's largest magnitude (tc_ptx.cuh:f16_scale_exp), splitting the scaled value into fp16 hi + lo
// and undoing both scales on the fp32 accumulator -- all exact.  The operators are quantised once at gpad_setup; the
// dual iterate y_v is quantised inside product 1 (its row maxima come from product 2's epilogue); the primal iterate
// zhat_v needs the maximum of a whole row before any of it can be split, which no tile of product 1 sees, so this
// file's quantize_rows_kernel runs between the two products (one read of zhat, 4 n bytes per instance).
#include <cuda_fp16.h>

#include "batch_tc.h"
#include "gpad_internal.h"
#include "tc_ptx.cuh"

namespace gpad {
namespace tc {

namespace {

constexpr int kRowWarps = 8;

// one warp per row.  CACHE float4 per lane are kept in registers between the two passes (rows up to CACHE * 128
// floats are read once); longer rows are read again from L1 / L2.
template <int CACHE>
__global__ void __launch_bounds__(32 * kRowWarps)
quantize_rows_kernel(const float* __restrict__ src, int ld, int rows, uint2* __restrict__ hi, uint2* __restrict__ lo,
                     float* __restrict__ inv, unsigned* __restrict__ zero_rows) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    // programmatic dependent launch (a no-op without the launch attribute): the previous kernel's zhat is needed from here
    // on; the next kernel may start its prologue (its own dependency wait covers this grid's writes)
    grid_dependency_wait();
    grid_launch_dependents();
    if (row >= rows) return;
    const float4* x = reinterpret_cast<const float4*>(src + (size_t)row * ld);
    const int nv = ld >> 2;
    float4 c[CACHE > 0 ? CACHE : 1];
    float mx = 0.f;
#pragma unroll
    for (int k = 0; k < CACHE; ++k) {
        const int i = lane + 32 * k;
        c[k] = i < nv ? x[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(c[k].x), fabsf(c[k].y)), fmaxf(fabsf(c[k].z), fabsf(c[k].w))));
    }
    for (int i = lane + 32 * CACHE; i < nv; i += 32) {
        const float4 v = x[i];
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    mx = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(mx)));      // |x| >= 0: bits order like values
    const int e = f16_scale_exp(mx);
    const float sc = pow2f(e);
    uint2* h = hi + (size_t)row * nv;
    uint2* l = lo + (size_t)row * nv;
#pragma unroll
    for (int k = 0; k < CACHE; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) {
            uint2 ph, pl;
            split_f16x2(c[k].x * sc, c[k].y * sc, ph.x, pl.x);
            split_f16x2(c[k].z * sc, c[k].w * sc, ph.y, pl.y);
            h[i] = ph;
            l[i] = pl;
        }
    }
    for (int i = lane + 32 * CACHE; i < nv; i += 32) {
        const float4 v = x[i];
        uint2 ph, pl;
        split_f16x2(v.x * sc, v.y * sc, ph.x, pl.x);
        split_f16x2(v.z * sc, v.w * sc, ph.y, pl.y);
        h[i] = ph;
        l[i] = pl;
    }
    if (lane == 0) {
        inv[row] = pow2f(-e);
        if (zero_rows) zero_rows[row] = 0u;
    }
}

__global__ void __launch_bounds__(32 * kRowWarps)
rowmax_kernel(const float* __restrict__ src, int ld, int rows, float* __restrict__ rowmax) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* x = reinterpret_cast<const float4*>(src + (size_t)row * ld);
    const int nv = ld >> 2;
    float mx = 0.f;
    for (int i = lane; i < nv; i += 32) {
        const float4 v = x[i];
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    mx = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(mx)));
    if (lane == 0) rowmax[row] = mx;
}

}  // namespace

int launch_quantize_rows(const float* src, int ld, int rows, uint16_t* hi, uint16_t* lo, float* inv, unsigned* zero_rows,
                         cudaStream_t s, bool pdl) {
    if (rows <= 0) return GPAD_OK;
    if (ld % 4) { set_error("quantize_rows: leading dimension %d is not a multiple of 4", ld); return GPAD_ERR_INVALID_ARG; }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((rows + kRowWarps - 1) / kRowWarps); lc.blockDim = dim3(32 * kRowWarps); lc.dynamicSmemBytes = 0; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
    const float* src_c = src;
    uint2* h2 = reinterpret_cast<uint2*>(hi);
    uint2* l2 = reinterpret_cast<uint2*>(lo);
    if (ld <= 512) GPAD_CUDA(cudaLaunchKernelEx(&lc, quantize_rows_kernel<4>, src_c, ld, rows, h2, l2, inv, zero_rows));
    else GPAD_CUDA(cudaLaunchKernelEx(&lc, quantize_rows_kernel<0>, src_c, ld, rows, h2, l2, inv, zero_rows));
    return GPAD_OK;
}

int launch_rowmax(const float* src, int ld, int rows, float* rowmax, cudaStream_t s) {
    if (rows <= 0) return GPAD_OK;
    if (ld % 4) { set_error("rowmax: leading dimension %d is not a multiple of 4", ld); return GPAD_ERR_INVALID_ARG; }
    rowmax_kernel<<<(rows + kRowWarps - 1) / kRowWarps, 32 * kRowWarps, 0, s>>>(src, ld, rows, rowmax);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

}  // namespace tc
}  // namespace gpad
