// batch_common.cuh -- per-element epilogues of the two batched products, shared by the
// CUDA-core (batch_simt.cu) and tcgen05 (batch_tc.cu) paths, so both apply exactly the same
// arithmetic after the accumulator:
//
//   product 1 (step 2 + 3, kernel_functions.cu:16-72):   acc = <M_G[i,:], w_b>
//        zhat = acc - g_P;  z = (1-theta) z + theta zhat
//   product 2 (step 1 + 4, kernel_functions.cu:7-14, 142-200):  acc = <G_L[i,:], zhat_b>
//        w = y + beta (y - y_prev);  s = acc + (w + p_D);  y+ = (s + |s|)/2
// w is never stored: product 1 rebuilds it (and its tf32 split) while staging its A operand,
// product 2 rebuilds it here, which saves three m-sized HBM arrays per iteration.
#pragma once
#include <cuda_runtime.h>

#include "gpad_internal.h"

namespace gpad {

// red[b][k]: 0 max sbar, 1 max rhat, 2 min w, 3 sum w*rhat, 4 sum w*dot, 5 sum f*zhat, 6 nonfinite
constexpr int kRedStride = 8;

struct BatchKernelArgs {
    // leading dimensions
    int n, m, np, mp;
    int B;                 // valid instances
    IterScalars it;
    int checking;          // termination enabled for this solve (sbar recurrence maintained)
    float L;
    // state (see BatchState)
    const float* g_P;
    const float* p_D;
    const float* f;
    const float* y_prev;   // y_{v-1}
    const float* y_cur;    // y_v
    float* y_next;         // y_{v+1}
    float* z;
    float* zhat;
    float* zh_hi;
    float* zh_lo;
    float* sbar;
    float* red;
    const int* done;       // per-instance "stopped" flag (null in fixed-iteration mode)
    // P-formulation of product 1 (tcgen05 path): the GEMM computes P_v = M_G y_v and the epilogue forms
    // M_G w_v = P_v + beta (P_v - P_{v-1}) -- linear in y, so y_{v-1} never has to be staged as an MMA operand
    int p_only;            // warm start: this launch only produces P_{-1} = M_G y_{-1}
    const float* P_prev;   // [Bp][np] P_{v-1}
    float* P_cur;          // [Bp][np] P_v
    // dual-gap evaluation launches (termination branch 3): product 1 forms z_y = M_G y_{v+1} - g_P for the
    // flagged instances (need[b]) and f'z_y, product 2 reduces y'(G_L z_y) and y'p_D; nothing else is updated
    int dual;
    const int* need;       // [Bp] instance takes the dual-gap branch at this check
    float* zy;             // [Bp][np] z_y (CUDA-core path; the tcgen05 path reuses zh_hi / zh_lo)
    // tolerance mode, tile retirement: the 128-row batch tiles that still hold a running instance, as a dense list
    // (batch_simt.cu:batch_tiles_*); null in fixed-iteration solves (every tile runs)
    const int* tile_list;
    const int* tile_count;
    const int* dual_count; // dual-gap launches return at once when no instance waits for the evaluation
    // GPAD_PREC_FP16X3 (fp16 hi / lo operands scaled per row by powers of two; tc_ptx.cuh:f16_scale_exp)
    const float* a_rowmax;   // product 1: [Bp] max_k y_v[b][k], the transform warps derive the row scale from it
    const float* a_rowinv;   // product 2: [Bp] 2^-e of the zhat row scale (written by zsplit_kernel)
    const float* b_colinv;   // [rows_pad] 2^-e of the operator row (= output column) scales
    unsigned* next_rowmax;   // product 2: [Bp] bit pattern of max_k y_{v+1}[b][k] (atomicMax; zeroed by zsplit_kernel)
};

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// x = hi + lo + O(2^-22 |x|): hi = RN_tf32(x), lo = RN_tf32(x - hi) (x - hi is exact in fp32)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tf32_rn(x);
    lo = tf32_rn(x - hi);
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    // monotone int mapping: valid for any mix of signs
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// ---- product 1 epilogue for one element (instance b, row i of M_G) ----
template <bool SPLIT>
__device__ __forceinline__ void epilogue1(const BatchKernelArgs& a, int b, int i, float acc, float& f_zhat) {
    const size_t o = (size_t)b * a.np + i;
    if (a.dual) {
        if (!a.need[b]) return;
        const float zy = acc - a.g_P[o];
        if (SPLIT) {
            float hi, lo;
            split_tf32(zy, hi, lo);
            a.zh_hi[o] = hi;
            a.zh_lo[o] = lo;
        } else {
            a.zy[o] = zy;
        }
        f_zhat = fmaf(a.f[o], zy, f_zhat);
        return;
    }
    const float zh = acc - a.g_P[o];
    a.z[o] = __fadd_rn(__fmul_rn(1.0f - a.it.theta, a.z[o]), __fmul_rn(a.it.theta, zh));   // unfused like the CPU build
    a.zhat[o] = zh;
    if (SPLIT) {
        float hi, lo;
        split_tf32(zh, hi, lo);
        a.zh_hi[o] = hi;
        a.zh_lo[o] = lo;
    }
    if (a.it.check && a.f) f_zhat = fmaf(a.f[o], zh, f_zhat);
}

struct Red2 {
    float max_sbar = -INFINITY, max_rhat = -INFINITY, min_w = INFINITY, w_rhat = 0.f, w_dot = 0.f, bad = 0.f;
};

__device__ __forceinline__ float momentum(float y, float y_prev, float beta) {
    return __fadd_rn(y, __fmul_rn(beta, __fsub_rn(y, y_prev)));      // step 1, unfused like the CPU build
}

// ---- product 2 epilogue for one element (instance b, row i of G_L) ----
__device__ __forceinline__ void epilogue2(const BatchKernelArgs& a, int b, int i, float acc, Red2& r) {
    const size_t o = (size_t)b * a.mp + i;
    if (a.dual) {           // y = y_{v+1} is an input here; red slots 3 / 4 collect y'(G_L z_y) and y'p_D
        if (a.need[b]) {
            const float y = a.y_next[o];
            r.w_rhat = fmaf(y, acc, r.w_rhat);
            r.w_dot = fmaf(y, a.p_D[o], r.w_dot);
        }
        return;
    }
    const float wv = momentum(a.y_cur[o], a.y_prev[o], a.it.beta), pd = a.p_D[o];
    const float s = acc + (wv + pd);
    const float yn = 0.5f * (s + fabsf(s));
    a.y_next[o] = yn;
    if (a.checking) {
        const float rhat = acc + pd;
        const float sb = __fadd_rn(__fmul_rn(1.0f - a.it.theta, a.sbar[o]), __fmul_rn(a.it.theta, rhat));
        a.sbar[o] = sb;
        if (a.it.check) {
            r.max_sbar = fmaxf(r.max_sbar, sb);
            r.max_rhat = fmaxf(r.max_rhat, rhat);
            r.min_w = fminf(r.min_w, wv);
            r.w_rhat = fmaf(wv, rhat, r.w_rhat);
            r.w_dot = fmaf(wv, acc, r.w_dot);
            if (!isfinite(yn)) r.bad = 1.f;
        }
    }
}

__device__ __forceinline__ void flush_red2(const BatchKernelArgs& a, int b, const Red2& r) {
    float* red = a.red + (size_t)b * kRedStride;
    atomic_max_float(red + 0, r.max_sbar);
    atomic_max_float(red + 1, r.max_rhat);
    atomic_min_float(red + 2, r.min_w);
    atomicAdd(red + 3, r.w_rhat);
    atomicAdd(red + 4, r.w_dot);
    if (r.bad > 0.f) atomicExch(red + 6, 1.0f);
}

}  // namespace gpad
