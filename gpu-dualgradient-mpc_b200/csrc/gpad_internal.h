// gpad_internal.h -- shared declarations of libgpad_b200.so (not part of the public C ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "gpad.h"

namespace gpad {

// ---- error plumbing: every failure records a thread-local message and returns a status ----
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define GPAD_CUDA(call)                                                         \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return ::gpad::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define GPAD_REQUIRE(cond, ...)                     \
    do {                                            \
        if (!(cond)) {                              \
            ::gpad::set_error(__VA_ARGS__);         \
            return GPAD_ERR_INVALID_ARG;            \
        }                                           \
    } while (0)

#define GPAD_TRY(expr) do { int rc__ = (expr); if (rc__ != GPAD_OK) return rc__; } while (0)
#define GPAD_TRY_RC(expr) GPAD_TRY(expr)

inline int round_up(int v, int q) { return (v + q - 1) / q * q; }
inline size_t round_up_sz(size_t v, size_t q) { return (v + q - 1) / q * q; }

// ---- per-iteration scalars handed to kernels by value (main.cu:163,170 does the same) ----
struct IterScalars {
    float theta;      // theta_v
    float beta;       // beta_v: w_v = y_v + beta_v (y_v - y_{v-1}) is recomputed wherever it is needed
    int check;        // 1 when the termination quantities are reduced this iteration
    int store_zhat;   // zhat itself is an output only: written on the last iteration / in tolerance mode
};

// ---- device state of a batched (shared-operator) solve; all row-major [rows][ld] ----
struct BatchState {
    int B = 0;        // instances in this call
    int Bp = 0;       // capacity rounded up to the batch tile
    int n = 0, m = 0;
    int np = 0, mp = 0;  // leading dimensions (padded to 32 floats)
    float* g_P = nullptr;   // [Bp][np]
    float* p_D = nullptr;   // [Bp][mp]
    float* f = nullptr;     // [Bp][np] (optional)
    float* yb[3] = {nullptr, nullptr, nullptr};  // rotating y_{v-1} / y_v / y_{v+1}  [Bp][mp]; y_v lives in yb[v % 3]
    float* z = nullptr;     // [Bp][np]
    float* zhat = nullptr;  // [Bp][np]
    float* zh_hi = nullptr; // TF32X3 only
    float* zh_lo = nullptr;
    float* Pb[2] = {nullptr, nullptr};  // TF32X3 only: P_v = M_G y_v lives in Pb[v & 1] (P-formulation of product 1)
    // FP16X3 only (fixed-iteration solves): fp16 hi / lo of the row-scaled zhat, the inverse row scales, and the row
    // maxima of y_v (bit patterns; y_v's live in ymax[v & 1])
    uint16_t* zq_hi = nullptr;   // [Bp][np]
    uint16_t* zq_lo = nullptr;
    float* zinv = nullptr;       // [Bp]
    unsigned* ymax[2] = {nullptr, nullptr};   // [Bp]
    float* sbar = nullptr;  // [Bp][mp] averaged residual (termination only)
    float* red = nullptr;   // [Bp][8] per-instance reductions (termination only)
    int* done = nullptr;    // [Bp] instance stopped (termination mode)
    int* iters = nullptr;   // [Bp]
    int* status = nullptr;  // [Bp]
    float* max_viol = nullptr;
    float* gap = nullptr;
    int* active_count = nullptr;  // [2] instances still iterating, instances waiting for the dual-gap evaluation
    int* need = nullptr;          // [Bp] instance takes the dual-gap branch at this check
    float* zy = nullptr;          // [Bp][np] z_y scratch of the dual-gap evaluation (CUDA-core path)
    // tile retirement (tolerance mode): dense ascending list of the 128-row tiles that still hold a running instance
    int* tile_flags = nullptr;    // [Bp / 128]
    int* tile_list = nullptr;     // [Bp / 128]
    int* tile_count = nullptr;    // [1]
    unsigned long long* stat = nullptr;   // [2] instance-iterations scheduled (tiles x 128 x iterations) / needed (sum of iters)
    // compaction (tolerance mode, batch_compact.cu): perm[row] = original instance of a working row (-1: dead row)
    int* perm = nullptr;          // [Bp]
    int* holes = nullptr;         // [Bp]
    int* movers = nullptr;        // [Bp]
    int* compact_counts = nullptr;  // [2] running rows, moves
};

struct Operators {
    // fp32 operators in the sequential (K-contiguous) layout, zero padded:
    float* M_G = nullptr;   // [n_rows_pad][mp]   rows = outputs of GEMM 1 (n), K = m
    float* G_L = nullptr;   // [m_rows_pad][np]   rows = outputs of GEMM 2 (m), K = n
    float* M_G_lo = nullptr;  // TF32X3: M_G holds RN-tf32(M_G), *_lo the tf32-rounded remainder
    float* G_L_lo = nullptr;
    int n_rows_pad = 0, m_rows_pad = 0;
    // FP16X3: fp16 hi / lo of the operators, every row scaled by its own power of two, and 2^-e per row
    uint16_t *M_Gq_hi = nullptr, *M_Gq_lo = nullptr, *G_Lq_hi = nullptr, *G_Lq_lo = nullptr;
    float *M_G_inv = nullptr, *G_L_inv = nullptr;
};

struct BatchKernelArgs;

// ---- kernel launchers (defined in the .cu files) ----
int launch_pad_rows(float* dst, int ld, int rows_total, const float* src, int len, int B, cudaStream_t s);
int launch_unpad_rows(float* dst, int len, int B, const float* src, int ld, cudaStream_t s);
int launch_unpad_y(float* dst_next, float* dst_cur, float* dst_w, int m, int B, const float* yb0, const float* yb1,
                   const float* yb2, int mp, const int* iters, const float* beta_dev, cudaStream_t s);
int launch_batch_init(const BatchState& st, bool checking, cudaStream_t s);
int launch_batch_reset_term(const BatchState& st, int max_iter, cudaStream_t s);
int launch_perm_identity(const BatchState& st, cudaStream_t s);
int launch_compact(const BatchState& st, const BatchState& archive, int rows_bound, bool have_f, cudaStream_t s);
int launch_archive_all(const BatchState& st, const BatchState& archive, int rows_bound, cudaStream_t s);
int launch_batch_tiles(const BatchState& st, int iterations, bool rebuild, cudaStream_t s);
int launch_batch_iter_sum(const BatchState& st, cudaStream_t s);
int launch_batch_decide(const BatchState& st, int iter_done, float L, float eps_g, float eps_V, bool have_f, cudaStream_t s);
int launch_batch_decide_dual(const BatchState& st, int iter_done, float L, float eps_V, cudaStream_t s);
int launch_batch_finite(const BatchState& st, const float* y_next, cudaStream_t s);
int launch_simt_product(int phase, const Operators& op, const BatchKernelArgs& args, int Bp, cudaStream_t s);
int launch_step_one(const float* y, const float* y_prev, float* w, float beta, int m, cudaStream_t s);
int launch_gemv_t(const float* A, const float* x, int rows, int cols, int mode, const float* v1,
                  const float* v2, float* out, cudaStream_t s);
int launch_copy(float* dst, const float* src, int n, cudaStream_t s);
int launch_step_three(float theta, const float* zhat, float* z, int n, cudaStream_t s);

}  // namespace gpad
