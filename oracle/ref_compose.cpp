/*
 * ref_compose.cpp -- drives the UNMODIFIED reference step functions
 * (/root/reference/Code/CUDA/FinalProject/src/seq_functions.cpp, compiled where it lies by
 * oracle/Makefile into oracle/_ref/libgpad_ref.so) in the order of the reference GPU loop,
 * main.cu:160-175.  The reference ships no sequential whole-solve driver,
 * so this composition is ours; every arithmetic operation happens inside the reference's
 * own object code.  TEST INFRASTRUCTURE ONLY: used to pin oracle/gpad_oracle.c and as the
 * "reference" CPU baseline of bench.py.
 */
#include <cstring>
#include <cstddef>
#include <atomic>
#include <thread>
#include <vector>
#include "seq_functions.h"

extern "C" {

/* one QP, fixed iteration count; layouts are the sequential ones (M_G [n][m], G_L [m][n]).
 * y_next/y/z/zhat/w are the five vectors main.cu:176-180 copies back. */
void ref_solve_fixed(int n_u, int N, int m, const float* M_G, const float* g_P, const float* G_L,
                     const float* p_D, const float* theta, const float* beta, int iters,
                     const float* y0, const float* y_prev0,
                     float* y_next, float* y, float* z, float* zhat, float* w) {
    const int n = n_u * N;
    if (y0) std::memcpy(y_next, y0, sizeof(float) * m); else std::memset(y_next, 0, sizeof(float) * m);
    if (y_prev0) std::memcpy(y, y_prev0, sizeof(float) * m); else std::memset(y, 0, sizeof(float) * m);
    std::memset(z, 0, sizeof(float) * n);
    std::memset(zhat, 0, sizeof(float) * n);
    std::memset(w, 0, sizeof(float) * m);
    for (int v = 0; v < iters; ++v) {
        StepOneGPADSequential(y_next, y, w, beta[v], m);
        StepTwoGPADSequential(M_G, w, g_P, zhat, N, n_u, m);
        std::memcpy(y, y_next, sizeof(float) * m);
        StepThreeGPADSequential(theta[v], n, z, zhat, z);
        StepFourGPADSequential(G_L, y_next, w, p_D, zhat, N, n_u, m);
    }
}

/* B independent QPs sharing operators, instance-major g_P [B][n], p_D [B][m]; std::thread
 * workers pull instances from a shared counter; returns threads used */
int ref_solve_fixed_batch(int n_u, int N, int m, const float* M_G, const float* G_L,
                          const float* theta, const float* beta, int iters, int B,
                          const float* g_P, const float* p_D,
                          float* y_next, float* y, float* z, float* zhat, float* w, int nthreads) {
    const int n = n_u * N;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads > B) nthreads = B;
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int b = next.fetch_add(1, std::memory_order_relaxed);
            if (b >= B) break;
            ref_solve_fixed(n_u, N, m, M_G, g_P + (size_t)b * n, G_L, p_D + (size_t)b * m, theta,
                            beta, iters, nullptr, nullptr, y_next + (size_t)b * m,
                            y + (size_t)b * m, z + (size_t)b * n, zhat + (size_t)b * n,
                            w + (size_t)b * m);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return nthreads;
}

}  // extern "C"
