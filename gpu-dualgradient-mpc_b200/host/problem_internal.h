// problem_internal.h -- the condensed problem behind gpad_problem_t, shared by host/problem.cpp, host/plants.cpp
// and csrc/closed_loop.cu (plain C++, no CUDA types).
#pragma once
#include <mutex>
#include <utility>
#include <vector>

#include "gpad.h"

namespace gpad {

struct Mat {
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};

}  // namespace gpad

struct gpad_problem_s {
    int n_u = 0, N = 0, n = 0, m = 0, n_par = 0, nx = 0;
    double L = 0.0;
    gpad::Mat H, G;        // n x n, m x n
    gpad::Mat MG;          // n x m   = -H^-1 G'
    gpad::Mat Ff, Kg;      // n x n_par  (f = Ff p, g_P = Kg p)
    gpad::Mat Bb;          // m x n_par
    std::vector<double> b0;
    gpad::Mat A, B;        // plant (nx x nx, nx x n_u)
    std::vector<std::pair<int, int>> blocks;    // dual blocks {offset, rows per stage}: the receding-horizon shift
    // device copies of Kg / Ff / Bb / b0 / A / B, one per device that evaluated instance maps (csrc/closed_loop.cu)
    std::mutex dev_mutex;
    std::vector<void*> dev_cache;
};

namespace gpad {

int build_battery(int n_u, int N, const double* cap_scale, gpad_problem_s** out);
void problem_dev_free(void* cache);
int closed_loop_device(gpad_problem_t p, gpad_handle_t h, int B, const double* x0, const double* xref, int samples,
                       const float* theta, const float* beta, int max_iter, int warm_start, double* x_traj, double* u_traj);

}  // namespace gpad
