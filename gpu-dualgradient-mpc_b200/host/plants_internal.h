// plants_internal.h -- B condensed battery plants behind gpad_plants_t (host/plants.cpp, csrc/closed_loop.cu)
#pragma once
#include <algorithm>
#include <utility>
#include <vector>

#include "problem_internal.h"

struct gpad_plants_s {
    int n_u = 0, N = 0, n = 0, m = 0, n_par = 0, nx = 0, B = 0;
    std::vector<float> MG, GL;           // [B][n][m], [B][m][n] sequential layout (G_L already divided by L_b)
    std::vector<double> Kg, Ff, Bb;      // [B][n][n_par], [B][n][n_par], [B][m][n_par]
    std::vector<double> L;               // [B]
    std::vector<double> Bm;              // [B][nx][n_u] plant input matrices
    std::vector<double> b0, A;           // common: [m], [nx][nx]
    std::vector<std::pair<int, int>> blocks;
    // device copies of a shard's instance maps, one per (device, first, count) that ran a closed loop (csrc/closed_loop.cu)
    std::mutex dev_mutex;
    std::vector<void*> dev_cache;
};

namespace gpad {
void plants_dev_free(void* cache);
int closed_loop_plants_device(gpad_plants_t p, gpad_handle_t h, int first, int count, const double* x0, int samples,
                              const float* theta, const float* beta, int max_iter, int warm_start, double* x_traj,
                              double* u_traj);
}
