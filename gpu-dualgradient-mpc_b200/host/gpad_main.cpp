// gpad_main.cpp -- the repaired equivalent of the reference driver Code/CUDA/FinalProject/main.cu
// (which does not compile as shipped: SURVEY fact 5).  Host code is plain C++ calling the device
// only through the C ABI of libgpad_b200.so (include/gpad.h).
//
//   gpad_main <data file>            solve the problem of a reference-format data file
//                                    (main.cu:29-67; operators in the flipped layout its kernels read)
//   gpad_main --battery n_u N        generate the battery-balancing problem of gpad.m instead
//                                    (the reference's inputs_manysets/*.txt are git-LFS stubs)
//   options: --iters K (default 100 = N_v, main.cu:87)   --eps E (tolerance mode, check every iteration)
//            --write FILE (also write the generated problem as a reference-format data file)
//
// Prints what main.cu:188-190 prints (sizes, total and average time per iteration) and, because
// the reference never checked its result, the norms of the five vectors it copies back
// (main.cu:176-180).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gpad.h"

static double norm_inf(const std::vector<float>& v) {
    double m = 0.0;
    for (float x : v) m = std::fmax(m, std::fabs((double)x));
    return m;
}

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int rc__ = (call);                                                                   \
        if (rc__ != GPAD_OK) {                                                               \
            fprintf(stderr, "%s failed: %s (%s)\n", #call, gpad_status_string(rc__), gpad_last_error()); \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

int main(int argc, char** argv) {
    const char* path = nullptr;
    const char* write_path = nullptr;
    int bat_nu = 0, bat_N = 0, iters = 100;
    float eps = 0.f;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--battery") && i + 2 < argc) { bat_nu = atoi(argv[++i]); bat_N = atoi(argv[++i]); }
        else if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--eps") && i + 1 < argc) eps = (float)atof(argv[++i]);
        else if (!strcmp(argv[i], "--write") && i + 1 < argc) write_path = argv[++i];
        else path = argv[i];
    }
    if (!path && bat_nu <= 0) {
        fprintf(stderr, "usage: %s <data file> | --battery n_u N [--iters K] [--eps E] [--write FILE]\n", argv[0]);
        return 2;
    }

    gpad_file_t file;
    memset(&file, 0, sizeof(file));
    if (path) {
        CHECK(gpad_file_read(path, &file));
    } else {
        gpad_problem_t prob;
        CHECK(gpad_problem_battery(bat_nu, bat_N, &prob));
        int n_par = 0;
        CHECK(gpad_problem_dims(prob, &file.n_u, &file.N, &file.m, &n_par, &file.L));
        const size_t n = (size_t)file.n_u * file.N, m = file.m;
        file.num_iterations = iters;
        file.M_G = (float*)malloc(sizeof(float) * n * m); file.G_L = (float*)malloc(sizeof(float) * n * m);
        file.g_P = (float*)malloc(sizeof(float) * n); file.p_D = (float*)malloc(sizeof(float) * m);
        file.theta = (float*)malloc(sizeof(float) * iters); file.beta = (float*)malloc(sizeof(float) * iters);
        CHECK(gpad_problem_operators(prob, GPAD_LAYOUT_FLIPPED, file.M_G, file.G_L));
        std::vector<double> x0(n_par);
        const double x0_10[10] = {-0.1, 0.45, -0.09, 0.05, 0, -0.05, 0.3, 0.2, 0.25, -0.45};   // gpad.m:10
        for (int i = 0; i < n_par; ++i) x0[i] = n_par == 10 ? x0_10[i] : 0.4 * std::sin(1.0 + 2.0 * i);
        CHECK(gpad_problem_instances(prob, 1, x0.data(), file.g_P, file.p_D, nullptr));
        CHECK(gpad_schedule(file.theta, file.beta, iters, GPAD_SCHEDULE_PAPER));
        gpad_problem_destroy(prob);
        if (write_path) CHECK(gpad_file_write(write_path, &file));
    }
    if (file.num_iterations < iters) {
        // main.cu reads theta[v]/beta[v] out of bounds here; we refuse instead
        fprintf(stderr, "data file holds %d schedule entries, %d iterations requested\n", file.num_iterations, iters);
        return 1;
    }
    const int n = file.n_u * file.N, m = file.m;

    gpad_config_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.n_u = file.n_u; cfg.N = file.N; cfg.m = m; cfg.L = file.L;
    cfg.layout = GPAD_LAYOUT_FLIPPED; cfg.mode = GPAD_MODE_LATENCY; cfg.precision = GPAD_PREC_FP32;
    cfg.max_batch = 1; cfg.device = -1; cfg.operators_mem = GPAD_MEM_HOST;
    gpad_handle_t h;
    CHECK(gpad_setup(&cfg, file.M_G, file.G_L, &h));

    std::vector<float> y_vp1(m), y_v(m), w_v(m), z_v(n), zhat_v(n);
    int it_done = 0, status = 0;
    float viol = 0.f, gap = 0.f;
    gpad_solve_args_t a;
    memset(&a, 0, sizeof(a));
    a.batch = 1; a.mem = GPAD_MEM_HOST;
    a.g_P = file.g_P; a.p_D = file.p_D; a.theta = file.theta; a.beta = file.beta; a.max_iter = iters;
    a.check_every = eps > 0.f ? 1 : 0; a.eps_g = eps; a.eps_V = eps;
    a.y_next = y_vp1.data(); a.y = y_v.data(); a.w = w_v.data(); a.z = z_v.data(); a.zhat = zhat_v.data();
    a.iters = &it_done; a.status = &status; a.max_viol = &viol; a.gap = &gap;
    CHECK(gpad_solve(h, &a));   // warm-up (module load, schedule upload)
    const auto t0 = std::chrono::steady_clock::now();
    CHECK(gpad_solve(h, &a));
    const auto t1 = std::chrono::steady_clock::now();
    const long usec = (long)std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();

    printf("%s\n", gpad_describe(h));
    printf("n_u = %d, N = %d, m = %d\n", file.n_u, file.N, m);                                      // main.cu:188
    printf("Total GPU Execution Time over %d trial(s) = %ld usec\n", it_done, usec);               // main.cu:189
    printf("Avg. GPU Execution Time over %d trial(s) = %ld usec\n", it_done, usec / (it_done ? it_done : 1));   // main.cu:190
    printf("status = %d, iterations = %d, max violation = %g, gap = %g\n", status, it_done, viol, gap);
    printf("|y_vp1| = %.8g  |y_v| = %.8g  |z_v| = %.8g  |zhat_v| = %.8g  |w_v| = %.8g\n", norm_inf(y_vp1), norm_inf(y_v),
           norm_inf(z_v), norm_inf(zhat_v), norm_inf(w_v));
    gpad_destroy(h);
    gpad_file_free(&file);
    return 0;
}
