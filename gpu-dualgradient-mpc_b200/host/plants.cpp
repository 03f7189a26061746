// plants.cpp -- per-instance plants (BASELINE config 5: "per-instance plant matrices"): B battery packs whose cell
// capacities differ per instance, condensed one by one with the same code as gpad_problem_battery (gpad.m:4-85 with
// gpad.m:18 scaled, acceldualgrad.m:9-23) on host threads.  Every plant has its own M_G, G_L, L and affine instance
// maps; the constraint offsets b0 and A = I are common.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "plants_internal.h"

extern "C" {

int gpad_plants_battery(int n_u, int N, int B, const double* capacity_scale, int threads, gpad_plants_t* out) {
    if (!out || n_u < 1 || N < 1 || B < 1 || !capacity_scale) return GPAD_ERR_INVALID_ARG;
    gpad_plants_s* P = new gpad_plants_s;
    const int n = n_u * N, m = 4 * n + 2 * N;
    P->n_u = n_u; P->N = N; P->n = n; P->m = m; P->n_par = n_u; P->nx = n_u; P->B = B;
    P->MG.resize((size_t)B * n * m); P->GL.resize((size_t)B * n * m);
    P->Kg.resize((size_t)B * n * n_u); P->Ff.resize((size_t)B * n * n_u); P->Bb.resize((size_t)B * m * n_u);
    P->L.resize(B); P->Bm.resize((size_t)B * n_u * n_u);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    threads = std::max(1, std::min(threads, B));
    std::atomic<int> next(0), failed(0);
    auto work = [&]() {
        for (;;) {
            const int b = next.fetch_add(1, std::memory_order_relaxed);
            if (b >= B || failed.load()) break;
            gpad_problem_s* q = nullptr;
            if (gpad::build_battery(n_u, N, capacity_scale + (size_t)b * n_u, &q) != GPAD_OK) { failed = 1; break; }
            const double invL = 1.0 / q->L;
            float* mg = &P->MG[(size_t)b * n * m];
            float* gl = &P->GL[(size_t)b * n * m];
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < m; ++j) mg[(size_t)i * m + j] = (float)q->MG(i, j);
            for (int i = 0; i < m; ++i)
                for (int j = 0; j < n; ++j) gl[(size_t)i * n + j] = (float)(q->G(i, j) * invL);        // acceldualgrad.m:22
            std::memcpy(&P->Kg[(size_t)b * n * n_u], q->Kg.a.data(), sizeof(double) * n * n_u);
            std::memcpy(&P->Ff[(size_t)b * n * n_u], q->Ff.a.data(), sizeof(double) * n * n_u);
            std::memcpy(&P->Bb[(size_t)b * m * n_u], q->Bb.a.data(), sizeof(double) * m * n_u);
            std::memcpy(&P->Bm[(size_t)b * n_u * n_u], q->B.a.data(), sizeof(double) * n_u * n_u);
            P->L[b] = q->L;
            if (b == 0) { P->b0 = q->b0; P->A = q->A.a; P->blocks = q->blocks; }
            delete q;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (failed.load()) { delete P; return GPAD_ERR_INVALID_ARG; }
    *out = P;
    return GPAD_OK;
}

int gpad_plants_destroy(gpad_plants_t p) {
    if (!p) return GPAD_OK;
    for (void* d : p->dev_cache) gpad::plants_dev_free(d);
    delete p;
    return GPAD_OK;
}

int gpad_plants_dims(gpad_plants_t p, int* n_u, int* N, int* m, int* n_par, int* B) {
    if (!p) return GPAD_ERR_INVALID_ARG;
    if (n_u) *n_u = p->n_u;
    if (N) *N = p->N;
    if (m) *m = p->m;
    if (n_par) *n_par = p->n_par;
    if (B) *B = p->B;
    return GPAD_OK;
}

int gpad_plants_operators(gpad_plants_t p, int layout, float* M_G, float* G_L, float* L) {
    if (!p || !M_G || !G_L) return GPAD_ERR_INVALID_ARG;
    const int n = p->n, m = p->m;
    const size_t per = (size_t)n * m;
    if (layout == GPAD_LAYOUT_SEQUENTIAL) {
        std::memcpy(M_G, p->MG.data(), sizeof(float) * per * p->B);
        std::memcpy(G_L, p->GL.data(), sizeof(float) * per * p->B);
    } else if (layout == GPAD_LAYOUT_FLIPPED) {
        for (int b = 0; b < p->B; ++b) {
            const float* mg = &p->MG[b * per]; const float* gl = &p->GL[b * per];
            float* fm = M_G + b * per; float* fg = G_L + b * per;
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < m; ++j) fm[(size_t)j * n + i] = mg[(size_t)i * m + j];
            for (int i = 0; i < m; ++i)
                for (int j = 0; j < n; ++j) fg[(size_t)j * m + i] = gl[(size_t)i * n + j];
        }
    } else {
        return GPAD_ERR_INVALID_ARG;
    }
    if (L) for (int b = 0; b < p->B; ++b) L[b] = (float)p->L[b];
    return GPAD_OK;
}

int gpad_plants_instances(gpad_plants_t p, const double* params, float* g_P, float* p_D, float* f) {
    if (!p || !params) return GPAD_ERR_INVALID_ARG;
    const int n = p->n, m = p->m, np = p->n_par;
    for (int b = 0; b < p->B; ++b) {
        const double* par = params + (size_t)b * np;
        const double* Kg = &p->Kg[(size_t)b * n * np];
        const double* Ff = &p->Ff[(size_t)b * n * np];
        const double* Bb = &p->Bb[(size_t)b * m * np];
        const double invL = 1.0 / p->L[b];
        if (g_P || f)
            for (int i = 0; i < n; ++i) {
                double sg = 0.0, sf = 0.0;
                for (int c = 0; c < np; ++c) { sg += Kg[i * np + c] * par[c]; sf += Ff[i * np + c] * par[c]; }
                if (g_P) g_P[(size_t)b * n + i] = (float)sg;                                     // acceldualgrad.m:21
                if (f) f[(size_t)b * n + i] = (float)sf;
            }
        if (p_D)
            for (int i = 0; i < m; ++i) {
                double s = p->b0[i];
                for (int c = 0; c < np; ++c) s += Bb[i * np + c] * par[c];
                p_D[(size_t)b * m + i] = (float)(-s * invL);                                     // acceldualgrad.m:23
            }
    }
    return GPAD_OK;
}

int gpad_closed_loop_plants(gpad_plants_t p, gpad_handle_t h, int first, int count, const double* x0, int samples,
                            const float* theta, const float* beta, int max_iter, int warm_start, double* x_traj,
                            double* u_traj) {
    if (!p || !h || !x0 || !theta || !beta || samples < 1 || max_iter < 1) return GPAD_ERR_INVALID_ARG;
    if (warm_start < GPAD_WARM_COLD || warm_start > GPAD_WARM_SHIFTED) return GPAD_ERR_INVALID_ARG;
    if (first <= 0 && count <= 0) { first = 0; count = p->B; }
    if (first < 0 || count < 1 || first + count > p->B) return GPAD_ERR_INVALID_ARG;
    int hn = 0, hN = 0, hm = 0, hmode = 0, hB = 0;
    if (gpad_handle_dims(h, &hn, &hN, &hm, &hmode, &hB, nullptr) != GPAD_OK) return GPAD_ERR_INVALID_ARG;
    if (hn != p->n_u || hN != p->N || hm != p->m || count > hB || hmode != GPAD_MODE_BATCH_PER_INSTANCE) return GPAD_ERR_INVALID_ARG;
    return gpad::closed_loop_plants_device(p, h, first, count, x0, samples, theta, beta, max_iter, warm_start, x_traj, u_traj);
}

}  // extern "C"
