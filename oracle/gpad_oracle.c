/*
 * gpad_oracle.c -- CPU oracle for the GPAD hot path.  TEST INFRASTRUCTURE ONLY
 * (see gpad_oracle.h for who may load it and how parity is pinned).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -pthread -fPIC -shared   (oracle/Makefile)
 * -ffp-contract=off keeps every multiply and add separately rounded, which is what
 * the reference's x86-64 "g++ -O2" build of seq_functions.cpp does (no FMA without
 * -march), so the restatement is bit-identical to it.
 */
#include "gpad_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* step 1: momentum extrapolation.  ref: seq_functions.cpp:45-51 (kernel_functions.cu:7-14) */
void oracle_step_one(const float* y, const float* y_prev, float* w, float beta, int m) {
    for (int i = 0; i < m; ++i) {
        float d = y[i] - y_prev[i];
        w[i] = y[i] + beta * d;
    }
}

/* step 2: primal recovery zhat = M_G w - g_P, M_G [n][m]; strict j = 0..m-1 fp32 sum.
 * ref: seq_functions.cpp:54-66 */
void oracle_step_two(const float* M_G, const float* w, const float* g_P, float* zhat,
                     int N, int n_u, int m) {
    const int n = n_u * N;
    for (int r = 0; r < n; ++r) {
        const float* row = M_G + (size_t)r * m;
        float acc = 0.0f;
        for (int c = 0; c < m; ++c) acc += row[c] * w[c];
        zhat[r] = acc - g_P[r];
    }
}

/* step 3: ergodic average z = (1-theta) z_prev + theta zhat.  ref: seq_functions.cpp:68-72.
 * (1 - theta) is int - float -> float, i.e. 1.0f - theta. */
void oracle_step_three(float theta, int n, const float* z_prev, const float* zhat, float* z) {
    const float one_minus = 1.0f - theta;
    for (int i = 0; i < n; ++i) z[i] = one_minus * z_prev[i] + theta * zhat[i];
}

/* step 4: y+ = max(G_L zhat + (w + p_D), 0), G_L [m][n]; the (w+p_D) pair is rounded
 * once before it joins the sum, the projection is (s+|s|)/2.  ref: seq_functions.cpp:75-87 */
void oracle_step_four(const float* G_L, float* y_next, const float* w, const float* p_D,
                      const float* zhat, int N, int n_u, int m) {
    const int n = n_u * N;
    for (int r = 0; r < m; ++r) {
        const float* row = G_L + (size_t)r * n;
        float acc = 0.0f;
        for (int c = 0; c < n; ++c) acc += row[c] * zhat[c];
        acc += w[r] + p_D[r];
        y_next[r] = (acc + fabsf(acc)) / 2;
    }
}

/* flat step 2: M_G stored [N][m]; stage i, input j gathers the stride-n_u columns of the
 * first 4 n_u N constraint rows plus all the sum-constraint columns.
 * ref: seq_functions.cpp:5-20 */
void oracle_step_two_flat(const float* M_G, const float* w, const float* g_P, float* zhat,
                          int N, int n_u, int m) {
    const int box = 4 * n_u * N;
    for (int s = 0; s < N; ++s) {
        const float* row = M_G + (size_t)s * m;
        for (int u = 0; u < n_u; ++u) {
            float acc = 0.0f;
            for (int c = u; c < box; c += n_u) acc += row[c] * w[c];
            for (int c = box; c < m; ++c) acc += row[c] * w[c];
            zhat[s * n_u + u] = acc - g_P[s * n_u + u];
        }
    }
}

/* flat step 4: G_L stored [m][N]; thresholded projection (y<0 -> 0).
 * ref: seq_functions.cpp:23-43 */
void oracle_step_four_flat(const float* G_L, float* y_next, const float* w, const float* p_D,
                           const float* zhat, int N, int n_u, int m) {
    const int box = 4 * n_u * N;
    for (int r = 0; r < m; ++r) {
        const float* row = G_L + (size_t)r * N;
        float acc = 0.0f;
        for (int s = 0; s < N; ++s) {
            if (r < box) {
                acc += row[s] * zhat[s * n_u + (r % n_u)];
            } else {
                for (int u = 0; u < n_u; ++u) acc += row[s] * zhat[s * n_u + u];
            }
        }
        y_next[r] = acc + w[r] + p_D[r];
    }
    for (int r = 0; r < m; ++r)
        if (y_next[r] < 0) y_next[r] = 0;
}

/* theta_{v+1} = (sqrt(theta^4 + 4 theta^2) - theta^2)/2, theta_0 = theta_{-1} = 1
 * (acceldualgrad.m:17,55; paper eq. 8e).  Evaluated in double like MATLAB, stored fp32.
 * PAPER:      beta_v = theta_v (1/theta_{v-1} - 1)
 * MATLAB_LAG: acceldualgrad.m:56 evaluates beta before the register shift (:60-64), so
 *             iteration v uses theta_{v-1} (1/theta_{v-2} - 1), i.e. the paper's beta_{v-1}. */
void oracle_schedule(float* theta, float* beta, int count, int variant) {
    double th_prev = 1.0, th = 1.0; /* theta_{v-1}, theta_v */
    double lagged = 0.0;            /* acceldualgrad.m:27 beta_v = 0 */
    for (int v = 0; v < count; ++v) {
        const double paper_beta = th * (1.0 / th_prev - 1.0);
        theta[v] = (float)th;
        beta[v] = (float)(variant == ORACLE_SCHEDULE_MATLAB_LAG ? lagged : paper_beta);
        lagged = paper_beta;
        const double t2 = th * th;
        const double th_next = (sqrt(t2 * t2 + 4.0 * t2) - t2) / 2.0;
        th_prev = th;
        th = th_next;
    }
}

static int all_finite_f(const float* v, int len) {
    for (int i = 0; i < len; ++i)
        if (!isfinite(v[i])) return 0;
    return 1;
}

/*
 * Whole solve: the reference has no sequential driver; this composes the step functions in
 * the order of the GPU loop main.cu:160-175 (1 -> 2 -> y_prev<-y -> 3 -> 4) from the zero
 * start of main.cu:69-77, and adds the termination test of acceldualgrad.m:66-79 / paper
 * Alg. 1, which the reference never ported to C.
 *
 * Termination, evaluated after step 4 of iteration v when (v+1) % check_every == 0, with
 *   dot_i  = (G_L zhat_v)_i                       (step-4 sum before the (w+p_D) term)
 *   rhat_i = dot_i + p_D_i      = g(zhat_v)_i / L
 *   sbar_i = (1-theta_v) sbar_i + theta_v rhat_i  = g(z_v)_i / L   (g affine, z_v averaged)
 *   1. L max_i sbar_i <= eps_g                                  -> CONVERGED_Z
 *   2. else if L max_i rhat_i <= eps_g:
 *        if min_i w_i >= 0:  gap = -L sum_i w_i rhat_i  (= -w'g(zhat))
 *             gap <= eps_V, or (f given) gap <= V(zhat) eps_V/(1+eps_V) -> CONVERGED_ZHAT
 *        else (f given): V(zhat) - Phi(y_{v+1}) <= eps_V max(Phi,1)      -> CONVERGED_DUAL
 *   V(zhat) = (f'zhat - L sum_i w_i dot_i)/2     (uses H zhat = -(G'w + f))
 *   Phi(y)  = f'z_y/2 + (L/2) sum_i y_i (G_L z_y)_i + L sum_i y_i p_D_i,  z_y = M_G y - g_P
 * Reductions (max, dot) are accumulated in double here; they are not reference arithmetic.
 */
int oracle_solve(const oracle_problem_t* p, oracle_result_t* r) {
    const int n = p->n_u * p->N, m = p->m;
    float* y = r->y_next; /* holds y_v on entry to an iteration, y_{v+1} after step 4 */
    float* y_prev = r->y;
    float* z = r->z;
    float* zhat = r->zhat;
    float* w = r->w;
    float* sbar = (float*)calloc((size_t)m, sizeof(float));
    float* dot = (float*)calloc((size_t)m, sizeof(float));
    float* scratch_z = NULL;
    if (!sbar || !dot) { free(sbar); free(dot); return -1; }

    if (p->y0) memcpy(y, p->y0, sizeof(float) * m); else memset(y, 0, sizeof(float) * m);
    if (p->y_prev0) memcpy(y_prev, p->y_prev0, sizeof(float) * m); else memset(y_prev, 0, sizeof(float) * m);
    if (p->z0) memcpy(z, p->z0, sizeof(float) * n); else memset(z, 0, sizeof(float) * n);
    memset(zhat, 0, sizeof(float) * n);
    memset(w, 0, sizeof(float) * m);

    r->iters = 0;
    r->status = ORACLE_STATUS_MAX_ITER;
    r->max_viol = NAN;
    r->gap = NAN;
    const int checking = p->check_every > 0;

    for (int v = 0; v < p->max_iter; ++v) {
        const float th = p->theta[v];
        oracle_step_one(y, y_prev, w, p->beta[v], m);                 /* main.cu:163 */
        oracle_step_two(p->M_G, w, p->g_P, zhat, p->N, p->n_u, m);     /* main.cu:166 */
        memcpy(y_prev, y, sizeof(float) * m);                          /* main.cu:167 */
        oracle_step_three(th, n, z, zhat, z);                          /* main.cu:170 */
        if (checking) {
            /* same row sums step 4 forms, kept before the (w+p_D) term joins them */
            for (int i = 0; i < m; ++i) {
                const float* row = p->G_L + (size_t)i * n;
                float acc = 0.0f;
                for (int c = 0; c < n; ++c) acc += row[c] * zhat[c];
                dot[i] = acc;
            }
        }
        oracle_step_four(p->G_L, y, w, p->p_D, zhat, p->N, p->n_u, m); /* main.cu:171 */
        r->iters = v + 1;

        if (!checking) continue;
        const float one_minus = 1.0f - th;
        for (int i = 0; i < m; ++i) {
            const float rhat = dot[i] + p->p_D[i];
            sbar[i] = one_minus * sbar[i] + th * rhat;
        }
        if ((v + 1) % p->check_every != 0) continue;

        if (!all_finite_f(y, m)) { r->status = ORACLE_STATUS_NONFINITE; break; }
        double max_sbar = -INFINITY, max_rhat = -INFINITY, min_w = INFINITY;
        double w_rhat = 0.0, w_dot = 0.0;
        for (int i = 0; i < m; ++i) {
            const float rhat = dot[i] + p->p_D[i];
            if (sbar[i] > max_sbar) max_sbar = sbar[i];
            if (rhat > max_rhat) max_rhat = rhat;
            if (w[i] < min_w) min_w = w[i];
            w_rhat += (double)w[i] * rhat;
            w_dot += (double)w[i] * dot[i];
        }
        const float viol_z = (float)(p->L * max_sbar);
        const float viol_zhat = (float)(p->L * max_rhat);
        r->max_viol = viol_z;
        if (viol_z <= p->eps_g) { r->status = ORACLE_STATUS_CONVERGED_Z; break; }
        if (viol_zhat > p->eps_g) continue;

        double f_zhat = 0.0;
        if (p->f) for (int j = 0; j < n; ++j) f_zhat += (double)p->f[j] * zhat[j];
        const double V = 0.5 * (f_zhat - p->L * w_dot);
        if (min_w >= 0.0) {
            const float gap = (float)(-p->L * w_rhat);
            r->gap = gap;
            if (gap <= p->eps_V || (p->f && gap <= (float)(V * p->eps_V / (1.0 + p->eps_V)))) {
                r->status = ORACLE_STATUS_CONVERGED_ZHAT;
                r->max_viol = viol_zhat;
                break;
            }
        } else if (p->f) {
            if (!scratch_z) scratch_z = (float*)malloc(sizeof(float) * n);
            oracle_step_two(p->M_G, y, p->g_P, scratch_z, p->N, p->n_u, m); /* z_y */
            double f_zy = 0.0, y_Gzy = 0.0, y_pD = 0.0;
            for (int j = 0; j < n; ++j) f_zy += (double)p->f[j] * scratch_z[j];
            for (int i = 0; i < m; ++i) {
                const float* row = p->G_L + (size_t)i * n;
                float acc = 0.0f;
                for (int c = 0; c < n; ++c) acc += row[c] * scratch_z[c];
                y_Gzy += (double)y[i] * acc;
                y_pD += (double)y[i] * p->p_D[i];
            }
            const double Phi = 0.5 * f_zy + 0.5 * p->L * y_Gzy + p->L * y_pD;
            const float gap = (float)(V - Phi);
            r->gap = gap;
            if (gap <= (float)(p->eps_V * (Phi > 1.0 ? Phi : 1.0))) {
                r->status = ORACLE_STATUS_CONVERGED_DUAL;
                r->max_viol = viol_zhat;
                break;
            }
        }
    }
    if (r->status == ORACLE_STATUS_MAX_ITER && !all_finite_f(y, m)) r->status = ORACLE_STATUS_NONFINITE;
    free(sbar); free(dot); free(scratch_z);
    return 0;
}

/* the arbiter: identical loop, every operation in double; termination identical in form */
int oracle_solve_f64(const oracle_problem_t* p, double* y, double* y_prev, double* z,
                     double* zhat, double* w, int* iters, int* status) {
    const int n = p->n_u * p->N, m = p->m;
    double* sbar = (double*)calloc((size_t)m, sizeof(double));
    double* dot = (double*)calloc((size_t)m, sizeof(double));
    double* zy = (double*)calloc((size_t)n, sizeof(double));
    if (!sbar || !dot || !zy) { free(sbar); free(dot); free(zy); return -1; }
    for (int i = 0; i < m; ++i) { y[i] = p->y0 ? p->y0[i] : 0.0; y_prev[i] = p->y_prev0 ? p->y_prev0[i] : 0.0; w[i] = 0.0; }
    for (int j = 0; j < n; ++j) { z[j] = p->z0 ? p->z0[j] : 0.0; zhat[j] = 0.0; }
    *iters = 0;
    *status = ORACLE_STATUS_MAX_ITER;
    const int checking = p->check_every > 0;
    const double L = p->L;
    for (int v = 0; v < p->max_iter; ++v) {
        const double th = p->theta[v], be = p->beta[v];
        for (int i = 0; i < m; ++i) w[i] = y[i] + be * (y[i] - y_prev[i]);
        for (int j = 0; j < n; ++j) {
            const float* row = p->M_G + (size_t)j * m;
            double acc = 0.0;
            for (int c = 0; c < m; ++c) acc += (double)row[c] * w[c];
            zhat[j] = acc - (double)p->g_P[j];
        }
        memcpy(y_prev, y, sizeof(double) * m);
        for (int j = 0; j < n; ++j) z[j] = (1.0 - th) * z[j] + th * zhat[j];
        for (int i = 0; i < m; ++i) {
            const float* row = p->G_L + (size_t)i * n;
            double acc = 0.0;
            for (int c = 0; c < n; ++c) acc += (double)row[c] * zhat[c];
            dot[i] = acc;
            const double s = acc + (w[i] + (double)p->p_D[i]);
            y[i] = s > 0.0 ? s : 0.0;
        }
        *iters = v + 1;
        if (!checking) continue;
        for (int i = 0; i < m; ++i) sbar[i] = (1.0 - th) * sbar[i] + th * (dot[i] + (double)p->p_D[i]);
        if ((v + 1) % p->check_every != 0) continue;
        double max_sbar = -INFINITY, max_rhat = -INFINITY, min_w = INFINITY, w_rhat = 0.0, w_dot = 0.0;
        for (int i = 0; i < m; ++i) {
            const double rhat = dot[i] + (double)p->p_D[i];
            if (sbar[i] > max_sbar) max_sbar = sbar[i];
            if (rhat > max_rhat) max_rhat = rhat;
            if (w[i] < min_w) min_w = w[i];
            w_rhat += w[i] * rhat;
            w_dot += w[i] * dot[i];
        }
        if (L * max_sbar <= p->eps_g) { *status = ORACLE_STATUS_CONVERGED_Z; break; }
        if (L * max_rhat > p->eps_g) continue;
        double f_zhat = 0.0;
        if (p->f) for (int j = 0; j < n; ++j) f_zhat += (double)p->f[j] * zhat[j];
        const double V = 0.5 * (f_zhat - L * w_dot);
        if (min_w >= 0.0) {
            const double gap = -L * w_rhat;
            if (gap <= p->eps_V || (p->f && gap <= V * p->eps_V / (1.0 + p->eps_V))) {
                *status = ORACLE_STATUS_CONVERGED_ZHAT; break;
            }
        } else if (p->f) {
            double f_zy = 0.0, y_Gzy = 0.0, y_pD = 0.0;
            for (int j = 0; j < n; ++j) {
                const float* row = p->M_G + (size_t)j * m;
                double acc = 0.0;
                for (int c = 0; c < m; ++c) acc += (double)row[c] * y[c];
                zy[j] = acc - (double)p->g_P[j];
                f_zy += (double)p->f[j] * zy[j];
            }
            for (int i = 0; i < m; ++i) {
                const float* row = p->G_L + (size_t)i * n;
                double acc = 0.0;
                for (int c = 0; c < n; ++c) acc += (double)row[c] * zy[c];
                y_Gzy += y[i] * acc;
                y_pD += y[i] * (double)p->p_D[i];
            }
            const double Phi = 0.5 * f_zy + 0.5 * L * y_Gzy + L * y_pD;
            if (V - Phi <= p->eps_V * (Phi > 1.0 ? Phi : 1.0)) { *status = ORACLE_STATUS_CONVERGED_DUAL; break; }
        }
    }
    free(sbar); free(dot); free(zy);
    return 0;
}

/* ---- batch driver: independent QPs over pthread workers pulling from a shared counter ---- */
typedef struct {
    const oracle_problem_t* shared;
    int B;
    const float *g_P, *p_D;
    float *y_next, *y, *z, *zhat, *w;
    int *iters, *status;
    int next; /* atomic work counter */
} batch_job_t;

static void* batch_worker(void* arg) {
    batch_job_t* j = (batch_job_t*)arg;
    const oracle_problem_t* shared = j->shared;
    const int n = shared->n_u * shared->N, m = shared->m;
    for (;;) {
        const int b = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (b >= j->B) break;
        oracle_problem_t p = *shared;
        p.g_P = j->g_P + (size_t)b * n;
        p.p_D = j->p_D + (size_t)b * m;
        if (shared->y0) p.y0 = shared->y0 + (size_t)b * m;
        if (shared->y_prev0) p.y_prev0 = shared->y_prev0 + (size_t)b * m;
        if (shared->z0) p.z0 = shared->z0 + (size_t)b * n;
        if (shared->f) p.f = shared->f + (size_t)b * n;
        oracle_result_t r;
        r.y_next = j->y_next + (size_t)b * m;
        r.y = j->y + (size_t)b * m;
        r.z = j->z + (size_t)b * n;
        r.zhat = j->zhat + (size_t)b * n;
        r.w = j->w + (size_t)b * m;
        oracle_solve(&p, &r);
        if (j->iters) j->iters[b] = r.iters;
        if (j->status) j->status[b] = r.status;
    }
    return NULL;
}

int oracle_solve_batch(const oracle_problem_t* shared, int B,
                       const float* g_P, const float* p_D,
                       float* y_next, float* y, float* z, float* zhat, float* w,
                       int* iters, int* status, int nthreads) {
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads > B) nthreads = B;
    if (nthreads < 1) nthreads = 1;
    batch_job_t job = {shared, B, g_P, p_D, y_next, y, z, zhat, w, iters, status, 0};
    pthread_t* tids = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    int started = 0;
    for (int t = 1; t < nthreads; ++t)
        if (pthread_create(&tids[started], NULL, batch_worker, &job) == 0) ++started;
    batch_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(tids[t], NULL);
    free(tids);
    return started + 1;
}
