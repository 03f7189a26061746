/*
 * gpad_oracle.h -- CPU oracle for the GPAD hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's sequential GPAD step functions
 * (Code/CUDA/FinalProject/src/seq_functions.cpp) composed in the order of the
 * reference driver loop (Code/CUDA/FinalProject/main.cu:160-175).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product library (libgpad_b200.so) never links or calls it.
 *
 * Parity pin: checked bit-for-bit against the reference's own compiled
 * seq_functions.cpp (oracle/_ref/libgpad_ref.so, built by oracle/Makefile) and
 * against the reference's five step-3 fixtures (FinalProject/build/step3/{1..5});
 * see tests/test_oracle.py and tests/golden/.
 *
 * Layout conventions (the *sequential* ones, seq_functions.cpp:61,82):
 *   M_G is [n][m] row-major, G_L is [m][n] row-major, n = n_u*N.
 */
#ifndef GPAD_ORACLE_H
#define GPAD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* status codes shared with include/gpad.h (GPAD_STATUS_*) */
enum {
    ORACLE_STATUS_MAX_ITER = 0,        /* ran all max_iter iterations              */
    ORACLE_STATUS_CONVERGED_Z = 1,     /* averaged iterate z_v feasible within eps_g*/
    ORACLE_STATUS_CONVERGED_ZHAT = 2,  /* zhat_v feasible and gap test passed (w>=0)*/
    ORACLE_STATUS_CONVERGED_DUAL = 3,  /* zhat_v feasible and V(zhat)-Phi(y) passed */
    ORACLE_STATUS_NONFINITE = 4
};

enum { ORACLE_SCHEDULE_PAPER = 0, ORACLE_SCHEDULE_MATLAB_LAG = 1 };

/* ---- single steps (seq_functions.cpp) ---------------------------------- */
void oracle_step_one(const float* y, const float* y_prev, float* w, float beta, int m);
void oracle_step_two(const float* M_G, const float* w, const float* g_P, float* zhat,
                     int N, int n_u, int m);
void oracle_step_three(float theta, int n, const float* z_prev, const float* zhat, float* z);
void oracle_step_four(const float* G_L, float* y_next, const float* w, const float* p_D,
                      const float* zhat, int N, int n_u, int m);
/* battery-structured ("flat") variants, operators stored [N][m] / [m][N] */
void oracle_step_two_flat(const float* M_G, const float* w, const float* g_P, float* zhat,
                          int N, int n_u, int m);
void oracle_step_four_flat(const float* G_L, float* y_next, const float* w, const float* p_D,
                           const float* zhat, int N, int n_u, int m);

/* ---- theta / beta schedule (acceldualgrad.m:55-56, paper eq. 8e) -------- */
void oracle_schedule(float* theta, float* beta, int count, int variant);

/* ---- whole solve -------------------------------------------------------- */
typedef struct {
    int n_u, N, m;
    const float* M_G;      /* [n][m] */
    const float* G_L;      /* [m][n] */
    const float* g_P;      /* [n]    */
    const float* p_D;      /* [m]    */
    const float* theta;    /* [max_iter] */
    const float* beta;     /* [max_iter] */
    int max_iter;
    const float* y0;       /* [m] or NULL (zeros): y_0      */
    const float* y_prev0;  /* [m] or NULL (zeros): y_{-1}   */
    const float* z0;       /* [n] or NULL (zeros): z_{-1}   */
    /* termination (row T of SURVEY section 8a); check_every <= 0 disables it */
    int check_every;
    float eps_g, eps_V, L;
    const float* f;        /* [n] or NULL: linear cost term, enables relative/dual gap tests */
} oracle_problem_t;

typedef struct {
    float* y_next;   /* [m]  y_{I}     (main.cu: y_vp1)  */
    float* y;        /* [m]  y_{I-1}   (main.cu: y_v)    */
    float* z;        /* [n]  z_{I-1}   (main.cu: z_v)    */
    float* zhat;     /* [n]  zhat_{I-1}                  */
    float* w;        /* [m]  w_{I-1}                     */
    int iters;       /* iterations executed               */
    int status;      /* ORACLE_STATUS_*                   */
    float max_viol;  /* max_i g(.)_i of the certified iterate at the last check (else NaN) */
    float gap;       /* duality-gap figure of the last check (else NaN)                    */
} oracle_result_t;

/* fp32, the reference's arithmetic: strict left-to-right sums */
int oracle_solve(const oracle_problem_t* p, oracle_result_t* r);
/* same loop with every operation in double (inputs promoted): the arbiter */
int oracle_solve_f64(const oracle_problem_t* p, double* y_next, double* y, double* z,
                     double* zhat, double* w, int* iters, int* status);

/* batch of independent QPs sharing operators; per-instance g_P [B][n], p_D [B][m]
 * (instance-major), optional per-instance y0/y_prev0 and f.  Outputs instance-major.
 * nthreads <= 0 -> all OpenMP threads. Returns threads used. */
int oracle_solve_batch(const oracle_problem_t* shared, int B,
                       const float* g_P, const float* p_D,
                       float* y_next, float* y, float* z, float* zhat, float* w,
                       int* iters, int* status, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
