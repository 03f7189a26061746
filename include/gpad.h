/*
 * gpad.h -- C ABI of libgpad_b200.so, the B200-native (sm_100a) GPAD solver.
 *
 * This is the drop-in boundary for the GPAD hot path of shreyasren/GPU-DualGradient-MPC.
 * The reference has no plugin/FFI layer: its boundary is "the symbols main.cu links against"
 * (the five __global__ kernels of Code/CUDA/FinalProject/include/kernel_functions.h:9-39,
 * launched by the loop at main.cu:160-175 on caller-owned device buffers) plus the text data
 * file main.cu:29-67 reads.  Every entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - plain C types only; every function returns a gpad_status (0 = ok) and never exits.
 *   - n = n_u*N decision variables, m constraints (file header "n_u N m num_iterations L").
 *   - operators follow the C/CUDA sign convention of the reference: zhat = M_G w - g_P
 *     (kernel_functions.cu:62, seq_functions.cpp:63), y+ = max(G_L zhat + (w + p_D), 0).
 *   - batched vectors are instance-major: v[b*len + i] (a batch is B reference problems'
 *     vectors back to back).
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point
 *     returns GPAD_ERR_NO_DEVICE / GPAD_ERR_CUDA.
 *   - a handle is not thread-safe and has ONE solve in flight: its scratch (schedule tables, batch
 *     state) is shared by all its solves, so a solve enqueued on one stream is ordered by the library
 *     after the handle's previous solve on any other stream (gpad_solve_async is the one pipelined
 *     exception and owns its ordering).  Distinct handles are independent.  The caller owns every
 *     buffer it passes; the library owns its handle and its converted operator copies.
 *   - every call selects the handle's device for its duration and restores the caller's current device.
 */
#ifndef GPAD_H
#define GPAD_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPAD_API_VERSION 2

typedef enum {
    GPAD_OK = 0,
    GPAD_ERR_INVALID_ARG = 1,
    GPAD_ERR_CUDA = 2,
    GPAD_ERR_UNSUPPORTED = 3,
    GPAD_ERR_ALLOC = 4,
    GPAD_ERR_NO_DEVICE = 5,
    GPAD_ERR_IO = 6
} gpad_status;

/* operator storage of the pointers given to gpad_setup / the step shims */
typedef enum {
    GPAD_LAYOUT_FLIPPED = 0,    /* M_G [m][n], G_L [n][m]: what the reference kernels read
                                   (ENABLE_FLIPPING, kernel_functions.cu:50,180)            */
    GPAD_LAYOUT_SEQUENTIAL = 1, /* M_G [n][m], G_L [m][n]: what seq_functions.cpp:61,82 read */
    GPAD_LAYOUT_FLAT = 2        /* battery-structured "flattened" operators M_G [N][m], G_L [m][N]
                                   (ENABLE_FLATTEN_MATRICES, main.cu:39-41,50-52; seq_functions.cpp:5-43,
                                   kernel_functions.cu:74-109); expanded to dense at gpad_setup          */
} gpad_layout;

typedef enum {
    GPAD_MODE_LATENCY = 1,            /* one QP, persistent kernel, operators resident on chip */
    GPAD_MODE_BATCH_SHARED = 2,       /* B QPs sharing M_G/G_L: per-iteration GEMMs            */
    GPAD_MODE_BATCH_PER_INSTANCE = 3  /* B QPs each with its own M_G/G_L: batched GEMV         */
} gpad_mode;

typedef enum {
    GPAD_PREC_FP32 = 0,    /* CUDA-core FFMA, fp32 accumulate                                   */
    GPAD_PREC_TF32X3 = 1,  /* tcgen05 kind::tf32, hi/lo split of both operands (3 MMAs/product),
                              fp32 accumulate in TMEM -- BATCH_SHARED only                      */
    GPAD_PREC_FP16X3 = 2   /* tcgen05 kind::f16: the same 11 + 11 bit hi/lo split held in fp16, every operand
                              row scaled by an exact power of two taken from its largest magnitude (scales
                              undone on the fp32 accumulator): half the operand bytes and MMA instructions
                              of TF32X3 at the same accuracy.  Fixed-iteration solves; tolerance-mode solves
                              of such a handle run the TF32X3 kernels -- BATCH_SHARED only          */
} gpad_precision;

typedef enum { GPAD_MEM_HOST = 0, GPAD_MEM_DEVICE = 1 } gpad_memspace;

/* per-instance termination status (iters/status/max_viol/gap are the termination outputs the
 * reference never had; SURVEY section 8a row T, acceldualgrad.m:66-79, paper Alg. 1) */
typedef enum {
    GPAD_STATUS_MAX_ITER = 0,
    GPAD_STATUS_CONVERGED_Z = 1,
    GPAD_STATUS_CONVERGED_ZHAT = 2,
    GPAD_STATUS_CONVERGED_DUAL = 3,
    GPAD_STATUS_NONFINITE = 4
} gpad_solve_status;

typedef enum { GPAD_SCHEDULE_PAPER = 0, GPAD_SCHEDULE_MATLAB_LAG = 1 } gpad_schedule_variant;

const char* gpad_status_string(int status);
/* message of the last failure on the calling thread (CUDA error text included) */
const char* gpad_last_error(void);
int gpad_api_version(void);
/* number of usable sm_100 devices (0 without a GPU; never fails) */
int gpad_device_count(void);

/* ------------------------------------------------------------------------------------------
 * 1. Step-compatible shims: host-callable replacements of the five kernel launches of the
 *    reference loop body, same argument lists plus a stream (the reference uses the default
 *    stream and cudaDeviceSynchronize, main.cu:164,168,172; here nothing synchronises).
 *    All pointers are DEVICE pointers; operators are in GPAD_LAYOUT_FLIPPED like the
 *    reference's data files.  stream is a cudaStream_t (NULL = default stream).
 * ---------------------------------------------------------------------------------------- */
/* replaces StepOneGPADKernel<<<>>>            kernel_functions.h:9,  main.cu:163 */
int gpad_step_one(const float* y_vec_in, const float* y_vec_minus_1_in, float* w_vec_out,
                  float beta_v, int m, void* stream);
/* replaces StepTwoGPADKernel<<<>>>            kernel_functions.h:10-20, main.cu:166 */
int gpad_step_two(const float* M_G, const float* w_v, const float* g_P, float* zhat,
                  int N, int n_u, int m, void* stream);
/* replaces DeviceArrayCopy<<<>>>              kernel_functions.h:39, main.cu:167 */
int gpad_array_copy(float* dest, const float* src, int size, void* stream);
/* replaces StepThreeGPADKernel<<<>>>          kernel_functions.h:21, main.cu:170 */
int gpad_step_three(float theta, const float* zhat_v, float* z_v, int length, void* stream);
/* replaces StepFourGPADFlippedParRows<<<>>>   kernel_functions.h:26-38, main.cu:171
 * (max_threads is accepted and ignored, as in the reference kernel) */
int gpad_step_four(const float* G_L, float* y_vp1, const float* w_v, const float* p_D,
                   const float* zhat_v, int N, int n_u, int m, int max_threads, void* stream);

/* ------------------------------------------------------------------------------------------
 * 2. Whole-solve entry points: replace the body of main() between readData and the D2H copies
 *    (main.cu:108-180): allocation, H2D, the iteration loop, D2H.
 * ---------------------------------------------------------------------------------------- */
typedef struct gpad_handle_s* gpad_handle_t;
typedef struct gpad_problem_s* gpad_problem_t;      /* host-side condensed problem, section 3 */

typedef struct {
    int n_u, N, m;          /* file header, main.cu:34                                        */
    float L;                /* Lipschitz constant from the file header (used by termination)  */
    int layout;             /* gpad_layout of M_G / G_L passed to gpad_setup                  */
    int mode;               /* gpad_mode                                                       */
    int precision;          /* gpad_precision                                                  */
    int max_batch;          /* capacity in instances (1 for GPAD_MODE_LATENCY)                 */
    int device;             /* CUDA device ordinal, -1 = current device                        */
    int operators_mem;      /* gpad_memspace of M_G / G_L                                      */
    int reserved[6];        /* must be zero                                                    */
} gpad_config_t;

/* Uploads / converts the operators once (the cudaMalloc + cudaMemcpy H2D block main.cu:126-147).
 * GPAD_MODE_BATCH_PER_INSTANCE: M_G / G_L hold max_batch operators back to back. */
int gpad_setup(const gpad_config_t* cfg, const float* M_G, const float* G_L, gpad_handle_t* out);
int gpad_destroy(gpad_handle_t h);

typedef struct {
    int batch;              /* instances in this call, 1..max_batch                            */
    int mem;                /* gpad_memspace of every in/out pointer below except theta/beta   */
    /* inputs */
    const float* g_P;       /* [batch][n]                                                      */
    const float* p_D;       /* [batch][m]                                                      */
    const float* f;         /* [batch][n] or NULL; enables the relative / dual gap tests       */
    const float* y0;        /* [batch][m] or NULL (zeros, main.cu:69-77): y_0   (warm start)   */
    const float* y_prev0;   /* [batch][m] or NULL (zeros): y_{-1}                              */
    const float* theta;     /* HOST [max_iter]  (passed by value per launch in main.cu:170)    */
    const float* beta;      /* HOST [max_iter]  (main.cu:163)                                  */
    int max_iter;           /* N_v = 100 in main.cu:87                                         */
    int check_every;        /* <= 0: fixed iteration count (the reference's behaviour)         */
    float eps_g, eps_V;     /* acceldualgrad.m:12-13                                           */
    /* outputs, any may be NULL: the five vectors main.cu:176-180 copies back ...              */
    float* y_next;          /* [batch][m]  y_I       (main.cu: y_vp1)                          */
    float* y;               /* [batch][m]  y_{I-1}   (main.cu: y_v)                            */
    float* z;               /* [batch][n]  z_{I-1}   (main.cu: z_v)                            */
    float* zhat;            /* [batch][n]  zhat_{I-1}                                          */
    float* w;               /* [batch][m]  w_{I-1}                                             */
    /* ... and the termination outputs */
    int* iters;             /* [batch] iterations executed                                     */
    int* status;            /* [batch] gpad_solve_status                                       */
    float* max_viol;        /* [batch] max_i g(.)_i at the last check (NaN if never checked)   */
    float* gap;             /* [batch] duality-gap figure at the last check (NaN if none)      */
    void* stream;           /* cudaStream_t for GPAD_MEM_DEVICE calls (NULL = default stream)  */
    /* on-device instance build (GPAD_MODE_BATCH_SHARED): when params != NULL, g_P / p_D above are ignored
     * and built on the device from the parameters, g_P = H^-1 F' p, p_D = -(b0 + Bb p) / L
     * (acceldualgrad.m:21,23; gpad.m:81,85): 8 n_par bytes per instance cross PCIe instead of 4 (n + m) */
    const double* params;   /* [batch][n_par] (memspace `mem`) or NULL                         */
    gpad_problem_t problem; /* the problem the parameters belong to (same n_u, N, m as the handle) */
    int build_f;            /* with params: also build f = F' p and enable the relative / dual gap tests */
    int reserved[3];        /* must be zero                                                    */
} gpad_solve_args_t;

/* GPAD_MEM_HOST: copies in, solves, copies out and synchronises before returning.  In GPAD_MODE_BATCH_SHARED a
 * fixed-iteration solve of >= 16384 instances on the tensor-core precisions runs as two halves over the
 * double-buffered path of gpad_solve_async (the second half's inputs arrive and the first half's results leave under
 * the other half's iterations); results are those of the single solve bit for bit, and the handle holds a second set of
 * batch state from the first such call on.
 * GPAD_MEM_DEVICE: enqueues everything on args->stream and returns without synchronising
 * (in tolerance mode the host follows the device's stop decisions a few checks behind; stopped
 * instances are frozen, so the extra iterations it may enqueue change nothing). */
int gpad_solve(gpad_handle_t h, const gpad_solve_args_t* args);

/* Asynchronous host-memory solves (GPAD_MODE_BATCH_SHARED, fixed iteration count): enqueues copy-in, iterations and
 * copy-out on three streams over two alternating sets of batch state and returns a ticket at once, so that with
 * back-to-back calls the H2D copies of solve k+1 and the D2H copies of solve k-1 run under the iterations of solve k
 * (main.cu:136-147 / 176-180 are serial with its loop).  Host buffers should be page-locked (cudaHostAlloc /
 * cudaHostRegister); pageable ones work but their copies serialise.  A solve's buffers belong to the library until
 * gpad_wait(ticket) returns; at most two tickets are in flight (a third call waits on the device for the first). */
int gpad_solve_async(gpad_handle_t h, const gpad_solve_args_t* args, long long* ticket);
int gpad_wait(gpad_handle_t h, long long ticket);

/* dimensions a handle was set up with (any pointer may be NULL) */
int gpad_handle_dims(gpad_handle_t h, int* n_u, int* N, int* m, int* mode, int* max_batch, int* device);

/* tolerance-mode bookkeeping of the handle's last GPAD_MODE_BATCH_SHARED solve (synchronises with it):
 * instance-iterations the GEMM kernels were scheduled for (128 x iterations per batch tile that still held a running
 * instance) and instance-iterations the instances needed (sum of iters); their ratio is the share of tensor work that
 * was useful, the rest rode along on stopped instances of partly finished tiles.  Instances stop at very different
 * iterations, so the library retires finished tiles and, whenever a fifth of the working rows has stopped, gathers
 * the running instances into dense tiles (results do not depend on it). */
typedef struct {
    double instance_iterations_scheduled;
    double instance_iterations_needed;
    int compactions;            /* times the running instances were gathered into dense tiles */
    int reserved;
} gpad_solve_stats_t;
int gpad_solve_stats(gpad_handle_t h, gpad_solve_stats_t* out);

/* Optional per-kernel device timing: when enabled, every hot-path kernel launch of this handle is
 * bracketed by CUDA events on the launching stream (bench.py's roofline figure).
 * which: 0 = latency persistent kernel (GPAD_PREC_FP16X3 batch handles: the zhat row-quantisation kernel between
 * the products), 1 = product-1 kernel, 2 = product-2 kernel.
 * gpad_profile_read synchronises, returns the accumulated milliseconds and launch count since the
 * last read, and resets them. */
int gpad_profile_enable(gpad_handle_t h, int enable);
int gpad_profile_read(gpad_handle_t h, int which, double* total_ms, long long* launches);

/* kernels launched by this handle since setup (bench.py's gpu_launches claim) */
long long gpad_launch_count(gpad_handle_t h);
/* human-readable description of the kernel path chosen for this handle */
const char* gpad_describe(gpad_handle_t h);

/* ------------------------------------------------------------------------------------------
 * 3. Host-side problem setup (C++ restatement of the MATLAB offline stage: gpad.m:4-85,
 *    acceldualgrad.m:9-23) and the theta/beta schedule the data file carries (main.cu:61-64).
 * ---------------------------------------------------------------------------------------- */
/* battery balancing, n_u cells, horizon N: m = 4 n_u N + 2 N, L = ||H||_F^2 */
int gpad_problem_battery(int n_u, int N, gpad_problem_t* out);
/* hover-linearised quadrotor, nx = 12, nu = 4, horizon N: m = 24 N, L = 1.02 lambda_max(G H^-1 G') */
int gpad_problem_quadrotor(int N, gpad_problem_t* out);
int gpad_problem_destroy(gpad_problem_t p);
/* n_par = length of the per-instance parameter vector (battery: x0 [n_u]; quadrotor: [x0;xref] [24]) */
int gpad_problem_dims(gpad_problem_t p, int* n_u, int* N, int* m, int* n_par, float* L);
/* Lipschitz constant of the dual gradient, the L of G_L = G / L and p_D = -b / L: GPAD_L_REFERENCE = ||H||_F^2 with H the
 * PRIMAL Hessian (acceldualgrad.m:11: the reference's choice, the battery default; only a valid bound by accident of the
 * battery scaling, 1.5x ... 50x above the dual Hessian's largest eigenvalue), GPAD_L_LAMBDA_MAX = 1.02 lambda_max(G H^-1 G')
 * (paper section 4; the quadrotor default).  Changes what gpad_problem_dims / _operators / _instances return and what the
 * on-device instance build uses from then on; handles built from the earlier operators keep the earlier L. */
enum { GPAD_L_REFERENCE = 0, GPAD_L_LAMBDA_MAX = 1 };
int gpad_problem_set_lipschitz(gpad_problem_t p, int which, float* L_out);
/* operators to host buffers of n*m floats each, in the requested gpad_layout */
int gpad_problem_operators(gpad_problem_t p, int layout, float* M_G, float* G_L);
/* per-instance vectors from parameters [B][n_par] (double): g_P [B][n], p_D [B][m], f [B][n]
 * (f may be NULL) -- all host buffers */
int gpad_problem_instances(gpad_problem_t p, int B, const double* params, float* g_P, float* p_D,
                           float* f);
/* the same maps evaluated on the CURRENT device: params [B][n_par] (double), g_P [B][n], p_D [B][m], f [B][n] or NULL,
 * all device pointers, dense rows; enqueued on `stream`.  The problem's matrices are uploaded once per device. */
int gpad_instances_device(gpad_problem_t p, int B, const double* params, float* g_P, float* p_D, float* f, void* stream);
/* plant matrices for closed-loop simulation: A [nx][nx], B [nx][n_u] row-major (double) */
int gpad_problem_plant(gpad_problem_t p, int* nx, double* A, double* B);

/* Flattened (battery-structured) operators <-> dense sequential operators.  The flat form exists when
 * M_G[(s n_u + u)][k] is zero unless k >= 4 n_u N or k % n_u == u (and likewise for G_L), i.e. identical
 * cells (Cookbook 2.2).  gpad_flatten_operators returns in *max_residual the largest |entry| that the flat
 * form cannot represent (0 for the reference's battery problem). */
int gpad_flatten_operators(int n_u, int N, int m, const float* M_G_seq, const float* G_L_seq, float* M_G_flat,
                           float* G_L_flat, float* max_residual);
int gpad_expand_operators(int n_u, int N, int m, const float* M_G_flat, const float* G_L_flat, float* M_G_seq,
                          float* G_L_seq);

/* Closed-loop receding-horizon simulation (gpad.m:79-95): every sample builds g_P / p_D from the current
 * states, solves the batch with h (max_iter iterations, fixed), applies u = z[0:n_u] and advances
 * x <- A x + B u in double.  warm_start: 0 = cold start every sample like the reference (acceldualgrad.m:16-18),
 * GPAD_WARM_PREVIOUS = the previous duals (y_I, y_{I-1}) start the next solve, GPAD_WARM_SHIFTED = the previous duals
 * moved one stage forward inside every constraint block (the last stage repeats), the receding-horizon shift.
 * The whole loop runs on the handle's device (instance build, solve, state advance; nx <= 32); only the trajectories
 * are copied back.  x0 [batch][nx]; xref [batch][n_par-nx] or NULL; x_traj [samples+1][batch][nx];
 * u_traj [samples][batch][n_u] (host, either may be NULL).  The handle must have been set up for the problem's
 * n_u, N, m in a batch mode with max_batch >= batch (GPAD_ERR_INVALID_ARG otherwise). */
enum { GPAD_WARM_COLD = 0, GPAD_WARM_PREVIOUS = 1, GPAD_WARM_SHIFTED = 2 };
int gpad_closed_loop(gpad_problem_t prob, gpad_handle_t h, int batch, const double* x0, const double* xref, int samples,
                     const float* theta, const float* beta, int max_iter, int warm_start, double* x_traj, double* u_traj);

/* ------------------------------------------------------------------------------------------
 * 3b. Per-instance plants (BASELINE config 5): B battery packs whose cell capacities differ per instance
 *     (gpad.m:18 scaled by capacity_scale [B][n_u]), hence B different M_G / G_L / L and affine instance maps.
 *     Condensed on `threads` host threads (<= 0: all cores).
 * ---------------------------------------------------------------------------------------- */
typedef struct gpad_plants_s* gpad_plants_t;
int gpad_plants_battery(int n_u, int N, int B, const double* capacity_scale, int threads, gpad_plants_t* out);
int gpad_plants_destroy(gpad_plants_t p);
int gpad_plants_dims(gpad_plants_t p, int* n_u, int* N, int* m, int* n_par, int* B);
/* operators of all plants back to back ([B][n*m] each) in the requested layout: what gpad_setup takes in
 * GPAD_MODE_BATCH_PER_INSTANCE; L [B] (may be NULL) */
int gpad_plants_operators(gpad_plants_t p, int layout, float* M_G, float* G_L, float* L);
/* per-instance vectors, one parameter row per plant: params [B][n_par] -> g_P [B][n], p_D [B][m], f [B][n] or NULL (host) */
int gpad_plants_instances(gpad_plants_t p, const double* params, float* g_P, float* p_D, float* f);
/* gpad_closed_loop with one plant per instance; h is a GPAD_MODE_BATCH_PER_INSTANCE handle built from
 * gpad_plants_operators of the same plants (max_batch >= B); first <= 0 with count <= 0 means all plants,
 * otherwise plants [first, first + count) -- the shard this handle owns */
int gpad_closed_loop_plants(gpad_plants_t p, gpad_handle_t h, int first, int count, const double* x0, int samples,
                            const float* theta, const float* beta, int max_iter, int warm_start, double* x_traj,
                            double* u_traj);

/* ------------------------------------------------------------------------------------------
 * 3c. One logical solver over several devices of one box: the batch is cut into contiguous shards, one per device,
 *     operators are replicated (per-instance operators are sharded with their instances), every device runs its own
 *     handle from its own host thread, and results land in the caller's host buffers by asynchronous D2H copies.
 *     No collective, no NCCL: the path has no exchange step.  cfg->max_batch is the capacity of the
 *     whole group; cfg->device is ignored.  devices may repeat an ordinal (several shards on one GPU).
 *     Results do not depend on the number of devices.
 * ---------------------------------------------------------------------------------------- */
typedef struct gpad_group_s* gpad_group_t;
int gpad_group_setup(const gpad_config_t* cfg, const int* devices, int device_count, const float* M_G, const float* G_L,
                     gpad_group_t* out);
int gpad_group_destroy(gpad_group_t g);
/* GPAD_MEM_HOST arguments only; shards [0, batch) over the devices and solves them concurrently */
int gpad_group_solve(gpad_group_t g, const gpad_solve_args_t* args);
int gpad_group_size(gpad_group_t g);
/* shard i: its handle and the instance range [first, first + count) it owns for a batch of `batch` */
int gpad_group_shard(gpad_group_t g, int i, int batch, gpad_handle_t* h, int* first, int* count);

/* theta_v, beta_v for v = 0..count-1 (paper eq. 8e / acceldualgrad.m:55-56) */
int gpad_schedule(float* theta, float* beta, int count, int variant);

/* ------------------------------------------------------------------------------------------
 * 4. The reference's text data file (main.cu:29-67): header "n_u N m num_iterations L", then
 *    M_G (n*m), g_P (n), G_L (n*m), p_D (m), theta[num_iterations], beta[num_iterations].
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int n_u, N, m, num_iterations;
    float L;
    float *M_G, *g_P, *G_L, *p_D, *theta, *beta;   /* malloc'ed by gpad_file_read */
} gpad_file_t;

int gpad_file_read(const char* path, gpad_file_t* out);
int gpad_file_write(const char* path, const gpad_file_t* in);
void gpad_file_free(gpad_file_t* f);
/* the ENABLE_FLATTEN_MATRICES variant of the same file (main.cu:39-41,50-52): the two operators hold N*m floats
 * each (GPAD_LAYOUT_FLAT) instead of n*m; everything else is identical */
int gpad_file_read_flat(const char* path, gpad_file_t* out);
int gpad_file_write_flat(const char* path, const gpad_file_t* in);

/* The per-step fixtures of the reference's harnesses, "<dir>/input.txt" + "<dir>/output.txt":
 *   step 2 (main_prof.cu:117-156)  input: n_u N m, M_G, w[m], g_P[n];  output: M_G w [n], zhat [n]
 *   step 3 (step3.cu:59-81)        input: n_u N m theta, z_prev[n], zhat[n];  output: z [n]
 *   step 4 (main_prof.cu:198-239)  input: n_u N m, w[m], zhat[n], p_D[m], G_L;  output: G_L zhat [m], sum [m], y_next [m]
 * `flat` != 0: the operator holds N*m floats (the default fixtures of main_prof.cu:10, ENABLE_FLATTEN_MATRICES).
 * Absent vectors are NULL; gpad_fixture_read mallocs, gpad_fixture_free releases. */
typedef struct {
    int step;                       /* 2, 3 or 4 */
    int n_u, N, m, flat;
    float theta;                    /* step 3 */
    float *op;                      /* step 2: M_G, step 4: G_L */
    float *w, *g_P, *p_D, *zhat_in, *z_prev;               /* inputs  */
    float *prod, *sum, *zhat_out, *z_out, *y_next;         /* outputs */
} gpad_fixture_t;
int gpad_fixture_read(const char* dir, int step, int flat, gpad_fixture_t* out);
int gpad_fixture_write(const char* dir, const gpad_fixture_t* in);
void gpad_fixture_free(gpad_fixture_t* f);

/* ------------------------------------------------------------------------------------------
 * 5. Test hook: C[M][N] = A[M][K] * B[N][K]^T through the tcgen05 3xTF32 mainloop used by
 *    GPAD_PREC_TF32X3 (device pointers, row-major, any M,N,K >= 1).
 * ---------------------------------------------------------------------------------------- */
int gpad_debug_gemm_tf32x3(const float* A, const float* B, float* C, int M, int N, int K,
                           void* stream);
/* The same through the GPAD_PREC_FP16X3 path: rows of A and B scaled and split into fp16 hi / lo, kind::f16
 * MMAs, scales undone in the epilogue.  kernel 0 = shared-memory-operand kernel (product 2's mainloop),
 * 1 = TMEM-operand kernel (product 1's: A quantised by the transform warps inside the kernel),
 * 2 = the same with one accumulator stage and tiles of up to 256 columns (its single-wave plan). */
int gpad_debug_gemm_f16x3(const float* A, const float* B, float* C, int M, int N, int K, int kernel,
                          void* stream);

/* Test hook, host only (no device call): the column tiling the batch kernels use for an operator with
 * `ncols` output columns. kernel 0 = shared-memory-operand kernel (tiles <= 256 columns), 1 = the
 * TMEM-operand kernel (tiles <= 208 columns, 96 TMEM columns kept for the state ring), 2 = the
 * fp16 product 2 with the TMA-streamed epilogue (tiles of whole 32-column blocks, <= 256).
 * Tile t covers columns [t*step, t*step + bn); tmem_cols = TMEM columns the plan occupies (<= 512). */
int gpad_debug_plan_tiles(int kernel, int ncols, int* bn, int* n_tiles, int* step, int* tmem_cols);

#ifdef __cplusplus
}
#endif
#endif /* GPAD_H */
