// latency.h -- parameters of the persistent single-QP kernel (latency.cu)
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace gpad {
namespace lat {

enum { SYNC_BLOCK = 0, SYNC_CLUSTER = 1, SYNC_GRID = 2 };
constexpr int kMaxThreads = 512;
constexpr int kRegChunks = 8;    // float4 operator fragments per lane and phase in the register-resident variant

struct Params {
    int n, m;
    int nld, mld;            // row strides of G_L / M_G: n, m rounded up to 4 * lanes-per-row floats
    int rows_a, rows_b;      // rows of M_G / G_L per CTA (ceil(n/G), ceil(m/G))
    int rows_a_pad, rows_b_pad;  // rounded up to 4
    int g_pad;               // G rounded up to 4
    int lg_a, lg_b;          // log2(lanes per row) in phase A / B
    int res_a, res_b;        // own rows of M_G / G_L resident in shared memory (memory variant)
    int sched_smem;          // latency_small.cu: theta/beta entries staged in shared memory
    int warp_rows, warp_ordered;  // latency_warp.cu schedule override (0 / -1: chosen by batch size)
    int warp_pack;                // fixed-iteration batches: two QPs per warp (1 default / 0 off)
    int batch;               // SYNC_BLOCK only: independent instances, one CTA each (per-instance operators)
    size_t op_stride_a, op_stride_b;   // elements between consecutive instances' M_G / G_L (0: shared)
    const float* M_G;        // [n][mld] sequential layout, zero padded
    const float* G_L;        // [m][nld]
    const float* g_P;        // [n]
    const float* p_D;        // [m]
    const float* f;          // [n] or null
    const float* y0;         // [m] or null
    const float* y_prev0;    // [m] or null
    const float* theta;      // device [max_iter]
    const float* beta;       // device [max_iter]
    int max_iter, check_every;
    float eps_g, eps_V, L;
    // outputs (device, always valid)
    float *out_y_next, *out_y, *out_z, *out_zhat, *out_w;
    int *out_iters, *out_status;
    float *out_max_viol, *out_gap;
    // grid-mode exchange buffers
    float* x_w;              // [m]
    float* x_zhat;           // [n]
    float* x_red;            // [3][8][g_pad]
    unsigned* barrier;       // zeroed before launch
    int* nonfinite_flag;     // zeroed before launch
};

// latency_flat.cu: one QP with battery-structured ("flattened") operators on one thread-block cluster
struct FlatParams {
    int n_u, N, m;
    int Q;                   // box multipliers per cell (4 N)
    int N4;                  // stage count rounded up to the phase-B fragment (4 * CHB floats)
    int SC;                  // stages per CTA in phase A
    int CHA, CHB;            // float4 chunks per lane (phase A) / per thread (phase B)
    int lenA;                // floats per phase-A operator row: [Q | tail | zero padding] = 128 * CHA
    int w_len;               // floats of the permuted multiplier vector in shared memory (with zero padding)
    int sched;               // theta / beta entries staged in shared memory
    int C, threads, rows_b;  // cluster size, threads per CTA, rows of G_L per CTA
    int xchg;                // 0: bulk copies of each CTA's block, 1: per-entry asynchronous stores
    const float* A_op;       // [n][lenA]   row (s, u): cell u's box entries of M_G, then the sum-constraint entries
    const float* B_op;       // [m][N4]     row i: its N entries of G_L
    const float* g_P; const float* p_D; const float* y0; const float* y_prev0;
    const float* theta; const float* beta;
    int max_iter;
    float *out_y_next, *out_y, *out_z, *out_zhat, *out_w;
    int *out_iters, *out_status;
    float *out_max_viol, *out_gap;
    int* nonfinite_flag;     // zeroed before launch
};
size_t smem_bytes(const Params& p, bool regs);
int launch(const Params& p, int sync_mode, bool regs, int G, int threads, cudaStream_t stream);
int launch_convert_ops(float* dst, const float* src, int B, int rows, int cols, int ld, bool flipped, cudaStream_t stream);
int max_cluster_size(int threads, size_t smem);
// latency_small.cu: lean one-CTA-per-QP kernel (operators + per-row state in registers)
size_t small_smem_bytes(const Params& p);
int small_sched_capacity();
int launch_small(const Params& p, int cha, int chb, int cluster, int threads, cudaStream_t stream);

// latency_flat.cu
bool plan_flat(int n_u, int N, int m, size_t smem_limit, int max_cluster, FlatParams* out);
int launch_flat(const FlatParams& p, cudaStream_t stream);
float build_flat_operators(const FlatParams& p, const float* MG, const float* GL, std::vector<float>& A_op, std::vector<float>& B_op);
// latency_warp.cu: one warp per QP for the tiny problems (n <= 16, m <= 64), latency and per-instance batch modes
int warp_supported(const Params& p);
int launch_warp(const Params& p, cudaStream_t stream);
// latency_grid2.cu: second-generation whole-chip kernel (column partition, vectors in registers, fixed-iteration solves)
size_t grid2_smem_bytes(const Params& p);
int grid2_supported(const Params& p, size_t smem_limit);
int launch_grid2(const Params& p, int G, cudaStream_t stream);

}  // namespace lat
}  // namespace gpad
