// handle.h -- the solver handle behind gpad_handle_t and the pieces of host logic the api_*.cu files share.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "batch_common.cuh"
#include "batch_tc.h"
#include "gpad_internal.h"
#include "latency.h"

namespace gpad {

// Experiment switches.  ONE environment variable, GPAD_DEBUG="key=value,key=value", parsed ONCE in gpad_setup and
// stored in the handle; nothing on the solve path reads the environment.  Product defaults need no variable.
struct Knobs {
    std::string latency_plan;        // "" | "block" | "cluster:<C>" | "grid:<G>" | "lean:<C>"
    int latency_threads = 0;         // generic kernel: threads per CTA
    int latency_no_smem_ops = 0;     // generic kernel: stream the operators from L2
    int latency_grid2 = -1;          // -1 auto; 0 forces the generic whole-chip kernel
    int latency_warp = -1;           // -1 auto; 0 forces the one-CTA kernel for tiny problems
    int latency_flat = -1;           // -1 auto; 0 expands flat operators to the dense kernels
    int flat_xchg = 0;               // flat kernel exchange: 0 bulk DSMEM copies, 1 per-entry st.async
    int warp_rows = 0, warp_ordered = -1;   // one-warp kernel schedule (0 / -1: chosen by batch size)
    int warp_pack = 1;               // fixed-iteration per-instance batches: two QPs per warp
    int tc_p1 = -1;                  // product 1: -1 waves model, 1 TMEM-operand kernel, 0 shared-memory-operand kernel
    int tc_stages = 0;               // cap on its ring depth (0: as many as fit)
    int tc_bn2 = 0;                  // product 2 tile width (0: tuned, cached per shape)
    int tc_autotune = 1;             // 0: first candidate width without timing
    int tc_pdl = -1;                 // programmatic dependent launch between the batch kernels: -1 auto (fp16 plan, solves of
                                     // <= 16K instances), 0 off, 1 on (measured: -4 % at 8K instances, within noise at 64K)
    int tc_cluster_attr = 0;         // launch the shared-memory-operand kernel as clusters of one CTA
    int tc_retire = 1;               // tolerance mode: skip batch tiles whose instances have all stopped
    int tc_compact = 1;              // tolerance mode: gather the running instances into dense tiles (needs tc_retire)
    int check_lag = 4;               // tolerance mode: checks the host may run ahead of the device
    int sync_split = 1;              // synchronous host-memory solves of >= 16K instances run as two pipelined halves
};
Knobs parse_knobs();

struct BatchSlot {                   // one complete set of batch state (two exist when gpad_solve_async is used)
    BatchState st;
    tc::GemmDesc g1, g2;
    tc::GemmDesc g1h, g2h;           // GPAD_PREC_FP16X3: the fp16-operand plans of fixed-iteration solves
    cudaEvent_t ev_in = nullptr, ev_comp = nullptr, ev_out = nullptr;
    long long ticket = -1;           // async solve occupying this slot (-1: free)
    bool allocated = false;
};

}  // namespace gpad

struct gpad_handle_s {
    gpad_config_t cfg{};
    gpad::Knobs knobs;
    int n = 0, device = 0, num_sms = 0;
    size_t smem_optin = 0;
    std::string desc;
    long long launches = 0;
    cudaStream_t own_stream = nullptr;
    std::vector<void*> allocs;

    // ---- stream ordering of per-handle scratch (theta/beta tables, flags, batch state): every solve makes its stream
    // wait for the previous solve of this handle, whatever stream that ran on; one solve in flight per handle ----
    cudaEvent_t ev_last = nullptr;
    bool ev_last_valid = false;

    // ---- optional per-kernel event timing (gpad_profile_*) ----
    bool profile = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_used[3];
    static constexpr size_t kMaxProfiled = 1 << 16;     // pairs kept per kernel between reads
    cudaEvent_t prof_begin(cudaStream_t s) {
        if (!profile) return nullptr;
        cudaEvent_t e = take_event();
        cudaEventRecord(e, s);
        return e;
    }
    void prof_end(int which, cudaEvent_t begin, cudaStream_t s) {
        if (!profile || !begin) return;
        if (ev_used[which].size() >= kMaxProfiled) { ev_pool.push_back(begin); return; }
        cudaEvent_t e = take_event();
        cudaEventRecord(e, s);
        ev_used[which].push_back({begin, e});
    }
    cudaEvent_t take_event() {
        if (ev_pool.empty()) { cudaEvent_t e; cudaEventCreate(&e); return e; }
        cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e;
    }

    // ---- latency mode ----
    gpad::lat::Params lp{};
    int sync_mode = 0, G = 1, threads = 256;
    bool ops_smem = false;
    bool small = false;                       // lean one-CTA / cluster kernel (latency_small.cu)
    bool warp = false;                        // tiny problems: one warp per QP (latency_warp.cu)
    bool grid2 = false;                       // whole-chip plans run latency_grid2.cu
    bool flat = false;                        // fixed-iteration solves run latency_flat.cu (battery-structured operators)
    gpad::lat::FlatParams fp{};
    int cha = 1, chb = 1;
    float *d_gP = nullptr, *d_pD = nullptr, *d_f = nullptr, *d_y0 = nullptr, *d_yprev0 = nullptr;
    float *d_theta = nullptr, *d_beta = nullptr;
    int sched_cap = 0;
    std::vector<float> h_theta, h_beta;       // last uploaded schedule
    float *o_ynext = nullptr, *o_y = nullptr, *o_z = nullptr, *o_zhat = nullptr, *o_w = nullptr;
    int *o_iters = nullptr, *o_status = nullptr;
    float *o_viol = nullptr, *o_gap = nullptr;
    unsigned* d_flags = nullptr;              // [0] barrier counter, [1] nonfinite flag
    float *lat_h_in = nullptr, *lat_h_out = nullptr;   // pinned mirrors of the input block (d_gP ...) and the output block (o_ynext ...)
    size_t lat_in_floats = 0, lat_out_floats = 0;

    // ---- per-instance mode ----
    float *pi_gP = nullptr, *pi_pD = nullptr, *pi_f = nullptr, *pi_y0 = nullptr, *pi_yprev0 = nullptr;   // host-mode staging
    float *pi_ynext = nullptr, *pi_y = nullptr, *pi_z = nullptr, *pi_zhat = nullptr, *pi_w = nullptr;
    int *pi_iters = nullptr, *pi_status = nullptr;
    float *pi_viol = nullptr, *pi_gap = nullptr;

    // ---- batch mode (shared operators) ----
    gpad::Operators op;
    gpad::BatchSlot slot[2];                  // slot[1] is allocated by the first gpad_solve_async
    float* stage_in = nullptr;                // staging for host-memory inputs/outputs [max_batch][max(n,m)]
    float* stage_out = nullptr;               // async solves: staging of the w output
    double* d_params = nullptr;               // [max_batch][n_par] parameters of on-device instance builds
    int params_cap = 0;
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    int* h_active = nullptr;                  // pinned ring of {running, waiting-for-dual} counters, one pair per check in flight
    std::vector<cudaEvent_t> ev_check;
    long long next_ticket = 0;
    unsigned long long* h_stat = nullptr;     // pinned [2]: tolerance-mode statistics of the last solve (gpad_solve_stats)
    gpad::BatchState arch;                    // compacted tolerance-mode solves: outputs by original instance index
    bool arch_allocated = false, out_from_archive = false;
    int compactions = 0;                      // of the last solve
};

namespace gpad {

template <typename T>
int dev_alloc(gpad_handle_s* h, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
        cudaGetLastError();
        return GPAD_ERR_ALLOC;
    }
    h->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return GPAD_OK;
}

int upload_padded(gpad_handle_s* h, const float* src, int rows, int cols, int rows_pad, int ld_pad, float** out);
int upload_schedule(gpad_handle_s* h, const float* theta, const float* beta, int count, cudaStream_t s);
// makes `s` wait for the previous solve of this handle / records the end of this one
int solve_begin(gpad_handle_s* h, cudaStream_t s);
int solve_end(gpad_handle_s* h, cudaStream_t s);

// api_batch.cu
int setup_batch(gpad_handle_s* h, const std::vector<float>& MG, const std::vector<float>& GL);
int solve_batch(gpad_handle_s* h, const gpad_solve_args_t* a);
int solve_batch_async(gpad_handle_s* h, const gpad_solve_args_t* a, long long* ticket);
int wait_batch(gpad_handle_s* h, long long ticket);
void destroy_batch(gpad_handle_s* h);

// closed_loop.cu: per-instance vectors from parameters on the device (problem data cached per device)
int instances_device(gpad_problem_t prob, int device, int B, const double* params_dev, float* g_P, int ld_g, float* p_D, int ld_p,
                     float* f, int ld_f, cudaStream_t s);

}  // namespace gpad
