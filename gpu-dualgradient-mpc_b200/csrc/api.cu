// api.cu -- the C ABI of libgpad_b200.so (include/gpad.h): handle management, operator
// conversion, path selection and the host loop that drives the kernels.  Replaces the body of
// the reference's main() between readData and the D2H copies (main.cu:108-180).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "batch_common.cuh"
#include "batch_tc.h"
#include "gpad_internal.h"
#include "latency.h"

namespace gpad {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? GPAD_ERR_NO_DEVICE : GPAD_ERR_CUDA;
}

}  // namespace gpad

using namespace gpad;

struct gpad_handle_s {
    gpad_config_t cfg{};
    int n = 0, device = 0, num_sms = 0;
    size_t smem_optin = 0;
    std::string desc;
    long long launches = 0;
    cudaStream_t own_stream = nullptr;
    std::vector<void*> allocs;

    // ---- optional per-kernel event timing (gpad_profile_*) ----
    bool profile = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_used[3];
    cudaEvent_t prof_begin(cudaStream_t s) {
        if (!profile) return nullptr;
        cudaEvent_t e = take_event();
        cudaEventRecord(e, s);
        return e;
    }
    void prof_end(int which, cudaEvent_t begin, cudaStream_t s) {
        if (!profile || !begin) return;
        cudaEvent_t e = take_event();
        cudaEventRecord(e, s);
        ev_used[which].push_back({begin, e});
    }
    cudaEvent_t take_event() {
        if (ev_pool.empty()) { cudaEvent_t e; cudaEventCreate(&e); return e; }
        cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e;
    }

    // ---- latency mode ----
    lat::Params lp{};
    int sync_mode = 0, G = 1, threads = 256;
    bool ops_smem = false;
    bool small = false;                       // lean one-CTA kernel (latency_small.cu)
    bool grid_lean = false;                   // lean whole-chip kernel (latency_grid.cu)
    bool warp = false;                        // tiny problems: one warp per QP (latency_warp.cu) for solves without f
    bool grid2 = false;                       // fixed-iteration solves run latency_grid2.cu (whole chip, vectors in registers)
    unsigned stamp_next = 16;                 // flag-in-data exchange epochs
    size_t ll_words = 0;
    int cha = 1, chb = 1;
    float *d_gP = nullptr, *d_pD = nullptr, *d_f = nullptr, *d_y0 = nullptr, *d_yprev0 = nullptr;
    float *d_theta = nullptr, *d_beta = nullptr;
    int sched_cap = 0;
    std::vector<float> h_theta, h_beta;       // last uploaded schedule
    float *o_ynext = nullptr, *o_y = nullptr, *o_z = nullptr, *o_zhat = nullptr, *o_w = nullptr;
    int *o_iters = nullptr, *o_status = nullptr;
    float *o_viol = nullptr, *o_gap = nullptr;
    unsigned* d_flags = nullptr;              // [0] barrier counter, [1] nonfinite flag

    // ---- per-instance mode (one CTA per QP, latency kernel in SYNC_BLOCK) ----
    float *pi_gP = nullptr, *pi_pD = nullptr, *pi_f = nullptr, *pi_y0 = nullptr, *pi_yprev0 = nullptr;   // host-mode staging
    float *pi_ynext = nullptr, *pi_y = nullptr, *pi_z = nullptr, *pi_zhat = nullptr, *pi_w = nullptr;
    int *pi_iters = nullptr, *pi_status = nullptr;
    float *pi_viol = nullptr, *pi_gap = nullptr;

    // ---- batch mode ----
    Operators op;
    BatchState st;
    tc::GemmDesc g1, g2;
    float* stage_in = nullptr;                // staging for host-memory inputs/outputs [max_batch][max(n,m)]
    float* stage_out = nullptr;               // pipelined host solves: separate staging for the outputs
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    int* h_active = nullptr;                  // pinned
};

namespace {

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
#define GPAD_TRY(expr) do { int rc__ = (expr); if (rc__ != GPAD_OK) return rc__; } while (0)



// operators -> host, sequential layout (M_G [n][m], G_L [m][n])
int fetch_operators(const gpad_config_t& c, const float* M_G, const float* G_L, size_t count_each, int copies,
                    std::vector<float>& MG, std::vector<float>& GL) {
    const size_t flat_each = (size_t)c.N * c.m;
    const size_t total = (c.layout == GPAD_LAYOUT_FLAT ? flat_each : count_each) * copies;
    std::vector<float> a(total), b(total);
    if (c.operators_mem == GPAD_MEM_DEVICE) {
        GPAD_CUDA(cudaMemcpy(a.data(), M_G, total * sizeof(float), cudaMemcpyDeviceToHost));
        GPAD_CUDA(cudaMemcpy(b.data(), G_L, total * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        memcpy(a.data(), M_G, total * sizeof(float));
        memcpy(b.data(), G_L, total * sizeof(float));
    }
    if (c.layout == GPAD_LAYOUT_SEQUENTIAL) { MG.swap(a); GL.swap(b); return GPAD_OK; }
    const int n = c.n_u * c.N, m = c.m;
    if (c.layout == GPAD_LAYOUT_FLAT) {      // flattened battery operators: expand to dense sequential
        MG.assign((size_t)n * m, 0.f); GL.assign((size_t)n * m, 0.f);
        return gpad_expand_operators(c.n_u, c.N, m, a.data(), b.data(), MG.data(), GL.data());
    }
    MG.resize(total); GL.resize(total);
    for (int k = 0; k < copies; ++k) {
        const float* fa = a.data() + k * count_each; const float* fb = b.data() + k * count_each;
        float* sa = MG.data() + k * count_each; float* sb = GL.data() + k * count_each;
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < n; ++i) sa[(size_t)i * m + j] = fa[(size_t)j * n + i];    // flipped M_G is [m][n]
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) sb[(size_t)i * n + j] = fb[(size_t)j * m + i];    // flipped G_L is [n][m]
    }
    return GPAD_OK;
}

// rows x cols (ld = cols) -> device rows_pad x ld_pad, zero padded
int upload_padded(gpad_handle_s* h, const float* src, int rows, int cols, int rows_pad, int ld_pad, float** out) {
    std::vector<float> tmp((size_t)rows_pad * ld_pad, 0.f);
    for (int r = 0; r < rows; ++r) memcpy(&tmp[(size_t)r * ld_pad], src + (size_t)r * cols, sizeof(float) * cols);
    GPAD_TRY(dev_alloc(h, out, tmp.size()));
    GPAD_CUDA(cudaMemcpy(*out, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice));
    return GPAD_OK;
}

// ---- latency plan: how many CTAs cooperate, how they synchronise, where the operators live ----
struct LatPlan {
    int sync = -1, G = 1, threads = 64;
    bool regs = false;
    bool small = false;
    bool grid_lean = false;
    int cha = 1, chb = 1;
    lat::Params p{};
};

// CTAs per cluster that share each operator tile through TMA multicast in the cta_group::1 tcgen05 kernels
int tc_multicast() {
    int mc = 1;
    if (const char* e = getenv("GPAD_TC_MC")) mc = atoi(e) == 2 ? 2 : 1;
    return mc;
}

int ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// fills row split, lanes-per-row, strides, residency for G cooperating CTAs; returns false if infeasible
bool plan_for(int n, int m, int G, bool want_regs, size_t smem_limit, bool no_resident, LatPlan& out) {
    lat::Params& p = out.p;
    p.n = n; p.m = m;
    p.rows_a = (n + G - 1) / G; p.rows_b = (m + G - 1) / G;
    p.rows_a_pad = round_up(p.rows_a, 4); p.rows_b_pad = round_up(p.rows_b, 4);
    p.g_pad = round_up(G, 4);
    const int m4 = (m + 3) / 4, n4 = (n + 3) / 4;
    out.G = G;
    if (want_regs) {
        // one pass per phase, operator fragments in registers: lanes-per-row = smallest power of two that
        // leaves <= 4 (else <= kRegChunks) float4 per lane while all rows of the CTA fit in 512 threads
        auto pick = [&](int len4, int rows, int& lg) {
            for (int target : {4, lat::kRegChunks})
                for (lg = 0; lg <= 5; ++lg)
                    if (((len4 + (1 << lg) - 1) >> lg) <= target && (rows << lg) <= lat::kMaxThreads) return true;
            return false;
        };
        if (!pick(m4, p.rows_a, p.lg_a) || !pick(n4, p.rows_b, p.lg_b)) return false;
        out.threads = std::max(32, round_up(std::max(p.rows_a << p.lg_a, p.rows_b << p.lg_b), 32));
        p.res_a = p.res_b = 0;
        out.regs = true;
    } else {
        p.lg_a = std::min(5, ilog2_ceil(m4)); p.lg_b = std::min(5, ilog2_ceil(n4));
        out.threads = std::min(lat::kMaxThreads, std::max(64, round_up(std::max(p.rows_a << p.lg_a, p.rows_b << p.lg_b), 32)));
        out.regs = false;
    }
    p.mld = round_up(m, 4 << p.lg_a); p.nld = round_up(n, 4 << p.lg_b);
    if (!out.regs) {
        // partial residency: as many own rows as fit next to the vectors, split evenly between the operators
        p.res_a = p.res_b = 0;
        const size_t fixed = lat::smem_bytes(p, false);
        if (fixed > smem_limit) return false;
        if (!no_resident) {
            const size_t budget = smem_limit - fixed;
            const size_t want = ((size_t)p.rows_a * p.mld + (size_t)p.rows_b * p.nld) * sizeof(float);
            const double frac = want ? std::min(1.0, (double)budget / (double)want) : 1.0;
            p.res_a = (int)(p.rows_a * frac);
            size_t left = budget - (size_t)p.res_a * p.mld * sizeof(float);
            p.res_b = (int)std::min<size_t>(p.rows_b, left / ((size_t)p.nld * sizeof(float)));
        }
    }
    return lat::smem_bytes(p, out.regs) <= smem_limit;
}

// lean one-CTA plan (latency_small.cu): every row of both operators gets lanes-per-row = 2^lg lanes with
// CH in {1,2,4,8} float4 fragments per lane; prefers <= 4 fragments, all rows must fit in 512 threads
bool plan_small(int n, int m, int C, LatPlan& out) {
    lat::Params& p = out.p;
    p = lat::Params{};
    p.n = n; p.m = m;
    const int rows_a = (n + C - 1) / C, rows_b = (m + C - 1) / C;
    const int m4 = (m + 3) / 4, n4 = (n + 3) / 4;
    auto pick = [&](int len4, int rows, int& lg, int& ch) {
        for (int target : {4, 8})
            for (lg = 0; lg <= 5; ++lg) {
                const int chunks = (len4 + (1 << lg) - 1) >> lg;
                if (chunks <= target && (rows << lg) <= lat::kMaxThreads) {
                    ch = 1;
                    while (ch < chunks) ch <<= 1;
                    return true;
                }
            }
        return false;
    };
    if (!pick(m4, rows_a, p.lg_a, out.cha) || !pick(n4, rows_b, p.lg_b, out.chb)) return false;
    p.mld = (4 << p.lg_a) * out.cha; p.nld = (4 << p.lg_b) * out.chb;
    out.threads = std::max(32, round_up(std::max(rows_a << p.lg_a, rows_b << p.lg_b), 32));
    p.rows_a = rows_a; p.rows_b = rows_b; p.rows_a_pad = round_up(rows_a, 4); p.rows_b_pad = round_up(rows_b, 4);
    p.g_pad = round_up(C, 4);
    p.res_a = p.res_b = 0;
    out.G = C; out.sync = C > 1 ? lat::SYNC_CLUSTER : lat::SYNC_BLOCK; out.regs = true; out.small = true;
    return true;
}

// lean whole-chip plan (latency_grid.cu): per phase the 16 warps form cw column slices x rg row groups with
// rb <= 4 rows per warp and ch <= 6 float4 chunks per lane; the option with the fewest shared-memory loads wins
bool plan_grid(int n, int m, int G, size_t smem_limit, bool no_resident, LatPlan& out) {
    lat::Params& p = out.p;
    p = lat::Params{};
    p.n = n; p.m = m;
    p.rows_a = (n + G - 1) / G; p.rows_b = (m + G - 1) / G;
    p.rows_a_pad = round_up(p.rows_a, 4); p.rows_b_pad = round_up(p.rows_b, 4);
    p.g_pad = round_up(G, 4);
    if (p.rows_a > 512 || p.rows_b > 512) return false;
    auto pick = [](int len4, int rows, int& cw, int& rg, int& rb, int& ch) {
        long best = -1;
        for (int g = 1; g <= 16; g <<= 1) {
            const int c = 16 / g;
            const int b = std::min(4, (rows + g - 1) / g);
            const int h = (len4 + c * 32 - 1) / (c * 32);
            if (h > 6) continue;
            const int passes = (rows + g * b - 1) / (g * b);
            const long cost = (long)passes * h * (b + 1);
            if (best < 0 || cost < best) { best = cost; cw = c; rg = g; rb = b; ch = h; }
        }
        return best >= 0;
    };
    if (!pick((m + 3) / 4, p.rows_a, p.cwa, p.rga, p.rba, p.cha) || !pick((n + 3) / 4, p.rows_b, p.cwb, p.rgb, p.rbb, p.chb)) return false;
    p.mld = p.cwa * 32 * 4 * p.cha; p.nld = p.cwb * 32 * 4 * p.chb;
    p.lg_a = p.lg_b = 5;
    p.res_a = p.res_b = 0;
    const size_t fixed = lat::grid_smem_bytes(p);
    if (fixed > smem_limit) return false;
    if (!no_resident) {
        const size_t budget = smem_limit - fixed;
        const size_t want = ((size_t)p.rows_a * p.mld + (size_t)p.rows_b * p.nld) * sizeof(float);
        const double frac = want ? std::min(1.0, (double)budget / (double)want) : 1.0;
        p.res_a = (int)(p.rows_a * frac);
        const size_t left = budget - (size_t)p.res_a * p.mld * sizeof(float);
        p.res_b = (int)std::min<size_t>(p.rows_b, left / ((size_t)p.nld * sizeof(float)));
    }
    out.G = G; out.sync = lat::SYNC_GRID; out.regs = false; out.small = false; out.grid_lean = true; out.threads = 512;
    return lat::grid_smem_bytes(p) <= smem_limit;
}

// ------------------------------------------------------------------ latency mode
int setup_latency(gpad_handle_s* h, const std::vector<float>& MG, const std::vector<float>& GL) {
    const int n = h->n, m = h->cfg.m;
    const size_t limit = h->smem_optin;
    const bool no_res = getenv("GPAD_LATENCY_NO_SMEM_OPS") != nullptr;
    LatPlan plan;
    const char* env = getenv("GPAD_LATENCY_PLAN");    // "block" | "cluster:<C>" | "grid:<G>" (experiments)
    bool ok = false;
    if (env) {
        int sync = lat::SYNC_GRID, G = h->num_sms;
        if (!strncmp(env, "block", 5)) { sync = lat::SYNC_BLOCK; G = 1; }
        else if (!strncmp(env, "cluster:", 8)) { sync = lat::SYNC_CLUSTER; G = std::max(1, std::min(16, atoi(env + 8))); }
        else if (!strncmp(env, "lean:", 5)) { sync = lat::SYNC_CLUSTER; G = std::max(1, std::min(16, atoi(env + 5))); }
        else if (!strncmp(env, "grid:", 5)) { G = std::max(1, std::min(h->num_sms, atoi(env + 5))); }
        else if (!strncmp(env, "leangrid:", 9)) { G = std::max(1, std::min(h->num_sms, atoi(env + 9))); }
        if (!strncmp(env, "lean:", 5)) {
            ok = plan_small(n, m, G, plan);
        } else if (!strncmp(env, "leangrid:", 9)) {
            ok = plan_grid(n, m, G, limit, no_res, plan);
        } else {
            ok = (!no_res && plan_for(n, m, G, true, limit, no_res, plan)) || plan_for(n, m, G, false, limit, no_res, plan);
            plan.sync = sync;
        }
    } else {
        // 1. one CTA with register-resident operators (lean kernel); 2. the smallest cluster that allows
        // register residency; 3. the whole chip, operators in shared memory as far as they fit
        if (plan_small(n, m, 1, plan)) ok = true;
        const int max_cluster = ok ? 1 : lat::max_cluster_size(lat::kMaxThreads, 64 * 1024);
        for (int C : {2, 4, 8, 16}) {
            if (ok || C > max_cluster) break;
            if (plan_small(n, m, C, plan)) ok = true;
        }
        // (the lean flag-in-data grid kernel, GPAD_LATENCY_PLAN=leangrid:<G>, measured 2x slower than this
        //  generic grid plan on B200 and is opt-in only: DESIGN.md 4.3)
        if (!ok) { ok = plan_for(n, m, h->num_sms, false, limit, no_res, plan); plan.sync = lat::SYNC_GRID; }
    }
    if (!ok) {
        set_error("latency mode: vectors of n=%d, m=%d do not fit in shared memory", n, m);
        return GPAD_ERR_UNSUPPORTED;
    }
    if (const char* t = getenv("GPAD_LATENCY_THREADS"))
        if (!plan.regs && atoi(t) > 0) plan.threads = std::max(32, std::min(lat::kMaxThreads, atoi(t) / 32 * 32));
    lat::Params& p = h->lp;
    p = plan.p;
    h->sync_mode = plan.sync; h->G = plan.G; h->threads = plan.threads; h->ops_smem = plan.regs;
    h->small = plan.small; h->cha = plan.cha; h->chb = plan.chb; h->grid_lean = plan.grid_lean;
    p.L = h->cfg.L;
    p.batch = 1; p.op_stride_a = p.op_stride_b = 0;
    if (plan.grid_lean) {
        h->ll_words = 2 * ((size_t)m + n + (size_t)plan.G * 8);
        unsigned long long* ll = nullptr;
        GPAD_TRY(dev_alloc(h, &ll, h->ll_words));
        GPAD_CUDA(cudaMemset(ll, 0, h->ll_words * sizeof(unsigned long long)));
        p.ll_w = ll; p.ll_z = ll + 2 * (size_t)m; p.ll_r = p.ll_z + 2 * (size_t)n;
    }
    float *dMG, *dGL;
    GPAD_TRY(upload_padded(h, MG.data(), n, m, n, p.mld, &dMG));
    GPAD_TRY(upload_padded(h, GL.data(), m, n, m, p.nld, &dGL));
    p.M_G = dMG; p.G_L = dGL;

    // per-solve device buffers
    GPAD_TRY(dev_alloc(h, &h->d_gP, n)); GPAD_TRY(dev_alloc(h, &h->d_pD, m)); GPAD_TRY(dev_alloc(h, &h->d_f, n));
    GPAD_TRY(dev_alloc(h, &h->d_y0, m)); GPAD_TRY(dev_alloc(h, &h->d_yprev0, m));
    GPAD_TRY(dev_alloc(h, &h->o_ynext, m)); GPAD_TRY(dev_alloc(h, &h->o_y, m)); GPAD_TRY(dev_alloc(h, &h->o_w, m));
    GPAD_TRY(dev_alloc(h, &h->o_z, n)); GPAD_TRY(dev_alloc(h, &h->o_zhat, n));
    GPAD_TRY(dev_alloc(h, &h->o_iters, 1)); GPAD_TRY(dev_alloc(h, &h->o_status, 1));
    GPAD_TRY(dev_alloc(h, &h->o_viol, 1)); GPAD_TRY(dev_alloc(h, &h->o_gap, 1));
    GPAD_TRY(dev_alloc(h, &p.x_w, p.mld)); GPAD_TRY(dev_alloc(h, &p.x_zhat, p.nld));
    GPAD_CUDA(cudaMemset(p.x_w, 0, sizeof(float) * p.mld));          // the zero padding is gathered too
    GPAD_CUDA(cudaMemset(p.x_zhat, 0, sizeof(float) * p.nld));
    GPAD_TRY(dev_alloc(h, &p.x_red, 3 * 8 * p.g_pad));
    GPAD_TRY(dev_alloc(h, &h->d_flags, 2));
    p.barrier = h->d_flags; p.nonfinite_flag = reinterpret_cast<int*>(h->d_flags + 1);

    char where[192];
    if (plan.grid_lean) snprintf(where, sizeof(where), "%d/%d + %d/%d rows per CTA in shared memory (lean grid kernel: warps %dx%d | %dx%d, flag-in-data exchange)",
                                 p.res_a, p.rows_a, p.res_b, p.rows_b, p.cwa, p.rga, p.cwb, p.rgb);
    else if (plan.small) snprintf(where, sizeof(where), "and per-row state in registers (lean kernel, %dx%d fragments)", plan.cha, plan.chb);
    else if (plan.regs) snprintf(where, sizeof(where), "in registers");
    else snprintf(where, sizeof(where), "%d/%d + %d/%d rows per CTA in shared memory, rest streamed from L2", p.res_a, p.rows_a, p.res_b, p.rows_b);
    char buf[448];
    snprintf(buf, sizeof(buf), "latency: persistent kernel, %s x%d CTAs, %d threads, lanes/row %d|%d, operators %s, smem %zu B/CTA",
             plan.sync == lat::SYNC_BLOCK ? "single-CTA" : plan.sync == lat::SYNC_CLUSTER ? "cluster(DSMEM)" : "cooperative-grid",
             plan.G, plan.threads, 1 << p.lg_a, 1 << p.lg_b, where,
             plan.grid_lean ? lat::grid_smem_bytes(p) : plan.small ? lat::small_smem_bytes(p) : lat::smem_bytes(p, plan.regs));
    h->desc = buf;
    // fixed-iteration solves of a whole-chip plan run the second-generation kernel when it covers the problem
    h->grid2 = plan.sync == lat::SYNC_GRID && !plan.small && !plan.grid_lean && plan.G == h->num_sms &&
               lat::grid2_supported(p, limit);
    if (const char* e = getenv("GPAD_LATENCY_GRID2")) h->grid2 = h->grid2 && atoi(e) != 0;
    h->warp = plan.small && plan.G == 1 && lat::warp_supported(p);
    if (const char* e = getenv("GPAD_LATENCY_WARP")) h->warp = h->warp && atoi(e) != 0;
    if (h->warp) h->desc = "latency: one warp, operators and state in registers, no shared memory or block barrier in the loop (latency_warp.cu)";
    if (h->grid2) {
        snprintf(buf, sizeof(buf), "latency: persistent kernel, cooperative-grid x%d CTAs, 512 threads, column-partitioned GEMV: exchanged "
                 "vectors in registers, M_G rows in shared memory (%zu B/CTA), G_L fragments in registers, counter barrier, "
                 "all termination branches in-kernel", plan.G, lat::grid2_smem_bytes(p));
        h->desc = buf;
    }
    return GPAD_OK;
}

int upload_schedule(gpad_handle_s* h, const float* theta, const float* beta, int count, cudaStream_t s) {
    if ((int)h->h_theta.size() == count && !memcmp(h->h_theta.data(), theta, sizeof(float) * count) &&
        !memcmp(h->h_beta.data(), beta, sizeof(float) * count))
        return GPAD_OK;
    if (count > h->sched_cap) {
        const int cap = std::max(count, 256);
        GPAD_TRY(dev_alloc(h, &h->d_theta, cap)); GPAD_TRY(dev_alloc(h, &h->d_beta, cap));
        h->sched_cap = cap;
    }
    h->h_theta.assign(theta, theta + count);
    h->h_beta.assign(beta, beta + count);
    GPAD_CUDA(cudaMemcpyAsync(h->d_theta, h->h_theta.data(), sizeof(float) * count, cudaMemcpyHostToDevice, s));
    GPAD_CUDA(cudaMemcpyAsync(h->d_beta, h->h_beta.data(), sizeof(float) * count, cudaMemcpyHostToDevice, s));
    return GPAD_OK;
}

int solve_latency(gpad_handle_s* h, const gpad_solve_args_t* a) {
    const int n = h->n, m = h->cfg.m;
    const bool host = a->mem == GPAD_MEM_HOST;
    cudaStream_t s = host ? h->own_stream : static_cast<cudaStream_t>(a->stream);
    GPAD_TRY(upload_schedule(h, a->theta, a->beta, a->max_iter, s));
    lat::Params p = h->lp;
    if (host) {
        GPAD_CUDA(cudaMemcpyAsync(h->d_gP, a->g_P, sizeof(float) * n, cudaMemcpyHostToDevice, s));
        GPAD_CUDA(cudaMemcpyAsync(h->d_pD, a->p_D, sizeof(float) * m, cudaMemcpyHostToDevice, s));
        if (a->f) GPAD_CUDA(cudaMemcpyAsync(h->d_f, a->f, sizeof(float) * n, cudaMemcpyHostToDevice, s));
        if (a->y0) GPAD_CUDA(cudaMemcpyAsync(h->d_y0, a->y0, sizeof(float) * m, cudaMemcpyHostToDevice, s));
        if (a->y_prev0) GPAD_CUDA(cudaMemcpyAsync(h->d_yprev0, a->y_prev0, sizeof(float) * m, cudaMemcpyHostToDevice, s));
        p.g_P = h->d_gP; p.p_D = h->d_pD; p.f = a->f ? h->d_f : nullptr;
        p.y0 = a->y0 ? h->d_y0 : nullptr; p.y_prev0 = a->y_prev0 ? h->d_yprev0 : nullptr;
    } else {
        p.g_P = a->g_P; p.p_D = a->p_D; p.f = a->f; p.y0 = a->y0; p.y_prev0 = a->y_prev0;
    }
    p.theta = h->d_theta; p.beta = h->d_beta;
    p.max_iter = a->max_iter; p.check_every = a->check_every > 0 ? a->check_every : 0;
    p.eps_g = a->eps_g; p.eps_V = a->eps_V;
    // in device mode the kernel writes straight into the caller's buffers; the two vectors the
    // dual-gap branch parks (w, zhat) always need real storage
    const bool dev = !host;
    p.out_y_next = dev && a->y_next ? a->y_next : h->o_ynext;
    p.out_y = dev && a->y ? a->y : h->o_y;
    p.out_z = dev && a->z ? a->z : h->o_z;
    p.out_zhat = dev && a->zhat ? a->zhat : h->o_zhat;
    p.out_w = dev && a->w ? a->w : h->o_w;
    p.out_iters = dev && a->iters ? a->iters : h->o_iters;
    p.out_status = dev && a->status ? a->status : h->o_status;
    p.out_max_viol = dev && a->max_viol ? a->max_viol : h->o_viol;
    p.out_gap = dev && a->gap ? a->gap : h->o_gap;
    cudaEvent_t pe = h->prof_begin(s);
    if (h->warp && p.max_iter >= 1) {
        p.batch = 1; p.op_stride_a = 0; p.op_stride_b = 0;
        GPAD_TRY(lat::launch_warp(p, s));
    } else if (h->small) {
        p.sched_smem = round_up(std::min(a->max_iter, lat::small_sched_capacity()), 4);
        GPAD_TRY(lat::launch_small(p, h->cha, h->chb, h->G, h->threads, s));
    } else if (h->grid_lean) {
        // every solve gets its own stamp range so words of earlier solves never match
        const unsigned need = 3u * (unsigned)(4 * a->max_iter + 64);
        if (h->stamp_next > 0xFFFFFFFFu - need - 16u) {
            GPAD_CUDA(cudaMemsetAsync(p.ll_w, 0, h->ll_words * sizeof(unsigned long long), s));
            h->stamp_next = 16;
        }
        p.stamp_base = h->stamp_next;
        h->stamp_next += need;
        GPAD_TRY(lat::launch_grid(p, h->G, s));
    } else if (h->grid2 && p.max_iter >= 1) {
        GPAD_CUDA(cudaMemsetAsync(h->d_flags, 0, 2 * sizeof(unsigned), s));
        GPAD_TRY(lat::launch_grid2(p, h->G, s));
    } else {
        GPAD_CUDA(cudaMemsetAsync(h->d_flags, 0, 2 * sizeof(unsigned), s));
        GPAD_TRY(lat::launch(p, h->sync_mode, h->ops_smem, h->G, h->threads, s));
    }
    h->prof_end(0, pe, s);
    h->launches += 1;
    if (host) {
        if (a->y_next) GPAD_CUDA(cudaMemcpyAsync(a->y_next, h->o_ynext, sizeof(float) * m, cudaMemcpyDeviceToHost, s));
        if (a->y) GPAD_CUDA(cudaMemcpyAsync(a->y, h->o_y, sizeof(float) * m, cudaMemcpyDeviceToHost, s));
        if (a->w) GPAD_CUDA(cudaMemcpyAsync(a->w, h->o_w, sizeof(float) * m, cudaMemcpyDeviceToHost, s));
        if (a->z) GPAD_CUDA(cudaMemcpyAsync(a->z, h->o_z, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
        if (a->zhat) GPAD_CUDA(cudaMemcpyAsync(a->zhat, h->o_zhat, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
        if (a->iters) GPAD_CUDA(cudaMemcpyAsync(a->iters, h->o_iters, sizeof(int), cudaMemcpyDeviceToHost, s));
        if (a->status) GPAD_CUDA(cudaMemcpyAsync(a->status, h->o_status, sizeof(int), cudaMemcpyDeviceToHost, s));
        if (a->max_viol) GPAD_CUDA(cudaMemcpyAsync(a->max_viol, h->o_viol, sizeof(float), cudaMemcpyDeviceToHost, s));
        if (a->gap) GPAD_CUDA(cudaMemcpyAsync(a->gap, h->o_gap, sizeof(float), cudaMemcpyDeviceToHost, s));
        GPAD_CUDA(cudaStreamSynchronize(s));
    }
    return GPAD_OK;
}

// ------------------------------------------------------------------ batch, per-instance operators
// Every QP has its own M_G / G_L, so there is nothing to share between instances: each CTA runs
// the persistent latency kernel on one QP (batched GEMV).  Operators are read from HBM once per
// solve (into registers when the plan allows), not once per iteration.
int setup_per_instance(gpad_handle_s* h, const float* M_G, const float* G_L) {
    const int n = h->n, m = h->cfg.m, B = h->cfg.max_batch;
    LatPlan plan;
    if (plan_small(n, m, 1, plan)) plan.sync = lat::SYNC_BLOCK;
    else if (plan_for(n, m, 1, true, h->smem_optin, false, plan)) plan.sync = lat::SYNC_BLOCK;
    else if (plan_for(n, m, 1, false, h->smem_optin, false, plan)) plan.sync = lat::SYNC_BLOCK;
    else { set_error("per-instance mode: n=%d, m=%d does not fit one CTA", n, m); return GPAD_ERR_UNSUPPORTED; }
    lat::Params& p = h->lp;
    p = plan.p;
    h->sync_mode = lat::SYNC_BLOCK; h->G = 1; h->threads = plan.threads; h->ops_smem = plan.regs;
    h->small = plan.small; h->cha = plan.cha; h->chb = plan.chb;
    p.L = h->cfg.L;
    p.batch = B;
    p.op_stride_a = (size_t)n * p.mld; p.op_stride_b = (size_t)m * p.nld;
    float *dMG, *dGL;
    GPAD_TRY(dev_alloc(h, &dMG, (size_t)B * p.op_stride_a));
    GPAD_TRY(dev_alloc(h, &dGL, (size_t)B * p.op_stride_b));
    const bool flipped = h->cfg.layout == GPAD_LAYOUT_FLIPPED;
    const size_t per = (size_t)n * m;
    if (h->cfg.operators_mem == GPAD_MEM_DEVICE) {
        GPAD_TRY(lat::launch_convert_ops(dMG, M_G, B, n, m, p.mld, flipped, nullptr));
        GPAD_TRY(lat::launch_convert_ops(dGL, G_L, B, m, n, p.nld, flipped, nullptr));
    } else {
        // stage through a bounded device buffer
        const int chunk = (int)std::max<size_t>(1, std::min<size_t>(B, (size_t)(256u << 20) / (per * sizeof(float))));
        float* tmp = nullptr;
        GPAD_CUDA(cudaMalloc(&tmp, (size_t)chunk * per * sizeof(float)));
        int rc = GPAD_OK;
        for (int b0 = 0; b0 < B && rc == GPAD_OK; b0 += chunk) {
            const int nb = std::min(chunk, B - b0);
            for (int which = 0; which < 2 && rc == GPAD_OK; ++which) {
                const float* src = (which == 0 ? M_G : G_L) + (size_t)b0 * per;
                if (cudaMemcpy(tmp, src, (size_t)nb * per * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { rc = GPAD_ERR_CUDA; break; }
                rc = which == 0 ? lat::launch_convert_ops(dMG + (size_t)b0 * p.op_stride_a, tmp, nb, n, m, p.mld, flipped, nullptr)
                                : lat::launch_convert_ops(dGL + (size_t)b0 * p.op_stride_b, tmp, nb, m, n, p.nld, flipped, nullptr);
                if (cudaDeviceSynchronize() != cudaSuccess) rc = GPAD_ERR_CUDA;
            }
        }
        cudaFree(tmp);
        if (rc != GPAD_OK) { set_error("per-instance operator upload failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    }
    GPAD_CUDA(cudaDeviceSynchronize());
    p.M_G = dMG; p.G_L = dGL;
    const size_t bn_ = (size_t)B * n, bm_ = (size_t)B * m;
    GPAD_TRY(dev_alloc(h, &h->pi_gP, bn_)); GPAD_TRY(dev_alloc(h, &h->pi_pD, bm_)); GPAD_TRY(dev_alloc(h, &h->pi_f, bn_));
    GPAD_TRY(dev_alloc(h, &h->pi_y0, bm_)); GPAD_TRY(dev_alloc(h, &h->pi_yprev0, bm_));
    GPAD_TRY(dev_alloc(h, &h->pi_ynext, bm_)); GPAD_TRY(dev_alloc(h, &h->pi_y, bm_)); GPAD_TRY(dev_alloc(h, &h->pi_w, bm_));
    GPAD_TRY(dev_alloc(h, &h->pi_z, bn_)); GPAD_TRY(dev_alloc(h, &h->pi_zhat, bn_));
    GPAD_TRY(dev_alloc(h, &h->pi_iters, B)); GPAD_TRY(dev_alloc(h, &h->pi_status, B));
    GPAD_TRY(dev_alloc(h, &h->pi_viol, B)); GPAD_TRY(dev_alloc(h, &h->pi_gap, B));
    GPAD_TRY(dev_alloc(h, &p.x_w, 4)); GPAD_TRY(dev_alloc(h, &p.x_zhat, 4)); GPAD_TRY(dev_alloc(h, &p.x_red, 4));
    GPAD_TRY(dev_alloc(h, &h->d_flags, 2));
    p.barrier = h->d_flags; p.nonfinite_flag = reinterpret_cast<int*>(h->d_flags + 1);
    char buf[320];
    snprintf(buf, sizeof(buf), "batch-per-instance: one CTA per QP (batched GEMV), %d threads, lanes/row %d|%d, operators %s, smem %zu B/CTA",
             plan.threads, 1 << p.lg_a, 1 << p.lg_b,
             plan.small ? "and per-row state read once into registers (lean kernel)" : plan.regs ? "read once into registers" : "in shared memory / streamed",
             plan.small ? lat::small_smem_bytes(p) : lat::smem_bytes(p, plan.regs));
    h->desc = buf;
    h->warp = lat::warp_supported(p);
    if (const char* e = getenv("GPAD_LATENCY_WARP")) h->warp = h->warp && atoi(e) != 0;
    if (h->warp) h->desc = "batch-per-instance: one WARP per QP (batched GEMV), 4 QPs per CTA, operators and state read once into registers, "
                           "no shared memory or block barrier in the loop (latency_warp.cu)";
    return GPAD_OK;
}

int solve_per_instance(gpad_handle_s* h, const gpad_solve_args_t* a) {
    const int n = h->n, m = h->cfg.m, B = a->batch;
    const bool host = a->mem == GPAD_MEM_HOST;
    cudaStream_t s = host ? h->own_stream : static_cast<cudaStream_t>(a->stream);
    GPAD_TRY(upload_schedule(h, a->theta, a->beta, a->max_iter, s));
    lat::Params p = h->lp;
    p.batch = B;
    const size_t bn_ = (size_t)B * n * sizeof(float), bm_ = (size_t)B * m * sizeof(float);
    if (host) {
        GPAD_CUDA(cudaMemcpyAsync(h->pi_gP, a->g_P, bn_, cudaMemcpyHostToDevice, s));
        GPAD_CUDA(cudaMemcpyAsync(h->pi_pD, a->p_D, bm_, cudaMemcpyHostToDevice, s));
        if (a->f) GPAD_CUDA(cudaMemcpyAsync(h->pi_f, a->f, bn_, cudaMemcpyHostToDevice, s));
        if (a->y0) GPAD_CUDA(cudaMemcpyAsync(h->pi_y0, a->y0, bm_, cudaMemcpyHostToDevice, s));
        if (a->y_prev0) GPAD_CUDA(cudaMemcpyAsync(h->pi_yprev0, a->y_prev0, bm_, cudaMemcpyHostToDevice, s));
        p.g_P = h->pi_gP; p.p_D = h->pi_pD; p.f = a->f ? h->pi_f : nullptr;
        p.y0 = a->y0 ? h->pi_y0 : nullptr; p.y_prev0 = a->y_prev0 ? h->pi_yprev0 : nullptr;
    } else {
        p.g_P = a->g_P; p.p_D = a->p_D; p.f = a->f; p.y0 = a->y0; p.y_prev0 = a->y_prev0;
    }
    p.theta = h->d_theta; p.beta = h->d_beta;
    p.max_iter = a->max_iter; p.check_every = a->check_every > 0 ? a->check_every : 0;
    p.eps_g = a->eps_g; p.eps_V = a->eps_V;
    const bool dev = !host;
    p.out_y_next = dev ? a->y_next : (a->y_next ? h->pi_ynext : nullptr);
    p.out_y = dev ? a->y : (a->y ? h->pi_y : nullptr);
    p.out_z = dev ? a->z : (a->z ? h->pi_z : nullptr);
    p.out_zhat = dev && a->zhat ? a->zhat : h->pi_zhat;       // also the scratch of the dual-gap branch
    p.out_w = dev && a->w ? a->w : h->pi_w;
    p.out_iters = dev ? a->iters : (a->iters ? h->pi_iters : nullptr);
    p.out_status = dev ? a->status : (a->status ? h->pi_status : nullptr);
    p.out_max_viol = dev ? a->max_viol : (a->max_viol ? h->pi_viol : nullptr);
    p.out_gap = dev ? a->gap : (a->gap ? h->pi_gap : nullptr);
    cudaEvent_t pe = h->prof_begin(s);
    if (h->warp && p.max_iter >= 1) {
        GPAD_TRY(lat::launch_warp(p, s));
    } else if (h->small) {
        p.sched_smem = round_up(std::min(a->max_iter, lat::small_sched_capacity()), 4);
        GPAD_TRY(lat::launch_small(p, h->cha, h->chb, 1, h->threads, s));
    } else {
        GPAD_TRY(lat::launch(p, lat::SYNC_BLOCK, h->ops_smem, 1, h->threads, s));
    }
    h->prof_end(0, pe, s);
    h->launches += 1;
    if (host) {
        if (a->y_next) GPAD_CUDA(cudaMemcpyAsync(a->y_next, h->pi_ynext, bm_, cudaMemcpyDeviceToHost, s));
        if (a->y) GPAD_CUDA(cudaMemcpyAsync(a->y, h->pi_y, bm_, cudaMemcpyDeviceToHost, s));
        if (a->w) GPAD_CUDA(cudaMemcpyAsync(a->w, h->pi_w, bm_, cudaMemcpyDeviceToHost, s));
        if (a->z) GPAD_CUDA(cudaMemcpyAsync(a->z, h->pi_z, bn_, cudaMemcpyDeviceToHost, s));
        if (a->zhat) GPAD_CUDA(cudaMemcpyAsync(a->zhat, h->pi_zhat, bn_, cudaMemcpyDeviceToHost, s));
        if (a->iters) GPAD_CUDA(cudaMemcpyAsync(a->iters, h->pi_iters, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
        if (a->status) GPAD_CUDA(cudaMemcpyAsync(a->status, h->pi_status, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
        if (a->max_viol) GPAD_CUDA(cudaMemcpyAsync(a->max_viol, h->pi_viol, sizeof(float) * B, cudaMemcpyDeviceToHost, s));
        if (a->gap) GPAD_CUDA(cudaMemcpyAsync(a->gap, h->pi_gap, sizeof(float) * B, cudaMemcpyDeviceToHost, s));
        GPAD_CUDA(cudaStreamSynchronize(s));
    }
    return GPAD_OK;
}

// ------------------------------------------------------------------ batch (shared operators)
int setup_batch(gpad_handle_s* h, const std::vector<float>& MG, const std::vector<float>& GL) {
    const int n = h->n, m = h->cfg.m;
    const bool tcp = h->cfg.precision == GPAD_PREC_TF32X3;
    BatchState& st = h->st;
    st.n = n; st.m = m; st.np = round_up(n, 32); st.mp = round_up(m, 32);
    // cta_group::1 by default: measured on B200 the CTA-pair kernel (GPAD_TC_CG=2) ties on product 2 (HBM-bound
    // epilogue) and loses on product 1 (cross-CTA hand-off of the in-kernel w split), see DESIGN.md 4.1
    int cg = 1;
    if (const char* e = getenv("GPAD_TC_CG")) cg = atoi(e) == 2 ? 2 : 1;
    st.Bp = round_up(h->cfg.max_batch, 256);
    int bn1 = 0, nt1 = 0, bn2 = 0, nt2 = 0;
    tc::plan_tiles(n, &bn1, &nt1);
    tc::plan_tiles(m, &bn2, &nt2);
    // product 1: second-generation kernel (P-formulation, A operand through TMEM) unless disabled / CTA pairs requested
    bool p1 = tcp && cg == 1;
    if (const char* e = getenv("GPAD_TC_P1")) p1 = p1 && atoi(e) != 0;
    if (const char* e = getenv("GPAD_TC_PFORM")) p1 = p1 && atoi(e) != 0;
    if (p1 && !getenv("GPAD_TC_P1")) {
        // the TMEM-fed kernel costs ~824 clk per 128-row tile and k-block whatever the tile width (<= 208 columns); the
        // first-generation kernel ~940 clk at 208 columns, growing with the width (<= 256) -- but it may need fewer tiles.
        // With few tiles the number of waves over the SMs decides (battery (10,100), 4096 QPs: 160 tiles = 2 waves against
        // 128 tiles = 1 wave: measured 98.7 k against 112 k solves/s), with many tiles the per-tile cost does.
        int bn_ts = 0, nt_ts = 0;
        tc::plan_tiles_p1(n, &bn_ts, &nt_ts);
        const int mt = (h->cfg.max_batch + 127) / 128;
        const double cost_ts = std::ceil((double)mt * nt_ts / h->num_sms) * 824.0;
        const double cost_ss = std::ceil((double)mt * nt1 / h->num_sms) * 940.0 * bn1 / 208.0;
        p1 = cost_ts <= cost_ss;
    }
    int step1 = 0;
    if (p1) tc::plan_tiles_p1(n, &bn1, &nt1, &step1);
    bool p2ts = false;                           // product 2 through the same TMEM-A kernel (GPAD_TC_P2TS=1)
    if (const char* e = getenv("GPAD_TC_P2TS")) p2ts = p1 && atoi(e) != 0;
    if (p2ts) tc::plan_tiles_p1(m, &bn2, &nt2);
    h->op.n_rows_pad = round_up(std::max(bn1 * nt1, n), 128);
    h->op.m_rows_pad = round_up(m + 256, 128);      // any product-2 tiling of width <= 256 stays inside (setup-time autotuning)
    GPAD_TRY(upload_padded(h, MG.data(), n, m, h->op.n_rows_pad, st.mp, &h->op.M_G));
    GPAD_TRY(upload_padded(h, GL.data(), m, n, h->op.m_rows_pad, st.np, &h->op.G_L));
    const size_t bm = (size_t)st.Bp * st.mp, bnn = (size_t)st.Bp * st.np;
    GPAD_TRY(dev_alloc(h, &st.g_P, bnn)); GPAD_TRY(dev_alloc(h, &st.p_D, bm)); GPAD_TRY(dev_alloc(h, &st.f, bnn));
    GPAD_TRY(dev_alloc(h, &st.yb[0], bm)); GPAD_TRY(dev_alloc(h, &st.yb[1], bm)); GPAD_TRY(dev_alloc(h, &st.yb[2], bm));
    GPAD_TRY(dev_alloc(h, &st.z, bnn)); GPAD_TRY(dev_alloc(h, &st.zhat, bnn)); GPAD_TRY(dev_alloc(h, &st.sbar, bm));
    GPAD_TRY(dev_alloc(h, &st.red, (size_t)st.Bp * kRedStride));
    GPAD_TRY(dev_alloc(h, &st.done, st.Bp)); GPAD_TRY(dev_alloc(h, &st.iters, st.Bp)); GPAD_TRY(dev_alloc(h, &st.status, st.Bp));
    GPAD_TRY(dev_alloc(h, &st.max_viol, st.Bp)); GPAD_TRY(dev_alloc(h, &st.gap, st.Bp));
    // the padding columns of the row-padded arrays are never written by a kernel; the DMA (cudaMemcpy2D) input path of
    // pipelined host solves relies on them being zero
    GPAD_CUDA(cudaMemset(st.g_P, 0, bnn * sizeof(float))); GPAD_CUDA(cudaMemset(st.p_D, 0, bm * sizeof(float)));
    GPAD_CUDA(cudaMemset(st.f, 0, bnn * sizeof(float)));
    for (int k = 0; k < 3; ++k) GPAD_CUDA(cudaMemset(st.yb[k], 0, bm * sizeof(float)));
    GPAD_TRY(dev_alloc(h, &st.active_count, 2));
    GPAD_TRY(dev_alloc(h, &st.need, st.Bp)); GPAD_TRY(dev_alloc(h, &st.zy, bnn));
    GPAD_CUDA(cudaMemset(st.zy, 0, bnn * sizeof(float)));
    GPAD_TRY(dev_alloc(h, &h->stage_in, (size_t)h->cfg.max_batch * std::max(n, m)));
    GPAD_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->h_active), 2 * sizeof(int)));
    char buf[640];
    if (tcp) {
        GPAD_TRY(dev_alloc(h, &st.zh_hi, bnn)); GPAD_TRY(dev_alloc(h, &st.zh_lo, bnn));
        GPAD_TRY(dev_alloc(h, &st.Pb[0], bnn)); GPAD_TRY(dev_alloc(h, &st.Pb[1], bnn));
        GPAD_CUDA(cudaMemset(st.zh_hi, 0, bnn * sizeof(float))); GPAD_CUDA(cudaMemset(st.zh_lo, 0, bnn * sizeof(float)));
        GPAD_CUDA(cudaMemset(st.zhat, 0, bnn * sizeof(float)));    // its K-padding columns feed product 2 when it stages zhat itself
        const size_t c1 = (size_t)h->op.n_rows_pad * st.mp, c2 = (size_t)h->op.m_rows_pad * st.np;
        GPAD_TRY(dev_alloc(h, &h->op.M_G_lo, c1)); GPAD_TRY(dev_alloc(h, &h->op.G_L_lo, c2));
        GPAD_TRY(tc::launch_split(h->op.M_G, h->op.M_G, h->op.M_G_lo, c1, nullptr));
        GPAD_TRY(tc::launch_split(h->op.G_L, h->op.G_L, h->op.G_L_lo, c2, nullptr));
        GPAD_CUDA(cudaDeviceSynchronize());
        int bk = 16;
        if (const char* e = getenv("GPAD_TC_BK")) bk = atoi(e) == 32 ? 32 : 16;
        tc::GemmDesc& g1 = h->g1; tc::GemmDesc& g2 = h->g2;
        g1.bk = g2.bk = bk;
        g1.cg = g2.cg = cg;
        g1.mc = g2.mc = cg == 1 ? tc_multicast() : 1;
        const int bdiv = cg * g1.mc;                 // operator rows per TMA box = bn / bdiv
        g1.k_pad = st.mp; g1.bn = bn1; g1.n_tiles = nt1; g1.ncols_valid = n;
        g2.k_pad = st.np; g2.bn = bn2; g2.n_tiles = nt2; g2.ncols_valid = m;
        g1.stages = cg == 2 ? tc::pick_stages2(bk, bn1, h->smem_optin) : tc::pick_stages(bk, bn1, h->smem_optin);
        g2.stages = cg == 2 ? tc::pick_stages2(bk, bn2, h->smem_optin) : tc::pick_stages(bk, bn2, h->smem_optin);
        if (p1) {
            g1.p1 = 1; g1.bk = 16; g1.mc = 1; g1.step = step1;
            GPAD_TRY(tc::plan_rings_p1(1, bn1, h->smem_optin, &g1.a_stages, &g1.stages));
        }
        if (p2ts) {
            g2.p1 = 1; g2.bk = 16; g2.mc = 1;
            GPAD_TRY(tc::plan_rings_p1(2, bn2, h->smem_optin, &g2.a_stages, &g2.stages));
        }
        if (const char* e = getenv("GPAD_TC_STAGES")) { g1.stages = std::min(g1.stages, std::max(2, atoi(e))); g2.stages = std::min(g2.stages, std::max(2, atoi(e))); }
        for (int k = 0; k < 3; ++k) GPAD_TRY(tc::make_tmap(&g1.tmY[k], st.yb[k], st.mp, st.Bp, st.mp, g1.bk, 128));
        GPAD_TRY(tc::make_tmap(&g1.tmB_hi, h->op.M_G, st.mp, h->op.n_rows_pad, st.mp, g1.bk, p1 ? bn1 : bn1 / bdiv));
        GPAD_TRY(tc::make_tmap(&g1.tmB_lo, h->op.M_G_lo, st.mp, h->op.n_rows_pad, st.mp, g1.bk, p1 ? bn1 : bn1 / bdiv));
        g2.xf2 = 0;      // measured slower (0.72 -> 0.81 ms): the in-place split adds shared-memory traffic to a kernel bound by it
        if (const char* e = getenv("GPAD_TC_XF2")) g2.xf2 = cg == 1 && atoi(e) != 0;
        if (p2ts) g2.xf2 = 1;                        // zhat is staged as one fp32 tile and split in registers
        GPAD_TRY(tc::make_tmap(&g2.tmA_hi, g2.xf2 ? st.zhat : st.zh_hi, st.np, st.Bp, st.np, g2.bk, 128));
        GPAD_TRY(tc::make_tmap(&g2.tmA_lo, st.zh_lo, st.np, st.Bp, st.np, bk, 128));
        GPAD_TRY(tc::make_tmap(&g2.tmB_hi, h->op.G_L, st.np, h->op.m_rows_pad, st.np, g2.bk, p2ts ? bn2 : bn2 / bdiv));
        GPAD_TRY(tc::make_tmap(&g2.tmB_lo, h->op.G_L_lo, st.np, h->op.m_rows_pad, st.np, g2.bk, p2ts ? bn2 : bn2 / bdiv));
        // ---- product 2 tile width: measured, not modelled (64K quadrotor batch, ms per launch: 240 -> 0.73, 208 -> 1.04,
        // 192 -> 0.69, 160 -> 0.66, 128 -> 0.81), so a few widths are timed on this handle's own
        // buffers (3 launches each on the zeroed state) and the fastest is kept.  Results do not depend on the width: every
        // output element sums over K in the same order.
        auto config_g2 = [&](int b) -> int {
            bn2 = b; nt2 = (m + b - 1) / b;
            g2.bn = bn2; g2.n_tiles = nt2;
            g2.stages = tc::pick_stages(bk, bn2, h->smem_optin);
            if (const char* e = getenv("GPAD_TC_STAGES")) g2.stages = std::min(g2.stages, std::max(2, atoi(e)));
            GPAD_TRY(tc::make_tmap(&g2.tmB_hi, h->op.G_L, st.np, h->op.m_rows_pad, st.np, g2.bk, bn2 / bdiv));
            GPAD_TRY(tc::make_tmap(&g2.tmB_lo, h->op.G_L_lo, st.np, h->op.m_rows_pad, st.np, g2.bk, bn2 / bdiv));
            return GPAD_OK;
        };
        bool tune = cg == 1 && !p2ts && !g2.xf2 && h->cfg.max_batch >= 1024;
        if (const char* e = getenv("GPAD_TC_AUTOTUNE")) tune = tune && atoi(e) != 0;
        if (const char* e = getenv("GPAD_TC_BN2")) {       // explicit width (multiple of 16, <= 256)
            if (cg == 1 && !p2ts) GPAD_TRY(config_g2(std::max(16, std::min(256, atoi(e) / 16 * 16))));
            tune = false;
        }
        if (tune) {
            const int bn_default = bn2;
            BatchKernelArgs k{};
            k.n = n; k.m = m; k.np = st.np; k.mp = st.mp; k.B = h->cfg.max_batch; k.L = h->cfg.L;
            k.g_P = st.g_P; k.p_D = st.p_D; k.z = st.z; k.zhat = st.zhat; k.zh_hi = st.zh_hi; k.zh_lo = st.zh_lo;
            k.sbar = st.sbar; k.red = st.red;
            k.y_prev = st.yb[2]; k.y_cur = st.yb[0]; k.y_next = st.yb[1];
            k.it.theta = 1.f; k.it.beta = 0.f;
            cudaEvent_t e0, e1;
            GPAD_CUDA(cudaEventCreate(&e0)); GPAD_CUDA(cudaEventCreate(&e1));
            float best_ms = 1e30f; int best_bn = bn_default;
            std::vector<int> cand = {bn_default};
            // multiples of 32 columns: every 32-column block of the epilogue then starts on a 128-byte line (widths that
            // are only multiples of 16 -- 144, 176, 208, 240 -- measured 0.73 .. 1.04 ms against 0.66 .. 0.69 for 160 / 192)
            for (int b : {256, 224, 192, 160, 128}) if (b != bn_default && (m + b - 1) / b <= 64) cand.push_back(b);
            for (int b : cand) {
                if (config_g2(b) != GPAD_OK) continue;
                g2.m_tiles = (h->cfg.max_batch + 127) / 128;
                float ms = 1e30f;
                bool ok = true;
                for (int rep = 0; rep < 4 && ok; ++rep) {          // first launch untimed
                    if (rep == 1) cudaEventRecord(e0, nullptr);
                    ok = tc::launch_gemm(2, g2, k, nullptr, 0, h->num_sms, nullptr) == GPAD_OK;
                }
                cudaEventRecord(e1, nullptr);
                if (cudaEventSynchronize(e1) != cudaSuccess || !ok) { cudaGetLastError(); continue; }
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best_ms) { best_ms = ms; best_bn = b; }
            }
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            GPAD_TRY(config_g2(best_bn));
            GPAD_CUDA(cudaMemset(st.yb[1], 0, bm * sizeof(float)));      // the timed launches wrote y_next
        }
        if (p1)
            snprintf(buf, sizeof(buf),
                     "batch-shared: tcgen05 cta_group::1 kind::tf32 x3; product1 = P-formulation (A = y_v only, split in registers, A operand "
                     "through a TMEM ring, state ring %d x 8 KB + operator ring %d stages), tiles 128x%d x%d; product2 tiles 128x%d x%d "
                     "(%d stages, TMA ring bk=%d, operator multicast x%d), TMEM 512 cols, persistent over %d SMs",
                     g1.a_stages, g1.stages, bn1, nt1, bn2, nt2, g2.stages, bk, g2.mc, h->num_sms);
        else
        snprintf(buf, sizeof(buf),
                 "batch-shared: tcgen05 cta_group::%d kind::tf32 x3 (hi/lo split, w built in-kernel), TMA ring bk=%d (operator tiles multicast x%d), product1 tiles "
                 "%dx%d x%d (%d stages), product2 tiles %dx%d x%d (%d stages), TMEM 2x256 cols, persistent over %d SMs",
                 cg, bk, g1.mc, 128 * cg, bn1, nt1, g1.stages, 128 * cg, bn2, nt2, g2.stages, h->num_sms);
    } else {
        snprintf(buf, sizeof(buf), "batch-shared: CUDA-core fp32 GEMM 128x128x16 tiles with fused GPAD epilogues");
    }
    h->desc = buf;
    return GPAD_OK;
}

// ---- a contiguous range of instances [b0, b0 + B) of the batch state: the same arrays, offset, with their own
// tensor maps.  The whole batch is one view; host-memory fixed-iteration solves are split into several views so that
// the PCIe copies of one range overlap the iterations of another (solve_batch_pipelined).
struct BatchView {
    BatchState st;             // pointers offset to row b0, B = rows in the range, Bp = rows rounded up to 256
    tc::GemmDesc g1, g2;
    int b0 = 0;
};

int make_view(gpad_handle_s* h, int b0, int B, BatchView& v) {
    const BatchState& s = h->st;
    const bool tcp = h->cfg.precision == GPAD_PREC_TF32X3;
    v.st = s; v.g1 = h->g1; v.g2 = h->g2; v.b0 = b0;
    v.st.B = B;
    if (b0 == 0 && round_up(B, 256) == s.Bp) return GPAD_OK;          // the whole allocation: the maps made at setup
    const size_t om = (size_t)b0 * s.mp, on = (size_t)b0 * s.np;
    BatchState& t = v.st;
    t.Bp = std::min(round_up(B, 256), s.Bp - b0);
    t.g_P += on; t.p_D += om; t.f += on; t.z += on; t.zhat += on; t.sbar += om;
    for (int k = 0; k < 3; ++k) t.yb[k] += om;
    if (t.zh_hi) { t.zh_hi += on; t.zh_lo += on; t.Pb[0] += on; t.Pb[1] += on; }
    t.zy += on;
    t.red += (size_t)b0 * kRedStride; t.done += b0; t.need += b0; t.iters += b0; t.status += b0; t.max_viol += b0; t.gap += b0;
    if (tcp) {
        for (int k = 0; k < 3; ++k) GPAD_TRY(tc::make_tmap(&v.g1.tmY[k], t.yb[k], s.mp, t.Bp, s.mp, v.g1.bk, 128));
        GPAD_TRY(tc::make_tmap(&v.g2.tmA_hi, v.g2.xf2 ? t.zhat : t.zh_hi, s.np, t.Bp, s.np, v.g2.bk, 128));
        GPAD_TRY(tc::make_tmap(&v.g2.tmA_lo, t.zh_lo, s.np, t.Bp, s.np, v.g2.bk, 128));
    }
    return GPAD_OK;
}

// user vector rows [b0, b0 + B) of [.][len] (host or device) -> padded device rows [Bp][ld]; null src -> zeros.
// stage == nullptr with host data: strided DMA copy straight into the padded rows (no SM work in the copy path)
int ingest(gpad_handle_s* h, const BatchView& v, float* dst, int ld, const float* src, int len, bool host, float* stage, cudaStream_t s) {
    const int B = v.st.B;
    const float* dsrc = src ? src + (size_t)v.b0 * len : nullptr;
    if (host && !stage) {
        if (dsrc) GPAD_CUDA(cudaMemcpy2DAsync(dst, sizeof(float) * ld, dsrc, sizeof(float) * len, sizeof(float) * len, B, cudaMemcpyHostToDevice, s));
        else GPAD_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)B * ld, s));
        return GPAD_OK;
    }
    if (src && host) {
        GPAD_CUDA(cudaMemcpyAsync(stage, dsrc, sizeof(float) * (size_t)B * len, cudaMemcpyHostToDevice, s));
        dsrc = stage;
    }
    GPAD_TRY(launch_pad_rows(dst, ld, v.st.Bp, dsrc, len, B, s));
    h->launches += 1;
    return GPAD_OK;
}

int emit(gpad_handle_s* h, const BatchView& v, float* dst, int len, const float* src, int ld, bool host, float* stage, cudaStream_t s) {
    if (!dst) return GPAD_OK;
    const int B = v.st.B;
    dst += (size_t)v.b0 * len;
    if (host && !stage) {
        GPAD_CUDA(cudaMemcpy2DAsync(dst, sizeof(float) * len, src, sizeof(float) * ld, sizeof(float) * len, B, cudaMemcpyDeviceToHost, s));
        return GPAD_OK;
    }
    float* ddst = host ? stage : dst;
    GPAD_TRY(launch_unpad_rows(ddst, len, B, src, ld, s));
    h->launches += 1;
    if (host) GPAD_CUDA(cudaMemcpyAsync(dst, ddst, sizeof(float) * (size_t)B * len, cudaMemcpyDeviceToHost, s));
    return GPAD_OK;
}

int view_inputs(gpad_handle_s* h, const BatchView& v, const gpad_solve_args_t* a, bool host, float* stage, cudaStream_t s) {
    const BatchState& st = v.st;
    GPAD_TRY(ingest(h, v, st.g_P, st.np, static_cast<const float*>(a->g_P), st.n, host, stage, s));
    GPAD_TRY(ingest(h, v, st.p_D, st.mp, static_cast<const float*>(a->p_D), st.m, host, stage, s));
    if (a->f) GPAD_TRY(ingest(h, v, st.f, st.np, static_cast<const float*>(a->f), st.n, host, stage, s));
    GPAD_TRY(ingest(h, v, st.yb[0], st.mp, static_cast<const float*>(a->y0), st.m, host, stage, s));         // y_0
    GPAD_TRY(ingest(h, v, st.yb[2], st.mp, static_cast<const float*>(a->y_prev0), st.m, host, stage, s));    // y_{-1}
    return GPAD_OK;
}

// the iterations of one view (everything between the input and the output copies)
int view_iterate(gpad_handle_s* h, BatchView& v, const gpad_solve_args_t* a, cudaStream_t s) {
    BatchState& st = v.st;
    const int n = st.n, m = st.m, B = st.B;
    const bool tcp = h->cfg.precision == GPAD_PREC_TF32X3;
    const bool checking = a->check_every > 0;
    GPAD_TRY(launch_batch_init(st, checking, s));
    GPAD_TRY(launch_batch_reset_term(st, a->max_iter, s));
    h->launches += 2;

    BatchKernelArgs k{};
    k.n = n; k.m = m; k.np = st.np; k.mp = st.mp; k.B = B; k.checking = checking ? 1 : 0; k.L = h->cfg.L;
    k.g_P = st.g_P; k.p_D = st.p_D; k.f = a->f ? st.f : nullptr;
    k.z = st.z; k.zhat = st.zhat; k.zh_hi = st.zh_hi; k.zh_lo = st.zh_lo;
    k.sbar = st.sbar; k.red = st.red; k.done = checking ? st.done : nullptr;
    k.prefetch = 0;   // measured slower on B200 (product 2: 0.71 -> 0.90 ms): the SM ingest path is the bound, extra requests cost
    if (const char* e = getenv("GPAD_TC_PREFETCH")) k.prefetch = atoi(e) != 0;
    const int tile_rows = (tcp && v.g1.cg == 2) ? 256 : 128;
    const int m_tiles = round_up(B, tile_rows) / tile_rows;
    v.g1.m_tiles = m_tiles; v.g2.m_tiles = m_tiles;
    const int Bp_call = m_tiles * tile_rows;
    // P-formulation of product 1 (cta_group::1 tcgen05 kernel): P_v = M_G y_v, M_G w_v = P_v + beta_v (P_v - P_{v-1})
    k.pform = (tcp && v.g1.cg == 1) ? 1 : 0;
    if (const char* e = getenv("GPAD_TC_PFORM")) k.pform = k.pform && atoi(e) != 0;
    if (v.g1.p1) k.pform = 1;
    k.zh_single = (tcp && v.g2.xf2) ? 1 : 0;
    if (k.pform && a->max_iter > 0) {
        if (a->y_prev0) {          // warm start: P_{-1} = M_G y_{-1} (one extra product-1 launch)
            BatchKernelArgs kp = k;
            kp.p_only = 1; kp.P_cur = st.Pb[1]; kp.P_prev = st.Pb[0];
            v.g1.tmA_hi = v.g1.tmY[2]; v.g1.tmA_lo = v.g1.tmY[2];
            GPAD_TRY(v.g1.p1 ? tc::launch_p1(1, v.g1, kp, h->num_sms, s) : tc::launch_gemm(1, v.g1, kp, nullptr, 0, h->num_sms, s));
            h->launches += 1;
        } else {
            GPAD_CUDA(cudaMemsetAsync(st.Pb[1], 0, sizeof(float) * (size_t)st.Bp * st.np, s));
        }
    }

    for (int it = 0; it < a->max_iter; ++it) {
        const bool check = checking && ((it + 1) % a->check_every == 0);
        k.it.theta = a->theta[it];
        k.it.beta = a->beta[it];
        k.it.check = check ? 1 : 0;
        k.it.store_zhat = (checking || it + 1 == a->max_iter) ? 1 : 0;
        k.y_prev = st.yb[(it + 2) % 3];           // y_{v-1}
        k.y_cur = st.yb[it % 3];                  // y_v
        k.y_next = st.yb[(it + 1) % 3];           // y_{v+1} overwrites y_{v-2}
        k.P_cur = st.Pb[it & 1]; k.P_prev = st.Pb[(it + 1) & 1];
        if (tcp) { v.g1.tmA_hi = v.g1.tmY[it % 3]; v.g1.tmA_lo = v.g1.tmY[(it + 2) % 3]; }
        if (tcp) {
            cudaEvent_t pe = h->prof_begin(s);
            GPAD_TRY(v.g1.p1 ? tc::launch_p1(1, v.g1, k, h->num_sms, s) : tc::launch_gemm(1, v.g1, k, nullptr, 0, h->num_sms, s));
            h->prof_end(1, pe, s);
            pe = h->prof_begin(s);
            GPAD_TRY(v.g2.p1 ? tc::launch_p1(2, v.g2, k, h->num_sms, s) : tc::launch_gemm(2, v.g2, k, nullptr, 0, h->num_sms, s));
            h->prof_end(2, pe, s);
        } else {
            cudaEvent_t pe = h->prof_begin(s);
            GPAD_TRY(launch_simt_product(1, h->op, k, Bp_call, s));
            h->prof_end(1, pe, s);
            pe = h->prof_begin(s);
            GPAD_TRY(launch_simt_product(2, h->op, k, Bp_call, s));
            h->prof_end(2, pe, s);
        }
        h->launches += 2;
        if (check) {
            GPAD_TRY(launch_batch_decide(st, it + 1, h->cfg.L, a->eps_g, a->eps_V, a->f != nullptr, s));
            h->launches += 1;
            GPAD_CUDA(cudaMemcpyAsync(h->h_active, st.active_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
            GPAD_CUDA(cudaStreamSynchronize(s));
            if (a->f && h->h_active[1] > 0) {
                // dual-gap branch: z_y = M_G y_{v+1} - g_P and G_L z_y for the flagged instances (two more products)
                BatchKernelArgs kd = k;
                kd.dual = 1; kd.need = st.need; kd.zy = st.zy; kd.p_only = 0;
                kd.it.check = 1; kd.it.beta = 0.f; kd.it.theta = 0.f; kd.it.store_zhat = 0;
                kd.y_cur = k.y_next; kd.y_prev = k.y_next;      // beta = 0: w = y_{v+1}
                if (tcp) { v.g1.tmA_hi = v.g1.tmY[(it + 1) % 3]; v.g1.tmA_lo = v.g1.tmY[(it + 1) % 3]; }
                if (tcp) {
                    GPAD_TRY(v.g1.p1 ? tc::launch_p1(1, v.g1, kd, h->num_sms, s) : tc::launch_gemm(1, v.g1, kd, nullptr, 0, h->num_sms, s));
                    GPAD_TRY(tc::launch_gemm(2, v.g2, kd, nullptr, 0, h->num_sms, s));
                } else {
                    GPAD_TRY(launch_simt_product(1, h->op, kd, Bp_call, s));
                    GPAD_TRY(launch_simt_product(2, h->op, kd, Bp_call, s));
                }
                GPAD_TRY(launch_batch_decide_dual(st, it + 1, h->cfg.L, a->eps_V, s));
                h->launches += 3;
                GPAD_CUDA(cudaMemcpyAsync(h->h_active, st.active_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
                GPAD_CUDA(cudaStreamSynchronize(s));
            }
            if (*h->h_active <= 0) break;
        }
    }
    if (!checking && a->max_iter > 0) {
        GPAD_TRY(launch_batch_finite(st, st.yb[a->max_iter % 3], s));
        h->launches += 1;
    }
    return GPAD_OK;
}

int view_outputs(gpad_handle_s* h, const BatchView& v, const gpad_solve_args_t* a, bool host, float* stage, cudaStream_t s, bool dma = false) {
    const BatchState& st = v.st;
    const int n = st.n, m = st.m, B = st.B;
    float* yo[3] = {static_cast<float*>(a->y_next), static_cast<float*>(a->y), static_cast<float*>(a->w)};
    if (host && dma) {
        // fixed iteration count I for every instance: y_I sits in yb[I % 3], y_{I-1} in yb[(I + 2) % 3]: strided DMA copies;
        // only w_{I-1} needs arithmetic (one kernel into the staging buffer)
        const int I = a->max_iter;
        if (yo[0]) GPAD_TRY(emit(h, v, yo[0], m, st.yb[I % 3], st.mp, true, nullptr, s));
        if (yo[1]) GPAD_TRY(emit(h, v, yo[1], m, st.yb[(I + 2) % 3], st.mp, true, nullptr, s));
        if (yo[2]) {
            GPAD_TRY(launch_unpad_y(nullptr, nullptr, stage, m, B, st.yb[0], st.yb[1], st.yb[2], st.mp, st.iters, h->d_beta, s));
            GPAD_CUDA(cudaMemcpyAsync(yo[2] + (size_t)v.b0 * m, stage, sizeof(float) * (size_t)B * m, cudaMemcpyDeviceToHost, s));
            h->launches += 1;
        }
        GPAD_TRY(emit(h, v, static_cast<float*>(a->z), n, st.z, st.np, true, nullptr, s));
        GPAD_TRY(emit(h, v, static_cast<float*>(a->zhat), n, st.zhat, st.np, true, nullptr, s));
    } else if (host) {      // host mode stages one vector at a time
        for (int k3 = 0; k3 < 3; ++k3) {
            if (!yo[k3]) continue;
            GPAD_TRY(launch_unpad_y(k3 == 0 ? stage : nullptr, k3 == 1 ? stage : nullptr, k3 == 2 ? stage : nullptr,
                                    m, B, st.yb[0], st.yb[1], st.yb[2], st.mp, st.iters, h->d_beta, s));
            GPAD_CUDA(cudaMemcpyAsync(yo[k3] + (size_t)v.b0 * m, stage, sizeof(float) * (size_t)B * m, cudaMemcpyDeviceToHost, s));
            h->launches += 1;
        }
    } else if (yo[0] || yo[1] || yo[2]) {
        const size_t o = (size_t)v.b0 * m;
        GPAD_TRY(launch_unpad_y(yo[0] ? yo[0] + o : nullptr, yo[1] ? yo[1] + o : nullptr, yo[2] ? yo[2] + o : nullptr, m, B,
                                st.yb[0], st.yb[1], st.yb[2], st.mp, st.iters, h->d_beta, s));
        h->launches += 1;
    }
    if (!(host && dma)) {
        GPAD_TRY(emit(h, v, static_cast<float*>(a->z), n, st.z, st.np, host, stage, s));
        GPAD_TRY(emit(h, v, static_cast<float*>(a->zhat), n, st.zhat, st.np, host, stage, s));
    }
    const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (a->iters) GPAD_CUDA(cudaMemcpyAsync(static_cast<int*>(a->iters) + v.b0, st.iters, sizeof(int) * B, kind, s));
    if (a->status) GPAD_CUDA(cudaMemcpyAsync(static_cast<int*>(a->status) + v.b0, st.status, sizeof(int) * B, kind, s));
    if (a->max_viol) GPAD_CUDA(cudaMemcpyAsync(static_cast<float*>(a->max_viol) + v.b0, st.max_viol, sizeof(float) * B, kind, s));
    if (a->gap) GPAD_CUDA(cudaMemcpyAsync(static_cast<float*>(a->gap) + v.b0, st.gap, sizeof(float) * B, kind, s));
    return GPAD_OK;
}

// host buffers, fixed iteration count, a large batch: ranges of `chunk` instances go through copy-in / iterate /
// copy-out on three streams, so the PCIe traffic of one range hides behind the iterations of its neighbours
int solve_batch_pipelined(gpad_handle_s* h, const gpad_solve_args_t* a, int chunk) {
    const int B = a->batch, C = (B + chunk - 1) / chunk;
    if (!h->stream_in) {
        // high priority: their short pad / unpad kernels must get SMs at the next boundary between the persistent GEMM
        // kernels of the compute stream instead of queueing behind all of them
        int lo = 0, hi = 0;
        GPAD_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        GPAD_CUDA(cudaStreamCreateWithPriority(&h->stream_in, cudaStreamNonBlocking, hi));
        GPAD_CUDA(cudaStreamCreateWithPriority(&h->stream_out, cudaStreamNonBlocking, hi));
        GPAD_TRY(dev_alloc(h, &h->stage_out, (size_t)h->cfg.max_batch * std::max(h->st.n, h->st.m)));
    }
    cudaStream_t s_in = h->stream_in, s_comp = h->own_stream, s_out = h->stream_out;
    GPAD_TRY(upload_schedule(h, a->theta, a->beta, a->max_iter, s_comp));
    GPAD_CUDA(cudaStreamSynchronize(s_comp));
    std::vector<cudaEvent_t> ev_in(C), ev_comp(C);
    for (int c = 0; c < C; ++c) {
        GPAD_CUDA(cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming));
        GPAD_CUDA(cudaEventCreateWithFlags(&ev_comp[c], cudaEventDisableTiming));
    }
    int rc = GPAD_OK;
    std::vector<BatchView> views(C);
    for (int c = 0; c < C && rc == GPAD_OK; ++c) rc = make_view(h, c * chunk, std::min(chunk, B - c * chunk), views[c]);
    // enqueue order matters: the ~200 launches of one range can fill the launch queue and block the host, so the input
    // copies of the NEXT range are queued before this range's iterations
    auto queue_inputs = [&](int c) {
        int r = view_inputs(h, views[c], a, true, nullptr, s_in);
        cudaEventRecord(ev_in[c], s_in);
        return r;
    };
    if (rc == GPAD_OK) rc = queue_inputs(0);
    for (int c = 0; c < C && rc == GPAD_OK; ++c) {
        if (c + 1 < C && (rc = queue_inputs(c + 1)) != GPAD_OK) break;
        cudaStreamWaitEvent(s_comp, ev_in[c], 0);
        if ((rc = view_iterate(h, views[c], a, s_comp)) != GPAD_OK) break;
        cudaEventRecord(ev_comp[c], s_comp);
        cudaStreamWaitEvent(s_out, ev_comp[c], 0);
        if ((rc = view_outputs(h, views[c], a, true, h->stage_out, s_out, true)) != GPAD_OK) break;
    }
    cudaError_t e1 = cudaStreamSynchronize(s_in), e2 = cudaStreamSynchronize(s_comp), e3 = cudaStreamSynchronize(s_out);
    for (int c = 0; c < C; ++c) { cudaEventDestroy(ev_in[c]); cudaEventDestroy(ev_comp[c]); }
    if (rc != GPAD_OK) return rc;
    GPAD_CUDA(e1); GPAD_CUDA(e2); GPAD_CUDA(e3);
    return GPAD_OK;
}

int solve_batch(gpad_handle_s* h, const gpad_solve_args_t* a) {
    const bool host = a->mem == GPAD_MEM_HOST;
    const bool checking = a->check_every > 0;
    if (host && !checking) {
        // opt-in (GPAD_HOST_CHUNK=<instances per range>, multiple of 256).  Measured on B200, 64K quadrotor batch with
        // 2.8 GB of copies per solve: one pass 354 k solves/s; 2 / 4 / 8 ranges 353 / 347 / 311 k -- the copies do overlap
        // (a 1.9 GB copy next to a full solve costs 1.4 ms), but 32K / 16K / 8K-instance solves are 9 / 19 / 44 % less
        // efficient per instance (wave quantisation of the persistent kernels, fixed cost per launch), which cancels it.
        if (const char* e = getenv("GPAD_HOST_CHUNK")) {
            const int chunk = std::max(256, atoi(e) / 256 * 256);
            if (a->batch >= 2 * chunk) return solve_batch_pipelined(h, a, chunk);
        }
    }
    cudaStream_t s = host ? h->own_stream : static_cast<cudaStream_t>(a->stream);
    BatchView v;
    GPAD_TRY(make_view(h, 0, a->batch, v));
    GPAD_TRY(view_inputs(h, v, a, host, h->stage_in, s));
    GPAD_TRY(upload_schedule(h, a->theta, a->beta, a->max_iter, s));    // beta[] on the device for the w output
    GPAD_TRY(view_iterate(h, v, a, s));
    GPAD_TRY(view_outputs(h, v, a, host, h->stage_in, s));
    if (host) GPAD_CUDA(cudaStreamSynchronize(s));
    return GPAD_OK;
}

}  // namespace

// =================================================================== C ABI
extern "C" {

const char* gpad_status_string(int status) {
    switch (status) {
        case GPAD_OK: return "ok";
        case GPAD_ERR_INVALID_ARG: return "invalid argument";
        case GPAD_ERR_CUDA: return "CUDA error";
        case GPAD_ERR_UNSUPPORTED: return "unsupported configuration";
        case GPAD_ERR_ALLOC: return "allocation failed";
        case GPAD_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
        case GPAD_ERR_IO: return "I/O error";
        default: return "unknown status";
    }
}

const char* gpad_last_error(void) { return g_last_error.c_str(); }
int gpad_api_version(void) { return GPAD_API_VERSION; }

int gpad_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    int usable = 0;
    for (int d = 0; d < count; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++usable;
    }
    return usable;
}

static int stream_of(void* s, cudaStream_t* out) { *out = static_cast<cudaStream_t>(s); return GPAD_OK; }

int gpad_step_one(const float* y, const float* y_prev, float* w, float beta, int m, void* stream) {
    GPAD_REQUIRE(y && y_prev && w && m > 0, "gpad_step_one: null pointer or m <= 0");
    cudaStream_t s; stream_of(stream, &s);
    return launch_step_one(y, y_prev, w, beta, m, s);
}

int gpad_step_two(const float* M_G, const float* w_v, const float* g_P, float* zhat, int N, int n_u, int m, void* stream) {
    GPAD_REQUIRE(M_G && w_v && g_P && zhat && N > 0 && n_u > 0 && m > 0, "gpad_step_two: bad argument");
    cudaStream_t s; stream_of(stream, &s);
    return launch_gemv_t(M_G, w_v, m, N * n_u, 0, g_P, nullptr, zhat, s);      // flipped M_G: [m][n]
}

int gpad_array_copy(float* dest, const float* src, int size, void* stream) {
    GPAD_REQUIRE(dest && src && size > 0, "gpad_array_copy: bad argument");
    cudaStream_t s; stream_of(stream, &s);
    return launch_copy(dest, src, size, s);
}

int gpad_step_three(float theta, const float* zhat_v, float* z_v, int length, void* stream) {
    GPAD_REQUIRE(zhat_v && z_v && length > 0, "gpad_step_three: bad argument");
    cudaStream_t s; stream_of(stream, &s);
    return launch_step_three(theta, zhat_v, z_v, length, s);
}

int gpad_step_four(const float* G_L, float* y_vp1, const float* w_v, const float* p_D, const float* zhat_v, int N, int n_u,
                   int m, int max_threads, void* stream) {
    (void)max_threads;
    GPAD_REQUIRE(G_L && y_vp1 && w_v && p_D && zhat_v && N > 0 && n_u > 0 && m > 0, "gpad_step_four: bad argument");
    cudaStream_t s; stream_of(stream, &s);
    return launch_gemv_t(G_L, zhat_v, N * n_u, m, 1, w_v, p_D, y_vp1, s);      // flipped G_L: [n][m]
}

int gpad_setup(const gpad_config_t* cfg, const float* M_G, const float* G_L, gpad_handle_t* out) {
    GPAD_REQUIRE(cfg && M_G && G_L && out, "gpad_setup: null argument");
    GPAD_REQUIRE(cfg->n_u > 0 && cfg->N > 0 && cfg->m > 0, "gpad_setup: n_u, N, m must be positive");
    GPAD_REQUIRE(cfg->layout == GPAD_LAYOUT_FLIPPED || cfg->layout == GPAD_LAYOUT_SEQUENTIAL || cfg->layout == GPAD_LAYOUT_FLAT,
                 "gpad_setup: bad layout");
    GPAD_REQUIRE(cfg->layout != GPAD_LAYOUT_FLAT || (cfg->mode != GPAD_MODE_BATCH_PER_INSTANCE && cfg->m >= 4 * cfg->n_u * cfg->N),
                 "gpad_setup: GPAD_LAYOUT_FLAT needs m >= 4 n_u N and shared operators");
    GPAD_REQUIRE(cfg->mode >= GPAD_MODE_LATENCY && cfg->mode <= GPAD_MODE_BATCH_PER_INSTANCE, "gpad_setup: bad mode");
    GPAD_REQUIRE(cfg->precision == GPAD_PREC_FP32 || cfg->precision == GPAD_PREC_TF32X3, "gpad_setup: bad precision");
    GPAD_REQUIRE(cfg->max_batch >= 1, "gpad_setup: max_batch must be >= 1");
    GPAD_REQUIRE(cfg->mode != GPAD_MODE_LATENCY || cfg->max_batch == 1, "gpad_setup: latency mode solves one QP (max_batch = 1)");
    GPAD_REQUIRE(cfg->precision == GPAD_PREC_FP32 || cfg->mode == GPAD_MODE_BATCH_SHARED,
                 "gpad_setup: GPAD_PREC_TF32X3 is only available in GPAD_MODE_BATCH_SHARED");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device: %s (libgpad_b200 has no CPU fallback)", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return GPAD_ERR_NO_DEVICE;
    }
    int dev = cfg->device;
    if (dev < 0) GPAD_CUDA(cudaGetDevice(&dev));
    GPAD_REQUIRE(dev < count, "gpad_setup: device %d out of range (%d devices)", dev, count);
    GPAD_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    GPAD_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libgpad_b200 contains sm_100a code only", dev, prop.major, prop.minor);
        return GPAD_ERR_UNSUPPORTED;
    }
    gpad_handle_s* h = new gpad_handle_s;
    h->cfg = *cfg; h->cfg.device = dev;
    h->n = cfg->n_u * cfg->N; h->device = dev; h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    int rc = GPAD_OK;
    std::vector<float> MG, GL;
    do {
        if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { rc = GPAD_ERR_CUDA; set_error("stream creation failed"); break; }
        if (cfg->mode == GPAD_MODE_BATCH_PER_INSTANCE) { rc = setup_per_instance(h, M_G, G_L); break; }
        rc = fetch_operators(*cfg, M_G, G_L, (size_t)h->n * cfg->m, 1, MG, GL);
        if (rc != GPAD_OK) break;
        rc = cfg->mode == GPAD_MODE_LATENCY ? setup_latency(h, MG, GL) : setup_batch(h, MG, GL);
    } while (0);
    if (rc != GPAD_OK) { gpad_destroy(h); return rc; }
    *out = h;
    return GPAD_OK;
}

int gpad_destroy(gpad_handle_t h) {
    if (!h) return GPAD_OK;
    cudaSetDevice(h->device);
    for (void* p : h->allocs) cudaFree(p);
    for (int k = 0; k < 3; ++k)
        for (auto& pr : h->ev_used[k]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->h_active) cudaFreeHost(h->h_active);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->stream_in) cudaStreamDestroy(h->stream_in);
    if (h->stream_out) cudaStreamDestroy(h->stream_out);
    cudaGetLastError();
    delete h;
    return GPAD_OK;
}

int gpad_solve(gpad_handle_t h, const gpad_solve_args_t* a) {
    GPAD_REQUIRE(h && a, "gpad_solve: null argument");
    GPAD_REQUIRE(a->batch >= 1 && a->batch <= h->cfg.max_batch, "gpad_solve: batch %d outside 1..%d", a->batch, h->cfg.max_batch);
    GPAD_REQUIRE(a->mem == GPAD_MEM_HOST || a->mem == GPAD_MEM_DEVICE, "gpad_solve: bad memspace");
    GPAD_REQUIRE(a->g_P && a->p_D, "gpad_solve: g_P and p_D are required");
    GPAD_REQUIRE(a->max_iter >= 1 && a->theta && a->beta, "gpad_solve: max_iter >= 1 and theta/beta are required");
    GPAD_REQUIRE(a->check_every <= 0 || (a->eps_g >= 0.f && a->eps_V >= 0.f), "gpad_solve: negative tolerance");
    if (h->cfg.mode == GPAD_MODE_BATCH_SHARED && a->check_every > 0 && a->f && (h->g2.xf2 || h->g2.p1)) {
        set_error("gpad_solve: the dual-gap branch (f != NULL) is not available with the experimental product-2 kernels (GPAD_TC_XF2 / GPAD_TC_P2TS)");
        return GPAD_ERR_UNSUPPORTED;
    }
    GPAD_CUDA(cudaSetDevice(h->device));
    if (h->cfg.mode == GPAD_MODE_BATCH_PER_INSTANCE) return solve_per_instance(h, a);
    return h->cfg.mode == GPAD_MODE_LATENCY ? solve_latency(h, a) : solve_batch(h, a);
}

int gpad_profile_enable(gpad_handle_t h, int enable) {
    GPAD_REQUIRE(h, "gpad_profile_enable: null handle");
    h->profile = enable != 0;
    return GPAD_OK;
}

int gpad_profile_read(gpad_handle_t h, int which, double* total_ms, long long* launches) {
    GPAD_REQUIRE(h && which >= 0 && which < 3, "gpad_profile_read: bad argument");
    GPAD_CUDA(cudaSetDevice(h->device));
    double ms = 0.0;
    for (auto& pr : h->ev_used[which]) {
        GPAD_CUDA(cudaEventSynchronize(pr.second));
        float t = 0.f;
        GPAD_CUDA(cudaEventElapsedTime(&t, pr.first, pr.second));
        ms += t;
        h->ev_pool.push_back(pr.first);
        h->ev_pool.push_back(pr.second);
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = (long long)h->ev_used[which].size();
    h->ev_used[which].clear();
    return GPAD_OK;
}

long long gpad_launch_count(gpad_handle_t h) { return h ? h->launches : 0; }
const char* gpad_describe(gpad_handle_t h) { return h ? h->desc.c_str() : ""; }

int gpad_debug_plan_tiles(int kernel, int ncols, int* bn, int* n_tiles, int* step, int* tmem_cols) {
    GPAD_REQUIRE(bn && n_tiles && step && tmem_cols && ncols > 0 && (kernel == 0 || kernel == 1), "gpad_debug_plan_tiles: bad argument");
    if (kernel == 1) {
        tc::plan_tiles_p1(ncols, bn, n_tiles, step);
        *tmem_cols = 2 * *bn + 96;                 // two accumulators + 3 state slots of 32 columns (batch_tc_p1.cu)
    } else {
        tc::plan_tiles(ncols, bn, n_tiles);
        *step = *bn;
        *tmem_cols = 2 * 256;                      // two accumulators at a fixed 256-column stride (batch_tc.cu)
    }
    return GPAD_OK;
}

int gpad_debug_gemm_tf32x3(const float* A, const float* B, float* C, int M, int N, int K, void* stream) {
    GPAD_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "gpad_debug_gemm_tf32x3: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0;
    GPAD_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    GPAD_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) { set_error("sm_100 device required"); return GPAD_ERR_UNSUPPORTED; }
    int bk = 16;
    if (const char* e = getenv("GPAD_TC_BK")) bk = atoi(e) == 32 ? 32 : 16;
    tc::GemmDesc g;
    g.bk = bk;
    g.cg = 1;
    if (const char* e = getenv("GPAD_TC_CG")) g.cg = atoi(e) == 2 ? 2 : 1;
    g.mc = g.cg == 1 ? tc_multicast() : 1;
    g.k_pad = round_up(K, 32);
    tc::plan_tiles(N, &g.bn, &g.n_tiles);
    g.m_tiles = round_up(M, 128 * g.cg) / (128 * g.cg);
    g.ncols_valid = N;
    g.stages = g.cg == 2 ? tc::pick_stages2(bk, g.bn, prop.sharedMemPerBlockOptin) : tc::pick_stages(bk, g.bn, prop.sharedMemPerBlockOptin);
    if (const char* e = getenv("GPAD_TC_STAGES")) g.stages = std::min(g.stages, std::max(2, atoi(e)));
    const int Mp = round_up(g.m_tiles, g.mc) * 128 * g.cg, Np = round_up(g.bn * g.n_tiles, 128);
    float *Ap, *Al, *Bp, *Bl;
    const size_t ca = (size_t)Mp * g.k_pad, cb = (size_t)Np * g.k_pad;
    GPAD_CUDA(cudaMalloc(&Ap, ca * 4)); GPAD_CUDA(cudaMalloc(&Al, ca * 4));
    GPAD_CUDA(cudaMalloc(&Bp, cb * 4)); GPAD_CUDA(cudaMalloc(&Bl, cb * 4));
    int rc = GPAD_OK;
    do {
        if ((rc = launch_pad_rows(Ap, g.k_pad, Mp, A, K, M, s)) != GPAD_OK) break;
        if ((rc = launch_pad_rows(Bp, g.k_pad, Np, B, K, N, s)) != GPAD_OK) break;
        if ((rc = tc::launch_split(Ap, Ap, Al, ca, s)) != GPAD_OK) break;
        if ((rc = tc::launch_split(Bp, Bp, Bl, cb, s)) != GPAD_OK) break;
        if ((rc = tc::make_tmap(&g.tmA_hi, Ap, g.k_pad, Mp, g.k_pad, bk, 128)) != GPAD_OK) break;
        if ((rc = tc::make_tmap(&g.tmA_lo, Al, g.k_pad, Mp, g.k_pad, bk, 128)) != GPAD_OK) break;
        if ((rc = tc::make_tmap(&g.tmB_hi, Bp, g.k_pad, Np, g.k_pad, bk, g.bn / (g.cg * g.mc))) != GPAD_OK) break;
        if ((rc = tc::make_tmap(&g.tmB_lo, Bl, g.k_pad, Np, g.k_pad, bk, g.bn / (g.cg * g.mc))) != GPAD_OK) break;
        BatchKernelArgs k{};
        k.B = M;
        rc = tc::launch_gemm(0, g, k, C, N, prop.multiProcessorCount, s);
    } while (0);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(Ap); cudaFree(Al); cudaFree(Bp); cudaFree(Bl);
    if (rc == GPAD_OK && e != cudaSuccess) return cuda_fail(e, "gpad_debug_gemm_tf32x3", __FILE__, __LINE__);
    return rc;
}

}  // extern "C"
