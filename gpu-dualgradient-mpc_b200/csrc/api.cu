// api.cu -- the C ABI of libgpad_b200.so (include/gpad.h): handle management, operator
// conversion, path selection and the host loop that drives the kernels.  Replaces the body of
// the reference's main() between readData and the D2H copies (main.cu:108-180).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "handle.h"

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

// GPAD_DEBUG="key=value,key=value": the one place the library reads the environment, once per gpad_setup
Knobs parse_knobs() {
    Knobs k;
    const char* env = getenv("GPAD_DEBUG");
    if (!env) return k;
    std::string all(env);
    size_t pos = 0;
    while (pos < all.size()) {
        size_t end = all.find(',', pos);
        if (end == std::string::npos) end = all.size();
        const std::string item = all.substr(pos, end - pos);
        pos = end + 1;
        const size_t eq = item.find('=');
        if (eq == std::string::npos) continue;
        const std::string key = item.substr(0, eq), val = item.substr(eq + 1);
        const int iv = atoi(val.c_str());
        if (key == "latency_plan") k.latency_plan = val;
        else if (key == "latency_threads") k.latency_threads = iv;
        else if (key == "latency_no_smem_ops") k.latency_no_smem_ops = iv;
        else if (key == "latency_grid2") k.latency_grid2 = iv;
        else if (key == "latency_warp") k.latency_warp = iv;
        else if (key == "latency_flat") k.latency_flat = iv;
        else if (key == "flat_xchg") k.flat_xchg = iv;
        else if (key == "warp_rows") k.warp_rows = iv;
        else if (key == "warp_ordered") k.warp_ordered = iv;
        else if (key == "warp_pack") k.warp_pack = iv;
        else if (key == "tc_p1") k.tc_p1 = iv;
        else if (key == "tc_stages") k.tc_stages = iv;
        else if (key == "tc_bn2") k.tc_bn2 = iv;
        else if (key == "tc_autotune") k.tc_autotune = iv;
        else if (key == "tc_pdl") k.tc_pdl = iv;
        else if (key == "sync_split") k.sync_split = iv;
        else if (key == "tc_cluster_attr") k.tc_cluster_attr = iv;
        else if (key == "tc_retire") k.tc_retire = iv;
        else if (key == "tc_compact") k.tc_compact = iv;
        else if (key == "check_lag") k.check_lag = std::max(0, std::min(64, iv));
    }
    return k;
}

// rows x cols (ld = cols) -> device rows_pad x ld_pad, zero padded
int upload_padded(gpad_handle_s* h, const float* src, int rows, int cols, int rows_pad, int ld_pad, float** out) {
    std::vector<float> tmp((size_t)rows_pad * ld_pad, 0.f);
    for (int r = 0; r < rows; ++r) memcpy(&tmp[(size_t)r * ld_pad], src + (size_t)r * cols, sizeof(float) * cols);
    GPAD_TRY(dev_alloc(h, out, tmp.size()));
    GPAD_CUDA(cudaMemcpy(*out, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice));
    return GPAD_OK;
}

// One solve in flight per handle: the per-handle scratch (schedule tables, flags, exchange buffers, batch state) is
// shared by all solves, so every solve first makes its stream wait for the end of the previous one -- whatever stream
// that ran on -- and records its own end.
int solve_begin(gpad_handle_s* h, cudaStream_t s) {
    if (h->ev_last_valid) GPAD_CUDA(cudaStreamWaitEvent(s, h->ev_last, 0));
    return GPAD_OK;
}
int solve_end(gpad_handle_s* h, cudaStream_t s) {
    GPAD_CUDA(cudaEventRecord(h->ev_last, s));
    h->ev_last_valid = true;
    return GPAD_OK;
}

// theta / beta tables on the device.  Called after solve_begin(h, s): the stream already waits for the previous solve,
// so rewriting the tables cannot race with a kernel that still reads them; a CHANGED schedule additionally waits for the
// asynchronous pipeline's copy streams (rare: schedules are normally constant).
int upload_schedule(gpad_handle_s* h, const float* theta, const float* beta, int count, cudaStream_t s) {
    if ((int)h->h_theta.size() == count && !memcmp(h->h_theta.data(), theta, sizeof(float) * count) &&
        !memcmp(h->h_beta.data(), beta, sizeof(float) * count))
        return GPAD_OK;
    if (h->stream_in) { GPAD_CUDA(cudaStreamSynchronize(h->stream_in)); GPAD_CUDA(cudaStreamSynchronize(h->stream_out)); }
    if (count > h->sched_cap) {
        GPAD_CUDA(cudaStreamSynchronize(s));          // the old tables may still be read by work queued on s
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

}  // namespace gpad

using namespace gpad;

namespace {

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

// ---- latency plan: how many CTAs cooperate, how they synchronise, where the operators live ----
struct LatPlan {
    int sync = -1, G = 1, threads = 64;
    bool regs = false;
    bool small = false;
    int cha = 1, chb = 1;
    lat::Params p{};
};

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

// ------------------------------------------------------------------ latency mode
int setup_latency(gpad_handle_s* h, const std::vector<float>& MG, const std::vector<float>& GL) {
    const int n = h->n, m = h->cfg.m;
    const size_t limit = h->smem_optin;
    const Knobs& kn = h->knobs;
    const bool no_res = kn.latency_no_smem_ops != 0;
    LatPlan plan;
    const char* env = kn.latency_plan.empty() ? nullptr : kn.latency_plan.c_str();    // "block" | "cluster:<C>" | "grid:<G>" | "lean:<C>"
    bool ok = false;
    if (env) {
        int sync = lat::SYNC_GRID, G = h->num_sms;
        if (!strncmp(env, "block", 5)) { sync = lat::SYNC_BLOCK; G = 1; }
        else if (!strncmp(env, "cluster:", 8)) { sync = lat::SYNC_CLUSTER; G = std::max(1, std::min(16, atoi(env + 8))); }
        else if (!strncmp(env, "lean:", 5)) { sync = lat::SYNC_CLUSTER; G = std::max(1, std::min(16, atoi(env + 5))); }
        else if (!strncmp(env, "grid:", 5)) { G = std::max(1, std::min(h->num_sms, atoi(env + 5))); }
        if (!strncmp(env, "lean:", 5)) {
            ok = plan_small(n, m, G, plan);
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
        if (!ok) { ok = plan_for(n, m, h->num_sms, false, limit, no_res, plan); plan.sync = lat::SYNC_GRID; }
    }
    if (!ok) {
        set_error("latency mode: vectors of n=%d, m=%d do not fit in shared memory", n, m);
        return GPAD_ERR_UNSUPPORTED;
    }
    if (kn.latency_threads > 0 && !plan.regs) plan.threads = std::max(32, std::min(lat::kMaxThreads, kn.latency_threads / 32 * 32));
    lat::Params& p = h->lp;
    p = plan.p;
    h->sync_mode = plan.sync; h->G = plan.G; h->threads = plan.threads; h->ops_smem = plan.regs;
    h->small = plan.small; h->cha = plan.cha; h->chb = plan.chb;
    p.L = h->cfg.L;
    p.batch = 1; p.op_stride_a = p.op_stride_b = 0;
    p.warp_rows = kn.warp_rows; p.warp_ordered = kn.warp_ordered; p.warp_pack = kn.warp_pack;
    float *dMG, *dGL;
    GPAD_TRY(upload_padded(h, MG.data(), n, m, n, p.mld, &dMG));
    GPAD_TRY(upload_padded(h, GL.data(), m, n, m, p.nld, &dGL));
    p.M_G = dMG; p.G_L = dGL;

    // per-solve device buffers: ONE input block [g_P | p_D | y0 | y_prev0 | f] and ONE output block [y_next | y | w | z | zhat |
    // iters, status, max_viol, gap] (sub-arrays on 16-byte boundaries), each mirrored by a pinned host block, so that a
    // host-memory solve costs one H2D and one D2H copy instead of up to five and nine (main.cu:136-147, 176-180 copy vector by
    // vector; on a 35 us solve those copies were 110 us)
    {
        const size_t na = round_up_sz(n, 4), ma = round_up_sz(m, 4);
        h->lat_in_floats = 2 * na + 3 * ma;
        h->lat_out_floats = 3 * ma + 2 * na + 4;
        float *din = nullptr, *dout = nullptr;
        GPAD_TRY(dev_alloc(h, &din, h->lat_in_floats)); GPAD_TRY(dev_alloc(h, &dout, h->lat_out_floats));
        h->d_gP = din; h->d_pD = din + na; h->d_y0 = h->d_pD + ma; h->d_yprev0 = h->d_y0 + ma; h->d_f = h->d_yprev0 + ma;
        h->o_ynext = dout; h->o_y = dout + ma; h->o_w = h->o_y + ma; h->o_z = h->o_w + ma; h->o_zhat = h->o_z + na;
        float* tail = h->o_zhat + na;
        h->o_iters = reinterpret_cast<int*>(tail); h->o_status = reinterpret_cast<int*>(tail + 1); h->o_viol = tail + 2; h->o_gap = tail + 3;
        GPAD_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->lat_h_in), sizeof(float) * h->lat_in_floats));
        GPAD_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->lat_h_out), sizeof(float) * h->lat_out_floats));
    }
    GPAD_TRY(dev_alloc(h, &p.x_w, p.mld)); GPAD_TRY(dev_alloc(h, &p.x_zhat, p.nld));
    GPAD_CUDA(cudaMemset(p.x_w, 0, sizeof(float) * p.mld));          // the zero padding is gathered too
    GPAD_CUDA(cudaMemset(p.x_zhat, 0, sizeof(float) * p.nld));
    GPAD_TRY(dev_alloc(h, &p.x_red, 3 * 8 * p.g_pad));
    GPAD_TRY(dev_alloc(h, &h->d_flags, 2));
    p.barrier = h->d_flags; p.nonfinite_flag = reinterpret_cast<int*>(h->d_flags + 1);

    char where[192];
    if (plan.small) snprintf(where, sizeof(where), "and per-row state in registers (lean kernel, %dx%d fragments)", plan.cha, plan.chb);
    else if (plan.regs) snprintf(where, sizeof(where), "in registers");
    else snprintf(where, sizeof(where), "%d/%d + %d/%d rows per CTA in shared memory, rest streamed from L2", p.res_a, p.rows_a, p.res_b, p.rows_b);
    char buf[448];
    snprintf(buf, sizeof(buf), "latency: persistent kernel, %s x%d CTAs, %d threads, lanes/row %d|%d, operators %s, smem %zu B/CTA",
             plan.sync == lat::SYNC_BLOCK ? "single-CTA" : plan.sync == lat::SYNC_CLUSTER ? "cluster(DSMEM)" : "cooperative-grid",
             plan.G, plan.threads, 1 << p.lg_a, 1 << p.lg_b, where,
             plan.small ? lat::small_smem_bytes(p) : lat::smem_bytes(p, plan.regs));
    h->desc = buf;
    // fixed-iteration solves of a whole-chip plan run the second-generation kernel when it covers the problem
    h->grid2 = plan.sync == lat::SYNC_GRID && !plan.small && plan.G == h->num_sms && lat::grid2_supported(p, limit);
    if (kn.latency_grid2 == 0) h->grid2 = false;
    h->warp = plan.small && plan.G == 1 && lat::warp_supported(p);
    if (kn.latency_warp == 0) h->warp = false;
    if (h->warp) h->desc = "latency: one warp, operators and state in registers, no shared memory or block barrier in the loop (latency_warp.cu)";
    // Battery structure (identical cells: Cookbook 2.2, seq_functions.cpp:5-43): when the operators are exactly "flat" the
    // fixed-iteration solves can run latency_flat.cu on n_u x fewer operator bytes, one cluster instead of the whole chip.
    // Measured on B200 (100 iterations): (10,100) 510 us against 506 us for the dense whole-chip kernel, the small
    // problems 1.3-2x slower than their dense one-CTA / cluster kernels.  So it is the default only for
    // operators that ARRIVE flat (GPAD_LAYOUT_FLAT) and would otherwise need the whole chip; GPAD_DEBUG latency_flat=1
    // forces it for any flat problem, =0 disables it.
    const bool flat_default = h->cfg.layout == GPAD_LAYOUT_FLAT && plan.sync == lat::SYNC_GRID && !env;
    if (kn.latency_flat != 0 && h->cfg.n_u >= 2 && (kn.latency_flat == 1 || flat_default)) {
        lat::FlatParams fp{};
        const int max_cluster = lat::max_cluster_size(lat::kMaxThreads, 64 * 1024);
        if (lat::plan_flat(h->cfg.n_u, h->cfg.N, m, limit, max_cluster, &fp)) {
            std::vector<float> A_op, B_op;
            const float resid = lat::build_flat_operators(fp, MG.data(), GL.data(), A_op, B_op);
            if (resid == 0.f) {
                float *dA, *dB;
                GPAD_TRY(dev_alloc(h, &dA, A_op.size())); GPAD_TRY(dev_alloc(h, &dB, B_op.size()));
                GPAD_CUDA(cudaMemcpy(dA, A_op.data(), A_op.size() * sizeof(float), cudaMemcpyHostToDevice));
                GPAD_CUDA(cudaMemcpy(dB, B_op.data(), B_op.size() * sizeof(float), cudaMemcpyHostToDevice));
                fp.A_op = dA; fp.B_op = dB;
                fp.xchg = kn.flat_xchg ? 1 : 0;
                fp.nonfinite_flag = reinterpret_cast<int*>(h->d_flags + 1);
                h->fp = fp; h->flat = true;
                snprintf(buf, sizeof(buf), "latency: flat battery operators (%dx fewer bytes) on one %d-CTA cluster, %d threads: phase A %d stages x %d cells "
                         "per CTA from shared memory (%d B), phase B one row per thread in registers, DSMEM st.async + mbarrier exchange "
                         "(latency_flat.cu); tolerance-mode solves: %s", h->cfg.n_u, fp.C, fp.threads, fp.SC, h->cfg.n_u,
                         (int)(sizeof(float) * fp.SC * h->cfg.n_u * fp.lenA), h->desc.c_str());
                h->desc = buf;
            }
        }
    }
    if (h->grid2 && !h->flat) {
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
    GPAD_TRY(solve_begin(h, s));
    GPAD_TRY(upload_schedule(h, a->theta, a->beta, a->max_iter, s));
    lat::Params p = h->lp;
    // the one-warp kernel reads every input once (into registers) and writes every output once: for host-memory solves it
    // works straight on the pinned mirrors (zero copy; a few hundred bytes cross PCIe), which removes both DMA copies
    // from a 35 us solve
    const bool zero_copy = host && h->warp && a->max_iter >= 1 && !(h->flat && a->check_every <= 0);
    if (host) {
        // gather the inputs that exist into the pinned mirror of the device block, one copy of its used prefix
        // (block order: g_P, p_D, y0, y_prev0, f -- the optional ones last)
        float* hi = h->lat_h_in;
        memcpy(hi + (h->d_gP - h->d_gP), a->g_P, sizeof(float) * n);
        memcpy(hi + (h->d_pD - h->d_gP), a->p_D, sizeof(float) * m);
        size_t used = (size_t)(h->d_y0 - h->d_gP);
        if (a->y0) { memcpy(hi + (h->d_y0 - h->d_gP), a->y0, sizeof(float) * m); used = (size_t)(h->d_yprev0 - h->d_gP); }
        if (a->y_prev0) { memcpy(hi + (h->d_yprev0 - h->d_gP), a->y_prev0, sizeof(float) * m); used = (size_t)(h->d_f - h->d_gP); }
        if (a->f) { memcpy(hi + (h->d_f - h->d_gP), a->f, sizeof(float) * n); used = h->lat_in_floats; }
        if (!zero_copy) GPAD_CUDA(cudaMemcpyAsync(h->d_gP, hi, sizeof(float) * used, cudaMemcpyHostToDevice, s));
        const float* in = zero_copy ? hi : h->d_gP;
        p.g_P = in; p.p_D = in + (h->d_pD - h->d_gP); p.f = a->f ? in + (h->d_f - h->d_gP) : nullptr;
        p.y0 = a->y0 ? in + (h->d_y0 - h->d_gP) : nullptr; p.y_prev0 = a->y_prev0 ? in + (h->d_yprev0 - h->d_gP) : nullptr;
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
    if (zero_copy) {
        float* ho = h->lat_h_out;
        p.out_y_next = ho; p.out_y = ho + (h->o_y - h->o_ynext); p.out_w = ho + (h->o_w - h->o_ynext);
        p.out_z = ho + (h->o_z - h->o_ynext); p.out_zhat = ho + (h->o_zhat - h->o_ynext);
        float* tail = ho + (h->o_viol - 2 - h->o_ynext);
        p.out_iters = reinterpret_cast<int*>(tail); p.out_status = reinterpret_cast<int*>(tail + 1);
        p.out_max_viol = tail + 2; p.out_gap = tail + 3;
    }
    cudaEvent_t pe = h->prof_begin(s);
    if (h->flat && p.check_every == 0 && p.max_iter >= 1) {
        lat::FlatParams fp = h->fp;
        fp.g_P = p.g_P; fp.p_D = p.p_D; fp.y0 = p.y0; fp.y_prev0 = p.y_prev0; fp.theta = p.theta; fp.beta = p.beta;
        fp.max_iter = p.max_iter;
        fp.out_y_next = p.out_y_next; fp.out_y = p.out_y; fp.out_z = p.out_z; fp.out_zhat = p.out_zhat; fp.out_w = p.out_w;
        fp.out_iters = p.out_iters; fp.out_status = p.out_status; fp.out_max_viol = p.out_max_viol; fp.out_gap = p.out_gap;
        GPAD_CUDA(cudaMemsetAsync(h->d_flags, 0, 2 * sizeof(unsigned), s));
        GPAD_TRY(lat::launch_flat(fp, s));
    } else if (h->warp && p.max_iter >= 1) {
        p.batch = 1; p.op_stride_a = 0; p.op_stride_b = 0;
        GPAD_TRY(lat::launch_warp(p, s));
    } else if (h->small) {
        p.sched_smem = round_up(std::min(a->max_iter, lat::small_sched_capacity()), 4);
        GPAD_TRY(lat::launch_small(p, h->cha, h->chb, h->G, h->threads, s));
    } else if (h->grid2 && p.max_iter >= 1) {
        GPAD_CUDA(cudaMemsetAsync(h->d_flags, 0, 2 * sizeof(unsigned), s));
        GPAD_TRY(lat::launch_grid2(p, h->G, s));
    } else {
        GPAD_CUDA(cudaMemsetAsync(h->d_flags, 0, 2 * sizeof(unsigned), s));
        GPAD_TRY(lat::launch(p, h->sync_mode, h->ops_smem, h->G, h->threads, s));
    }
    h->prof_end(0, pe, s);
    h->launches += 1;
    if (host && !zero_copy) GPAD_CUDA(cudaMemcpyAsync(h->lat_h_out, h->o_ynext, sizeof(float) * h->lat_out_floats, cudaMemcpyDeviceToHost, s));
    GPAD_TRY(solve_end(h, s));
    if (host) {
        GPAD_CUDA(cudaStreamSynchronize(s));
        const float* ho = h->lat_h_out;
        if (a->y_next) memcpy(a->y_next, ho + (h->o_ynext - h->o_ynext), sizeof(float) * m);
        if (a->y) memcpy(a->y, ho + (h->o_y - h->o_ynext), sizeof(float) * m);
        if (a->w) memcpy(a->w, ho + (h->o_w - h->o_ynext), sizeof(float) * m);
        if (a->z) memcpy(a->z, ho + (h->o_z - h->o_ynext), sizeof(float) * n);
        if (a->zhat) memcpy(a->zhat, ho + (h->o_zhat - h->o_ynext), sizeof(float) * n);
        const float* tail = ho + (h->o_viol - 2 - h->o_ynext);
        if (a->iters) memcpy(a->iters, tail, sizeof(int));
        if (a->status) memcpy(a->status, tail + 1, sizeof(int));
        if (a->max_viol) *a->max_viol = tail[2];
        if (a->gap) *a->gap = tail[3];
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
    p.warp_rows = h->knobs.warp_rows; p.warp_ordered = h->knobs.warp_ordered; p.warp_pack = h->knobs.warp_pack;
    h->warp = lat::warp_supported(p);
    if (h->knobs.latency_warp == 0) h->warp = false;
    if (h->warp) h->desc = h->knobs.warp_pack
        ? "batch-per-instance: batched GEMV in registers (latency_warp.cu): fixed-iteration solves pack TWO QPs per warp (16 lanes each, "
          "8 QPs per CTA), tolerance-mode solves one QP per warp; operators and state read once, no shared memory or block barrier in the loop"
        : "batch-per-instance: one WARP per QP (batched GEMV), 4 QPs per CTA, operators and state read once into registers, "
          "no shared memory or block barrier in the loop (latency_warp.cu)";
    return GPAD_OK;
}

int solve_per_instance(gpad_handle_s* h, const gpad_solve_args_t* a) {
    const int n = h->n, m = h->cfg.m, B = a->batch;
    const bool host = a->mem == GPAD_MEM_HOST;
    cudaStream_t s = host ? h->own_stream : static_cast<cudaStream_t>(a->stream);
    GPAD_TRY(solve_begin(h, s));
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
    }
    GPAD_TRY(solve_end(h, s));
    if (host) GPAD_CUDA(cudaStreamSynchronize(s));
    return GPAD_OK;
}

// selects a device for the duration of a call and restores the caller's current device afterwards
struct DeviceGuard {
    int prev = -1;
    cudaError_t status = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != dev) status = cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

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
    GPAD_REQUIRE(cfg->precision == GPAD_PREC_FP32 || cfg->precision == GPAD_PREC_TF32X3 || cfg->precision == GPAD_PREC_FP16X3,
                 "gpad_setup: bad precision");
    GPAD_REQUIRE(cfg->max_batch >= 1, "gpad_setup: max_batch must be >= 1");
    GPAD_REQUIRE(cfg->mode != GPAD_MODE_LATENCY || cfg->max_batch == 1, "gpad_setup: latency mode solves one QP (max_batch = 1)");
    GPAD_REQUIRE(cfg->precision == GPAD_PREC_FP32 || cfg->mode == GPAD_MODE_BATCH_SHARED,
                 "gpad_setup: GPAD_PREC_TF32X3 / GPAD_PREC_FP16X3 are only available in GPAD_MODE_BATCH_SHARED");
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
    DeviceGuard guard(dev);
    GPAD_CUDA(guard.status);
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
    h->knobs = parse_knobs();
    int rc = GPAD_OK;
    std::vector<float> MG, GL;
    do {
        if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming) != cudaSuccess) {
            rc = GPAD_ERR_CUDA; set_error("stream / event creation failed"); break;
        }
        if (cfg->mode == GPAD_MODE_BATCH_PER_INSTANCE) { rc = setup_per_instance(h, M_G, G_L); break; }
        rc = fetch_operators(*cfg, M_G, G_L, (size_t)h->n * cfg->m, 1, MG, GL);
        if (rc != GPAD_OK) break;
        rc = cfg->mode == GPAD_MODE_LATENCY ? setup_latency(h, MG, GL) : setup_batch(h, MG, GL);
    } while (0);
    // setup work ran on several streams (legacy default stream included): everything has landed before the first solve
    if (rc == GPAD_OK && cudaDeviceSynchronize() != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "gpad_setup", __FILE__, __LINE__); }
    if (rc != GPAD_OK) { gpad_destroy(h); return rc; }
    *out = h;
    return GPAD_OK;
}

int gpad_destroy(gpad_handle_t h) {
    if (!h) return GPAD_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();                       // nothing of this handle may still be running when its memory goes
    for (void* p : h->allocs) cudaFree(p);
    for (int k = 0; k < 3; ++k)
        for (auto& pr : h->ev_used[k]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    destroy_batch(h);
    if (h->lat_h_in) cudaFreeHost(h->lat_h_in);
    if (h->lat_h_out) cudaFreeHost(h->lat_h_out);
    if (h->ev_last) cudaEventDestroy(h->ev_last);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->stream_in) cudaStreamDestroy(h->stream_in);
    if (h->stream_out) cudaStreamDestroy(h->stream_out);
    cudaGetLastError();
    delete h;
    return GPAD_OK;
}

static int check_solve_args(gpad_handle_t h, const gpad_solve_args_t* a, const char* who) {
    GPAD_REQUIRE(h && a, "%s: null argument", who);
    GPAD_REQUIRE(a->batch >= 1 && a->batch <= h->cfg.max_batch, "%s: batch %d outside 1..%d", who, a->batch, h->cfg.max_batch);
    GPAD_REQUIRE(a->mem == GPAD_MEM_HOST || a->mem == GPAD_MEM_DEVICE, "%s: bad memspace", who);
    if (a->params) {
        GPAD_REQUIRE(a->problem, "%s: params need the problem they parametrise", who);
        GPAD_REQUIRE(h->cfg.mode == GPAD_MODE_BATCH_SHARED, "%s: on-device instance build is available in GPAD_MODE_BATCH_SHARED", who);
        int nu = 0, N = 0, m = 0;
        GPAD_REQUIRE(gpad_problem_dims(a->problem, &nu, &N, &m, nullptr, nullptr) == GPAD_OK && nu == h->cfg.n_u && N == h->cfg.N && m == h->cfg.m,
                     "%s: the problem's dimensions differ from the handle's", who);
    } else {
        GPAD_REQUIRE(a->g_P && a->p_D, "%s: g_P and p_D (or params + problem) are required", who);
        GPAD_REQUIRE(!a->build_f, "%s: build_f needs params + problem", who);
    }
    GPAD_REQUIRE(a->max_iter >= 1 && a->theta && a->beta, "%s: max_iter >= 1 and theta/beta are required", who);
    GPAD_REQUIRE(a->check_every <= 0 || (a->eps_g >= 0.f && a->eps_V >= 0.f), "%s: negative tolerance", who);
    return GPAD_OK;
}

int gpad_solve(gpad_handle_t h, const gpad_solve_args_t* a) {
    GPAD_TRY(check_solve_args(h, a, "gpad_solve"));
    DeviceGuard guard(h->device);
    GPAD_CUDA(guard.status);
    if (h->cfg.mode == GPAD_MODE_BATCH_PER_INSTANCE) return solve_per_instance(h, a);
    return h->cfg.mode == GPAD_MODE_LATENCY ? solve_latency(h, a) : solve_batch(h, a);
}

int gpad_solve_async(gpad_handle_t h, const gpad_solve_args_t* a, long long* ticket) {
    GPAD_TRY(check_solve_args(h, a, "gpad_solve_async"));
    GPAD_REQUIRE(h->cfg.mode == GPAD_MODE_BATCH_SHARED, "gpad_solve_async: GPAD_MODE_BATCH_SHARED only");
    DeviceGuard guard(h->device);
    GPAD_CUDA(guard.status);
    return solve_batch_async(h, a, ticket);
}

int gpad_wait(gpad_handle_t h, long long ticket) {
    GPAD_REQUIRE(h, "gpad_wait: null handle");
    DeviceGuard guard(h->device);
    GPAD_CUDA(guard.status);
    return wait_batch(h, ticket);
}

int gpad_handle_dims(gpad_handle_t h, int* n_u, int* N, int* m, int* mode, int* max_batch, int* device) {
    GPAD_REQUIRE(h, "gpad_handle_dims: null handle");
    if (n_u) *n_u = h->cfg.n_u;
    if (N) *N = h->cfg.N;
    if (m) *m = h->cfg.m;
    if (mode) *mode = h->cfg.mode;
    if (max_batch) *max_batch = h->cfg.max_batch;
    if (device) *device = h->device;
    return GPAD_OK;
}

int gpad_solve_stats(gpad_handle_t h, gpad_solve_stats_t* out) {
    GPAD_REQUIRE(h && out, "gpad_solve_stats: null argument");
    GPAD_REQUIRE(h->cfg.mode == GPAD_MODE_BATCH_SHARED, "gpad_solve_stats: GPAD_MODE_BATCH_SHARED only");
    DeviceGuard guard(h->device);
    GPAD_CUDA(guard.status);
    if (h->ev_last_valid) GPAD_CUDA(cudaEventSynchronize(h->ev_last));
    out->instance_iterations_scheduled = (double)h->h_stat[0];
    out->instance_iterations_needed = (double)h->h_stat[1];
    out->compactions = h->compactions;
    return GPAD_OK;
}

int gpad_profile_enable(gpad_handle_t h, int enable) {
    GPAD_REQUIRE(h, "gpad_profile_enable: null handle");
    h->profile = enable != 0;
    return GPAD_OK;
}

int gpad_profile_read(gpad_handle_t h, int which, double* total_ms, long long* launches) {
    GPAD_REQUIRE(h && which >= 0 && which < 3, "gpad_profile_read: bad argument");
    DeviceGuard guard(h->device);
    GPAD_CUDA(guard.status);
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
    GPAD_REQUIRE(bn && n_tiles && step && tmem_cols && ncols > 0 && kernel >= 0 && kernel <= 2, "gpad_debug_plan_tiles: bad argument");
    if (kernel == 2) {
        tc::plan_tiles_p2(ncols, bn, n_tiles);
        *step = *bn;
        *tmem_cols = 2 * 256;                      // two accumulators at a fixed 256-column stride (batch_tc_p2.cu)
    } else if (kernel == 1) {
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
    const Knobs kn = parse_knobs();
    const int bk = 16;
    tc::GemmDesc g;
    g.bk = bk;
    g.k_pad = round_up(K, 32);
    tc::plan_tiles(N, &g.bn, &g.n_tiles);
    g.m_tiles = round_up(M, 128) / 128;
    g.ncols_valid = N;
    g.stages = tc::pick_stages(bk, g.bn, prop.sharedMemPerBlockOptin);
    if (kn.tc_stages > 0) g.stages = std::min(g.stages, std::max(2, kn.tc_stages));
    const int Mp = g.m_tiles * 128, Np = round_up(g.bn * g.n_tiles, 128);
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
        if ((rc = tc::make_tmap(&g.tmB_hi, Bp, g.k_pad, Np, g.k_pad, bk, g.bn)) != GPAD_OK) break;
        if ((rc = tc::make_tmap(&g.tmB_lo, Bl, g.k_pad, Np, g.k_pad, bk, g.bn)) != GPAD_OK) break;
        BatchKernelArgs k{};
        k.B = M;
        rc = tc::launch_gemm(0, g, k, C, N, prop.multiProcessorCount, s);
    } while (0);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(Ap); cudaFree(Al); cudaFree(Bp); cudaFree(Bl);
    if (rc == GPAD_OK && e != cudaSuccess) return cuda_fail(e, "gpad_debug_gemm_tf32x3", __FILE__, __LINE__);
    return rc;
}

int gpad_debug_gemm_f16x3(const float* A, const float* B, float* C, int M, int N, int K, int kernel, void* stream) {
    GPAD_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && kernel >= 0 && kernel <= 2, "gpad_debug_gemm_f16x3: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0;
    GPAD_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    GPAD_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) { set_error("sm_100 device required"); return GPAD_ERR_UNSUPPORTED; }
    tc::GemmDesc g;
    g.f16 = 1; g.bk = 16;
    g.k_pad = round_up(K, 32);
    if (kernel == 2) { tc::plan_tiles_p1(N, &g.bn, &g.n_tiles, &g.step, 256); g.p1 = 1; g.acc_stages = 1; }
    else if (kernel == 1) { tc::plan_tiles_p1(N, &g.bn, &g.n_tiles, &g.step); g.p1 = 1; }
    else tc::plan_tiles(N, &g.bn, &g.n_tiles);
    g.m_tiles = round_up(M, 128) / 128;
    g.ncols_valid = N;
    const int Mp = g.m_tiles * 128, Np = round_up(g.bn * g.n_tiles, 128);
    const size_t ca = (size_t)Mp * g.k_pad, cb = (size_t)Np * g.k_pad;
    // one scratch allocation, carved up (256-byte aligned pieces): nothing to leak on an early error
    auto piece = [](size_t bytes) { return round_up_sz(bytes, 256); };
    const size_t total = piece(ca * 4) + piece(cb * 4) + 2 * piece(ca * 2) + 2 * piece(cb * 2) + 2 * piece((size_t)Mp * 4) + piece((size_t)Np * 4);
    char* scratch = nullptr;
    GPAD_CUDA(cudaMalloc(&scratch, total));
    char* cur = scratch;
    auto take = [&](size_t bytes) { char* p = cur; cur += piece(bytes); return p; };
    float* Ap = reinterpret_cast<float*>(take(ca * 4)); float* Bp = reinterpret_cast<float*>(take(cb * 4));
    uint16_t* Ah = reinterpret_cast<uint16_t*>(take(ca * 2)); uint16_t* Al = reinterpret_cast<uint16_t*>(take(ca * 2));
    uint16_t* Bh = reinterpret_cast<uint16_t*>(take(cb * 2)); uint16_t* Bl = reinterpret_cast<uint16_t*>(take(cb * 2));
    float* ainv = reinterpret_cast<float*>(take((size_t)Mp * 4)); float* amax = reinterpret_cast<float*>(take((size_t)Mp * 4));
    float* binv = reinterpret_cast<float*>(take((size_t)Np * 4));
    int rc = GPAD_OK;
    do {
        if ((rc = launch_pad_rows(Ap, g.k_pad, Mp, A, K, M, s)) != GPAD_OK) break;
        if ((rc = launch_pad_rows(Bp, g.k_pad, Np, B, K, N, s)) != GPAD_OK) break;
        if ((rc = tc::launch_quantize_rows(Bp, g.k_pad, Np, Bh, Bl, binv, nullptr, s)) != GPAD_OK) break;
        if ((rc = tc::make_tmap_bytes(&g.tmB_hi, Bh, 2, g.k_pad, Np, g.k_pad, 32, g.bn)) != GPAD_OK) break;
        if ((rc = tc::make_tmap_bytes(&g.tmB_lo, Bl, 2, g.k_pad, Np, g.k_pad, 32, g.bn)) != GPAD_OK) break;
        BatchKernelArgs k{};
        k.B = M;
        k.b_colinv = binv;
        if (kernel == 0) {
            g.stages = tc::pick_stages(16, g.bn, prop.sharedMemPerBlockOptin);
            if ((rc = tc::launch_quantize_rows(Ap, g.k_pad, Mp, Ah, Al, ainv, nullptr, s)) != GPAD_OK) break;
            if ((rc = tc::make_tmap_bytes(&g.tmA_hi, Ah, 2, g.k_pad, Mp, g.k_pad, 32, 128)) != GPAD_OK) break;
            if ((rc = tc::make_tmap_bytes(&g.tmA_lo, Al, 2, g.k_pad, Mp, g.k_pad, 32, 128)) != GPAD_OK) break;
            k.a_rowinv = ainv;
            rc = tc::launch_gemm(0, g, k, C, N, prop.multiProcessorCount, s);
        } else {
            // the TMEM-operand kernel quantises A itself; its warm-start launch (p_only) is a plain P = A B^T
            if ((rc = tc::plan_rings_p1(g.bn, prop.sharedMemPerBlockOptin, &g.a_stages, &g.stages, true)) != GPAD_OK) break;
            if ((rc = tc::launch_rowmax(Ap, g.k_pad, Mp, amax, s)) != GPAD_OK) break;
            if ((rc = tc::make_tmap(&g.tmA_hi, Ap, g.k_pad, Mp, g.k_pad, 32, 128)) != GPAD_OK) break;
            k.a_rowmax = amax;
            k.p_only = 1; k.np = N; k.P_cur = C;
            rc = tc::launch_p1(g, k, prop.multiProcessorCount, s);
        }
    } while (0);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(scratch);
    if (rc == GPAD_OK && e != cudaSuccess) return cuda_fail(e, "gpad_debug_gemm_f16x3", __FILE__, __LINE__);
    return rc;
}

}  // extern "C"
