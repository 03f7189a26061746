// group.cu -- one logical solver over several devices of one box (SURVEY 8b "one handle may span 1-8 devices", 8e).
// The batch is cut into contiguous shards, one per device; every shard has its own gpad handle (operators replicated,
// per-instance operators sharded with their instances) driven by its own host thread, and results land in the
// caller's host buffers through that handle's own asynchronous D2H copies.  The path has no exchange step, so there is
// no collective and no NCCL here: a shard's rows depend on nothing outside the shard, which is also why results do not
// depend on the number of devices.
#include <string>
#include <thread>
#include <vector>

#include "gpad_internal.h"

struct gpad_group_s {
    gpad_config_t cfg{};
    std::vector<int> devices;
    std::vector<gpad_handle_t> handles;
    int cap = 0;                 // capacity of every shard (multiple of 128)
};

namespace {

// shard i of a batch: shared operators balance the batch over the devices in multiples of 128 (the batch tile);
// per-instance operators are resident per device, so instance b belongs to shard b / cap
void shard_range(const gpad_group_s* g, int i, int batch, int* first, int* count) {
    const int D = (int)g->handles.size();
    int per = g->cap;
    if (g->cfg.mode != GPAD_MODE_BATCH_PER_INSTANCE) per = std::min(g->cap, gpad::round_up((batch + D - 1) / D, 128));
    const int lo = std::min(batch, i * per), hi = std::min(batch, (i + 1) * per);
    *first = lo; *count = hi - lo;
}

template <typename T> T* offset(T* p, size_t elems) { return p ? p + elems : nullptr; }

}  // namespace

extern "C" {

int gpad_group_setup(const gpad_config_t* cfg, const int* devices, int device_count, const float* M_G, const float* G_L,
                     gpad_group_t* out) {
    using namespace gpad;
    GPAD_REQUIRE(cfg && devices && device_count >= 1 && device_count <= 64 && M_G && G_L && out, "gpad_group_setup: bad argument");
    GPAD_REQUIRE(cfg->mode != GPAD_MODE_LATENCY || device_count == 1, "gpad_group_setup: latency mode solves one QP on one device");
    GPAD_REQUIRE(cfg->operators_mem == GPAD_MEM_HOST, "gpad_group_setup: operators must be in host memory");
    gpad_group_s* g = new gpad_group_s;
    g->cfg = *cfg;
    g->devices.assign(devices, devices + device_count);
    g->handles.assign(device_count, nullptr);
    g->cap = cfg->mode == GPAD_MODE_LATENCY ? 1 : round_up((cfg->max_batch + device_count - 1) / device_count, 128);
    const size_t per_op = (size_t)cfg->n_u * cfg->N * cfg->m;
    std::vector<int> rc(device_count, GPAD_OK);
    std::vector<std::string> msg(device_count);
    std::vector<std::thread> pool;
    for (int i = 0; i < device_count; ++i)
        pool.emplace_back([&, i]() {
            gpad_config_t c = *cfg;
            c.device = devices[i];
            c.max_batch = g->cap;
            const float *mg = M_G, *gl = G_L;
            if (cfg->mode == GPAD_MODE_BATCH_PER_INSTANCE) {
                const int first = std::min(cfg->max_batch, i * g->cap);
                c.max_batch = std::max(1, std::min(g->cap, cfg->max_batch - first));
                mg += (size_t)first * per_op; gl += (size_t)first * per_op;
                if (first >= cfg->max_batch) return;          // more devices than instances: this shard stays empty
            }
            rc[i] = gpad_setup(&c, mg, gl, &g->handles[i]);
            if (rc[i] != GPAD_OK) msg[i] = gpad_last_error();
        });
    for (auto& t : pool) t.join();
    for (int i = 0; i < device_count; ++i)
        if (rc[i] != GPAD_OK) {
            set_error("gpad_group_setup: shard %d (device %d): %s", i, devices[i], msg[i].c_str());
            const int r = rc[i];
            gpad_group_destroy(g);
            return r;
        }
    *out = g;
    return GPAD_OK;
}

int gpad_group_destroy(gpad_group_t g) {
    if (!g) return GPAD_OK;
    for (gpad_handle_t h : g->handles) gpad_destroy(h);
    delete g;
    return GPAD_OK;
}

int gpad_group_size(gpad_group_t g) { return g ? (int)g->handles.size() : 0; }

int gpad_group_shard(gpad_group_t g, int i, int batch, gpad_handle_t* h, int* first, int* count) {
    using namespace gpad;
    GPAD_REQUIRE(g && i >= 0 && i < (int)g->handles.size() && batch >= 0 && batch <= g->cfg.max_batch, "gpad_group_shard: bad argument");
    int f = 0, c = 0;
    shard_range(g, i, batch, &f, &c);
    if (h) *h = g->handles[i];
    if (first) *first = f;
    if (count) *count = c;
    return GPAD_OK;
}

int gpad_group_solve(gpad_group_t g, const gpad_solve_args_t* a) {
    using namespace gpad;
    GPAD_REQUIRE(g && a, "gpad_group_solve: null argument");
    GPAD_REQUIRE(a->mem == GPAD_MEM_HOST, "gpad_group_solve: host-memory arguments only (device buffers belong to one device)");
    GPAD_REQUIRE(a->batch >= 1 && a->batch <= g->cfg.max_batch, "gpad_group_solve: batch %d outside 1..%d", a->batch, g->cfg.max_batch);
    const int D = (int)g->handles.size();
    const size_t n = (size_t)g->cfg.n_u * g->cfg.N, m = (size_t)g->cfg.m;
    int n_par = 0;
    if (a->params) GPAD_REQUIRE(a->problem && gpad_problem_dims(a->problem, nullptr, nullptr, nullptr, &n_par, nullptr) == GPAD_OK,
                                "gpad_group_solve: params need their problem");
    std::vector<int> rc(D, GPAD_OK);
    std::vector<std::string> msg(D);
    std::vector<std::thread> pool;
    for (int i = 0; i < D; ++i) {
        int first = 0, count = 0;
        shard_range(g, i, a->batch, &first, &count);
        if (count <= 0 || !g->handles[i]) continue;
        pool.emplace_back([&, i, first, count]() {
            gpad_solve_args_t s = *a;
            const size_t f = (size_t)first;
            s.batch = count;
            s.g_P = offset(a->g_P, f * n); s.p_D = offset(a->p_D, f * m); s.f = offset(a->f, f * n);
            s.y0 = offset(a->y0, f * m); s.y_prev0 = offset(a->y_prev0, f * m);
            s.params = offset(a->params, f * (size_t)n_par);
            s.y_next = offset(a->y_next, f * m); s.y = offset(a->y, f * m); s.w = offset(a->w, f * m);
            s.z = offset(a->z, f * n); s.zhat = offset(a->zhat, f * n);
            s.iters = offset(a->iters, f); s.status = offset(a->status, f);
            s.max_viol = offset(a->max_viol, f); s.gap = offset(a->gap, f);
            rc[i] = gpad_solve(g->handles[i], &s);
            if (rc[i] != GPAD_OK) msg[i] = gpad_last_error();
        });
    }
    for (auto& t : pool) t.join();
    for (int i = 0; i < D; ++i)
        if (rc[i] != GPAD_OK) {
            set_error("gpad_group_solve: shard %d (device %d): %s", i, g->devices[i], msg[i].c_str());
            return rc[i];
        }
    return GPAD_OK;
}

}  // extern "C"
