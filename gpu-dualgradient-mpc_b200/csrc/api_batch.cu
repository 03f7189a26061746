// api_batch.cu -- throughput mode with shared operators (GPAD_MODE_BATCH_SHARED): setup, the host loop that drives
// the two GEMM kernels per iteration (main.cu:160-175 for a whole batch), and the synchronous / asynchronous solve
// entry points behind gpad_solve / gpad_solve_async / gpad_wait.
//
// State lives in one or two SLOTS (BatchSlot): a complete set of batch arrays with its tensor maps.  gpad_solve uses
// slot 0.  gpad_solve_async alternates between two slots and three streams, so that for back-to-back host-memory
// solves the H2D copies of solve k+1 and the D2H copies of solve k-1 run under the iterations of solve k.
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include "handle.h"

namespace gpad {

namespace {

// ---- product 2 tile width: measured once per (device, n, m, batch) in this process, not modelled: the landscape is
// irregular (64K quadrotor batch, ms per launch: 256 -> 0.74, 224 -> 0.93, 192 -> 0.69, 160 -> 0.66, 128 -> 0.81) ----
std::mutex g_tune_mutex;
std::map<std::tuple<int, int, int, int, int>, std::pair<int, std::string>> g_tune_cache;     // (device, n, m, batch, fp16 kernel)

// tensor maps over a slot's own state for the fp16 plans: y_v as fp32 tiles of 128 x 32 (quantised in-kernel), zhat hi / lo
// as fp16 tiles of 128 x 32
int make_f16_state_maps(BatchSlot& sl) {
    const BatchState& st = sl.st;
    for (int k = 0; k < 3; ++k) GPAD_TRY(tc::make_tmap(&sl.g1h.tmY[k], st.yb[k], st.mp, st.Bp, st.mp, 32, 128));
    GPAD_TRY(tc::make_tmap_bytes(&sl.g2h.tmA_hi, st.zq_hi, 2, st.np, st.Bp, st.np, 32, 128));
    GPAD_TRY(tc::make_tmap_bytes(&sl.g2h.tmA_lo, st.zq_lo, 2, st.np, st.Bp, st.np, 32, 128));
    // epilogue operands of product 2: only the m real columns exist for the maps (zero fill / clipped stores beyond)
    for (int k = 0; k < 3; ++k) GPAD_TRY(tc::make_tmap(&sl.g2h.tmEy[k], st.yb[k], st.m, st.Bp, st.mp, 32, 128));
    GPAD_TRY(tc::make_tmap(&sl.g2h.tmEpd, st.p_D, st.m, st.Bp, st.mp, 32, 128));
    return GPAD_OK;
}

int alloc_slot(gpad_handle_s* h, BatchSlot& sl, const BatchSlot* like) {
    const int n = h->n, m = h->cfg.m;
    const bool f16 = h->cfg.precision == GPAD_PREC_FP16X3;
    const bool tcp = h->cfg.precision == GPAD_PREC_TF32X3 || f16;
    BatchState& st = sl.st;
    st.n = n; st.m = m; st.np = round_up(n, 32); st.mp = round_up(m, 32);
    st.Bp = round_up(h->cfg.max_batch, 256);
    const size_t bm = (size_t)st.Bp * st.mp, bnn = (size_t)st.Bp * st.np;
    GPAD_TRY(dev_alloc(h, &st.g_P, bnn)); GPAD_TRY(dev_alloc(h, &st.p_D, bm)); GPAD_TRY(dev_alloc(h, &st.f, bnn));
    GPAD_TRY(dev_alloc(h, &st.yb[0], bm)); GPAD_TRY(dev_alloc(h, &st.yb[1], bm)); GPAD_TRY(dev_alloc(h, &st.yb[2], bm));
    GPAD_TRY(dev_alloc(h, &st.z, bnn)); GPAD_TRY(dev_alloc(h, &st.zhat, bnn)); GPAD_TRY(dev_alloc(h, &st.sbar, bm));
    GPAD_TRY(dev_alloc(h, &st.red, (size_t)st.Bp * kRedStride));
    GPAD_TRY(dev_alloc(h, &st.done, st.Bp)); GPAD_TRY(dev_alloc(h, &st.iters, st.Bp)); GPAD_TRY(dev_alloc(h, &st.status, st.Bp));
    GPAD_TRY(dev_alloc(h, &st.max_viol, st.Bp)); GPAD_TRY(dev_alloc(h, &st.gap, st.Bp));
    // the padding columns of the row-padded arrays are never written by a kernel; the strided DMA copies of host-memory
    // solves and the on-device instance build rely on them being zero
    GPAD_CUDA(cudaMemset(st.g_P, 0, bnn * sizeof(float))); GPAD_CUDA(cudaMemset(st.p_D, 0, bm * sizeof(float)));
    GPAD_CUDA(cudaMemset(st.f, 0, bnn * sizeof(float)));
    for (int k = 0; k < 3; ++k) GPAD_CUDA(cudaMemset(st.yb[k], 0, bm * sizeof(float)));
    GPAD_TRY(dev_alloc(h, &st.active_count, 2));
    GPAD_TRY(dev_alloc(h, &st.need, st.Bp)); GPAD_TRY(dev_alloc(h, &st.zy, bnn));
    GPAD_CUDA(cudaMemset(st.zy, 0, bnn * sizeof(float)));
    GPAD_TRY(dev_alloc(h, &st.tile_flags, st.Bp / 128)); GPAD_TRY(dev_alloc(h, &st.tile_list, st.Bp / 128));
    GPAD_TRY(dev_alloc(h, &st.tile_count, 1)); GPAD_TRY(dev_alloc(h, &st.stat, 2));
    GPAD_TRY(dev_alloc(h, &st.perm, st.Bp)); GPAD_TRY(dev_alloc(h, &st.holes, st.Bp)); GPAD_TRY(dev_alloc(h, &st.movers, st.Bp));
    GPAD_TRY(dev_alloc(h, &st.compact_counts, 2));
    GPAD_CUDA(cudaMemset(st.stat, 0, 2 * sizeof(unsigned long long)));
    if (tcp) {
        GPAD_TRY(dev_alloc(h, &st.zh_hi, bnn)); GPAD_TRY(dev_alloc(h, &st.zh_lo, bnn));
        GPAD_TRY(dev_alloc(h, &st.Pb[0], bnn)); GPAD_TRY(dev_alloc(h, &st.Pb[1], bnn));
        GPAD_CUDA(cudaMemset(st.zh_hi, 0, bnn * sizeof(float))); GPAD_CUDA(cudaMemset(st.zh_lo, 0, bnn * sizeof(float)));
        GPAD_CUDA(cudaMemset(st.zhat, 0, bnn * sizeof(float)));
        if (like) {             // second slot: same tiling, its own state maps
            sl.g1 = like->g1; sl.g2 = like->g2;
            for (int k = 0; k < 3; ++k) GPAD_TRY(tc::make_tmap(&sl.g1.tmY[k], st.yb[k], st.mp, st.Bp, st.mp, sl.g1.bk, 128));
            GPAD_TRY(tc::make_tmap(&sl.g2.tmA_hi, st.zh_hi, st.np, st.Bp, st.np, sl.g2.bk, 128));
            GPAD_TRY(tc::make_tmap(&sl.g2.tmA_lo, st.zh_lo, st.np, st.Bp, st.np, sl.g2.bk, 128));
        }
        if (f16) {
            GPAD_TRY(dev_alloc(h, &st.zq_hi, bnn)); GPAD_TRY(dev_alloc(h, &st.zq_lo, bnn));
            GPAD_CUDA(cudaMemset(st.zq_hi, 0, bnn * sizeof(uint16_t))); GPAD_CUDA(cudaMemset(st.zq_lo, 0, bnn * sizeof(uint16_t)));
            GPAD_TRY(dev_alloc(h, &st.zinv, st.Bp)); GPAD_TRY(dev_alloc(h, &st.ymax[0], st.Bp)); GPAD_TRY(dev_alloc(h, &st.ymax[1], st.Bp));
            GPAD_CUDA(cudaMemset(st.zinv, 0, st.Bp * sizeof(float)));
            GPAD_CUDA(cudaMemset(st.ymax[0], 0, st.Bp * sizeof(unsigned))); GPAD_CUDA(cudaMemset(st.ymax[1], 0, st.Bp * sizeof(unsigned)));
            if (like) {
                sl.g1h = like->g1h; sl.g2h = like->g2h;
                GPAD_TRY(make_f16_state_maps(sl));
            }
        }
    }
    GPAD_CUDA(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
    GPAD_CUDA(cudaEventCreateWithFlags(&sl.ev_comp, cudaEventDisableTiming));
    GPAD_CUDA(cudaEventCreateWithFlags(&sl.ev_out, cudaEventDisableTiming));
    // the memsets above ran on the legacy default stream, which the library's non-blocking streams do not wait for:
    // they must have landed before the first solve copies inputs into these arrays
    GPAD_CUDA(cudaDeviceSynchronize());
    sl.allocated = true;
    return GPAD_OK;
}

// output archive of compacted tolerance-mode solves: stopped instances' results by ORIGINAL index (batch_compact.cu)
int alloc_archive(gpad_handle_s* h) {
    if (h->arch_allocated) return GPAD_OK;
    const BatchState& st = h->slot[0].st;
    BatchState& ar = h->arch;
    ar = BatchState{};
    ar.n = st.n; ar.m = st.m; ar.np = st.np; ar.mp = st.mp; ar.Bp = st.Bp;
    const size_t bm = (size_t)st.Bp * st.mp, bnn = (size_t)st.Bp * st.np;
    for (int k = 0; k < 3; ++k) GPAD_TRY(dev_alloc(h, &ar.yb[k], bm));
    GPAD_TRY(dev_alloc(h, &ar.z, bnn)); GPAD_TRY(dev_alloc(h, &ar.zhat, bnn));
    GPAD_TRY(dev_alloc(h, &ar.iters, st.Bp)); GPAD_TRY(dev_alloc(h, &ar.status, st.Bp));
    GPAD_TRY(dev_alloc(h, &ar.max_viol, st.Bp)); GPAD_TRY(dev_alloc(h, &ar.gap, st.Bp));
    h->arch_allocated = true;
    return GPAD_OK;
}

BatchKernelArgs kernel_args(const gpad_handle_s* h, const BatchState& st, const gpad_solve_args_t* a, bool checking) {
    BatchKernelArgs k{};
    k.n = st.n; k.m = st.m; k.np = st.np; k.mp = st.mp; k.B = st.B; k.checking = checking ? 1 : 0; k.L = h->cfg.L;
    k.g_P = st.g_P; k.p_D = st.p_D; k.f = (a && (a->f || a->build_f)) ? st.f : nullptr;
    k.z = st.z; k.zhat = st.zhat; k.zh_hi = st.zh_hi; k.zh_lo = st.zh_lo;
    k.sbar = st.sbar; k.red = st.red; k.done = checking ? st.done : nullptr;
    const bool retire = checking && h->knobs.tc_retire;
    k.tile_list = retire ? st.tile_list : nullptr;
    k.tile_count = retire ? st.tile_count : nullptr;
    k.dual_count = st.active_count + 1;
    return k;
}

// user vector [B][len] (host or device) -> padded device rows [.][ld]; null src -> zeros.  Host data go by strided DMA
// copy straight into the padded rows (no SM work in the copy path; the padding columns stay zero from setup).
int ingest(gpad_handle_s* h, const BatchState& st, float* dst, int ld, const float* src, int len, bool host, cudaStream_t s) {
    const int B = st.B;
    if (!src) { GPAD_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)B * ld, s)); return GPAD_OK; }
    if (host) {
        GPAD_CUDA(cudaMemcpy2DAsync(dst, sizeof(float) * ld, src, sizeof(float) * len, sizeof(float) * len, B, cudaMemcpyHostToDevice, s));
        return GPAD_OK;
    }
    GPAD_TRY(launch_pad_rows(dst, ld, B, src, len, B, s));
    h->launches += 1;
    return GPAD_OK;
}

int inputs(gpad_handle_s* h, BatchSlot& sl, const gpad_solve_args_t* a, bool host, cudaStream_t s) {
    const BatchState& st = sl.st;
    if (a->params) {
        // g_P = Kg p, p_D = -(b0 + Bb p) / L (and f = Ff p) evaluated on the device from the parameters
        // (acceldualgrad.m:21,23): 8 n_par bytes per instance cross PCIe instead of 4 (n + m)
        int n_par = 0;
        GPAD_TRY(gpad_problem_dims(a->problem, nullptr, nullptr, nullptr, &n_par, nullptr));
        const double* pd = a->params;
        if (host) {
            if (h->params_cap < h->cfg.max_batch * n_par) {
                GPAD_TRY(dev_alloc(h, &h->d_params, (size_t)h->cfg.max_batch * n_par));
                h->params_cap = h->cfg.max_batch * n_par;
            }
            GPAD_CUDA(cudaMemcpyAsync(h->d_params, a->params, sizeof(double) * (size_t)st.B * n_par, cudaMemcpyHostToDevice, s));
            pd = h->d_params;
        }
        GPAD_TRY(instances_device(a->problem, h->device, st.B, pd, st.g_P, st.np, st.p_D, st.mp, a->build_f ? st.f : nullptr, st.np, s));
        h->launches += a->build_f ? 3 : 2;
    } else {
        GPAD_TRY(ingest(h, st, st.g_P, st.np, a->g_P, st.n, host, s));
        GPAD_TRY(ingest(h, st, st.p_D, st.mp, a->p_D, st.m, host, s));
        if (a->f) GPAD_TRY(ingest(h, st, st.f, st.np, a->f, st.n, host, s));
    }
    GPAD_TRY(ingest(h, st, st.yb[0], st.mp, a->y0, st.m, host, s));         // y_0
    GPAD_TRY(ingest(h, st, st.yb[2], st.mp, a->y_prev0, st.m, host, s));    // y_{-1}
    return GPAD_OK;
}

// `it`: the iteration (selects the rotating y buffers for the TMA-fed epilogue of the fp16 product 2)
int launch_product(gpad_handle_s* h, BatchSlot& sl, int phase, const BatchKernelArgs& k, cudaStream_t s, bool f16 = false, int it = 0) {
    if (f16 && phase == 1) return tc::launch_p1(sl.g1h, k, h->num_sms, s);
    if (f16) return tc::launch_p2(sl.g2h, k, it % 3, (it + 2) % 3, (it + 1) % 3, h->num_sms, s);
    if (h->cfg.precision == GPAD_PREC_TF32X3 || h->cfg.precision == GPAD_PREC_FP16X3) {
        if (phase == 1 && sl.g1.p1) return tc::launch_p1(sl.g1, k, h->num_sms, s);
        return tc::launch_gemm(phase, phase == 1 ? sl.g1 : sl.g2, k, nullptr, 0, h->num_sms, s);
    }
    return launch_simt_product(phase, h->op, k, round_up(k.B, 128), s);
}

// the iterations of one solve (everything between the input and the output copies)
int iterate(gpad_handle_s* h, BatchSlot& sl, const gpad_solve_args_t* a, cudaStream_t s) {
    BatchState& st = sl.st;
    const int B = st.B;
    const bool checking = a->check_every > 0;
    const bool f16 = h->cfg.precision == GPAD_PREC_FP16X3 && !checking;      // tolerance mode runs the tf32 kernels
    const bool tcp = h->cfg.precision == GPAD_PREC_TF32X3 || h->cfg.precision == GPAD_PREC_FP16X3;
    const bool have_f = a->f != nullptr || a->build_f;
    // compaction needs the tile list (the compacted layout is only visible to the kernels through it)
    const bool compacting = checking && h->knobs.tc_retire && h->knobs.tc_compact && B > 256;
    h->out_from_archive = false;
    h->compactions = 0;
    if (compacting) { GPAD_TRY(alloc_archive(h)); GPAD_TRY(launch_perm_identity(st, s)); h->launches += 1; }
    int work_rows = B;                 // upper bound of the rows that still hold a live instance
    int last_running = B;              // newest stop count the host has seen
    GPAD_TRY(launch_batch_init(st, checking, s));
    GPAD_TRY(launch_batch_reset_term(st, a->max_iter, s));
    h->launches += 2;

    BatchKernelArgs k = kernel_args(h, st, a, checking);
    const int m_tiles = round_up(B, 128) / 128;
    sl.g1.m_tiles = m_tiles; sl.g2.m_tiles = m_tiles;
    sl.g1h.m_tiles = m_tiles; sl.g2h.m_tiles = m_tiles;
    // fp16 plan: three launches per iteration; below ~16K instances the launch boundaries are 4 % of a solve (measured: 8K
    // instances 15.3 -> 14.7 ms, 64K within noise), so small solves launch their kernels programmatically dependent
    sl.g1h.pdl = sl.g2h.pdl = (h->knobs.tc_pdl > 0 || (h->knobs.tc_pdl < 0 && m_tiles <= 128)) ? 1 : 0;
    const int q_rows = m_tiles * 128;          // rows the quantisation kernels cover: whole batch tiles
    if (f16 && a->max_iter > 0) {
        // row maxima of y_0 (product 2 reduces those of every later iterate)
        if (a->y0) { GPAD_TRY(tc::launch_rowmax(st.yb[0], st.mp, q_rows, reinterpret_cast<float*>(st.ymax[0]), s)); h->launches += 1; }
        else GPAD_CUDA(cudaMemsetAsync(st.ymax[0], 0, sizeof(unsigned) * q_rows, s));
    }
    // programmatic dependent launch: a GEMM kernel's CTAs start while the previous kernel drains; nothing it reads from
    // global memory (tile lists and counters included, see TileSched) is read before its dependency wait
    sl.g1.pdl = sl.g2.pdl = h->knobs.tc_pdl > 0 ? 1 : 0;
    sl.g1.cluster_attr = sl.g2.cluster_attr = h->knobs.tc_cluster_attr ? 1 : 0;
    if (tcp && a->max_iter > 0) {
        if (a->y_prev0) {          // warm start: P_{-1} = M_G y_{-1} (one extra product-1 launch)
            BatchKernelArgs kp = k;
            kp.p_only = 1; kp.P_cur = st.Pb[1]; kp.P_prev = st.Pb[0];
            kp.tile_list = nullptr; kp.tile_count = nullptr;
            sl.g1.tmA_hi = sl.g1.tmY[2]; sl.g1.tmA_lo = sl.g1.tmY[2];
            if (f16) {
                GPAD_TRY(tc::launch_rowmax(st.yb[2], st.mp, q_rows, reinterpret_cast<float*>(st.ymax[1]), s));
                h->launches += 1;
                sl.g1h.tmA_hi = sl.g1h.tmY[2];
                kp.a_rowmax = reinterpret_cast<const float*>(st.ymax[1]); kp.b_colinv = h->op.M_G_inv;
            }
            GPAD_TRY(launch_product(h, sl, 1, kp, s, f16));
            h->launches += 1;
        } else {
            GPAD_CUDA(cudaMemsetAsync(st.Pb[1], 0, sizeof(float) * (size_t)st.Bp * st.np, s));
        }
    }

    // tolerance mode: the host runs up to `lag` decisions ahead of the device.  Stopped instances are frozen (their rows
    // are skipped, their tiles retired), so iterations enqueued past the point where everything has stopped change nothing.
    const int lag = std::max(0, std::min(h->knobs.check_lag, (int)h->ev_check.size() - 1));
    const int ring = (int)h->ev_check.size();
    int checks_issued = 0, checks_seen = 0, since_check = 0;
    bool all_done = false;
    auto poll = [&](bool drain) -> int {
        while (checks_seen < checks_issued) {
            cudaEvent_t e = h->ev_check[checks_seen % ring];
            if (drain || checks_issued - checks_seen > lag) GPAD_CUDA(cudaEventSynchronize(e));
            else {
                cudaError_t q = cudaEventQuery(e);
                if (q == cudaErrorNotReady) break;
                GPAD_CUDA(q);
            }
            last_running = h->h_active[2 * (checks_seen % ring)];
            if (last_running <= 0) all_done = true;
            ++checks_seen;
        }
        return GPAD_OK;
    };

    for (int it = 0; it < a->max_iter && !all_done; ++it) {
        const bool check = checking && ((it + 1) % a->check_every == 0);
        k.it.theta = a->theta[it];
        k.it.beta = a->beta[it];
        k.it.check = check ? 1 : 0;
        k.it.store_zhat = (check || it + 1 == a->max_iter) ? 1 : 0;      // instances stop at checks only
        k.y_prev = st.yb[(it + 2) % 3];           // y_{v-1}
        k.y_cur = st.yb[it % 3];                  // y_v
        k.y_next = st.yb[(it + 1) % 3];           // y_{v+1} overwrites y_{v-2}
        k.P_cur = st.Pb[it & 1]; k.P_prev = st.Pb[(it + 1) & 1];
        if (tcp) { sl.g1.tmA_hi = sl.g1.tmY[it % 3]; sl.g1.tmA_lo = sl.g1.tmY[it % 3]; }
        if (f16) {
            sl.g1h.tmA_hi = sl.g1h.tmY[it % 3];
            k.a_rowmax = reinterpret_cast<const float*>(st.ymax[it & 1]);
            k.b_colinv = h->op.M_G_inv;
        }
        cudaEvent_t pe = h->prof_begin(s);
        GPAD_TRY(launch_product(h, sl, 1, k, s, f16));
        h->prof_end(1, pe, s);
        if (f16) {
            // zhat_v -> row scale, fp16 hi / lo; also clears the row maxima product 2 is about to reduce
            pe = h->prof_begin(s);
            GPAD_TRY(tc::launch_quantize_rows(st.zhat, st.np, q_rows, st.zq_hi, st.zq_lo, st.zinv, st.ymax[(it + 1) & 1], s,
                                              sl.g1h.pdl != 0));
            h->prof_end(0, pe, s);
            h->launches += 1;
            k.a_rowinv = st.zinv; k.b_colinv = h->op.G_L_inv; k.next_rowmax = st.ymax[(it + 1) & 1];
        }
        pe = h->prof_begin(s);
        GPAD_TRY(launch_product(h, sl, 2, k, s, f16, it));
        h->prof_end(2, pe, s);
        h->launches += 2;
        ++since_check;
        if (!check) continue;
        GPAD_TRY(launch_batch_decide(st, it + 1, h->cfg.L, a->eps_g, a->eps_V, have_f, s));
        h->launches += 1;
        if (have_f) {
            // dual-gap branch: z_y = M_G y_{v+1} - g_P and G_L z_y for the flagged instances (two more products, whose
            // CTAs return at once when no instance is flagged)
            BatchKernelArgs kd = k;
            kd.dual = 1; kd.need = st.need; kd.zy = st.zy; kd.p_only = 0;
            kd.it.check = 1; kd.it.beta = 0.f; kd.it.theta = 0.f; kd.it.store_zhat = 0;
            kd.y_cur = k.y_next; kd.y_prev = k.y_next;      // beta = 0: w = y_{v+1}
            if (tcp) { sl.g1.tmA_hi = sl.g1.tmY[(it + 1) % 3]; sl.g1.tmA_lo = sl.g1.tmY[(it + 1) % 3]; }
            GPAD_TRY(launch_product(h, sl, 1, kd, s));     // (reads the decision kernel's counter after its dependency wait)
            GPAD_TRY(launch_product(h, sl, 2, kd, s));
            GPAD_TRY(launch_batch_decide_dual(st, it + 1, h->cfg.L, a->eps_V, s));
            h->launches += 3;
        }
        if (compacting && last_running > 0 && work_rows > 256 && work_rows - last_running >= std::max(128, work_rows / 5)) {
            // a fifth of the working rows (and at least one tile's worth) has stopped, by a count that can only be stale on
            // the high side: archive the stopped rows, move the running ones into the holes below
            GPAD_TRY(launch_compact(st, h->arch, work_rows, have_f, s));
            h->launches += 3;
            work_rows = last_running;
            h->out_from_archive = true;
            h->compactions += 1;
        }
        GPAD_TRY(launch_batch_tiles(st, since_check, h->knobs.tc_retire != 0, s));
        h->launches += h->knobs.tc_retire ? 2 : 1;
        since_check = 0;
        const int slot = checks_issued % ring;
        GPAD_CUDA(cudaMemcpyAsync(h->h_active + 2 * slot, st.active_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        GPAD_CUDA(cudaEventRecord(h->ev_check[slot], s));
        ++checks_issued;
        GPAD_TRY(poll(false));
    }
    if (checking) {
        GPAD_TRY(poll(true));
        if (since_check > 0) { GPAD_TRY(launch_batch_tiles(st, since_check, false, s)); h->launches += 1; }
        if (h->out_from_archive) {
            GPAD_TRY(launch_archive_all(st, h->arch, work_rows, s));
            h->launches += 1;
        }
        BatchState sum = st;
        if (h->out_from_archive) sum.iters = h->arch.iters;
        GPAD_TRY(launch_batch_iter_sum(sum, s));
        h->launches += 1;
    } else if (a->max_iter > 0) {
        GPAD_TRY(launch_batch_finite(st, st.yb[a->max_iter % 3], s));
        h->launches += 1;
    }
    return GPAD_OK;
}

int emit(gpad_handle_s* h, const BatchState& st, float* dst, int len, const float* src, int ld, bool host, cudaStream_t s) {
    if (!dst) return GPAD_OK;
    if (host) {
        GPAD_CUDA(cudaMemcpy2DAsync(dst, sizeof(float) * len, src, sizeof(float) * ld, sizeof(float) * len, st.B, cudaMemcpyDeviceToHost, s));
        return GPAD_OK;
    }
    GPAD_TRY(launch_unpad_rows(dst, len, st.B, src, ld, s));
    h->launches += 1;
    return GPAD_OK;
}

int outputs(gpad_handle_s* h, BatchSlot& sl, const gpad_solve_args_t* a, bool host, float* stage, cudaStream_t s) {
    BatchState st = sl.st;
    if (h->out_from_archive && a->check_every > 0) {
        // compacted solve: results were archived by original instance index
        for (int k = 0; k < 3; ++k) st.yb[k] = h->arch.yb[k];
        st.z = h->arch.z; st.zhat = h->arch.zhat;
        st.iters = h->arch.iters; st.status = h->arch.status; st.max_viol = h->arch.max_viol; st.gap = h->arch.gap;
    }
    const int n = st.n, m = st.m, B = st.B;
    float* yo[3] = {a->y_next, a->y, a->w};
    if (host && a->check_every <= 0) {
        // fixed iteration count I for every instance: y_I sits in yb[I % 3], y_{I-1} in yb[(I + 2) % 3]: strided DMA copies;
        // only w_{I-1} needs arithmetic (one kernel into the staging buffer)
        const int I = a->max_iter;
        GPAD_TRY(emit(h, st, yo[0], m, st.yb[I % 3], st.mp, true, s));
        GPAD_TRY(emit(h, st, yo[1], m, st.yb[(I + 2) % 3], st.mp, true, s));
        if (yo[2]) {
            GPAD_TRY(launch_unpad_y(nullptr, nullptr, stage, m, B, st.yb[0], st.yb[1], st.yb[2], st.mp, st.iters, h->d_beta, s));
            GPAD_CUDA(cudaMemcpyAsync(yo[2], stage, sizeof(float) * (size_t)B * m, cudaMemcpyDeviceToHost, s));
            h->launches += 1;
        }
    } else if (host) {      // per-instance iteration counts: one vector at a time through the staging buffer
        for (int k3 = 0; k3 < 3; ++k3) {
            if (!yo[k3]) continue;
            GPAD_TRY(launch_unpad_y(k3 == 0 ? stage : nullptr, k3 == 1 ? stage : nullptr, k3 == 2 ? stage : nullptr,
                                    m, B, st.yb[0], st.yb[1], st.yb[2], st.mp, st.iters, h->d_beta, s));
            GPAD_CUDA(cudaMemcpyAsync(yo[k3], stage, sizeof(float) * (size_t)B * m, cudaMemcpyDeviceToHost, s));
            h->launches += 1;
        }
    } else if (yo[0] || yo[1] || yo[2]) {
        GPAD_TRY(launch_unpad_y(yo[0], yo[1], yo[2], m, B, st.yb[0], st.yb[1], st.yb[2], st.mp, st.iters, h->d_beta, s));
        h->launches += 1;
    }
    GPAD_TRY(emit(h, st, a->z, n, st.z, st.np, host, s));
    GPAD_TRY(emit(h, st, a->zhat, n, st.zhat, st.np, host, s));
    const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (a->iters) GPAD_CUDA(cudaMemcpyAsync(a->iters, st.iters, sizeof(int) * B, kind, s));
    if (a->status) GPAD_CUDA(cudaMemcpyAsync(a->status, st.status, sizeof(int) * B, kind, s));
    if (a->max_viol) GPAD_CUDA(cudaMemcpyAsync(a->max_viol, st.max_viol, sizeof(float) * B, kind, s));
    if (a->gap) GPAD_CUDA(cudaMemcpyAsync(a->gap, st.gap, sizeof(float) * B, kind, s));
    if (a->check_every > 0)
        GPAD_CUDA(cudaMemcpyAsync(h->h_stat, sl.st.stat, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    return GPAD_OK;
}

int wait_all_async(gpad_handle_s* h) {
    for (BatchSlot& sl : h->slot)
        if (sl.allocated && sl.ticket >= 0) { GPAD_CUDA(cudaEventSynchronize(sl.ev_out)); sl.ticket = -1; }
    return GPAD_OK;
}

}  // namespace

// ------------------------------------------------------------------ setup
int setup_batch(gpad_handle_s* h, const std::vector<float>& MG, const std::vector<float>& GL) {
    const int n = h->n, m = h->cfg.m;
    const bool f16 = h->cfg.precision == GPAD_PREC_FP16X3;
    const bool tcp = h->cfg.precision == GPAD_PREC_TF32X3 || f16;
    const Knobs& kn = h->knobs;
    BatchSlot& sl = h->slot[0];
    const int np = round_up(n, 32), mp = round_up(m, 32);
    int bn1 = 0, nt1 = 0, bn2 = 0, nt2 = 0, step1 = 0;
    tc::plan_tiles(n, &bn1, &nt1);
    tc::plan_tiles(m, &bn2, &nt2);
    // product 1: the TMEM-operand kernel costs ~824 clk per 128-row tile and k-block whatever the tile width (<= 208 columns),
    // the shared-memory-operand kernel ~940 clk at 208 columns, growing with the width (<= 256) -- but it may need fewer
    // tiles.  With few tiles the number of waves over the SMs decides (battery (10,100), 4096 QPs: 160 tiles = 2 waves
    // against 128 tiles = 1 wave: measured 98.7 k against 112 k solves/s), with many tiles the per-tile cost does.
    bool p1 = tcp;
    if (f16) p1 = true;       // the fp16 product 1 exists as the TMEM-operand kernel only: one tiling for both families
    else if (tcp && kn.tc_p1 >= 0) p1 = kn.tc_p1 != 0;
    else if (tcp) {
        int bn_ts = 0, nt_ts = 0;
        tc::plan_tiles_p1(n, &bn_ts, &nt_ts);
        const int mt = (h->cfg.max_batch + 127) / 128;
        const double cost_ts = std::ceil((double)mt * nt_ts / h->num_sms) * 824.0;
        const double cost_ss = std::ceil((double)mt * nt1 / h->num_sms) * 940.0 * bn1 / 208.0;
        p1 = cost_ts <= cost_ss;
    }
    if (p1) tc::plan_tiles_p1(n, &bn1, &nt1, &step1);
    // fp16 product 1: when wider tiles (<= 256 columns, ONE accumulator stage) bring the whole launch into a single wave over
    // the SMs and the default plan needs more, take them -- a single wave has no next tile to overlap the epilogue with
    // (battery (10,100), 4096 QPs: 160 tiles of 208 = 2 waves against 128 tiles of 256 = 1 wave)
    int bn1h = bn1, nt1h = nt1, step1h = step1, acc1h = 2;
    if (f16) {
        int bw = 0, nw = 0, sw = 0;
        tc::plan_tiles_p1(n, &bw, &nw, &sw, 256);
        const int mt = (h->cfg.max_batch + 127) / 128;
        if (mt * nw <= h->num_sms && mt * nt1 > h->num_sms) { bn1h = bw; nt1h = nw; step1h = sw; acc1h = 1; }
    }
    h->op.n_rows_pad = round_up(std::max(std::max(bn1 * nt1, bn1h * nt1h), n), 128);
    h->op.m_rows_pad = round_up(m + 256, 128);      // any product-2 tiling of width <= 256 stays inside
    GPAD_TRY(upload_padded(h, MG.data(), n, m, h->op.n_rows_pad, mp, &h->op.M_G));
    GPAD_TRY(upload_padded(h, GL.data(), m, n, h->op.m_rows_pad, np, &h->op.G_L));
    GPAD_TRY(alloc_slot(h, sl, nullptr));
    BatchState& st = sl.st;
    GPAD_TRY(dev_alloc(h, &h->stage_in, (size_t)h->cfg.max_batch * std::max(n, m)));
    const int ring = std::max(1, kn.check_lag) + 1;
    GPAD_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->h_active), 2 * ring * sizeof(int)));
    GPAD_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->h_stat), 2 * sizeof(unsigned long long)));
    h->h_stat[0] = h->h_stat[1] = 0;
    h->ev_check.resize(ring);
    for (cudaEvent_t& e : h->ev_check) GPAD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    char buf[768];
    if (!tcp) {
        h->desc = "batch-shared: CUDA-core fp32 GEMM 128x128x16 tiles with fused GPAD epilogues";
        return GPAD_OK;
    }
    const size_t c1 = (size_t)h->op.n_rows_pad * mp, c2 = (size_t)h->op.m_rows_pad * np;
    if (f16) {
        // fp16 hi / lo of the row-scaled operators, from the fp32 values (before the tf32 split below rewrites them)
        GPAD_TRY(dev_alloc(h, &h->op.M_Gq_hi, c1)); GPAD_TRY(dev_alloc(h, &h->op.M_Gq_lo, c1));
        GPAD_TRY(dev_alloc(h, &h->op.G_Lq_hi, c2)); GPAD_TRY(dev_alloc(h, &h->op.G_Lq_lo, c2));
        GPAD_TRY(dev_alloc(h, &h->op.M_G_inv, h->op.n_rows_pad)); GPAD_TRY(dev_alloc(h, &h->op.G_L_inv, h->op.m_rows_pad));
        GPAD_TRY(tc::launch_quantize_rows(h->op.M_G, mp, h->op.n_rows_pad, h->op.M_Gq_hi, h->op.M_Gq_lo, h->op.M_G_inv, nullptr, h->own_stream));
        GPAD_TRY(tc::launch_quantize_rows(h->op.G_L, np, h->op.m_rows_pad, h->op.G_Lq_hi, h->op.G_Lq_lo, h->op.G_L_inv, nullptr, h->own_stream));
        GPAD_CUDA(cudaStreamSynchronize(h->own_stream));
    }
    GPAD_TRY(dev_alloc(h, &h->op.M_G_lo, c1)); GPAD_TRY(dev_alloc(h, &h->op.G_L_lo, c2));
    GPAD_TRY(tc::launch_split(h->op.M_G, h->op.M_G, h->op.M_G_lo, c1, h->own_stream));
    GPAD_TRY(tc::launch_split(h->op.G_L, h->op.G_L, h->op.G_L_lo, c2, h->own_stream));
    GPAD_CUDA(cudaStreamSynchronize(h->own_stream));
    const int bk = 16;
    tc::GemmDesc& g1 = sl.g1; tc::GemmDesc& g2 = sl.g2;
    g1.bk = g2.bk = bk;
    g1.k_pad = mp; g1.bn = bn1; g1.n_tiles = nt1; g1.ncols_valid = n;
    g2.k_pad = np; g2.bn = bn2; g2.n_tiles = nt2; g2.ncols_valid = m;
    auto cap_stages = [&](int s) { return kn.tc_stages > 0 ? std::min(s, std::max(2, kn.tc_stages)) : s; };
    g1.stages = cap_stages(tc::pick_stages(bk, bn1, h->smem_optin));
    g2.stages = cap_stages(tc::pick_stages(bk, bn2, h->smem_optin));
    if (p1) {
        g1.p1 = 1; g1.bk = 16; g1.step = step1;
        GPAD_TRY(tc::plan_rings_p1(bn1, h->smem_optin, &g1.a_stages, &g1.stages));
    }
    for (int k = 0; k < 3; ++k) GPAD_TRY(tc::make_tmap(&g1.tmY[k], st.yb[k], mp, st.Bp, mp, g1.bk, 128));
    GPAD_TRY(tc::make_tmap(&g1.tmB_hi, h->op.M_G, mp, h->op.n_rows_pad, mp, g1.bk, bn1));
    GPAD_TRY(tc::make_tmap(&g1.tmB_lo, h->op.M_G_lo, mp, h->op.n_rows_pad, mp, g1.bk, bn1));
    GPAD_TRY(tc::make_tmap(&g2.tmA_hi, st.zh_hi, np, st.Bp, np, g2.bk, 128));
    GPAD_TRY(tc::make_tmap(&g2.tmA_lo, st.zh_lo, np, st.Bp, np, g2.bk, 128));
    // product 2 tiling of width b for the tf32 (half = false) or the fp16 plan
    auto config_g2 = [&](tc::GemmDesc& g, bool half, int b) -> int {
        g.bn = b; g.n_tiles = (m + b - 1) / b;
        g.stages = cap_stages(tc::pick_stages(bk, b, h->smem_optin));
        if (half && g.p2) {
            GPAD_TRY(tc::plan_rings_p2(b, h->smem_optin, &g.stages, &g.e_stages, kn.tc_stages));
        }
        if (half) {
            GPAD_TRY(tc::make_tmap_bytes(&g.tmB_hi, h->op.G_Lq_hi, 2, np, h->op.m_rows_pad, np, 32, b));
            GPAD_TRY(tc::make_tmap_bytes(&g.tmB_lo, h->op.G_Lq_lo, 2, np, h->op.m_rows_pad, np, 32, b));
        } else {
            GPAD_TRY(tc::make_tmap(&g.tmB_hi, h->op.G_L, np, h->op.m_rows_pad, np, g.bk, b));
            GPAD_TRY(tc::make_tmap(&g.tmB_lo, h->op.G_L_lo, np, h->op.m_rows_pad, np, g.bk, b));
        }
        return GPAD_OK;
    };
    // ---- product 2 tile width: timed once per (device, n, m, batch, kernel family) in this process on the handle's own
    // zeroed buffers and own stream (first launch untimed, three timed); results do not depend on the width: every output
    // element sums over K in the same order.  GPAD_DEBUG tc_bn2=<w> fixes it, tc_autotune=0 keeps the first candidate. ----
    auto tune_g2 = [&](tc::GemmDesc& g, bool half, int bn_default, std::string& tune_note) -> int {
        GPAD_TRY(config_g2(g, half, bn_default));
        if (kn.tc_bn2 > 0) {
            GPAD_TRY(config_g2(g, half, std::max(g.p2 ? 32 : 16, std::min(256, kn.tc_bn2 / (g.p2 ? 32 : 16) * (g.p2 ? 32 : 16)))));
            tune_note = "fixed by tc_bn2";
            return GPAD_OK;
        }
        if (h->cfg.max_batch < 1024 || !kn.tc_autotune) return GPAD_OK;
        const auto key = std::make_tuple(h->device, n, m, round_up(h->cfg.max_batch, 128), half ? 1 : 0);
        std::lock_guard<std::mutex> lock(g_tune_mutex);
        auto hit = g_tune_cache.find(key);
        if (hit != g_tune_cache.end()) {
            GPAD_TRY(config_g2(g, half, hit->second.first));
            tune_note = hit->second.second + " (cached)";
            return GPAD_OK;
        }
        st.B = h->cfg.max_batch;
        BatchKernelArgs k = kernel_args(h, st, nullptr, false);
        k.y_prev = st.yb[2]; k.y_cur = st.yb[0]; k.y_next = st.yb[1];
        k.it.theta = 1.f; k.it.beta = 0.f;
        k.a_rowinv = st.zinv; k.b_colinv = h->op.G_L_inv; k.next_rowmax = st.ymax[1];
        cudaEvent_t e0, e1;
        GPAD_CUDA(cudaEventCreate(&e0)); GPAD_CUDA(cudaEventCreate(&e1));
        float best_ms = 1e30f; int best_bn = bn_default;
        std::vector<int> cand = {bn_default};
        // multiples of 32 columns: every 32-column block of the epilogue then starts on a 128-byte line (widths that
        // are only multiples of 16 -- 144, 176, 208, 240 -- measured 0.73 .. 1.04 ms against 0.66 .. 0.69 for 160 / 192)
        for (int b : {256, 224, 192, 160, 128}) if (b != bn_default && (m + b - 1) / b <= 64) cand.push_back(b);
        for (int b : cand) {
            if (config_g2(g, half, b) != GPAD_OK) continue;
            g.m_tiles = (h->cfg.max_batch + 127) / 128;
            float ms = 1e30f;
            bool ok = true;
            for (int rep = 0; rep < 4 && ok; ++rep) {
                if (rep == 1) cudaEventRecord(e0, h->own_stream);
                ok = (g.p2 ? tc::launch_p2(g, k, 0, 2, 1, h->num_sms, h->own_stream)
                           : tc::launch_gemm(2, g, k, nullptr, 0, h->num_sms, h->own_stream)) == GPAD_OK;
            }
            cudaEventRecord(e1, h->own_stream);
            if (cudaEventSynchronize(e1) != cudaSuccess || !ok) { cudaGetLastError(); continue; }
            cudaEventElapsedTime(&ms, e0, e1);
            char t[48];
            snprintf(t, sizeof(t), "%s%d:%.3f", tune_note.empty() ? "" : " ", b, ms / 3.0f);
            tune_note += t;
            if (ms < best_ms) { best_ms = ms; best_bn = b; }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        GPAD_TRY(config_g2(g, half, best_bn));
        GPAD_CUDA(cudaMemset(st.yb[1], 0, (size_t)st.Bp * mp * sizeof(float)));      // the timed launches wrote y_next
        if (half) GPAD_CUDA(cudaMemset(st.ymax[1], 0, st.Bp * sizeof(unsigned)));
        tune_note = "timed, ms per launch by width: " + tune_note;
        g_tune_cache[key] = {best_bn, tune_note};
        return GPAD_OK;
    };
    std::string tune_note, tune_note_h;
    GPAD_TRY(tune_g2(g2, false, bn2, tune_note));
    bn2 = g2.bn; nt2 = g2.n_tiles;
    if (f16) {
        tc::GemmDesc& g1h = sl.g1h; tc::GemmDesc& g2h = sl.g2h;
        g1h = tc::GemmDesc{}; g2h = tc::GemmDesc{};
        g1h.f16 = g2h.f16 = 1; g1h.bk = g2h.bk = bk;
        g1h.k_pad = mp; g1h.bn = bn1h; g1h.n_tiles = nt1h; g1h.ncols_valid = n; g1h.p1 = 1; g1h.step = step1h; g1h.acc_stages = acc1h;
        GPAD_TRY(tc::plan_rings_p1(bn1h, h->smem_optin, &g1h.a_stages, &g1h.stages, true));
        g2h.k_pad = np; g2h.ncols_valid = m;
        GPAD_TRY(make_f16_state_maps(sl));
        GPAD_TRY(tc::make_tmap_bytes(&g1h.tmB_hi, h->op.M_Gq_hi, 2, mp, h->op.n_rows_pad, mp, 32, bn1h));
        GPAD_TRY(tc::make_tmap_bytes(&g1h.tmB_lo, h->op.M_Gq_lo, 2, mp, h->op.n_rows_pad, mp, 32, bn1h));
        int bnh = 0, nth = 0;
        g2h.p2 = 1;
        tc::plan_tiles_p2(m, &bnh, &nth);
        GPAD_TRY(tune_g2(g2h, true, bnh, tune_note_h));
    }
    if (p1)
        snprintf(buf, sizeof(buf),
                 "batch-shared: tcgen05 cta_group::1 kind::tf32 x3; product1 = P-formulation (A = y_v only, split in registers, A operand "
                 "through a TMEM ring, state ring %d x 8 KB + operator ring %d stages), tiles 128x%d x%d; product2 tiles 128x%d x%d "
                 "(%d stages, TMA ring bk=%d; width %s), TMEM 512 cols, persistent over %d SMs, programmatic dependent launch %s",
                 g1.a_stages, g1.stages, bn1, nt1, bn2, nt2, g2.stages, bk, tune_note.empty() ? "by plan" : tune_note.c_str(),
                 h->num_sms, kn.tc_pdl > 0 ? "on" : "off");
    else
        snprintf(buf, sizeof(buf),
                 "batch-shared: tcgen05 cta_group::1 kind::tf32 x3 (P-formulation, y_v split in shared memory), TMA ring bk=%d, product1 tiles "
                 "128x%d x%d (%d stages), product2 tiles 128x%d x%d (%d stages; width %s), TMEM 2x256 cols, persistent over %d SMs, "
                 "programmatic dependent launch %s",
                 bk, bn1, nt1, g1.stages, bn2, nt2, g2.stages, tune_note.empty() ? "by plan" : tune_note.c_str(), h->num_sms,
                 kn.tc_pdl > 0 ? "on" : "off");
    h->desc = buf;
    if (f16) {
        snprintf(buf, sizeof(buf),
                 "batch-shared, fixed-iteration solves: tcgen05 cta_group::1 kind::f16 x3 (fp16 hi/lo of power-of-two row-scaled operands, "
                 "scales undone on the fp32 accumulator); product1 = P-formulation, y_v quantised in registers into a TMEM A ring (state "
                 "ring %d x 16 KB + operator ring %d stages), tiles 128x%d x%d (%d accumulator stage%s), k-blocks of 32; zhat row quantisation kernel; product2 "
                 "with TMA-streamed epilogue operands and stores, tiles 128x%d x%d (%d + %d stages; width %s); tolerance-mode solves: ",
                 sl.g1h.a_stages, sl.g1h.stages, sl.g1h.bn, sl.g1h.n_tiles, sl.g1h.acc_stages, sl.g1h.acc_stages > 1 ? "s" : "",
                 sl.g2h.bn, sl.g2h.n_tiles,
                 sl.g2h.stages, sl.g2h.e_stages, tune_note_h.empty() ? "by plan" : tune_note_h.c_str());
        h->desc = std::string(buf) + h->desc;
    }
    return GPAD_OK;
}

void destroy_batch(gpad_handle_s* h) {
    for (BatchSlot& sl : h->slot) {
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_comp) cudaEventDestroy(sl.ev_comp);
        if (sl.ev_out) cudaEventDestroy(sl.ev_out);
    }
    for (cudaEvent_t e : h->ev_check) cudaEventDestroy(e);
    if (h->h_active) cudaFreeHost(h->h_active);
    if (h->h_stat) cudaFreeHost(h->h_stat);
}

// ------------------------------------------------------------------ synchronous solve (slot 0)
int solve_batch_async(gpad_handle_s* h, const gpad_solve_args_t* a, long long* ticket);
int wait_batch(gpad_handle_s* h, long long ticket);

int solve_batch(gpad_handle_s* h, const gpad_solve_args_t* a) {
    const bool host = a->mem == GPAD_MEM_HOST;
    GPAD_TRY(wait_all_async(h));
    if (host && a->check_every <= 0 && a->max_iter > 0 && a->batch >= 16384 && h->knobs.sync_split &&
        (h->cfg.precision == GPAD_PREC_TF32X3 || h->cfg.precision == GPAD_PREC_FP16X3)) {
        // a big host-memory solve runs as two halves through the double-buffered path: the second half's inputs arrive and
        // the first half's results leave under the other half's iterations (main.cu copies in, iterates, copies out
        // serially).  Instances are independent, so the results are those of the single solve bit for bit; a half of >= 8K
        // instances still runs at >= 83 % of the full batch's per-instance rate.
        int n_par = 0;
        if (a->params) GPAD_TRY(gpad_problem_dims(a->problem, nullptr, nullptr, nullptr, &n_par, nullptr));
        const int n = h->n, m = h->cfg.m;
        const int b0 = std::min(a->batch, round_up(a->batch / 2, 128));
        gpad_solve_args_t half[2] = {*a, *a};
        half[0].batch = b0; half[1].batch = a->batch - b0;
        auto shift = [&](auto*& ptr, size_t per_instance) { if (ptr) ptr += (size_t)b0 * per_instance; };
        shift(half[1].g_P, n); shift(half[1].p_D, m); shift(half[1].f, n); shift(half[1].y0, m); shift(half[1].y_prev0, m);
        shift(half[1].y_next, m); shift(half[1].y, m); shift(half[1].w, m); shift(half[1].z, n); shift(half[1].zhat, n);
        shift(half[1].iters, 1); shift(half[1].status, 1); shift(half[1].max_viol, 1); shift(half[1].gap, 1);
        shift(half[1].params, n_par);
        long long t[2] = {-1, -1};
        GPAD_TRY(solve_batch_async(h, &half[0], &t[0]));
        if (half[1].batch > 0) GPAD_TRY(solve_batch_async(h, &half[1], &t[1]));
        GPAD_TRY(wait_batch(h, t[0]));
        if (t[1] >= 0) GPAD_TRY(wait_batch(h, t[1]));
        return GPAD_OK;
    }
    cudaStream_t s = host ? h->own_stream : static_cast<cudaStream_t>(a->stream);
    BatchSlot& sl = h->slot[0];
    sl.st.B = a->batch;
    GPAD_TRY(solve_begin(h, s));
    GPAD_TRY(upload_schedule(h, a->theta, a->beta, a->max_iter, s));    // beta[] on the device for the w output
    GPAD_TRY(inputs(h, sl, a, host, s));
    GPAD_TRY(iterate(h, sl, a, s));
    GPAD_TRY(outputs(h, sl, a, host, h->stage_in, s));
    GPAD_TRY(solve_end(h, s));
    if (host) GPAD_CUDA(cudaStreamSynchronize(s));
    return GPAD_OK;
}

// ------------------------------------------------------------------ asynchronous, double-buffered host-memory solves
int solve_batch_async(gpad_handle_s* h, const gpad_solve_args_t* a, long long* ticket) {
    if (a->mem != GPAD_MEM_HOST || a->check_every > 0) {
        // device-memory solves are asynchronous already; tolerance mode needs the host in its loop
        set_error("gpad_solve_async: host-memory, fixed-iteration solves only (device-memory solves already only enqueue)");
        return GPAD_ERR_UNSUPPORTED;
    }
    if (!h->stream_in) {
        // high priority: their short kernels (instance build, w output) must get SMs at the next boundary between the
        // persistent GEMM kernels of the compute stream instead of queueing behind all of them
        int lo = 0, hi = 0;
        GPAD_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        GPAD_CUDA(cudaStreamCreateWithPriority(&h->stream_in, cudaStreamNonBlocking, hi));
        GPAD_CUDA(cudaStreamCreateWithPriority(&h->stream_out, cudaStreamNonBlocking, hi));
        GPAD_TRY(dev_alloc(h, &h->stage_out, (size_t)h->cfg.max_batch * h->cfg.m));
    }
    const long long t = h->next_ticket++;
    BatchSlot& sl = h->slot[t & 1];
    if (!sl.allocated) GPAD_TRY(alloc_slot(h, sl, &h->slot[0]));
    cudaStream_t s_in = h->stream_in, s_comp = h->own_stream, s_out = h->stream_out;
    if (sl.ticket >= 0) {
        // the slot's previous occupant: its outputs must have left before new inputs land in the same arrays
        GPAD_CUDA(cudaStreamWaitEvent(s_in, sl.ev_out, 0));
    }
    sl.ticket = t;
    sl.st.B = a->batch;
    GPAD_TRY(solve_begin(h, s_comp));
    GPAD_TRY(upload_schedule(h, a->theta, a->beta, a->max_iter, s_comp));
    GPAD_TRY(inputs(h, sl, a, true, s_in));
    GPAD_CUDA(cudaEventRecord(sl.ev_in, s_in));
    GPAD_CUDA(cudaStreamWaitEvent(s_comp, sl.ev_in, 0));
    GPAD_TRY(iterate(h, sl, a, s_comp));
    GPAD_CUDA(cudaEventRecord(sl.ev_comp, s_comp));
    GPAD_CUDA(cudaStreamWaitEvent(s_out, sl.ev_comp, 0));
    GPAD_TRY(outputs(h, sl, a, true, h->stage_out, s_out));
    GPAD_CUDA(cudaEventRecord(sl.ev_out, s_out));
    GPAD_TRY(solve_end(h, s_comp));      // synchronous solves additionally wait for the tickets (wait_all_async)
    if (ticket) *ticket = t;
    return GPAD_OK;
}

int wait_batch(gpad_handle_s* h, long long ticket) {
    if (ticket < 0 || ticket >= h->next_ticket) { set_error("gpad_wait: unknown ticket %lld", ticket); return GPAD_ERR_INVALID_ARG; }
    BatchSlot& sl = h->slot[ticket & 1];
    if (sl.allocated && sl.ticket == ticket) {
        GPAD_CUDA(cudaEventSynchronize(sl.ev_out));
        sl.ticket = -1;
    }
    return GPAD_OK;              // an older ticket of this slot has completed: its successor waited for it on the device
}

}  // namespace gpad
