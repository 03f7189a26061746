// latency_warp.cu -- ONE WARP per QP for the tiny problems (n <= 16, m <= 64; the reference's default battery
// problem is n = 12, m = 56, gpad.m:4-5): latency mode with one warp, GPAD_MODE_BATCH_PER_INSTANCE with four warps
// (= four independent QPs) per CTA.
//
// The lean one-CTA kernel (latency_small.cu) spends two __syncthreads and two shared-memory exchanges per iteration on
// a problem whose whole state fits one warp.  Here lane l owns dual entries i = l and l + 32:
//   phase A (step 2, kernel_functions.cu:16-64): lane l holds COLUMNS l, l+32 of M_G, so its own w_i never leaves
//            the lane; the 16 (padded) row sums are reduced with the transposed butterfly (16 shuffles; the rows sit in
//            each lane's registers in the lane-dependent order r ^ ((l >> 1) & 15), which makes the butterfly free of
//            selects: lat_util.cuh warp_sum_prepermuted) and the owner lanes apply -g_P and the z average (step 3);
//   phase B (step 4 + step 1, kernel_functions.cu:142-200, 7-14): zhat is broadcast with n shuffles, lane l holds
//            ROWS l, l+32 of G_L and finishes y_{v+1}, w_{v+1} for its own entries.
// No shared memory, no block barrier, no global traffic in the loop except the (cached) theta / beta schedule.
// Termination: all three branches of SURVEY row T; the dual-gap branch runs the two phases once more on y_{v+1}.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "gpad_internal.h"
#include "lat_util.cuh"
#include "latency.h"

namespace gpad {
namespace lat {

namespace {

constexpr int kWR = 16;    // rows of M_G, padded
constexpr int kMR = 2;     // dual entries per lane
constexpr int kWarpsPerCta = 4;

// NR = rows of M_G that can be non-zero (n <= NR <= kWR): rows beyond it skip their products and their zhat broadcast
// ORD: shuffles issued in program order (lat_util.cuh warp_sum_transposed_v) -- see launch_warp for which mode uses what
template <bool CHECK, int NR, bool ORD>
__global__ void __launch_bounds__(32 * kWarpsPerCta) gpad_warp_kernel(const Params p_in, int warps_per_cta) {
    Params p = p_in;
    const int lane = threadIdx.x & 31;
    const size_t inst = (size_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    if (inst >= (size_t)(p.batch > 1 ? p.batch : 1)) return;       // whole warps leave together; nothing below is block-wide
    p.M_G += inst * p.op_stride_a; p.G_L += inst * p.op_stride_b;
    p.g_P += inst * p.n; p.p_D += inst * p.m;
    if (p.y0) p.y0 += inst * p.m;
    if (p.y_prev0) p.y_prev0 += inst * p.m;
    if (p.f) p.f += inst * p.n;
    const int n = p.n, m = p.m;

    // ---- operators and state into registers ----
    float mg[kMR][kWR], gl[kMR][kWR];      // mg[j][k] = M_G[row k ^ rmask][column], gl[j][c] = G_L[row][column c]
    const int rmask = (lane >> 1) & (kWR - 1);
    float yv[kMR], yp[kMR], pd[kMR], w[kMR], yn[kMR], sb[kMR];
    bool own[kMR];
    const float beta0 = p.beta[0];
#pragma unroll
    for (int j = 0; j < kMR; ++j) {
        const int i = lane + 32 * j;
        own[j] = i < m;
#pragma unroll
        for (int r = 0; r < kWR; ++r) {
            const int rp = r ^ rmask;          // the row this register slot holds for the reduction
            mg[j][r] = (own[j] && rp < n) ? __ldg(p.M_G + (size_t)rp * p.mld + i) : 0.f;
            gl[j][r] = (own[j] && r < n) ? __ldg(p.G_L + (size_t)i * p.nld + r) : 0.f;
        }
        yv[j] = (own[j] && p.y0) ? p.y0[i] : 0.f;
        yp[j] = (own[j] && p.y_prev0) ? p.y_prev0[i] : 0.f;
        pd[j] = own[j] ? p.p_D[i] : 0.f;
        w[j] = __fadd_rn(yv[j], __fmul_rn(beta0, __fsub_rn(yv[j], yp[j])));     // step 1 of iteration 0
        yn[j] = yv[j];
        sb[j] = 0.f;
    }
    const int r_me = rmask;                            // the row whose total this lane holds after the reduction
    const float gp_me = r_me < n ? p.g_P[r_me] : 0.f;
    const float f_me = (p.f && r_me < n && (lane & 1) == 0) ? p.f[r_me] : 0.f;      // one lane per row carries f_r
    float z_me = 0.f, zh_me = 0.f;

    int iters = 0, status = GPAD_STATUS_MAX_ITER;
    float out_viol = __int_as_float(0x7fc00000), out_gap = __int_as_float(0x7fc00000);
    int until_check = CHECK ? p.check_every : 0x7fffffff;
    float theta_pf = __ldg(p.theta), beta_pf = p.max_iter > 1 ? __ldg(p.beta + 1) : 0.f;
    for (int v = 0; v < p.max_iter; ++v) {
        const float theta = theta_pf, one_minus = 1.0f - theta;
        const bool last = (v + 1 == p.max_iter);
        const float beta_next = last ? 0.f : beta_pf;
        if (!last) {
            theta_pf = __ldg(p.theta + v + 1);
            beta_pf = (v + 2 < p.max_iter) ? __ldg(p.beta + v + 2) : 0.f;
        }
        const bool check = CHECK && (--until_check == 0);
        if (check) until_check = p.check_every;

        // ---------------- phase A ----------------
        float acc[kWR];
#pragma unroll
        for (int r = 0; r < kWR; ++r) acc[r] = fmaf(mg[1][r], w[1], mg[0][r] * w[0]);
        const float tot = warp_sum_prepermuted<kWR, ORD>(acc);
        zh_me = tot - gp_me;
        z_me = __fadd_rn(__fmul_rn(one_minus, z_me), __fmul_rn(theta, zh_me));

        // ---------------- phase B ----------------
        float d[kMR] = {0.f, 0.f};
        float zc[NR];
#pragma unroll
        for (int c = 0; c < NR; ++c) zc[c] = ORD ? shfl_idx_v(zh_me, 2 * c) : __shfl_sync(0xffffffffu, zh_me, 2 * c);
#pragma unroll
        for (int c = 0; c < NR; ++c) {
#pragma unroll
            for (int j = 0; j < kMR; ++j) d[j] = fmaf(gl[j][c], zc[c], d[j]);
        }
        float r_max_sbar = -INFINITY, r_max_rhat = -INFINITY, r_min_w = INFINITY, r_w_rhat = 0.f, r_w_dot = 0.f, r_bad = 0.f;
#pragma unroll
        for (int j = 0; j < kMR; ++j) {
            const float s = d[j] + (w[j] + pd[j]);
            yn[j] = 0.5f * (s + fabsf(s));
            if (CHECK) {
                const float rhat = d[j] + pd[j];
                sb[j] = __fadd_rn(__fmul_rn(one_minus, sb[j]), __fmul_rn(theta, rhat));
                if (check && own[j]) {
                    r_max_sbar = fmaxf(r_max_sbar, sb[j]); r_max_rhat = fmaxf(r_max_rhat, rhat); r_min_w = fminf(r_min_w, w[j]);
                    r_w_rhat = fmaf(w[j], rhat, r_w_rhat);
                    r_w_dot = fmaf(w[j], d[j], r_w_dot);
                    if (!isfinite(yn[j])) r_bad = 1.f;
                }
            }
        }
        iters = v + 1;
        if (check) {
            float fz = f_me * zh_me;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                r_max_sbar = fmaxf(r_max_sbar, __shfl_xor_sync(0xffffffffu, r_max_sbar, o));
                r_max_rhat = fmaxf(r_max_rhat, __shfl_xor_sync(0xffffffffu, r_max_rhat, o));
                r_min_w = fminf(r_min_w, __shfl_xor_sync(0xffffffffu, r_min_w, o));
                r_w_rhat += __shfl_xor_sync(0xffffffffu, r_w_rhat, o);
                r_w_dot += __shfl_xor_sync(0xffffffffu, r_w_dot, o);
                r_bad = fmaxf(r_bad, __shfl_xor_sync(0xffffffffu, r_bad, o));
                fz += __shfl_xor_sync(0xffffffffu, fz, o);
            }
            const float viol_z = p.L * r_max_sbar, viol_zhat = p.L * r_max_rhat;
            out_viol = viol_z;
            bool stop = false;
            if (r_bad > 0.f) { status = GPAD_STATUS_NONFINITE; stop = true; }
            else if (viol_z <= p.eps_g) { status = GPAD_STATUS_CONVERGED_Z; stop = true; }
            else if (viol_zhat <= p.eps_g) {
                const float V = 0.5f * (fz - p.L * r_w_dot);
                if (r_min_w >= 0.f) {
                    const float gapv = -p.L * r_w_rhat;
                    out_gap = gapv;
                    if (gapv <= p.eps_V || (p.f && gapv <= V * p.eps_V / (1.0f + p.eps_V))) {
                        status = GPAD_STATUS_CONVERGED_ZHAT; out_viol = viol_zhat; stop = true;
                    }
                } else if (p.f) {
                    // dual branch: Phi(y_{v+1}) with z_y = M_G y+ - g_P and G_L z_y -- the two phases once more, on y+
                    float a2[kWR];
#pragma unroll
                    for (int r = 0; r < kWR; ++r) a2[r] = fmaf(mg[1][r], yn[1], mg[0][r] * yn[0]);
                    const float zy_me = warp_sum_prepermuted<kWR, false>(a2) - gp_me;
                    float fzy = f_me * zy_me, y_gz = 0.f, y_pd = 0.f;
                    float d2[kMR] = {0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < NR; ++c) {
                        const float zc = __shfl_sync(0xffffffffu, zy_me, 2 * c);
#pragma unroll
                        for (int j = 0; j < kMR; ++j) d2[j] = fmaf(gl[j][c], zc, d2[j]);
                    }
#pragma unroll
                    for (int j = 0; j < kMR; ++j)
                        if (own[j]) { y_gz = fmaf(yn[j], d2[j], y_gz); y_pd = fmaf(yn[j], pd[j], y_pd); }
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        fzy += __shfl_xor_sync(0xffffffffu, fzy, o); y_gz += __shfl_xor_sync(0xffffffffu, y_gz, o); y_pd += __shfl_xor_sync(0xffffffffu, y_pd, o);
                    }
                    const float Phi = 0.5f * fzy + 0.5f * p.L * y_gz + p.L * y_pd;
                    const float gapv = V - Phi;
                    out_gap = gapv;
                    if (gapv <= p.eps_V * fmaxf(Phi, 1.0f)) { status = GPAD_STATUS_CONVERGED_DUAL; out_viol = viol_zhat; stop = true; }
                }
            }
            if (stop) break;
        }
        if (!last) {
            // advance: w_{v+1}, y_{v-1} <- y_v <- y_{v+1}; not on the last iteration: w_v / y_v are outputs
#pragma unroll
            for (int j = 0; j < kMR; ++j) {
                w[j] = __fadd_rn(yn[j], __fmul_rn(beta_next, __fsub_rn(yn[j], yv[j])));
                yp[j] = yv[j];
                yv[j] = yn[j];
            }
        }
    }

    // ---------------- outputs (main.cu:176-180 + termination outputs) ----------------
    bool bad = false;
#pragma unroll
    for (int j = 0; j < kMR; ++j) {
        const int i = lane + 32 * j;
        if (own[j]) {
            if (p.out_y_next) p.out_y_next[inst * m + i] = yn[j];
            if (p.out_y) p.out_y[inst * m + i] = yv[j];
            if (p.out_w) p.out_w[inst * m + i] = w[j];
            bad = bad || !isfinite(yn[j]);
        }
    }
    if ((lane & 1) == 0 && r_me < n) {
        if (p.out_z) p.out_z[inst * n + r_me] = z_me;
        if (p.out_zhat) p.out_zhat[inst * n + r_me] = zh_me;
    }
    if (status == GPAD_STATUS_MAX_ITER && __any_sync(0xffffffffu, bad)) status = GPAD_STATUS_NONFINITE;
    if (lane == 0) {
        if (p.out_iters) p.out_iters[inst] = iters;
        if (p.out_status) p.out_status[inst] = status;
        if (p.out_max_viol) p.out_max_viol[inst] = out_viol;
        if (p.out_gap) p.out_gap[inst] = out_gap;
    }
}

// ---- TWO QPs per warp (fixed-iteration batches): 16 lanes per QP, lane h = lane & 15 owns dual entries h, h+16, h+32, h+48.
// The reduction and the zhat broadcast then serve two QPs per shuffle: per QP-iteration 15 + 12 shuffles become 7.5 + 6
// (the products, 2 n m / 32 per QP, stay).  Same arithmetic per QP as gpad_warp_kernel except the order of the four
// products per row and of the 16-lane butterfly; results sit within rounding of it (tests compare both with the oracle).
constexpr int kMR2 = 4;

template <int NR, bool ORD>
__global__ void __launch_bounds__(32 * kWarpsPerCta) gpad_warp2_kernel(const Params p_in, int warps_per_cta) {
    Params p = p_in;
    const int lane = threadIdx.x & 31, h = lane & 15;
    const size_t B = (size_t)p.batch;
    const size_t pair = (size_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    if (2 * pair >= B) return;                                  // whole warps leave together
    const size_t inst_raw = 2 * pair + (lane >> 4);
    const bool live = inst_raw < B;                              // an odd batch leaves the last half-warp idle
    const size_t inst = live ? inst_raw : B - 1;                 // (it shadows the last QP and stores nothing)
    p.M_G += inst * p.op_stride_a; p.G_L += inst * p.op_stride_b;
    p.g_P += inst * p.n; p.p_D += inst * p.m;
    if (p.y0) p.y0 += inst * p.m;
    if (p.y_prev0) p.y_prev0 += inst * p.m;
    const int n = p.n, m = p.m;

    float mg[kMR2][kWR], gl[kMR2][NR];     // mg[j][k] = M_G[row k ^ h][column], gl[j][c] = G_L[row][column c]
    float yv[kMR2], yp[kMR2], pd[kMR2], w[kMR2], yn[kMR2];
    bool own[kMR2];
    const float beta0 = p.beta[0];
#pragma unroll
    for (int j = 0; j < kMR2; ++j) {
        const int i = h + 16 * j;
        own[j] = i < m;
#pragma unroll
        for (int r = 0; r < kWR; ++r) {
            const int rp = r ^ h;
            mg[j][r] = (own[j] && rp < n) ? __ldg(p.M_G + (size_t)rp * p.mld + i) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < NR; ++c) gl[j][c] = (own[j] && c < n) ? __ldg(p.G_L + (size_t)i * p.nld + c) : 0.f;
        yv[j] = (own[j] && p.y0) ? p.y0[i] : 0.f;
        yp[j] = (own[j] && p.y_prev0) ? p.y_prev0[i] : 0.f;
        pd[j] = own[j] ? p.p_D[i] : 0.f;
        w[j] = __fadd_rn(yv[j], __fmul_rn(beta0, __fsub_rn(yv[j], yp[j])));     // step 1 of iteration 0
        yn[j] = yv[j];
    }
    const float gp_me = h < n ? p.g_P[h] : 0.f;              // lane h ends the reduction with the total of row h
    float z_me = 0.f, zh_me = 0.f;

    float theta_pf = __ldg(p.theta), beta_pf = p.max_iter > 1 ? __ldg(p.beta + 1) : 0.f;
    for (int v = 0; v < p.max_iter; ++v) {
        const float theta = theta_pf, one_minus = 1.0f - theta;
        const bool last = (v + 1 == p.max_iter);
        const float beta_next = last ? 0.f : beta_pf;
        if (!last) {
            theta_pf = __ldg(p.theta + v + 1);
            beta_pf = (v + 2 < p.max_iter) ? __ldg(p.beta + v + 2) : 0.f;
        }
        // ---------------- phase A ----------------
        float acc[kWR];
#pragma unroll
        for (int r = 0; r < kWR; ++r) acc[r] = fmaf(mg[3][r], w[3], fmaf(mg[2][r], w[2], fmaf(mg[1][r], w[1], mg[0][r] * w[0])));
        const float tot = halfwarp_sum_prepermuted<ORD>(acc);
        zh_me = tot - gp_me;
        z_me = __fadd_rn(__fmul_rn(one_minus, z_me), __fmul_rn(theta, zh_me));
        // ---------------- phase B ----------------
        float d[kMR2] = {0.f, 0.f, 0.f, 0.f};
        float zc[NR];
#pragma unroll
        for (int c = 0; c < NR; ++c) zc[c] = __shfl_sync(0xffffffffu, zh_me, c, 16);       // lane c of this half
#pragma unroll
        for (int c = 0; c < NR; ++c) {
#pragma unroll
            for (int j = 0; j < kMR2; ++j) d[j] = fmaf(gl[j][c], zc[c], d[j]);
        }
#pragma unroll
        for (int j = 0; j < kMR2; ++j) {
            const float s = d[j] + (w[j] + pd[j]);
            yn[j] = 0.5f * (s + fabsf(s));
        }
        if (!last) {
#pragma unroll
            for (int j = 0; j < kMR2; ++j) {
                w[j] = __fadd_rn(yn[j], __fmul_rn(beta_next, __fsub_rn(yn[j], yv[j])));
                yp[j] = yv[j];
                yv[j] = yn[j];
            }
        }
    }

    // ---------------- outputs (main.cu:176-180 + termination outputs) ----------------
    bool bad = false;
#pragma unroll
    for (int j = 0; j < kMR2; ++j) {
        const int i = h + 16 * j;
        if (own[j] && live) {
            if (p.out_y_next) p.out_y_next[inst * m + i] = yn[j];
            if (p.out_y) p.out_y[inst * m + i] = yv[j];
            if (p.out_w) p.out_w[inst * m + i] = w[j];
            bad = bad || !isfinite(yn[j]);
        }
    }
    if (h < n && live) {
        if (p.out_z) p.out_z[inst * n + h] = z_me;
        if (p.out_zhat) p.out_zhat[inst * n + h] = zh_me;
    }
    const unsigned bad_lanes = __ballot_sync(0xffffffffu, bad);
    const bool bad_half = (bad_lanes >> (lane & 16)) & 0xffffu;
    if (h == 0 && live) {
        if (p.out_iters) p.out_iters[inst] = p.max_iter;
        if (p.out_status) p.out_status[inst] = bad_half ? GPAD_STATUS_NONFINITE : GPAD_STATUS_MAX_ITER;
        if (p.out_max_viol) p.out_max_viol[inst] = __int_as_float(0x7fc00000);
        if (p.out_gap) p.out_gap[inst] = __int_as_float(0x7fc00000);
    }
}

}  // namespace

int warp_supported(const Params& p) { return p.n <= kWR && p.m <= 32 * kMR; }

int launch_warp(const Params& p, cudaStream_t stream) {
    const int B = p.batch > 1 ? p.batch : 1;
    const int wpc = B > 1 ? kWarpsPerCta : 1;
    const int grid = (B + wpc - 1) / wpc;
    const bool chk = p.check_every > 0;
    // NR = 12 covers the reference's default battery problem (n = 12): 12 instead of 16 zhat broadcasts and 12-long FMA
    // chains in phase B; program-ordered shuffles keep the reduction breadth-first.  Measured with the select-free reduction,
    // one QP, 100 iterations: (12, ordered) 35.3 us, (12, ptxas) 40.4 us, (16, either) 55 us; batches are issue-bound and
    // prefer the same plan.
    int nr = p.n <= 12 ? 12 : kWR;
    bool ord = true;
    // experiments (GPAD_DEBUG warp_rows / warp_ordered / warp_pack, read at gpad_setup)
    if ((p.warp_rows == 12 || p.warp_rows == kWR) && p.warp_rows >= p.n) nr = p.warp_rows;
    if (p.warp_ordered >= 0) ord = p.warp_ordered != 0;
    // fixed-iteration batches: two QPs per warp (halves the reduction and broadcast shuffles per QP)
    if (B > 1 && !chk && p.warp_pack != 0 && p.max_iter >= 1) {
        const int grid2 = (int)(((size_t)(B + 1) / 2 + kWarpsPerCta - 1) / kWarpsPerCta);
        if (nr == 12) { if (ord) gpad_warp2_kernel<12, true><<<grid2, 32 * kWarpsPerCta, 0, stream>>>(p, kWarpsPerCta);
                        else gpad_warp2_kernel<12, false><<<grid2, 32 * kWarpsPerCta, 0, stream>>>(p, kWarpsPerCta); }
        else          { if (ord) gpad_warp2_kernel<kWR, true><<<grid2, 32 * kWarpsPerCta, 0, stream>>>(p, kWarpsPerCta);
                        else gpad_warp2_kernel<kWR, false><<<grid2, 32 * kWarpsPerCta, 0, stream>>>(p, kWarpsPerCta); }
        GPAD_CUDA(cudaGetLastError());
        return GPAD_OK;
    }
#define GPAD_WARP_LAUNCH(C, N, O) gpad_warp_kernel<C, N, O><<<grid, 32 * wpc, 0, stream>>>(p, wpc)
    if (nr == 12) {
        if (ord) { if (chk) GPAD_WARP_LAUNCH(true, 12, true); else GPAD_WARP_LAUNCH(false, 12, true); }
        else     { if (chk) GPAD_WARP_LAUNCH(true, 12, false); else GPAD_WARP_LAUNCH(false, 12, false); }
    } else {
        if (ord) { if (chk) GPAD_WARP_LAUNCH(true, kWR, true); else GPAD_WARP_LAUNCH(false, kWR, true); }
        else     { if (chk) GPAD_WARP_LAUNCH(true, kWR, false); else GPAD_WARP_LAUNCH(false, kWR, false); }
    }
#undef GPAD_WARP_LAUNCH
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

}  // namespace lat
}  // namespace gpad
