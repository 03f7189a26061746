// latency.cu -- latency mode: ONE QP, the whole GPAD loop inside one persistent kernel.
//
// Replaces the reference's 5 launches + 3 cudaDeviceSynchronize per iteration
// (main.cu:160-175) with zero launches per iteration.  G cooperating CTAs split the rows of
// both operators; each CTA keeps its operator rows, the full extrapolated dual w and the full
// zhat in shared memory.  Per iteration:
//
//   phase A (steps 2+3, kernel_functions.cu:16-72):  own rows r of M_G:
//        zhat_r = <M_G[r,:], w> - g_P[r];  z_r = (1-theta) z_r + theta zhat_r;  publish zhat_r
//   -- barrier --
//   phase B (steps 4+1, kernel_functions.cu:142-200, 7-14): own rows i of G_L:
//        s = <G_L[i,:], zhat> + (w_i + p_D_i);  y+_i = (s+|s|)/2;
//        w+_i = y+_i + beta_{v+1} (y+_i - y_i);  publish w+_i   (y_prev <- y is a register move,
//        which removes the reference's DeviceArrayCopy launch, kernel_functions.cu:260-264)
//   -- barrier --
//
// The barrier / publish mechanism is the only thing that differs between the three variants:
//   SYNC_BLOCK   G = 1            __syncthreads, vectors live in this CTA's shared memory
//   SYNC_CLUSTER G <= 16, 1 cluster  barrier.cluster + distributed-shared-memory stores
//   SYNC_GRID    G = #SMs, cooperative launch, global-memory exchange + atomic-counter barrier
// Operators are cached in shared memory when the CTA's slice fits (OPS_SMEM), otherwise they
// stream from L2 every iteration.  Dot products: LPR lanes per row, float4 loads, shuffle tree.
//
// Termination (acceldualgrad.m:66-79): reductions are warp-shuffle trees into
// per-CTA partials, exchanged with the same mechanism, so every CTA takes the same decision.
#include <cooperative_groups.h>

#include "gpad_internal.h"
#include "latency.h"

namespace cg = cooperative_groups;

namespace gpad {
namespace lat {

namespace {

constexpr int kNumRed = 8;  // max_sbar, max_rhat, min_w, w.rhat, w.dot, f.zhat, nonfinite, spare

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ int round_up_dev(int v, int q) { return (v + q - 1) / q * q; }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// <row[0..len4*4), x> with `lpr` lanes per row (lpr in {1,2,4,8,16,32}); every lane of the
// lpr-group returns the full sum.
template <bool OPS_SMEM>
__device__ __forceinline__ float row_dot(const float* __restrict__ row, const float* __restrict__ x,
                                         int len4, int sub, int lpr) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = sub;
    for (; k + 3 * lpr < len4; k += 4 * lpr) {
        float4 m0, m1, m2, m3;
        if (OPS_SMEM) { m0 = r4[k]; m1 = r4[k + lpr]; m2 = r4[k + 2 * lpr]; m3 = r4[k + 3 * lpr]; }
        else { m0 = __ldg(r4 + k); m1 = __ldg(r4 + k + lpr); m2 = __ldg(r4 + k + 2 * lpr); m3 = __ldg(r4 + k + 3 * lpr); }
        const float4 v0 = x4[k], v1 = x4[k + lpr], v2 = x4[k + 2 * lpr], v3 = x4[k + 3 * lpr];
        a0 = fmaf(m0.x, v0.x, a0); a0 = fmaf(m0.y, v0.y, a0); a0 = fmaf(m0.z, v0.z, a0); a0 = fmaf(m0.w, v0.w, a0);
        a1 = fmaf(m1.x, v1.x, a1); a1 = fmaf(m1.y, v1.y, a1); a1 = fmaf(m1.z, v1.z, a1); a1 = fmaf(m1.w, v1.w, a1);
        a2 = fmaf(m2.x, v2.x, a2); a2 = fmaf(m2.y, v2.y, a2); a2 = fmaf(m2.z, v2.z, a2); a2 = fmaf(m2.w, v2.w, a2);
        a3 = fmaf(m3.x, v3.x, a3); a3 = fmaf(m3.y, v3.y, a3); a3 = fmaf(m3.z, v3.z, a3); a3 = fmaf(m3.w, v3.w, a3);
    }
    for (; k < len4; k += lpr) {
        const float4 m0 = OPS_SMEM ? r4[k] : __ldg(r4 + k);
        const float4 v0 = x4[k];
        a0 = fmaf(m0.x, v0.x, a0); a0 = fmaf(m0.y, v0.y, a0); a0 = fmaf(m0.z, v0.z, a0); a0 = fmaf(m0.w, v0.w, a0);
    }
    float acc = (a0 + a1) + (a2 + a3);
    for (int o = lpr >> 1; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

template <int SYNC>
struct Sync {
    unsigned epoch = 0;
    // full barrier across the G cooperating CTAs; makes everything published before it visible
    __device__ __forceinline__ void barrier(const Params& p) {
        if (SYNC == SYNC_BLOCK) {
            __syncthreads();
        } else if (SYNC == SYNC_CLUSTER) {
            cg::this_cluster().sync();
        } else {
            __syncthreads();
            if (threadIdx.x == 0) {
                epoch += gridDim.x;
                __threadfence();
                atomicAdd(p.barrier, 1u);
                while (ld_acquire(p.barrier) < epoch) {}
            }
            __syncthreads();
        }
    }
};

// publish value v of element idx of an exchanged vector (local smem copy `loc`, global copy `glob`)
template <int SYNC>
__device__ __forceinline__ void publish(float* loc, float* glob, int idx, float v) {
    if (SYNC == SYNC_BLOCK) {
        loc[idx] = v;
    } else if (SYNC == SYNC_CLUSTER) {
        cg::cluster_group cl = cg::this_cluster();
        const unsigned C = cl.num_blocks();
        for (unsigned r = 0; r < C; ++r) cl.map_shared_rank(loc, r)[idx] = v;
    } else {
        glob[idx] = v;
    }
}

// after the barrier: bring the full exchanged vector into local shared memory (grid mode only)
template <int SYNC>
__device__ __forceinline__ void gather(float* loc, const float* glob, int len) {
    if (SYNC == SYNC_GRID) {
        for (int i = threadIdx.x; i < len; i += blockDim.x) loc[i] = __ldcg(glob + i);
        __syncthreads();
    }
}

template <int SYNC, bool OPS_SMEM>
__global__ void __launch_bounds__(kMaxThreads, 1) gpad_latency_kernel(const Params p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int G = (SYNC == SYNC_BLOCK) ? 1 : (int)gridDim.x;
    const int c = (SYNC == SYNC_BLOCK) ? 0 : (int)blockIdx.x;
    const int n = p.n, m = p.m, nld = p.nld, mld = p.mld;

    // row ownership
    const int a0 = min(n, c * p.rows_a), a1 = min(n, a0 + p.rows_a);
    const int b0 = min(m, c * p.rows_b), b1 = min(m, b0 + p.rows_b);
    const int na = a1 - a0, nb = b1 - b0;

    // shared memory carve-up (all sizes multiples of 4 floats)
    float* w_s = smem;                     // [mld] full w_v (zero padded)
    float* zh_s = w_s + mld;               // [nld] full zhat_v (zero padded)
    float* yv_s = zh_s + nld;              // [rows_b] own y_v
    float* yp_s = yv_s + p.rows_b_pad;     // [rows_b] own y_{v-1}
    float* yn_s = yp_s + p.rows_b_pad;     // [rows_b] own y_{v+1}
    float* sb_s = yn_s + p.rows_b_pad;     // [rows_b] own sbar
    float* dt_s = sb_s + p.rows_b_pad;     // [rows_b] own dot = (G_L zhat)_i
    float* pd_s = dt_s + p.rows_b_pad;     // [rows_b] own p_D rows (no global loads inside the loop)
    float* z_s = pd_s + p.rows_b_pad;      // [rows_a] own z
    float* gp_s = z_s + p.rows_a_pad;      // [rows_a] own g_P rows
    float* f_s = gp_s + p.rows_a_pad;      // [rows_a] own f rows (termination only)
    float* red_s = f_s + p.rows_a_pad;     // [kNumRed * G_pad] reductions of all CTAs + [32*kNumRed] scratch
    float* scr_s = red_s + kNumRed * p.g_pad;
    float* ops_s = scr_s + 32 * kNumRed;   // OPS_SMEM: [rows_a][mld] then [rows_b][nld]
    const float* MGrows = OPS_SMEM ? ops_s : p.M_G + (size_t)a0 * mld;
    const float* GLrows = OPS_SMEM ? ops_s + (size_t)p.rows_a * mld : p.G_L + (size_t)b0 * nld;

    // ---- prologue: operators to smem, initial vectors ----
    if (OPS_SMEM) {
        const float4* src = reinterpret_cast<const float4*>(p.M_G + (size_t)a0 * mld);
        float4* dst = reinterpret_cast<float4*>(ops_s);
        for (int i = tid; i < na * (mld >> 2); i += nthr) dst[i] = __ldg(src + i);
        src = reinterpret_cast<const float4*>(p.G_L + (size_t)b0 * nld);
        dst = reinterpret_cast<float4*>(ops_s + (size_t)p.rows_a * mld);
        for (int i = tid; i < nb * (nld >> 2); i += nthr) dst[i] = __ldg(src + i);
    }
    const float beta0 = p.beta[0];
    for (int i = tid; i < mld; i += nthr) {
        float wv = 0.f;
        if (i < m) {
            const float y = p.y0 ? p.y0[i] : 0.f, yp = p.y_prev0 ? p.y_prev0[i] : 0.f;
            wv = __fadd_rn(y, __fmul_rn(beta0, __fsub_rn(y, yp)));   // step 1 of iteration 0 (unfused like the CPU build)
        }
        w_s[i] = wv;
    }
    for (int i = tid; i < nld; i += nthr) zh_s[i] = 0.f;
    for (int i = tid; i < nb; i += nthr) {
        yv_s[i] = p.y0 ? p.y0[b0 + i] : 0.f;
        yp_s[i] = p.y_prev0 ? p.y_prev0[b0 + i] : 0.f;
        yn_s[i] = yv_s[i];
        sb_s[i] = 0.f;
        dt_s[i] = 0.f;
        pd_s[i] = p.p_D[b0 + i];
    }
    for (int i = tid; i < na; i += nthr) {
        z_s[i] = 0.f;
        gp_s[i] = p.g_P[a0 + i];
        f_s[i] = p.f ? p.f[a0 + i] : 0.f;
    }
    Sync<SYNC> sync;
    if (SYNC == SYNC_CLUSTER) cg::this_cluster().sync();   // peers' smem must exist before remote stores
    else __syncthreads();

    const int lpr_a = p.lpr_a, lpr_b = p.lpr_b;
    const int mld4 = mld >> 2, nld4 = nld >> 2;
    const bool checking = p.check_every > 0;
    int iters = 0, status = GPAD_STATUS_MAX_ITER;
    float out_viol = __int_as_float(0x7fc00000), out_gap = __int_as_float(0x7fc00000);

    // theta_v / beta_{v+1} are fetched one iteration ahead so their latency never sits on the
    // critical path (cluster / grid barriers invalidate L1)
    float theta_pf = p.theta[0];
    float beta_pf = p.max_iter > 1 ? p.beta[1] : 0.f;
    for (int v = 0; v < p.max_iter; ++v) {
        const float theta = theta_pf;
        const float one_minus = 1.0f - theta;
        const bool last = (v + 1 == p.max_iter);
        const float beta_next = last ? 0.f : beta_pf;
        if (!last) {
            theta_pf = __ldg(p.theta + v + 1);
            beta_pf = (v + 2 < p.max_iter) ? __ldg(p.beta + v + 2) : 0.f;
        }
        const bool check = checking && ((v + 1) % p.check_every == 0);

        // ---------------- phase A: zhat rows, z average ----------------
        float f_zhat = 0.f;
        {
            const int rows_per_pass = (nthr / lpr_a);
            const int sub = tid % lpr_a, slot = tid / lpr_a;
            for (int r = slot; r < round_up_dev(na, rows_per_pass); r += rows_per_pass) {
                // all lanes of a warp iterate together (shuffles); rows beyond na are masked
                const bool valid = r < na;
                const float* row = MGrows + (size_t)(valid ? r : 0) * mld;
                const float d = row_dot<OPS_SMEM>(row, w_s, mld4, sub, lpr_a);
                if (valid && sub == 0) {
                    const float zh = d - gp_s[r];
                    z_s[r] = __fadd_rn(__fmul_rn(one_minus, z_s[r]), __fmul_rn(theta, zh));
                    publish<SYNC>(zh_s, p.x_zhat, a0 + r, zh);
                    if (check && p.f) f_zhat = fmaf(f_s[r], zh, f_zhat);
                }
            }
        }
        if (check && p.f) {     // block partial of f'zhat -> red slot of this CTA
            f_zhat = warp_sum(f_zhat);
            if (lane == 0) scr_s[warp] = f_zhat;
            __syncthreads();
            if (tid == 0) {
                float s = 0.f;
                for (int k = 0; k < nwarps; ++k) s += scr_s[k];
                publish<SYNC>(red_s, p.x_red + ((v / p.check_every) & 1) * kNumRed * p.g_pad, 5 * p.g_pad + c, s);
            }
        }
        sync.barrier(p);
        gather<SYNC>(zh_s, p.x_zhat, n);

        // ---------------- phase B: dual step, projection, momentum ----------------
        float r_max_sbar = -INFINITY, r_max_rhat = -INFINITY, r_min_w = INFINITY;
        float r_w_rhat = 0.f, r_w_dot = 0.f, r_bad = 0.f;
        {
            const int rows_per_pass = (nthr / lpr_b);
            const int sub = tid % lpr_b, slot = tid / lpr_b;
            for (int r = slot; r < round_up_dev(nb, rows_per_pass); r += rows_per_pass) {
                const bool valid = r < nb;
                const float* row = GLrows + (size_t)(valid ? r : 0) * nld;
                const float d = row_dot<OPS_SMEM>(row, zh_s, nld4, sub, lpr_b);
                if (valid && sub == 0) {
                    const int i = b0 + r;
                    const float wi = w_s[i], pd = pd_s[r];
                    float s = d + (wi + pd);
                    const float yn = 0.5f * (s + fabsf(s));
                    yn_s[r] = yn;
                    dt_s[r] = d;
                    if (!check && !last) {
                        // advance in place: w_{v+1} (only this lane reads w_s[i] during phase B),
                        // y_{v-1} <- y_v <- y_{v+1}.  Not on the last iteration: w_v / y_v are outputs.
                        const float yv = yv_s[r];
                        publish<SYNC>(w_s, p.x_w, i, __fadd_rn(yn, __fmul_rn(beta_next, __fsub_rn(yn, yv))));
                        yp_s[r] = yv;
                        yv_s[r] = yn;
                    }
                    if (checking) {
                        const float rhat = d + pd;
                        const float sb = __fadd_rn(__fmul_rn(one_minus, sb_s[r]), __fmul_rn(theta, rhat));
                        sb_s[r] = sb;
                        if (check) {
                            r_max_sbar = fmaxf(r_max_sbar, sb);
                            r_max_rhat = fmaxf(r_max_rhat, rhat);
                            r_min_w = fminf(r_min_w, wi);
                            r_w_rhat = fmaf(wi, rhat, r_w_rhat);
                            r_w_dot = fmaf(wi, d, r_w_dot);
                            if (!isfinite(yn)) r_bad = 1.f;
                        }
                    }
                }
            }
        }
        iters = v + 1;

        if (!check) {
            if (!last) {
                sync.barrier(p);
                gather<SYNC>(w_s, p.x_w, m);
            }
            continue;
        }

        // ---------------- termination test (all CTAs take the same decision) ----------------
        float* xred = p.x_red + ((v / p.check_every) & 1) * kNumRed * p.g_pad;
        {
            r_max_sbar = warp_max(r_max_sbar); r_max_rhat = warp_max(r_max_rhat); r_min_w = warp_min(r_min_w);
            r_w_rhat = warp_sum(r_w_rhat); r_w_dot = warp_sum(r_w_dot); r_bad = warp_max(r_bad);
            if (lane == 0) {
                scr_s[0 * 32 + warp] = r_max_sbar; scr_s[1 * 32 + warp] = r_max_rhat; scr_s[2 * 32 + warp] = r_min_w;
                scr_s[3 * 32 + warp] = r_w_rhat;   scr_s[4 * 32 + warp] = r_w_dot;    scr_s[6 * 32 + warp] = r_bad;
            }
            __syncthreads();
            if (tid == 0) {
                float a = -INFINITY, b = -INFINITY, cmin = INFINITY, d = 0.f, e = 0.f, g = 0.f;
                for (int k = 0; k < nwarps; ++k) {
                    a = fmaxf(a, scr_s[k]); b = fmaxf(b, scr_s[32 + k]); cmin = fminf(cmin, scr_s[64 + k]);
                    d += scr_s[96 + k]; e += scr_s[128 + k]; g = fmaxf(g, scr_s[192 + k]);
                }
                publish<SYNC>(red_s, xred, 0 * p.g_pad + c, a);
                publish<SYNC>(red_s, xred, 1 * p.g_pad + c, b);
                publish<SYNC>(red_s, xred, 2 * p.g_pad + c, cmin);
                publish<SYNC>(red_s, xred, 3 * p.g_pad + c, d);
                publish<SYNC>(red_s, xred, 4 * p.g_pad + c, e);
                publish<SYNC>(red_s, xred, 6 * p.g_pad + c, g);
            }
        }
        sync.barrier(p);
        gather<SYNC>(red_s, xred, kNumRed * p.g_pad);
        float max_sbar = -INFINITY, max_rhat = -INFINITY, min_w = INFINITY, w_rhat = 0.f, w_dot = 0.f, fz = 0.f, bad = 0.f;
        for (int k = 0; k < G; ++k) {       // identical order in every thread of every CTA
            max_sbar = fmaxf(max_sbar, red_s[k]);
            max_rhat = fmaxf(max_rhat, red_s[p.g_pad + k]);
            min_w = fminf(min_w, red_s[2 * p.g_pad + k]);
            w_rhat += red_s[3 * p.g_pad + k];
            w_dot += red_s[4 * p.g_pad + k];
            fz += red_s[5 * p.g_pad + k];
            bad = fmaxf(bad, red_s[6 * p.g_pad + k]);
        }
        const float viol_z = p.L * max_sbar, viol_zhat = p.L * max_rhat;
        out_viol = viol_z;
        bool stop = false;
        if (bad > 0.f) { status = GPAD_STATUS_NONFINITE; stop = true; }
        else if (viol_z <= p.eps_g) { status = GPAD_STATUS_CONVERGED_Z; stop = true; }
        else if (viol_zhat <= p.eps_g) {
            const float V = 0.5f * (fz - p.L * w_dot);
            if (min_w >= 0.f) {
                const float gapv = -p.L * w_rhat;
                out_gap = gapv;
                if (gapv <= p.eps_V || (p.f && gapv <= V * p.eps_V / (1.0f + p.eps_V))) {
                    status = GPAD_STATUS_CONVERGED_ZHAT; out_viol = viol_zhat; stop = true;
                }
            } else if (p.f) {
                // dual branch: Phi(y_{v+1}) needs z_y = M_G y+ - g_P and G_L z_y: two more phases.
                // Uses w_s / zh_s as scratch, so the current w_v / zhat_v are parked in global.
                __syncthreads();
                for (int r = tid; r < nb; r += nthr) { p.out_w[b0 + r] = w_s[b0 + r]; }
                for (int r = tid; r < na; r += nthr) { p.out_zhat[a0 + r] = zh_s[a0 + r]; }
                sync.barrier(p);      // everyone finished reading w_s / zh_s of iteration v
                for (int r = tid; r < nb; r += nthr) publish<SYNC>(w_s, p.x_w, b0 + r, yn_s[r]);
                sync.barrier(p);
                gather<SYNC>(w_s, p.x_w, m);
                float fzy = 0.f;
                {
                    const int rows_per_pass = (nthr / lpr_a);
                    const int sub = tid % lpr_a, slot = tid / lpr_a;
                    for (int r = slot; r < round_up_dev(na, rows_per_pass); r += rows_per_pass) {
                        const bool valid = r < na;
                        const float* row = MGrows + (size_t)(valid ? r : 0) * mld;
                        const float d = row_dot<OPS_SMEM>(row, w_s, mld4, sub, lpr_a);
                        if (valid && sub == 0) {
                            const float zy = d - gp_s[r];
                            publish<SYNC>(zh_s, p.x_zhat, a0 + r, zy);
                            fzy = fmaf(f_s[r], zy, fzy);
                        }
                    }
                }
                sync.barrier(p);
                gather<SYNC>(zh_s, p.x_zhat, n);
                float y_gz = 0.f, y_pd = 0.f;
                {
                    const int rows_per_pass = (nthr / lpr_b);
                    const int sub = tid % lpr_b, slot = tid / lpr_b;
                    for (int r = slot; r < round_up_dev(nb, rows_per_pass); r += rows_per_pass) {
                        const bool valid = r < nb;
                        const float* row = GLrows + (size_t)(valid ? r : 0) * nld;
                        const float d = row_dot<OPS_SMEM>(row, zh_s, nld4, sub, lpr_b);
                        if (valid && sub == 0) {
                            y_gz = fmaf(yn_s[r], d, y_gz);
                            y_pd = fmaf(yn_s[r], pd_s[r], y_pd);
                        }
                    }
                }
                fzy = warp_sum(fzy); y_gz = warp_sum(y_gz); y_pd = warp_sum(y_pd);
                __syncthreads();
                if (lane == 0) { scr_s[warp] = fzy; scr_s[32 + warp] = y_gz; scr_s[64 + warp] = y_pd; }
                __syncthreads();
                float* xred2 = p.x_red + 2 * kNumRed * p.g_pad;
                if (tid == 0) {
                    float a = 0.f, b = 0.f, d = 0.f;
                    for (int k = 0; k < nwarps; ++k) { a += scr_s[k]; b += scr_s[32 + k]; d += scr_s[64 + k]; }
                    publish<SYNC>(red_s, xred2, 0 * p.g_pad + c, a);
                    publish<SYNC>(red_s, xred2, 1 * p.g_pad + c, b);
                    publish<SYNC>(red_s, xred2, 2 * p.g_pad + c, d);
                }
                sync.barrier(p);
                gather<SYNC>(red_s, xred2, 3 * p.g_pad);
                float s_fzy = 0.f, s_ygz = 0.f, s_ypd = 0.f;
                for (int k = 0; k < G; ++k) { s_fzy += red_s[k]; s_ygz += red_s[p.g_pad + k]; s_ypd += red_s[2 * p.g_pad + k]; }
                const float Phi = 0.5f * s_fzy + 0.5f * p.L * s_ygz + p.L * s_ypd;
                const float gapv = V - Phi;
                out_gap = gapv;
                if (gapv <= p.eps_V * fmaxf(Phi, 1.0f)) { status = GPAD_STATUS_CONVERGED_DUAL; out_viol = viol_zhat; stop = true; }
                // restore w_v / zhat_v into shared memory
                sync.barrier(p);
                for (int i = tid; i < m; i += nthr) w_s[i] = __ldcg(p.out_w + i);
                for (int i = tid; i < n; i += nthr) zh_s[i] = __ldcg(p.out_zhat + i);
                __syncthreads();
            }
        }
        if (stop) break;
        if (!last) {
            __syncthreads();
            for (int r = tid; r < nb; r += nthr) {
                const float yn = yn_s[r], yv = yv_s[r];
                publish<SYNC>(w_s, p.x_w, b0 + r, __fadd_rn(yn, __fmul_rn(beta_next, __fsub_rn(yn, yv))));
                yp_s[r] = yv;
                yv_s[r] = yn;
            }
            sync.barrier(p);
            gather<SYNC>(w_s, p.x_w, m);
        }
    }

    // ---------------- epilogue: the five vectors of main.cu:176-180 + termination outputs ----------------
    __syncthreads();
    float bad_local = 0.f;
    for (int r = tid; r < nb; r += nthr) {
        const float yn = yn_s[r];
        if (!isfinite(yn)) bad_local = 1.f;
        if (p.out_y_next) p.out_y_next[b0 + r] = yn;
        if (p.out_y) p.out_y[b0 + r] = yv_s[r];
        if (p.out_w) p.out_w[b0 + r] = w_s[b0 + r];
    }
    for (int r = tid; r < na; r += nthr) {
        if (p.out_z) p.out_z[a0 + r] = z_s[r];
        if (p.out_zhat) p.out_zhat[a0 + r] = zh_s[a0 + r];
    }
    if (status == GPAD_STATUS_MAX_ITER) {
        // a non-finite iterate anywhere turns MAX_ITER into NONFINITE (one flag in global memory)
        bad_local = warp_max(bad_local);
        if (lane == 0 && bad_local > 0.f) atomicExch(p.nonfinite_flag, 1);
        sync.barrier(p);
        if (c == 0 && tid == 0 && *(volatile int*)p.nonfinite_flag) status = GPAD_STATUS_NONFINITE;
    }
    if (c == 0 && tid == 0) {
        if (p.out_iters) *p.out_iters = iters;
        if (p.out_status) *p.out_status = status;
        if (p.out_max_viol) *p.out_max_viol = out_viol;
        if (p.out_gap) *p.out_gap = out_gap;
    }
    if (SYNC == SYNC_CLUSTER) cg::this_cluster().sync();   // no CTA exits while peers may still store to it
}

}  // namespace

size_t smem_bytes(const Params& p, bool ops_smem) {
    size_t fl = (size_t)p.mld + p.nld + 6 * (size_t)p.rows_b_pad + 3 * (size_t)p.rows_a_pad + (size_t)kNumRed * p.g_pad + 32 * kNumRed;
    if (ops_smem) fl += (size_t)p.rows_a * p.mld + (size_t)p.rows_b * p.nld;
    return fl * sizeof(float);
}

template <int SYNC, bool OPS>
static int launch_variant(const Params& p, int G, int threads, size_t smem, cudaStream_t stream) {
    auto kern = gpad_latency_kernel<SYNC, OPS>;
    GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (SYNC == SYNC_GRID) {
        void* args[] = {(void*)&p};
        GPAD_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(G), dim3(threads), args, smem, stream));
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(G);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        int nattr = 0;
        if (SYNC == SYNC_CLUSTER) {
            if (G > 8) GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = G;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            nattr = 1;
        }
        cfg.attrs = attr;
        cfg.numAttrs = nattr;
        GPAD_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    }
    return GPAD_OK;
}

int launch(const Params& p, int sync_mode, bool ops_smem, int G, int threads, cudaStream_t stream) {
    const size_t smem = smem_bytes(p, ops_smem);
    switch (sync_mode) {
        case SYNC_BLOCK:
            return ops_smem ? launch_variant<SYNC_BLOCK, true>(p, 1, threads, smem, stream)
                            : launch_variant<SYNC_BLOCK, false>(p, 1, threads, smem, stream);
        case SYNC_CLUSTER:
            return ops_smem ? launch_variant<SYNC_CLUSTER, true>(p, G, threads, smem, stream)
                            : launch_variant<SYNC_CLUSTER, false>(p, G, threads, smem, stream);
        default:
            return ops_smem ? launch_variant<SYNC_GRID, true>(p, G, threads, smem, stream)
                            : launch_variant<SYNC_GRID, false>(p, G, threads, smem, stream);
    }
}

int max_cluster_size(bool ops_smem, int threads, size_t smem) {
    // largest cluster (<=16) the driver will co-schedule with this much shared memory
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    auto kern = ops_smem ? (void*)gpad_latency_kernel<SYNC_CLUSTER, true> : (void*)gpad_latency_kernel<SYNC_CLUSTER, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cfg.gridDim = dim3(16);
    int cs = 0;
    if (cudaOccupancyMaxPotentialClusterSize(&cs, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 8; }
    return cs;
}

}  // namespace lat
}  // namespace gpad
