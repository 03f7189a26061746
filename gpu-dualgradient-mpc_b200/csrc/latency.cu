// latency.cu -- latency mode: ONE QP, the whole GPAD loop inside one persistent kernel.
//
// Replaces the reference's 5 launches + 3 cudaDeviceSynchronize per iteration
// (main.cu:160-175) with zero launches per iteration.  G cooperating CTAs split the rows of
// both operators; each CTA keeps the full extrapolated dual w and the full zhat in shared
// memory.  Per iteration:
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
// Three synchronisation variants of one template:
//   SYNC_BLOCK   G = 1              __syncthreads, vectors live in this CTA's shared memory
//   SYNC_CLUSTER G <= 16, 1 cluster barrier.cluster + distributed-shared-memory stores
//   SYNC_GRID    G = #SMs, cooperative launch, global exchange + red.release / ld.acquire barrier
// Two operator-residency variants:
//   REGS = true   every thread keeps its operator fragments in registers for the whole solve
//                 (small problems / cluster slices: one pass per phase, <= kRegChunks float4 per lane)
//   REGS = false  the first res_a / res_b own rows live in shared memory, the remainder streams
//                 from L2 every iteration (partial residency: (10,100) keeps ~85 % on chip)
// All work assignment (row, lane-in-row, pointers) is computed once before the iteration loop;
// dot products are `lpr` lanes per row, float4 loads, shuffle tree.
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
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    return acc;
}
// sum over the lpr lanes of a row group (lpr = 1 << lg)
__device__ __forceinline__ float group_sum(float v, int lg) {
    switch (lg) {
        case 5: v += __shfl_xor_sync(0xffffffffu, v, 16);
        case 4: v += __shfl_xor_sync(0xffffffffu, v, 8);
        case 3: v += __shfl_xor_sync(0xffffffffu, v, 4);
        case 2: v += __shfl_xor_sync(0xffffffffu, v, 2);
        case 1: v += __shfl_xor_sync(0xffffffffu, v, 1);
        default: break;
    }
    return v;
}

// <row, x> over `chunks` float4 per lane (row / x padded so that no bounds check is needed);
// generic loads: the row may live in shared memory (resident) or in global memory (streamed)
__device__ __forceinline__ float row_dot_mem(const float4* __restrict__ row, const float4* __restrict__ x, int chunks, int lpr) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int c = 0;
    for (; c + 8 <= chunks; c += 8) {           // 8 independent 16-byte loads in flight per lane
        const float4 m0 = row[(c + 0) * lpr], m1 = row[(c + 1) * lpr], m2 = row[(c + 2) * lpr], m3 = row[(c + 3) * lpr];
        const float4 m4 = row[(c + 4) * lpr], m5 = row[(c + 5) * lpr], m6 = row[(c + 6) * lpr], m7 = row[(c + 7) * lpr];
        a0 = dot4(m0, x[(c + 0) * lpr], a0); a1 = dot4(m1, x[(c + 1) * lpr], a1);
        a2 = dot4(m2, x[(c + 2) * lpr], a2); a3 = dot4(m3, x[(c + 3) * lpr], a3);
        a0 = dot4(m4, x[(c + 4) * lpr], a0); a1 = dot4(m5, x[(c + 5) * lpr], a1);
        a2 = dot4(m6, x[(c + 6) * lpr], a2); a3 = dot4(m7, x[(c + 7) * lpr], a3);
    }
    for (; c + 2 <= chunks; c += 2) {
        const float4 m0 = row[c * lpr], m1 = row[(c + 1) * lpr];
        a0 = dot4(m0, x[c * lpr], a0); a1 = dot4(m1, x[(c + 1) * lpr], a1);
    }
    if (c < chunks) a2 = dot4(row[c * lpr], x[c * lpr], a2);
    return (a0 + a1) + (a2 + a3);
}

template <int SYNC>
struct Sync {
    unsigned epoch = 0;
    // full barrier across the G cooperating CTAs; makes everything published before it visible
    __device__ __forceinline__ void barrier(const Params& p) {
        if (SYNC == SYNC_BLOCK) {
            __syncthreads();
        } else if (SYNC == SYNC_CLUSTER) {
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else {
            __syncthreads();                      // orders this CTA's published stores before thread 0's release
            if (threadIdx.x == 0) {
                epoch += gridDim.x;
                red_release_add(p.barrier, 1u);
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
        const float4* g4 = reinterpret_cast<const float4*>(glob);
        float4* l4 = reinterpret_cast<float4*>(loc);
        for (int i = threadIdx.x; i < (len >> 2); i += blockDim.x) l4[i] = __ldcg(g4 + i);
        __syncthreads();
    }
}

template <int SYNC, bool REGS>
__global__ void __launch_bounds__(kMaxThreads, 1) gpad_latency_kernel(const Params p_in) {
    extern __shared__ __align__(16) float smem[];
    Params p = p_in;
    if (SYNC == SYNC_BLOCK && p.batch > 1) {
        // batched-GEMV mode (per-instance operators): every CTA is an independent QP
        const size_t inst = blockIdx.x;
        p.M_G += inst * p.op_stride_a; p.G_L += inst * p.op_stride_b;
        p.g_P += inst * p.n; p.p_D += inst * p.m;
        if (p.f) p.f += inst * p.n;
        if (p.y0) p.y0 += inst * p.m;
        if (p.y_prev0) p.y_prev0 += inst * p.m;
        if (p.out_y_next) p.out_y_next += inst * p.m;
        if (p.out_y) p.out_y += inst * p.m;
        if (p.out_w) p.out_w += inst * p.m;
        if (p.out_z) p.out_z += inst * p.n;
        if (p.out_zhat) p.out_zhat += inst * p.n;
        if (p.out_iters) p.out_iters += inst;
        if (p.out_status) p.out_status += inst;
        if (p.out_max_viol) p.out_max_viol += inst;
        if (p.out_gap) p.out_gap += inst;
    }
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
    float* ops_s = scr_s + 32 * kNumRed;   // !REGS: [res_a][mld] then [res_b][nld] resident operator rows

    // ---- static work assignment (computed once) ----
    const int lpr_a = 1 << p.lg_a, lpr_b = 1 << p.lg_b;
    const int sub_a = tid & (lpr_a - 1), slot_a = tid >> p.lg_a, rpp_a = nthr >> p.lg_a;
    const int sub_b = tid & (lpr_b - 1), slot_b = tid >> p.lg_b, rpp_b = nthr >> p.lg_b;
    const int chunks_a = (mld >> 2) >> p.lg_a, chunks_b = (nld >> 2) >> p.lg_b;
    const float4* w4 = reinterpret_cast<const float4*>(w_s) + sub_a;
    const float4* zh4 = reinterpret_cast<const float4*>(zh_s) + sub_b;

    // ---- prologue: operators to registers / shared memory, initial vectors ----
    float4 ra[REGS ? kRegChunks : 1], rb[REGS ? kRegChunks : 1];
    if (REGS) {
        const bool va = slot_a < na, vb = slot_b < nb;
        const float4* rowa = reinterpret_cast<const float4*>(p.M_G + (size_t)(a0 + (va ? slot_a : 0)) * mld) + sub_a;
        const float4* rowb = reinterpret_cast<const float4*>(p.G_L + (size_t)(b0 + (vb ? slot_b : 0)) * nld) + sub_b;
#pragma unroll
        for (int k = 0; k < kRegChunks; ++k) {
            ra[k] = (va && k < chunks_a) ? __ldg(rowa + k * lpr_a) : make_float4(0.f, 0.f, 0.f, 0.f);
            rb[k] = (vb && k < chunks_b) ? __ldg(rowb + k * lpr_b) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        const int ra_rows = min(na, p.res_a), rb_rows = min(nb, p.res_b);
        const float4* src = reinterpret_cast<const float4*>(p.M_G + (size_t)a0 * mld);
        float4* dst = reinterpret_cast<float4*>(ops_s);
        for (int i = tid; i < ra_rows * (mld >> 2); i += nthr) dst[i] = __ldg(src + i);
        src = reinterpret_cast<const float4*>(p.G_L + (size_t)b0 * nld);
        dst = reinterpret_cast<float4*>(ops_s + (size_t)p.res_a * mld);
        for (int i = tid; i < rb_rows * (nld >> 2); i += nthr) dst[i] = __ldg(src + i);
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
    for (int i = tid; i < p.rows_b_pad; i += nthr) {
        const bool v = i < nb;
        yv_s[i] = (v && p.y0) ? p.y0[b0 + i] : 0.f;
        yp_s[i] = (v && p.y_prev0) ? p.y_prev0[b0 + i] : 0.f;
        yn_s[i] = yv_s[i];
        sb_s[i] = 0.f;
        dt_s[i] = 0.f;
        pd_s[i] = v ? p.p_D[b0 + i] : 0.f;
    }
    for (int i = tid; i < p.rows_a_pad; i += nthr) {
        const bool v = i < na;
        z_s[i] = 0.f;
        gp_s[i] = v ? p.g_P[a0 + i] : 0.f;
        f_s[i] = (v && p.f) ? p.f[a0 + i] : 0.f;
    }
    Sync<SYNC> sync;
    if (SYNC == SYNC_CLUSTER) sync.barrier(p);   // peers' smem must exist before remote stores
    else __syncthreads();

    // per-pass row pointers of the memory path (resident rows in smem, the rest in global)
    auto row_a_ptr = [&](int r) -> const float4* {
        const float* base = r < p.res_a ? ops_s + (size_t)r * mld : p.M_G + (size_t)(a0 + r) * mld;
        return reinterpret_cast<const float4*>(base) + sub_a;
    };
    auto row_b_ptr = [&](int r) -> const float4* {
        const float* base = r < p.res_b ? ops_s + (size_t)p.res_a * mld + (size_t)r * nld : p.G_L + (size_t)(b0 + r) * nld;
        return reinterpret_cast<const float4*>(base) + sub_b;
    };
    // dot of own row r of M_G (phase A) / G_L (phase B) with the shared vector; full sum in every lane of the group
    auto dot_a = [&](int r, bool valid) -> float {
        float acc;
        if (REGS) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int k = 0; k < kRegChunks; k += 2) {
                if (k < chunks_a) s0 = dot4(ra[k], w4[k * lpr_a], s0);
                if (k + 1 < chunks_a) s1 = dot4(ra[k + 1], w4[(k + 1) * lpr_a], s1);
            }
            acc = s0 + s1;
        } else {
            acc = row_dot_mem(row_a_ptr(valid ? r : 0), w4, chunks_a, lpr_a);
        }
        return group_sum(acc, p.lg_a);
    };
    auto dot_b = [&](int r, bool valid) -> float {
        float acc;
        if (REGS) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int k = 0; k < kRegChunks; k += 2) {
                if (k < chunks_b) s0 = dot4(rb[k], zh4[k * lpr_b], s0);
                if (k + 1 < chunks_b) s1 = dot4(rb[k + 1], zh4[(k + 1) * lpr_b], s1);
            }
            acc = s0 + s1;
        } else {
            acc = row_dot_mem(row_b_ptr(valid ? r : 0), zh4, chunks_b, lpr_b);
        }
        return group_sum(acc, p.lg_b);
    };
    const int passes_a = REGS ? 1 : (p.rows_a + rpp_a - 1) / rpp_a;   // identical in every thread of every CTA
    const int passes_b = REGS ? 1 : (p.rows_b + rpp_b - 1) / rpp_b;

    const bool checking = p.check_every > 0;
    int iters = 0, status = GPAD_STATUS_MAX_ITER;
    float out_viol = __int_as_float(0x7fc00000), out_gap = __int_as_float(0x7fc00000);
    int until_check = checking ? p.check_every : 0x7fffffff;
    int check_count = 0;

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
        const bool check = (--until_check == 0);
        if (check) { until_check = p.check_every; ++check_count; }
        float* xred = p.x_red + (check_count & 1) * kNumRed * p.g_pad;   // reduction slots alternate between checks

        // ---------------- phase A: zhat rows, z average ----------------
        float f_zhat = 0.f;
        for (int ps = 0, r = slot_a; ps < passes_a; ++ps, r += rpp_a) {
            const bool valid = r < na;
            const float d = dot_a(r, valid);
            if (valid && sub_a == 0) {
                const float zh = d - gp_s[r];
                z_s[r] = __fadd_rn(__fmul_rn(one_minus, z_s[r]), __fmul_rn(theta, zh));
                publish<SYNC>(zh_s, p.x_zhat, a0 + r, zh);
                if (check) f_zhat = fmaf(f_s[r], zh, f_zhat);
            }
        }
        if (check && p.f) {     // block partial of f'zhat -> red slot of this CTA
            f_zhat = warp_sum(f_zhat);
            if (lane == 0) scr_s[warp] = f_zhat;
            __syncthreads();
            if (tid == 0) {
                float s = 0.f;
                for (int k = 0; k < nwarps; ++k) s += scr_s[k];
                publish<SYNC>(red_s, xred, 5 * p.g_pad + c, s);
            }
        }
        sync.barrier(p);
        gather<SYNC>(zh_s, p.x_zhat, nld);

        // ---------------- phase B: dual step, projection, momentum ----------------
        float r_max_sbar = -INFINITY, r_max_rhat = -INFINITY, r_min_w = INFINITY;
        float r_w_rhat = 0.f, r_w_dot = 0.f, r_bad = 0.f;
        for (int ps = 0, r = slot_b; ps < passes_b; ++ps, r += rpp_b) {
            const bool valid = r < nb;
            const float d = dot_b(r, valid);
            if (valid && sub_b == 0) {
                const int i = b0 + r;
                const float wi = w_s[i], pd = pd_s[r];
                const float s = d + (wi + pd);
                const float yn = 0.5f * (s + fabsf(s));
                yn_s[r] = yn;
                if (checking) {
                    dt_s[r] = d;
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
                if (!check && !last) {
                    // advance in place: w_{v+1} (only this lane reads w_s[i] during phase B),
                    // y_{v-1} <- y_v <- y_{v+1}.  Not on the last iteration: w_v / y_v are outputs.
                    const float yv = yv_s[r];
                    publish<SYNC>(w_s, p.x_w, i, __fadd_rn(yn, __fmul_rn(beta_next, __fsub_rn(yn, yv))));
                    yp_s[r] = yv;
                    yv_s[r] = yn;
                }
            }
        }
        iters = v + 1;

        if (!check) {
            if (!last) {
                sync.barrier(p);
                gather<SYNC>(w_s, p.x_w, mld);
            }
            continue;
        }

        // ---------------- termination test (all CTAs take the same decision) ----------------
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
                gather<SYNC>(w_s, p.x_w, mld);
                float fzy = 0.f;
                for (int ps = 0, r = slot_a; ps < passes_a; ++ps, r += rpp_a) {
                    const bool valid = r < na;
                    const float d = dot_a(r, valid);
                    if (valid && sub_a == 0) {
                        const float zy = d - gp_s[r];
                        publish<SYNC>(zh_s, p.x_zhat, a0 + r, zy);
                        fzy = fmaf(f_s[r], zy, fzy);
                    }
                }
                sync.barrier(p);
                gather<SYNC>(zh_s, p.x_zhat, nld);
                float y_gz = 0.f, y_pd = 0.f;
                for (int ps = 0, r = slot_b; ps < passes_b; ++ps, r += rpp_b) {
                    const bool valid = r < nb;
                    const float d = dot_b(r, valid);
                    if (valid && sub_b == 0) {
                        y_gz = fmaf(yn_s[r], d, y_gz);
                        y_pd = fmaf(yn_s[r], pd_s[r], y_pd);
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
            gather<SYNC>(w_s, p.x_w, mld);
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
        // a non-finite iterate anywhere turns MAX_ITER into NONFINITE
        if (SYNC == SYNC_BLOCK) {
            if (__syncthreads_or(bad_local > 0.f)) status = GPAD_STATUS_NONFINITE;
        } else {
            bad_local = warp_max(bad_local);
            if (lane == 0 && bad_local > 0.f) atomicExch(p.nonfinite_flag, 1);
            sync.barrier(p);
            if (c == 0 && tid == 0 && *(volatile int*)p.nonfinite_flag) status = GPAD_STATUS_NONFINITE;
        }
    }
    if (c == 0 && tid == 0) {
        if (p.out_iters) *p.out_iters = iters;
        if (p.out_status) *p.out_status = status;
        if (p.out_max_viol) *p.out_max_viol = out_viol;
        if (p.out_gap) *p.out_gap = out_gap;
    }
    if (SYNC == SYNC_CLUSTER) sync.barrier(p);   // no CTA exits while peers may still store to it
}

}  // namespace

size_t smem_bytes(const Params& p, bool regs) {
    size_t fl = (size_t)p.mld + p.nld + 6 * (size_t)p.rows_b_pad + 3 * (size_t)p.rows_a_pad + (size_t)kNumRed * p.g_pad + 32 * kNumRed;
    if (!regs) fl += (size_t)p.res_a * p.mld + (size_t)p.res_b * p.nld;
    return fl * sizeof(float);
}

template <int SYNC, bool REGS>
static int launch_variant(const Params& p, int G, int threads, size_t smem, cudaStream_t stream) {
    auto kern = gpad_latency_kernel<SYNC, REGS>;
    GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (SYNC == SYNC_BLOCK) {
        kern<<<p.batch > 1 ? p.batch : 1, threads, smem, stream>>>(p);
        GPAD_CUDA(cudaGetLastError());
    } else if (SYNC == SYNC_GRID) {
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

int launch(const Params& p, int sync_mode, bool regs, int G, int threads, cudaStream_t stream) {
    const size_t smem = smem_bytes(p, regs);
    switch (sync_mode) {
        case SYNC_BLOCK:
            return regs ? launch_variant<SYNC_BLOCK, true>(p, 1, threads, smem, stream)
                        : launch_variant<SYNC_BLOCK, false>(p, 1, threads, smem, stream);
        case SYNC_CLUSTER:
            return regs ? launch_variant<SYNC_CLUSTER, true>(p, G, threads, smem, stream)
                        : launch_variant<SYNC_CLUSTER, false>(p, G, threads, smem, stream);
        default:
            return regs ? launch_variant<SYNC_GRID, true>(p, G, threads, smem, stream)
                        : launch_variant<SYNC_GRID, false>(p, G, threads, smem, stream);
    }
}

// dst[b][r][0..ld) <- instance b's rows x cols operator, from the sequential ([rows][cols]) or the
// flipped ([cols][rows]) user layout, zero padded to the row stride ld
__global__ void convert_ops_kernel(float* __restrict__ dst, const float* __restrict__ src, int B, int rows, int cols, int ld,
                                   int flipped) {
    const size_t total = (size_t)B * rows * ld;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(idx % ld);
        const size_t br = idx / ld;
        const int r = (int)(br % rows);
        const size_t b = br / rows;
        float v = 0.f;
        if (cc < cols) v = src[b * rows * cols + (flipped ? (size_t)cc * rows + r : (size_t)r * cols + cc)];
        dst[idx] = v;
    }
}

int launch_convert_ops(float* dst, const float* src, int B, int rows, int cols, int ld, bool flipped, cudaStream_t stream) {
    const size_t total = (size_t)B * rows * ld;
    size_t g = (total + 255) / 256;
    if (g > 148 * 32) g = 148 * 32;
    convert_ops_kernel<<<(int)(g ? g : 1), 256, 0, stream>>>(dst, src, B, rows, cols, ld, flipped ? 1 : 0);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int max_cluster_size(int threads, size_t smem) {
    // largest cluster (<=16) the driver will co-schedule with this much shared memory
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    auto kern = (void*)gpad_latency_kernel<SYNC_CLUSTER, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cfg.gridDim = dim3(16);
    int cs = 0;
    if (cudaOccupancyMaxPotentialClusterSize(&cs, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 8; }
    return cs;
}

}  // namespace lat
}  // namespace gpad
