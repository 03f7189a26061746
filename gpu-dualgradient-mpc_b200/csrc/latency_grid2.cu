// latency_grid2.cu -- latency mode for ONE large QP on the whole chip, second generation (fixed-iteration solves).
//
// ncu on the generic grid plan (latency.cu, battery (10,100): 9.2 us per iteration) showed ~70 % of all warp samples
// inside the two GEMV phases -- generic index arithmetic, a whole residual operator row streamed from L2 by a single
// warp, both operand streams read from shared memory -- and ~28 % in the two grid barriers.  This kernel keeps the
// reference's iteration (steps 1-4 of kernel_functions.cu:7-200 in main.cu:160-175 order) but lays the work out for
// the machine:
//   * COLUMN partition inside the CTA: thread t owns the same few float4 column chunks of every own row, so the
//     exchanged vector (w_v, zhat_v) lives in REGISTERS -- each thread loads only its own chunks straight from the
//     global exchange buffer after the barrier; nothing is gathered into shared memory;
//   * phase A (own rows of M_G, K = m): operator rows resident in shared memory, one LDS.128 per 4 FMAs;
//     phase B (own rows of G_L, K = n): the thread's operator fragments live in registers for the whole solve
//     (RBH float4), so phase B reads no shared memory at all; together the operators of a (10,100) problem
//     (33.6 MB) are on chip: 117.6 KB of shared memory + 116 KB of registers per SM;
//   * per-row partial sums: warp butterflies, one shared-memory hop across warps, finalised by one thread per row
//     which also owns that row's state (z, g_P / y_v, y_{v-1}, p_D, w_i) in registers;
//   * exchange: plain stores to a global vector + the red.release / ld.acquire counter barrier of latency.cu
//     (flag-per-CTA and multi-counter barriers measured slower, see GridBarrier).
// Termination: all three branches of SURVEY row T (z / zhat feasibility, absolute and relative gap, and the dual-gap branch,
// which runs the two phases once more on y_{v+1}).
#include <cuda_runtime.h>

#include "gpad_internal.h"
#include "latency.h"
#include "lat_util.cuh"

namespace gpad {
namespace lat {

namespace {

constexpr int kT = 512;           // threads per CTA
constexpr int kRA = 8;            // max own rows of M_G per CTA

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    return acc;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid-wide barrier: one release-add per CTA on a single counter, thread 0 spins with ld.acquire.  Measured on B200
// (battery (10,100), 100 iterations): this counter 578 us; arrivals spread over 4 counters polled with relaxed loads
// + one fence 703 us; one epoch flag per CTA polled by 148 threads 917 us, by one warp 1125 us; {value, tag} records
// validated by the consumers + relaxed arrive (no fence) 610 us after the shuffle rework (this barrier: 514 us).
struct GridBarrier {
    unsigned epoch = 0;
    __device__ __forceinline__ void sync(unsigned* counter) {
        __syncthreads();                      // this CTA's published stores precede thread 0's release
        if (threadIdx.x == 0) {
            epoch += gridDim.x;
            red_release_add(counter, 1u);
            while (ld_acquire(counter) < epoch) {}
        }
        __syncthreads();
    }
};

// KA: float4 chunks of w per thread (ceil(mld / 4 / 512)); RBH: own rows of G_L per thread group
template <int KA, int RBH, bool CHECK>
__global__ void __launch_bounds__(kT, 1) gpad_grid2_kernel(const Params p, int H) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x;
    const int n = p.n, m = p.m, nld = p.nld, mld = p.mld;
    const int a0 = min(n, c * p.rows_a), na = min(n, a0 + p.rows_a) - a0;
    const int b0 = min(m, c * p.rows_b), nb = min(m, b0 + p.rows_b) - b0;
    const int CA = mld >> 2, CB = nld >> 2;
    const int TH = kT / H;                     // threads per phase-B row group
    const int h = tid / TH, tc = tid - h * TH; // phase B: row group and column chunk of this thread
    const int wih = (tid >> 5) - h * (TH >> 5), wph = TH >> 5;   // warp inside its group, warps per group

    float* ops_a = smem;                              // [kRA][mld] own rows of M_G (zero rows beyond na)
    float* scr_a = ops_a + (size_t)kRA * mld;         // [kRA][16]
    float* scr_b = scr_a + kRA * 16;                  // [H * RBH][wph]
    float* st_a = scr_b + H * RBH * wph;              // [4][kRA]   z, zhat, g_P, f of the own phase-A rows
    float* st_b = st_a + 4 * kRA;                     // [6][64]    y_v, y_{v-1}, p_D, w_i, y_{v+1}, sbar of the own phase-B rows
    float* red_l = st_b + 6 * 64;                     // [3][8]     termination partials: two phase-B finalising warps, phase A
    float* red_g = red_l + 24;                        // [8]        grid-wide termination quantities

    // ---- prologue: operators on chip, initial state ----
    for (int i = tid; i < kRA * CA; i += kT) {
        const int r = i / CA;
        reinterpret_cast<float4*>(ops_a)[i] = r < na ? __ldg(reinterpret_cast<const float4*>(p.M_G + (size_t)(a0 + r) * mld) + (i - r * CA))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 gl[RBH];                                   // G_L[b0 + h * RBH + j][4 tc .. 4 tc + 3]
#pragma unroll
    for (int j = 0; j < RBH; ++j) {
        const int r = h * RBH + j;
        gl[j] = (r < nb && tc < CB) ? __ldg(reinterpret_cast<const float4*>(p.G_L + (size_t)(b0 + r) * nld) + tc)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // step 1 of iteration 0 straight from the caller's vectors (every CTA computes its own chunks of w_0)
    float4 wv[KA];
    {
        const float beta0 = p.beta[0];
#pragma unroll
        for (int k = 0; k < KA; ++k) {
            float e[4] = {0.f, 0.f, 0.f, 0.f};
            const int ch = tid + k * kT;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = 4 * ch + q;
                if (ch < CA && i < m && p.y0) {
                    const float y = p.y0[i], yp = p.y_prev0 ? p.y_prev0[i] : 0.f;
                    e[q] = __fadd_rn(y, __fmul_rn(beta0, __fsub_rn(y, yp)));
                }
            }
            wv[k] = make_float4(e[0], e[1], e[2], e[3]);
        }
    }
    // row state lives in shared memory, touched only by the finalising threads (phase A: threads 0..na-1, phase B:
    // threads 0..nb-1): keeping it in registers of all 512 threads pushed the kernel into local-memory spills
    float* z_r = st_a + tid; float* zh_r = st_a + kRA + tid; float* gp_r = st_a + 2 * kRA + tid; float* f_r = st_a + 3 * kRA + tid;
    float* yv_r = st_b + tid; float* yp_r = st_b + 64 + tid; float* pd_r = st_b + 128 + tid;
    float* w_r = st_b + 192 + tid; float* yn_r = st_b + 256 + tid; float* sb_r = st_b + 320 + tid;
    if (tid < kRA) { *z_r = 0.f; *zh_r = 0.f; *gp_r = (tid < na) ? p.g_P[a0 + tid] : 0.f; *f_r = (tid < na && p.f) ? p.f[a0 + tid] : 0.f; }
    if (tid < 64) {
        const bool ok = tid < nb;
        const float yv = (ok && p.y0) ? p.y0[b0 + tid] : 0.f, yp = (ok && p.y_prev0) ? p.y_prev0[b0 + tid] : 0.f;
        *yv_r = yv; *yp_r = yp; *pd_r = ok ? p.p_D[b0 + tid] : 0.f;
        *w_r = __fadd_rn(yv, __fmul_rn(p.beta[0], __fsub_rn(yv, yp)));
        *yn_r = yv;
        *sb_r = 0.f;
    }
    GridBarrier bar;
    int iters = 0, status = GPAD_STATUS_MAX_ITER;
    float out_viol = __int_as_float(0x7fc00000), out_gap = __int_as_float(0x7fc00000);
    int until_check = CHECK ? p.check_every : 0x7fffffff, check_count = 0;
    __syncthreads();

    // own rows of M_G times a full m-vector held as register chunks: row sums land in scr_a[r][0..15]
    auto rows_a = [&](const float4 (&vec)[KA]) {
        float acc[kRA];
#pragma unroll
        for (int r = 0; r < kRA; ++r) acc[r] = 0.f;
#pragma unroll
        for (int k = 0; k < KA; ++k) {
            const int ch = tid + k * kT;
            if (ch < CA) {
#pragma unroll
                for (int r = 0; r < kRA; ++r)
                    acc[r] = dot4(reinterpret_cast<const float4*>(ops_a + (size_t)r * mld)[ch], vec[k], acc[r]);
            }
        }
        const float tot = warp_sum_transposed<kRA>(acc, lane);       // lanes 4r .. 4r+3 hold row r
        if ((lane & 3) == 0) scr_a[(lane >> 2) * 16 + warp] = tot;
        __syncthreads();
    };
    auto sum_a = [&]() -> float {       // finalising thread tid < na
        const float4* s4 = reinterpret_cast<const float4*>(scr_a + tid * 16);
        const float4 s0 = s4[0], s1 = s4[1], s2 = s4[2], s3 = s4[3];
        return ((s0.x + s0.y) + (s0.z + s0.w)) + ((s1.x + s1.y) + (s1.z + s1.w)) +
               ((s2.x + s2.y) + (s2.z + s2.w)) + ((s3.x + s3.y) + (s3.z + s3.w));
    };
    // own rows of G_L (register fragments) times the n-vector chunk zc: row sums land in scr_b[row][0..wph-1]
    auto rows_b = [&](const float4 zc) {
        float acc[RBH];
#pragma unroll
        for (int j = 0; j < RBH; ++j) acc[j] = dot4(gl[j], zc, 0.f);
        const float tot = warp_sum_transposed<RBH>(acc, lane);       // lanes (32 / RBH) j .. hold row j of the group
        constexpr int kSh = RBH == 8 ? 2 : 1;
        if ((lane & ((1 << kSh) - 1)) == 0) scr_b[(h * RBH + (lane >> kSh)) * wph + wih] = tot;
        __syncthreads();
    };
    auto sum_b = [&]() -> float {       // finalising thread tid < nb
        float d = 0.f;
        for (int k = 0; k < wph; ++k) d += scr_b[tid * wph + k];
        return d;
    };
    // sum over all CTAs of up to 3 per-CTA partials published in xr[k * g_pad + c]; every CTA gets the same totals
    auto grid_sum3 = [&](float* xr, float a, float b, float c3) {      // a, b, c3 valid in thread 0
        if (tid == 0) { xr[0 * p.g_pad + c] = a; xr[1 * p.g_pad + c] = b; xr[2 * p.g_pad + c] = c3; }
        bar.sync(p.barrier);
        if (warp == 0) {
            float x0 = 0.f, x1 = 0.f, x2 = 0.f;
            for (int k = lane; k < (int)gridDim.x; k += 32) {
                x0 += __ldcg(xr + 0 * p.g_pad + k); x1 += __ldcg(xr + 1 * p.g_pad + k); x2 += __ldcg(xr + 2 * p.g_pad + k);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                x0 += __shfl_xor_sync(0xffffffffu, x0, o); x1 += __shfl_xor_sync(0xffffffffu, x1, o); x2 += __shfl_xor_sync(0xffffffffu, x2, o);
            }
            if (lane == 0) { red_g[0] = x0; red_g[1] = x1; red_g[2] = x2; }
        }
        __syncthreads();
    };

    float theta_pf = p.theta[0];
    float beta_pf = p.max_iter > 1 ? p.beta[1] : 0.f;
    for (int v = 0; v < p.max_iter; ++v) {
        const float theta = theta_pf, one_minus = 1.0f - theta;
        const bool last = (v + 1 == p.max_iter);
        const float beta_next = last ? 0.f : beta_pf;
        if (!last) {
            theta_pf = __ldg(p.theta + v + 1);
            beta_pf = (v + 2 < p.max_iter) ? __ldg(p.beta + v + 2) : 0.f;
        }
        const bool check = CHECK && (--until_check == 0);
        if (check) { until_check = p.check_every; ++check_count; }

        // ---------------- phase A: zhat rows (step 2), z average (step 3) ----------------
        rows_a(wv);
        if (tid < 32) {
            float fz = 0.f;
            if (tid < na) {
                const float zh = sum_a() - *gp_r;
                *zh_r = zh;
                *z_r = __fadd_rn(__fmul_rn(one_minus, *z_r), __fmul_rn(theta, zh));
                p.x_zhat[a0 + tid] = zh;
                fz = *f_r * zh;
            }
            if (check && p.f) {       // f'zhat partial of this CTA (rows live in lanes 0..7 of warp 0)
                fz += __shfl_xor_sync(0xffffffffu, fz, 4); fz += __shfl_xor_sync(0xffffffffu, fz, 2); fz += __shfl_xor_sync(0xffffffffu, fz, 1);
                if (tid == 0) red_l[16] = fz;
            }
        }
        bar.sync(p.barrier);
        float4 zc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tc < CB) zc = ld_cg4(p.x_zhat + 4 * tc);

        // ---------------- phase B: dual step + projection (step 4), momentum (step 1 of v+1) ----------------
        {
            rows_b(zc);
            // termination partials of this row (acceldualgrad.m:66-79; same quantities as latency.cu)
            float r_max_sbar = -INFINITY, r_max_rhat = -INFINITY, r_min_w = INFINITY, r_w_rhat = 0.f, r_w_dot = 0.f, r_bad = 0.f;
            if (tid < nb) {
                const float d = sum_b();
                const float wi = *w_r, pd = *pd_r;
                const float s = d + (wi + pd);
                const float yn = 0.5f * (s + fabsf(s));
                *yn_r = yn;
                if (CHECK) {
                    const float rhat = d + pd;
                    const float sb = __fadd_rn(__fmul_rn(one_minus, *sb_r), __fmul_rn(theta, rhat));
                    *sb_r = sb;
                    if (check) {
                        r_max_sbar = sb; r_max_rhat = rhat; r_min_w = wi;
                        r_w_rhat = wi * rhat; r_w_dot = wi * d;
                        r_bad = isfinite(yn) ? 0.f : 1.f;
                    }
                }
                if (!check && !last) {
                    // advance: w_{v+1}, y_{v-1} <- y_v <- y_{v+1}; not on the last iteration: w_v / y_v are outputs
                    const float yv = *yv_r;
                    const float wn = __fadd_rn(yn, __fmul_rn(beta_next, __fsub_rn(yn, yv)));
                    *w_r = wn;
                    p.x_w[b0 + tid] = wn;
                    *yp_r = yv;
                    *yv_r = yn;
                }
            }
            iters = v + 1;
            if (check) {
                // ---- termination test: every CTA reduces the same numbers in the same order and takes the same decision ----
                float* xred = p.x_red + (check_count & 1) * 8 * p.g_pad;
                if (warp < 2) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        r_max_sbar = fmaxf(r_max_sbar, __shfl_xor_sync(0xffffffffu, r_max_sbar, o));
                        r_max_rhat = fmaxf(r_max_rhat, __shfl_xor_sync(0xffffffffu, r_max_rhat, o));
                        r_min_w = fminf(r_min_w, __shfl_xor_sync(0xffffffffu, r_min_w, o));
                        r_w_rhat += __shfl_xor_sync(0xffffffffu, r_w_rhat, o);
                        r_w_dot += __shfl_xor_sync(0xffffffffu, r_w_dot, o);
                        r_bad = fmaxf(r_bad, __shfl_xor_sync(0xffffffffu, r_bad, o));
                    }
                    if (lane == 0) {
                        float* r = red_l + warp * 8;
                        r[0] = r_max_sbar; r[1] = r_max_rhat; r[2] = r_min_w; r[3] = r_w_rhat; r[4] = r_w_dot; r[5] = r_bad;
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    xred[0 * p.g_pad + c] = fmaxf(red_l[0], red_l[8]);
                    xred[1 * p.g_pad + c] = fmaxf(red_l[1], red_l[9]);
                    xred[2 * p.g_pad + c] = fminf(red_l[2], red_l[10]);
                    xred[3 * p.g_pad + c] = red_l[3] + red_l[11];
                    xred[4 * p.g_pad + c] = red_l[4] + red_l[12];
                    xred[5 * p.g_pad + c] = fmaxf(red_l[5], red_l[13]);
                    xred[6 * p.g_pad + c] = p.f ? red_l[16] : 0.f;
                }
                bar.sync(p.barrier);
                if (warp == 0) {
                    float a = -INFINITY, b = -INFINITY, cm = INFINITY, d = 0.f, e = 0.f, g = 0.f, fzs = 0.f;
                    for (int k = lane; k < (int)gridDim.x; k += 32) {
                        fzs += __ldcg(xred + 6 * p.g_pad + k);
                        a = fmaxf(a, __ldcg(xred + 0 * p.g_pad + k)); b = fmaxf(b, __ldcg(xred + 1 * p.g_pad + k));
                        cm = fminf(cm, __ldcg(xred + 2 * p.g_pad + k)); d += __ldcg(xred + 3 * p.g_pad + k);
                        e += __ldcg(xred + 4 * p.g_pad + k); g = fmaxf(g, __ldcg(xred + 5 * p.g_pad + k));
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
                        cm = fminf(cm, __shfl_xor_sync(0xffffffffu, cm, o)); d += __shfl_xor_sync(0xffffffffu, d, o);
                        e += __shfl_xor_sync(0xffffffffu, e, o); g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, o));
                        fzs += __shfl_xor_sync(0xffffffffu, fzs, o);
                    }
                    if (lane == 0) { red_g[3] = d; red_g[4] = e; red_g[5] = g; red_g[6] = fzs; red_g[7] = cm; red_g[0] = a; red_g[1] = b; }
                }
                __syncthreads();
                const float viol_z = p.L * red_g[0], viol_zhat = p.L * red_g[1];
                const float min_w = red_g[7], w_rhat = red_g[3], w_dot = red_g[4], bad = red_g[5], fzhat = red_g[6];
                out_viol = viol_z;
                bool stop = false;
                if (bad > 0.f) { status = GPAD_STATUS_NONFINITE; stop = true; }
                else if (viol_z <= p.eps_g) { status = GPAD_STATUS_CONVERGED_Z; stop = true; }
                else if (viol_zhat <= p.eps_g) {
                    const float V = 0.5f * (fzhat - p.L * w_dot);
                    if (min_w >= 0.f) {
                        const float gapv = -p.L * w_rhat;
                        out_gap = gapv;
                        if (gapv <= p.eps_V || (p.f && gapv <= V * p.eps_V / (1.0f + p.eps_V))) {
                            status = GPAD_STATUS_CONVERGED_ZHAT; out_viol = viol_zhat; stop = true;
                        }
                    } else if (p.f) {
                        // dual branch: Phi(y_{v+1}) needs z_y = M_G y+ - g_P and G_L z_y: the two phases once more on y+,
                        // through the same exchange buffers (nothing of iteration v is still read from them)
                        if (tid < nb) p.x_w[b0 + tid] = *yn_r;
                        bar.sync(p.barrier);
                        float4 yv4[KA];
#pragma unroll
                        for (int k = 0; k < KA; ++k) {
                            const int ch = tid + k * kT;
                            yv4[k] = ch < CA ? ld_cg4(p.x_w + 4 * ch) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        rows_a(yv4);
                        if (tid < 32) {
                            float fzy = 0.f;
                            if (tid < na) {
                                const float zy = sum_a() - *gp_r;
                                p.x_zhat[a0 + tid] = zy;
                                fzy = *f_r * zy;
                            }
                            fzy += __shfl_xor_sync(0xffffffffu, fzy, 4); fzy += __shfl_xor_sync(0xffffffffu, fzy, 2); fzy += __shfl_xor_sync(0xffffffffu, fzy, 1);
                            if (tid == 0) red_l[16] = fzy;
                        }
                        bar.sync(p.barrier);
                        float4 zy4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (tc < CB) zy4 = ld_cg4(p.x_zhat + 4 * tc);
                        rows_b(zy4);
                        float y_gz = 0.f, y_pd = 0.f;
                        if (tid < nb) {
                            const float dd = sum_b(), yn = *yn_r;
                            y_gz = yn * dd; y_pd = yn * *pd_r;
                        }
                        if (warp < 2) {
#pragma unroll
                            for (int o = 16; o; o >>= 1) { y_gz += __shfl_xor_sync(0xffffffffu, y_gz, o); y_pd += __shfl_xor_sync(0xffffffffu, y_pd, o); }
                            if (lane == 0) { red_l[warp * 8 + 0] = y_gz; red_l[warp * 8 + 1] = y_pd; }
                        }
                        __syncthreads();
                        grid_sum3(p.x_red + 2 * 8 * p.g_pad, red_l[16], red_l[0] + red_l[8], red_l[1] + red_l[9]);
                        const float Phi = 0.5f * red_g[0] + 0.5f * p.L * red_g[1] + p.L * red_g[2];
                        const float gapv = V - Phi;
                        out_gap = gapv;
                        if (gapv <= p.eps_V * fmaxf(Phi, 1.0f)) { status = GPAD_STATUS_CONVERGED_DUAL; out_viol = viol_zhat; stop = true; }
                    }
                }
                if (stop) break;
                if (!last && tid < nb) {
                    const float yn = *yn_r, yv = *yv_r;
                    const float wn = __fadd_rn(yn, __fmul_rn(beta_next, __fsub_rn(yn, yv)));
                    *w_r = wn;
                    p.x_w[b0 + tid] = wn;
                    *yp_r = yv;
                    *yv_r = yn;
                }
            }
        }
        if (!last) {
            bar.sync(p.barrier);
#pragma unroll
            for (int k = 0; k < KA; ++k) {
                const int ch = tid + k * kT;
                wv[k] = ch < CA ? ld_cg4(p.x_w + 4 * ch) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }

    // ---------------- outputs (main.cu:176-180): y_I, y_{I-1}, z_{I-1}, zhat_{I-1}, w_{I-1} ----------------
    if (tid < nb) {
        p.out_y_next[b0 + tid] = *yn_r;
        p.out_y[b0 + tid] = *yv_r;
        p.out_w[b0 + tid] = *w_r;
        if (!isfinite(*yn_r)) atomicExch(p.nonfinite_flag, 1);
    }
    if (tid < na) {
        p.out_z[a0 + tid] = *z_r;
        p.out_zhat[a0 + tid] = *zh_r;
    }
    bar.sync(p.barrier);
    if (c == 0 && tid == 0) {
        if (status == GPAD_STATUS_MAX_ITER && *reinterpret_cast<volatile int*>(p.nonfinite_flag)) status = GPAD_STATUS_NONFINITE;
        *p.out_iters = iters;
        *p.out_status = status;
        *p.out_max_viol = out_viol;
        *p.out_gap = out_gap;
    }
}

template <int KA, int RBH>
int launch_t(const Params& p, int G, int H, size_t smem, cudaStream_t s) {
    auto kern = p.check_every > 0 ? gpad_grid2_kernel<KA, RBH, true> : gpad_grid2_kernel<KA, RBH, false>;
    GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Params pc = p;
    void* args[] = {&pc, &H};
    GPAD_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(G), dim3(kT), args, smem, s));
    return GPAD_OK;
}

}  // namespace

// phase-B row groups: the n/4 column chunks of zhat are spread over 512 / H threads
static int groups_for(int nld) {
    const int CB = nld >> 2;
    return CB <= 128 ? 4 : CB <= 256 ? 2 : CB <= 512 ? 1 : 0;
}

size_t grid2_smem_bytes(const Params& p) {
    const int H = groups_for(p.nld);
    const int rbh = H ? (p.rows_b + H - 1) / H : 0;
    return ((size_t)kRA * p.mld + kRA * 16 + (size_t)H * (rbh <= 8 ? 8 : 16) * (kT / (H ? H : 1) / 32) + 4 * kRA + 6 * 64 + 24 + 8) * sizeof(float);
}

// 1 when this plan covers the problem (p.rows_a / rows_b / mld / nld of the generic grid plan with G CTAs)
int grid2_supported(const Params& p, size_t smem_limit) {
    const int H = groups_for(p.nld);
    if (!H) return 0;
    if (p.rows_a > kRA) return 0;
    if ((p.mld >> 2) > 4 * kT) return 0;
    if ((p.rows_b + H - 1) / H > 16) return 0;
    return grid2_smem_bytes(p) <= smem_limit;
}

int launch_grid2(const Params& p, int G, cudaStream_t s) {
    const int H = groups_for(p.nld);
    const int KA = ((p.mld >> 2) + kT - 1) / kT;
    const int rbh = (p.rows_b + H - 1) / H;
    const size_t smem = grid2_smem_bytes(p);
    if (rbh <= 8) {
        switch (KA) {
            case 1: return launch_t<1, 8>(p, G, H, smem, s);
            case 2: return launch_t<2, 8>(p, G, H, smem, s);
            case 3: return launch_t<3, 8>(p, G, H, smem, s);
            default: return launch_t<4, 8>(p, G, H, smem, s);
        }
    }
    switch (KA) {
        case 1: return launch_t<1, 16>(p, G, H, smem, s);
        case 2: return launch_t<2, 16>(p, G, H, smem, s);
        case 3: return launch_t<3, 16>(p, G, H, smem, s);
        default: return launch_t<4, 16>(p, G, H, smem, s);
    }
}

}  // namespace lat
}  // namespace gpad
