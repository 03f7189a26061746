// latency_grid.cu -- latency mode for ONE LARGE QP (operators of tens of MB: the reference's main.cu
// dataset, battery n_u=10, N=100: n=1000, m=4200, 33.6 MB): the whole chip cooperates on one solve.
//
// G = #SMs CTAs (cooperative launch, 512 threads) split the rows of M_G and G_L; as many of a CTA's rows
// as fit stay in shared memory for the whole solve, the rest streams from L2.  Two things differ from
// the generic kernel in latency.cu:
//
//  * register-blocked GEMV: the 16 warps of a CTA form cw column slices x rg row groups; a warp loads each
//    float4 chunk of the shared vector ONCE and reuses it for its rb rows (the generic kernel re-read the
//    vector for every row, doubling shared-memory traffic); column-slice partials are combined through
//    shared memory in a fixed order.
//  * barrier-free global exchange ("flag in data", as in NCCL's LL protocol): every exchanged entry is one
//    64-bit word {fp32 value, 32-bit stamp} written with a single relaxed 8-byte store; consumers poll the
//    words they need until the stamp of the current exchange appears.  No atomic counter, no release
//    fence, no separate gather pass: one store -> L2 -> load per exchange instead of
//    store / fence / atomic / poll / load.
//
// Same arithmetic, outputs and termination test as latency.cu (row T of SURVEY 8a).
#include <algorithm>

#include "gpad_internal.h"
#include "latency.h"

namespace gpad {
namespace lat {

namespace {

constexpr int kWarps = 16;
constexpr int kMaxRB = 4;       // rows per warp and pass
constexpr int kMaxCH = 6;       // float4 chunks per lane

__device__ __forceinline__ float dot4g(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    return acc;
}
__device__ __forceinline__ float wsum_g(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void ll_store(unsigned long long* p, float v, unsigned stamp) {
    const unsigned long long word = ((unsigned long long)stamp << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(word) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
// poll len {value, stamp} words into dst (shared memory); 8 loads in flight per thread
__device__ __forceinline__ void ll_gather(const unsigned long long* src, float* dst, int len, unsigned stamp) {
    constexpr int U = 8;
    for (int base = threadIdx.x; base < len; base += blockDim.x * U) {
        unsigned long long e[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int i = base + j * blockDim.x;
            e[j] = i < len ? ll_load(src + i) : ((unsigned long long)stamp << 32);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int i = base + j * blockDim.x;
            if (i < len) {
                while ((unsigned)(e[j] >> 32) != stamp) { __nanosleep(40); e[j] = ll_load(src + i); }
                dst[i] = __uint_as_float((unsigned)e[j]);
            }
        }
    }
}

// acc[r] = partial <row (r0 + r), x> over this warp's column slice; rows beyond nrows give 0
__device__ __forceinline__ void block_dot(const float* ops_s, const float* ops_g, int res_rows, int row_base_g, int ld,
                                          int r0, int nrows, int rb, const float4* x4, int cw, int ncw, int ch, int lane,
                                          float (&acc)[kMaxRB]) {
    float4 xk[kMaxCH];
#pragma unroll
    for (int k = 0; k < kMaxCH; ++k)
        if (k < ch) xk[k] = x4[(k * ncw + cw) * 32 + lane];
#pragma unroll
    for (int r = 0; r < kMaxRB; ++r) {
        acc[r] = 0.f;
        const int row = r0 + r;
        if (r < rb && row < nrows) {
            const float* base = row < res_rows ? ops_s + (size_t)row * ld : ops_g + (size_t)(row_base_g + row) * ld;
            const float4* o4 = reinterpret_cast<const float4*>(base);
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int k = 0; k < kMaxCH; k += 2) {
                if (k < ch) s0 = dot4g(o4[(k * ncw + cw) * 32 + lane], xk[k], s0);
                if (k + 1 < ch) s1 = dot4g(o4[((k + 1) * ncw + cw) * 32 + lane], xk[k + 1], s1);
            }
            acc[r] = s0 + s1;
        }
    }
}

__global__ void __launch_bounds__(kWarps * 32, 1) gpad_grid_kernel(const Params p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, c = blockIdx.x;
    const int n = p.n, m = p.m, nld = p.nld, mld = p.mld;
    const int a0 = min(n, c * p.rows_a), na = min(n, a0 + p.rows_a) - a0;
    const int b0 = min(m, c * p.rows_b), nb = min(m, b0 + p.rows_b) - b0;

    float* w_s = smem;                          // [mld]
    float* zh_s = w_s + mld;                    // [nld]
    float* part_s = zh_s + nld;                 // [max(rows_a, rows_b) padded][16] column-slice partials
    float* red_s = part_s + max(p.rows_a_pad, p.rows_b_pad) * 16;  // [G_pad * 8] termination partials of all CTAs
    float* scr = red_s + p.g_pad * 8;           // [8 * 16]
    float* ops_s = scr + 8 * 16;                // [res_a][mld] then [res_b][nld]
    const float* opsA_s = ops_s;
    const float* opsB_s = ops_s + (size_t)p.res_a * mld;

    // ---- prologue ----
    {
        const int ra_rows = min(na, p.res_a), rb_rows = min(nb, p.res_b);
        const float4* src = reinterpret_cast<const float4*>(p.M_G + (size_t)a0 * mld);
        float4* dst = reinterpret_cast<float4*>(ops_s);
        for (int i = tid; i < ra_rows * (mld >> 2); i += blockDim.x) dst[i] = __ldg(src + i);
        src = reinterpret_cast<const float4*>(p.G_L + (size_t)b0 * nld);
        dst = reinterpret_cast<float4*>(ops_s + (size_t)p.res_a * mld);
        for (int i = tid; i < rb_rows * (nld >> 2); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    const float beta0 = p.beta[0];
    for (int i = tid; i < mld; i += blockDim.x) {
        float wv = 0.f;
        if (i < m) {
            const float y = p.y0 ? p.y0[i] : 0.f, yp = p.y_prev0 ? p.y_prev0[i] : 0.f;
            wv = __fadd_rn(y, __fmul_rn(beta0, __fsub_rn(y, yp)));
        }
        w_s[i] = wv;
    }
    for (int i = tid; i < nld; i += blockDim.x) zh_s[i] = 0.f;
    // per-row state lives in the registers of thread t (row t of the CTA's slice)
    const bool own_a = tid < na, own_b = tid < nb;
    float z_r = 0.f, gp_r = 0.f, f_r = 0.f, zh_r = 0.f;
    float yv = 0.f, yp = 0.f, yn = 0.f, pd_r = 0.f, w_r = 0.f, sb_r = 0.f, dot_r = 0.f;
    if (own_a) { gp_r = p.g_P[a0 + tid]; if (p.f) f_r = p.f[a0 + tid]; }
    if (own_b) {
        yv = p.y0 ? p.y0[b0 + tid] : 0.f;
        yp = p.y_prev0 ? p.y_prev0[b0 + tid] : 0.f;
        yn = yv;
        pd_r = p.p_D[b0 + tid];
        w_r = __fadd_rn(yv, __fmul_rn(beta0, __fsub_rn(yv, yp)));
    }
    __syncthreads();

    // warp roles in the two phases
    const int cw_a = warp % p.cwa, rg_a = warp / p.cwa;
    const int cw_b = warp % p.cwb, rg_b = warp / p.cwb;
    const int passes_a = (p.rows_a + p.rga * p.rba - 1) / (p.rga * p.rba);
    const int passes_b = (p.rows_b + p.rgb * p.rbb - 1) / (p.rgb * p.rbb);
    const float4* w4 = reinterpret_cast<const float4*>(w_s);
    const float4* zh4 = reinterpret_cast<const float4*>(zh_s);

    auto product_a = [&]() -> float {          // returns <M_G row tid, w_s> for tid < na
        for (int ps = 0; ps < passes_a; ++ps) {
            const int r0 = (ps * p.rga + rg_a) * p.rba;
            float acc[kMaxRB];
            block_dot(opsA_s, p.M_G, p.res_a, a0, mld, r0, na, p.rba, w4, cw_a, p.cwa, p.cha, lane, acc);
#pragma unroll
            for (int r = 0; r < kMaxRB; ++r) {
                if (r < p.rba) {
                    const float s = wsum_g(acc[r]);
                    if (lane == 0 && r0 + r < p.rows_a) part_s[(r0 + r) * 16 + cw_a] = s;
                }
            }
        }
        __syncthreads();
        float d = 0.f;
        if (own_a) for (int k = 0; k < p.cwa; ++k) d += part_s[tid * 16 + k];
        return d;
    };
    auto product_b = [&]() -> float {
        for (int ps = 0; ps < passes_b; ++ps) {
            const int r0 = (ps * p.rgb + rg_b) * p.rbb;
            float acc[kMaxRB];
            block_dot(opsB_s, p.G_L, p.res_b, b0, nld, r0, nb, p.rbb, zh4, cw_b, p.cwb, p.chb, lane, acc);
#pragma unroll
            for (int r = 0; r < kMaxRB; ++r) {
                if (r < p.rbb) {
                    const float s = wsum_g(acc[r]);
                    if (lane == 0 && r0 + r < p.rows_b) part_s[(r0 + r) * 16 + cw_b] = s;
                }
            }
        }
        __syncthreads();
        float d = 0.f;
        if (own_b) for (int k = 0; k < p.cwb; ++k) d += part_s[tid * 16 + k];
        return d;
    };

    unsigned seq_z = 0, seq_w = 0, seq_r = 0;      // exchange counters (uniform across CTAs)
    auto exchange_z = [&](float v) {               // publish own rows of a length-n vector, gather all of it into zh_s
        const unsigned stamp = p.stamp_base + 3 * seq_z + 1;
        unsigned long long* buf = p.ll_z + (size_t)(seq_z & 1) * n;
        if (own_a) ll_store(buf + a0 + tid, v, stamp);
        ll_gather(buf, zh_s, n, stamp);
        ++seq_z;
        __syncthreads();
    };
    auto exchange_w = [&](float v) {
        const unsigned stamp = p.stamp_base + 3 * seq_w + 2;
        unsigned long long* buf = p.ll_w + (size_t)(seq_w & 1) * m;
        if (own_b) ll_store(buf + b0 + tid, v, stamp);
        ll_gather(buf, w_s, m, stamp);
        ++seq_w;
        __syncthreads();
    };

    const bool checking = p.check_every > 0;
    int iters = 0, status = GPAD_STATUS_MAX_ITER, until_check = checking ? p.check_every : 0x7fffffff;
    float out_viol = __int_as_float(0x7fc00000), out_gap = __int_as_float(0x7fc00000);
    float theta_pf = p.theta[0];
    float beta_pf = p.max_iter > 1 ? p.beta[1] : 0.f;

    for (int v = 0; v < p.max_iter; ++v) {
        const float theta = theta_pf, one_minus = 1.0f - theta;
        const bool last = v + 1 == p.max_iter;
        const float beta_next = last ? 0.f : beta_pf;
        if (!last) {
            theta_pf = __ldg(p.theta + v + 1);
            beta_pf = v + 2 < p.max_iter ? __ldg(p.beta + v + 2) : 0.f;
        }
        const bool check = (--until_check == 0);
        if (check) until_check = p.check_every;

        // ---------------- phase A ----------------
        {
            const float d = product_a();
            if (own_a) {
                zh_r = d - gp_r;
                z_r = __fadd_rn(__fmul_rn(one_minus, z_r), __fmul_rn(theta, zh_r));
            }
            exchange_z(zh_r);
        }
        // ---------------- phase B ----------------
        {
            const float d = product_b();
            if (own_b) {
                const float s = d + (w_r + pd_r);
                yn = 0.5f * (s + fabsf(s));
                dot_r = d;
                if (checking) sb_r = __fadd_rn(__fmul_rn(one_minus, sb_r), __fmul_rn(theta, d + pd_r));
            }
        }
        iters = v + 1;
        bool stop = false;
        if (check) {
            // CTA partials -> flag-in-data exchange of 8 values per CTA -> identical ordered combine everywhere
            const float rhat = dot_r + pd_r;
            float vals[7] = {own_b ? sb_r : -INFINITY, own_b ? rhat : -INFINITY, own_b ? w_r : INFINITY,
                             own_b ? w_r * rhat : 0.f, own_b ? w_r * dot_r : 0.f, own_a ? f_r * zh_r : 0.f,
                             (own_b && !isfinite(yn)) ? 1.f : 0.f};
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                float x = vals[k];
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const float y2 = __shfl_xor_sync(0xffffffffu, x, o);
                    x = (k == 0 || k == 1 || k == 6) ? fmaxf(x, y2) : k == 2 ? fminf(x, y2) : x + y2;
                }
                if (lane == 0) scr[k * 16 + warp] = x;
            }
            __syncthreads();
            const unsigned stamp = p.stamp_base + 3 * seq_r + 3;
            unsigned long long* buf = p.ll_r + (size_t)(seq_r & 1) * G * 8;
            if (tid < 7) {
                float x = scr[tid * 16];
                for (int w2 = 1; w2 < kWarps; ++w2) {
                    const float y2 = scr[tid * 16 + w2];
                    x = (tid == 0 || tid == 1 || tid == 6) ? fmaxf(x, y2) : tid == 2 ? fminf(x, y2) : x + y2;
                }
                ll_store(buf + (size_t)c * 8 + tid, x, stamp);
            }
            if (tid == 7) ll_store(buf + (size_t)c * 8 + 7, 0.f, stamp);
            ll_gather(buf, red_s, G * 8, stamp);
            ++seq_r;
            __syncthreads();
            float max_sbar = -INFINITY, max_rhat = -INFINITY, min_w = INFINITY, w_rhat = 0.f, w_dot = 0.f, fz = 0.f, bad = 0.f;
            for (int k = 0; k < G; ++k) {
                max_sbar = fmaxf(max_sbar, red_s[k * 8 + 0]); max_rhat = fmaxf(max_rhat, red_s[k * 8 + 1]);
                min_w = fminf(min_w, red_s[k * 8 + 2]); w_rhat += red_s[k * 8 + 3]; w_dot += red_s[k * 8 + 4];
                fz += red_s[k * 8 + 5]; bad = fmaxf(bad, red_s[k * 8 + 6]);
            }
            const float viol_z = p.L * max_sbar, viol_zhat = p.L * max_rhat;
            out_viol = viol_z;
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
                    // dual-gap branch: z_y = M_G y+ - g_P and G_L z_y through the same exchanges; w_s / zh_s are
                    // scratch here and are rebuilt afterwards from the owners' registers
                    __syncthreads();
                    exchange_w(yn);                                   // w_s <- y_{v+1}
                    const float zy = product_a() - gp_r;
                    exchange_z(zy);                                   // zh_s <- z_y
                    const float gz = product_b();
                    float r3[3] = {own_a ? f_r * zy : 0.f, own_b ? yn * gz : 0.f, own_b ? yn * pd_r : 0.f};
#pragma unroll
                    for (int k = 0; k < 3; ++k) { r3[k] = wsum_g(r3[k]); if (lane == 0) scr[k * 16 + warp] = r3[k]; }
                    __syncthreads();
                    const unsigned stamp2 = p.stamp_base + 3 * seq_r + 3;
                    unsigned long long* buf2 = p.ll_r + (size_t)(seq_r & 1) * G * 8;
                    if (tid < 8) {
                        float x = 0.f;
                        if (tid < 3) for (int w2 = 0; w2 < kWarps; ++w2) x += scr[tid * 16 + w2];
                        ll_store(buf2 + (size_t)c * 8 + tid, x, stamp2);
                    }
                    ll_gather(buf2, red_s, G * 8, stamp2);
                    ++seq_r;
                    __syncthreads();
                    float s_fzy = 0.f, s_ygz = 0.f, s_ypd = 0.f;
                    for (int k = 0; k < G; ++k) { s_fzy += red_s[k * 8]; s_ygz += red_s[k * 8 + 1]; s_ypd += red_s[k * 8 + 2]; }
                    const float Phi = 0.5f * s_fzy + 0.5f * p.L * s_ygz + p.L * s_ypd;
                    const float gapv = V - Phi;
                    out_gap = gapv;
                    if (gapv <= p.eps_V * fmaxf(Phi, 1.0f)) { status = GPAD_STATUS_CONVERGED_DUAL; out_viol = viol_zhat; stop = true; }
                    __syncthreads();
                    exchange_w(w_r);                                  // restore w_v
                    exchange_z(zh_r);                                 // restore zhat_v
                }
            }
        }
        if (stop) break;
        if (!last) {
            float wn = 0.f;
            if (own_b) {
                wn = __fadd_rn(yn, __fmul_rn(beta_next, __fsub_rn(yn, yv)));
                w_r = wn; yp = yv; yv = yn;
            }
            exchange_w(wn);
        }
    }

    // ---------------- outputs ----------------
    if (own_b) {
        if (p.out_y_next) p.out_y_next[b0 + tid] = yn;
        if (p.out_y) p.out_y[b0 + tid] = yv;
        if (p.out_w) p.out_w[b0 + tid] = w_r;
    }
    if (own_a) {
        if (p.out_z) p.out_z[a0 + tid] = z_r;
        if (p.out_zhat) p.out_zhat[a0 + tid] = zh_r;
    }
    if (status == GPAD_STATUS_MAX_ITER) {
        // a non-finite iterate anywhere turns MAX_ITER into NONFINITE
        const int badf = __syncthreads_or(own_b && !isfinite(yn));
        const unsigned stamp = p.stamp_base + 3 * seq_r + 3;
        unsigned long long* buf = p.ll_r + (size_t)(seq_r & 1) * G * 8;
        if (tid < 8) ll_store(buf + (size_t)c * 8 + tid, badf ? 1.f : 0.f, stamp);
        if (c == 0) {
            ll_gather(buf, red_s, G * 8, stamp);
            __syncthreads();
            if (tid == 0) {
                float b = 0.f;
                for (int k = 0; k < G; ++k) b = fmaxf(b, red_s[k * 8]);
                if (b > 0.f) status = GPAD_STATUS_NONFINITE;
            }
        }
    }
    if (c == 0 && tid == 0) {
        if (p.out_iters) *p.out_iters = iters;
        if (p.out_status) *p.out_status = status;
        if (p.out_max_viol) *p.out_max_viol = out_viol;
        if (p.out_gap) *p.out_gap = out_gap;
    }
}

}  // namespace

size_t grid_smem_bytes(const Params& p) {
    const size_t fl = (size_t)p.mld + p.nld + (size_t)std::max(p.rows_a_pad, p.rows_b_pad) * 16 + (size_t)p.g_pad * 8 + 8 * 16 +
                      (size_t)p.res_a * p.mld + (size_t)p.res_b * p.nld;
    return fl * sizeof(float);
}

int launch_grid(const Params& p, int G, cudaStream_t stream) {
    const size_t smem = grid_smem_bytes(p);
    GPAD_CUDA(cudaFuncSetAttribute(gpad_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&p};
    GPAD_CUDA(cudaLaunchCooperativeKernel((void*)gpad_grid_kernel, dim3(G), dim3(kWarps * 32), args, smem, stream));
    return GPAD_OK;
}

}  // namespace lat
}  // namespace gpad
