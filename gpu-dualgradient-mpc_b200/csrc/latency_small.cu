// latency_small.cu -- the lean single-CTA GPAD kernel: one QP per CTA, operators in registers,
// per-row state (z, g_P, y, y_prev, p_D, w_i, sbar) in the owning lane's registers, only the two
// exchanged vectors (w, zhat) in shared memory, two __syncthreads per iteration and no global
// memory traffic inside the loop (theta/beta come from shared memory).
//
// Used for (a) latency mode on problems small enough for one CTA (the reference's default
// battery problem n=12, m=56, gpad.m:4-5) and (b) GPAD_MODE_BATCH_PER_INSTANCE, the batched-GEMV
// mode of BASELINE config 5: gridDim.x independent QPs, each with its own M_G / G_L, which are read
// from HBM exactly once per solve.  Same arithmetic and termination test as latency.cu
// (steps: kernel_functions.cu:7-14,16-72,142-200; row T: acceldualgrad.m:66-79).
//
// CLUSTER = true runs the same loop on one thread-block cluster of C <= 16 CTAs for a single mid-size QP:
// each CTA owns ceil(n/C) rows of M_G and ceil(m/C) rows of G_L, the exchanged entries of zhat / w are
// stored straight into every CTA's shared memory (DSMEM) and the two __syncthreads become two
// barrier.cluster; the termination partials travel the same way.
//
// CHA / CHB = float4 operator fragments per lane in phase A / B (compile time, so the dot
// products are straight-line code); rows are zero padded to 4 * lanes-per-row * CH floats.
#include <cooperative_groups.h>

#include "gpad_internal.h"
#include "latency.h"

namespace cg = cooperative_groups;

namespace gpad {
namespace lat {

namespace {

constexpr int kMaxSchedSmem = 1024;     // theta/beta entries staged in shared memory

__device__ __forceinline__ float dot4s(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    return acc;
}
__device__ __forceinline__ float group_sum_s(float v, int lpr) {
    for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float wmin(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide reduction of up to 6 values: op 0 = sum, 1 = max, 2 = min; result in every thread
template <int NV>
__device__ __forceinline__ void block_reduce(float (&v)[NV], const int (&op)[NV], float* scr, int warp, int lane, int nwarps) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = op[k] == 0 ? wsum(v[k]) : op[k] == 1 ? wmax(v[k]) : wmin(v[k]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) scr[k * 32 + warp] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        float r = scr[k * 32];
        for (int w = 1; w < nwarps; ++w) {
            const float x = scr[k * 32 + w];
            r = op[k] == 0 ? r + x : op[k] == 1 ? fmaxf(r, x) : fminf(r, x);
        }
        v[k] = r;
    }
}

template <bool CLUSTER>
__device__ __forceinline__ void cta_sync() {
    if (CLUSTER) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
}
// store one exchanged vector entry: locally, or into every CTA of the cluster (distributed shared memory)
template <bool CLUSTER>
__device__ __forceinline__ void put(float* loc, int idx, float v) {
    if (CLUSTER) {
        cg::cluster_group cl = cg::this_cluster();
        const unsigned C = cl.num_blocks();
        for (unsigned r = 0; r < C; ++r) cl.map_shared_rank(loc, r)[idx] = v;
    } else {
        loc[idx] = v;
    }
}

// ---- cluster exchange without cluster-wide barriers: every exchanged entry is an asynchronous 4-byte store
// into each CTA's shared memory that also credits 4 bytes to that CTA's mbarrier (st.async ... complete_tx);
// the consumer waits on its OWN mbarrier until the whole vector (len * 4 bytes) has landed.
__device__ __forceinline__ uint32_t smem_addr(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
// cluster-window addresses are affine in the CTA rank: addr(rank) = addr(0) + rank * stride, so the two mapa
// per destination are hoisted out of the iteration loop (rank0 addresses + stride computed once)
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_all(uint32_t data0, uint32_t bar0, uint32_t stride, float v, int C) {
    const uint32_t bits = __float_as_uint(v);
    for (int r = 0; r < C; ++r, data0 += stride, bar0 += stride)
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                     ::"r"(data0), "r"(bits), "r"(bar0) : "memory");
}
__device__ __forceinline__ void mbar_init_s(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

template <int CHA, int CHB, bool CLUSTER>
__global__ void __launch_bounds__(kMaxThreads) gpad_small_kernel(const Params p_in) {
    extern __shared__ __align__(16) float smem[];
    Params p = p_in;
    if (!CLUSTER) {   // per-instance operators / vectors: every CTA is an independent QP
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
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int n = p.n, m = p.m, mld = p.mld, nld = p.nld;
    float* w_s = smem;                       // [mld]
    float* zh_s = w_s + mld;                 // [nld]
    float* scr = zh_s + nld;                 // [6*32] reduction scratch
    float* cred = scr + 6 * 32;              // [16][8] per-CTA termination partials (cluster variant)
    uint64_t* xbar = reinterpret_cast<uint64_t*>(cred + 16 * 8);   // [2] mbarriers: zhat exchange, w exchange (cluster variant)
    float* th_s = cred + 16 * 8 + 4;         // [sched] theta
    float* be_s = th_s + p.sched_smem;       // [sched] beta

    const int lpr_a = 1 << p.lg_a, lpr_b = 1 << p.lg_b;
    const int rank = CLUSTER ? (int)blockIdx.x : 0, C = CLUSTER ? (int)gridDim.x : 1;
    const int sub_a = tid & (lpr_a - 1), loc_a = tid >> p.lg_a;
    const int sub_b = tid & (lpr_b - 1), loc_b = tid >> p.lg_b;
    // global row indices of this thread's rows (rows_a / rows_b rows per CTA)
    const int row_a = rank * p.rows_a + loc_a, row_b = rank * p.rows_b + loc_b;
    const bool has_a = loc_a < p.rows_a && row_a < n, has_b = loc_b < p.rows_b && row_b < m;
    const bool own_a = sub_a == 0 && has_a, own_b = sub_b == 0 && has_b;
    const float4* w4 = reinterpret_cast<const float4*>(w_s) + sub_a;
    const float4* zh4 = reinterpret_cast<const float4*>(zh_s) + sub_b;

    // ---- operators: HBM -> registers, once ----
    float4 ra[CHA], rb[CHB];
    {
        const float4* src = reinterpret_cast<const float4*>(p.M_G + (size_t)min(row_a, n - 1) * mld) + sub_a;
#pragma unroll
        for (int k = 0; k < CHA; ++k) ra[k] = has_a ? __ldg(src + k * lpr_a) : make_float4(0.f, 0.f, 0.f, 0.f);
        src = reinterpret_cast<const float4*>(p.G_L + (size_t)min(row_b, m - 1) * nld) + sub_b;
#pragma unroll
        for (int k = 0; k < CHB; ++k) rb[k] = has_b ? __ldg(src + k * lpr_b) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- per-row state in the owner's registers ----
    float z_r = 0.f, gp_r = 0.f, f_r = 0.f, zh_r = 0.f;
    float yv = 0.f, yp = 0.f, yn = 0.f, pd_r = 0.f, w_r = 0.f, sb_r = 0.f, dot_r = 0.f;
    const float beta0 = p.beta[0];
    if (own_a) { gp_r = p.g_P[row_a]; if (p.f) f_r = p.f[row_a]; }
    if (own_b) {
        yv = p.y0 ? p.y0[row_b] : 0.f;
        yp = p.y_prev0 ? p.y_prev0[row_b] : 0.f;
        yn = yv;
        pd_r = p.p_D[row_b];
        w_r = __fadd_rn(yv, __fmul_rn(beta0, __fsub_rn(yv, yp)));      // step 1 of iteration 0
    }
    for (int i = tid; i < mld; i += nthr) w_s[i] = 0.f;
    for (int i = tid; i < nld; i += nthr) zh_s[i] = 0.f;
    const bool sched_in_smem = p.max_iter <= p.sched_smem;
    if (sched_in_smem)
        for (int i = tid; i < p.max_iter; i += nthr) { th_s[i] = p.theta[i]; be_s[i] = p.beta[i]; }
    const uint32_t zbar = smem_addr(xbar), wbar = smem_addr(xbar + 1);
    const uint32_t zh_addr = smem_addr(zh_s), w_addr = smem_addr(w_s);
    uint32_t zbar0 = 0, wbar0 = 0, zh0 = 0, w0 = 0, cstride = 0;      // rank-0 views of the exchange targets
    if (CLUSTER) {
        zbar0 = mapa_rank(zbar, 0); wbar0 = mapa_rank(wbar, 0);
        zh0 = mapa_rank(zh_addr, 0) + 4u * row_a; w0 = mapa_rank(w_addr, 0) + 4u * row_b;
        cstride = C > 1 ? mapa_rank(zbar, 1) - zbar0 : 0;
    }
    uint32_t zpar = 0, wpar = 0;             // phase parities of the two exchange barriers
    if (CLUSTER && tid == 0) {
        mbar_init_s(zbar, 1); mbar_init_s(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cta_sync<CLUSTER>();                     // every CTA's vectors are zeroed / barriers initialised before anyone stores into them
    if (own_b) put<CLUSTER>(w_s, row_b, w_r);
    cta_sync<CLUSTER>();

    const bool checking = p.check_every > 0;
    int iters = 0, status = GPAD_STATUS_MAX_ITER, until_check = checking ? p.check_every : 0x7fffffff;
    float out_viol = __int_as_float(0x7fc00000), out_gap = __int_as_float(0x7fc00000);

    for (int v = 0; v < p.max_iter; ++v) {
        const float theta = sched_in_smem ? th_s[v] : __ldg(p.theta + v);
        const bool last = v + 1 == p.max_iter;
        const float beta_next = last ? 0.f : (sched_in_smem ? be_s[v + 1] : __ldg(p.beta + v + 1));
        const float one_minus = 1.0f - theta;
        const bool check = (--until_check == 0);
        if (check) until_check = p.check_every;

        // ---------------- phase A: zhat = M_G w - g_P, z average ----------------
        {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int k = 0; k < CHA; k += 2) {
                s0 = dot4s(ra[k], w4[k * lpr_a], s0);
                if (k + 1 < CHA) s1 = dot4s(ra[k + 1], w4[(k + 1) * lpr_a], s1);
            }
            const float d = group_sum_s(s0 + s1, lpr_a);
            if (own_a) {
                zh_r = d - gp_r;
                z_r = __fadd_rn(__fmul_rn(one_minus, z_r), __fmul_rn(theta, zh_r));
                if (CLUSTER) st_async_all(zh0, zbar0, cstride, zh_r, C);
                else zh_s[row_a] = zh_r;
            }
        }
        if (CLUSTER) {
            if (tid == 0) mbar_expect_s(zbar, 4u * n);       // this CTA expects all n entries of zhat_v
            mbar_wait_s(zbar, zpar); zpar ^= 1;
        } else {
            __syncthreads();
        }
        // ---------------- phase B: y+ = max(G_L zhat + (w + p_D), 0), momentum ----------------
        {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int k = 0; k < CHB; k += 2) {
                s0 = dot4s(rb[k], zh4[k * lpr_b], s0);
                if (k + 1 < CHB) s1 = dot4s(rb[k + 1], zh4[(k + 1) * lpr_b], s1);
            }
            const float d = group_sum_s(s0 + s1, lpr_b);
            if (own_b) {
                const float s = d + (w_r + pd_r);
                yn = 0.5f * (s + fabsf(s));
                dot_r = d;
                if (checking) sb_r = __fadd_rn(__fmul_rn(one_minus, sb_r), __fmul_rn(theta, d + pd_r));
                if (!check && !last) {      // advance; on check / last iterations w_v, y_v stay (they are outputs)
                    const float wn = __fadd_rn(yn, __fmul_rn(beta_next, __fsub_rn(yn, yv)));
                    if (CLUSTER) st_async_all(w0, wbar0, cstride, wn, C);
                    else w_s[row_b] = wn;
                    w_r = wn;
                    yp = yv; yv = yn;
                }
            }
        }
        iters = v + 1;
        if (!check) {
            if (CLUSTER) {
                if (!last) {
                    if (tid == 0) mbar_expect_s(wbar, 4u * m);   // all m entries of w_{v+1}
                    mbar_wait_s(wbar, wpar); wpar ^= 1;
                }
            } else {
                __syncthreads();
            }
            continue;
        }
        if (CLUSTER) cta_sync<CLUSTER>();     // check iterations fall back to cluster barriers (rare)

        // ---------------- termination test ----------------
        const float rhat = dot_r + pd_r;
        float red[6] = {own_b ? sb_r : -INFINITY, own_b ? rhat : -INFINITY, own_b ? w_r : INFINITY,
                        own_b ? w_r * rhat : 0.f, own_b ? w_r * dot_r : 0.f, own_a ? f_r * zh_r : 0.f};
        const int ops[6] = {1, 1, 2, 0, 0, 0};
        float bad[1] = {(own_b && !isfinite(yn)) ? 1.f : 0.f};
        const int opb[1] = {1};
        block_reduce<6>(red, ops, scr, warp, lane, nwarps);
        block_reduce<1>(bad, opb, scr, warp, lane, nwarps);
        if (CLUSTER) {      // per-CTA partials -> every CTA, then the same ordered combine everywhere
            if (tid < 7) put<CLUSTER>(cred, rank * 8 + tid, tid < 6 ? red[tid] : bad[0]);
            cta_sync<CLUSTER>();
            for (int k = 0; k < 6; ++k) {
                float r = cred[k];
                for (int c = 1; c < C; ++c) { const float x = cred[c * 8 + k]; r = ops[k] == 0 ? r + x : ops[k] == 1 ? fmaxf(r, x) : fminf(r, x); }
                red[k] = r;
            }
            float b = cred[6];
            for (int c = 1; c < C; ++c) b = fmaxf(b, cred[c * 8 + 6]);
            bad[0] = b;
            cta_sync<CLUSTER>();        // cred may be reused by the dual-gap branch
        }
        const float viol_z = p.L * red[0], viol_zhat = p.L * red[1];
        out_viol = viol_z;
        bool stop = false;
        if (bad[0] > 0.f) { status = GPAD_STATUS_NONFINITE; stop = true; }
        else if (viol_z <= p.eps_g) { status = GPAD_STATUS_CONVERGED_Z; stop = true; }
        else if (viol_zhat <= p.eps_g) {
            const float V = 0.5f * (red[5] - p.L * red[4]);
            if (red[2] >= 0.f) {
                const float gapv = -p.L * red[3];
                out_gap = gapv;
                if (gapv <= p.eps_V || (p.f && gapv <= V * p.eps_V / (1.0f + p.eps_V))) {
                    status = GPAD_STATUS_CONVERGED_ZHAT; out_viol = viol_zhat; stop = true;
                }
            } else if (p.f) {
                // dual-gap branch: z_y = M_G y+ - g_P, then y+'(G_L z_y): two extra products through w_s / zh_s
                cta_sync<CLUSTER>();
                if (own_b) put<CLUSTER>(w_s, row_b, yn);
                cta_sync<CLUSTER>();
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int k = 0; k < CHA; k += 2) {
                    s0 = dot4s(ra[k], w4[k * lpr_a], s0);
                    if (k + 1 < CHA) s1 = dot4s(ra[k + 1], w4[(k + 1) * lpr_a], s1);
                }
                const float zy = group_sum_s(s0 + s1, lpr_a) - gp_r;
                cta_sync<CLUSTER>();
                if (own_a) put<CLUSTER>(zh_s, row_a, zy);
                cta_sync<CLUSTER>();
                s0 = 0.f; s1 = 0.f;
#pragma unroll
                for (int k = 0; k < CHB; k += 2) {
                    s0 = dot4s(rb[k], zh4[k * lpr_b], s0);
                    if (k + 1 < CHB) s1 = dot4s(rb[k + 1], zh4[(k + 1) * lpr_b], s1);
                }
                const float gz = group_sum_s(s0 + s1, lpr_b);
                float r3[3] = {own_a ? f_r * zy : 0.f, own_b ? yn * gz : 0.f, own_b ? yn * pd_r : 0.f};
                const int op3[3] = {0, 0, 0};
                block_reduce<3>(r3, op3, scr, warp, lane, nwarps);
                if (CLUSTER) {
                    if (tid < 3) put<CLUSTER>(cred, rank * 8 + tid, r3[tid]);
                    cta_sync<CLUSTER>();
                    for (int k = 0; k < 3; ++k) { float r = cred[k]; for (int c = 1; c < C; ++c) r += cred[c * 8 + k]; r3[k] = r; }
                }
                const float Phi = 0.5f * r3[0] + 0.5f * p.L * r3[1] + p.L * r3[2];
                const float gapv = V - Phi;
                out_gap = gapv;
                if (gapv <= p.eps_V * fmaxf(Phi, 1.0f)) { status = GPAD_STATUS_CONVERGED_DUAL; out_viol = viol_zhat; stop = true; }
                cta_sync<CLUSTER>();
                if (own_b) put<CLUSTER>(w_s, row_b, w_r);          // restore w_v / zhat_v
                if (own_a) put<CLUSTER>(zh_s, row_a, zh_r);
                cta_sync<CLUSTER>();
            }
        }
        if (stop) break;
        if (!last && own_b) {
            const float wn = __fadd_rn(yn, __fmul_rn(beta_next, __fsub_rn(yn, yv)));
            put<CLUSTER>(w_s, row_b, wn); w_r = wn;
            yp = yv; yv = yn;
        }
        cta_sync<CLUSTER>();
    }

    // ---------------- outputs (main.cu:176-180 + termination outputs) ----------------
    if (own_b) {
        if (p.out_y_next) p.out_y_next[row_b] = yn;
        if (p.out_y) p.out_y[row_b] = yv;
        if (p.out_w) p.out_w[row_b] = w_r;
    }
    if (own_a) {
        if (p.out_z) p.out_z[row_a] = z_r;
        if (p.out_zhat) p.out_zhat[row_a] = zh_r;
    }
    if (status == GPAD_STATUS_MAX_ITER) {
        int badf = __syncthreads_or(own_b && !isfinite(yn));
        if (CLUSTER) {
            if (tid == 0) put<CLUSTER>(cred, rank * 8 + 7, badf ? 1.f : 0.f);
            cta_sync<CLUSTER>();
            for (int c = 0; c < C; ++c) badf |= cred[c * 8 + 7] > 0.f;
        }
        if (badf) status = GPAD_STATUS_NONFINITE;
    }
    if (CLUSTER) cta_sync<CLUSTER>();          // no CTA exits while a peer may still store into it
    if (rank == 0 && tid == 0) {
        if (p.out_iters) *p.out_iters = iters;
        if (p.out_status) *p.out_status = status;
        if (p.out_max_viol) *p.out_max_viol = out_viol;
        if (p.out_gap) *p.out_gap = out_gap;
    }
}

template <int CHA, int CHB>
int launch_small_t(const Params& p, int cluster, int threads, size_t smem, cudaStream_t stream) {
    if (cluster <= 1) {
        auto kern = gpad_small_kernel<CHA, CHB, false>;
        if (smem > 48 * 1024) GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<p.batch > 1 ? p.batch : 1, threads, smem, stream>>>(p);
        GPAD_CUDA(cudaGetLastError());
        return GPAD_OK;
    }
    auto kern = gpad_small_kernel<CHA, CHB, true>;
    if (smem > 48 * 1024) GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cluster > 8) GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GPAD_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return GPAD_OK;
}

template <int CHA>
int launch_small_a(const Params& p, int chb, int cluster, int threads, size_t smem, cudaStream_t stream) {
    switch (chb) {
        case 1: return launch_small_t<CHA, 1>(p, cluster, threads, smem, stream);
        case 2: return launch_small_t<CHA, 2>(p, cluster, threads, smem, stream);
        case 4: return launch_small_t<CHA, 4>(p, cluster, threads, smem, stream);
        default: return launch_small_t<CHA, 8>(p, cluster, threads, smem, stream);
    }
}

}  // namespace

size_t small_smem_bytes(const Params& p) {
    return ((size_t)p.mld + p.nld + 6 * 32 + 16 * 8 + 4 + 2 * (size_t)p.sched_smem) * sizeof(float);
}

int small_sched_capacity() { return kMaxSchedSmem; }

// cha / chb in {1, 2, 4, 8}: float4 fragments per lane, rows padded accordingly (plan_small in api.cu)
int launch_small(const Params& p, int cha, int chb, int cluster, int threads, cudaStream_t stream) {
    const size_t smem = small_smem_bytes(p);
    switch (cha) {
        case 1: return launch_small_a<1>(p, chb, cluster, threads, smem, stream);
        case 2: return launch_small_a<2>(p, chb, cluster, threads, smem, stream);
        case 4: return launch_small_a<4>(p, chb, cluster, threads, smem, stream);
        default: return launch_small_a<8>(p, chb, cluster, threads, smem, stream);
    }
}

}  // namespace lat
}  // namespace gpad
