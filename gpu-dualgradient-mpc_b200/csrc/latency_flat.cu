// latency_flat.cu -- latency mode on the battery-structured ("flattened") operators of the reference
// (ENABLE_FLATTEN_MATRICES: seq_functions.cpp:5-43, kernel_functions.cu:74-109, Cookbook 2.2), SURVEY 8(f) row 4.
//
// Identical cells decouple the battery problem: row (s, u) of M_G (stage s, cell u) only touches the multipliers of
// cell u's box constraints and of the per-stage sum constraints, and a box row of G_L only touches cell u's inputs:
//     zhat[s,u] = sum_{k < 4 n_u N, k % n_u = u} Mf[s][k] w[k]  +  sum_{k >= 4 n_u N} Mf[s][k] w[k]  -  g_P[s,u]
//     acc_i     = sum_s Gf[i][s] zhat[s, i % n_u]                       (box rows,  i <  4 n_u N)
//     acc_i     = sum_s Gf[i][s] (sum_u zhat[s,u])                      (sum rows,  i >= 4 n_u N)
// with Mf [N][m], Gf [m][N]: n_u times fewer operator bytes and flops than the dense form.  For (10,100) that is
// 3.4 MB instead of 33.6 MB: the whole problem lives in ONE 16-CTA cluster (shared memory + registers) and the two
// whole-chip grid barriers per iteration of latency_grid2.cu become two DSMEM exchanges.
//
// Layout inside the kernel (everything permuted so that every dot product runs over contiguous, 16-byte aligned data
// and every CTA PRODUCES a contiguous block of each exchanged vector):
//   w_s   [n_u][Q] cell-major box multipliers (Q = 4N entries per cell), then the tail (sum-constraint multipliers);
//         CTA c owns POSITIONS [c R, (c+1) R) of it, thread t the row of G_L whose multiplier lives at position c R + t
//   zh_s  [C][n_u + 1][8]: block c holds the (<= 8) stages CTA c owns, cell-major, plus S[s] = sum_u zhat[s,u]
//   phase A: CTA c owns SC consecutive stages; WARP u handles cell u of those stages: a lane holds float4 chunks of the
//            row vector [w_s segment u | tail] and accumulates SC rows against them (operator rows in shared memory,
//            one vector load serves SC rows); lane sl then owns row (s_sl, u): zhat, z average, into the CTA's own block
//   S:       one block barrier, SC threads sum their stage over the cells in a fixed order
//   phase B: one THREAD per row of G_L (its operator entries in registers, stage-padded like zh_s: 2 float4 per CTA
//            block), box rows read their cell's 8 floats of every block, sum rows read S
//   exchange: every CTA writes what it produced into its OWN shared memory and ships the block to its 15 peers with one
//            bulk asynchronous copy each (cp.async.bulk shared::cta -> shared::cluster, completing bytes on the peer's
//            mbarrier); consumers wait on their own mbarrier, no cluster barrier in the loop.  (The first version stored
//            every entry into every CTA with st.async.b32, like latency_small.cu: 5.3 k four-byte remote stores per CTA
//            and iteration took 8 us per iteration -- the remote-store issue rate, ~2.5 clk each, was the bound.)
// Fixed-iteration solves (the reference's behaviour, main.cu:87); tolerance-mode solves of the same handle run the dense
// kernels.  Same arithmetic as everywhere else: unfused step 1 / step 3 / (w + p_D) / projection, tree-ordered dots.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "gpad_internal.h"
#include "lat_util.cuh"
#include "latency.h"

namespace gpad {
namespace lat {

namespace {

constexpr int kFlatMaxSC = 8;        // stages per CTA in phase A
constexpr int kFlatMaxCHA = 8;       // float4 chunks per lane of a phase-A row vector

__device__ __forceinline__ uint32_t f_smem_addr(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ uint32_t f_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// one 4-byte asynchronous store into every OTHER CTA of the cluster, each crediting that CTA's mbarrier
__device__ __forceinline__ void f_st_async_peers(uint32_t data0, uint32_t bar0, uint32_t stride, float v, int C, int self) {
    const uint32_t bits = __float_as_uint(v);
    for (int r = 0; r < C; ++r, data0 += stride, bar0 += stride)
        if (r != self)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                         ::"r"(data0), "r"(bits), "r"(bar0) : "memory");
}
// one bulk copy of `bytes` (multiple of 16) from this CTA's shared memory to the same offset in CTA `rank`, completing on
// that CTA's mbarrier
__device__ __forceinline__ void f_bulk_to_peer(uint32_t src_cta, uint32_t dst_rank0, uint32_t bar_rank0, uint32_t stride, int rank,
                                               uint32_t bytes) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_rank0 + (uint32_t)rank * stride), "r"(src_cta), "r"(bytes), "r"(bar_rank0 + (uint32_t)rank * stride) : "memory");
}
__device__ __forceinline__ void f_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void f_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void f_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float f_dot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    return acc;
}

// XCHG 0: bulk copies of the CTA's block; XCHG 1: every producing thread stores its entry straight into the peers
// (st.async.b32; a warp's entries are contiguous in the destination)
// big clusters hold 2 * CL float4 of G_L per thread in registers: at most 384 threads there, so that ptxas gets 168 registers
constexpr int flat_max_threads(int cl) { return cl >= 8 ? 384 : kMaxThreads; }

template <int CL, int XCHG>
__global__ void __launch_bounds__(flat_max_threads(CL)) gpad_flat_kernel(const FlatParams p) {
    constexpr int CHB = 2 * CL;                          // float4 fragments of a phase-B row: 8 stage slots per CTA block
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)blockIdx.x;
    const int n_u = p.n_u, N = p.N, m = p.m, Q = p.Q;
    const int box = n_u * Q, SC = p.SC, CHA = p.CHA, lenA = p.lenA, R = p.rows_b;
    const int zblk = (n_u + 1) * 8;                      // floats per CTA block of zh_s
    float* w_s = smem;                                  // [w_len]  permuted multipliers (+ zero padding)
    float* zh_s = w_s + p.w_len;                        // [CL][n_u + 1][8]
    float* th_s = zh_s + CL * zblk;                     // [sched] theta
    float* be_s = th_s + p.sched;                       // [sched] beta
    uint64_t* xbar = reinterpret_cast<uint64_t*>(be_s + p.sched);      // [2] mbarriers
    float* a_s = reinterpret_cast<float*>(xbar + 2);    // [SC * n_u][lenA] this CTA's rows of the phase-A operator

    // ---- phase-A operator rows of this CTA: global -> shared memory, once ----
    const int s0 = rank * SC;                            // first stage of this CTA
    {
        const int rows = SC * n_u;
        const float4* src = reinterpret_cast<const float4*>(p.A_op);
        float4* dst = reinterpret_cast<float4*>(a_s);
        const int l4 = lenA / 4;
        for (int k = tid; k < rows * l4; k += nthr) {
            const int r = k / l4, c4 = k % l4;
            const int s = s0 + r / n_u;
            dst[k] = s < N ? __ldg(src + (size_t)(s * n_u + r % n_u) * l4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    // ---- phase-B row of this thread: the row of G_L whose multiplier lives at position rank R + tid of w_s ----
    const int pos_b = rank * R + tid;
    int row_b = -1;
    if (tid < R) {
        if (pos_b < box) row_b = (pos_b % Q) * n_u + pos_b / Q;
        else if (pos_b < m) row_b = pos_b;
    }
    const bool own_b = row_b >= 0;
    float4 rb[CHB];
    {
        const float4* src = reinterpret_cast<const float4*>(p.B_op + (size_t)max(row_b, 0) * (4 * CHB));
#pragma unroll
        for (int k = 0; k < CHB; ++k) rb[k] = own_b ? __ldg(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int seg_b = (own_b && row_b < box) ? row_b % n_u : n_u;               // cell segment, or S for the sum rows
    const float4* zh4 = reinterpret_cast<const float4*>(zh_s + seg_b * 8);

    // ---- phase-A role: warp u < n_u handles cell u of the CTA's stages; after the transposed butterfly lane 4 sl holds
    //      the total of stage slot sl and owns row (s0 + sl, u) ----
    const bool warp_a = warp < n_u;
    const int u = warp;
    const int sl_me = lane >> 2;                         // stage slot whose row total this lane ends up with
    const int s_me = s0 + sl_me;
    const bool own_a = warp_a && (lane & 3) == 0 && sl_me < SC && s_me < N;
    const int row_a = s_me * n_u + u;
    int voff[kFlatMaxCHA];                               // float offset into w_s of this lane's vector chunks
#pragma unroll
    for (int ch = 0; ch < kFlatMaxCHA; ++ch) {
        const int q = (ch * 32 + lane) * 4;              // position inside the row [segment u | tail]
        voff[ch] = q < Q ? u * Q + q : box + (q - Q);
    }

    // ---- per-row state ----
    float z_r = 0.f, gp_r = 0.f, zh_r = 0.f;
    float yv = 0.f, yp = 0.f, yn = 0.f, pd_r = 0.f, w_r = 0.f;
    const float beta0 = p.beta[0];
    if (own_a) gp_r = p.g_P[row_a];
    if (own_b) {
        yv = p.y0 ? p.y0[row_b] : 0.f;
        yp = p.y_prev0 ? p.y_prev0[row_b] : 0.f;
        yn = yv;
        pd_r = p.p_D[row_b];
        w_r = __fadd_rn(yv, __fmul_rn(beta0, __fsub_rn(yv, yp)));      // step 1 of iteration 0
    }
    for (int i = tid; i < p.w_len; i += nthr) w_s[i] = 0.f;
    for (int i = tid; i < CL * zblk; i += nthr) zh_s[i] = 0.f;
    const bool sched_in_smem = p.max_iter <= p.sched;
    if (sched_in_smem)
        for (int i = tid; i < p.max_iter; i += nthr) { th_s[i] = p.theta[i]; be_s[i] = p.beta[i]; }
    const uint32_t zbar = f_smem_addr(xbar), wbar = f_smem_addr(xbar + 1);
    const uint32_t zbar0 = f_mapa(zbar, 0), wbar0 = f_mapa(wbar, 0);
    const uint32_t cstride = CL > 1 ? f_mapa(zbar, 1) - zbar0 : 0;
    // this CTA's blocks of the two exchanged vectors: local address (source) and rank-0 view of the same offset (destination)
    float* w_mine = w_s + rank * R;
    float* zh_mine = zh_s + rank * zblk;
    const uint32_t w_src = f_smem_addr(w_mine), zh_src = f_smem_addr(zh_mine);
    const uint32_t w_dst0 = f_mapa(w_src, 0), zh_dst0 = f_mapa(zh_src, 0);
    const uint32_t w_bytes = 4u * R, zh_bytes = 4u * zblk;
    uint32_t zpar = 0, wpar = 0;
    if (tid == 0) {
        f_mbar_init(zbar, 1); f_mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    f_cluster_sync();                         // every CTA's vectors are zeroed / barriers initialised before anyone copies into them

    // ships this CTA's block of an exchanged vector to the peers (after every thread's writes to it are visible to the
    // asynchronous proxy) and waits until the peers' blocks have landed here
    auto exchange = [&](uint32_t src, uint32_t dst0, uint32_t bar_local, uint32_t bar0, uint32_t bytes, uint32_t& parity) {
        __syncthreads();
        if (CL > 1) {
            if (tid == 0) f_mbar_expect(bar_local, (uint32_t)(CL - 1) * bytes);
            if (XCHG == 0 && tid < CL && tid != rank) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the block's (barrier-ordered) writes -> async proxy
                f_bulk_to_peer(src, dst0, bar0, cstride, tid, bytes);
            }
            f_mbar_wait(bar_local, parity); parity ^= 1;
        }
    };
    // XCHG 1: rank-0 views of this thread's own entries
    const uint32_t w_ent0 = w_dst0 + 4u * (uint32_t)tid;                        // tid < R
    const uint32_t zh_ent0 = zh_dst0 + 4u * (uint32_t)(u * 8 + sl_me);          // lanes 4 sl (zero for the padding slots)
    const uint32_t s_ent0 = zh_dst0 + 4u * (uint32_t)(n_u * 8 + tid);           // tid < 8

    if (tid < R) {
        w_mine[tid] = own_b ? w_r : 0.f;
        if (XCHG == 1) f_st_async_peers(w_ent0, wbar0, cstride, own_b ? w_r : 0.f, CL, rank);
    }
    exchange(w_src, w_dst0, wbar, wbar0, w_bytes, wpar);

    for (int v = 0; v < p.max_iter; ++v) {
        const float theta = sched_in_smem ? th_s[v] : __ldg(p.theta + v);
        const bool last = v + 1 == p.max_iter;
        const float beta_next = last ? 0.f : (sched_in_smem ? be_s[v + 1] : __ldg(p.beta + v + 1));
        const float one_minus = 1.0f - theta;

        // ---------------- phase A: zhat = M_G w - g_P, z average ----------------
        // straight-line code: loads and products of absent chunks / stage slots are predicated off (a branch per slot
        // would serialise the eight reductions: measured 1.3 us per iteration on a problem that needs 0.4)
        if (warp_a) {
            float acc[kFlatMaxSC];
#pragma unroll
            for (int sl = 0; sl < kFlatMaxSC; ++sl) acc[sl] = 0.f;
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int ch = 0; ch < kFlatMaxCHA; ++ch) {
                const bool have = ch < CHA;
                const float4 x = have ? *reinterpret_cast<const float4*>(w_s + voff[ch]) : zero4;
                const float* a = a_s + (size_t)u * lenA + (ch * 32 + lane) * 4;
#pragma unroll
                for (int sl = 0; sl < kFlatMaxSC; ++sl) {
                    const float4 av = (have && sl < SC) ? *reinterpret_cast<const float4*>(a + (size_t)sl * n_u * lenA) : zero4;
                    acc[sl] = f_dot4(av, x, acc[sl]);
                }
            }
            const float mine = warp_sum_transposed_v<kFlatMaxSC>(acc, lane);      // lane l: total of slot (l >> 2) & 7
            if (own_a) {
                zh_r = mine - gp_r;
                z_r = __fadd_rn(__fmul_rn(one_minus, z_r), __fmul_rn(theta, zh_r));
                zh_mine[u * 8 + sl_me] = zh_r;
            }
            // XCHG 1: all 8 slots of the cell's row travel (the padding slots carry zeros), so that every CTA receives
            // exactly one block's worth of bytes
            if (XCHG == 1 && (lane & 3) == 0) f_st_async_peers(zh_ent0, zbar0, cstride, own_a ? zh_r : 0.f, CL, rank);
        }
        __syncthreads();
        if (tid < 8) {                        // S[s] = sum_u zhat[s,u], fixed order (zero for the padding slots)
            float part[16];
#pragma unroll
            for (int uu = 0; uu < 16; ++uu) part[uu] = uu < n_u ? zh_mine[uu * 8 + tid] : 0.f;
            float S = 0.f;
#pragma unroll
            for (int uu = 0; uu < 16; ++uu) S += part[uu];
            zh_mine[n_u * 8 + tid] = S;
            if (XCHG == 1) f_st_async_peers(s_ent0, zbar0, cstride, S, CL, rank);
        }
        exchange(zh_src, zh_dst0, zbar, zbar0, zh_bytes, zpar);

        // ---------------- phase B: y+ = max(G_L zhat + (w + p_D), 0), momentum ----------------
        {
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int c = 0; c < CL; ++c) {
                const float4* zc = zh4 + c * (zblk / 4);
                d0 = f_dot4(rb[2 * c], zc[0], d0);
                d1 = f_dot4(rb[2 * c + 1], zc[1], d1);
            }
            if (own_b) {
                const float s = (d0 + d1) + (w_r + pd_r);
                yn = 0.5f * (s + fabsf(s));
                if (!last) {            // advance; on the last iteration w_v, y_v stay (they are outputs)
                    const float wn = __fadd_rn(yn, __fmul_rn(beta_next, __fsub_rn(yn, yv)));
                    w_mine[tid] = wn;
                    w_r = wn;
                    yp = yv; yv = yn;
                }
            }
            // XCHG 1: every position of the block travels (positions without a row carry their zero)
            if (XCHG == 1 && !last && tid < R) f_st_async_peers(w_ent0, wbar0, cstride, own_b ? w_r : 0.f, CL, rank);
        }
        if (!last) exchange(w_src, w_dst0, wbar, wbar0, w_bytes, wpar);
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
    int bad = __syncthreads_or(own_b && !isfinite(yn));
    if (tid == 0 && bad) atomicExch(p.nonfinite_flag, 1);
    f_cluster_sync();                          // no CTA exits while a peer may still copy into it; flags are published
    if (rank == 0 && tid == 0) {
        __threadfence();
        if (p.out_iters) *p.out_iters = p.max_iter;
        if (p.out_status) *p.out_status = atomicAdd(p.nonfinite_flag, 0) ? GPAD_STATUS_NONFINITE : GPAD_STATUS_MAX_ITER;
        if (p.out_max_viol) *p.out_max_viol = __int_as_float(0x7fc00000);
        if (p.out_gap) *p.out_gap = __int_as_float(0x7fc00000);
    }
}

size_t flat_smem_bytes(const FlatParams& p) {
    return ((size_t)p.w_len + (size_t)p.C * (p.n_u + 1) * 8 + 2 * (size_t)p.sched + 4 + (size_t)p.SC * p.n_u * p.lenA) * sizeof(float) + 16;
}

template <int CL>
int launch_flat_t(const FlatParams& p, cudaStream_t stream) {
    auto kern = p.xchg == 0 ? gpad_flat_kernel<CL, 0> : gpad_flat_kernel<CL, 1>;
    const size_t smem = flat_smem_bytes(p);
    // function attributes are set when they have to grow (per device): cudaFuncSetAttribute costs tens of microseconds of
    // host time, which a single-QP latency solve would pay on every call
    static size_t smem_set[2][64] = {};
    int dev = 0;
    GPAD_CUDA(cudaGetDevice(&dev));
    size_t& have = smem_set[p.xchg ? 1 : 0][dev & 63];
    if (smem > have) {
        GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (p.C > 8) GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        have = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.C); cfg.blockDim = dim3(p.threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GPAD_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return GPAD_OK;
}

}  // namespace

// plans the cluster for a flat problem; false when it does not fit (n_u > 16 cells, rows beyond 512 threads, shared memory)
bool plan_flat(int n_u, int N, int m, size_t smem_limit, int max_cluster, FlatParams* out) {
    FlatParams p{};
    const int box = 4 * n_u * N;
    if (n_u < 1 || n_u > 16 || m < box || N < 1) return false;
    p.n_u = n_u; p.N = N; p.m = m;
    p.Q = 4 * N;
    const int T4 = round_up(m - box, 4);
    p.lenA = round_up(p.Q + T4, 128);                  // 32 lanes x float4
    p.CHA = p.lenA / 128;
    if (p.CHA > kFlatMaxCHA) return false;
    p.sched = 256;
    for (int C : {1, 2, 4, 8, 16}) {
        if (C > max_cluster) break;
        p.C = C;
        p.SC = (N + C - 1) / C;
        p.rows_b = round_up((m + C - 1) / C, 4);       // positions of w_s per CTA: a 16-byte multiple for the bulk copies
        p.N4 = 8 * C; p.CHB = 2 * C;
        p.w_len = std::max(box + T4, C * p.rows_b) + 128;      // the last chunk of a row vector may read past the tail: zero padding
        p.threads = std::max(32 * n_u, round_up(p.rows_b, 32));
        if (p.SC > kFlatMaxSC || p.threads > flat_max_threads(C)) continue;
        if (flat_smem_bytes(p) > smem_limit) continue;
        *out = p;
        return true;
    }
    return false;
}

// dense sequential operators whose structure is exactly flat -> the kernel's two operator arrays (host); returns the
// largest |entry| of the dense operators that the flat structure cannot represent (0: the problem IS flat)
float build_flat_operators(const FlatParams& p, const float* MG, const float* GL, std::vector<float>& A_op, std::vector<float>& B_op) {
    const int n_u = p.n_u, N = p.N, m = p.m, n = n_u * N, box = n_u * p.Q;
    A_op.assign((size_t)n * p.lenA, 0.f);
    for (int s = 0; s < N; ++s)
        for (int u = 0; u < n_u; ++u) {
            const float* src = MG + (size_t)(s * n_u + u) * m;
            float* dst = &A_op[(size_t)(s * n_u + u) * p.lenA];
            for (int q = 0; q < p.Q; ++q) dst[q] = src[q * n_u + u];              // cell u's box multipliers
            for (int k = box; k < m; ++k) dst[p.Q + (k - box)] = src[k];          // the sum-constraint multipliers
        }
    // row i of G_L, stage s = c SC + sl -> slot c * 8 + sl (the stage padding of zh_s)
    B_op.assign((size_t)m * p.N4, 0.f);
    auto slot = [&](int s) { return (s / p.SC) * 8 + s % p.SC; };
    for (int i = 0; i < m; ++i)
        for (int s = 0; s < N; ++s) B_op[(size_t)i * p.N4 + slot(s)] = GL[(size_t)i * n + s * n_u + (i < box ? i % n_u : 0)];
    float resid = 0.f;
    for (int r = 0; r < n; ++r)
        for (int k = 0; k < box; ++k)
            if (k % n_u != r % n_u) resid = std::max(resid, std::fabs(MG[(size_t)r * m + k]));
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            const float flat = (i >= box || j % n_u == i % n_u) ? B_op[(size_t)i * p.N4 + slot(j / n_u)] : 0.f;
            resid = std::max(resid, std::fabs(GL[(size_t)i * n + j] - flat));
        }
    return resid;
}

int launch_flat(const FlatParams& p, cudaStream_t stream) {
    switch (p.C) {
        case 1: return launch_flat_t<1>(p, stream);
        case 2: return launch_flat_t<2>(p, stream);
        case 4: return launch_flat_t<4>(p, stream);
        case 8: return launch_flat_t<8>(p, stream);
        default: return launch_flat_t<16>(p, stream);
    }
}

}  // namespace lat
}  // namespace gpad
