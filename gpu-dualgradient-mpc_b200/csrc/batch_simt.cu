// batch_simt.cu -- throughput mode on CUDA cores (GPAD_PREC_FP32) plus the batch bookkeeping
// kernels both precisions share (init / pad / decide / gather outputs).
//
// Each GPAD iteration over a batch of B QPs that share M_G and G_L is two GEMMs with the batch
// on the M dimension:
//     Zhat[B x n] = W[B x m]    * M_G^T[m x n]   (+ epilogue1: -g_P, z average)
//     Y+  [B x m] = Zhat[B x n] * G_L^T[n x m]   (+ epilogue2: +w+p_D, projection, momentum)
// Both operands are K-contiguous (instance-major state, sequential-layout operators), all
// buffers are zero padded to the tile sizes, so the mainloop has no bounds checks.
// Tile 128x128x16, 256 threads, 8x8 outputs per thread, double-buffered shared memory.
#include <algorithm>

#include "batch_common.cuh"
#include "gpad_internal.h"

namespace gpad {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

// C[b][i] = sum_k A[b*lda + k] * Bop[i*ldb + k], K multiple of BK, rows padded to BM / BN.
// PHASE 1: A = w_v built on the fly from y_v (A) and y_{v-1} (A2): step 1 fused into the operand load.
template <int PHASE>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const float* __restrict__ A, const float* __restrict__ A2, int lda,
                                                        const float* __restrict__ Bop, int ldb, int K,
                                                        const BatchKernelArgs args) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 thread grid, 8x8 outputs each
    if (args.dual && args.dual_count && *args.dual_count == 0) return;     // nobody waits for the dual-gap evaluation
    if (args.tile_count && (int)blockIdx.y >= *args.tile_count) return;     // tolerance mode: retired batch tiles
    const int row0 = (args.tile_list ? args.tile_list[blockIdx.y] : (int)blockIdx.y) * BM, col0 = blockIdx.x * BN;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // global -> register staging: 2 float4 per operand per thread
    const int lr0 = tid >> 2, lk = (tid & 3) * 4;    // rows lr0 and lr0+64, k offset lk
    const float* Ag = A + (size_t)(row0 + lr0) * lda + lk;
    const float* Ag2 = PHASE == 1 ? A2 + (size_t)(row0 + lr0) * lda + lk : nullptr;
    const float* Bg = Bop + (size_t)(col0 + lr0) * ldb + lk;
    float4 ra0, ra1, rb0, rb1;
    const float beta = args.it.beta;
    auto load_g = [&](int k0) {
        ra0 = *reinterpret_cast<const float4*>(Ag + k0);
        ra1 = *reinterpret_cast<const float4*>(Ag + (size_t)64 * lda + k0);
        if (PHASE == 1) {
            const float4 p0 = *reinterpret_cast<const float4*>(Ag2 + k0);
            const float4 p1 = *reinterpret_cast<const float4*>(Ag2 + (size_t)64 * lda + k0);
            ra0 = make_float4(momentum(ra0.x, p0.x, beta), momentum(ra0.y, p0.y, beta), momentum(ra0.z, p0.z, beta), momentum(ra0.w, p0.w, beta));
            ra1 = make_float4(momentum(ra1.x, p1.x, beta), momentum(ra1.y, p1.y, beta), momentum(ra1.z, p1.z, beta), momentum(ra1.w, p1.w, beta));
        }
        rb0 = __ldg(reinterpret_cast<const float4*>(Bg + k0));
        rb1 = __ldg(reinterpret_cast<const float4*>(Bg + (size_t)64 * ldb + k0));
    };
    auto store_s = [&](int buf) {
        As[buf][lk + 0][lr0] = ra0.x; As[buf][lk + 1][lr0] = ra0.y; As[buf][lk + 2][lr0] = ra0.z; As[buf][lk + 3][lr0] = ra0.w;
        As[buf][lk + 0][lr0 + 64] = ra1.x; As[buf][lk + 1][lr0 + 64] = ra1.y; As[buf][lk + 2][lr0 + 64] = ra1.z; As[buf][lk + 3][lr0 + 64] = ra1.w;
        Bs[buf][lk + 0][lr0] = rb0.x; Bs[buf][lk + 1][lr0] = rb0.y; Bs[buf][lk + 2][lr0] = rb0.z; Bs[buf][lk + 3][lr0] = rb0.w;
        Bs[buf][lk + 0][lr0 + 64] = rb1.x; Bs[buf][lk + 1][lr0 + 64] = rb1.y; Bs[buf][lk + 2][lr0 + 64] = rb1.z; Bs[buf][lk + 3][lr0 + 64] = rb1.w;
    };

    load_g(0);
    store_s(0);
    __syncthreads();
    const int nk = K / BK;
    for (int kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load_g((kb + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            // thread owns rows ty*4..+3 and 64+ty*4..+3, cols tx*4..+3 and 64+tx*4..+3
            const float4 a_lo = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a_hi = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b_lo = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b_hi = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
            const float bv[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < nk) {
            store_s(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- fused epilogue ----
    const int ncols = PHASE == 1 ? args.n : args.m;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int b = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (b >= args.B) continue;
        if (args.done && args.done[b]) continue;
        float f_zhat = 0.f;
        Red2 red;
        bool any = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (c >= ncols) continue;
            any = true;
            if (PHASE == 1) epilogue1<false>(args, b, c, acc[i][j], f_zhat);
            else epilogue2(args, b, c, acc[i][j], red);
        }
        if (args.it.check && any) {
            if (PHASE == 1) { if (args.f) atomicAdd(args.red + (size_t)b * kRedStride + 5, f_zhat); }
            else flush_red2(args, b, red);
        }
    }
}

// dst[Bp][ld] <- src[B][len] (zero padding), src may be null (all zeros)
__global__ void pad_rows_kernel(float* __restrict__ dst, int ld, int rows_total, const float* __restrict__ src,
                                int len, int B) {
    const size_t total = (size_t)rows_total * ld;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / ld), i = (int)(idx % ld);
        dst[idx] = (src && b < B && i < len) ? src[(size_t)b * len + i] : 0.f;
    }
}

// dst[B][len] <- src[Bp][ld]
__global__ void unpad_rows_kernel(float* __restrict__ dst, int len, int B, const float* __restrict__ src, int ld) {
    const size_t total = (size_t)B * len;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / len), i = (int)(idx % len);
        dst[idx] = src[(size_t)b * ld + i];
    }
}

// zeroed z / zhat / sbar (y_0, y_{-1} were written into yb[0], yb[2] by the padding copies)
__global__ void batch_init_kernel(int Bp, int np, int mp, float* __restrict__ z, float* __restrict__ zhat,
                                  float* __restrict__ sbar) {
    const size_t tm = (size_t)Bp * mp, tn = (size_t)Bp * np;
    if (sbar)
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < tm; idx += (size_t)gridDim.x * blockDim.x) sbar[idx] = 0.f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < tn; idx += (size_t)gridDim.x * blockDim.x) {
        z[idx] = 0.f;
        zhat[idx] = 0.f;
    }
}

__global__ void batch_reset_term_kernel(int Bp, float* __restrict__ red, int* __restrict__ done, int* __restrict__ iters,
                                        int* __restrict__ status, float* __restrict__ max_viol, float* __restrict__ gap,
                                        int* __restrict__ active, int* __restrict__ need, int B, int max_iter,
                                        int* __restrict__ tile_list, int* __restrict__ tile_count,
                                        unsigned long long* __restrict__ stat) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) { active[0] = B; active[1] = 0; *tile_count = (B + 127) / 128; stat[0] = 0; stat[1] = 0; }
    if (b < (B + 127) / 128) tile_list[b] = b;        // every batch tile runs until the first decision
    if (b >= Bp) return;
    float* r = red + (size_t)b * kRedStride;
    r[0] = -INFINITY; r[1] = -INFINITY; r[2] = INFINITY; r[3] = 0.f; r[4] = 0.f; r[5] = 0.f; r[6] = 0.f; r[7] = 0.f;
    done[b] = 0;
    need[b] = 0;
    iters[b] = max_iter;
    status[b] = GPAD_STATUS_MAX_ITER;
    max_viol[b] = __int_as_float(0x7fc00000);
    gap[b] = __int_as_float(0x7fc00000);
}

// termination decision per instance after a check iteration (same tests as oracle_solve).
// Instances whose zhat is feasible but whose w has a negative entry take the dual-gap branch when f is given: they
// are flagged in need[] (V(zhat) and the violation parked in red[7] / red[6]) and decided by
// batch_decide_dual_kernel after the two extra operator products.
__global__ void batch_decide_kernel(int B, int iter_done, float L, float eps_g, float eps_V, int have_f,
                                    float* __restrict__ red, int* __restrict__ done, int* __restrict__ iters,
                                    int* __restrict__ status, float* __restrict__ max_viol, float* __restrict__ gap,
                                    int* __restrict__ active, int* __restrict__ need) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || done[b]) return;
    float* r = red + (size_t)b * kRedStride;
    const float viol_z = L * r[0], viol_zhat = L * r[1];
    int st = -1;
    float mv = viol_z, gp = gap[b];
    bool dual = false;
    float V = 0.f;
    if (r[6] > 0.f) st = GPAD_STATUS_NONFINITE;
    else if (viol_z <= eps_g) st = GPAD_STATUS_CONVERGED_Z;
    else if (viol_zhat <= eps_g) {
        V = 0.5f * (r[5] - L * r[4]);
        if (r[2] >= 0.f) {
            gp = -L * r[3];
            if (gp <= eps_V || (have_f && gp <= V * eps_V / (1.0f + eps_V))) { st = GPAD_STATUS_CONVERGED_ZHAT; mv = viol_zhat; }
        } else if (have_f) {
            dual = true;
        }
    }
    max_viol[b] = mv;
    gap[b] = gp;
    if (st >= 0) {
        status[b] = st;
        iters[b] = iter_done;
        done[b] = 1;
        atomicSub(active, 1);
    }
    r[0] = -INFINITY; r[1] = -INFINITY; r[2] = INFINITY; r[3] = 0.f; r[4] = 0.f; r[5] = 0.f; r[6] = 0.f;
    if (dual) {
        need[b] = 1;
        r[6] = viol_zhat;
        r[7] = V;
        atomicAdd(active + 1, 1);
    }
}

// second half of the dual-gap branch: red[5] = f'z_y, red[3] = y'(G_L z_y), red[4] = y'p_D with y = y_{v+1}
//   Phi(y) = f'z_y / 2 + (L/2) y'(G_L z_y) + L y'p_D;   stop when V(zhat) - Phi <= eps_V max(Phi, 1)
__global__ void batch_decide_dual_kernel(int B, int iter_done, float L, float eps_V, float* __restrict__ red,
                                         int* __restrict__ done, int* __restrict__ iters, int* __restrict__ status,
                                         float* __restrict__ max_viol, float* __restrict__ gap, int* __restrict__ active,
                                         int* __restrict__ need) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) active[1] = 0;
    if (b >= B || !need[b]) return;
    float* r = red + (size_t)b * kRedStride;
    const float Phi = 0.5f * r[5] + 0.5f * L * r[3] + L * r[4];
    const float gapv = r[7] - Phi;
    gap[b] = gapv;
    if (gapv <= eps_V * fmaxf(Phi, 1.0f)) {
        status[b] = GPAD_STATUS_CONVERGED_DUAL;
        iters[b] = iter_done;
        max_viol[b] = r[6];
        done[b] = 1;
        atomicSub(active, 1);
    }
    need[b] = 0;
    r[3] = 0.f; r[4] = 0.f; r[5] = 0.f; r[6] = 0.f; r[7] = 0.f;
}

// ---- tile retirement (tolerance mode): after every decision the 128-row batch tiles that still hold a running
// instance are listed densely, in ascending order; the GEMM kernels then schedule only those ----
__global__ void batch_tile_flags_kernel(int B, const int* __restrict__ done, int* __restrict__ flags) {
    const int b = blockIdx.x * 128 + threadIdx.x;
    const int run = __syncthreads_or(b < B && !done[b]);
    if (threadIdx.x == 0) flags[blockIdx.x] = run;
}

// one block: stat[0] += (tiles that ran since the last decision) x 128 x iterations, then the new list
__global__ void __launch_bounds__(1024)
batch_tile_list_kernel(int m_tiles, const int* __restrict__ flags, int* __restrict__ list, int* __restrict__ count,
                       unsigned long long* __restrict__ stat, int iterations, int rebuild) {
    __shared__ int warp_sum[32];
    __shared__ int base;
    if (threadIdx.x == 0) {
        stat[0] += (unsigned long long)*count * 128ull * (unsigned long long)iterations;
        base = 0;
    }
    __syncthreads();
    if (!rebuild) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t0 = 0; t0 < m_tiles; t0 += 1024) {
        const int t = t0 + threadIdx.x;
        const int f = (t < m_tiles && flags[t]) ? 1 : 0;
        int incl = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int ws = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += v; }
            warp_sum[lane] = ws;                   // inclusive over warps
        }
        __syncthreads();
        const int before = base + (warp ? warp_sum[warp - 1] : 0) + incl - f;
        if (f) list[before] = t;
        __syncthreads();
        if (threadIdx.x == 0) base += warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

// stat[1] = sum over instances of the iterations they needed
__global__ void batch_iter_sum_kernel(int B, const int* __restrict__ iters, unsigned long long* __restrict__ stat) {
    unsigned long long s = 0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) s += (unsigned long long)iters[b];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(stat + 1, s);
}

// outputs of instance b after I_b iterations: y_I in yb[I % 3], y_{I-1} in yb[(I-1) % 3], and
// w_{I-1} = y_{I-1} + beta_{I-1} (y_{I-1} - y_{I-2}) with y_{I-2} in yb[(I+1) % 3]
__global__ void unpad_y_kernel(float* __restrict__ dst_next, float* __restrict__ dst_cur, float* __restrict__ dst_w, int m, int B,
                               const float* __restrict__ yb0, const float* __restrict__ yb1, const float* __restrict__ yb2,
                               int mp, const int* __restrict__ iters, const float* __restrict__ beta) {
    const size_t total = (size_t)B * m;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / m), i = (int)(idx % m);
        const int I = iters[b];
        const float* buf[3] = {yb0, yb1, yb2};
        const size_t o = (size_t)b * mp + i;
        const float yI = buf[I % 3][o], yIm1 = buf[(I + 2) % 3][o], yIm2 = buf[(I + 1) % 3][o];
        if (dst_next) dst_next[idx] = yI;
        if (dst_cur) dst_cur[idx] = yIm1;
        if (dst_w) dst_w[idx] = I > 0 ? momentum(yIm1, yIm2, beta[I - 1]) : 0.f;
    }
}

// fixed-iteration mode: MAX_ITER unless an iterate went non-finite
__global__ void batch_finite_kernel(int B, int mp, int m, const float* __restrict__ y_next, int* __restrict__ status) {
    const int b = blockIdx.x;
    if (b >= B) return;
    int bad = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x)
        if (!isfinite(y_next[(size_t)b * mp + i])) bad = 1;
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0 && bad) status[b] = GPAD_STATUS_NONFINITE;
}

inline int grid_for(size_t total) {
    size_t g = (total + 255) / 256;
    return (int)(g > 148 * 16 ? 148 * 16 : (g ? g : 1));
}

}  // namespace

int launch_pad_rows(float* dst, int ld, int rows_total, const float* src, int len, int B, cudaStream_t s) {
    pad_rows_kernel<<<grid_for((size_t)rows_total * ld), 256, 0, s>>>(dst, ld, rows_total, src, len, B);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_unpad_rows(float* dst, int len, int B, const float* src, int ld, cudaStream_t s) {
    unpad_rows_kernel<<<grid_for((size_t)B * len), 256, 0, s>>>(dst, len, B, src, ld);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_batch_init(const BatchState& st, bool checking, cudaStream_t s) {
    batch_init_kernel<<<grid_for((size_t)st.Bp * st.mp), 256, 0, s>>>(st.Bp, st.np, st.mp, st.z, st.zhat, checking ? st.sbar : nullptr);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_batch_reset_term(const BatchState& st, int max_iter, cudaStream_t s) {
    batch_reset_term_kernel<<<(st.Bp + 255) / 256, 256, 0, s>>>(st.Bp, st.red, st.done, st.iters, st.status, st.max_viol,
                                                              st.gap, st.active_count, st.need, st.B, max_iter, st.tile_list,
                                                              st.tile_count, st.stat);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

// accounts `iterations` for the tiles that ran since the last decision and (rebuild) lists the tiles still running
int launch_batch_tiles(const BatchState& st, int iterations, bool rebuild, cudaStream_t s) {
    const int m_tiles = (st.B + 127) / 128;
    if (rebuild) batch_tile_flags_kernel<<<m_tiles, 128, 0, s>>>(st.B, st.done, st.tile_flags);
    batch_tile_list_kernel<<<1, 1024, 0, s>>>(m_tiles, st.tile_flags, st.tile_list, st.tile_count, st.stat, iterations, rebuild ? 1 : 0);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_batch_iter_sum(const BatchState& st, cudaStream_t s) {
    batch_iter_sum_kernel<<<std::min(148, (st.B + 255) / 256), 256, 0, s>>>(st.B, st.iters, st.stat);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_batch_decide(const BatchState& st, int iter_done, float L, float eps_g, float eps_V, bool have_f, cudaStream_t s) {
    batch_decide_kernel<<<(st.B + 255) / 256, 256, 0, s>>>(st.B, iter_done, L, eps_g, eps_V, have_f ? 1 : 0, st.red, st.done,
                                                          st.iters, st.status, st.max_viol, st.gap, st.active_count, st.need);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_batch_decide_dual(const BatchState& st, int iter_done, float L, float eps_V, cudaStream_t s) {
    batch_decide_dual_kernel<<<(st.B + 255) / 256, 256, 0, s>>>(st.B, iter_done, L, eps_V, st.red, st.done, st.iters, st.status,
                                                               st.max_viol, st.gap, st.active_count, st.need);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_unpad_y(float* dst_next, float* dst_cur, float* dst_w, int m, int B, const float* yb0, const float* yb1,
                   const float* yb2, int mp, const int* iters, const float* beta_dev, cudaStream_t s) {
    unpad_y_kernel<<<grid_for((size_t)B * m), 256, 0, s>>>(dst_next, dst_cur, dst_w, m, B, yb0, yb1, yb2, mp, iters, beta_dev);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_batch_finite(const BatchState& st, const float* y_next, cudaStream_t s) {
    batch_finite_kernel<<<st.B, 128, 0, s>>>(st.B, st.mp, st.m, y_next, st.status);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

// one product of a GPAD iteration on CUDA cores (fused GEMM + epilogue)
int launch_simt_product(int phase, const Operators& op, const BatchKernelArgs& args, int Bp, cudaStream_t s) {
    if (phase == 1) {
        dim3 g1((args.n + BN - 1) / BN, Bp / BM);
        simt_gemm_kernel<1><<<g1, 256, 0, s>>>(args.y_cur, args.y_prev, args.mp, op.M_G, args.mp, args.mp, args);
    } else {
        dim3 g2((args.m + BN - 1) / BN, Bp / BM);
        simt_gemm_kernel<2><<<g2, 256, 0, s>>>(args.dual ? args.zy : args.zhat, nullptr, args.np, op.G_L, args.np, args.np, args);
    }
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

}  // namespace gpad
