// batch_tc_p2.cu -- product 2 of the throughput mode, second generation (GPAD_PREC_FP16X3, fixed-iteration solves).
//
//     Y+[B x m] = Zhat_v[B x n] * G_L^T          y_{v+1} = max(acc + (w_v + p_D), 0),  w_v = y_v + beta_v (y_v - y_{v-1})
//
// With fp16 operands the mainloop of this product costs half of what it did, and the first-generation kernel
// (batch_tc.cu) did not get faster: it is bound by its EPILOGUE -- three m-sized arrays in, one out, 2.5 GB per launch at
// the 64K quadrotor batch -- which twelve warps fetched with ordinary loads after a shared-memory transpose, stalling
// on HBM latency between the accumulator and the store (3.4 TB/s).  Here the epilogue operands are streamed by TMA:
//   warp 0      mainloop producer: zhat hi / lo [128 x 32] and G_L hi / lo [bn x 32] fp16 k-blocks (SWIZZLE_64B ring)
//   warp 1      TMEM allocator + MMA issuer (kind::f16, hi*lo + lo*hi + hi*hi into one fp32 accumulator, two stages)
//   warp 2      epilogue-operand producer: for every 32-column block of the tile three boxes [128 rows x 32 fp32] of
//               y_v, y_{v-1}, p_D (SWIZZLE_128B) into an E ring -- it runs a whole ring ahead of the epilogue warps, so
//               the operands of a tile arrive while its MMAs are still running
//   warp 3      store issuer: cp.async.bulk.tensor store of a finished block, frees the E slot once the store has read it
//   warps 4..11 two epilogue groups of 128 threads (thread = batch row = TMEM lane): tcgen05.ld of the row's 32
//               accumulator columns, the three operand rows from shared memory (conflict-free under the 128-byte
//               swizzle, no transpose), y_{v+1} written IN PLACE over the p_D row, fence, arrive
// Row maxima of y_{v+1} (the next product 1 scales its fp16 operand rows with them) are a per-thread running maximum:
// one atomicMax per row and tile.  Columns beyond m are clipped by the tensor maps (zero fill on load, dropped on
// store); rows beyond the batch compute on zeros.
#include <cuda.h>

#include <algorithm>

#include "batch_common.cuh"
#include "batch_tc.h"
#include "gpad_internal.h"
#include "tc_ptx.cuh"

namespace gpad {
namespace tc {

namespace {

constexpr int kP2Groups = 2;                 // epilogue groups of 4 warps
constexpr int kP2FirstEpi = 4;
constexpr int kP2Threads = 32 * (kP2FirstEpi + 4 * kP2Groups);
constexpr int kP2BK = 16;                    // operand tile rows are 64 bytes: 32 fp16
constexpr uint32_t kEOp = kBM * 32 * 4;      // one epilogue operand block: 128 rows x 32 fp32 = 16 KB
constexpr uint32_t kEStage = 3 * kEOp;

struct P2Sched {
    int unit, step, total, n_tiles;
    __device__ P2Sched(int m_tiles, int n_tiles_) : unit(blockIdx.x), step(gridDim.x), total(m_tiles * n_tiles_), n_tiles(n_tiles_) {}
    __device__ bool valid() const { return unit < total; }
    __device__ void next() { unit += step; }
    __device__ int m_tile() const { return unit / n_tiles; }      // operator tile fastest: the tiles of one batch tile run
    __device__ int n_tile() const { return unit % n_tiles; }      // on neighbouring CTAs at the same time (DRAM locality)
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(kP2Threads, 1)
tc_p2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
             const __grid_constant__ CUtensorMap tmYcur, const __grid_constant__ CUtensorMap tmYprev,
             const __grid_constant__ CUtensorMap tmPd, const __grid_constant__ CUtensorMap tmYnext,
             int num_k_blocks, int m_tiles, int n_tiles, int bn, int stages, int e_stages, const BatchKernelArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int BK = kP2BK;
    const uint32_t a_bytes = kBM * BK * 4, b_bytes = (uint32_t)bn * BK * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;           // a multiple of 1024: bn is a multiple of 32
    uint8_t* e_ring = smem + (size_t)stages * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(e_ring + (size_t)e_stages * kEStage);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = full_bar + stages;
    uint64_t* e_full = empty_bar + stages;          // operands of a block have landed
    uint64_t* e_done = e_full + e_stages;           // the block's y_{v+1} is written (128 arrivals)
    uint64_t* e_empty = e_done + e_stages;          // the store has read the slot
    uint64_t* tfull_bar = e_empty + e_stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = bn / 32;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
        tma_prefetch_desc(&tmYcur); tma_prefetch_desc(&tmYprev); tma_prefetch_desc(&tmPd); tma_prefetch_desc(&tmYnext);
        for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
        for (int s = 0; s < e_stages; ++s) {
            mbar_init(smem_u32(e_full + s), 1); mbar_init(smem_u32(e_done + s), 128); mbar_init(smem_u32(e_empty + s), 1);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(tfull_bar + s), 1); mbar_init(smem_u32(tempty_bar + s), 4 * kP2Groups); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    grid_dependency_wait();
    grid_launch_dependents();

    if (warp == 0) {
        // ============================ mainloop producer ============================
        int stage = 0; uint32_t phase = 0;
        for (P2Sched ts(m_tiles, n_tiles); ts.valid(); ts.next()) {
            const int row_a = ts.m_tile() * kBM, row_b = ts.n_tile() * bn;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                const uint32_t fb = smem_u32(full_bar + stage);
                const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
                if (elect_one()) {
                    mbar_expect_tx(fb, stage_bytes);
                    tma_load_2d(base, &tmA_hi, kb * 2 * BK, row_a, fb);
                    tma_load_2d(base + a_bytes, &tmA_lo, kb * 2 * BK, row_a, fb);
                    tma_load_2d(base + 2 * a_bytes, &tmB_hi, kb * 2 * BK, row_b, fb);
                    tma_load_2d(base + 2 * a_bytes + b_bytes, &tmB_lo, kb * 2 * BK, row_b, fb);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        const uint32_t idesc = make_idesc_f16(bn);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (P2Sched ts(m_tiles, n_tiles); ts.valid(); ts.next()) {
            mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32(full_bar + stage), phase);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t a_hi = make_smem_desc<BK>(base + ks * 32);
                        const uint64_t a_lo = make_smem_desc<BK>(base + a_bytes + ks * 32);
                        const uint64_t b_hi = make_smem_desc<BK>(base + 2 * a_bytes + ks * 32);
                        const uint64_t b_lo = make_smem_desc<BK>(base + 2 * a_bytes + b_bytes + ks * 32);
                        umma_f16(d_tmem, a_hi, b_lo, idesc, (kb | ks) != 0 ? 1u : 0u);
                        umma_f16(d_tmem, a_lo, b_hi, idesc, 1u);
                        umma_f16(d_tmem, a_hi, b_hi, idesc, 1u);
                    }
                    umma_commit(smem_u32(empty_bar + stage));
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(smem_u32(tfull_bar + acc));
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp == 2) {
        // ============================ epilogue-operand producer ============================
        int es = 0; uint32_t eph = 0;
        for (P2Sched ts(m_tiles, n_tiles); ts.valid(); ts.next()) {
            const int row = ts.m_tile() * kBM, col0 = ts.n_tile() * bn;
            for (int blk = 0; blk < nblk; ++blk) {
                mbar_wait(smem_u32(e_empty + es), eph ^ 1);
                if (elect_one()) {
                    const uint32_t fb = smem_u32(e_full + es);
                    const uint32_t base = smem_u32(e_ring + (size_t)es * kEStage);
                    mbar_expect_tx(fb, kEStage);
                    tma_load_2d(base, &tmYcur, col0 + blk * 32, row, fb);
                    tma_load_2d(base + kEOp, &tmYprev, col0 + blk * 32, row, fb);
                    tma_load_2d(base + 2 * kEOp, &tmPd, col0 + blk * 32, row, fb);
                }
                __syncwarp();
                if (++es == e_stages) { es = 0; eph ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ============================ store issuer ============================
        int es = 0; uint32_t eph = 0;
        for (P2Sched ts(m_tiles, n_tiles); ts.valid(); ts.next()) {
            const int row = ts.m_tile() * kBM, col0 = ts.n_tile() * bn;
            for (int blk = 0; blk < nblk; ++blk) {
                mbar_wait(smem_u32(e_done + es), eph);
                if (elect_one()) {
                    tma_store_2d(&tmYnext, smem_u32(e_ring + (size_t)es * kEStage + 2 * kEOp), col0 + blk * 32, row);
                    tma_store_commit();
                    tma_store_wait_read();                          // the slot may be refilled
                    mbar_arrive(smem_u32(e_empty + es));
                }
                __syncwarp();
                if (++es == e_stages) { es = 0; eph ^= 1; }
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else {
        // ============================ epilogue groups ============================
        const int g = (warp - kP2FirstEpi) >> 2;
        const int q = warp & 3, row = q * 32 + lane;                // TMEM lane = batch row of the tile
        const uint32_t sw = (uint32_t)(row & 7);
        const float beta = args.it.beta;
        int es = 0; uint32_t eph = 0;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t blk_count = 0;                                     // blocks alternate between the groups
        for (P2Sched ts(m_tiles, n_tiles); ts.valid(); ts.next()) {
            const int row_g = ts.m_tile() * kBM + row, col0 = ts.n_tile() * bn;
            const float row_inv = __ldg(args.a_rowinv + row_g);
            uint32_t mx = 0;
            mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
            tc_fence_after();
            for (int blk = 0; blk < nblk; ++blk, ++blk_count) {
                if ((int)(blk_count % kP2Groups) == g) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride + blk * 32), v);
                    const float4* cinv = reinterpret_cast<const float4*>(args.b_colinv + col0 + blk * 32);
                    mbar_wait(smem_u32(e_full + es), eph);
                    uint8_t* yc = e_ring + (size_t)es * kEStage + row * 128;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const uint32_t off = (ch ^ sw) << 4;        // SWIZZLE_128B: chunk c of row r at r * 128 + ((c ^ (r & 7)) << 4)
                        const float4 a = *reinterpret_cast<const float4*>(yc + off);
                        const float4 b = *reinterpret_cast<const float4*>(yc + kEOp + off);
                        const float4 c = *reinterpret_cast<const float4*>(yc + 2 * kEOp + off);
                        const float4 ci = __ldg(cinv + ch);
                        float4 o;
                        {
                            const float s = __uint_as_float(v[ch * 4 + 0]) * row_inv * ci.x + (momentum(a.x, b.x, beta) + c.x);
                            o.x = 0.5f * (s + fabsf(s));
                        }
                        {
                            const float s = __uint_as_float(v[ch * 4 + 1]) * row_inv * ci.y + (momentum(a.y, b.y, beta) + c.y);
                            o.y = 0.5f * (s + fabsf(s));
                        }
                        {
                            const float s = __uint_as_float(v[ch * 4 + 2]) * row_inv * ci.z + (momentum(a.z, b.z, beta) + c.z);
                            o.z = 0.5f * (s + fabsf(s));
                        }
                        {
                            const float s = __uint_as_float(v[ch * 4 + 3]) * row_inv * ci.w + (momentum(a.w, b.w, beta) + c.w);
                            o.w = 0.5f * (s + fabsf(s));
                        }
                        mx = max(mx, max(max(__float_as_uint(o.x), __float_as_uint(o.y)), max(__float_as_uint(o.z), __float_as_uint(o.w))));
                        *reinterpret_cast<float4*>(yc + 2 * kEOp + off) = o;
                    }
                    fence_proxy_async_smem();                       // generic-proxy writes -> visible to the TMA store
                    mbar_arrive(smem_u32(e_done + es));
                }
                if (++es == e_stages) { es = 0; eph ^= 1; }
            }
            // columns beyond m were computed on zero-filled operands (and are dropped by the store): they are >= 0 only
            // through p_D = 0, acc = 0, so they cannot raise the maximum
            if (mx != 0u && row_g < args.B) atomicMax(args.next_rowmax + row_g, mx);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty_bar + acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

size_t p2_smem_bytes(int bn, int stages, int e_stages) {
    return 1024 + (size_t)stages * (2 * kBM + 2 * bn) * kP2BK * 4 + (size_t)e_stages * kEStage + (2 * stages + 3 * e_stages + 4) * 8 + 16;
}

}  // namespace

// tiles of a multiple of 32 columns (the epilogue works in whole 32-column blocks), at most 256
void plan_tiles_p2(int ncols, int* bn, int* n_tiles) {
    const int nt = (ncols + 255) / 256;
    int b = ((ncols + nt - 1) / nt + 31) / 32 * 32;
    *bn = std::max(32, std::min(256, b));
    *n_tiles = (ncols + *bn - 1) / *bn;
}

int plan_rings_p2(int bn, size_t smem_limit, int* stages, int* e_stages, int max_stages) {
    int e = 2, s = 6;
    while (s > 2 && p2_smem_bytes(bn, s, e) > smem_limit) --s;
    if (p2_smem_bytes(bn, s, e) > smem_limit) return GPAD_ERR_UNSUPPORTED;
    if (s > 4) s = 4;
    if (max_stages >= 2 && s > max_stages) s = max_stages;
    while (e < 4 && p2_smem_bytes(bn, s, e + 1) <= smem_limit) ++e;      // spare shared memory deepens the HBM-facing ring
    *stages = s; *e_stages = e;
    return GPAD_OK;
}

// y_{v+1} = max(Zhat G_L^T + (w_v + p_D), 0); cur / prev / next index the rotating y buffers (GemmDesc::tmEy)
int launch_p2(const GemmDesc& g, const BatchKernelArgs& args, int cur, int prev, int next, int num_sms, cudaStream_t s) {
    if (g.bn % 32 || g.bn > 256) { set_error("tcgen05 product 2: tile width %d is not a multiple of 32 <= 256", g.bn); return GPAD_ERR_UNSUPPORTED; }
    const size_t smem = p2_smem_bytes(g.bn, g.stages, g.e_stages);
    GPAD_CUDA(cudaFuncSetAttribute(tc_p2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int units = g.m_tiles * g.n_tiles;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(std::min(units, num_sms)); lc.blockDim = dim3(kP2Threads); lc.dynamicSmemBytes = smem; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = g.pdl ? 1 : 0;
    GPAD_CUDA(cudaLaunchKernelEx(&lc, tc_p2_kernel, g.tmA_hi, g.tmA_lo, g.tmB_hi, g.tmB_lo, g.tmEy[cur], g.tmEy[prev], g.tmEpd, g.tmEy[next],
                                 g.k_pad / (2 * kP2BK), g.m_tiles, g.n_tiles, g.bn, g.stages, g.e_stages, args));
    return GPAD_OK;
}

}  // namespace tc
}  // namespace gpad
