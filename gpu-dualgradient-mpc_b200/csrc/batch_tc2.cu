// batch_tc2.cu -- the throughput-mode GEMMs on CTA PAIRS (tcgen05 cta_group::2).
//
// Same math, warp roles and epilogues as batch_tc.cu, but two CTAs of a 2-CTA cluster (one TPC)
// compute one 256 x bn tile: each CTA stages ITS OWN 128 batch rows of A and only HALF of the
// operator tile B (bn/2 rows), the leader's single MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256) and both tensor cores read the peer's B half over the pair
// link.  That halves the L2 -> shared-memory operator traffic per SM, which is what limited the
// cta_group::1 kernel (both GEMMs sat at ~56 % tensor-pipe active with DRAM and L2 unsaturated,
// profiles/r1_ncu_tc_gemm_full.csv).
//
// Synchronisation (mbarriers live at identical offsets in both CTAs):
//   full[s]   leader's: its producer arrives with expect_tx for BOTH CTAs' bytes; the peer's TMA
//             loads are 2-SM loads whose completion bytes are credited to the leader's barrier
//   fullA[s]  product 1 only, per CTA: the two A tiles (y_v, y_{v-1}) land on the CTA's own barrier so
//             its transform warps can rewrite them as hi/lo of w; they then arrive on
//   ready[s]  leader's (count = transform warps of both CTAs; the peer arrives remotely)
//   empty[s]  per CTA, released by the leader's multicast tcgen05.commit
//   tfull[a]  per CTA (multicast commit), tempty[a] leader's (epilogue warps of both CTAs)
#include <cuda.h>

#include "batch_common.cuh"
#include "batch_tc.h"
#include "gpad_internal.h"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace gpad {
namespace tc {

namespace {

struct PairSched {
    int tile, step, total, n_tiles;
    __device__ PairSched(int total_, int n_tiles_) : tile(blockIdx.x >> 1), step(gridDim.x >> 1), total(total_), n_tiles(n_tiles_) {}
    __device__ bool valid() const { return tile < total; }
    __device__ void next() { tile += step; }
    __device__ int m_tile() const { return tile / n_tiles; }     // 256-row pair tile
    __device__ int n_tile() const { return tile % n_tiles; }
};

template <int PHASE, int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                int num_k_blocks, int m_tiles, int n_tiles, int bn, int stages,
                const BatchKernelArgs args, float* __restrict__ Cdbg, int ldc, int ncols_valid) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int bh = bn >> 1;                                    // operator rows staged by this CTA
    const uint32_t a_bytes = kBM * BK * 4, bh_bytes = (uint32_t)bh * BK * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * bh_bytes;   // per CTA
    constexpr bool kXform = PHASE == 1;
    constexpr int kEpi = kXform ? kWorkWarps - kXformWarps : kWorkWarps;
    constexpr int kFirstEpiWarp = 2 + (kXform ? kXformWarps : 0);
    float* epi_buf = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + kWorkWarps * kEpiBufFloats);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + stages;
    uint64_t* ready_bar = bars + 2 * stages;
    uint64_t* fulla_bar = bars + 3 * stages;
    uint64_t* tfull_bar = bars + 4 * stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                   // 0 = leader (issues the MMAs)
    const int total_tiles = m_tiles * n_tiles;                 // m_tiles counts 256-row pair tiles

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(full_bar + s), 1);
            mbar_init(smem_u32(empty_bar + s), 1);
            mbar_init(smem_u32(ready_bar + s), 2 * kXformWarps);
            mbar_init(smem_u32(fulla_bar + s), 1);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(tfull_bar + s), 1); mbar_init(smem_u32(tempty_bar + s), 2 * kEpi); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_slot), 512);   // one warp of each CTA, same warp id
    tc_fence_before();
    cluster_sync_all();                                        // barrier inits + allocation visible pair-wide
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================ TMA producer (both CTAs) ============================
        // warp-uniform loop, one elected lane issues (see elect_one)
        int stage = 0; uint32_t phase = 0;
        for (PairSched ts(total_tiles, n_tiles); ts.valid(); ts.next()) {
            const int row_a = ts.m_tile() * 256 + (int)rank * kBM;
            const int row_b = ts.n_tile() * bn + (int)rank * bh;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                const uint32_t fb = smem_u32(full_bar + stage);
                const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
                if (elect_one()) {
                    if (kXform) {
                        // A tiles feed this CTA's transform warps: own barrier; B halves feed the pair's MMA
                        const uint32_t fa = smem_u32(fulla_bar + stage);
                        mbar_expect_tx(fa, 2 * a_bytes);
                        tma_load_2d(base, &tmA_hi, kb * BK, row_a, fa);
                        tma_load_2d(base + a_bytes, &tmA_lo, kb * BK, row_a, fa);
                        if (rank == 0) mbar_expect_tx(fb, 4 * bh_bytes);
                    } else {
                        if (rank == 0) mbar_expect_tx(fb, 2 * stage_bytes);
                        tma_load_2d_2sm(base, &tmA_hi, kb * BK, row_a, fb);
                        tma_load_2d_2sm(base + a_bytes, &tmA_lo, kb * BK, row_a, fb);
                    }
                    tma_load_2d_2sm(base + 2 * a_bytes, &tmB_hi, kb * BK, row_b, fb);
                    tma_load_2d_2sm(base + 2 * a_bytes + bh_bytes, &tmB_lo, kb * BK, row_b, fb);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer (leader CTA only) ============================
        // warp-uniform loop, one elected lane issues (see elect_one)
        if (rank == 0) {
            const uint32_t idesc = make_idesc_2sm(bn);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (PairSched ts(total_tiles, n_tiles); ts.valid(); ts.next()) {
                mbar_wait_cluster(smem_u32(tempty_bar + acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    if (kXform) mbar_wait_cluster(smem_u32(ready_bar + stage), phase);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < BK / 8; ++ks) {
                            const uint64_t a_hi = make_smem_desc<BK>(base + ks * 32);
                            const uint64_t a_lo = make_smem_desc<BK>(base + a_bytes + ks * 32);
                            const uint64_t b_hi = make_smem_desc<BK>(base + 2 * a_bytes + ks * 32);
                            const uint64_t b_lo = make_smem_desc<BK>(base + 2 * a_bytes + bh_bytes + ks * 32);
                            umma_tf32_2sm(d_tmem, a_hi, b_lo, idesc, (kb | ks) != 0 ? 1u : 0u);
                            umma_tf32_2sm(d_tmem, a_lo, b_hi, idesc, 1u);
                            umma_tf32_2sm(d_tmem, a_hi, b_hi, idesc, 1u);
                        }
                        umma_commit_2sm(smem_u32(empty_bar + stage));     // frees the slot in both CTAs
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit_2sm(smem_u32(tfull_bar + acc));   // accumulator complete, both CTAs
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (kXform && warp < kFirstEpiWarp) {
        // ============================ transform warps (product 1, both CTAs) ============================
        const int xt = threadIdx.x - 64;
        const float beta = args.it.beta;
        constexpr int kVec = kBM * BK / 4;
        int stage = 0; uint32_t phase = 0;
        for (PairSched ts(total_tiles, n_tiles); ts.valid(); ts.next()) {
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32(fulla_bar + stage), phase);
                float4* t0 = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes);
                float4* t1 = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes + a_bytes);
#pragma unroll
                for (int i = xt; i < kVec; i += 32 * kXformWarps) {
                    const float4 y = t0[i], yp = t1[i];
                    float4 hi, lo;
                    split_tf32(momentum(y.x, yp.x, beta), hi.x, lo.x);
                    split_tf32(momentum(y.y, yp.y, beta), hi.y, lo.y);
                    split_tf32(momentum(y.z, yp.z, beta), hi.z, lo.z);
                    split_tf32(momentum(y.w, yp.w, beta), hi.w, lo.w);
                    t0[i] = hi;
                    t1[i] = lo;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(smem_u32(ready_bar + stage), 0);   // the leader's barrier
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ============================ epilogue warps (both CTAs, own 128 rows) ============================
        const int ew = warp - kFirstEpiWarp;
        const int q = warp & 3;
        const int part = ew >> 2;
        constexpr int kParts = kEpi / 4;
        float* buf = epi_buf + (warp - 2) * kEpiBufFloats;
        const int nblk = (bn + 31) / 32;
        int acc = 0; uint32_t acc_phase = 0;
        for (PairSched ts(total_tiles, n_tiles); ts.valid(); ts.next()) {
            mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
            tc_fence_after();
            const int row_base = ts.m_tile() * 256 + (int)rank * kBM + q * 32;
            for (int blk = part; blk < nblk; blk += kParts) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride + blk * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) buf[lane * 33 + j] = __uint_as_float(v[j]);
                __syncwarp();
                epilogue_block<PHASE>(args, buf, lane, row_base, blk, ts.n_tile(), bn, ncols_valid, Cdbg, ldc);
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(smem_u32(tempty_bar + acc), 0);     // the leader's barrier
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    __syncwarp();
    tc_fence_before();
    cluster_sync_all();                 // nobody leaves (or frees TMEM) while the peer may still use this CTA
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

}  // namespace

size_t smem_bytes2(int bk, int bn, int stages) {
    const size_t stage = (size_t)(2 * kBM + bn) * bk * 4;      // two A tiles + two half-B tiles per CTA
    return 1024 + stages * stage + (size_t)kWorkWarps * kEpiBufFloats * 4 + (4 * stages + 4) * 8 + 16;
}

int pick_stages2(int bk, int bn, size_t smem_limit) {
    int s = 8;
    while (s > 2 && smem_bytes2(bk, bn, s) > smem_limit) --s;
    return s;
}

template <int PHASE, int BK>
static int launch_one2(const GemmDesc& g, const BatchKernelArgs& args, float* C, int ldc, int num_sms, cudaStream_t s) {
    auto kern = tc_gemm2_kernel<PHASE, BK>;
    const size_t smem = smem_bytes2(BK, g.bn, g.stages);
    GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = g.m_tiles * g.n_tiles;                   // pair tiles
    const int pairs = std::min(tiles, num_sms / 2);
    kern<<<2 * pairs, kThreads, smem, s>>>(g.tmA_hi, g.tmA_lo, g.tmB_hi, g.tmB_lo, g.k_pad / BK, g.m_tiles, g.n_tiles, g.bn,
                                           g.stages, args, C, ldc, g.ncols_valid);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int launch_gemm2(int phase, const GemmDesc& g, const BatchKernelArgs& args, float* C, int ldc, int num_sms, cudaStream_t s) {
    if (g.bk == 16) {
        if (phase == 0) return launch_one2<0, 16>(g, args, C, ldc, num_sms, s);
        if (phase == 1) return launch_one2<1, 16>(g, args, C, ldc, num_sms, s);
        return launch_one2<2, 16>(g, args, C, ldc, num_sms, s);
    }
    if (phase == 0) return launch_one2<0, 32>(g, args, C, ldc, num_sms, s);
    if (phase == 1) return launch_one2<1, 32>(g, args, C, ldc, num_sms, s);
    return launch_one2<2, 32>(g, args, C, ldc, num_sms, s);
}

}  // namespace tc
}  // namespace gpad
