// batch_tc.cu -- throughput mode on the 5th-generation tensor cores (GPAD_PREC_TF32X3).
//
// Each GPAD iteration over a batch of B QPs sharing M_G / G_L is two GEMMs with the batch on
// the MMA M dimension:
//     Zhat[B x n] = W[B x m]    * M_G^T      (epilogue1: -g_P, z average, tf32 split of zhat)
//     Y+  [B x m] = Zhat[B x n] * G_L^T      (epilogue2: +(w+p_D), projection, momentum, split of w+)
// fp32 accuracy from kind::tf32: both operands are split x = hi + lo (hi = RN_tf32(x),
// lo = RN_tf32(x - hi)) and every product is three MMAs into the same fp32 TMEM accumulator:
//     hi*lo + lo*hi + hi*hi          (lo*lo ~ 2^-22 relative is dropped: "3xTF32")
//
// Kernel anatomy (one CTA per SM, persistent over output tiles of 128 x bn, bn <= 256):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of two A tiles [128 x BK] and B_hi/B_lo [bn x BK]
//               into a `stages`-deep shared-memory ring (K-major, hardware swizzle), mbarrier tx-count
//   warp 1      TMEM allocator + MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::tf32,
//               tcgen05.commit releases ring slots and publishes finished accumulators
//   warps 2..13 product 2 / test hook: 12 epilogue warps: tcgen05.ld 32x32b -> registers -> per-warp
//               smem transpose -> coalesced global reads/writes of the fused GPAD epilogue
//               product 1: warps 2..5 are TRANSFORM warps -- the A tile that arrives is y_v (fp32); they
//               rewrite it in place as tf32 hi and write lo next to it (the momentum is applied to the
//               product in the epilogue, P-formulation), fence the generic->async proxy and hand the slot
//               to the MMA warp; warps 6..13 are the epilogue
//   TMEM: 512 columns = 2 accumulator stages x 256 fp32 columns, so the epilogue of tile t
//   overlaps the mainloop of tile t+1.
// Launch boundaries: every kernel is launched with programmatic stream serialization (PDL): its CTAs start as soon
// as an SM frees up, set up barriers / TMEM / descriptors, and only then wait for the previous kernel's memory
// (griddepcontrol.wait), so the tail of one product overlaps the prologue of the next.
// Tolerance mode: the batch tiles that still hold a running instance arrive as a dense list (tile retirement).
#include <cuda.h>

#include <algorithm>

#include "batch_common.cuh"
#include "gpad_internal.h"
#include "batch_tc.h"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

namespace gpad {
namespace tc {

namespace {

// work units are (batch tile, operator tile), operator tile fastest so that the tiles of one batch tile run on
// neighbouring CTAs at the same time (DRAM locality of the m-sized epilogue rows).  With a tile
// list (tolerance mode) only the listed batch tiles exist.
struct TileSched {
    int unit, step, total, n_tiles;
    const int* list;
    __device__ TileSched(int m_tiles, int n_tiles_, const BatchKernelArgs& a)
        : unit(blockIdx.x), step(gridDim.x), total((a.tile_count ? *a.tile_count : m_tiles) * n_tiles_), n_tiles(n_tiles_), list(a.tile_list) {
        if (a.dual && a.dual_count && *a.dual_count == 0) total = 0;      // nobody waits for the dual-gap evaluation
    }
    __device__ bool valid() const { return unit < total; }
    __device__ void next() { unit += step; }
    __device__ int m_tile() const { return list ? list[unit / n_tiles] : unit / n_tiles; }
    __device__ int n_tile() const { return unit % n_tiles; }
};

// warp roles per instantiation: warps 0/1 producer + MMA, then transform warps, then epilogue warps
//   product 1 (first generation): 4 transform + 8 epilogue; product 2 / test hook: 12 epilogue
__host__ __device__ constexpr int xform_warps(int phase) { return phase == 1 ? kXformWarps : 0; }
__host__ __device__ constexpr int epi_warps(int phase) { return kWorkWarps - xform_warps(phase); }
__host__ __device__ constexpr int cta_threads(int phase) { return 32 * (2 + kWorkWarps); }

// PHASE 0: plain store of C (test hook); 1: GPAD product 1; 2: GPAD product 2
// F16 (GPAD_PREC_FP16X3, test hook of the mainloop batch_tc_p2.cu shares): the operand tiles hold fp16 hi / lo of row-scaled
// values -- the same 64-byte rows, twice the K per row (BK counts 4-byte units), kind::f16 MMAs; the epilogue undoes the row
// and column scales on the accumulator
template <int PHASE, int BK, bool TOL, bool F16>
__global__ void __launch_bounds__(cta_threads(PHASE), 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               int num_k_blocks, int m_tiles, int n_tiles, int bn, int stages,
               const BatchKernelArgs args, float* __restrict__ Cdbg, int ldc, int ncols_valid) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_bytes = kBM * BK * 4, b_bytes = (uint32_t)bn * BK * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    constexpr int kElemsPerBlock = F16 ? 2 * BK : BK;      // K elements of one k-block (TMA coordinates count elements)
    constexpr bool kXform = PHASE == 1;    // transform warps: the A tile is y_v (fp32) -> tf32 hi in place, lo next to it
    constexpr int kXW = xform_warps(PHASE);
    constexpr int kEpi = epi_warps(PHASE);
    constexpr int kFirstEpiWarp = 2 + kXW;
    float* epi_buf = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + kWorkWarps * kEpiBufFloats);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + stages;
    uint64_t* ready_bar = bars + 2 * stages;                  // product 1: transform warps -> MMA warp
    uint64_t* tfull_bar = bars + 3 * stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); mbar_init(smem_u32(ready_bar + s), kXW > 0 ? kXW : 1);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(tfull_bar + s), 1); mbar_init(smem_u32(tempty_bar + s), kEpi); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // PDL: everything above touched no global memory; from here on the previous kernel's writes are needed
    grid_dependency_wait();
    grid_launch_dependents();

    if (warp == 0) {
        // ============================ TMA producer ============================
        // warp-uniform loop, one elected lane issues (see elect_one)
        int stage = 0; uint32_t phase = 0;
        for (TileSched ts(m_tiles, n_tiles, args); ts.valid(); ts.next()) {
            const int row_a = ts.m_tile() * kBM, row_b = ts.n_tile() * bn;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                const uint32_t fb = smem_u32(full_bar + stage);
                const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
                if (elect_one()) {
                    // product 1 stages one fp32 tile (y_v); its lo half is produced in place by the transform warps
                    mbar_expect_tx(fb, kXform ? stage_bytes - a_bytes : stage_bytes);
                    tma_load_2d(base, &tmA_hi, kb * kElemsPerBlock, row_a, fb);
                    if (!kXform) tma_load_2d(base + a_bytes, &tmA_lo, kb * kElemsPerBlock, row_a, fb);
                    tma_load_2d(base + 2 * a_bytes, &tmB_hi, kb * kElemsPerBlock, row_b, fb);
                    tma_load_2d(base + 2 * a_bytes + b_bytes, &tmB_lo, kb * kElemsPerBlock, row_b, fb);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        // warp-uniform loop (waits and descriptor arithmetic in uniform registers), one elected lane issues
        const uint32_t idesc = make_idesc_k<F16>(bn);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (TileSched ts(m_tiles, n_tiles, args); ts.valid(); ts.next()) {
            mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32((kXform ? ready_bar : full_bar) + stage), phase);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint64_t a_hi = make_smem_desc<BK>(base + ks * 32);
                        const uint64_t a_lo = make_smem_desc<BK>(base + a_bytes + ks * 32);
                        const uint64_t b_hi = make_smem_desc<BK>(base + 2 * a_bytes + ks * 32);
                        const uint64_t b_lo = make_smem_desc<BK>(base + 2 * a_bytes + b_bytes + ks * 32);
                        umma_ss<F16>(d_tmem, a_hi, b_lo, idesc, (kb | ks) != 0 ? 1u : 0u);
                        umma_ss<F16>(d_tmem, a_lo, b_hi, idesc, 1u);
                        umma_ss<F16>(d_tmem, a_hi, b_hi, idesc, 1u);
                    }
                    umma_commit(smem_u32(empty_bar + stage));     // frees the ring slot when these MMAs retire
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(smem_u32(tfull_bar + acc));   // accumulator complete
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (kXform && warp < kFirstEpiWarp) {
        // ============================ transform warps (product 1) ============================
        // slot layout: [y_v tile | (lo) | B_hi | B_lo]; tile0 <- RN_tf32(y), tile1 <- RN_tf32(y - tile0): same swizzle,
        // so the rewrite is purely elementwise
        const int xt = threadIdx.x - 64;                        // 0..127
        constexpr int kVec = kBM * BK / 4;                      // float4 per A tile
        int stage = 0; uint32_t phase = 0;
        for (TileSched ts(m_tiles, n_tiles, args); ts.valid(); ts.next()) {
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32(full_bar + stage), phase);
                float4* t0 = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes);
                float4* t1 = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes + a_bytes);
#pragma unroll
                for (int i = xt; i < kVec; i += 32 * kXW) {
                    const float4 y = t0[i];
                    float4 hi, lo;
                    split_tf32(y.x, hi.x, lo.x);
                    split_tf32(y.y, hi.y, lo.y);
                    split_tf32(y.z, hi.z, lo.z);
                    split_tf32(y.w, hi.w, lo.w);
                    t0[i] = hi;
                    t1[i] = lo;
                }
                fence_proxy_async_smem();                       // generic-proxy writes -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(ready_bar + stage));
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ============================ epilogue warps ============================
        const int ew = warp - kFirstEpiWarp;
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int part = ew >> 2;               // warps sharing a quarter alternate over column blocks
        constexpr int kParts = kEpi / 4;
        float* buf = epi_buf + ew * kEpiBufFloats;
        const int nblk = (bn + 31) / 32;
        int acc = 0; uint32_t acc_phase = 0;
        for (TileSched ts(m_tiles, n_tiles, args); ts.valid(); ts.next()) {
            mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
            tc_fence_after();
            const int row_base = ts.m_tile() * kBM + q * 32;
            // F16: this thread's accumulator row (TMEM lane = batch row) carries the row scale of the A operand
            float row_inv = 1.f;
            if (F16) row_inv = __ldg(args.a_rowinv + row_base + lane);
            for (int blk = part; blk < nblk; blk += kParts) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride + blk * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) buf[lane * 33 + j] = F16 ? __uint_as_float(v[j]) * row_inv : __uint_as_float(v[j]);
                __syncwarp();
                epilogue_block<PHASE, TOL, F16>(args, buf, lane, row_base, blk, ts.n_tile(), bn, ncols_valid, Cdbg, ldc);
                __syncwarp();
            }
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

// x -> (RN_tf32(x), RN_tf32(x - hi)), elementwise
__global__ void split_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        float h, l;
        split_tf32(src[i], h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

int make_tmap(CUtensorMap* map, const float* ptr, int k_elems, int rows, int ld, int box_k, int box_rows) {
    return make_tmap_bytes(map, ptr, 4, k_elems, rows, ld, box_k, box_rows);
}

// elem_bytes 4: fp32 rows, 2: fp16 rows; k_elems / ld / box_k count elements.  The swizzle span is the box row.
int make_tmap_bytes(CUtensorMap* map, const void* ptr, int elem_bytes, int k_elems, int rows, int ld, int box_k, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return GPAD_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)k_elems, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const int row_bytes = box_k * elem_bytes;
    const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return GPAD_ERR_CUDA; }
    return GPAD_OK;
}

void plan_tiles(int ncols, int* bn, int* n_tiles) {
    const int nt = (ncols + 255) / 256;
    int b = ((ncols + nt - 1) / nt + 15) / 16 * 16;
    if (b < 16) b = 16;
    *bn = b;
    *n_tiles = nt;
}

size_t smem_bytes(int bk, int bn, int stages) {
    const size_t stage = (size_t)(2 * kBM + 2 * bn) * bk * 4;
    return 1024 + stages * stage + (size_t)kWorkWarps * kEpiBufFloats * 4 + (3 * stages + 4) * 8 + 16;
}

int pick_stages(int bk, int bn, size_t smem_limit) {
    int s = 8;
    while (s > 2 && smem_bytes(bk, bn, s) > smem_limit) --s;
    return s;
}

template <int PHASE, int BK, bool TOL, bool F16 = false>
static int launch_one(const GemmDesc& g, const BatchKernelArgs& args, float* C, int ldc, int num_sms, cudaStream_t s) {
    auto kern = tc_gemm_kernel<PHASE, BK, TOL, F16>;
    const size_t smem = smem_bytes(BK, g.bn, g.stages);
    GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int units = g.m_tiles * g.n_tiles;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(std::min(units, num_sms)); lc.blockDim = dim3(cta_threads(PHASE)); lc.dynamicSmemBytes = smem; lc.stream = s;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (g.cluster_attr) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = 1; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    if (g.pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    lc.attrs = at; lc.numAttrs = na;
    GPAD_CUDA(cudaLaunchKernelEx(&lc, kern, g.tmA_hi, g.tmA_lo, g.tmB_hi, g.tmB_lo, g.k_pad / (F16 ? 2 * BK : BK), g.m_tiles,
                                 g.n_tiles, g.bn, g.stages, args, C, ldc, g.ncols_valid));
    return GPAD_OK;
}

int launch_gemm(int phase, const GemmDesc& g, const BatchKernelArgs& args, float* C, int ldc, int num_sms, cudaStream_t s) {
    if (g.bk != 16) { set_error("tcgen05 GEMM: K block %d is not built (16 only)", g.bk); return GPAD_ERR_UNSUPPORTED; }
    // fixed-iteration solves run the lean instantiation; tolerance mode (stopped rows, reductions, dual-gap launches) its own
    const bool tol = args.checking || args.dual || args.done;
    if (g.f16) {
        // fp16 hi / lo operands: the test hook of this mainloop only (product 1: batch_tc_p1.cu, product 2: batch_tc_p2.cu)
        if (phase != 0) { set_error("tcgen05 GEMM: no fp16 instantiation for this launch"); return GPAD_ERR_UNSUPPORTED; }
        return launch_one<0, 16, false, true>(g, args, C, ldc, num_sms, s);
    }
    if (phase == 0) return launch_one<0, 16, false>(g, args, C, ldc, num_sms, s);
    if (phase == 1) return tol ? launch_one<1, 16, true>(g, args, C, ldc, num_sms, s) : launch_one<1, 16, false>(g, args, C, ldc, num_sms, s);
    return tol ? launch_one<2, 16, true>(g, args, C, ldc, num_sms, s) : launch_one<2, 16, false>(g, args, C, ldc, num_sms, s);
}

int launch_split(const float* src, float* hi, float* lo, size_t count, cudaStream_t s) {
    size_t g = (count + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g == 0) g = 1;
    split_kernel<<<(int)g, 256, 0, s>>>(src, hi, lo, count);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

}  // namespace tc
}  // namespace gpad
