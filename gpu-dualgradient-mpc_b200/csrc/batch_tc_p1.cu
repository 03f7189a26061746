// batch_tc_p1.cu -- product 1 of the throughput mode, second generation (GPAD_PREC_TF32X3, cta_group::1).
//
//     P_v[B x n] = Y_v[B x m] * M_G^T            zhat_v = P_v + beta_v (P_v - P_{v-1}) - g_P     (P-formulation)
//
// What the microbenchmark (tests/ubench/ubench_tc.cu, profiles/r1_ubench_*.log) showed about the first-generation
// kernel (batch_tc.cu, PHASE 1): it is bound by shared-memory bandwidth (TMA fill + in-place hi/lo rewrite of the A
// tiles + six MMAs per k-block that each re-read A and B from shared memory ~ 220 B/clk against 128 B/clk), and its
// state tiles come from HBM with ~2 us of loaded latency while sharing ring slots with operator tiles from L2.
// This kernel therefore
//   * stages ONLY y_v (8 KB per k-block) -- M_G w_v is linear in y, the momentum is applied to P in the epilogue;
//   * keeps the state tiles in their own deep ring of small slots and the operator tiles in a shallow ring of
//     large slots, each with its own producer warp;
//   * never writes the split A operand back to shared memory: four transform warps read a y tile (thread = batch
//     row), split it into tf32 hi | lo in registers and tcgen05.st it into a 3-slot TMEM ring; the MMAs take A from
//     TMEM (tcgen05.mma [d], [a_tmem], b_desc), so only B is read from shared memory.
// TMEM: 2 accumulator stages x bn (<= 208) columns + 3 x 32 columns of A ring = 512.
// Warps: 0 operator producer, 1 TMEM allocator + MMA issuer, 2 state producer, 3..6 transform (TMEM lane
// quarter = warp % 4), 7.. epilogue (8 warps for product 1).  All single-thread instructions are issued from warp-uniform loops (elect.sync).
#include <cuda.h>

#include <algorithm>

#include "batch_common.cuh"
#include "batch_tc.h"
#include "gpad_internal.h"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace gpad {
namespace tc {

namespace {

constexpr int kP1FirstXform = 3;       // warps 3..6 transform (any 4 consecutive warps cover the 4 TMEM lane quarters)
constexpr int kP1FirstEpi = 7;
constexpr int kP1TStages = 3;          // TMEM A ring slots (32 columns each: hi 16 | lo 16)
constexpr int kP1BK = 16;

// tiles in (batch tile, operator tile) order; with a tile list (tolerance mode) only the listed batch tiles exist
struct P1Sched {
    int tile, step, total, n_tiles;
    const int* list;
    __device__ P1Sched(int m_tiles, int n_tiles_, const BatchKernelArgs& a)
        : tile(blockIdx.x), step(gridDim.x), total((a.tile_count ? *a.tile_count : m_tiles) * n_tiles_), n_tiles(n_tiles_), list(a.tile_list) {
        if (a.dual && a.dual_count && *a.dual_count == 0) total = 0;      // nobody waits for the dual-gap evaluation
    }
    __device__ bool valid() const { return tile < total; }
    __device__ void next() { tile += step; }
    __device__ int m_tile() const { return list ? list[tile / n_tiles] : tile / n_tiles; }
    __device__ int n_tile() const { return tile % n_tiles; }
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// F16 (GPAD_PREC_FP16X3, fixed-iteration solves): a k-block is 32 K elements.  The y tile arrives as 128 x 32 fp32
// (128-byte rows, SWIZZLE_128B), the transform warps scale each row by its power of two (from the row maximum the
// previous product 2 reduced), split it into fp16 hi | lo pairs -- the same 32 TMEM columns per slot -- and the MMAs are
// kind::f16 with K = 16: half the instructions and half the operator bytes per flop of the tf32 form.
template <int EPI, bool TOL, bool F16>
__global__ void __launch_bounds__(32 * (kP1FirstEpi + EPI), 1)
tc_p1_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmB_hi,
             const __grid_constant__ CUtensorMap tmB_lo, int num_k_blocks, int m_tiles, int n_tiles, int bn,
             int a_stages, int b_stages, const BatchKernelArgs args, int ncols_valid, int step, int acc_stages) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int BK = kP1BK;                                       // operator tile rows are BK * 4 = 64 bytes
    constexpr int kElems = F16 ? 2 * BK : BK;                       // K elements per k-block
    constexpr uint32_t a_bytes = kBM * kElems * 4;                  // fp32 y tile
    const uint32_t b_bytes = (uint32_t)bn * BK * 4;                 // one of hi / lo
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + (size_t)a_stages * a_bytes;
    float* epi_buf = reinterpret_cast<float*>(b_ring + (size_t)b_stages * 2 * b_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + EPI * kEpiBufFloats);
    uint64_t* afull = bars;
    uint64_t* aempty = afull + a_stages;
    uint64_t* bfull = aempty + a_stages;
    uint64_t* bempty = bfull + b_stages;
    uint64_t* ready = bempty + b_stages;            // TMEM A slot written
    uint64_t* tfree = ready + kP1TStages;           // TMEM A slot consumed
    uint64_t* tfull_bar = tfree + kP1TStages;       // accumulator complete
    uint64_t* tempty_bar = tfull_bar + 2;           // accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // two accumulator stages of bn <= 208 columns (2 * 208 + 96 = 512), or ONE of up to 256 columns for plans whose tiles
    // fit a single wave over the SMs (nothing to overlap the epilogue with; fewer, wider tiles)
    const uint32_t acc_stride = (uint32_t)bn;
    const uint32_t a_col0 = (uint32_t)acc_stages * acc_stride;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
        for (int s = 0; s < a_stages; ++s) { mbar_init(smem_u32(afull + s), 1); mbar_init(smem_u32(aempty + s), 4); }
        for (int s = 0; s < b_stages; ++s) { mbar_init(smem_u32(bfull + s), 1); mbar_init(smem_u32(bempty + s), 1); }
        for (int s = 0; s < kP1TStages; ++s) { mbar_init(smem_u32(ready + s), 4); mbar_init(smem_u32(tfree + s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(tfull_bar + s), 1); mbar_init(smem_u32(tempty_bar + s), EPI); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    grid_dependency_wait();          // PDL: the previous kernel's memory is needed from here on
    grid_launch_dependents();

    if (warp == 0) {
        // ============================ operator producer (M_G hi / lo, L2 resident) ============================
        int s = 0; uint32_t ph = 0;
        for (P1Sched ts(m_tiles, n_tiles, args); ts.valid(); ts.next()) {
            const int row_b = ts.n_tile() * step;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32(bempty + s), ph ^ 1);
                if (elect_one()) {
                    const uint32_t fb = smem_u32(bfull + s);
                    const uint32_t bb = smem_u32(b_ring + (size_t)s * 2 * b_bytes);
                    mbar_expect_tx(fb, 2 * b_bytes);
                    tma_load_2d(bb, &tmB_hi, kb * kElems, row_b, fb);
                    tma_load_2d(bb + b_bytes, &tmB_lo, kb * kElems, row_b, fb);
                }
                __syncwarp();
                if (++s == b_stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ============================ state producer (y_v tiles, HBM) ============================
        int s = 0; uint32_t ph = 0;
        for (P1Sched ts(m_tiles, n_tiles, args); ts.valid(); ts.next()) {
            const int row_a = ts.m_tile() * kBM;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32(aempty + s), ph ^ 1);
                if (elect_one()) {
                    const uint32_t fb = smem_u32(afull + s);
                    mbar_expect_tx(fb, a_bytes);
                    tma_load_2d(smem_u32(a_ring + (size_t)s * a_bytes), &tmY, kb * kElems, row_a, fb);
                }
                __syncwarp();
                if (++s == a_stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        const uint32_t idesc = make_idesc_k<F16>(bn);
        int s = 0; uint32_t ph = 0;
        int t = 0; uint32_t tph = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (P1Sched ts(m_tiles, n_tiles, args); ts.valid(); ts.next()) {
            mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32(bfull + s), ph);
                mbar_wait(smem_u32(ready + t), tph);
                tc_fence_after();
                const uint32_t bb = smem_u32(b_ring + (size_t)s * 2 * b_bytes);
                const uint32_t at = tmem_base + a_col0 + (uint32_t)t * 32u;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint64_t b_hi = make_smem_desc<BK>(bb + ks * 32);
                        const uint64_t b_lo = make_smem_desc<BK>(bb + b_bytes + ks * 32);
                        umma_ts<F16>(d_tmem, at + ks * 8, b_lo, idesc, (kb | ks) != 0 ? 1u : 0u);        // hi * lo
                        umma_ts<F16>(d_tmem, at + 16 + ks * 8, b_hi, idesc, 1u);                         // lo * hi
                        umma_ts<F16>(d_tmem, at + ks * 8, b_hi, idesc, 1u);                              // hi * hi
                    }
                    umma_commit(smem_u32(bempty + s));
                    umma_commit(smem_u32(tfree + t));
                }
                __syncwarp();
                if (++s == b_stages) { s = 0; ph ^= 1; }
                if (++t == kP1TStages) { t = 0; tph ^= 1; }
            }
            if (elect_one()) umma_commit(smem_u32(tfull_bar + acc));
            __syncwarp();
            if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= kP1FirstXform && warp < kP1FirstEpi) {
        // ============================ transform warps: y tile -> tf32 hi | lo -> TMEM A ring ============================
        // TMA wrote the tile with SWIZZLE_64B: 16-byte chunk c of row r sits at r * 64 + ((c ^ ((r >> 1) & 3)) << 4)
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        int s = 0; uint32_t ph = 0;
        int t = 0; uint32_t tph = 0;
        for (P1Sched ts(m_tiles, n_tiles, args); ts.valid(); ts.next()) {
            float row_scale = 1.f;
            if (F16) row_scale = pow2f(f16_scale_exp(__ldg(args.a_rowmax + ts.m_tile() * kBM + row)));
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(smem_u32(afull + s), ph);
                uint32_t v[32];
                if (F16) {
                    // SWIZZLE_128B: 16-byte chunk c of row r sits at r * 128 + ((c ^ (r & 7)) << 4)
                    const uint8_t* tile = a_ring + (size_t)s * a_bytes + row * 128;
                    float4 y[8];
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) y[ch] = *reinterpret_cast<const float4*>(tile + ((ch ^ (row & 7)) << 4));
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(aempty + s));       // the shared-memory slot is free again
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        // TMEM column j of the slot holds K elements 2j, 2j + 1 (hi), column 16 + j their lo parts
                        split_f16x2(y[ch].x * row_scale, y[ch].y * row_scale, v[ch * 2], v[16 + ch * 2]);
                        split_f16x2(y[ch].z * row_scale, y[ch].w * row_scale, v[ch * 2 + 1], v[16 + ch * 2 + 1]);
                    }
                } else {
                const uint8_t* tile = a_ring + (size_t)s * a_bytes + row * 64;
                float4 y[4];
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) y[ch] = *reinterpret_cast<const float4*>(tile + ((ch ^ ((row >> 1) & 3)) << 4));
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(aempty + s));       // the shared-memory slot is free again
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const float e[4] = {y[ch].x, y[ch].y, y[ch].z, y[ch].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float hi, lo;
                        split_tf32(e[j], hi, lo);
                        v[ch * 4 + j] = __float_as_uint(hi);
                        v[16 + ch * 4 + j] = __float_as_uint(lo);
                    }
                }
                }
                mbar_wait(smem_u32(tfree + t), tph ^ 1);                // the MMAs that read this TMEM slot have retired
                tc_fence_after();
                tmem_st32(tmem_base + lane_addr + a_col0 + (uint32_t)t * 32u, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(ready + t));
                if (++s == a_stages) { s = 0; ph ^= 1; }
                if (++t == kP1TStages) { t = 0; tph ^= 1; }
            }
        }
    } else if (warp >= kP1FirstEpi) {
        // ============================ epilogue warps ============================
        const int ew = warp - kP1FirstEpi;
        const int q = warp & 3;
        const int part = ew >> 2;
        constexpr int kParts = EPI / 4;
        float* buf = epi_buf + ew * kEpiBufFloats;
        const int nblk = (bn + 31) / 32;
        int acc = 0; uint32_t acc_phase = 0;
        for (P1Sched ts(m_tiles, n_tiles, args); ts.valid(); ts.next()) {
            mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
            tc_fence_after();
            const int row_base = ts.m_tile() * kBM + q * 32;
            float row_inv = 1.f;         // F16: this thread's accumulator row carries the scale of its y row
            if (F16) row_inv = pow2f(-f16_scale_exp(__ldg(args.a_rowmax + row_base + lane)));
            for (int blk = part; blk < nblk; blk += kParts) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_stride + (uint32_t)(blk * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) buf[lane * 33 + j] = F16 ? __uint_as_float(v[j]) * row_inv : __uint_as_float(v[j]);
                __syncwarp();
                epilogue_block<1, TOL, F16>(args, buf, lane, row_base, blk, ts.n_tile(), bn, ncols_valid, nullptr, 0, step);
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty_bar + acc));
            if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
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

size_t p1_smem_bytes(int bn, int a_stages, int b_stages, int epi, bool f16 = false) {
    return 1024 + (size_t)a_stages * kBM * (f16 ? 2 : 1) * kP1BK * 4 + (size_t)b_stages * 2 * bn * kP1BK * 4 +
           (size_t)epi * kEpiBufFloats * 4 + (2 * a_stages + 2 * b_stages + 2 * kP1TStages + 4) * 8 + 16;
}

}  // namespace

// tiles of at most 208 columns: 2 x 208 accumulator columns + 96 columns of A ring fill the 512 TMEM columns
// (max_bn = 256 for the single-accumulator plan)
void plan_tiles_p1(int ncols, int* bn, int* n_tiles, int* step, int max_bn) {
    const int nt = (ncols + max_bn - 1) / max_bn;
    int b = ((ncols + nt - 1) / nt + 15) / 16 * 16;
    if (b < 16) b = 16;
    *bn = b;
    *n_tiles = nt;
    // start the tiles on 128-byte lines when the same number of tiles still covers all columns that way
    const int s32 = b / 32 * 32;
    if (step) *step = (nt > 1 && s32 > 0 && (nt - 1) * s32 + b >= ncols) ? s32 : b;
}

int plan_rings_p1(int bn, size_t smem_limit, int* a_stages, int* b_stages, bool f16) {
    const int epi = 8;
    int b = 5, a = f16 ? 4 : 8;
    while (b > 2 && p1_smem_bytes(bn, a, b, epi, f16) > smem_limit) --b;
    while (a > 2 && p1_smem_bytes(bn, a, b, epi, f16) > smem_limit) --a;
    if (p1_smem_bytes(bn, a, b, epi, f16) > smem_limit) return GPAD_ERR_UNSUPPORTED;
    while (a < 12 && p1_smem_bytes(bn, a + 1, b, epi, f16) <= smem_limit) ++a;     // spare shared memory deepens the HBM-facing ring
    *a_stages = a; *b_stages = b;
    return GPAD_OK;
}

// P_v = Y_v M_G^T (A = y_v)
int launch_p1(const GemmDesc& g, const BatchKernelArgs& args, int num_sms, cudaStream_t s) {
    const size_t smem = p1_smem_bytes(g.bn, g.a_stages, g.stages, 8, g.f16 != 0);
    const int tiles = g.m_tiles * g.n_tiles;
    const bool tol = args.checking || args.dual || args.done;
    if (g.f16 && tol) { set_error("tcgen05 product 1: the fp16 kernel serves fixed-iteration solves only"); return GPAD_ERR_UNSUPPORTED; }
    auto kern = g.f16 ? tc_p1_kernel<8, false, true> : tol ? tc_p1_kernel<8, true, false> : tc_p1_kernel<8, false, false>;
    GPAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(std::min(tiles, num_sms)); lc.blockDim = dim3(32 * (kP1FirstEpi + 8)); lc.dynamicSmemBytes = smem; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = g.pdl ? 1 : 0;
    GPAD_CUDA(cudaLaunchKernelEx(&lc, kern, g.tmA_hi, g.tmB_hi, g.tmB_lo, g.k_pad / (g.f16 ? 2 * kP1BK : kP1BK), g.m_tiles, g.n_tiles, g.bn, g.a_stages,
                                 g.stages, args, g.ncols_valid, g.step > 0 ? g.step : g.bn, g.acc_stages));
    return GPAD_OK;
}

}  // namespace tc
}  // namespace gpad
