// ubench_tc.cu -- microbenchmark behind the tile-shape decisions of the tcgen05 3xTF32 GEMM kernels
// (DESIGN.md 4.1): what bounds a persistent TMA -> smem ring -> tcgen05.mma.kind::tf32 pipeline on
// this B200 -- the tensor pipe, the L2 -> SM fabric, or shared-memory bandwidth -- and how far TMA
// multicast across a cluster moves that bound.  Not product code, not a bench value.
//
// One CTA per SM streams k-blocks of [a_tiles x 128 x 16] "state" rows (unique per CTA, HBM resident)
// and [2 x bn x 16] "operator" rows (shared by all CTAs, L2 resident) and, per k-block, issues
// 3 x (16/8) x (bn / bn_mma) MMAs of shape 128 x bn_mma x 8.  Switches: MMA on/off, TMA on/off,
// in-place transform warps on/off (product 1's w split), cluster size (operator boxes multicast).
//
//   build:  make ubench        run (on the GPU box):  build/ubench_tc
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_ptx.cuh"

using namespace gpad::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Cfg {
    int bn;          // operator rows per k-block (per hi / lo copy)
    int bn_mma;      // N of one MMA (bn % bn_mma == 0)
    int box_rows;    // operator rows per TMA box
    int a_tiles;     // 128-row state tiles per k-block (2: y_v and y_{v-1}, or hi and lo)
    int stages;
    int kblocks;     // k-blocks per tile
    int tiles;       // tiles per CTA
    int do_tma, do_mma, do_xform;
    int cluster;     // 1, 2 or 4: operator boxes are multicast to the whole cluster
    int m_tiles;     // distinct state tiles in the buffer
    int order;       // 0: the three MMAs of one accumulator back to back; 1: accumulators interleaved
    int bulk_b;      // 1: operator k-blocks are contiguous pre-packed images fetched with ONE cp.async.bulk (no per-row TMA requests)
    int bk;          // K floats per k-block: 16 (64 B rows, SWIZZLE_64B) or 32 (128 B rows, SWIZZLE_128B)
    int warp_issue;  // 0: the whole issue loop runs under `if (lane == 0)`; 1: warp-uniform loop, elect.sync around each MMA
    int f16 = 0;     // 1 (warp_issue = 1 only): kind::f16 MMAs (K = 16 per instruction from the same 32 operand bytes per row)
};

// cluster helpers (the product kernels no longer use clusters / multicast: measured no gain)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
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

template <int BK, bool F16 = false>
__global__ void __launch_bounds__(192, 1)
ub_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
          const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, const Cfg c,
          unsigned long long* __restrict__ cycles, const float* __restrict__ bsrc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_bytes = 128 * BK * 4, b_bytes = (uint32_t)c.bn * BK * 4, box_bytes = (uint32_t)c.box_rows * BK * 4;
    const uint32_t stage_bytes = c.a_tiles * a_bytes + 2 * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)c.stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + c.stages;
    uint64_t* ready_bar = bars + 2 * c.stages;
    uint64_t* done_bar = bars + 3 * c.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = c.cluster > 1 ? cluster_ctarank() : 0;
    const uint16_t mask = (uint16_t)((1u << c.cluster) - 1);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < c.stages; ++s) {
            mbar_init(smem_u32(full_bar + s), 1);
            mbar_init(smem_u32(empty_bar + s), c.cluster);
            mbar_init(smem_u32(ready_bar + s), 4);
        }
        mbar_init(smem_u32(done_bar), 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    if (c.cluster > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long t0 = clock64();

    const int nboxes = c.bn / c.box_rows;
    if (warp == 0) {
        if (lane == 0 && c.do_tma) {
            int stage = 0; uint32_t phase = 0;
            for (int t = 0; t < c.tiles; ++t) {
                const int m_tile = (int)((blockIdx.x + (long long)t * gridDim.x) % c.m_tiles);
                for (int kb = 0; kb < c.kblocks; ++kb) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                    const uint32_t fb = smem_u32(full_bar + stage);
                    mbar_expect_tx(fb, stage_bytes);
                    const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
                    tma_load_2d(base, &tmA0, kb * BK, m_tile * 128, fb);
                    if (c.a_tiles > 1) tma_load_2d(base + a_bytes, &tmA1, kb * BK, m_tile * 128, fb);
                    const uint32_t bb = base + c.a_tiles * a_bytes;
                    if (c.bulk_b) {
                        const char* src = reinterpret_cast<const char*>(bsrc) + (size_t)(kb % 64) * 2 * b_bytes;
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(bb), "l"(src), "r"(2 * b_bytes), "r"(fb) : "memory");
                    } else
                    for (int bx = (int)rank; bx < nboxes; bx += c.cluster) {
                        if (c.cluster > 1) {
                            tma_load_2d_mc(bb + bx * box_bytes, &tmB0, kb * BK, bx * c.box_rows, fb, mask);
                            tma_load_2d_mc(bb + b_bytes + bx * box_bytes, &tmB1, kb * BK, bx * c.box_rows, fb, mask);
                        } else {
                            tma_load_2d(bb + bx * box_bytes, &tmB0, kb * BK, bx * c.box_rows, fb);
                            tma_load_2d(bb + b_bytes + bx * box_bytes, &tmB1, kb * BK, bx * c.box_rows, fb);
                        }
                    }
                    if (++stage == c.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && c.warp_issue) {
        // CUTLASS-style: every lane runs the loop (uniform control flow and address arithmetic),
        // one elected lane executes each tcgen05 instruction
        const uint32_t idesc = F16 ? make_idesc_f16(c.bn_mma) : make_idesc(c.bn_mma);
        const int nh = c.bn / c.bn_mma;
        int stage = 0; uint32_t phase = 0;
        for (int t = 0; t < c.tiles; ++t) {
            for (int kb = 0; kb < c.kblocks; ++kb) {
                if (c.do_tma) {
                    mbar_wait(smem_u32((c.do_xform ? ready_bar : full_bar) + stage), phase);
                    tc_fence_after();
                }
                const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t bb = base + c.a_tiles * a_bytes;
                if constexpr (F16) {
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint64_t a_hi = make_smem_desc<BK>(base + ks * 32);
                        const uint64_t a_lo = make_smem_desc<BK>(base + a_bytes + ks * 32);
                        for (int h = 0; h < nh; ++h) {
                            const uint32_t off = (uint32_t)h * c.bn_mma * BK * 4;
                            const uint64_t b_hi = make_smem_desc<BK>(bb + off + ks * 32);
                            const uint64_t b_lo = make_smem_desc<BK>(bb + b_bytes + off + ks * 32);
                            const uint32_t d = tmem_base + (uint32_t)h * c.bn_mma;
                            if (elect_one()) {
                                if (c.do_xform == 2) {
                                    const uint32_t at = tmem_base + (uint32_t)((c.bn + 31) & ~31) + (uint32_t)stage * 32u + (uint32_t)ks * 8u;
                                    umma_f16_ts(d, at, b_lo, idesc, (t | kb | ks) != 0 ? 1u : 0u);
                                    umma_f16_ts(d, at + 16u, b_hi, idesc, 1u);
                                    umma_f16_ts(d, at, b_hi, idesc, 1u);
                                } else {
                                    umma_f16(d, a_hi, b_lo, idesc, (t | kb | ks) != 0 ? 1u : 0u);
                                    umma_f16(d, a_lo, b_hi, idesc, 1u);
                                    umma_f16(d, a_hi, b_hi, idesc, 1u);
                                }
                            }
                        }
                    }
                } else
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t a_hi = make_smem_desc<BK>(base + ks * 32);
                    const uint64_t a_lo = make_smem_desc<BK>(base + a_bytes + ks * 32);
                    for (int h = 0; h < nh; ++h) {
                        const uint32_t off = (uint32_t)h * c.bn_mma * BK * 4;
                        const uint64_t b_hi = make_smem_desc<BK>(bb + off + ks * 32);
                        const uint64_t b_lo = make_smem_desc<BK>(bb + b_bytes + off + ks * 32);
                        const uint32_t d = tmem_base + (uint32_t)h * c.bn_mma;
                        if (elect_one()) {
                            if (c.do_xform == 2) {      // A (hi | lo, 16 columns each) from the TMEM ring slot of this stage
                                const uint32_t at = tmem_base + (uint32_t)((c.bn + 31) & ~31) + (uint32_t)stage * 32u + (uint32_t)ks * 8u;
                                umma_tf32_ts(d, at, b_lo, idesc, (t | kb | ks) != 0 ? 1u : 0u);
                                umma_tf32_ts(d, at + 16u, b_hi, idesc, 1u);
                                umma_tf32_ts(d, at, b_hi, idesc, 1u);
                            } else {
                                umma_tf32(d, a_hi, b_lo, idesc, (t | kb | ks) != 0 ? 1u : 0u);
                                umma_tf32(d, a_lo, b_hi, idesc, 1u);
                                umma_tf32(d, a_hi, b_hi, idesc, 1u);
                            }
                        }
                    }
                }
                if (c.do_tma && elect_one()) {
                    if (c.cluster > 1) umma_commit_mc(smem_u32(empty_bar + stage), mask);
                    else umma_commit(smem_u32(empty_bar + stage));
                }
                __syncwarp();
                if (++stage == c.stages) { stage = 0; phase ^= 1; }
            }
        }
        if (elect_one()) umma_commit(smem_u32(done_bar));
        __syncwarp();
        mbar_wait(smem_u32(done_bar), 0);
        if (lane == 0) atomicMax(cycles + blockIdx.x, (unsigned long long)(clock64() - t0));
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(c.bn_mma);
            const int nh = c.bn / c.bn_mma;
            int stage = 0; uint32_t phase = 0;
            for (int t = 0; t < c.tiles; ++t) {
                for (int kb = 0; kb < c.kblocks; ++kb) {
                    if (c.do_tma) {
                        mbar_wait(smem_u32((c.do_xform ? ready_bar : full_bar) + stage), phase);
                        tc_fence_after();
                    }
                    const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
                    if (c.do_mma) {
                        const uint32_t bb = base + c.a_tiles * a_bytes;
#pragma unroll
                        for (int ks = 0; ks < BK / 8; ++ks) {
                            const uint64_t a_hi = make_smem_desc<BK>(base + ks * 32);
                            const uint64_t a_lo = make_smem_desc<BK>(base + (c.a_tiles > 1 ? a_bytes : 0) + ks * 32);
                            if (c.order == 0) {
                                for (int h = 0; h < nh; ++h) {
                                    const uint32_t off = (uint32_t)h * c.bn_mma * BK * 4;
                                    const uint64_t b_hi = make_smem_desc<BK>(bb + off + ks * 32);
                                    const uint64_t b_lo = make_smem_desc<BK>(bb + b_bytes + off + ks * 32);
                                    const uint32_t d = tmem_base + (uint32_t)h * c.bn_mma;
                                    umma_tf32(d, a_hi, b_lo, idesc, (t | kb | ks) != 0 ? 1u : 0u);
                                    umma_tf32(d, a_lo, b_hi, idesc, 1u);
                                    umma_tf32(d, a_hi, b_hi, idesc, 1u);
                                }
                            } else {
                                for (int j = 0; j < 3; ++j) {
                                    for (int h = 0; h < nh; ++h) {
                                        const uint32_t off = (uint32_t)h * c.bn_mma * BK * 4;
                                        const uint64_t b_hi = make_smem_desc<BK>(bb + off + ks * 32);
                                        const uint64_t b_lo = make_smem_desc<BK>(bb + b_bytes + off + ks * 32);
                                        const uint32_t d = tmem_base + (uint32_t)h * c.bn_mma;
                                        umma_tf32(d, j == 1 ? a_lo : a_hi, j == 0 ? b_lo : b_hi, idesc, (t | kb | ks | j) != 0 ? 1u : 0u);
                                    }
                                }
                            }
                        }
                        if (c.do_tma) {
                            if (c.cluster > 1) umma_commit_mc(smem_u32(empty_bar + stage), mask);
                            else umma_commit(smem_u32(empty_bar + stage));
                        }
                    } else if (c.do_tma) {
                        for (int r = 0; r < c.cluster; ++r) {
                            if (c.cluster > 1) mbar_arrive_remote(smem_u32(empty_bar + stage), r);
                            else mbar_arrive(smem_u32(empty_bar + stage));
                        }
                    }
                    if (++stage == c.stages) { stage = 0; phase ^= 1; }
                }
            }
            if (c.do_mma) {
                umma_commit(smem_u32(done_bar));
                mbar_wait(smem_u32(done_bar), 0);
            }
            atomicMax(cycles + blockIdx.x, (unsigned long long)(clock64() - t0));
        }
    } else if (c.do_xform == 2 && c.do_tma) {
        // y tile (one state tile, 64B-swizzled rows) -> registers -> hi | lo -> TMEM A ring slot (thread = row)
        const int q = warp & 3, row = q * 32 + lane;
        int stage = 0; uint32_t phase = 0;
        for (int t = 0; t < c.tiles; ++t) {
            for (int kb = 0; kb < c.kblocks; ++kb) {
                mbar_wait(smem_u32(full_bar + stage), phase);
                const uint8_t* tile = smem + (size_t)stage * stage_bytes + row * 64;
                uint32_t v[32];
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const float4 y = *reinterpret_cast<const float4*>(tile + ((ch ^ ((row >> 1) & 3)) << 4));
                    const float e[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t hi;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(e[j]));
                        uint32_t lo;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(e[j] - __uint_as_float(hi)));
                        v[ch * 4 + j] = hi;
                        v[16 + ch * 4 + j] = lo;
                    }
                }
                tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((c.bn + 31) & ~31) + (uint32_t)stage * 32u, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(ready_bar + stage));
                if (++stage == c.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (c.do_xform && c.do_tma) {
        const int xt = threadIdx.x - 64;
        constexpr int kVec = 128 * BK / 4;
        int stage = 0; uint32_t phase = 0;
        for (int t = 0; t < c.tiles; ++t) {
            for (int kb = 0; kb < c.kblocks; ++kb) {
                mbar_wait(smem_u32(full_bar + stage), phase);
                float4* t0p = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes);
                float4* t1p = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes + a_bytes);
#pragma unroll
                for (int i = xt; i < kVec; i += 128) {
                    const float4 y = t0p[i], yp = t1p[i];
                    float4 hi, lo;
                    hi.x = y.x + 0.5f * (y.x - yp.x); lo.x = y.x - hi.x;
                    hi.y = y.y + 0.5f * (y.y - yp.y); lo.y = y.y - hi.y;
                    hi.z = y.z + 0.5f * (y.z - yp.z); lo.z = y.z - hi.z;
                    hi.w = y.w + 0.5f * (y.w - yp.w); lo.w = y.w - hi.w;
                    t0p[i] = hi;
                    t1p[i] = lo;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(ready_bar + stage));
                if (++stage == c.stages) { stage = 0; phase ^= 1; }
            }
        }
    }
    __syncwarp();          // single-lane role loops: reconverge before the aligned CTA barrier
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(cycles + blockIdx.x, (unsigned long long)(clock64() - t0));
    if (c.cluster > 1) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// ---- candidate product-1 pipeline: split rings -------------------------------------------------------
// state tiles (y, HBM, ~2 us loaded latency) ride a deep ring of small slots, operator tiles (L2) a shallow ring
// of large slots; transform warps turn each y tile into hi | lo in a TMEM A ring; MMAs are the TS form.
struct Cfg2 { int bn, box_rows, a_stages, b_stages, t_stages, kblocks, tiles, cluster, m_tiles; };

__global__ void __launch_bounds__(256, 1)
ub2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
           const __grid_constant__ CUtensorMap tmB1, const Cfg2 c, unsigned long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int BK = 16;
    const uint32_t a_bytes = 128 * BK * 4, b_bytes = (uint32_t)c.bn * BK * 4, box_bytes = (uint32_t)c.box_rows * BK * 4;
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + (size_t)c.a_stages * a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)c.b_stages * 2 * b_bytes);
    uint64_t* afull = bars;
    uint64_t* aempty = afull + c.a_stages;
    uint64_t* bfull = aempty + c.a_stages;
    uint64_t* bempty = bfull + c.b_stages;
    uint64_t* ready = bempty + c.b_stages;
    uint64_t* tfree = ready + c.t_stages;
    uint64_t* done_bar = tfree + c.t_stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = c.cluster > 1 ? cluster_ctarank() : 0;
    const uint16_t mask = (uint16_t)((1u << c.cluster) - 1);
    const uint32_t a_col0 = (uint32_t)((c.bn + 31) & ~31);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < c.a_stages; ++s) { mbar_init(smem_u32(afull + s), 1); mbar_init(smem_u32(aempty + s), 4); }
        for (int s = 0; s < c.b_stages; ++s) { mbar_init(smem_u32(bfull + s), 1); mbar_init(smem_u32(bempty + s), c.cluster); }
        for (int s = 0; s < c.t_stages; ++s) { mbar_init(smem_u32(ready + s), 4); mbar_init(smem_u32(tfree + s), 1); }
        mbar_init(smem_u32(done_bar), 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    if (c.cluster > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long t0 = clock64();
    const int total = c.tiles * c.kblocks;
    const int nboxes = c.bn / c.box_rows;

    if (warp == 0) {            // operator producer
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < total; ++it) {
            const int kb = it % c.kblocks;
            mbar_wait(smem_u32(bempty + s), ph ^ 1);
            if (elect_one()) {
                const uint32_t fb = smem_u32(bfull + s);
                mbar_expect_tx(fb, 2 * b_bytes);
                const uint32_t bb = smem_u32(b_ring + (size_t)s * 2 * b_bytes);
                for (int bx = (int)rank; bx < nboxes; bx += c.cluster) {
                    if (c.cluster > 1) {
                        tma_load_2d_mc(bb + bx * box_bytes, &tmB0, kb * BK, bx * c.box_rows, fb, mask);
                        tma_load_2d_mc(bb + b_bytes + bx * box_bytes, &tmB1, kb * BK, bx * c.box_rows, fb, mask);
                    } else {
                        tma_load_2d(bb + bx * box_bytes, &tmB0, kb * BK, bx * c.box_rows, fb);
                        tma_load_2d(bb + b_bytes + bx * box_bytes, &tmB1, kb * BK, bx * c.box_rows, fb);
                    }
                }
            }
            __syncwarp();
            if (++s == c.b_stages) { s = 0; ph ^= 1; }
        }
    } else if (warp == 2) {     // state producer
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < total; ++it) {
            const int t = it / c.kblocks, kb = it % c.kblocks;
            const int m_tile = (int)((blockIdx.x + (long long)t * gridDim.x) % c.m_tiles);
            mbar_wait(smem_u32(aempty + s), ph ^ 1);
            if (elect_one()) {
                const uint32_t fb = smem_u32(afull + s);
                mbar_expect_tx(fb, a_bytes);
                tma_load_2d(smem_u32(a_ring + (size_t)s * a_bytes), &tmA, kb * BK, m_tile * 128, fb);
            }
            __syncwarp();
            if (++s == c.a_stages) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {     // MMA issuer
        const uint32_t idesc = make_idesc(c.bn);
        int s = 0; uint32_t ph = 0; int ts = 0; uint32_t tph = 0;
        for (int it = 0; it < total; ++it) {
            mbar_wait(smem_u32(bfull + s), ph);
            mbar_wait(smem_u32(ready + ts), tph);
            tc_fence_after();
            const uint32_t bb = smem_u32(b_ring + (size_t)s * 2 * b_bytes);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t b_hi = make_smem_desc<BK>(bb + ks * 32);
                    const uint64_t b_lo = make_smem_desc<BK>(bb + b_bytes + ks * 32);
                    const uint32_t at = tmem_base + a_col0 + (uint32_t)ts * 32u + (uint32_t)ks * 8u;
                    umma_tf32_ts(tmem_base, at, b_lo, idesc, (it | ks) != 0 ? 1u : 0u);
                    umma_tf32_ts(tmem_base, at + 16u, b_hi, idesc, 1u);
                    umma_tf32_ts(tmem_base, at, b_hi, idesc, 1u);
                }
                if (c.cluster > 1) umma_commit_mc(smem_u32(bempty + s), mask);
                else umma_commit(smem_u32(bempty + s));
                umma_commit(smem_u32(tfree + ts));
            }
            __syncwarp();
            if (++s == c.b_stages) { s = 0; ph ^= 1; }
            if (++ts == c.t_stages) { ts = 0; tph ^= 1; }
        }
        if (elect_one()) umma_commit(smem_u32(done_bar));
        __syncwarp();
        mbar_wait(smem_u32(done_bar), 0);
        if (lane == 0) atomicMax(cycles + blockIdx.x, (unsigned long long)(clock64() - t0));
    } else if (warp >= 4) {     // transform warps: thread = row
        const int q = warp & 3, row = q * 32 + lane;
        int s = 0; uint32_t ph = 0; int ts = 0; uint32_t tph = 0;
        for (int it = 0; it < total; ++it) {
            mbar_wait(smem_u32(afull + s), ph);
            const uint8_t* tile = a_ring + (size_t)s * a_bytes + row * 64;
            float4 y[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) y[ch] = *reinterpret_cast<const float4*>(tile + ((ch ^ ((row >> 1) & 3)) << 4));
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(aempty + s));
            uint32_t v[32];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const float e[4] = {y[ch].x, y[ch].y, y[ch].z, y[ch].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t hi, lo;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(e[j]));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(e[j] - __uint_as_float(hi)));
                    v[ch * 4 + j] = hi;
                    v[16 + ch * 4 + j] = lo;
                }
            }
            mbar_wait(smem_u32(tfree + ts), tph ^ 1);
            tc_fence_after();
            tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + a_col0 + (uint32_t)ts * 32u, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(ready + ts));
            if (++s == c.a_stages) { s = 0; ph ^= 1; }
            if (++ts == c.t_stages) { ts = 0; tph ^= 1; }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (c.cluster > 1) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static CUtensorMap make_map(const float* ptr, int k_elems, int rows, int ld, int box_rows, int bk) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)k_elems, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : bk == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensor map encode failed %d\n", (int)r); exit(1); }
    return m;
}

static float* dA0; static float* dA1; static float* dB0; static float* dB1;
static unsigned long long* dCyc;
static const int K = 2400, ROWS_A = 65536, ROWS_B = 512;

static void run(const char* name, Cfg c) {
    CUtensorMap a0 = make_map(dA0, K, ROWS_A, K, 128, c.bk), a1 = make_map(dA1, K, ROWS_A, K, 128, c.bk);
    CUtensorMap b0 = make_map(dB0, K, ROWS_B, K, c.box_rows, c.bk), b1 = make_map(dB1, K, ROWS_B, K, c.box_rows, c.bk);
    const size_t stage_bytes = (size_t)c.a_tiles * 128 * c.bk * 4 + 2 * (size_t)c.bn * c.bk * 4;
    auto kern = c.f16 ? ub_kernel<16, true> : c.bk == 32 ? ub_kernel<32> : c.bk == 8 ? ub_kernel<8> : ub_kernel<16>;
    const size_t smem = 1024 + c.stages * stage_bytes + (3 * c.stages + 2) * 8 + 16;
    if (smem > 232448) { printf("%-44s skipped (smem %zu)\n", name, smem); return; }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int sms = 148;
    cudaLaunchConfig_t lc = {};
    lc.blockDim = dim3(192); lc.dynamicSmemBytes = smem; lc.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int grid = sms / c.cluster * c.cluster;
    if (c.cluster > 1) {
        lc.gridDim = dim3(grid);
        int ncl = 0;
        CK(cudaOccupancyMaxActiveClusters(&ncl, kern, &lc));
        if (ncl * c.cluster < grid) grid = ncl * c.cluster;
    }
    lc.gridDim = dim3(grid);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemsetAsync(dCyc, 0, 1024 * 8));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&lc, kern, a0, a1, b0, b1, c, dCyc, (const float*)dB0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    std::vector<unsigned long long> cyc(grid);
    CK(cudaMemcpy(cyc.data(), dCyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long cmax = 0; for (auto v : cyc) if (v > cmax) cmax = v;
    const double kb_total = (double)c.tiles * c.kblocks;
    const double clk_per_kb = cmax / kb_total;
    // 2048 tf32 MAC / clk / SM; kind::f16: twice the K per instruction, util is quoted against 4096 fp16 MAC / clk / SM
    const double ideal_mma_clk = 3.0 * (c.bk / 8) * 128.0 * c.bn * 8 / 2048.0;
    const double bytes_per_kb_sm = c.a_tiles * 128 * 4.0 * c.bk + 2.0 * c.bn * 4.0 * c.bk / c.cluster;   // L2 -> SM fabric bytes
    const double smem_wr = c.a_tiles * 128 * 4.0 * c.bk + 2.0 * c.bn * 4.0 * c.bk;
    const double flops = 2.0 * 128 * c.bn * c.bk * kb_total * grid;                 // algorithmic (one product)
    printf("%-44s grid %3d  %7.3f ms  clk/kblock %7.1f  (mma ideal %6.1f -> util %5.1f%%)  L2->SM %5.1f B/clk/SM (chip %6.0f)  smem-fill %5.1f B/clk  alg %6.1f TF/s  eff-clock %4.0f MHz\n",
           name, grid, best, clk_per_kb, c.do_mma ? ideal_mma_clk : 0.0, c.do_mma ? 100.0 * ideal_mma_clk / clk_per_kb : 0.0,
           c.do_tma ? bytes_per_kb_sm / clk_per_kb : 0.0, c.do_tma ? bytes_per_kb_sm / clk_per_kb * grid : 0.0,
           c.do_tma ? smem_wr / clk_per_kb : 0.0, c.do_mma ? flops / (best * 1e-3) / 1e12 : 0.0, cmax / (best * 1e3));
    fflush(stdout);
}


static void run2(const char* name, Cfg2 c) {
    CUtensorMap a = make_map(dA0, K, ROWS_A, K, 128, 16);
    CUtensorMap b0 = make_map(dB0, K, ROWS_B, K, c.box_rows, 16), b1 = make_map(dB1, K, ROWS_B, K, c.box_rows, 16);
    const size_t smem = 1024 + (size_t)c.a_stages * 8192 + (size_t)c.b_stages * 2 * c.bn * 64 + (2 * c.a_stages + 2 * c.b_stages + 2 * c.t_stages + 2) * 8 + 16;
    if (smem > 232448) { printf("%-44s skipped (smem %zu)\n", name, smem); return; }
    CK(cudaFuncSetAttribute(ub2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(ub2_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t lc = {};
    lc.blockDim = dim3(256); lc.dynamicSmemBytes = smem; lc.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int grid = 148 / c.cluster * c.cluster;
    lc.gridDim = dim3(grid);
    if (c.cluster > 1) {
        int ncl = 0;
        CK(cudaOccupancyMaxActiveClusters(&ncl, ub2_kernel, &lc));
        if (ncl * c.cluster < grid) grid = ncl * c.cluster;
    }
    lc.gridDim = dim3(grid);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemsetAsync(dCyc, 0, 1024 * 8));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&lc, ub2_kernel, a, b0, b1, c, dCyc));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    std::vector<unsigned long long> cyc(grid);
    CK(cudaMemcpy(cyc.data(), dCyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long cmax = 0; for (auto v : cyc) if (v > cmax) cmax = v;
    const double kb_total = (double)c.tiles * c.kblocks;
    const double clk_per_kb = cmax / kb_total;
    const double ideal = 3.0 * 2 * 128.0 * c.bn * 8 / 2048.0;
    printf("%-44s grid %3d  %7.3f ms  clk/kblock %7.1f  (mma ideal %6.1f -> util %5.1f%%)  eff-clock %4.0f MHz\n", name, grid, best,
           clk_per_kb, ideal, 100.0 * ideal / clk_per_kb, cmax / (best * 1e3));
    fflush(stdout);
}

int main(int argc, char** argv) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    g_encode = reinterpret_cast<EncodeTiledFn>(p);
    CK(cudaMalloc(&dA0, (size_t)ROWS_A * K * 4)); CK(cudaMalloc(&dA1, (size_t)ROWS_A * K * 4));
    CK(cudaMalloc(&dB0, (size_t)ROWS_B * K * 4)); CK(cudaMalloc(&dB1, (size_t)ROWS_B * K * 4));
    CK(cudaMalloc(&dCyc, 1024 * 8));
    CK(cudaMemset(dA0, 0, (size_t)ROWS_A * K * 4)); CK(cudaMemset(dA1, 0, (size_t)ROWS_A * K * 4));
    CK(cudaMemset(dB0, 0, (size_t)ROWS_B * K * 4)); CK(cudaMemset(dB1, 0, (size_t)ROWS_B * K * 4));
    const int KB = K / 16, T = 4;
    auto f16_section = [&]() {
        printf("---- kind::f16 (GPAD_PREC_FP16X3): the same operand bytes per MMA, K = 16 (twice the MACs); clk per k-block of 6 MMAs\n");
        for (int f16 = 0; f16 < 2; ++f16) {
            for (int n : {128, 160, 192, 208, 256}) {
                char nm[64]; snprintf(nm, sizeof nm, "%s mma only SS 128x%d", f16 ? "f16 " : "tf32", n);
                run(nm, Cfg{n, n, n / 2, 2, 2, KB, T, 0, 1, 0, 1, 512, 0, 0, 16, 1, f16});
            }
            run(f16 ? "f16  mma only TS 128x208" : "tf32 mma only TS 128x208", Cfg{208, 208, 104, 1, 3, KB, T, 0, 1, 2, 1, 512, 0, 0, 16, 1, f16});
            run(f16 ? "f16  mma only TS 128x256" : "tf32 mma only TS 128x256", Cfg{256, 256, 128, 1, 3, KB, T, 0, 1, 2, 1, 512, 0, 0, 16, 1, f16});
            run(f16 ? "f16  mma only TS 2x208" : "tf32 mma only TS 2x208",     Cfg{416, 208, 104, 1, 3, KB, T, 0, 1, 2, 1, 512, 0, 0, 16, 1, f16});
        }
    };
    if (argc == 2 && !strcmp(argv[1], "f16")) { f16_section(); return 0; }
    if (argc > 1) {      // name bn bn_mma box a_tiles stages kblocks tiles tma mma xform cluster order warp_issue
        for (int i = 1; i + 13 < argc; i += 14) {
            Cfg c{atoi(argv[i + 1]), atoi(argv[i + 2]), atoi(argv[i + 3]), atoi(argv[i + 4]), atoi(argv[i + 5]), atoi(argv[i + 6]),
                  atoi(argv[i + 7]), atoi(argv[i + 8]), atoi(argv[i + 9]), atoi(argv[i + 10]), atoi(argv[i + 11]), 512, atoi(argv[i + 12]),
                  0, 16, atoi(argv[i + 13])};
            run(argv[i], c);
        }
        return 0;
    }
    for (int wi = 0; wi < 2; ++wi) {
        printf("---- issue loop: %s\n", wi ? "warp-uniform, elect.sync per MMA group" : "under if (lane == 0)");
        //                                  bn  bn_mma box a st  kb  tiles tma mma xf cl m_tiles order issue
        const int ns[] = {64, 128, 208, 240, 256};
        for (int n : ns) {
            char nm[64]; snprintf(nm, sizeof nm, "mma only 128x%d", n);
            run(nm, Cfg{n, n, n / 2, 2, 2, KB, T, 0, 1, 0, 1, 512, 0, 0, 16, wi});
        }
        run("mma only 2x128",                 Cfg{256, 128, 128, 2, 2, KB, T, 0, 1, 0, 1, 512, 0, 0, 16, wi});
        run("mma only 2x208",                 Cfg{416, 208, 104, 2, 2, KB, T, 0, 1, 0, 1, 512, 0, 0, 16, wi});
        run("mma only 2x256",                 Cfg{512, 256, 128, 2, 2, KB, T, 0, 1, 0, 1, 512, 0, 0, 16, wi});
        run("mma only 4x104",                 Cfg{416, 104, 104, 2, 2, KB, T, 0, 1, 0, 1, 512, 0, 0, 16, wi});
        run("tma+mma bn=208 (today, no xform)", Cfg{208, 208, 104, 2, 4, KB, T, 1, 1, 0, 1, 512, 0, 0, 16, wi});
        run("tma+mma+xform bn=208 (today)",   Cfg{208, 208, 104, 2, 4, KB, T, 1, 1, 1, 1, 512, 0, 0, 16, wi});
        run("tma+mma bn=416",                 Cfg{416, 208, 104, 2, 3, KB, T, 1, 1, 0, 1, 512, 0, 0, 16, wi});
        run("tma+mma+xform bn=416",           Cfg{416, 208, 104, 2, 3, KB, T, 1, 1, 1, 1, 512, 0, 0, 16, wi});
        run("tma+mma bn=416 cluster2",        Cfg{416, 208, 104, 2, 3, KB, T, 1, 1, 0, 2, 512, 0, 0, 16, wi});
        run("tma+mma+xform bn=416 cluster2",  Cfg{416, 208, 104, 2, 3, KB, T, 1, 1, 1, 2, 512, 0, 0, 16, wi});
        run("tma+mma bn=416 cluster4",        Cfg{416, 208, 104, 2, 3, KB, T, 1, 1, 0, 4, 512, 0, 0, 16, wi});
        // product 2 shapes: K = 416 (26 k-blocks), ten operator tiles of 240 per state tile
        run("p2: mma only bn=240",            Cfg{240, 240, 120, 2, 3, 26, 40, 0, 1, 0, 1, 512, 0, 0, 16, wi});
        run("p2: tma+mma bn=240",             Cfg{240, 240, 120, 2, 3, 26, 40, 1, 1, 0, 1, 512, 0, 0, 16, wi});
        run("p2: tma+mma bn=240 cluster2",    Cfg{240, 240, 120, 2, 3, 26, 40, 1, 1, 0, 2, 512, 0, 0, 16, wi});
        run("p2: tma+mma bn=480",             Cfg{480, 240, 120, 2, 2, 26, 20, 1, 1, 0, 1, 512, 0, 0, 16, wi});
        run("p2: tma+mma bn=480 cluster2",    Cfg{480, 240, 120, 2, 2, 26, 20, 1, 1, 0, 2, 512, 0, 0, 16, wi});
    }
    f16_section();
    printf("---- product 1 candidates: A = y only (one state tile), hi/lo built by transform warps into a TMEM A ring (TS MMAs)\n");
    run("y-only smem-xform bn=208 (A2 layout)", Cfg{208, 208, 104, 2, 3, KB, T, 1, 1, 1, 1, 512, 0, 0, 16, 1});
    run("y-only tmem-A bn=208",            Cfg{208, 208, 104, 1, 3, KB, T, 1, 1, 2, 1, 512, 0, 0, 16, 1});
    run("y-only tmem-A bn=208 cluster2",   Cfg{208, 208, 104, 1, 3, KB, T, 1, 1, 2, 2, 512, 0, 0, 16, 1});
    run("mma only TS 128x208",             Cfg{208, 208, 104, 1, 3, KB, T, 0, 1, 2, 1, 512, 0, 0, 16, 1});
    run("mma only TS 2x208",               Cfg{416, 208, 104, 1, 3, KB, T, 0, 1, 2, 1, 512, 0, 0, 16, 1});
    run("y-only tmem-A bn=208 4 stages",   Cfg{208, 208, 104, 1, 4, KB, T, 1, 1, 2, 1, 512, 0, 0, 16, 1});
    run("y-only tmem-A bn=208 5 stages",   Cfg{208, 208, 104, 1, 5, KB, T, 1, 1, 2, 1, 512, 0, 0, 16, 1});
    run("y-only tmem-A bn=208 5 st cluster2", Cfg{208, 208, 104, 1, 5, KB, T, 1, 1, 2, 2, 512, 0, 0, 16, 1});
    run("y-only tmem-A bn=240 5 stages",   Cfg{240, 240, 120, 1, 5, KB, T, 1, 1, 2, 1, 512, 0, 0, 16, 1});
    run("y-only tmem-A bn=240 5 st cluster2", Cfg{240, 240, 120, 1, 5, KB, T, 1, 1, 2, 2, 512, 0, 0, 16, 1});
    run("tma+mma (no xform) bn=208 a1 5 st", Cfg{208, 208, 104, 1, 5, KB, T, 1, 1, 0, 1, 512, 0, 0, 16, 1});
    run("tma+mma (no xform) bn=208 a1 5 st cl2", Cfg{208, 208, 104, 1, 5, KB, T, 1, 1, 0, 2, 512, 0, 0, 16, 1});
    run("y-only tmem-A bn=416 (2x208)",    Cfg{416, 208, 104, 1, 3, KB, T, 1, 1, 2, 1, 512, 0, 0, 16, 1});
    run("y-only tmem-A bn=416 cluster2",   Cfg{416, 208, 104, 1, 3, KB, T, 1, 1, 2, 2, 512, 0, 0, 16, 1});
    run("y-only tmem-A bn=416 cluster4",   Cfg{416, 208, 104, 1, 3, KB, T, 1, 1, 2, 4, 512, 0, 0, 16, 1});
    printf("---- split rings: deep state ring (HBM), shallow operator ring (L2), TMEM A ring\n");
    //                                         bn  box  a_st b_st t_st kb  tiles cl m_tiles
    run2("split bn=208 a4 b3 t3",         Cfg2{208, 104, 4, 3, 3, KB, T, 1, 512});
    run2("split bn=208 a8 b3 t3",         Cfg2{208, 104, 8, 3, 3, KB, T, 1, 512});
    run2("split bn=208 a8 b4 t3",         Cfg2{208, 104, 8, 4, 3, KB, T, 1, 512});
    run2("split bn=208 a8 b5 t3",         Cfg2{208, 104, 8, 5, 3, KB, T, 1, 512});
    run2("split bn=208 a6 b5 t3",         Cfg2{208, 104, 6, 5, 3, KB, T, 1, 512});
    run2("split bn=208 a8 b3 t4",         Cfg2{208, 104, 8, 3, 4, KB, T, 1, 512});
    run2("split bn=208 a8 b4 t4",         Cfg2{208, 104, 8, 4, 4, KB, T, 1, 512});
    run2("split bn=208 a12 b4 t6",        Cfg2{208, 104, 12, 4, 6, KB, T, 1, 512});
    run2("split bn=208 a12 b4 t6 cluster2", Cfg2{208, 104, 12, 4, 6, KB, T, 2, 512});
    run2("split bn=208 a8 b5 t8",         Cfg2{208, 104, 8, 5, 8, KB, T, 1, 512});
    run2("split bn=240 a8 b4 t6",         Cfg2{240, 120, 8, 4, 6, KB, T, 1, 512});
    run2("split bn=240 a8 b4 t6 cluster2", Cfg2{240, 120, 8, 4, 6, KB, T, 2, 512});
    run2("split bn=256 a8 b4 t6",         Cfg2{256, 128, 8, 4, 6, KB, T, 1, 512});
    run2("split bn=256 a8 b4 t6 cluster2", Cfg2{256, 128, 8, 4, 6, KB, T, 2, 512});
    printf("---- 128 B rows (BK = 32, SWIZZLE_128B) against 64 B rows (BK = 16)\n");
    run("bk32 tma only bn=208 A2",          Cfg{208, 208, 104, 2, 2, K / 32, T, 1, 0, 0, 1, 512, 0, 0, 32, 1});
    run("bk32 tma only bn=416 A2 (2 st)",   Cfg{416, 208, 104, 2, 1, K / 32, T, 1, 0, 0, 1, 512, 0, 0, 32, 1});
    run("bk32 tma+mma bn=208 A2 2 st",      Cfg{208, 208, 104, 2, 2, K / 32, T, 1, 1, 0, 1, 512, 0, 0, 32, 1});
    run("bk32 tma+mma+xform bn=208 A2 2 st", Cfg{208, 208, 104, 2, 2, K / 32, T, 1, 1, 1, 1, 512, 0, 0, 32, 1});
    run("bk32 tma+mma bn=208 a1 3 st",      Cfg{208, 208, 104, 1, 3, K / 32, T, 1, 1, 0, 1, 512, 0, 0, 32, 1});
    printf("---- operator k-blocks as contiguous pre-packed images, one cp.async.bulk each (no per-row TMA requests)\n");
    run("p2 bn=240 A2 3 st, tensor-map B",  Cfg{240, 240, 120, 2, 3, 26, 40, 1, 1, 0, 1, 512, 0, 0, 16, 1});
    run("p2 bn=240 A2 3 st, bulk B",        Cfg{240, 240, 120, 2, 3, 26, 40, 1, 1, 0, 1, 512, 0, 1, 16, 1});
    run("p2 bn=240 A2 3 st, bulk B, tma only", Cfg{240, 240, 120, 2, 3, 26, 40, 1, 0, 0, 1, 512, 0, 1, 16, 1});
    run("p2 bn=240 A2 3 st, tensor-map B, tma only", Cfg{240, 240, 120, 2, 3, 26, 40, 1, 0, 0, 1, 512, 0, 0, 16, 1});
    run("p1 bn=208 A2 4 st, tensor-map B",  Cfg{208, 208, 104, 2, 4, KB, T, 1, 1, 0, 1, 512, 0, 0, 16, 1});
    run("p1 bn=208 A2 4 st, bulk B",        Cfg{208, 208, 104, 2, 4, KB, T, 1, 1, 0, 1, 512, 0, 1, 16, 1});
    printf("---- same with the state tiles L2 resident (8 distinct batch tiles), as zhat is in product 2\n");
    run("p2 L2-A bn=240 3 st, tensor-map B",  Cfg{240, 240, 120, 2, 3, 26, 40, 1, 1, 0, 1, 8, 0, 0, 16, 1});
    run("p2 L2-A bn=240 3 st, bulk B",        Cfg{240, 240, 120, 2, 3, 26, 40, 1, 1, 0, 1, 8, 0, 1, 16, 1});
    run("p2 L2-A bn=240 3 st, tensor-map B, tma only", Cfg{240, 240, 120, 2, 3, 26, 40, 1, 0, 0, 1, 8, 0, 0, 16, 1});
    run("p2 L2-A bn=240 3 st, bulk B, tma only", Cfg{240, 240, 120, 2, 3, 26, 40, 1, 0, 0, 1, 8, 0, 1, 16, 1});
    run("p2 L2-A bn=240 mma only",            Cfg{240, 240, 120, 2, 3, 26, 40, 0, 1, 0, 1, 8, 0, 0, 16, 1});
    printf("---- 32 B rows (BK = 8, SWIZZLE_32B): does the MMA read whole 128 B shared-memory lines?\n");
    run("bk8  tma+mma bn=240 A2 6 st (p2)", Cfg{240, 240, 120, 2, 6, 52, 40, 1, 1, 0, 1, 512, 0, 0, 8, 1});
    run("bk8  tma+mma bn=240 A2 8 st (p2)", Cfg{240, 240, 120, 2, 8, 52, 40, 1, 1, 0, 1, 512, 0, 0, 8, 1});
    run("bk8  tma only bn=240 A2 8 st",     Cfg{240, 240, 120, 2, 8, 52, 40, 1, 0, 0, 1, 512, 0, 0, 8, 1});
    run("bk16 tma+mma bn=240 A2 3 st (p2)", Cfg{240, 240, 120, 2, 3, 26, 40, 1, 1, 0, 1, 512, 0, 0, 16, 1});
    run("bk16 tma+mma bn=240 A2 4 st (p2)", Cfg{240, 240, 120, 2, 4, 26, 40, 1, 1, 0, 1, 512, 0, 0, 16, 1});
    run("bk8  tma+mma bn=208 A2 8 st",      Cfg{208, 208, 104, 2, 8, K / 8, T, 1, 1, 0, 1, 512, 0, 0, 8, 1});
    run("bk32 tma+mma bn=240 A2 2 st (p2)", Cfg{240, 240, 120, 2, 2, 13, 40, 1, 1, 0, 1, 512, 0, 0, 32, 1});
    run("bk16 tma+mma bn=208 A2 2 st",      Cfg{208, 208, 104, 2, 2, K / 16, T, 1, 1, 0, 1, 512, 0, 0, 16, 1});
    run("bk16 tma+mma bn=208 A2 4 st",      Cfg{208, 208, 104, 2, 4, K / 16, T, 1, 1, 0, 1, 512, 0, 0, 16, 1});
    run("tma only A only (bn=16)",        Cfg{16, 16, 16, 2, 8, KB, T, 1, 0, 0, 1, 512, 0, 0, 16, 0});
    run("tma only bn=208 A2",             Cfg{208, 208, 104, 2, 4, KB, T, 1, 0, 0, 1, 512, 0, 0, 16, 0});
    run("tma only bn=416 A2",             Cfg{416, 208, 104, 2, 3, KB, T, 1, 0, 0, 1, 512, 0, 0, 16, 0});
    return 0;
}
