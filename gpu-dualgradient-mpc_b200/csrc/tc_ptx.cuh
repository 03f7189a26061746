// tc_ptx.cuh -- inline-PTX wrappers shared by the tcgen05 kernels (batch_tc.cu, batch_tc_p1.cu): mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 alloc / mma / commit / ld, programmatic dependent launch, UMMA shared-memory and
// instruction descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace gpad {
namespace tc {

constexpr int kBM = 128;            // batch rows per CTA tile == UMMA M per CTA
constexpr int kAccStride = 256;     // TMEM columns per accumulator stage
constexpr int kXformWarps = 4;      // product 1 only
constexpr int kWorkWarps = 12;      // warps 2..13: 12 epilogue warps, or 4 transform + 8 epilogue
constexpr int kThreads = 32 * (2 + kWorkWarps);
constexpr int kEpiBufFloats = 32 * 33;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one elected lane of a fully converged warp (always the same lane while all 32 are active).  The single-thread
// tcgen05 / TMA instructions are issued under this predicate from warp-UNIFORM loops: a role loop that runs
// entirely under `if (lane == 0)` makes ptxas serialise every descriptor through a per-thread R2UR loop, which
// measured ~145 clk per tcgen05.mma (tests/ubench/ubench_tc.cu) -- above the 104-128 clk the MMA itself needs.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 (fp16 operands, fp32 accumulate): K = 16 per instruction, twice the MACs of a kind::tf32 instruction for
// the same operand bytes
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (TS form)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (F16) umma_f16(d, a, b, idesc, acc); else umma_tf32(d, a, b, idesc, acc);
}
template <bool F16>
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (F16) umma_f16_ts(d, a, b, idesc, acc); else umma_tf32_ts(d, a, b, idesc, acc);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor of a K-major operand tile whose rows are BK floats wide and
// hardware-swizzled with a span of one row (BK = 32: SWIZZLE_128B, BK = 16: SWIZZLE_64B);
// canonical layout ((8,rows/8),(T,2)) : ((row bytes, SBO),(1, LBO)) -- cute mma_sm100_desc.hpp
template <int BK>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    constexpr uint64_t row_bytes = BK * 4;
    constexpr uint64_t sbo = 8 * row_bytes;                     // next group of 8 rows
    constexpr uint64_t layout = (BK == 32) ? 2 : (BK == 16 ? 4 : 6);  // SWIZZLE_128B / 64B / 32B
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                    // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                     // LBO (unused for swizzled K-major)
    d |= (sbo >> 4) << 32;                                      // SBO, bits [32,46)
    d |= (uint64_t)1 << 46;                                     // descriptor version (Blackwell)
    d |= layout << 61;                                          // swizzle mode, bits [61,64)
    return d;
}

// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc(int bn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
// same for kind::f16 with fp16 operands (a_format = b_format = 0)
__device__ __forceinline__ uint32_t make_idesc_f16(int bn) {
    return (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
template <bool F16>
__device__ __forceinline__ uint32_t make_idesc_k(int bn) { return F16 ? make_idesc_f16(bn) : make_idesc(bn); }

// ---- GPAD_PREC_FP16X3: x * 2^e = hi + lo with fp16 hi, lo (11 + 11 significant bits, what the tf32 split keeps) ----
// The power-of-two scale brings the largest magnitude of a row into [2^14, 2^15): entries down to 2^-18 of it keep
// the full 22 bits, smaller ones an absolute error of 2^-40 of the row maximum (fp16 subnormal spacing 2^-24).
// Scaling by powers of two is exact, and so is undoing it on the fp32 accumulator.
__device__ __forceinline__ int f16_scale_exp(float row_max) {
    // row_max >= 0.  exponent field of row_max -> e with row_max * 2^e in [2^14, 2^15); zero / subnormal / non-finite
    // maxima (nothing to scale, or a row that has already diverged) get e = 0
    const int ex = (int)((__float_as_uint(row_max) >> 23) & 0xffu);
    if (ex == 0 || ex == 255) return 0;
    int e = 14 - (ex - 127);
    return max(-100, min(100, e));
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((uint32_t)(127 + e) << 23); }
// two scaled fp32 values -> packed fp16x2 hi and lo (low half = first element)
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    float h0, h1;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    const float r0 = x0 - h0, r1 = x1 - h1;          // exact in fp32
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}


// programmatic dependent launch: wait until the kernels this one depends on have completed and their memory is
// visible (a no-op when the launch carried no programmatic attribute) / allow the next kernel in the stream to be
// scheduled as SMs free up
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace tc
}  // namespace gpad
