// lat_util.cuh -- small device helpers shared by the latency-mode kernels
#pragma once
#include <cuda_runtime.h>

namespace gpad {
namespace lat {

// sums R per-lane values over the 32 lanes of a warp with R - 1 + log2(32 / R) shuffles instead of 5 R: every
// step pairs lanes that differ in one bit, each keeps one half of the values and hands over the other half.
// On return lane l holds the warp total of value (l >> log2(32 / R)) & (R - 1).
template <int R>
__device__ __forceinline__ float warp_sum_transposed(float (&v)[R], int lane) {
    static_assert(R == 8 || R == 16 || R == 32, "R");
    int o = 16;
#pragma unroll
    for (int s = R / 2; s >= 1; s >>= 1, o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    float t = v[0];
#pragma unroll
    for (; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}
// the same reduction with the shuffles as volatile asm: ptxas keeps volatile statements in program order, so every
// shuffle of one level is issued before the first of the next (breadth-first).  Left to itself ptxas sometimes walks
// the tree depth-first to save registers, which puts ~4 R dependent shuffles on the critical path of a lone warp.
__device__ __forceinline__ float shfl_bfly_v(float x, int o) {
    float r;
    asm volatile("shfl.sync.bfly.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=f"(r) : "f"(x), "r"(o));
    return r;
}
__device__ __forceinline__ float shfl_idx_v(float x, int src) {
    float r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=f"(r) : "f"(x), "r"(src));
    return r;
}
template <int R>
__device__ __forceinline__ float warp_sum_transposed_v(float (&v)[R], int lane) {
    static_assert(R == 8 || R == 16 || R == 32, "R");
    int o = 16;
#pragma unroll
    for (int s = R / 2; s >= 1; s >>= 1, o >>= 1) {
        const bool up = (lane & o) != 0;
        float got[R / 2];
#pragma unroll
        for (int i = 0; i < s; ++i) got[i] = shfl_bfly_v(up ? v[i] : v[i + s], o);
#pragma unroll
        for (int i = 0; i < s; ++i) v[i] = (up ? v[i + s] : v[i]) + got[i];
    }
    float t = v[0];
#pragma unroll
    for (; o >= 1; o >>= 1) t += shfl_bfly_v(t, o);
    return t;
}
// The same reduction WITHOUT the selects: when lane l keeps its R values in the order v'[k] = v[k ^ mask(l)], with mask(l) the
// lane bits the butterfly pairs on (for R = 16 over 32 lanes: mask = (l >> 1) & 15), the element to hand over is at a fixed
// index at every level -- the partner's mask differs in exactly the bit that swaps the halves.  The kernel gets the permuted
// order for free by loading its operator rows in that order once.  Same pairing, hence the same sums bit for bit; lane l
// ends with the total of value mask(l).  ORD: shuffles as volatile asm (program order), see above.
template <int R, bool ORD>
__device__ __forceinline__ float warp_sum_prepermuted(float (&v)[R]) {
    static_assert(R == 8 || R == 16 || R == 32, "R");
    int o = 16;
#pragma unroll
    for (int s = R / 2; s >= 1; s >>= 1, o >>= 1) {
        float got[R / 2];
#pragma unroll
        for (int i = 0; i < s; ++i) got[i] = ORD ? shfl_bfly_v(v[i + s], o) : __shfl_xor_sync(0xffffffffu, v[i + s], o);
#pragma unroll
        for (int i = 0; i < s; ++i) v[i] = v[i] + got[i];
    }
    float t = v[0];
#pragma unroll
    for (; o >= 1; o >>= 1) t += ORD ? shfl_bfly_v(t, o) : __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}
// the same over HALF a warp (two independent problems per warp, 16 lanes each): R = 16 values, four levels, no tail;
// lane h = lane & 15 keeps its values in the order v'[k] = v[k ^ h] and ends with the total of value h
template <bool ORD>
__device__ __forceinline__ float halfwarp_sum_prepermuted(float (&v)[16]) {
    int o = 8;
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1, o >>= 1) {
        float got[8];
#pragma unroll
        for (int i = 0; i < s; ++i) got[i] = ORD ? shfl_bfly_v(v[i + s], o) : __shfl_xor_sync(0xffffffffu, v[i + s], o);
#pragma unroll
        for (int i = 0; i < s; ++i) v[i] = v[i] + got[i];
    }
    return v[0];
}
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }


}  // namespace lat
}  // namespace gpad
