// tc_epilogue.cuh -- the fused GPAD epilogue of one transposed 32x32 accumulator block, shared by the two
// tcgen05 kernels (batch_tc.cu, batch_tc_p1.cu).  `buf` holds the block transposed in shared memory
// (buf[row * 33 + col]); lane = output column, rows row_base..row_base+31 of the batch.
#pragma once
#include "batch_common.cuh"

namespace gpad {
namespace tc {

// fast path (every iteration of a fixed-iteration solve; in tolerance mode the iterations between two checks): rows in
// chunks of kChunk with every global load of the chunk issued before the first use, so each warp keeps
// kChunk * (2 or 3) x 128 B in flight; streaming cache hints keep the operators in L2.
// TOL = false is the fixed-iteration code (nothing but the three operand loads and the store per element: its load
// batching is what the kernel's speed hangs on -- a predicate that depends on a loaded flag inside the load loop cost
// product 2 10-20 %, measured).  TOL = true (tolerance mode) first gathers which rows still run (done[b]), skips the
// stopped ones and carries the averaged residual sbar along.
template <int PHASE, bool TOL>
__device__ __forceinline__ void fast_rows(const BatchKernelArgs& args, const float* buf, int lane, int row_base, int c, bool col_ok) {
    constexpr int kChunk = 16;
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += kChunk) {
        unsigned live = 0;
        if (TOL) {
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const int b = row_base + r0 + j;
                live |= ((col_ok && b < args.B && __ldg(args.done + min(b, args.B - 1)) == 0) ? 1u : 0u) << j;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kChunk; ++j) live |= ((col_ok && row_base + r0 + j < args.B) ? 1u : 0u) << j;
        }
        if (PHASE == 1) {
            float gp[kChunk], zo[kChunk], pp[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const bool ok = (live >> j) & 1u;
                const size_t o = (size_t)(row_base + r0 + j) * args.np + c;
                gp[j] = ok ? __ldcs(args.g_P + o) : 0.f;
                zo[j] = ok ? __ldcs(args.z + o) : 0.f;
                pp[j] = ok ? __ldcs(args.P_prev + o) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                if (!((live >> j) & 1u)) continue;
                const size_t o = (size_t)(row_base + r0 + j) * args.np + c;
                float acc = buf[(r0 + j) * 33 + lane];
                __stcs(args.P_cur + o, acc);
                acc = momentum(acc, pp[j], args.it.beta);          // M_G w_v from P_v, P_{v-1}
                const float zh = acc - gp[j];
                __stcs(args.z + o, __fadd_rn(__fmul_rn(1.0f - args.it.theta, zo[j]), __fmul_rn(args.it.theta, zh)));
                if (args.it.store_zhat) __stcs(args.zhat + o, zh);
                float hi, lo;
                split_tf32(zh, hi, lo);
                args.zh_hi[o] = hi;      // re-read by product 2 of this iteration: default caching
                args.zh_lo[o] = lo;
            }
        } else {
            float yc[kChunk], yp[kChunk], pd[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const bool ok = (live >> j) & 1u;
                const size_t o = (size_t)(row_base + r0 + j) * args.mp + c;
                yc[j] = ok ? __ldcs(args.y_cur + o) : 0.f;
                yp[j] = ok ? __ldcs(args.y_prev + o) : 0.f;
                pd[j] = ok ? __ldcs(args.p_D + o) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                if (!((live >> j) & 1u)) continue;
                const size_t o = (size_t)(row_base + r0 + j) * args.mp + c;
                const float wv = momentum(yc[j], yp[j], args.it.beta);
                const float sacc = buf[(r0 + j) * 33 + lane] + (wv + pd[j]);
                args.y_next[o] = 0.5f * (sacc + fabsf(sacc));   // read back by the next two kernels
            }
            if (TOL) {
                // sbar <- (1 - theta) sbar + theta (acc + p_D), the residual of the averaged iterate
                float sb[kChunk];
#pragma unroll
                for (int j = 0; j < kChunk; ++j)
                    sb[j] = ((live >> j) & 1u) ? __ldcs(args.sbar + (size_t)(row_base + r0 + j) * args.mp + c) : 0.f;
#pragma unroll
                for (int j = 0; j < kChunk; ++j) {
                    if (!((live >> j) & 1u)) continue;
                    const float rhat = buf[(r0 + j) * 33 + lane] + pd[j];
                    __stcs(args.sbar + (size_t)(row_base + r0 + j) * args.mp + c,
                           __fadd_rn(__fmul_rn(1.0f - args.it.theta, sb[j]), __fmul_rn(args.it.theta, rhat)));
                }
            }
        }
    }
}

// GPAD_PREC_FP16X3 (fixed-iteration solves), product 1: the accumulator block arrives with the A row scale already
// undone (the epilogue warp applies it while transposing: TMEM lane = batch row); `cinv` undoes the operator row's
// scale (lane = output column).  zhat is stored in fp32 only -- its fp16 split needs the maximum of the WHOLE row,
// which no single tile sees, and is made by quantize_rows_kernel between the products (batch_f16.cu).  Product 2 of
// this precision has its own kernel and epilogue (batch_tc_p2.cu).
template <int PHASE>
__device__ __forceinline__ void fast_rows_f16(const BatchKernelArgs& args, const float* buf, int lane, int row_base, int c, bool col_ok,
                                              float cinv) {
    static_assert(PHASE == 1, "fp16 product 2 runs batch_tc_p2.cu");
    constexpr int kChunk = 16;
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += kChunk) {
        unsigned live = 0;
#pragma unroll
        for (int j = 0; j < kChunk; ++j) live |= ((col_ok && row_base + r0 + j < args.B) ? 1u : 0u) << j;
        if (PHASE == 1) {
            float gp[kChunk], zo[kChunk], pp[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const bool ok = (live >> j) & 1u;
                const size_t o = (size_t)(row_base + r0 + j) * args.np + c;
                gp[j] = ok ? __ldcs(args.g_P + o) : 0.f;
                zo[j] = ok ? __ldcs(args.z + o) : 0.f;
                pp[j] = ok ? __ldcs(args.P_prev + o) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                if (!((live >> j) & 1u)) continue;
                const size_t o = (size_t)(row_base + r0 + j) * args.np + c;
                float acc = buf[(r0 + j) * 33 + lane] * cinv;
                __stcs(args.P_cur + o, acc);
                acc = momentum(acc, pp[j], args.it.beta);          // M_G w_v from P_v, P_{v-1}
                const float zh = acc - gp[j];
                __stcs(args.z + o, __fadd_rn(__fmul_rn(1.0f - args.it.theta, zo[j]), __fmul_rn(args.it.theta, zh)));
                args.zhat[o] = zh;       // re-read by zsplit_kernel right after this launch: default caching
            }
        }
    }
}

// TOL = false: the kernel of fixed-iteration solves carries nothing but the fast path (the epilogue code of the
// tolerance mode -- stopped rows, residual average, reductions, dual-gap launches -- lives in the TOL = true
// instantiation: a 30 % larger kernel measured 10-17 % slower on product 2, whose 14 warps run four different roles
// out of one instruction cache)
template <int PHASE, bool TOL, bool F16 = false>
__device__ __forceinline__ void epilogue_block(const BatchKernelArgs& args, const float* buf, int lane, int row_base, int blk,
                                               int n_tile, int bn, int ncols_valid, float* __restrict__ Cdbg, int ldc,
                                               int step = 0) {
    // tiles start every `step` columns (default: bn).  step < bn (a multiple of 32) keeps every 32-column block on a
    // 128-byte line; the bn - step columns a tile shares with its predecessor belong to the predecessor
    const int cin = blk * 32 + lane;            // column inside the tile
    const int stp = step > 0 ? step : bn;
    const int c = n_tile * stp + cin;      // global output column
    const bool col_ok = cin < bn && c < ncols_valid && (n_tile == 0 || cin >= bn - stp);
    float cinv = 1.f;
    if (F16) cinv = col_ok ? __ldg(args.b_colinv + c) : 0.f;
    if (PHASE == 0) {
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
            const int b = row_base + rr;
            if (col_ok && b < args.B) Cdbg[(size_t)b * ldc + c] = F16 ? buf[rr * 33 + lane] * cinv : buf[rr * 33 + lane];
        }
    } else if (PHASE == 1 && args.p_only) {
        // warm start: P_{-1} = M_G y_{-1}, nothing else
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
            const int b = row_base + rr;
            if (col_ok && b < args.B) args.P_cur[(size_t)b * args.np + c] = F16 ? buf[rr * 33 + lane] * cinv : buf[rr * 33 + lane];
        }
    } else if (F16) {
        if constexpr (PHASE == 1) fast_rows_f16<1>(args, buf, lane, row_base, c, col_ok, cinv);
    } else if (!TOL) {
        fast_rows<PHASE, false>(args, buf, lane, row_base, c, col_ok);
    } else if (!args.it.check && !args.dual && args.done) {
        fast_rows<PHASE, true>(args, buf, lane, row_base, c, col_ok);
    } else {
        // general path: termination bookkeeping (per-row reductions, stopped instances)
#pragma unroll 2
        for (int rr = 0; rr < 32; ++rr) {
            const int b = row_base + rr;
            const float val = buf[rr * 33 + lane];
            bool ok = col_ok && b < args.B;
            if (ok && args.done) ok = args.done[b] == 0;
            if (PHASE == 1) {
                float f_zhat = 0.f;
                if (ok) {
                    float acc = val;
                    if (!args.dual) {
                        const size_t o = (size_t)b * args.np + c;
                        args.P_cur[o] = val;
                        acc = momentum(val, args.P_prev[o], args.it.beta);
                    }
                    epilogue1<true>(args, b, c, acc, f_zhat);
                }
                if (args.it.check && args.f) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) f_zhat += __shfl_xor_sync(0xffffffffu, f_zhat, o);
                    if (lane == 0 && b < args.B) atomicAdd(args.red + (size_t)b * kRedStride + 5, f_zhat);
                }
            } else {
                Red2 red;
                if (ok) epilogue2(args, b, c, val, red);
                if (args.it.check) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        red.max_sbar = fmaxf(red.max_sbar, __shfl_xor_sync(0xffffffffu, red.max_sbar, o));
                        red.max_rhat = fmaxf(red.max_rhat, __shfl_xor_sync(0xffffffffu, red.max_rhat, o));
                        red.min_w = fminf(red.min_w, __shfl_xor_sync(0xffffffffu, red.min_w, o));
                        red.w_rhat += __shfl_xor_sync(0xffffffffu, red.w_rhat, o);
                        red.w_dot += __shfl_xor_sync(0xffffffffu, red.w_dot, o);
                        red.bad = fmaxf(red.bad, __shfl_xor_sync(0xffffffffu, red.bad, o));
                    }
                    if (lane == 0 && b < args.B && !(args.done && args.done[b])) flush_red2(args, b, red);
                }
            }
        }
    }
}

}  // namespace tc
}  // namespace gpad
