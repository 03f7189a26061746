// batch_compact.cu -- tolerance mode of the shared-operator batch: gathering the instances that still run into dense
// batch tiles.  Instances stop at very different iterations (quadrotor, eps = 1e-3: 40 ... 3300, BASELINE config 4), and a
// 128-row tile keeps riding every MMA until its slowest instance stops; tile retirement alone removes almost nothing
// when instances are spread at random.  So whenever a fifth of the working rows has stopped, the batch is compacted:
//   plan     (one block)   n_run = running rows; holes = stopped rows below n_run, movers = running rows at or above it
//   archive  (block / row) every stopped row's outputs (the three rotating y buffers, z, zhat, iters, status, ...) go to
//                          the archive at the instance's ORIGINAL index perm[row]; the row is then dead (perm = -1)
//   move     (block / move) mover i -> hole i: all per-row state (p_D, g_P, f, y x3, z, zhat, sbar, P x2, reductions,
//                          bookkeeping) and its perm entry; the vacated row is marked stopped
// after which rows [0, n_run) all run and the tile list shrinks to ceil(n_run / 128) tiles.  Rows are independent in
// both products, so an instance's iterates do not depend on the row it occupies: results are bit-identical to the
// uncompacted solve (tests/test_gpu_parity.py).  Everything is enqueued on the solve's stream; the host only decides
// WHEN, from the stop counts it already receives.
#include "batch_common.cuh"
#include "gpad_internal.h"

namespace gpad {

namespace {

// block-wide exclusive scan helper over chunks of 1024 flags; returns the running total in *base (shared)
__device__ __forceinline__ int block_scan_1024(int flag, int* warp_sum, int* base, int& total_before) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int ws = warp_sum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += v; }
        warp_sum[lane] = ws;
    }
    __syncthreads();
    total_before = *base + (warp ? warp_sum[warp - 1] : 0) + incl - flag;
    const int chunk_total = warp_sum[31];
    __syncthreads();
    if (threadIdx.x == 0) *base += chunk_total;
    __syncthreads();
    return chunk_total;
}

// counts[0] = running rows, counts[1] = moves; holes / movers ascending
__global__ void __launch_bounds__(1024)
compact_plan_kernel(int rows, const int* __restrict__ done, const int* __restrict__ perm, int* __restrict__ holes,
                    int* __restrict__ movers, int* __restrict__ counts) {
    __shared__ int warp_sum[32];
    __shared__ int base;
    __shared__ int n_run_s;
    // pass 1: running rows
    int mine = 0;
    for (int r = threadIdx.x; r < rows; r += 1024) mine += (perm[r] >= 0 && !done[r]) ? 1 : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < 32; ++w) s += warp_sum[w]; n_run_s = s; base = 0; }
    __syncthreads();
    const int n_run = n_run_s;
    // pass 2: holes = rows below n_run that do not run (stopped or dead)
    for (int r0 = 0; r0 < n_run; r0 += 1024) {
        const int r = r0 + threadIdx.x;
        const int f = (r < n_run && !(perm[r] >= 0 && !done[r])) ? 1 : 0;
        int before;
        block_scan_1024(f, warp_sum, &base, before);
        if (f) holes[before] = r;
    }
    __syncthreads();
    const int n_holes = base;
    __syncthreads();
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    // pass 3: movers = running rows at or above n_run
    for (int r0 = n_run; r0 < rows; r0 += 1024) {
        const int r = r0 + threadIdx.x;
        const int f = (r < rows && perm[r] >= 0 && !done[r]) ? 1 : 0;
        int before;
        block_scan_1024(f, warp_sum, &base, before);
        if (f) movers[before] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) { counts[0] = n_run; counts[1] = min(n_holes, base); }
}

struct RowSet {                  // row-major arrays that travel with an instance
    float* ptr[14];
    int len[14];                 // floats per row (the leading dimension)
    int count;
};

struct Bookkeeping { int *done, *need, *iters, *status, *perm; float *max_viol, *gap; };

// stopped (or, at the end of the solve, all live) rows -> archive at their original index; the row dies
__global__ void __launch_bounds__(256)
compact_archive_kernel(int rows, int all, RowSet src, RowSet dst, Bookkeeping bk, Bookkeeping ar) {
    const int r = blockIdx.x;
    if (r >= rows) return;
    const int o = bk.perm[r];
    if (o < 0 || (!all && !bk.done[r])) return;
    for (int a = 0; a < src.count; ++a) {
        const float4* s = reinterpret_cast<const float4*>(src.ptr[a] + (size_t)r * src.len[a]);
        float4* d = reinterpret_cast<float4*>(dst.ptr[a] + (size_t)o * dst.len[a]);
        for (int i = threadIdx.x; i < src.len[a] / 4; i += blockDim.x) d[i] = s[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ar.iters[o] = bk.iters[r]; ar.status[o] = bk.status[r]; ar.max_viol[o] = bk.max_viol[r]; ar.gap[o] = bk.gap[r];
        bk.perm[r] = -1;
    }
}

__global__ void __launch_bounds__(256)
compact_move_kernel(const int* __restrict__ counts, const int* __restrict__ holes, const int* __restrict__ movers, RowSet rows,
                    float* __restrict__ red, Bookkeeping bk) {
    const int i = blockIdx.x;
    if (i >= counts[1]) return;
    const int s = movers[i], d = holes[i];
    for (int a = 0; a < rows.count; ++a) {
        const float4* sp = reinterpret_cast<const float4*>(rows.ptr[a] + (size_t)s * rows.len[a]);
        float4* dp = reinterpret_cast<float4*>(rows.ptr[a] + (size_t)d * rows.len[a]);
        for (int k = threadIdx.x; k < rows.len[a] / 4; k += blockDim.x) dp[k] = sp[k];
    }
    if (threadIdx.x < kRedStride) red[(size_t)d * kRedStride + threadIdx.x] = red[(size_t)s * kRedStride + threadIdx.x];
    if (threadIdx.x == 0) {
        bk.iters[d] = bk.iters[s]; bk.status[d] = bk.status[s]; bk.max_viol[d] = bk.max_viol[s]; bk.gap[d] = bk.gap[s];
        bk.need[d] = bk.need[s]; bk.need[s] = 0;
        bk.perm[d] = bk.perm[s]; bk.perm[s] = -1;
        bk.done[d] = 0; bk.done[s] = 1;
    }
}

__global__ void perm_identity_kernel(int* __restrict__ perm, int Bp, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < Bp) perm[b] = b < B ? b : -1;
}

RowSet work_rows(const BatchState& st, bool have_f) {
    RowSet r{};
    auto add = [&](float* p, int len) { if (p) { r.ptr[r.count] = p; r.len[r.count] = len; ++r.count; } };
    add(st.p_D, st.mp); add(st.g_P, st.np); if (have_f) add(st.f, st.np);
    add(st.yb[0], st.mp); add(st.yb[1], st.mp); add(st.yb[2], st.mp);
    add(st.z, st.np); add(st.zhat, st.np); add(st.sbar, st.mp);
    add(st.Pb[0], st.np); add(st.Pb[1], st.np);
    return r;
}

RowSet output_rows(const BatchState& st) {
    RowSet r{};
    float* p[5] = {st.yb[0], st.yb[1], st.yb[2], st.z, st.zhat};
    const int len[5] = {st.mp, st.mp, st.mp, st.np, st.np};
    for (int a = 0; a < 5; ++a) { r.ptr[a] = p[a]; r.len[a] = len[a]; }
    r.count = 5;
    return r;
}

Bookkeeping books(const BatchState& st) { return {st.done, st.need, st.iters, st.status, st.perm, st.max_viol, st.gap}; }

}  // namespace

int launch_perm_identity(const BatchState& st, cudaStream_t s) {
    perm_identity_kernel<<<(st.Bp + 255) / 256, 256, 0, s>>>(st.perm, st.Bp, st.B);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

// rows_bound: an upper bound of the working rows (rows at or above it are dead already)
int launch_compact(const BatchState& st, const BatchState& archive, int rows_bound, bool have_f, cudaStream_t s) {
    compact_plan_kernel<<<1, 1024, 0, s>>>(rows_bound, st.done, st.perm, st.holes, st.movers, st.compact_counts);
    compact_archive_kernel<<<rows_bound, 256, 0, s>>>(rows_bound, 0, output_rows(st), output_rows(archive), books(st), books(archive));
    compact_move_kernel<<<(rows_bound + 1) / 2, 256, 0, s>>>(st.compact_counts, st.holes, st.movers, work_rows(st, have_f), st.red, books(st));
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

// end of a compacted solve: every row that is still live goes to the archive
int launch_archive_all(const BatchState& st, const BatchState& archive, int rows_bound, cudaStream_t s) {
    compact_archive_kernel<<<rows_bound, 256, 0, s>>>(rows_bound, 1, output_rows(st), output_rows(archive), books(st), books(archive));
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

}  // namespace gpad
