// batch_tc.h -- host interface of the tcgen05 3xTF32 GEMM path (batch_tc.cu)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "batch_common.cuh"

namespace gpad {
namespace tc {

struct GemmDesc {
    CUtensorMap tmA_hi, tmA_lo;   // A [rows = batch][K] tiles of 128 x bk (product 1: y_v and y_{v-1}, split in-kernel)
    CUtensorMap tmY[3];           // product 1: the three rotating y buffers
    CUtensorMap tmB_hi, tmB_lo;   // B [rows = outputs][K] tiles of bn x bk
    int cg = 1;                   // 1: one CTA per 128 x bn tile; 2: CTA pair (cta_group::2) per 256 x bn tile
    int mc = 1;                   // cg == 1: CTAs per cluster sharing each operator tile through TMA multicast (1 or 2)
    int bk = 16;                  // K block in floats: 16 (SWIZZLE_64B) or 32 (SWIZZLE_128B)
    int k_pad = 0;                // K rounded up to bk
    int m_tiles = 0, n_tiles = 0, bn = 0, stages = 0;
    int xf2 = 0;                  // product 2 stages zhat as ONE fp32 tile and splits it in shared memory (transform warps)
    int p1 = 0;                   // product 1 runs the second-generation kernel (batch_tc_p1.cu): stages = operator ring,
    int a_stages = 0;             // a_stages = state ring
    int step = 0;                 // p1: column distance between tile starts (<= bn, see plan_tiles_p1); 0 = bn
    int ncols_valid = 0;          // output columns that exist (n or m)
};

int make_tmap(CUtensorMap* map, const float* ptr, int k_elems, int rows, int ld, int box_k, int box_rows);
void plan_tiles(int ncols, int* bn, int* n_tiles);
size_t smem_bytes(int bk, int bn, int stages);
int pick_stages(int bk, int bn, size_t smem_limit);
int launch_gemm(int phase, const GemmDesc& g, const BatchKernelArgs& args, float* C, int ldc, int num_sms, cudaStream_t s);
size_t smem_bytes2(int bk, int bn, int stages);
int pick_stages2(int bk, int bn, size_t smem_limit);
int launch_gemm2(int phase, const GemmDesc& g, const BatchKernelArgs& args, float* C, int ldc, int num_sms, cudaStream_t s);
void plan_tiles_p1(int ncols, int* bn, int* n_tiles, int* step = nullptr);
int plan_rings_p1(int phase, int bn, size_t smem_limit, int* a_stages, int* b_stages);
int launch_p1(int phase, const GemmDesc& g, const BatchKernelArgs& args, int num_sms, cudaStream_t s);
int launch_split(const float* src, float* hi, float* lo, size_t count, cudaStream_t s);

}  // namespace tc
}  // namespace gpad
