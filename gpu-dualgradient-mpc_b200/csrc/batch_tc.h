// batch_tc.h -- host interface of the tcgen05 3xTF32 GEMM path (batch_tc.cu)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "batch_common.cuh"

namespace gpad {
namespace tc {

struct GemmDesc {
    CUtensorMap tmA_hi, tmA_lo;   // A [rows = batch][K] tiles of 128 x bk (product 1: y_v, split in-kernel; product 2: zhat hi / lo)
    CUtensorMap tmY[3];           // product 1: the three rotating y buffers
    CUtensorMap tmB_hi, tmB_lo;   // B [rows = outputs][K] tiles of bn x bk
    int bk = 16;                  // K block in floats (SWIZZLE_64B operand tiles)
    int k_pad = 0;                // K rounded up to bk
    int m_tiles = 0, n_tiles = 0, bn = 0, stages = 0;
    int pdl = 0;                  // launch with programmatic stream serialization
    int cluster_attr = 0;         // launch as clusters of one CTA (changes the CTA -> SM assignment order)
    int p1 = 0;                   // product 1 runs the second-generation kernel (batch_tc_p1.cu): stages = operator ring,
    int a_stages = 0;             // a_stages = state ring
    int step = 0;                 // p1: column distance between tile starts (<= bn, see plan_tiles_p1); 0 = bn
    int acc_stages = 2;           // p1: TMEM accumulator stages (1: single-wave plans with tiles of up to 256 columns)
    int ncols_valid = 0;          // output columns that exist (n or m)
    int f16 = 0;                  // GPAD_PREC_FP16X3: operand maps over fp16 hi / lo arrays (k_pad counts K elements)
    // second-generation product 2 (batch_tc_p2.cu): epilogue operands by TMA -- the rotating y buffers and p_D as
    // [rows = batch][m] fp32 maps with boxes of 128 rows x 32 columns (loads and the store of y_{v+1})
    int p2 = 0;
    CUtensorMap tmEy[3], tmEpd;
    int e_stages = 0;
};

int make_tmap(CUtensorMap* map, const float* ptr, int k_elems, int rows, int ld, int box_k, int box_rows);
int make_tmap_bytes(CUtensorMap* map, const void* ptr, int elem_bytes, int k_elems, int rows, int ld, int box_k, int box_rows);
void plan_tiles(int ncols, int* bn, int* n_tiles);
size_t smem_bytes(int bk, int bn, int stages);
int pick_stages(int bk, int bn, size_t smem_limit);
int launch_gemm(int phase, const GemmDesc& g, const BatchKernelArgs& args, float* C, int ldc, int num_sms, cudaStream_t s);
void plan_tiles_p1(int ncols, int* bn, int* n_tiles, int* step = nullptr, int max_bn = 208);
int plan_rings_p1(int bn, size_t smem_limit, int* a_stages, int* b_stages, bool f16 = false);
// GPAD_PREC_FP16X3 helpers (batch_f16.cu)
// rows of fp32 -> per-row power-of-two scale, fp16 hi / lo of the scaled rows, inv[r] = 2^-e
int launch_quantize_rows(const float* src, int ld, int rows, uint16_t* hi, uint16_t* lo, float* inv, unsigned* zero_rows,
                         cudaStream_t s, bool pdl = false);
// rowmax[r] = max_k |x[r][k]|
int launch_rowmax(const float* src, int ld, int rows, float* rowmax, cudaStream_t s);
int launch_p1(const GemmDesc& g, const BatchKernelArgs& args, int num_sms, cudaStream_t s);
void plan_tiles_p2(int ncols, int* bn, int* n_tiles);
int plan_rings_p2(int bn, size_t smem_limit, int* stages, int* e_stages, int max_stages = 0);
int launch_p2(const GemmDesc& g, const BatchKernelArgs& args, int cur, int prev, int next, int num_sms, cudaStream_t s);
int launch_split(const float* src, float* hi, float* lo, size_t count, cudaStream_t s);

}  // namespace tc
}  // namespace gpad
