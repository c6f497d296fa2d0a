// Tensor-map encoders shared by the GEMM kernels (defined in gemm_fp4.cu; memoised per thread, see encode()).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace b200q {

// [rows, row_bytes] uint8 operand, box = [box_rows, 128 bytes], 128B swizzle, zero fill out of bounds
int make_operand_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t row_bytes, int box_rows, const char* what);
// the same operand as {128 bytes, rows, k-tiles}: one box = all k-tiles of box_rows rows (row_bytes % 128 == 0, <= 256 k-tiles)
int make_operand_ktile_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t row_bytes, int box_rows, const char* what);
// blocked scale buffer as a 3-D tensor {128 x u32 (one 512-B block), col_blocks, row_blocks}; box = {128, box_kb, box_rb};
// out-of-range blocks read as zero (scale 2^-127 / 0.0: never NaN)
int make_sf_tmap(CUtensorMap* tm, const void* ptr, int64_t row_blocks, int64_t col_blocks, int box_kb, int box_rb,
                 const char* what);

// gemm_decode.cu: the weight-streaming kernel for M <= 32 (operands swapped: weights are the 128-row MMA operand)
bool decode_eligible(int M, int N, int K, int ldd, int kind);
int launch_gemm_decode(const void* A, const void* B, const void* SFA, const void* SFB, const float* alpha, void* D, int M, int N,
                       int K, int ldd, int kind, bool static_w, cudaStream_t stream);

// the decode step in one launch (activations rotated + quantised inside every CTA of the decode kernel)
struct QuantParams;
bool decode_fuse_eligible(int M, int N, int K, int had, int method, int kind);
int launch_gemm_decode_fused(const QuantParams& q, int had, int method, const void* B, const void* SFB, const float* alpha, void* D,
                             int M, int N, int K, int kind, cudaStream_t stream);

}  // namespace b200q
