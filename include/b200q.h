/*
 * b200q.h -- C-ABI of the B200-native microscaled-FP4 hot path (libb200q.so).
 *
 * This is the drop-in boundary: plain pointers, sizes and a CUDA stream; no torch
 * types.  Each entry point names the reference interface it replaces (paths are
 * relative to the IST-DASLab/qutlass checkout).  The reference's own precedent for a
 * raw C boundary is qutlass/csrc/include/backward_host.h:4-46 (device pointers, ints,
 * cudaStream_t, int return).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers unless the name ends in _host.
 *   - Nothing here allocates or frees device memory and nothing synchronises the
 *     host: every call only enqueues work on `stream` (CUDA-graph capturable).
 *   - Return value: 0 on success, negative B200Q_E* on failure; b200q_last_error()
 *     returns a thread-local human-readable message for the last failure.
 *   - e2m1: two codes per byte, element 2i in the low nibble (tests/mxfp4_test.py:80).
 *   - "blocked" scale layout = cuBLAS block-scaled layout, 128x4 tiles of 512 B,
 *     K-blocks fastest (qutlass/utils.py:160-193).
 */
#ifndef B200Q_H_
#define B200Q_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* opaque here so the header does not need cuda_runtime.h; it IS a cudaStream_t */
typedef struct CUstream_st* b200q_stream_t;

#define B200Q_OK 0
#define B200Q_EINVAL (-1)   /* bad argument (shape / alignment / enum)        */
#define B200Q_ECUDA (-2)    /* a CUDA runtime / driver call failed             */
#define B200Q_EUNSUPPORTED (-3) /* device is not sm_100                        */

#define B200Q_METHOD_QUEST 0   /* Quartet / "quest": std-based scale           */
#define B200Q_METHOD_ABSMAX 1  /* abs-max scale                                */
/* OR into `method`: the caller asserts rot == c * Sylvester-Hadamard (c any bf16 scalar, e.g.
 * scipy.linalg.hadamard(H) * H**-0.5 as every reference test/benchmark builds it).  The kernel then skips its
 * device-side structure check.  Without the flag the kernel verifies R itself (exact, on the device, graph-safe)
 * and falls back to a generic x @ R for any other matrix. */
#define B200Q_ROT_TRUSTED_HADAMARD 0x100
/* OR into `method`: the caller knows rot is NOT of that form (e.g. identity, a learned rotation): the rotation then
 * runs on the tensor cores (tcgen05 kernel from 1 M elements, mma.sync kernel below that, had >= 32) instead of the
 * butterfly kernel's scalar fallback.  Large Hadamard inputs take the tcgen05 kernel too (it streams at the HBM rate
 * for any R); the choice never changes the results beyond fp32 summation order. */
#define B200Q_ROT_GENERIC 0x200
/* NVFP4 abs_max with Hadamard-128 -- the ONE case where the reference's sm_100 dispatch (bindings.cpp:413-415 ->
 * fused_quantize_nv_sm100.cu:192-207) deviates from its other kernels (mma.sync, H = 16/32/64, sm_120) and from its own test
 * oracle (tests/nvfp4_test.py:132-170): that kernel stores the e4m3-ROUNDED scale but derives the codes from the UNROUNDED one
 * (cutlass_extensions/epilogue/fusion/sm100_visitor_store_tma_warpspecialized.hpp:141-148,567-591).  This library targets
 * sm_100 only, so b200q_quantize_nv reproduces that kernel by DEFAULT (validated on B200 against the compiled reference:
 * scales identical, dequantised values identical up to <= 2e-4 +0/-0 codes).  OR B200Q_NV_ORACLE_CODES into `method` to get the
 * oracle / mma.sync arithmetic instead (codes from the rounded scale: slightly more accurate dequantisation; differs from
 * the reference's sm_100 output in 4.8 % of the values / 9.2 % of the code bytes, scales identical).  Ignored for every other
 * method / size.  B200Q_NV_SM100_CODES (round-1 opt-in for the now-default behaviour) is still accepted and has no effect. */
#define B200Q_NV_SM100_CODES 0x400
#define B200Q_NV_ORACLE_CODES 0x800

#define B200Q_KIND_MXF4 0      /* e2m1 x e2m1, ue8m0 scales, group 32          */
#define B200Q_KIND_NVF4 1      /* e2m1 x e2m1, ue4m3 scales, group 16          */
#define B200Q_KIND_MXF8 2      /* e4m3 x e4m3 (one byte / element), ue8m0 scales, group 32 -- the "next"
                                  row of the scope table: replaces matmul_host_mxf8_bf16_tn (gemm.cu:328-380);
                                  A [M, K], B [N, K] bytes, same blocked scale layout                         */
#define B200Q_KIND_MXF8_NN 3   /* as MXF8 but A is stored [K, M] (M contiguous; M % 16 == 0): replaces
                                  matmul_host_mxf8_bf16_nn (gemm.cu:388-434).  D[m,n] = sum_k A[k,m] B[n,k]; the scales
                                  of A stay in the blocked layout of the LOGICAL [M, K/32] matrix                */

/* OR into `kind` of b200q_gemm_fp4 / b200q_gemm_fp4_cfg: the caller guarantees that B and SFB are NOT written by the kernels
 * in front of this GEMM on `stream` (inference: weights quantised once).  The GEMM is a programmatic dependent launch; with
 * this flag its first ring of weight loads is issued before the grid dependency resolves (overlaps the tail of the
 * activation quantiser: -1.5 .. -2.3 us per call).  Without it (default) every global read waits for the preceding kernel,
 * which is what the reference's own pattern quantise(a); quantise(b); matmul(...) and QAT (weights re-quantised every step)
 * need.  b200q_linear_fp4 / b200q_linear_fp4_host take pre-quantised weights by contract and always set it. */
#define B200Q_GEMM_STATIC_WEIGHTS 0x100

/* ABI version of this header (bumped on incompatible change). */
int b200q_abi_version(void);

/* Thread-local message for the most recent failure on this thread ("" if none). */
const char* b200q_last_error(void);

/* The library reads its environment switches (B200Q_NO_PDL, B200Q_QUANT_TC, B200Q_FUSE, ... -- README "environment switches")
 * ONCE, at first use; call this after changing one inside a running process (tests, probe tools). */
void b200q_reload_env(void);
/* 1 if this library was built with -DB200Q_PROFILING (B200Q_GEMM_DEBUG_FLAGS honoured: timing-only switches that skip
 * loads / copies / stores and therefore produce WRONG results), 0 for the product build (the variable is ignored). */
int b200q_profiling_build(void);
/* Host-only: a counter that changes every time one of the quantise entry points of this library is asked to WRITE the
 * row-major scale buffer at `sf_rowmajor` (spuriously also when an unrelated buffer hashing to the same slot is written).
 * The Python surface remembers it next to the blocked copy a quantiser wrote and refuses to hand that copy out from
 * to_blocked() after ANY later writer went through this library -- including the raw torch.ops._qutlass_C ops, which write
 * OUT_sf through its data pointer without touching torch's version counter. */
unsigned b200q_sf_write_generation(const void* sf_rowmajor);
/* Host-only: hits / misses of the calling thread's tensor-map cache since it was last read (then reset). */
int b200q_debug_tmap_cache_stats(unsigned long long* hits, unsigned long long* misses);

/*
 * Fused rotate + MXFP4 quantise.
 * Replaces: fusedQuantizeMx{Quest,AbsMax}[Had64|Had128]_host
 *           (qutlass/csrc/include/fused_quantize_host.h:22-70, fused_quantize_mx.cu:107-207),
 *           fusedQuantizeMxAbsMax_host_sm100 (fused_quantize_mx_sm100.cu:192-207),
 *           fusedQuantizeMxQuestWithMask_host (fused_quantize_mx_mask.cu:107-121)
 *           and the separate to_blocked launch (qutlass/utils.py:16-133).
 *
 *   x_bf16      [numel] bf16, contiguous; rotated in consecutive groups of `had`
 *   rot_bf16    [had, had] bf16 row-major [k, n]; xh = x_group @ rot
 *   q_e2m1      [numel/2] bytes out
 *   sf_rowmajor ue8m0 bytes out; scale of 32-group g stored at byte g (flat, as the
 *               reference kernels do -- row-major [rows, row_len/32] when row_len%128==0);
 *               may be NULL if only the blocked layout is wanted
 *   sf_blocked  NULL, or the [pad128(rows) * pad4(row_len/32)] blocked buffer; written
 *               completely (pad rows/cols are zero-filled) so to_blocked is a no-op
 *   clip_mask   NULL, or [numel/32] uint32 (bit i = |xh_i/s| < 6), quest only
 *   numel       total elements; row_len = size of the last dim (numel % row_len == 0)
 *   had         32, 64 or 128 (numel % had == 0);  method = B200Q_METHOD_*
 */
int b200q_quantize_mx(const void* x_bf16, const void* rot_bf16, void* q_e2m1,
                      void* sf_rowmajor, void* sf_blocked, void* clip_mask,
                      int64_t numel, int64_t row_len, int had, int method,
                      b200q_stream_t stream);

/*
 * Fused rotate + NVFP4 quantise (group 16, e4m3 scales, fp32 global scale on device).
 * Replaces: fusedQuantizeNv{Quest,AbsMax}[Had32|Had64|Had128]_host
 *           (fused_quantize_host.h:72-116, fused_quantize_nv.cu:109-252) and
 *           fusedQuantizeNvAbsMax_host_sm100 (fused_quantize_nv_sm100.cu:192-207).
 *   had in {16, 32, 64, 128}; global_scale_dev points at ONE float on the device.
 */
int b200q_quantize_nv(const void* x_bf16, const void* rot_bf16, void* q_e2m1,
                      void* sf_rowmajor, void* sf_blocked, const float* global_scale_dev,
                      int64_t numel, int64_t row_len, int had, int method,
                      b200q_stream_t stream);

/*
 * Row-major scales -> blocked layout (standalone; only needed for scales that did
 * not come from b200q_quantize_*).  Replaces triton_scale_swizzle (qutlass/utils.py:16-133).
 *   sf_rowmajor [rows, cols] bytes (row stride = cols); sf_blocked [pad128(rows)*pad4(cols)].
 */
int b200q_swizzle_sf(const void* sf_rowmajor, void* sf_blocked, int64_t rows, int64_t cols,
                     b200q_stream_t stream);

/*
 * Block-scaled FP4 GEMM:  D[M,N] = bf16( alpha * (A .* SFA) @ (B .* SFB)^T ), fp32 accumulate.
 * Replaces: matmul_host_mxf4_bf16_tn / matmul_host_nvf4_bf16_tn
 *           (qutlass/csrc/include/gemm.h:21-35, gemm.cu:174-326).
 *   A [M, K/2], B [N, K/2] packed e2m1 (K-major, "tn"); SFA/SFB blocked scale buffers
 *   (ue8m0 group 32 for MXF4, ue4m3 group 16 for NVF4) of pad128(M|N) x pad4(K/group);
 *   alpha_dev: one fp32 on the device; D [M, N] bf16 row-major.
 *   K % 32 == 0 (MX) / K % 32 == 0 (NV, 16-byte TMA row pitch), all pointers 16-byte aligned.
 */
int b200q_gemm_fp4(const void* A, const void* B, const void* SFA, const void* SFB,
                   const float* alpha_dev, void* D_bf16, int M, int N, int K, int kind,
                   b200q_stream_t stream);

/*
 * Same, with an explicit kernel configuration (tuning / tests):
 *   cta_group 1|2, block_n in {64,128,192,256} (0 = heuristic for both); cta_group 4 with block_n 192|256 = CTA pairs in
 *   clusters of four whose two pairs multicast their A tiles to each other (experiment: bit-identical, not faster).
 */
int b200q_gemm_fp4_cfg(const void* A, const void* B, const void* SFA, const void* SFB,
                       const float* alpha_dev, void* D_bf16, int M, int N, int K, int kind,
                       int cta_group, int block_n, b200q_stream_t stream);

/* The configuration b200q_gemm_fp4 picks for this problem (host-only: no device work, no stream). */
int b200q_gemm_fp4_plan(int M, int N, int K, int kind, int* cta_group, int* block_n);

/* Number of kernels b200q_gemm_fp4 launches for this problem (1, or 2 when the last 256-column block of N is peeled
 * into a second launch of small tiles to fill the final, mostly empty wave of CTA pairs). */
int b200q_gemm_fp4_launches(int M, int N, int K, int kind);

/*
 * The whole device-side path in ONE call: rotate + quantise x (as b200q_quantize_mx / _nv, clip mask excluded), then
 * D = bf16(alpha * xq @ Wq^T) (as b200q_gemm_fp4).  Same outputs, bit for bit, as the two calls in sequence: xq_e2m1,
 * x_sf_rowmajor (nullable) and x_sf_blocked (required) are written like the quantiser writes them.
 * Replaces the reference's per-layer sequence fusedQuantizeMx -> to_blocked -> matmul_mxf4_bf16_tn
 * (qutlass/__init__.py:149-180, utils.py:160-193, __init__.py:34-43; benchmarks/bench_mxfp4_sm100.py:93-104).
 *
 * Default: the two launches (the GEMM's prologue and weight loads overlap the quantiser's tail through programmatic
 * dependent launch).  With B200Q_FUSE=1 in the environment, `method` carrying B200Q_ROT_TRUSTED_HADAMARD, K % 1024 == 0,
 * N % 8 == 0 and a problem large enough for the CTA-pair GEMM (M > 256) it is ONE persistent kernel: 4 extra warps per
 * CTA (plus the epilogue warps until their first accumulator is ready) quantise the activations and publish 256-row
 * blocks through progress counters in `ws`; the GEMM's TMA producer acquires a block's counter before loading it.
 * Measured on B200 the single kernel is bit-identical but not faster (the part is power-limited under FP4 MMA load:
 * profiles/r01_notes.md), hence opt-in.
 *   ws: NULL (never fuse) or b200q_linear_fp4_workspace_bytes(M) bytes of device memory that the caller zeroes ONCE
 *       (cudaMemset) after allocating; every call leaves it zeroed.  One workspace per stream.
 *   had / method / global_scale_dev as for b200q_quantize_*; kind B200Q_KIND_MXF4 or B200Q_KIND_NVF4.
 */
int64_t b200q_linear_fp4_workspace_bytes(int M);
int b200q_linear_fp4(const void* x_bf16, const void* rot_bf16, void* xq_e2m1, void* x_sf_rowmajor, void* x_sf_blocked,
                     const void* Wq, const void* Wsf_blocked, const float* alpha_dev, const float* global_scale_dev,
                     void* D_bf16, void* ws, int M, int N, int K, int had, int method, int kind, b200q_stream_t stream);
/* Kernels b200q_linear_fp4 launches for this problem (1 when fused, else quantiser + b200q_gemm_fp4_launches). */
int b200q_linear_fp4_launches(int M, int N, int K, int had, int method, int kind);

/*
 * Host-buffer convenience for the whole path (what bench.py's e2e leg times):
 *   x_host [M,K] bf16 (pinned) -> H2D -> rotate+quantise (abs_max, MX or NV) ->
 *   GEMM against pre-quantised weights (device) -> D2H into d_host [M,N] bf16.
 * `ws` is a caller-owned device workspace of b200q_linear_workspace_bytes(M,N,K,kind) bytes.
 * Enqueues on `stream`; the caller synchronises.
 */
int64_t b200q_linear_workspace_bytes(int M, int N, int K, int kind);
int b200q_linear_fp4_host(const void* x_host, const void* rot_bf16, const void* Wq, const void* Wsf_blocked,
                          const float* alpha_dev, const float* global_scale_dev,
                          void* d_host, void* ws, int M, int N, int K, int had, int kind,
                          b200q_stream_t stream);
/* Host-only query: the row slabs b200q_linear_fp4_host pipelines for M rows -- slab i = rows [bounds[i], bounds[i+1]),
 * every bound but the last a multiple of 128 (block-aligned scales); a short first slab starts the result copy early.
 * `bounds` needs room for 17 ints (`capacity`); returns the slab count (<= 16) or a negative error code. */
int b200q_linear_host_slabs(int M, int* bounds, int capacity);

/* ------------------------------------------------------------------------------------------------------------------
 * Transposing re-quantisers of the QAT backward pass.  Same argument order as the
 * reference's own raw-pointer entry points in qutlass/csrc/include/backward_host.h:4-46, plus `flags`
 * (0 or B200Q_ROT_TRUSTED_HADAMARD: rot == c * Sylvester-Hadamard(32), in-register butterflies; without it the
 * rotation is a generic fp32 x @ rot) where a rotation matrix is involved.  All tensors contiguous.
 */

/*
 * MXFP4 abs-max quantisation of rotate(x^T).  Replaces backward_t_bf16_cuda (backward_host.h:17-26,
 * quartet_bwd_sm120.cu:237-318,414-440; Python qutlass/__init__.py:206-244).
 *   x_bf16 [size_b, size_n, size_m] bf16; rot_bf16 [32, 32]
 *   xh_e2m1 [size_b, size_m, size_n/2], xh_e8m0 [size_b, size_m, size_n/32]: for every (b, m) and 32-group of n,
 *   xh = x[b, n-group, m] @ rot; s = 2^floor(log2(amax)); codes e2m1(xh * 3 / s); scale byte of s.
 *   size_n % 32 == 0, size_m % 8 == 0.
 */
int b200q_backward_t_bf16(const void* x_bf16, const void* rot_bf16, void* xh_e2m1, void* xh_e8m0, int size_m,
                          int size_n, int size_b, int flags, b200q_stream_t stream);

/*
 * Same on an MXFP4 input: dequantise (code * 2^(e-127)), transpose, rotate, s = 2^floor(log2(amax / alpha)),
 * codes e2m1(xh * 3 / (s * alpha)).  Replaces backward_qt_bf16_cuda (backward_host.h:4-15,
 * quartet_bwd_sm120.cu:320-412,443-470; Python qutlass/__init__.py:247-283).
 *   x_e2m1 [size_b, size_n, size_m/2], x_e8m0 [size_b, size_n, size_m/32], alpha_dev: one fp32 on the device
 *   outputs as above; size_n % 32 == 0, size_m % 32 == 0.
 */
int b200q_backward_qt_bf16(const void* x_e2m1, const void* x_e8m0, const void* rot_bf16, const float* alpha_dev,
                           void* xh_e2m1, void* xh_e8m0, int size_m, int size_n, int size_b, int flags,
                           b200q_stream_t stream);

/*
 * bf16 [m, n] -> e4m3 with one ue8m0 scale per 32 x 32 tile (biased exponent of the tile's abs-max minus 7; 127 for an
 * all-zero tile), the scale written twice: row_scales [m_pad, n/32] and column_scales [n, m_pad/32], m_pad = m rounded
 * up to 128 (rows >= m are treated as zero -- the reference pads on the host, qutlass/__init__.py:285-288).
 * Replaces backward_bf16_square_double_mxfp8_cuda (backward_host.h:28-36, quartet_bwd_sm120.cu:497-623).
 *   x_fp8 [m_pad, n] bytes; n % 32 == 0.
 */
int b200q_backward_bf16_square_double_mxfp8(const void* x_bf16, int m, int n, void* x_fp8, void* row_scales,
                                            void* column_scales, b200q_stream_t stream);

/*
 * MXFP4 [m, n/2] + ue8m0 scales [>= m, n/32] -> MXFP8 of the TRANSPOSE: x_fp8 [n, m_pad] e4m3 and shared_exps
 * [n, m_pad/32] (32-groups along m, exponent as above), m_pad = m rounded up to 256; rows >= m read as zero (the
 * reference pads on the host, qutlass/__init__.py:296-304).  Replaces mxfp4_transpose_mxfp8_cuda
 * (backward_host.h:38-46, quartet_bwd_sm120.cu:625-734).   n % 32 == 0.
 */
int b200q_mxfp4_transpose_mxfp8(const void* x_fp4, const void* scales_e8m0, int m, int n, void* x_fp8,
                                void* shared_exps, b200q_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200Q_H_ */
