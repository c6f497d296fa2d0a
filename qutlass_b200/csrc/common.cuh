// Shared host/device helpers for libb200q (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>

#include "b200q.h"

namespace b200q {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int check_device_sm100();          // 0 or B200Q_EUNSUPPORTED (cached per device)
int num_sms();                     // SM count of the current device (cached)
int current_device();              // ordinal of the current device, 0 on error
void note_sf_write(const void* sf_rowmajor);   // every quantise call that writes a row-major scale buffer (b200q_sf_write_generation)

// Environment switches, read ONCE (first use) instead of getenv() on every launch; b200q_reload_env() re-reads them
// (tests and probe tools that flip a switch inside one process call it).  Switches that change RESULTS (the GEMM's
// profiling flags: skip loads / copies / stores) exist only in a library built with -DB200Q_PROFILING; the product
// build ignores B200Q_GEMM_DEBUG_FLAGS entirely.
struct Env {
  int no_pdl;        // B200Q_NO_PDL: 0 = PDL everywhere, 1 = off, 2 = GEMM only
  int tail_split;    // B200Q_TAIL_SPLIT=1
  int gemm_hybrid;   // B200Q_GEMM_HYBRID=1
  int fuse;          // B200Q_FUSE=1
  int fuse_warps;    // B200Q_FUSE_WARPS: 4 (default) or 2
  int quant_mma;     // B200Q_QUANT_MMA=1
  int quant_tc;      // B200Q_QUANT_TC: -1 unset, 0, 1
  int gemm_flags;    // B200Q_GEMM_DEBUG_FLAGS (profiling builds only, else 0)
  int verbose;       // B200Q_GEMM_VERBOSE
  int gemm_skew;     // B200Q_GEMM_SKEW: -1 unset (planner decides), else forced k-tile skew of the split accumulator
  int no_tmap_cache; // B200Q_NO_TMAP_CACHE=1
  int decode_pace;   // B200Q_DECODE_PACE: SM cycles between the weight-stage requests of a decode-kernel CTA (-1 unset: library rule, 0: all at once)
  int bwd_pipe;      // B200Q_BWD_PIPE: -1 unset (library default), 0 = one-shot CTAs, 1 = persistent double-buffered transposing kernels
  int bwd_t_tc;      // B200Q_BWD_T_TC: -1 unset (library rule), 0 = CUDA-core backward_t_bf16, 1 = tcgen05 kernel (backward_tc.cu)
  int bwd_qt_tc;     // B200Q_BWD_QT_TC: -1 unset (library rule), 0 = CUDA-core backward_qt_bf16, 1 = tcgen05 kernel
  int fuse_decode;   // B200Q_FUSE_DECODE=1: b200q_linear_fp4 runs the decode step (M <= 32) as ONE launch (measured slower: opt-in)
};
const Env& env();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: set it once per (kernel, device).
// `done` is a per-instantiation bitmask of device ordinals (static in the caller).
template <typename Kern>
inline int ensure_dynamic_smem(Kern kern, int bytes, std::atomic<unsigned long long>& done) {
  const int dev = current_device() & 63;
  if (!((done.load(std::memory_order_acquire) >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %d) failed: %s", bytes, cudaGetErrorString(e));
      return B200Q_ECUDA;
    }
    done.fetch_or(1ull << dev, std::memory_order_release);   // setting the attribute twice from two threads is harmless
  }
  return 0;
}

// Programmatic dependent launch of the quantise kernels: fills attr[0] and returns the attribute count (0 with
// B200Q_NO_PDL=1, which switches PDL off everywhere, or =2, which keeps it for the GEMM only).  A kernel launched with it
// may be scheduled while the previous kernel in the stream drains; inside, everything that touches global memory sits
// behind griddepcontrol.wait (ptx::pdl_wait), which returns once that kernel has completed and its writes are visible.
inline unsigned pdl_attribute(cudaLaunchAttribute* attr) {
  if (env().no_pdl) return 0;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  return 1;
}

#define B200Q_REQUIRE(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::b200q::set_error(__VA_ARGS__);           \
      return B200Q_EINVAL;                       \
    }                                            \
  } while (0)

#define B200Q_CUDA(expr)                                                          \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::b200q::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                         __FILE__, __LINE__);                                     \
      return B200Q_ECUDA;                                                         \
    }                                                                             \
  } while (0)

// ---------------------------------------------------------------- scale layout
// Byte offset of scale (r, c) in the block-scaled layout: 128x4 tiles of 512 B,
// K-blocks fastest (reference: qutlass/utils.py:178-193).
__host__ __device__ __forceinline__ int64_t sf_blocked_offset(int64_t r, int64_t c, int64_t padded_cols) {
  return ((r >> 7) * (padded_cols >> 2) + (c >> 2)) * 512 + (r & 31) * 16 + ((r & 127) >> 5) * 4 + (c & 3);
}

__host__ __device__ __forceinline__ int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
__host__ __device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace b200q
