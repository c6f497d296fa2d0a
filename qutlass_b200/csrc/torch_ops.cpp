// Compiled torch op layer over the C-ABI of libb200q.so.
//
// Registers the reference's op names and schemas under torch.ops._qutlass_C with the torch STABLE ABI, the way the
// reference's own binding does (qutlass/csrc/bindings.cpp:498-540: STABLE_TORCH_LIBRARY_FRAGMENT + TORCH_BOX), so a call
// goes Python -> torch dispatcher -> this file -> b200q_* with no interpreter work and no ctypes in between.  This file
// contains NO kernels and no CUDA: argument checks (messages follow bindings.cpp:32-102,140-216,218-480), output
// allocation, device guard, current stream, then one call into include/b200q.h.  Built by qutlass_b200/build.py with g++
// into lib/b200q_torch_ops.so (needs only torch's header-only / C-shim headers); loaded with torch.ops.load_library.
//
// Besides the reference's 14 ops there is a private namespace, torch.ops._b200q_C, with the extended entry points the
// Python surface uses (blocked scale output + rotation hint in the quantisers, kernel configuration / static-weights flag in
// the GEMM, the one-call linear).
#include <torch/csrc/stable/accelerator.h>
#include <torch/csrc/stable/library.h>
#include <torch/csrc/stable/ops.h>
#include <torch/csrc/stable/tensor.h>
#include <torch/headeronly/core/ScalarType.h>
#include <torch/csrc/inductor/aoti_torch/c/shim.h>

#include <optional>
#include <string>
#include <tuple>

#include "b200q.h"

namespace b200q_ops {

using torch::headeronly::ScalarType;
using torch::stable::Tensor;

#define B200Q_CHECK(cond, ...) STD_TORCH_CHECK(cond, __VA_ARGS__)

static b200q_stream_t current_stream(const Tensor& t) {
  void* s = nullptr;
  TORCH_ERROR_CODE_CHECK(aoti_torch_get_current_cuda_stream(t.get_device_index(), &s));
  return reinterpret_cast<b200q_stream_t>(s);
}

static void rc_check(int rc) {
  if (rc != 0) {
    const char* msg = b200q_last_error();
    STD_TORCH_CHECK(false, (msg && msg[0]) ? msg : "libb200q call failed");
  }
}

struct Named {
  const Tensor& t;
  const char* name;
};

static void check_cuda_same(const char* op, std::initializer_list<Named> ts) {
  int dev = -1;
  for (const Named& n : ts) {
    B200Q_CHECK(n.t.is_cuda(), op, ": expected all tensors to be on CUDA, but ", n.name, " is not");
    if (dev < 0) dev = n.t.get_device_index();
    B200Q_CHECK(n.t.get_device_index() == dev, op, ": expected all tensors on the same GPU, but ", n.name, " is on cuda:",
                (int)n.t.get_device_index(), " (vs cuda:", dev, ")");
  }
}
static void check_contig(const char* op, std::initializer_list<Named> ts) {
  for (const Named& n : ts) B200Q_CHECK(n.t.is_contiguous(), op, ": expected ", n.name, " to be contiguous");
}

static int64_t blocked_elems(int64_t rows, int64_t k, int64_t group) {
  return ((rows + 127) / 128) * 128 * ((((k / group) + 3) / 4) * 4);
}

// ------------------------------------------------------------------------------------------ GEMM
// kind_flags = B200Q_KIND_* | optional B200Q_GEMM_STATIC_WEIGHTS (reference checks: bindings.cpp:32-102,140-216)
static Tensor gemm_impl(const char* op, const Tensor& A, const Tensor& B, const Tensor& A_sf, const Tensor& B_sf, const Tensor& alpha,
                        int64_t kind_flags, int64_t cta_group, int64_t block_n) {
  const int kind = (int)(kind_flags & 0xff);
  check_contig(op, {{A, "A"}, {B, "B"}, {A_sf, "A_sf"}, {B_sf, "B_sf"}});
  check_cuda_same(op, {{A, "A"}, {B, "B"}, {A_sf, "A_sf"}, {B_sf, "B_sf"}, {alpha, "alpha"}});
  const bool f8 = kind == B200Q_KIND_MXF8 || kind == B200Q_KIND_MXF8_NN;
  const bool nn = kind == B200Q_KIND_MXF8_NN;
  const bool nv = kind == B200Q_KIND_NVF4;
  const ScalarType op_dt = f8 ? ScalarType::Float8_e4m3fn : ScalarType::Byte;
  B200Q_CHECK(A.scalar_type() == op_dt, f8 ? "A must be float8_e4m3fn" : "A must be uint8");
  B200Q_CHECK(B.scalar_type() == op_dt, f8 ? "B must be float8_e4m3fn" : "B must be uint8");
  const ScalarType sf_dt = nv ? ScalarType::Float8_e4m3fn : ScalarType::Float8_e8m0fnu;
  B200Q_CHECK(A_sf.scalar_type() == sf_dt, nv ? "A_sf must be float8_e4m3fn" : "A_sf must be float8_e8m0fnu");
  B200Q_CHECK(B_sf.scalar_type() == sf_dt, nv ? "B_sf must be float8_e4m3fn" : "B_sf must be float8_e8m0fnu");
  B200Q_CHECK(A.dim() == 2 && B.dim() == 2, "A and B must be 2D");
  const int64_t min_k = nv ? 16 : 32;
  if (nn) {
    B200Q_CHECK(A.size(0) == B.size(1), "Inner dimensions must match for A.T @ B.T");
    B200Q_CHECK(A.size(0) >= min_k, "A K-dim must be >= ", min_k);
    B200Q_CHECK(A.size(1) % 16 == 0, "M (", A.size(1), ") must be a multiple of 16");
  } else {
    B200Q_CHECK(A.size(1) == B.size(1), "Inner dimensions must match for A @ B.T");
    B200Q_CHECK(A.size(1) >= min_k, "A K-dim must be >= ", min_k);
  }
  B200Q_CHECK(B.size(1) >= min_k, "B K-dim must be >= ", min_k);
  B200Q_CHECK(alpha.scalar_type() == ScalarType::Float && alpha.numel() >= 1, "alpha must be a float32 tensor with one element");
  const int64_t m = nn ? A.size(1) : A.size(0), n = B.size(0), k = B.size(1) * (f8 ? 1 : 2);
  const int64_t group = nv ? 16 : 32;
  B200Q_CHECK(k % 32 == 0, "K (", k, ") must be a multiple of 32");
  B200Q_CHECK(A_sf.numel() >= blocked_elems(m, k, group), "A_sf has ", A_sf.numel(), " scales, the blocked layout needs ",
              blocked_elems(m, k, group));
  B200Q_CHECK(B_sf.numel() >= blocked_elems(n, k, group), "B_sf has ", B_sf.numel(), " scales, the blocked layout needs ",
              blocked_elems(n, k, group));
  B200Q_CHECK(m < (1ll << 31) && n < (1ll << 31) && k < (1ll << 31), "M, N, K must fit 32 bits");
  Tensor out = torch::stable::new_empty(A, {m, n}, ScalarType::BFloat16);
  torch::stable::accelerator::DeviceGuard guard(A.get_device_index());
  rc_check(b200q_gemm_fp4_cfg(A.const_data_ptr(), B.const_data_ptr(), A_sf.const_data_ptr(), B_sf.const_data_ptr(),
                              (const float*)alpha.const_data_ptr(), out.mutable_data_ptr(), (int)m, (int)n, (int)k, (int)kind_flags,
                              (int)cta_group, (int)block_n, current_stream(A)));
  return out;
}

Tensor matmul_mxf4_bf16_tn(const Tensor& A, const Tensor& B, const Tensor& A_sf, const Tensor& B_sf, const Tensor& alpha) {
  return gemm_impl("matmul_mxf4_bf16_tn", A, B, A_sf, B_sf, alpha, B200Q_KIND_MXF4, 0, 0);
}
Tensor matmul_nvf4_bf16_tn(const Tensor& A, const Tensor& B, const Tensor& A_sf, const Tensor& B_sf, const Tensor& alpha) {
  return gemm_impl("matmul_nvf4_bf16_tn", A, B, A_sf, B_sf, alpha, B200Q_KIND_NVF4, 0, 0);
}
Tensor matmul_mxf8_bf16_tn(const Tensor& A, const Tensor& B, const Tensor& A_sf, const Tensor& B_sf, const Tensor& alpha) {
  return gemm_impl("matmul_mxf8_bf16_tn", A, B, A_sf, B_sf, alpha, B200Q_KIND_MXF8, 0, 0);
}
Tensor matmul_mxf8_bf16_nn(const Tensor& A, const Tensor& B, const Tensor& A_sf, const Tensor& B_sf, const Tensor& alpha) {
  return gemm_impl("matmul_mxf8_bf16_nn", A, B, A_sf, B_sf, alpha, B200Q_KIND_MXF8_NN, 0, 0);
}
// sm_120-only prototype in the reference (gemm_ada.cu): the schema exists, the call fails on sm_100 there as well
Tensor matmul_ada_mxf4_bf16_tn(const Tensor&, const Tensor&, const Tensor&, const Tensor&, const Tensor&) {
  STD_TORCH_CHECK(false, "matmul_ada_mxf4_bf16_tn: sm_120-only prototype in the reference; outside the sm_100a hot path of qutlass_b200");
  return Tensor();
}
// private: explicit kind flags / kernel configuration
Tensor gemm_fp4(const Tensor& A, const Tensor& B, const Tensor& A_sf, const Tensor& B_sf, const Tensor& alpha, int64_t kind_flags,
                int64_t cta_group, int64_t block_n) {
  return gemm_impl("gemm_fp4", A, B, A_sf, B_sf, alpha, kind_flags, cta_group, block_n);
}

// ------------------------------------------------------------------------------------------ quantisers
// reference checks: bindings.cpp:218-252,292-333,335-426
static int64_t quant_checks(const char* op, const Tensor& A, const Tensor& R, const Tensor& OUT, const Tensor& OUT_sf) {
  check_contig(op, {{A, "A"}, {R, "B"}, {OUT, "OUT0"}, {OUT_sf, "OUT1"}});
  check_cuda_same(op, {{A, "A"}, {R, "B"}, {OUT, "OUT0"}, {OUT_sf, "OUT1"}});
  B200Q_CHECK(A.scalar_type() == ScalarType::BFloat16, "A must be bf16");
  B200Q_CHECK(R.scalar_type() == ScalarType::BFloat16, "B must be bf16");
  B200Q_CHECK(R.dim() == 2 && R.size(0) == R.size(1), "Rotation matrix must be square");
  const int64_t had = R.size(0);
  B200Q_CHECK(A.numel() % had == 0, "A must be divisible by", had);
  B200Q_CHECK(A.dim() >= 1 && A.size(A.dim() - 1) % 32 == 0, "last dimension of A must be a multiple of 32");
  return had;
}

static void quantize_mx_impl(const Tensor& A, const Tensor& R, Tensor& OUT, Tensor& OUT_sf, const std::optional<Tensor>& OUT_blocked,
                             const std::optional<Tensor>& OUT_mask, int64_t method_flags) {
  const int64_t had = quant_checks("fusedQuantizeMx", A, R, OUT, OUT_sf);
  B200Q_CHECK(had == 32 || had == 64 || had == 128, "Unsupported rotation size ", had, "; expected 32, 64, or 128.");
  torch::stable::accelerator::DeviceGuard guard(A.get_device_index());
  rc_check(b200q_quantize_mx(A.const_data_ptr(), R.const_data_ptr(), OUT.mutable_data_ptr(), OUT_sf.mutable_data_ptr(),
                             OUT_blocked ? OUT_blocked->mutable_data_ptr() : nullptr, OUT_mask ? OUT_mask->mutable_data_ptr() : nullptr,
                             A.numel(), A.size(A.dim() - 1), (int)had, (int)method_flags, current_stream(A)));
}

static void quantize_nv_impl(const Tensor& A, const Tensor& R, Tensor& OUT, Tensor& OUT_sf, const std::optional<Tensor>& OUT_blocked,
                             const Tensor& global_scale, int64_t method_flags) {
  const int64_t had = quant_checks("fusedQuantizeNv", A, R, OUT, OUT_sf);
  check_cuda_same("fusedQuantizeNv", {{A, "A"}, {global_scale, "global_scale"}});
  B200Q_CHECK(global_scale.scalar_type() == ScalarType::Float, "global_scale must be float");
  B200Q_CHECK(global_scale.dim() == 1 && global_scale.size(0) == 1, "global_scale must be a scalar");
  B200Q_CHECK(had == 16 || had == 32 || had == 64 || had == 128, "Unsupported rotation size ", had, "; expected 16, 32, 64, or 128.");
  torch::stable::accelerator::DeviceGuard guard(A.get_device_index());
  rc_check(b200q_quantize_nv(A.const_data_ptr(), R.const_data_ptr(), OUT.mutable_data_ptr(), OUT_sf.mutable_data_ptr(),
                             OUT_blocked ? OUT_blocked->mutable_data_ptr() : nullptr, (const float*)global_scale.const_data_ptr(),
                             A.numel(), A.size(A.dim() - 1), (int)had, (int)method_flags, current_stream(A)));
}

std::tuple<Tensor, Tensor> fusedQuantizeMxQuest(const Tensor& A, const Tensor& R, Tensor OUT, Tensor OUT_sf) {
  quantize_mx_impl(A, R, OUT, OUT_sf, std::nullopt, std::nullopt, B200Q_METHOD_QUEST);
  return {OUT, OUT_sf};
}
std::tuple<Tensor, Tensor> fusedQuantizeMxAbsMax(const Tensor& A, const Tensor& R, Tensor OUT, Tensor OUT_sf) {
  quantize_mx_impl(A, R, OUT, OUT_sf, std::nullopt, std::nullopt, B200Q_METHOD_ABSMAX);
  return {OUT, OUT_sf};
}
std::tuple<Tensor, Tensor, Tensor> fusedQuantizeMxQuestWithMask(const Tensor& A, const Tensor& R, Tensor OUT, Tensor OUT_sf, Tensor OUT_mask) {
  check_contig("fusedQuantizeMxQuestWithMask", {{OUT_mask, "OUT2"}});
  check_cuda_same("fusedQuantizeMxQuestWithMask", {{A, "A"}, {OUT_mask, "OUT2"}});
  quantize_mx_impl(A, R, OUT, OUT_sf, std::nullopt, OUT_mask, B200Q_METHOD_QUEST);
  return {OUT, OUT_sf, OUT_mask};
}
std::tuple<Tensor, Tensor> fusedQuantizeNvQuest(const Tensor& A, const Tensor& R, Tensor OUT, Tensor OUT_sf, const Tensor& global_scale) {
  quantize_nv_impl(A, R, OUT, OUT_sf, std::nullopt, global_scale, B200Q_METHOD_QUEST);
  return {OUT, OUT_sf};
}
std::tuple<Tensor, Tensor> fusedQuantizeNvAbsMax(const Tensor& A, const Tensor& R, Tensor OUT, Tensor OUT_sf, const Tensor& global_scale) {
  quantize_nv_impl(A, R, OUT, OUT_sf, std::nullopt, global_scale, B200Q_METHOD_ABSMAX);
  return {OUT, OUT_sf};
}
// private: + blocked scale output, clip mask, method | rotation-hint flags (include/b200q.h)
void quantize_mx(const Tensor& A, const Tensor& R, Tensor OUT, Tensor OUT_sf, std::optional<Tensor> OUT_blocked, std::optional<Tensor> OUT_mask,
                 int64_t method_flags) {
  quantize_mx_impl(A, R, OUT, OUT_sf, OUT_blocked, OUT_mask, method_flags);
}
void quantize_nv(const Tensor& A, const Tensor& R, Tensor OUT, Tensor OUT_sf, std::optional<Tensor> OUT_blocked, const Tensor& global_scale,
                 int64_t method_flags) {
  quantize_nv_impl(A, R, OUT, OUT_sf, OUT_blocked, global_scale, method_flags);
}
// private: row-major scales -> blocked layout (qutlass/utils.py:160-193)
Tensor swizzle_sf(const Tensor& sf) {
  B200Q_CHECK(sf.dim() == 2, "to_blocked expects a 2-D scale matrix");
  B200Q_CHECK(sf.element_size() == 1, "Expected element size to be 1 byte (8 bits)");
  B200Q_CHECK(sf.is_cuda(), "to_blocked: input must be a CUDA tensor (no CPU path in qutlass_b200)");
  B200Q_CHECK(sf.is_contiguous(), "to_blocked: input must be contiguous");
  const int64_t rows = sf.size(0), cols = sf.size(1);
  Tensor out = torch::stable::new_empty(sf, {((rows + 127) / 128) * 128 * (((cols + 3) / 4) * 4)});
  torch::stable::accelerator::DeviceGuard guard(sf.get_device_index());
  rc_check(b200q_swizzle_sf(sf.const_data_ptr(), out.mutable_data_ptr(), rows, cols, current_stream(sf)));
  return out;
}

// ------------------------------------------------------------------------------------------ backward re-quantisers
// (bindings.cpp:428-480 forwards raw pointers without checks; the shape checks live in the Python wrappers there and here)
void backward_t_bf16(const Tensor& x, const Tensor& h, Tensor xh_e2m1, Tensor xh_e8m0) {
  check_cuda_same("backward_t_bf16", {{x, "x"}, {h, "h"}, {xh_e2m1, "xh_e2m1"}, {xh_e8m0, "xh_e8m0"}});
  check_contig("backward_t_bf16", {{x, "x"}, {h, "h"}, {xh_e2m1, "xh_e2m1"}, {xh_e8m0, "xh_e8m0"}});
  B200Q_CHECK(x.scalar_type() == ScalarType::BFloat16 && h.scalar_type() == ScalarType::BFloat16, "backward_t_bf16: x and h must be bf16");
  B200Q_CHECK(x.dim() >= 2, "backward_t_bf16: x must have at least 2 dimensions");
  B200Q_CHECK(h.dim() == 2 && h.size(0) == 32 && h.size(1) == 32, "backward_t_bf16: h must be a 32 x 32 rotation matrix");
  const int64_t size_m = x.size(x.dim() - 1), size_n = x.size(x.dim() - 2);
  const int64_t size_b = x.numel() / std::max<int64_t>(size_m * size_n, 1);
  B200Q_CHECK((int64_t)(xh_e2m1.numel() * xh_e2m1.element_size()) == size_b * size_m * size_n / 2, "backward_t_bf16: xh_e2m1 has the wrong size");
  B200Q_CHECK(xh_e8m0.numel() == size_b * size_m * size_n / 32, "backward_t_bf16: xh_e8m0 has the wrong size");
  torch::stable::accelerator::DeviceGuard guard(x.get_device_index());
  rc_check(b200q_backward_t_bf16(x.const_data_ptr(), h.const_data_ptr(), xh_e2m1.mutable_data_ptr(), xh_e8m0.mutable_data_ptr(), (int)size_m,
                                 (int)size_n, (int)size_b, 0, current_stream(x)));
}
void backward_qt_bf16(const Tensor& x_e2m1, const Tensor& x_e8m0, const Tensor& h, const Tensor& alpha, Tensor xh_e2m1, Tensor xh_e8m0) {
  check_cuda_same("backward_qt_bf16", {{x_e2m1, "x_e2m1"}, {x_e8m0, "x_e8m0"}, {h, "h"}, {alpha, "alpha"}, {xh_e2m1, "xh_e2m1"}, {xh_e8m0, "xh_e8m0"}});
  check_contig("backward_qt_bf16", {{x_e2m1, "x_e2m1"}, {x_e8m0, "x_e8m0"}, {h, "h"}, {xh_e2m1, "xh_e2m1"}, {xh_e8m0, "xh_e8m0"}});
  B200Q_CHECK(x_e2m1.element_size() == 1 && x_e8m0.element_size() == 1, "backward_qt_bf16: x_e2m1 / x_e8m0 must be 1-byte dtypes");
  B200Q_CHECK(alpha.scalar_type() == ScalarType::Float && alpha.numel() >= 1, "backward_qt_bf16: alpha must be a float32 tensor with one element");
  B200Q_CHECK(h.scalar_type() == ScalarType::BFloat16 && h.dim() == 2 && h.size(0) == 32 && h.size(1) == 32,
              "backward_qt_bf16: h must be a 32 x 32 bf16 rotation matrix");
  B200Q_CHECK(x_e2m1.dim() >= 2, "backward_qt_bf16: x_e2m1 must have at least 2 dimensions");
  const int64_t size_m = x_e2m1.size(x_e2m1.dim() - 1) * 2, size_n = x_e2m1.size(x_e2m1.dim() - 2);
  const int64_t size_b = x_e2m1.numel() / std::max<int64_t>(x_e2m1.size(x_e2m1.dim() - 1) * size_n, 1);
  B200Q_CHECK(x_e8m0.numel() == size_b * size_n * size_m / 32, "backward_qt_bf16: x_e8m0 has the wrong size");
  B200Q_CHECK((int64_t)(xh_e2m1.numel() * xh_e2m1.element_size()) == size_b * size_m * size_n / 2, "backward_qt_bf16: xh_e2m1 has the wrong size");
  B200Q_CHECK(xh_e8m0.numel() == size_b * size_m * size_n / 32, "backward_qt_bf16: xh_e8m0 has the wrong size");
  torch::stable::accelerator::DeviceGuard guard(h.get_device_index());
  rc_check(b200q_backward_qt_bf16(x_e2m1.const_data_ptr(), x_e8m0.const_data_ptr(), h.const_data_ptr(), (const float*)alpha.const_data_ptr(),
                                  xh_e2m1.mutable_data_ptr(), xh_e8m0.mutable_data_ptr(), (int)size_m, (int)size_n, (int)size_b, 0,
                                  current_stream(h)));
}
void backward_bf16_square_double_mxfp8(const Tensor& x_bf16, Tensor x_fp8, Tensor row_scales, Tensor column_scales) {
  check_cuda_same("backward_bf16_square_double_mxfp8", {{x_bf16, "x_bf16"}, {x_fp8, "x_fp8"}, {row_scales, "row_scales"}, {column_scales, "column_scales"}});
  check_contig("backward_bf16_square_double_mxfp8", {{x_bf16, "x_bf16"}, {x_fp8, "x_fp8"}, {row_scales, "row_scales"}, {column_scales, "column_scales"}});
  B200Q_CHECK(x_bf16.scalar_type() == ScalarType::BFloat16 && x_bf16.dim() == 2, "backward_bf16_square_double_mxfp8: x_bf16 must be a 2-D bf16 tensor");
  const int64_t m = x_bf16.size(0), n = x_bf16.size(1), m_pad = (m + 127) / 128 * 128;
  B200Q_CHECK(n % 32 == 0, "backward_bf16_square_double_mxfp8: x_bf16.size(1) (", n, ") must be a multiple of 32");
  B200Q_CHECK(x_fp8.numel() == m_pad * n && x_fp8.element_size() == 1, "backward_bf16_square_double_mxfp8: x_fp8 must hold ", m_pad, " x ", n, " bytes");
  B200Q_CHECK(row_scales.numel() == m_pad * n / 32, "backward_bf16_square_double_mxfp8: row_scales has the wrong size");
  B200Q_CHECK(column_scales.numel() == n * m_pad / 32, "backward_bf16_square_double_mxfp8: column_scales has the wrong size");
  torch::stable::accelerator::DeviceGuard guard(x_bf16.get_device_index());
  rc_check(b200q_backward_bf16_square_double_mxfp8(x_bf16.const_data_ptr(), (int)m, (int)n, x_fp8.mutable_data_ptr(), row_scales.mutable_data_ptr(),
                                                   column_scales.mutable_data_ptr(), current_stream(x_bf16)));
}
// like the reference binding (bindings.cpp:466-479) the raw op takes x_fp4 already padded to 256 rows: m = x_fp4.size(0)
void mxfp4_transpose_mxfp8(const Tensor& x_fp4, const Tensor& scales, Tensor x_fp8, Tensor shared_exps) {
  check_cuda_same("mxfp4_transpose_mxfp8", {{x_fp4, "x_fp4"}, {scales, "scales"}, {x_fp8, "x_fp8"}, {shared_exps, "shared_exps"}});
  check_contig("mxfp4_transpose_mxfp8", {{x_fp4, "x_fp4"}, {scales, "scales"}, {x_fp8, "x_fp8"}, {shared_exps, "shared_exps"}});
  B200Q_CHECK(x_fp4.dim() == 2 && x_fp4.element_size() == 1, "mxfp4_transpose_mxfp8: x_fp4 must be a 2-D 1-byte tensor");
  const int64_t m = x_fp4.size(0), n = x_fp4.size(1) * 2, m_pad = (m + 255) / 256 * 256;
  B200Q_CHECK(n % 32 == 0, "mxfp4_transpose_mxfp8: 2 * x_fp4.size(1) (", n, ") must be a multiple of 32");
  B200Q_CHECK(scales.numel() >= m * (n / 32), "mxfp4_transpose_mxfp8: scales must hold at least ", m, " x ", n / 32, " bytes");
  B200Q_CHECK(x_fp8.numel() == n * m_pad, "mxfp4_transpose_mxfp8: x_fp8 must hold ", n, " x ", m_pad, " bytes");
  B200Q_CHECK(shared_exps.numel() == n * m_pad / 32, "mxfp4_transpose_mxfp8: shared_exps has the wrong size");
  torch::stable::accelerator::DeviceGuard guard(x_fp4.get_device_index());
  rc_check(b200q_mxfp4_transpose_mxfp8(x_fp4.const_data_ptr(), scales.const_data_ptr(), (int)m, (int)n, x_fp8.mutable_data_ptr(),
                                       shared_exps.mutable_data_ptr(), current_stream(x_fp4)));
}

}  // namespace b200q_ops

STABLE_TORCH_LIBRARY_FRAGMENT(_qutlass_C, ops) {
  ops.def("matmul_mxf4_bf16_tn(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha) -> Tensor");
  ops.def("matmul_nvf4_bf16_tn(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha) -> Tensor");
  ops.def("matmul_ada_mxf4_bf16_tn(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha) -> Tensor");
  ops.def("matmul_mxf8_bf16_tn(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha) -> Tensor");
  ops.def("matmul_mxf8_bf16_nn(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha) -> Tensor");
  ops.def("fusedQuantizeMxQuest(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf) -> (Tensor, Tensor)");
  ops.def("fusedQuantizeMxAbsMax(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf) -> (Tensor, Tensor)");
  ops.def("fusedQuantizeNvQuest(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf, Tensor global_scale) -> (Tensor, Tensor)");
  ops.def("fusedQuantizeNvAbsMax(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf, Tensor global_scale) -> (Tensor, Tensor)");
  ops.def("fusedQuantizeMxQuestWithMask(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf, Tensor OUT_mask) -> (Tensor, Tensor, Tensor)");
  ops.def("backward_t_bf16(Tensor x, Tensor h, Tensor xh_e2m1, Tensor xh_e8m0) -> ()");
  ops.def("backward_qt_bf16(Tensor x_e2m1, Tensor x_e8m0, Tensor h, Tensor alpha, Tensor xh_e2m1, Tensor xh_e8m0) -> ()");
  ops.def("backward_bf16_square_double_mxfp8(Tensor x_bf16, Tensor x_fp8, Tensor row_scales, Tensor column_scales) -> ()");
  ops.def("mxfp4_transpose_mxfp8(Tensor x_fp4, Tensor scales, Tensor x_fp8, Tensor shared_exps) -> ()");
}

STABLE_TORCH_LIBRARY_IMPL(_qutlass_C, CUDA, ops) {
  ops.impl("matmul_mxf4_bf16_tn", TORCH_BOX(&b200q_ops::matmul_mxf4_bf16_tn));
  ops.impl("matmul_nvf4_bf16_tn", TORCH_BOX(&b200q_ops::matmul_nvf4_bf16_tn));
  ops.impl("matmul_ada_mxf4_bf16_tn", TORCH_BOX(&b200q_ops::matmul_ada_mxf4_bf16_tn));
  ops.impl("matmul_mxf8_bf16_tn", TORCH_BOX(&b200q_ops::matmul_mxf8_bf16_tn));
  ops.impl("matmul_mxf8_bf16_nn", TORCH_BOX(&b200q_ops::matmul_mxf8_bf16_nn));
  ops.impl("fusedQuantizeMxQuest", TORCH_BOX(&b200q_ops::fusedQuantizeMxQuest));
  ops.impl("fusedQuantizeMxAbsMax", TORCH_BOX(&b200q_ops::fusedQuantizeMxAbsMax));
  ops.impl("fusedQuantizeNvQuest", TORCH_BOX(&b200q_ops::fusedQuantizeNvQuest));
  ops.impl("fusedQuantizeNvAbsMax", TORCH_BOX(&b200q_ops::fusedQuantizeNvAbsMax));
  ops.impl("fusedQuantizeMxQuestWithMask", TORCH_BOX(&b200q_ops::fusedQuantizeMxQuestWithMask));
  ops.impl("backward_t_bf16", TORCH_BOX(&b200q_ops::backward_t_bf16));
  ops.impl("backward_qt_bf16", TORCH_BOX(&b200q_ops::backward_qt_bf16));
  ops.impl("backward_bf16_square_double_mxfp8", TORCH_BOX(&b200q_ops::backward_bf16_square_double_mxfp8));
  ops.impl("mxfp4_transpose_mxfp8", TORCH_BOX(&b200q_ops::mxfp4_transpose_mxfp8));
}

STABLE_TORCH_LIBRARY_FRAGMENT(_b200q_C, ops) {
  ops.def("gemm_fp4(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha, int kind_flags, int cta_group, int block_n) -> Tensor");
  ops.def("quantize_mx(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf, Tensor? OUT_blocked, Tensor? OUT_mask, int method_flags) -> ()");
  ops.def("quantize_nv(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf, Tensor? OUT_blocked, Tensor global_scale, int method_flags) -> ()");
  ops.def("swizzle_sf(Tensor sf) -> Tensor");
}

STABLE_TORCH_LIBRARY_IMPL(_b200q_C, CUDA, ops) {
  ops.impl("gemm_fp4", TORCH_BOX(&b200q_ops::gemm_fp4));
  ops.impl("quantize_mx", TORCH_BOX(&b200q_ops::quantize_mx));
  ops.impl("quantize_nv", TORCH_BOX(&b200q_ops::quantize_nv));
  ops.impl("swizzle_sf", TORCH_BOX(&b200q_ops::swizzle_sf));
}
