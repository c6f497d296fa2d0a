// Host-buffer entry point for the whole hot path (what bench.py's e2e leg times):
//   pinned host bf16 activations -> H2D -> fused rotate+quantise (abs_max) -> block-scaled FP4 GEMM
//   against pre-quantised device weights -> D2H of the bf16 result.
// Mirrors the reference's "actual" benchmark iteration (benchmarks/bench_mxfp4_sm100.py:93-106:
// fusedQuantizeMx -> to_blocked -> matmul_mxf4_bf16_tn) with the copies a host caller pays.
#include "common.cuh"

namespace b200q {
struct LinearWs {
  int64_t off_x, off_q, off_sf, off_d, total;
};
static LinearWs layout(int M, int N, int K, int kind) {
  const int group = kind == B200Q_KIND_NVF4 ? 16 : 32;
  auto al = [](int64_t v) { return round_up(v, 256); };
  LinearWs w;
  w.off_x = 0;
  w.off_q = al((int64_t)M * K * 2);
  w.off_sf = w.off_q + al((int64_t)M * K / 2);
  w.off_d = w.off_sf + al(round_up(M, 128) * round_up(ceil_div(K, group), 4));
  w.total = w.off_d + al((int64_t)M * N * 2);
  return w;
}
}  // namespace b200q

using namespace b200q;

extern "C" int64_t b200q_linear_workspace_bytes(int M, int N, int K, int kind) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return layout(M, N, K, kind).total;
}

extern "C" int b200q_linear_fp4_host(const void* x_host, const void* rot_bf16, const void* Wq, const void* Wsf_blocked,
                                     const float* alpha_dev, const float* global_scale_dev, void* d_host, void* ws,
                                     int M, int N, int K, int had, int kind, b200q_stream_t stream) {
  B200Q_REQUIRE(x_host && rot_bf16 && Wq && Wsf_blocked && alpha_dev && d_host && ws, "null pointer argument");
  B200Q_REQUIRE(kind == B200Q_KIND_MXF4 || kind == B200Q_KIND_NVF4, "invalid kind %d", kind);
  B200Q_REQUIRE(M > 0 && N > 0 && K > 0 && K % 32 == 0, "bad shape M=%d N=%d K=%d", M, N, K);
  const LinearWs w = layout(M, N, K, kind);
  uint8_t* base = (uint8_t*)ws;
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_CUDA(cudaMemcpyAsync(base + w.off_x, x_host, (size_t)M * K * 2, cudaMemcpyHostToDevice, s));
  int rc;
  if (kind == B200Q_KIND_NVF4) {
    B200Q_REQUIRE(global_scale_dev, "global_scale_dev is required for NVFP4");
    rc = b200q_quantize_nv(base + w.off_x, rot_bf16, base + w.off_q, nullptr, base + w.off_sf, global_scale_dev,
                           (int64_t)M * K, K, had, B200Q_METHOD_ABSMAX, stream);
  } else {
    rc = b200q_quantize_mx(base + w.off_x, rot_bf16, base + w.off_q, nullptr, base + w.off_sf, nullptr,
                           (int64_t)M * K, K, had, B200Q_METHOD_ABSMAX, stream);
  }
  if (rc) return rc;
  rc = b200q_gemm_fp4(base + w.off_q, Wq, base + w.off_sf, Wsf_blocked, alpha_dev, base + w.off_d, M, N, K, kind, stream);
  if (rc) return rc;
  B200Q_CUDA(cudaMemcpyAsync(d_host, base + w.off_d, (size_t)M * N * 2, cudaMemcpyDeviceToHost, s));
  return 0;
}
