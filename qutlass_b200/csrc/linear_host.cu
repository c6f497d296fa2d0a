// Host-buffer entry point for the whole hot path (what bench.py's e2e leg times):
//   pinned host bf16 activations -> H2D -> fused rotate+quantise (abs_max) -> block-scaled FP4 GEMM
//   against pre-quantised device weights -> D2H of the bf16 result.
// Mirrors the reference's "actual" benchmark iteration (benchmarks/bench_mxfp4_sm100.py:93-106:
// fusedQuantizeMx -> to_blocked -> matmul_mxf4_bf16_tn) with the copies a host caller pays.
//
// The rows are processed in slabs on three streams (H2D / compute on the caller's stream / D2H) joined with events,
// so the 117 MB result copy overlaps the input copy and the kernels: the call is bound by the D2H direction of PCIe.
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

constexpr int kMaxSlabs = 16;
struct SideStreams {
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t fork = nullptr, in_ready[kMaxSlabs] = {}, out_ready[kMaxSlabs] = {}, done = nullptr;
  bool ok = false;
};
static SideStreams& side_streams() {
  static thread_local SideStreams s[16];
  int dev = 0;
  cudaGetDevice(&dev);
  SideStreams& r = s[dev & 15];
  if (!r.ok) {
    bool good = cudaStreamCreateWithFlags(&r.h2d, cudaStreamNonBlocking) == cudaSuccess &&
                cudaStreamCreateWithFlags(&r.d2h, cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&r.fork, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&r.done, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < kMaxSlabs && good; ++i)
      good = cudaEventCreateWithFlags(&r.in_ready[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&r.out_ready[i], cudaEventDisableTiming) == cudaSuccess;
    r.ok = good;
  }
  return r;
}

// Slabs of whole 128-row blocks (each slab's blocked scales are self-contained): slab i = rows [bounds[i], bounds[i+1]).
// The call is bound by the result copy (D2H), which can only start once the first slab has been uploaded and computed:
// so the first slab is a single 128-row block, the second one fills up to the regular slab size, the rest are regular.
static int slab_bounds(int M, int* bounds) {
  int slab_rows = 512;
  if (ceil_div(M, slab_rows) + 1 > kMaxSlabs) slab_rows = (int)round_up(ceil_div(M, kMaxSlabs - 1), 128);
  int n = 0;
  bounds[0] = 0;
  if (M > slab_rows) {
    bounds[++n] = 128;
    bounds[++n] = slab_rows;
  }
  while (bounds[n] < M) {
    const int64_t next = (int64_t)bounds[n] + slab_rows;
    bounds[n + 1] = next < M ? (int)next : M;
    ++n;
  }
  return n;
}
}  // namespace b200q

using namespace b200q;

extern "C" int b200q_linear_host_slabs(int M, int* bounds, int capacity) {
  B200Q_REQUIRE(M > 0 && bounds && capacity >= kMaxSlabs + 1, "need M > 0 and room for %d row bounds", kMaxSlabs + 1);
  return slab_bounds(M, bounds);
}

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
  if (kind == B200Q_KIND_NVF4) B200Q_REQUIRE(global_scale_dev, "global_scale_dev is required for NVFP4");
  const LinearWs w = layout(M, N, K, kind);
  const int group = kind == B200Q_KIND_NVF4 ? 16 : 32;
  uint8_t* base = (uint8_t*)ws;
  cudaStream_t s = (cudaStream_t)stream;
  SideStreams& ss = side_streams();
  B200Q_REQUIRE(ss.ok, "could not create helper streams/events");

  int bounds[kMaxSlabs + 1];
  const int n_slabs = slab_bounds(M, bounds);
  const int64_t sf_cols = round_up(ceil_div(K, group), 4);

  B200Q_CUDA(cudaEventRecord(ss.fork, s));
  B200Q_CUDA(cudaStreamWaitEvent(ss.h2d, ss.fork, 0));
  B200Q_CUDA(cudaStreamWaitEvent(ss.d2h, ss.fork, 0));
  for (int i = 0; i < n_slabs; ++i) {
    const int r0 = bounds[i];
    const int rows = bounds[i + 1] - r0;
    uint8_t* xs = base + w.off_x + (int64_t)r0 * K * 2;
    uint8_t* qs = base + w.off_q + (int64_t)r0 * K / 2;
    uint8_t* sfs = base + w.off_sf + (int64_t)r0 * sf_cols;       // r0 is a multiple of 128: block-aligned
    uint8_t* ds = base + w.off_d + (int64_t)r0 * N * 2;
    B200Q_CUDA(cudaMemcpyAsync(xs, (const uint8_t*)x_host + (int64_t)r0 * K * 2, (size_t)rows * K * 2,
                               cudaMemcpyHostToDevice, ss.h2d));
    B200Q_CUDA(cudaEventRecord(ss.in_ready[i], ss.h2d));
    B200Q_CUDA(cudaStreamWaitEvent(s, ss.in_ready[i], 0));
    int rc;
    if (kind == B200Q_KIND_NVF4)
      rc = b200q_quantize_nv(xs, rot_bf16, qs, nullptr, sfs, global_scale_dev, (int64_t)rows * K, K, had,
                             B200Q_METHOD_ABSMAX, stream);
    else
      rc = b200q_quantize_mx(xs, rot_bf16, qs, nullptr, sfs, nullptr, (int64_t)rows * K, K, had, B200Q_METHOD_ABSMAX,
                             stream);
    if (rc) return rc;
    rc = b200q_gemm_fp4(qs, Wq, sfs, Wsf_blocked, alpha_dev, ds, rows, N, K, kind | B200Q_GEMM_STATIC_WEIGHTS, stream);   // pre-quantised weights by contract
    if (rc) return rc;
    B200Q_CUDA(cudaEventRecord(ss.out_ready[i], s));
    B200Q_CUDA(cudaStreamWaitEvent(ss.d2h, ss.out_ready[i], 0));
    B200Q_CUDA(cudaMemcpyAsync((uint8_t*)d_host + (int64_t)r0 * N * 2, ds, (size_t)rows * N * 2, cudaMemcpyDeviceToHost,
                               ss.d2h));
  }
  // join: the caller's stream completes only after the last result copy
  B200Q_CUDA(cudaEventRecord(ss.done, ss.d2h));
  B200Q_CUDA(cudaStreamWaitEvent(s, ss.done, 0));
  return 0;
}
