// Error plumbing + device queries for libb200q.
#include "common.cuh"

#include <string.h>

namespace b200q {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static Env g_env;
static std::atomic<int> g_env_ready{0};

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

static void load_env() {
  Env e;
  e.no_pdl = env_int("B200Q_NO_PDL", 0);
  if (e.no_pdl < 0 || e.no_pdl > 2) e.no_pdl = 0;
  e.tail_split = env_int("B200Q_TAIL_SPLIT", 0) == 1;
  e.gemm_hybrid = env_int("B200Q_GEMM_HYBRID", 0) == 1;
  e.fuse = env_int("B200Q_FUSE", 0) == 1;
  e.fuse_warps = env_int("B200Q_FUSE_WARPS", 4) == 2 ? 2 : 4;
  e.quant_mma = env_int("B200Q_QUANT_MMA", 0) == 1;
  e.quant_tc = env_int("B200Q_QUANT_TC", -1);
  if (e.quant_tc != 0 && e.quant_tc != 1) e.quant_tc = -1;
#ifdef B200Q_PROFILING
  e.gemm_flags = env_int("B200Q_GEMM_DEBUG_FLAGS", 0);
#else
  e.gemm_flags = 0;
#endif
  e.verbose = getenv("B200Q_GEMM_VERBOSE") != nullptr;
  e.gemm_skew = env_int("B200Q_GEMM_SKEW", -1);
  e.no_tmap_cache = env_int("B200Q_NO_TMAP_CACHE", 0) == 1;
  e.fuse_decode = env_int("B200Q_FUSE_DECODE", 0) == 1;
  e.bwd_pipe = env_int("B200Q_BWD_PIPE", -1);
  if (e.bwd_pipe != 0 && e.bwd_pipe != 1) e.bwd_pipe = -1;
  e.bwd_t_tc = env_int("B200Q_BWD_T_TC", -1);
  if (e.bwd_t_tc != 0 && e.bwd_t_tc != 1) e.bwd_t_tc = -1;
  e.bwd_qt_tc = env_int("B200Q_BWD_QT_TC", -1);
  if (e.bwd_qt_tc != 0 && e.bwd_qt_tc != 1) e.bwd_qt_tc = -1;
  e.decode_pace = env_int("B200Q_DECODE_PACE", -1);
  if (e.decode_pace < -1 || e.decode_pace > 100000) e.decode_pace = -1;
  g_env = e;
  g_env_ready.store(1, std::memory_order_release);
}

const Env& env() {
  if (!g_env_ready.load(std::memory_order_acquire)) load_env();   // racing first calls store identical values
  return g_env;
}

static int g_sm_major[64];
static int g_sm_count[64];
static bool g_dev_init[64];

static int query_device(int* dev_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return B200Q_ECUDA;
  }
  if (dev < 0 || dev >= 64) {
    set_error("device ordinal %d out of range", dev);
    return B200Q_EINVAL;
  }
  if (!g_dev_init[dev]) {
    int major = 0, minor = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_sm_major[dev] = major * 10 + minor;
    g_sm_count[dev] = sms;
    g_dev_init[dev] = true;
  }
  *dev_out = dev;
  return 0;
}

int check_device_sm100() {
  int dev;
  int rc = query_device(&dev);
  if (rc) return rc;
  if (g_sm_major[dev] != 100) {
    set_error("libb200q is built for sm_100a only; current device is sm_%d", g_sm_major[dev]);
    return B200Q_EUNSUPPORTED;
  }
  return 0;
}

int current_device() {
  int dev;
  return query_device(&dev) ? 0 : dev;
}

int num_sms() {
  int dev;
  if (query_device(&dev)) return 148;
  return g_sm_count[dev] > 0 ? g_sm_count[dev] : 148;
}

// Write generations of row-major scale buffers (b200q.h: b200q_sf_write_generation): a small table of counters indexed by a
// hash of the pointer.  Two buffers sharing a slot only cause a spurious "changed" answer, never a missed one.
static std::atomic<unsigned> g_sf_gen[4096];
static inline unsigned sf_slot(const void* p) {
  return (unsigned)((((uintptr_t)p >> 4) * 0x9E3779B97F4A7C15ull) >> 52);
}
void note_sf_write(const void* sf_rowmajor) {
  if (sf_rowmajor) g_sf_gen[sf_slot(sf_rowmajor)].fetch_add(1u, std::memory_order_relaxed);
}
unsigned sf_write_generation(const void* sf_rowmajor) { return g_sf_gen[sf_slot(sf_rowmajor)].load(std::memory_order_relaxed); }

}  // namespace b200q

extern "C" unsigned b200q_sf_write_generation(const void* sf_rowmajor) { return b200q::sf_write_generation(sf_rowmajor); }
extern "C" int b200q_abi_version(void) { return 1; }
extern "C" void b200q_reload_env(void) { b200q::load_env(); }
extern "C" int b200q_profiling_build(void) {
#ifdef B200Q_PROFILING
  return 1;
#else
  return 0;
#endif
}
extern "C" const char* b200q_last_error(void) { return b200q::g_err; }
