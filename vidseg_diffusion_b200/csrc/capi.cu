// Library-wide state of libvidseg_b200.so: error string, launch counter, version probes.
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace vidseg {
thread_local char g_last_error[512] = "";
std::atomic<long long> g_launch_count{0};
std::atomic<int> g_profile_on{0};
bool pdl_enabled(int link) {
  static const int mask = [] { const char* e = getenv("VIDSEG_KM_PDL"); return e ? atoi(e) : 0xff; }();
  return (mask >> link) & 1;
}
std::atomic<int> g_operand_mode{1};

namespace {
struct ProfRecord { cudaEvent_t start, stop; int family; double work; };
std::mutex g_prof_mutex;
std::vector<ProfRecord> g_prof_pending;
std::vector<cudaEvent_t> g_prof_pool;
double g_prof_ms[kNumFamilies];
double g_prof_work[kNumFamilies];
long long g_prof_launches[kNumFamilies];

cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
// fold every finished record into the accumulators (synchronises on the recorded events)
void prof_drain() {
  for (ProfRecord& r : g_prof_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.stop) == cudaSuccess && cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
      g_prof_ms[r.family] += ms;
      g_prof_work[r.family] += r.work;
      g_prof_launches[r.family] += 1;
    }
    g_prof_pool.push_back(r.start);
    g_prof_pool.push_back(r.stop);
  }
  g_prof_pending.clear();
}
}  // namespace

void profile_before(int family, double work, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  ProfRecord r{prof_event(), prof_event(), family, work};
  cudaEventRecord(r.start, stream);
  g_prof_pending.push_back(r);
}
void profile_after(cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (!g_prof_pending.empty()) cudaEventRecord(g_prof_pending.back().stop, stream);
}
}  // namespace vidseg

VS_API int vidseg_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(vidseg::g_prof_mutex);
  if (on) {
    vidseg::prof_drain();
    for (int f = 0; f < vidseg::kNumFamilies; ++f) {
      vidseg::g_prof_ms[f] = 0.0; vidseg::g_prof_work[f] = 0.0; vidseg::g_prof_launches[f] = 0;
    }
  }
  vidseg::g_profile_on.store(on ? 1 : 0);
  return 0;
}
VS_API int vidseg_profile_read(int family, double* ms_total, long long* launches, double* work_total) {
  VS_REQUIRE(family >= 0 && family < vidseg::kNumFamilies, "bad kernel family");
  std::lock_guard<std::mutex> lock(vidseg::g_prof_mutex);
  vidseg::prof_drain();
  if (ms_total) *ms_total = vidseg::g_prof_ms[family];
  if (launches) *launches = vidseg::g_prof_launches[family];
  if (work_total) *work_total = vidseg::g_prof_work[family];
  return 0;
}

VS_API int vidseg_set_operand_mode(int mode) {
  VS_REQUIRE(mode == 0 || mode == 1, "operand mode must be 0 (fp16 pairs) or 1 (fp16 + fp8 corrections)");
  vidseg::g_operand_mode.store(mode);
  return 0;
}
VS_API int vidseg_get_operand_mode(void) { return vidseg::g_operand_mode.load(); }

VS_API const char* vidseg_last_error(void) { return vidseg::g_last_error; }
VS_API int vidseg_abi_version(void) { return 7; }
VS_API long long vidseg_launch_count(void) { return vidseg::g_launch_count.load(); }
VS_API int vidseg_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}
