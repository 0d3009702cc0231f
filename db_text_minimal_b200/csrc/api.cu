// api.cu -- error reporting, version and launch accounting of libdbb200.so
#include "common.cuh"
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace dbb {
thread_local char g_last_error[512] = "";
uint64_t g_launch_count = 0;

int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
  return DBB_ECUDA;
}
int set_error(int code, const char* msg) {
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
  return code;
}

// ---- optional per-kernel timing with CUDA events on the launching stream
struct ProfSlot { const char* name; cudaEvent_t e0, e1; };
static std::vector<ProfSlot> g_slots;
static std::vector<std::string> g_labels;     // storage for dynamic labels
static int g_prof_on = 0;
static size_t g_prof_used = 0;
static std::mutex g_prof_mu;

const char* prof_label(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& l : g_labels) if (l == s) return l.c_str();
  g_labels.reserve(4096);
  g_labels.push_back(s);
  return g_labels.back().c_str();
}
bool prof_enabled() { return g_prof_on != 0; }
int prof_begin(const char* name, cudaStream_t s) {
  if (!g_prof_on) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_used == g_slots.size()) {
    ProfSlot p{};
    if (cudaEventCreate(&p.e0) != cudaSuccess || cudaEventCreate(&p.e1) != cudaSuccess) return -1;
    g_slots.push_back(p);
  }
  ProfSlot& p = g_slots[g_prof_used];
  p.name = name;
  cudaEventRecord(p.e0, s);
  return (int)g_prof_used++;
}
void prof_end(int slot, cudaStream_t s) { cudaEventRecord(g_slots[slot].e1, s); }
}  // namespace dbb

extern "C" void dbb_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(dbb::g_prof_mu);
  dbb::g_prof_on = on;
  dbb::g_prof_used = 0;
}
// Synchronises the device, aggregates the recorded launches by kernel label and writes JSON:
// {"kernels": [{"name": ..., "launches": n, "ms": total}, ...]}.  Returns the number of bytes written (0 if buf is too small).
extern "C" size_t dbb_profile_report(char* buf, size_t cap) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(dbb::g_prof_mu);
  std::map<std::string, std::pair<int, double>> agg;
  for (size_t i = 0; i < dbb::g_prof_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, dbb::g_slots[i].e0, dbb::g_slots[i].e1) != cudaSuccess) continue;
    auto& a = agg[dbb::g_slots[i].name];
    a.first += 1; a.second += ms;
  }
  std::string out = "{\"kernels\": [";
  bool first = true;
  for (auto& kv : agg) {
    char tmp[512];
    snprintf(tmp, sizeof(tmp), "%s{\"name\": \"%s\", \"launches\": %d, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
    out += tmp; first = false;
  }
  out += "]}";
  dbb::g_prof_used = 0;
  if (out.size() + 1 > cap) return 0;
  memcpy(buf, out.c_str(), out.size() + 1);
  return out.size();
}

extern "C" int dbb_version(void) { return DBB_VERSION; }

extern "C" const char* dbb_strerror(int code) {
  switch (code) {
    case DBB_OK: return "ok";
    case DBB_EINVAL: return "invalid argument or shape";
    case DBB_EALIGN: return "pointer not 16-byte aligned";
    case DBB_EWORKSPACE: return "workspace too small";
    case DBB_ECUDA: return "CUDA error";
    case DBB_EUNSUPPORTED: return "configuration not supported";
    default: return "unknown error";
  }
}
extern "C" const char* dbb_last_cuda_error(void) { return dbb::g_last_error; }
extern "C" uint64_t dbb_launch_count(void) { return dbb::g_launch_count; }
