// Version, error string and device check of libtruncgptq.
#include <stdarg.h>
#include <stdlib.h>
#include <time.h>

#include <vector>

#include "common.cuh"

namespace tq {

static thread_local char g_err[512] = "";
thread_local int64_t g_launch_count = 0;

struct ProfSlot {
  cudaEvent_t e0, e1;
  double bytes;
  int kind;
  int sms;       // SM budget of the launching thread (tq_set_sm_budget) when the launch was made
};
static thread_local bool g_prof_on = false;
static thread_local int g_prof_every = 1;
static thread_local int64_t g_prof_seen = 0;                       // kinds 0..2 (tq_profile_end's total)
static thread_local int64_t g_prof_seen_kind[TQ_PROF_KINDS] = {};
static thread_local double g_prof_work_all[TQ_PROF_KINDS] = {};    // work of EVERY launch of the kind, sampled or not
static thread_local std::vector<ProfSlot> g_prof_slots;

int prof_begin_launch(cudaStream_t st, double work, int kind) {
  if (!g_prof_on || kind < 0 || kind >= TQ_PROF_KINDS) return -1;
  const int64_t idx = g_prof_seen_kind[kind]++;
  g_prof_work_all[kind] += work;
  if (kind <= TQ_PROF_QRCP_PANEL) ++g_prof_seen;
  if (idx % g_prof_every != 0 || g_prof_slots.size() >= 16384) return -1;
  ProfSlot s;
  if (cudaEventCreate(&s.e0) != cudaSuccess || cudaEventCreate(&s.e1) != cudaSuccess) return -1;
  s.bytes = work;
  s.kind = kind;
  s.sms = num_sms();
  cudaEventRecord(s.e0, st);
  g_prof_slots.push_back(s);
  return int(g_prof_slots.size()) - 1;
}

void prof_end_launch(cudaStream_t st, int slot) {
  if (slot >= 0) cudaEventRecord(g_prof_slots[slot].e1, st);
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (libtruncgptq has no CPU fallback)", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return TQ_ERR_CUDA;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    return TQ_ERR_CUDA;
  }
  if (major != 10) {
    set_error("device compute capability %d.x is not sm_100 (B200): libtruncgptq is sm_100a only", major);
    return TQ_ERR_UNSUPPORTED;
  }
  return TQ_OK;
}

// TQ_TRACE=1: synchronous stage timers (the stream is synchronised at every stage boundary);
// TQ_TRACE=2: event-based stage timers - nothing is synchronised, the stages are printed by trace_flush() when
//             the solve returns (for timing a solve that shares the GPU with others)
static int trace_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TQ_TRACE");
    v = (e && e[0] && e[0] != '0') ? (e[0] == '2' ? 2 : 1) : 0;
  }
  return v;
}
bool trace_enabled() { return trace_mode() == 1; }

struct TraceEv {
  const char* name;
  cudaEvent_t e0, e1;
};
static thread_local std::vector<TraceEv> g_trace_events;

void trace_flush(cudaStream_t st) {
  if (trace_mode() != 2 || g_trace_events.empty()) return;
  cudaStreamSynchronize(st);
  for (auto& t : g_trace_events) {
    float ms = 0;
    cudaEventElapsedTime(&ms, t.e0, t.e1);
    fprintf(stderr, "[tq-trace2] %-18s %10.3f ms\n", t.name, ms);
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  g_trace_events.clear();
}

static double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

StageTimer::StageTimer(cudaStream_t s, const char* n) : st(s), name(n), t0(0) {
  if (trace_mode() == 1) {
    cudaStreamSynchronize(st);
    t0 = now_ms();
  } else if (trace_mode() == 2) {
    TraceEv t{name, nullptr, nullptr};
    cudaEventCreate(&t.e0);
    cudaEventCreate(&t.e1);
    cudaEventRecord(t.e0, st);
    g_trace_events.push_back(t);
    t0 = double(g_trace_events.size());       // 1-based index of this timer's entry
  }
}
StageTimer::~StageTimer() {
  if (trace_mode() == 1) {
    cudaStreamSynchronize(st);
    fprintf(stderr, "[tq-trace] %-18s %10.3f ms\n", name, now_ms() - t0);
  } else if (trace_mode() == 2 && t0 >= 1.0 && size_t(t0) <= g_trace_events.size()) {
    cudaEventRecord(g_trace_events[size_t(t0) - 1].e1, st);
  }
}

// -1: automatic (two-stage from kTwoStageMinN on), 0 / 1: tq_set_eigh_two_stage
static thread_local int g_two_stage = -1;
int two_stage_setting() { return g_two_stage; }

static thread_local int g_sm_budget = 0;
static thread_local void (*g_stage_cb)(int, void*) = nullptr;
static thread_local void* g_stage_user = nullptr;

bool stage_callback_set() { return g_stage_cb != nullptr; }
void notify_stage(int stage) {
  if (g_stage_cb) g_stage_cb(stage, g_stage_user);
}

// SMs the calling thread's persistent (co-resident) kernels may occupy: the device's count, or the
// budget set with tq_set_sm_budget so that several solves can be in flight on one GPU.
int num_sms() {
  static thread_local int cached_sms[kMaxDevices] = {};
  int& cached = cached_sms[device_slot()];
  if (!cached) {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached = n;
  }
  return (g_sm_budget > 0 && g_sm_budget < cached) ? g_sm_budget : cached;
}

}  // namespace tq

extern "C" int64_t tq_launch_count(void) { return tq::g_launch_count; }

extern "C" int tq_profile_begin(int sample_every) {
  using namespace tq;
  for (auto& s : g_prof_slots) {
    cudaEventDestroy(s.e0);
    cudaEventDestroy(s.e1);
  }
  g_prof_slots.clear();
  g_prof_on = true;
  g_prof_every = sample_every > 0 ? sample_every : 1;
  g_prof_seen = 0;
  for (int k = 0; k < TQ_PROF_KINDS; ++k) {
    g_prof_seen_kind[k] = 0;
    g_prof_work_all[k] = 0.0;
  }
  return TQ_OK;
}

extern "C" int tq_profile_kernel(int kind, double* work, double* ms, int64_t* sampled, int64_t* total,
                                 double* work_all, double* sm_ms) {
  using namespace tq;
  TQ_REQUIRE(kind >= 0 && kind < TQ_PROF_KINDS, "tq_profile_kernel: unknown kind %d", kind);
  double b = 0, t = 0, smt = 0;
  int64_t cnt = 0;
  for (auto& s : g_prof_slots) {
    if (s.kind != kind) continue;
    float f = 0;
    if (cudaEventSynchronize(s.e1) == cudaSuccess && cudaEventElapsedTime(&f, s.e0, s.e1) == cudaSuccess) {
      b += s.bytes;
      t += f;
      smt += double(f) * s.sms;
      ++cnt;
    }
  }
  if (sm_ms) *sm_ms = smt;
  if (work) *work = b;
  if (ms) *ms = t;
  if (sampled) *sampled = cnt;
  if (total) *total = g_prof_seen_kind[kind];
  if (work_all) *work_all = g_prof_work_all[kind];
  return TQ_OK;
}

extern "C" int tq_profile_end(double* alg_bytes, double* ms, int64_t* sampled, int64_t* total) {
  using namespace tq;
  double b = 0, t = 0;
  int64_t cnt = 0;
  for (auto& s : g_prof_slots) {
    float f = 0;
    if (s.kind <= TQ_PROF_QRCP_PANEL && cudaEventSynchronize(s.e1) == cudaSuccess &&
        cudaEventElapsedTime(&f, s.e0, s.e1) == cudaSuccess) {
      b += s.bytes;
      t += f;
      ++cnt;
    }
    cudaEventDestroy(s.e0);
    cudaEventDestroy(s.e1);
  }
  if (alg_bytes) *alg_bytes = b;
  if (ms) *ms = t;
  if (sampled) *sampled = cnt;
  if (total) *total = g_prof_seen;
  g_prof_slots.clear();
  g_prof_on = false;
  return TQ_OK;
}

extern "C" int tq_set_sm_budget(int sms) {
  tq::g_sm_budget = sms > 0 ? sms : 0;
  return TQ_OK;
}

extern "C" int tq_set_eigh_two_stage(int on) {
  tq::g_two_stage = on < 0 ? -1 : (on ? 1 : 0);
  return TQ_OK;
}

extern "C" int tq_set_stage_callback(void (*cb)(int, void*), void* user) {
  tq::g_stage_cb = cb;
  tq::g_stage_user = user;
  return TQ_OK;
}

extern "C" int tq_version(void) { return TQ_VERSION; }
extern "C" const char* tq_last_error(void) { return tq::g_err; }
