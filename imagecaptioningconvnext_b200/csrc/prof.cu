// prof.cu — event pool behind ccx_prof_begin / ccx_prof_end (see ccx_prof.h).  Host code only.
#include <vector>

#include "../../include/ccx.h"
#include "ccx_prof.h"

namespace ccx {

bool g_prof_on = false;

namespace {
struct Span {
  cudaEvent_t a, b;
  int kind;
  double work;
};
std::vector<Span> g_spans;
std::vector<cudaEvent_t> g_pool;
size_t g_pool_next = 0;

cudaEvent_t take_event() {
  if (g_pool_next == g_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_pool.push_back(e);
  }
  return g_pool[g_pool_next++];
}
}  // namespace

void prof_record(int kind, cudaStream_t stream, bool begin, double work) {
  if (begin) {
    Span s;
    s.a = take_event();
    s.b = nullptr;
    s.kind = kind;
    s.work = work;
    cudaEventRecord(s.a, stream);
    g_spans.push_back(s);
  } else {
    // scopes nest strictly (a launcher never calls another profiled launcher), so the open span is the last one
    Span& s = g_spans.back();
    s.b = take_event();
    cudaEventRecord(s.b, stream);
  }
}

}  // namespace ccx

extern "C" {

int ccx_prof_begin(void) {
  ccx::g_spans.clear();
  ccx::g_pool_next = 0;
  ccx::g_prof_on = true;
  return CCX_OK;
}

// Per-launch detail of the spans recorded since ccx_prof_begin (call BEFORE ccx_prof_end, after a device sync):
// fills up to `max` entries, returns the number of spans recorded (or a negative status).
int ccx_prof_spans(int32_t* kind, double* ms, double* work, int32_t max) {
  if (cudaDeviceSynchronize() != cudaSuccess) return CCX_ERR_CUDA;
  int n = 0;
  for (const auto& s : ccx::g_spans) {
    if (s.b == nullptr) continue;
    if (n < max) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, s.a, s.b) != cudaSuccess) return CCX_ERR_CUDA;
      kind[n] = s.kind;
      ms[n] = t;
      work[n] = s.work;
    }
    ++n;
  }
  return n;
}

int ccx_prof_end(double* ms_per_kind, double* work_per_kind, int64_t* launches_per_kind, int32_t n_kinds) {
  ccx::g_prof_on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) return CCX_ERR_CUDA;
  for (int k = 0; k < n_kinds; ++k) {
    if (ms_per_kind) ms_per_kind[k] = 0.0;
    if (work_per_kind) work_per_kind[k] = 0.0;
    if (launches_per_kind) launches_per_kind[k] = 0;
  }
  for (const auto& s : ccx::g_spans) {
    if (s.kind < 0 || s.kind >= n_kinds || s.b == nullptr) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) != cudaSuccess) return CCX_ERR_CUDA;
    if (ms_per_kind) ms_per_kind[s.kind] += ms;
    if (work_per_kind) work_per_kind[s.kind] += s.work;
    if (launches_per_kind) launches_per_kind[s.kind] += 1;
  }
  ccx::g_spans.clear();
  return CCX_OK;
}

}  // extern "C"
