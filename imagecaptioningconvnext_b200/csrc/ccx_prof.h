// ccx_prof.h — optional per-launch CUDA-event timing inside the library (used by bench.py for the roofline).
// Off by default: a launcher pays one predictable branch.  When on, every kernel launch is bracketed by two
// events on the launch stream and attributed to a kernel kind; ccx_prof_end() sums them.
#pragma once
#include <cuda_runtime.h>

namespace ccx {

enum ProfKind : int {
  PROF_GEMM = 0,
  PROF_DWCONV_LN = 1,
  PROF_STEM = 2,
  PROF_LN_ROWS = 3,
  PROF_POOL = 4,
  PROF_ELEMENTWISE = 5,
  PROF_ATTENTION = 6,
  PROF_LSTM = 7,
  PROF_LOSS = 8,
  PROF_OPTIM = 9,
  PROF_GEMM_SKINNY = 10,   // M <= 32 recurrent GEMMs (gemm_skinny.cu): latency-bound, kept apart from the tcgen05 roofline
  PROF_NUM_KINDS = 11
};

extern bool g_prof_on;
void prof_record(int kind, cudaStream_t stream, bool begin, double work);

struct ProfScope {
  int kind;
  cudaStream_t stream;
  // work: algorithmic FLOPs (tensor-bound kinds) or bytes (HBM-bound kinds) of this launch
  ProfScope(int k, cudaStream_t s, double work) : kind(k), stream(s) {
    if (g_prof_on) prof_record(kind, stream, true, work);
  }
  ~ProfScope() {
    if (g_prof_on) prof_record(kind, stream, false, 0.0);
  }
};

}  // namespace ccx
