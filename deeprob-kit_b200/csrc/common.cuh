// common.cuh -- shared helpers of libdeeprob_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/deeprob_b200.h"

namespace dpk {

// ---- host side -------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);  // stores a thread-local message, returns `code`
int sm_count();                                 // SMs of the current device (cached per device)
int max_dynamic_smem();                         // opt-in shared memory per block (bytes)

#define DPK_CUDA_TRY(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::dpk::set_error(DPK_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                           \
  } while (0)

#define DPK_LAUNCH_CHECK(name)                                                               \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return ::dpk::set_error(DPK_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

// ---- lightweight launch accounting / per-category CUDA-event timing (bench.py, tests) -----------
enum ProfCat { CAT_PREP = 0, CAT_LEAF = 1, CAT_EINSUM = 2, CAT_ROOT = 3, CAT_BWD_EINSUM = 4, CAT_BWD_LEAF = 5,
               CAT_FINALIZE = 6, CAT_LAYER = 7, CAT_DGC = 8, CAT_DGC_BWD = 9, CAT_FLOW = 10, CAT_FLOW_BWD = 11,
               CAT_GEMM = 12, CAT_LEAF_MMA = 13, CAT_LEAF_MMA_PREP = 14, CAT_COUNT = 16 };
// RAII: counts `launches` kernel launches in category `cat`; when profiling is enabled also brackets
// them with CUDA events on `st` (dpk_profile_read sums the elapsed times).
struct ProfScope {
  ProfScope(int cat, cudaStream_t st, int launches = 1);
  ~ProfScope();
  int cat_; cudaStream_t st_; void* rec_;
};

// ---- programmatic dependent launch ---------------------------------------------------------------
// The kernels of an inference step are short (10-350 us) and strictly dependent.  A kernel launched through launch_pdl
// may be scheduled while its predecessor on the stream is still running (CTAs start as SMs free up) and must call
// pdl_wait() before it touches anything the predecessor -- or, transitively, anything before it -- wrote; the predecessor
// calls pdl_launch_dependents() once all its CTAs are resident.  Launch latency and prologues (barrier init, TMEM
// allocation, weight images into shared memory) then overlap the predecessor's tail.  Stream capture records the edge.
// Both device calls are no-ops in a kernel that was launched the ordinary way.  Measured inside a replayed CUDA graph:
// the narrow-model step (four kernels of 5-50 us) gains 7 % (0.117 -> 0.109 ms), the config-2 step (0.34 + 0.16 ms
// kernels) LOSES 15 us -- so callers ask for it only where kernels are short (`want`); DPK_PDL=1 / 0 forces it on / off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int pdl_mode();   // -1 auto (caller decides), 0 off, 1 on
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(bool want, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const int mode = pdl_mode();
  cfg.numAttrs = (mode == 1 || (mode < 0 && want)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t v, int64_t m) { return (v + m - 1) / m; }

// channel chunking used by every register-tiled kernel: chunk in {2,4,8,10,16}
struct Chunking {
  int chunk;   // compile-time tile width the kernel is instantiated for
  int count;   // number of chunks
  int padded;  // chunk * count
};
static inline Chunking pick_chunk(int n) {
  static const int opts[5] = {2, 4, 8, 10, 16};
  Chunking c;
  if (n <= 16) {
    for (int i = 0; i < 5; ++i)
      if (opts[i] >= n) { c.chunk = opts[i]; c.count = 1; c.padded = opts[i]; return c; }
  }
  int best = 16; int64_t best_pad = round_up(n, 16);
  for (int i = 3; i >= 2; --i) {  // 10, 8
    int64_t p = round_up(n, opts[i]);
    if (p < best_pad) { best_pad = p; best = opts[i]; }
  }
  c.chunk = best; c.padded = (int)best_pad; c.count = c.padded / best;
  return c;
}

// ---- device side -----------------------------------------------------------------------------
constexpr float kLogSqrt2Pi = 0.918938533204672741780329736406f;

// torch.nan_to_num with default arguments: NaN -> 0, +-inf -> +-FLT_MAX
__device__ __forceinline__ float nan_to_num(float v) {
  if (v != v) return 0.f;
  if (fabsf(v) > FLT_MAX) return copysignf(FLT_MAX, v);
  return v;
}

// read N consecutive floats (N even; 16-byte aligned when N % 4 == 0, else 8-byte) through the
// read-only path.  With a warp-uniform address this is a broadcast: one sector per request.
template <int N>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
    static_assert(N % 2 == 0, "row width must be even");
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      float2 t = __ldg(reinterpret_cast<const float2*>(p) + i);
      v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
  }
}

// same, from shared memory
template <int N>
__device__ __forceinline__ void load_row_smem(const float* p, float (&v)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      float4 t = reinterpret_cast<const float4*>(p)[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      float2 t = reinterpret_cast<const float2*>(p)[i];
      v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dpk
