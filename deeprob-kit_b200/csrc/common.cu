// common.cu -- error reporting and cached device attributes.
#include "common.cuh"

namespace dpk {

static thread_local char g_error[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

struct DevAttr { int sms; int smem; bool ok; };
static DevAttr g_attr[64];

static DevAttr& attr() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  DevAttr& a = g_attr[dev];
  if (!a.ok) {
    int sms = 0, smem = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    a.sms = sms > 0 ? sms : 148;
    a.smem = smem > 0 ? smem : 232448;  // no device (host-only planning): B200 opt-in limit
    a.ok = true;  // benign race: every thread writes the same values
  }
  return a;
}

// ---- profiling -------------------------------------------------------------------------------
}  // namespace dpk
#include <atomic>
#include <mutex>
#include <vector>
namespace dpk {
struct ProfRec { int cat; cudaEvent_t a, b; };
static std::atomic<long long> g_launches[CAT_COUNT];
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec*> g_prof_recs;

ProfScope::ProfScope(int cat, cudaStream_t st, int launches) : cat_(cat), st_(st), rec_(nullptr) {
  g_launches[cat].fetch_add(launches, std::memory_order_relaxed);
  if (g_prof_on.load(std::memory_order_relaxed)) {
    ProfRec* r = new ProfRec{cat, nullptr, nullptr};
    if (cudaEventCreate(&r->a) == cudaSuccess && cudaEventCreate(&r->b) == cudaSuccess) {
      cudaEventRecord(r->a, st);
      rec_ = r;
    } else {
      delete r;
    }
  }
}
ProfScope::~ProfScope() {
  if (rec_) {
    ProfRec* r = static_cast<ProfRec*>(rec_);
    cudaEventRecord(r->b, st_);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(r);
  }
}

int pdl_mode() {
  const char* v = getenv("DPK_PDL");      // read per call: tests and A/B runs switch it
  return (v && *v) ? (*v == '0' ? 0 : 1) : -1;
}

int sm_count() { return attr().sms; }
int max_dynamic_smem() { return attr().smem; }

}  // namespace dpk

extern "C" int dpk_abi_version(void) { return DPK_ABI_VERSION; }
extern "C" const char* dpk_last_error(void) { return dpk::g_error; }

// enable/disable CUDA-event bracketing of every kernel category (off by default: zero overhead)
extern "C" int dpk_profile_enable(int on) { dpk::g_prof_on.store(on ? 1 : 0); return DPK_OK; }

// Sum of elapsed milliseconds and number of kernel launches per category since the last read
// (arrays of `ncat` entries, see ProfCat in csrc/common.cuh); synchronises the recorded events.
extern "C" int dpk_profile_read(double* ms, int64_t* launches, int32_t ncat) {
  using namespace dpk;
  for (int c = 0; c < ncat; ++c) {
    if (ms) ms[c] = 0.0;
    if (launches) launches[c] = (c < CAT_COUNT) ? g_launches[c].exchange(0) : 0;
  }
  std::vector<ProfRec*> recs;
  { std::lock_guard<std::mutex> lk(g_prof_mu); recs.swap(g_prof_recs); }
  int rc = DPK_OK;
  for (ProfRec* r : recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r->b) != cudaSuccess || cudaEventElapsedTime(&t, r->a, r->b) != cudaSuccess)
      rc = set_error(DPK_E_CUDA, "profile event read failed");
    else if (ms && r->cat < ncat) ms[r->cat] += t;
    cudaEventDestroy(r->a); cudaEventDestroy(r->b);
    delete r;
  }
  return rc;
}

// ---- TMA tensor maps ---------------------------------------------------------------------------------
#include <cudaTypedefs.h>

#include "tc_common.cuh"
namespace dpk {
namespace tc {
int make_tensor_map_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                           uint32_t box_rows, uint32_t box_cols, int swizzle128) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
      return set_error(DPK_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return set_error(DPK_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
  return DPK_OK;
}
}  // namespace tc
}  // namespace dpk
