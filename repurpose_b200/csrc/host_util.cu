#include <stdlib.h>

#include "host_util.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace rp {

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
}  // namespace

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

int num_sms() {
  static int sms_of[kMaxDevices];  // 0 = not queried yet
  int& sms = sms_of[current_device()];
  if (sms <= 0) {
    int dev = 0, n = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      set_last_error("no CUDA device available");
      (void)cudaGetLastError();
      return 0;
    }
    if (major != 10) {
      set_last_error("device compute capability %d.x is not sm_100 (B200 required)", major);
      return 0;
    }
    sms = n;
  }
  return sms;
}

int make_tensor_map(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return RP_ERR_NO_DEVICE;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(out, dtype, cuuint32_t(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (CUresult %d): rank=%d dims=[%llu,%llu,%llu] "
                   "box=[%u,%u,%u] stride0=%llu",
                   int(r), rank, (unsigned long long)dims[0],
                   (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0,
                   rank > 2 ? box[2] : 0, (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return RP_ERR_CUDA;
  }
  return RP_OK;
}


bool pdl_enabled() {
  static const bool on = !(getenv("RP_PDL") && atoi(getenv("RP_PDL")) == 0);
  return on;
}

}  // namespace rp
