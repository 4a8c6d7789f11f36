// Host-side helpers shared by the kernel launchers: status codes, TMA descriptor (CUtensorMap)
// construction through the driver entry point (so the library has no link-time libcuda
// dependency and still dlopen()s on a box without a GPU).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/repurpose_b200.h"  // status codes RP_OK / RP_ERR_*

namespace rp {

void set_last_error(const char* fmt, ...);
const char* last_error();

#define RP_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::rp::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                    \
      return RP_ERR_CUDA;                                                          \
    }                                                                                    \
  } while (0)

#define RP_CHECK(cond, ...)               \
  do {                                    \
    if (!(cond)) {                        \
      ::rp::set_last_error(__VA_ARGS__);  \
      return RP_ERR_INVALID;        \
    }                                     \
  } while (0)

// dtype: CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 / FLOAT32.  All maps use SWIZZLE_128B; box inner extent
// must therefore be exactly 128 bytes.  dims/strides are innermost-first; strides[i] is the byte
// stride of dim i+1.
int make_tensor_map(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

inline int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, uint64_t cols,
                        uint64_t rows, uint64_t row_pitch_bytes, uint32_t box_cols,
                        uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {row_pitch_bytes};
  uint32_t box[2] = {box_cols, box_rows};
  return make_tensor_map(out, dtype, 2, base, dims, strides, box);
}
inline int make_tmap_3d(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, uint64_t cols,
                        uint64_t rows, uint64_t batch, uint64_t row_pitch_bytes,
                        uint64_t batch_pitch_bytes, uint32_t box_cols, uint32_t box_rows) {
  uint64_t dims[3] = {cols, rows, batch};
  uint64_t strides[2] = {row_pitch_bytes, batch_pitch_bytes};
  uint32_t box[3] = {box_cols, box_rows, 1};
  return make_tensor_map(out, dtype, 3, base, dims, strides, box);
}

// Function attributes (dynamic shared memory opt-in) and device properties are per device: launchers keep
// their "already configured" state per device so that one process may drive several GPUs.
constexpr int kMaxDevices = 64;
int current_device();  // clamped to [0, kMaxDevices)
int num_sms();
void count_launch(int n = 1);
int64_t launch_count();


// Programmatic dependent launch: a kernel launched through launch_pdl may become resident while the
// previous kernel of the stream is still draining; its prologue (barrier init, TMEM allocation,
// descriptor prefetch) then overlaps that tail.  Such kernels call pdl_wait() (ptx.cuh) before they
// touch global memory.  RP_PDL=0 falls back to plain stream order.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace rp
