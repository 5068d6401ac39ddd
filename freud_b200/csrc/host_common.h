// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptor creation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace freud {

void set_error(const std::string& msg);  // stored per thread; read through freud_last_error()

#define FREUD_CHECK_CUDA(expr)                                                                     \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::freud::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
      return 1;                                                                                    \
    }                                                                                              \
  } while (0)

#define FREUD_REQUIRE(cond, msg)                                                                   \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      ::freud::set_error(std::string("freud_b200: ") + msg + " [" #cond "]");                      \
      return 2;                                                                                    \
    }                                                                                              \
  } while (0)

// 2-D row-major tensor [rows, cols] (cols contiguous) -> tiled map with a {128 bytes, box_rows} box and the
// 128-byte swizzle.  elem_bytes: 2 (bf16) or 4 (fp32 / tf32).  Returns 0 on success.
int make_tensor_map_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t row_pitch_elems,
                       int elem_bytes, int box_rows);

int sm_count();

}  // namespace freud
