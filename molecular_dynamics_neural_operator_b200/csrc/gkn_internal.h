// Internal host-side helpers shared by the .cu translation units (not part of the C ABI).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gkn {

// Error reporting: every extern "C" entry point returns 0 on success and a non-zero code after
// storing a message retrievable with gkn_last_error() (thread-local).
int set_error(const char* msg);
int set_cuda_error(const char* what, cudaError_t err);

// Kernel-launch accounting (gkn_launch_count): one tick per kernel launched by this library.
void count_launch(int n = 1);

// Cached per-device SM count of the current device.
int device_sm_count(int* sms);

// Encode a 2-D bf16 row-major tensor map [rows, cols] (cols contiguous) with a {box_cols, box_rows}
// box and 128-byte swizzle; out-of-bounds elements read as zero.
int make_tensor_map_bf16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows,
                            uint32_t box_cols, uint32_t box_rows);

inline int ceil_div(int64_t a, int64_t b) { return int((a + b - 1) / b); }

}  // namespace gkn

#define GKN_CHECK_LAUNCH(name)                                            \
    do {                                                                  \
        cudaError_t _e = cudaGetLastError();                              \
        if (_e != cudaSuccess) return gkn::set_cuda_error(name, _e);      \
        gkn::count_launch();                                              \
    } while (0)
