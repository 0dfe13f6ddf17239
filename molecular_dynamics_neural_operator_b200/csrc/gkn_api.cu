// Housekeeping half of the C ABI: versioning, error strings, launch accounting, device queries and
// the TMA tensor-map encoder (driver entry point resolved at run time, so the library has no
// link-time dependency on libcuda and loads on a CPU-only box for the symbol-export test).
#include "gkn_internal.h"

#include <cudaTypedefs.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/gkn_b200.h"

namespace gkn {

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

int set_error(const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return GKN_ERR_INVALID;
}

int set_cuda_error(const char* what, cudaError_t err) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(err),
             cudaGetErrorString(err));
    return GKN_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int device_sm_count(int* sms) {
    static int cached[64] = {0};
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return set_cuda_error("cudaGetDevice", err);
    if (dev < 0 || dev >= 64) return set_error("device ordinal out of range");
    if (cached[dev] == 0) {
        int v = 0;
        err = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (err != cudaSuccess) return set_cuda_error("cudaDeviceGetAttribute(SM count)", err);
        cached[dev] = v;
    }
    *sms = cached[dev];
    return 0;
}

int make_tensor_map_bf16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows,
                            uint32_t box_cols, uint32_t box_rows) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault,
                                                  &qres);
        if (err == cudaSuccess && qres == cudaDriverEntryPointSuccess) {
            encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
        }
    });
    if (encode == nullptr) return set_error("cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) {
        return set_error("TMA operand base address must be 16-byte aligned");
    }
    if ((cols * 2) % 16 != 0) return set_error("TMA operand row pitch must be a multiple of 16 B");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
        return set_error(buf);
    }
    return 0;
}

}  // namespace gkn

extern "C" int gkn_version(void) { return GKN_B200_VERSION; }

extern "C" const char* gkn_last_error(void) { return gkn::g_err; }

extern "C" long long gkn_launch_count(void) {
    return gkn::g_launches.load(std::memory_order_relaxed);
}

extern "C" int gkn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return gkn::set_cuda_error("cudaGetDevice", err);
    int v = 0;
    if (sm_count) {
        int rc = gkn::device_sm_count(&v);
        if (rc) return rc;
        *sm_count = v;
    }
    if (cc_major) {
        err = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev);
        if (err != cudaSuccess) return gkn::set_cuda_error("cudaDeviceGetAttribute", err);
        *cc_major = v;
    }
    if (cc_minor) {
        err = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev);
        if (err != cudaSuccess) return gkn::set_cuda_error("cudaDeviceGetAttribute", err);
        *cc_minor = v;
    }
    return 0;
}
