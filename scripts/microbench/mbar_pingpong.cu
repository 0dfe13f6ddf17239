// Round-trip latency of an mbarrier hand-off between one "issuer" thread and W waiter warps
// (lane 0 polls, the rest of the warp parks at __syncwarp), the pattern used between the roles of
// the fused kernels.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mbar_pingpong mbar_pingpong.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ uint32_t test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}

__device__ __forceinline__ uint32_t try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return ok;
}

template <int MODE>  // 0 try_wait, 1 test_wait spin, 2 try_wait + nanosleep(32), 3 try_wait with a 20 us suspend hint
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    if (MODE == 0) { while (!try_wait(bar, parity)) {} }
    else if (MODE == 1) { while (!test_wait(bar, parity)) {} }
    else if (MODE == 2) { while (!try_wait(bar, parity)) __nanosleep(32); }
    else { while (!try_wait_hint(bar, parity, 20000u)) {} }
}

template <int MODE>
__global__ void pingpong(int waiters, int rounds, int stages, long long* out) {
    __shared__ uint64_t bars[16];  // go[s] = bars[s], back[s] = bars[8 + s]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[8 + s]), waiters); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64();
    if (warp == 0) {
        if (lane == 0) {
            // issuer: keeps `stages` hand-offs in flight (like an MMA thread with that many stages)
            for (int r = 0; r < rounds; ++r) {
                const int s = r % stages; const uint32_t ph = (r / stages) & 1u;
                if (r >= stages) wait<MODE>(smem_u32(&bars[8 + s]), ph ^ 1u);  // previous use acknowledged
                mbar_arrive(smem_u32(&bars[s]));
            }
            for (int r = rounds - stages; r < rounds; ++r) {
                if (r < 0) continue;
                wait<MODE>(smem_u32(&bars[8 + r % stages]), (r / stages) & 1u);
            }
        }
    } else if (warp <= waiters) {
        for (int r = 0; r < rounds; ++r) {
            const int s = r % stages; const uint32_t ph = (r / stages) & 1u;
            if (lane == 0) wait<MODE>(smem_u32(&bars[s]), ph);
            __syncwarp();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[8 + s]));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
    long long* d; cudaMalloc(&d, 8 * 148);
    for (int mode = 0; mode < 4; ++mode)
        for (int waiters : {1, 4, 16})
            for (int stages : {1, 4}) {
                const int rounds = 20000;
                for (int rep = 0; rep < 2; ++rep) {
                    if (mode == 0) pingpong<0><<<148, 32 * (waiters + 1)>>>(waiters, rounds, stages, d);
                    if (mode == 1) pingpong<1><<<148, 32 * (waiters + 1)>>>(waiters, rounds, stages, d);
                    if (mode == 2) pingpong<2><<<148, 32 * (waiters + 1)>>>(waiters, rounds, stages, d);
                    if (mode == 3) pingpong<3><<<148, 32 * (waiters + 1)>>>(waiters, rounds, stages, d);
                    cudaDeviceSynchronize();
                }
                long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                printf("mode=%d waiters=%2d stages=%d: %.0f cycles per hand-off round (err=%s)\n", mode, waiters, stages,
                       double(h[0]) / rounds, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
