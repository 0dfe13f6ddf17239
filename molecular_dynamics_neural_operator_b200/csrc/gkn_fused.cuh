// Shared device helpers of the fused producer/consumer kernels (gkn_node_fused.cu,
// gkn_node_fused_bwd.cu): the work-item record, cross-CTA counters, epilogue staging primitives.
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

namespace gkn {
namespace fused {

// One work item = two int4 words; the meaning of r0..r3 depends on the kind (see the kernels).
//   a = (kind, r0, r1, r2)      b = (wait_idx, wait_target, signal_idx, r3)
// wait_idx >= 0: the item waits until flags[wait_idx] >= wait_target; signal_idx >= 0: each of the
// four epilogue warps adds 1 to flags[signal_idx] when the item's output is in memory.
struct Item {
    int4 a, b;
};

__device__ __forceinline__ Item load_item(const Item* items, int64_t v) {
    Item it;
    const int4* p = reinterpret_cast<const int4*>(items + v);
    it.a = __ldg(p);
    it.b = __ldg(p + 1);
    return it;
}

// Software pipeline over a CTA's item queue.  An item record is a global load (an L2 round trip);
// loading "the next one" at the top of an iteration and branching on its kind in the following
// iteration stalls a role for that whole latency whenever an iteration is short, and anything
// derived from the record by further loads (the row pointers of a D item) adds a second, dependent
// latency.  ncu's source page showed the single-thread roles 60 % stalled on exactly these loads.
// So records are loaded kItemDepth items ahead and rotated through registers; derived loads are
// issued for the record that becomes current in the NEXT iteration, from a record that arrived
// iterations ago - nothing a role branches on is younger than one full iteration.
constexpr int kItemDepth = 3;

struct ItemPipe {
    Item q[kItemDepth];
    const Item* items;
    int64_t v_next, v_end, vstep;

    __device__ __forceinline__ static Item nop() {
        Item it;
        it.a = make_int4(2 /* no-op kind in both kernels */, 0, 0, 0);
        it.b = make_int4(-1, 0, -1, 0);
        return it;
    }
    __device__ __forceinline__ Item fetch() {
        const Item it = v_next < v_end ? load_item(items, v_next) : nop();
        v_next += vstep;
        return it;
    }
    __device__ __forceinline__ void init(const Item* it, int64_t v0, int64_t v1, int64_t step) {
        items = it;
        v_next = v0;
        v_end = v1;
        vstep = step;
#pragma unroll
        for (int i = 0; i < kItemDepth; ++i) q[i] = fetch();
    }
    // current record; afterwards q[0] is the record of the next iteration (already resident)
    __device__ __forceinline__ Item pop() {
        const Item it = q[0];
#pragma unroll
        for (int i = 0; i + 1 < kItemDepth; ++i) q[i] = q[i + 1];
        q[kItemDepth - 1] = fetch();
        return it;
    }
    __device__ __forceinline__ const Item& peek() const { return q[0]; }
};

__device__ __forceinline__ int32_t ld_acquire(const int32_t* p) {
    int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// counter += 1 with release semantics at GPU scope: everything this thread has observed (its own
// stores, and - after a __syncwarp - those of the other lanes of its warp) is visible to whoever
// acquires the new value.  One MEMBAR.ALL.GPU in one lane; __threadfence() in all 32 lanes is a
// sequentially-consistent fence plus an L1 invalidation each (measured: +40 % on producer tiles).
__device__ __forceinline__ void red_release_add(int32_t* p, int32_t v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// orders this thread's later async-proxy (TMA) accesses to global memory after the generic-proxy
// writes it has just observed through an acquire load
__device__ __forceinline__ void fence_proxy_async_global() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

__device__ __forceinline__ int32_t ld_relaxed(const int32_t* p) {
    int32_t v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Spin until a counter written by other CTAs reaches `target` (watchdog: trap instead of hanging).
// The polls are relaxed loads and ONE acquire fence follows the successful one: an acquire load
// per poll invalidates the SM's L1 every time (CCTL.IVALL) and with it the item records and row
// pointers every role of the CTA keeps there.
__device__ __forceinline__ void wait_counter(const int32_t* flag, int32_t target) {
    uint32_t spins = 0;
    while (ld_relaxed(flag) < target) {
        __nanosleep(64);
        if (++spins > (1u << 25)) __trap();
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(addr)
                 : "memory");
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace fused
}  // namespace gkn
