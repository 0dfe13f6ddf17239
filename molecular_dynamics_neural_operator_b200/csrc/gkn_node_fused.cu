// Fused producer/consumer kernels of the node-first factorisation (tcgen05 / TMEM / TMA, sm_100a).
//
// gkn_node_gemm.cu computes NNConv_old.message (reference graph_kernel.py:200-202) as
//     U[n][o, k] = sum_i x[n, i] * W3[i*w + o, k]        (dense GEMM, once per node)
//     msg[e, o]  = sum_k h[e, k] * U[src e][o, k]         (GEMM grouped by source node)
// with U ([N, w, K] bf16: 128 KB per node at w = 64, K = 1024) written to HBM by a library GEMM and
// read back by the grouped kernel - about 40 % of the traffic of an application.  Here both stages
// are work items of ONE persistent kernel: "P" items produce [128 nodes x 256 columns] tiles of U
// with the tensor core (A = x tile, B = W3 read in place as an MN-major operand), "C" items are the
// 128-row message tiles that consume U[n] through TMA as before.  The item list is ordered (by the
// host) so that the tiles of a node block are produced shortly before the units of that block
// consume them: U crosses the L2, not HBM.  A C item waits on a per-node-block counter that the
// epilogue warps of the P items release (st.global -> __threadfence -> atomicAdd; the consumer's TMA
// thread polls with ld.acquire.gpu and issues fence.proxy.async before its loads).  In "ring" mode
// U lives in a small ring of node blocks that is overwritten as soon as its consumers are done
// (a second counter, released by the C epilogues), so U never needs HBM capacity or write-back
// bandwidth at all; in "keep" mode it is written to a full-size buffer that the backward reuses.
//
// All CTAs of the launch are co-resident (grid <= SMs, one CTA per SM), so a CTA that spins on a
// counter only ever waits for CTAs that are running; every CTA processes its items in list order
// and P items wait (ring mode) only on C items that precede them in every CTA's order, so the
// waits cannot form a cycle.
#include <stdio.h>
#include <stdlib.h>

#include "gkn_internal.h"
#include "gkn_ptx.cuh"
#include "gkn_fused.cuh"

#include <cuda_bf16.h>

namespace gkn {

namespace {

constexpr int kM = 128;           // accumulator rows per item
constexpr int kKB = 64;           // reduction elements per k-block (one 128-byte swizzle row)
constexpr int kUmmaK = 16;
// TWO independent in-order item streams share one CTA (twelve warps):
//   P stream  warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue    (producer tiles of U)
//   C stream  warp 6 TMA producer, warp 7 MMA issuer, warps 8-11 epilogue   (message units)
// Every warp walks the CTA's whole queue and skips the other kind; the only coupling between the
// streams is the per-node-block counters.  (With one producer thread and one issuer thread for
// both kinds, an item that was not ready held up the items of the other kind queued behind it.)
constexpr int kThreads = 12 * 32;
constexpr int kPAccCols = 256;    // TMEM: one P accumulator [0, 256), then the C accumulators
constexpr int kABytes = kM * kKB * 2;        // 16 KB: one K-major A box (128 rows x 64)
constexpr int kMnBlockBytes = 64 * kKB * 2;  // 8 KB: one MN-major box (64 K-rows x 64 MN elements)
constexpr int kPCols = 256;       // columns of U per P tile

enum { IT_P = 0, IT_C = 1, IT_NOP = 2 };

// Item fields (fused::Item, gkn_fused.cuh):
//   P: r0 = first node (row of x), r1 = first column of U (multiple of 256), r2 = first row of the
//      U buffer this tile is written to, r3 = valid rows (<= 128).
//      wait (ring mode, before the stores): flags[wait_idx] >= wait_target; wait_idx < 0: none
//   C: r0 = first padded edge row, r1 = rows to store (<= 128), r2 = row of the U buffer holding
//      the unit's source node.  wait (before the loads of U): flags[wait_idx] >= wait_target
using fused::Item;
using fused::load_item;
using fused::ItemPipe;
using fused::wait_counter;
using fused::fence_proxy_async_global;
using fused::sts128;
using fused::lds128;
using fused::pack_bf16;
using fused::red_release_add;

// Shared memory.  P stream: one 48 KB fill buffer (x tile 16 KB + four 8 KB boxes of W3) - the
// tensor-core work of a tile is tiny, the stream is paced by its epilogue.  C stream: a ring of
// stages holding one k-block each (hs 16 KB + U[n] W x 128 B): six 24 KB stages at w <= 64, four
// 32 KB stages at w = 128.
template <int W>
struct FCfg {
    static constexpr int kWp = W < 64 ? 64 : W;             // reduction width of P (x zero-padded)
    static constexpr int kPFills = kWp / kKB;               // fills per P tile
    static constexpr int kPBytes = kABytes + 4 * kMnBlockBytes;   // 48 KB
    static constexpr int kCBBytes = W * kKB * 2;            // U[n] k-block: W rows x 64
    static constexpr int kCKbBytes = kABytes + kCBBytes;    // one C k-block
    static constexpr int kStageBytes = W <= 64 ? 24 * 1024 : 32 * 1024;
    static constexpr int kStages = W <= 64 ? 6 : 4;
    static constexpr int kCSlots = 256 / W > 4 ? 4 : 256 / W;    // C accumulators in TMEM
    static constexpr int kSmemBytes = kPBytes + kStages * kStageBytes + 1024 + 256 + 2 * 4 * 4096;
    static_assert(kCKbBytes <= kStageBytes, "a C k-block must fit one stage");
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory per CTA");
};

struct FusedParams {
    const Item* __restrict__ items;
    int64_t num_items;
    int32_t* flags;
    __nv_bfloat16* __restrict__ u;  // [u_rows, W*K]
    float* __restrict__ msg;        // [E_pad, W]
    int32_t K;
    int32_t tune;  // GKN_FUSED_TUNE bit 0 (experiment): plain (unhinted) loads of hs
};

// tmap_x : x bf16 [N, Wp]            box 64 x 128   (A of P, K-major)
// tmap_w3: W3r bf16 [Wp, W*K]        box 64 x 64    (B of P, MN-major: the columns are contiguous)
// tmap_hs: hs bf16 [E_pad, K]        box 64 x 128   (A of C, K-major)
// tmap_u : U bf16 [u_rows*W, K]      box 64 x W     (B of C, K-major)
template <int W>
__global__ void __launch_bounds__(kThreads, 1)
node_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x,
                      const __grid_constant__ CUtensorMap tmap_w3,
                      const __grid_constant__ CUtensorMap tmap_hs,
                      const __grid_constant__ CUtensorMap tmap_u, const FusedParams p) {
    using C = FCfg<W>;
    constexpr int kStages = C::kStages;
    constexpr int kStageBytes = C::kStageBytes;
    constexpr int kCSlots = C::kCSlots;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // P fill buffer
    const uint32_t cs_base = smem_base + C::kPBytes;                     // C ring
    const uint32_t bar_base = cs_base + kStages * kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto cfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto cempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + kCSlots + s); };
    const uint32_t pfull_bar = bar_base + 8u * (2 * kStages + 2 * kCSlots);   // P fill buffer
    const uint32_t pempty_bar = pfull_bar + 8u;
    const uint32_t tfull_bar = pfull_bar + 16u;                                // P accumulator
    const uint32_t tempty_bar = pfull_bar + 24u;
    const uint32_t tmem_ptr_smem = pfull_bar + 32u;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t v0 = blockIdx.x, v1 = p.num_items, vstep = gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < kCSlots; ++s) {
            mbar_init(cfull_bar(s), 1);
            mbar_init(cempty_bar(s), 4);
        }
        mbar_init(pfull_bar, 1);
        mbar_init(pempty_bar, 1);
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 4);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w3);
        tma_prefetch_desc(&tmap_hs);
        tma_prefetch_desc(&tmap_u);
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

    const int nkb = p.K / kKB;  // k-blocks of a C item

    // 32 rows x 128 B of this warp's staging tile -> global, 8 lanes per row (whole 128 B lines)
    auto store_rows = [&](uint32_t stage_smem, const uint32_t (&w)[32], uint8_t* gdst, int64_t pitch,
                          int rows_here) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            sts128(stage_smem + lane * 128 + ((c ^ (lane & 7)) << 4),
                   make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = j * 4 + (lane >> 3), c = lane & 7;
            const uint4 t = lds128(stage_smem + r * 128 + ((c ^ (r & 7)) << 4));
            if (r < rows_here) *reinterpret_cast<uint4*>(gdst + r * pitch + c * 16) = t;
        }
        __syncwarp();
    };

    if (warp == 0) {
        // ===================== P stream: TMA producer =====================
        if (lane == 0) {
            uint32_t n_fill = 0;
            ItemPipe pipe;
            pipe.init(p.items, v0, v1, vstep);
            for (int64_t v = v0; v < v1; v += vstep) {
                const Item it = pipe.pop();
                if (it.a.x != IT_P) continue;
#pragma unroll
                for (int f = 0; f < C::kPFills; ++f, ++n_fill) {
                    mbar_wait(pempty_bar, (n_fill & 1u) ^ 1u);
                    mbar_arrive_expect_tx(pfull_bar, C::kPBytes);
                    tma_load_2d(smem_base, &tmap_x, pfull_bar, f * kKB, it.a.y);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        tma_load_2d(smem_base + kABytes + j * kMnBlockBytes, &tmap_w3, pfull_bar,
                                    it.a.z + j * 64, f * kKB);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== P stream: MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_p = umma_idesc_bf16(kM, kPCols, 0, 1);
            uint32_t n_fill = 0, n_p = 0;
            ItemPipe pipe;
            pipe.init(p.items, v0, v1, vstep);
            for (int64_t v = v0; v < v1; v += vstep) {
                const Item it = pipe.pop();
                if (it.a.x != IT_P) continue;
                mbar_wait(tempty_bar, (n_p & 1u) ^ 1u);
                tc_fence_after();
#pragma unroll
                for (int f = 0; f < C::kPFills; ++f, ++n_fill) {
                    mbar_wait(pfull_bar, n_fill & 1u);
                    tc_fence_after();
                    const uint64_t adesc = umma_smem_desc(smem_base, 16, 1024, kSwizzle128B);
                    const uint64_t bdesc =
                        umma_smem_desc(smem_base + kABytes, kMnBlockBytes, 1024, kSwizzle128B);
#pragma unroll
                    for (int k = 0; k < kKB / kUmmaK; ++k) {
                        // K-major A: 16 elements = 32 B; MN-major B: 16 K-rows = 2048 B
                        umma_bf16_ss(tmem_base, adesc + 2ull * k, bdesc + 128ull * k, idesc_p,
                                     (f | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(pempty_bar);
                }
                umma_commit(tfull_bar);
                ++n_p;
            }
        }
    } else if (warp < 6) {
        // ===================== P stream: epilogue (warps 2..5) =====================
        const int q = warp & 3;           // TMEM lane quarter this warp may access
        const uint32_t lane_addr = uint32_t(q * 32) << 16;
        const uint32_t stage_smem = bar_base + 256 + q * 4096;  // per-warp 32 x 128 B staging tile
        const int64_t u_pitch = int64_t(W) * p.K * 2;           // bytes per row of U
        uint32_t n_p = 0;
        // The counter of a producer tile moves with release semantics after the tile's stores, and
        // a release right behind the stores waits for their full round trip.  So the update of
        // tile t is issued when this warp next has to wait anyway or is about to store tile t+1.
        int32_t pending = -1;
        auto flush = [&]() {
            if (pending >= 0) {
                __syncwarp();
                if (lane == 0) red_release_add(p.flags + pending, 1);
                pending = -1;
            }
        };
        ItemPipe pipe;
        pipe.init(p.items, v0, v1, vstep);
        for (int64_t v = v0; v < v1; v += vstep) {
            const Item it = pipe.pop();
            if (it.a.x != IT_P) continue;
            const uint32_t tphase = n_p & 1u;
            ++n_p;
            // never block with a counter update outstanding: somebody may be waiting for it
            uint32_t ready = 0;
            if (lane == 0) ready = mbar_try_wait(tfull_bar, tphase);
            ready = __shfl_sync(0xffffffffu, ready, 0);
            if (!ready || it.b.x >= 0) flush();
            if (lane == 0) {
                if (!ready) mbar_wait_backoff(tfull_bar, tphase, 32);
                // ring mode: the slot this tile overwrites must have been consumed
                if (it.b.x >= 0) wait_counter(p.flags + it.b.x, it.b.y);
            }
            __syncwarp();
            tc_fence_after();
            const uint32_t tcol = tmem_base + lane_addr;
            uint8_t* gbase = reinterpret_cast<uint8_t*>(p.u) + int64_t(it.a.w + q * 32) * u_pitch +
                             int64_t(it.a.z) * 2;
            const int rows_here = it.b.w - q * 32;
#pragma unroll 1
            for (int piece = 0; piece < kPCols / 64; ++piece) {
                uint32_t w[32];
                float f0[32], f1[32];
                tmem_ld32(tcol + piece * 64, f0);
                tmem_ld32(tcol + piece * 64 + 32, f1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    w[j] = pack_bf16(f0[2 * j], f0[2 * j + 1]);
                    w[16 + j] = pack_bf16(f1[2 * j], f1[2 * j + 1]);
                }
                if (piece == kPCols / 64 - 1) {  // accumulator fully read: hand it back now
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar);
                }
                if (piece == 0) flush();  // the previous tile's counter, before this tile's stores
                store_rows(stage_smem, w, gbase + piece * 128, u_pitch, rows_here);
            }
            pending = it.b.z;
        }
        flush();
    } else if (warp == 6) {
        // ===================== C stream: TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint64_t pol_stream = l2_policy_evict_first();
            ItemPipe pipe;
            pipe.init(p.items, v0, v1, vstep);
            for (int64_t v = v0; v < v1; v += vstep) {
                const Item it = pipe.pop();
                if (it.a.x != IT_C) continue;
                bool ready = it.b.x < 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = cs_base + stage * kStageBytes;
                    mbar_arrive_expect_tx(full_bar(stage), C::kCKbBytes);
                    // the edge hidden state does not depend on this launch: its load goes out
                    // before the counter is polled
                    if (p.tune & 1) {
                        tma_load_2d(sa, &tmap_hs, full_bar(stage), kb * kKB, it.a.y);
                    } else {
                        tma_load_2d_hint(sa, &tmap_hs, full_bar(stage), kb * kKB, it.a.y, pol_stream);
                    }
                    if (!ready) {
                        wait_counter(p.flags + it.b.x, it.b.y);
                        fence_proxy_async_global();
                        ready = true;
                    }
                    tma_load_2d(sa + kABytes, &tmap_u, full_bar(stage), kb * kKB, it.a.w * W);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 7) {
        // ===================== C stream: MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_c = umma_idesc_bf16(kM, W, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t n_c = 0;
            ItemPipe pipe;
            pipe.init(p.items, v0, v1, vstep);
            for (int64_t v = v0; v < v1; v += vstep) {
                const Item it = pipe.pop();
                if (it.a.x != IT_C) continue;
                const uint32_t cs = n_c % kCSlots, cphase = (n_c / kCSlots) & 1u;
                ++n_c;
                mbar_wait(cempty_bar(cs), cphase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + kPAccCols + cs * W;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = cs_base + stage * kStageBytes;
                    const uint64_t adesc = umma_smem_desc(sa, 16, 1024, kSwizzle128B);
                    const uint64_t bdesc = umma_smem_desc(sa + kABytes, 16, 1024, kSwizzle128B);
#pragma unroll
                    for (int k = 0; k < kKB / kUmmaK; ++k) {
                        umma_bf16_ss(tmem_d, adesc + 2ull * k, bdesc + 2ull * k, idesc_c,
                                     (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(cfull_bar(cs));
            }
        }
    } else {
        // ===================== C stream: epilogue (warps 8..11) =====================
        const int q = warp & 3;
        const uint32_t lane_addr = uint32_t(q * 32) << 16;
        const uint32_t stage_smem = bar_base + 256 + (4 + q) * 4096;
        uint32_t n_c = 0;
        ItemPipe pipe;
        pipe.init(p.items, v0, v1, vstep);
        for (int64_t v = v0; v < v1; v += vstep) {
            const Item it = pipe.pop();
            if (it.a.x != IT_C) continue;
            const uint32_t cs = n_c % kCSlots, cphase = (n_c / kCSlots) & 1u;
            ++n_c;
            if (lane == 0) mbar_wait_backoff(cfull_bar(cs), cphase, 32);
            __syncwarp();
            tc_fence_after();
            // a consumer's counter only says "U[n] has been read" (every MMA of the item is
            // complete): no data is published, so it moves at once and without a release
            if (it.b.z >= 0 && lane == 0) atomicAdd(p.flags + it.b.z, 1);
            const uint32_t tcol = tmem_base + lane_addr + kPAccCols + cs * W;
            uint8_t* gbase = reinterpret_cast<uint8_t*>(p.msg + int64_t(it.a.y + q * 32) * W);
            const int rows_here = it.a.z - q * 32;
#pragma unroll
            for (int piece = 0; piece < W / 32; ++piece) {
                uint32_t w[32];
                float f[32];
                tmem_ld32(tcol + piece * 32, f);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(f[j]);
                if (piece == W / 32 - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(cempty_bar(cs));
                }
                store_rows(stage_smem, w, gbase + piece * 128, W * 4, rows_here);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int W>
int launch_fused_fwd(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& th,
                     const CUtensorMap& tu, const FusedParams& p, int32_t want_grid,
                     cudaStream_t stream) {
    auto kern = node_fused_fwd_kernel<W>;
    constexpr int smem = FCfg<W>::kSmemBytes;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(node_fused_fwd)", err);
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    // co-residency of the whole grid is what makes the in-kernel waits safe: one CTA per SM
    if (want_grid > sms) {
        return set_error("gkn_node_fused_fwd_bf16: the item list was dealt for more CTAs than "
                         "this device has SMs");
    }
    int64_t grid = want_grid > 0 ? want_grid : (p.num_items < sms ? p.num_items : sms);
    if (grid < 1) return 0;
    kern<<<dim3(unsigned(grid)), dim3(kThreads), smem, stream>>>(tx, tw, th, tu, p);
    GKN_CHECK_LAUNCH("node_fused_fwd launch");
    return 0;
}

}  // namespace

}  // namespace gkn

extern "C" int gkn_node_fused_fwd_bf16(const void* x_bf16, const void* w3r_bf16, const void* hs_bf16,
                                       void* u_bf16, const int32_t* items, int64_t num_items,
                                       int32_t* flags, float* msg, int64_t N, int64_t E_pad,
                                       int64_t u_rows, int32_t K, int32_t W, int32_t grid,
                                       void* stream) {
    using namespace gkn;
    if (num_items <= 0) return 0;
    if (x_bf16 == nullptr || w3r_bf16 == nullptr || u_bf16 == nullptr || items == nullptr ||
        flags == nullptr) {
        return set_error("gkn_node_fused_fwd_bf16: NULL pointer argument");
    }
    if (W != 32 && W != 64 && W != 128) {
        return set_error("gkn_node_fused_fwd_bf16: width must be 32, 64 or 128");
    }
    if (K <= 0 || K % 256 != 0) {
        return set_error("gkn_node_fused_fwd_bf16: kernel width K must be a positive multiple of 256");
    }
    if (E_pad < 0 || E_pad % 16 != 0 || E_pad >= (int64_t(1) << 31) - 256 || N <= 0 ||
        u_rows <= 0 || u_rows * W >= (int64_t(1) << 31)) {
        return set_error("gkn_node_fused_fwd_bf16: bad sizes");
    }
    if ((E_pad > 0) != (hs_bf16 != nullptr && msg != nullptr)) {
        return set_error("gkn_node_fused_fwd_bf16: hs and msg must be given exactly when E_pad > 0");
    }
    const int Wp = W < 64 ? 64 : W;
    CUtensorMap tx, tw, th, tu;
    int rc = make_tensor_map_bf16_2d(&tx, x_bf16, uint64_t(Wp), uint64_t(N), kKB, kM);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tw, w3r_bf16, uint64_t(W) * K, uint64_t(Wp), 64, kKB);
    if (rc) return rc;
    // without edges the launch is a plain dense GEMM (P items only): the maps of the consumer side
    // still have to be valid descriptors
    rc = make_tensor_map_bf16_2d(&th, E_pad > 0 ? hs_bf16 : u_bf16, uint64_t(K),
                                 E_pad > 0 ? uint64_t(E_pad) : uint64_t(u_rows) * W, kKB, kM);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tu, u_bf16, uint64_t(K), uint64_t(u_rows) * W, kKB, W);
    if (rc) return rc;
    FusedParams p{};
    p.items = reinterpret_cast<const Item*>(items);
    p.num_items = num_items;
    p.flags = flags;
    p.u = static_cast<__nv_bfloat16*>(u_bf16);
    p.msg = msg;
    p.K = K;
    {
        static int tune = -1;
        if (tune < 0) {
            const char* e = getenv("GKN_FUSED_TUNE");
            tune = e != nullptr ? atoi(e) : 0;
        }
        p.tune = tune;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (W) {
        case 32: return launch_fused_fwd<32>(tx, tw, th, tu, p, grid, st);
        case 64: return launch_fused_fwd<64>(tx, tw, th, tu, p, grid, st);
        default: return launch_fused_fwd<128>(tx, tw, th, tu, p, grid, st);
    }
}
