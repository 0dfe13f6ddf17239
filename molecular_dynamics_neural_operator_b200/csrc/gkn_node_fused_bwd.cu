// Fused backward of one NNConv application on the node-first path (tcgen05 / TMEM / TMA, sm_100a):
// the gradient w.r.t. U is produced and consumed inside ONE persistent kernel and crosses the L2
// only - no [N, w*K] tensor is written to or read from HBM, and the weight gradient of the last
// edge-MLP layer (the autograd of NNConv_old.message, reference graph_kernel.py:200-202, w.r.t.
// net.layers.4.weight) needs no library GEMM.
//
//   "D" items (producer)   dU[n][k, o] = sum_{e: src e = n} h[e, k] * gm[e, o]      (grouped GEMM;
//                          A = hs, B = gs read in place as MN-major operands, K = padded degree)
//                          written as bf16 into a ring of node blocks, 16 KB per item
//   "X" items (consumer)   every CTA owns a fixed slice of the w*K = 65536 columns (k, o) of dU.
//                          For node block b it reads its slice of the block ONCE and feeds two
//                          accumulators from the same shared-memory tiles:
//                            dW3k^T[col, i] += sum_n dU[n, col] * x[n, i]    (TMEM-resident for the
//                                              whole launch: 64 columns per 128 columns of dU)
//                            dx_part[n, i]   = sum_{col in slice} dU[n, col] * W3k[i, col]
//                          dx_part goes to a per-CTA partial buffer that gkn_sum_partials folds in
//                          a fixed order (deterministic, no atomics).
//
// Ordering is the host's: D items of block b+1 surround the X items of block b in the list, X items
// are aligned so that item index = CTA index (mod grid), every CTA runs its items in order.  An X
// item waits for the counter of its block (all D items stored), a D item of block b+R waits (ring
// of R blocks) for the X items of block b - always items that precede it in every CTA's order, and
// all CTAs are co-resident (grid = one CTA per SM), so the waits cannot form a cycle.
#include <stdio.h>
#include <stdlib.h>

#include "gkn_internal.h"
#include "gkn_ptx.cuh"
#include "gkn_fused.cuh"

namespace gkn {

namespace {

using fused::fence_proxy_async_global;
using fused::Item;
using fused::lds128;
using fused::load_item;
using fused::ItemPipe;
using fused::pack_bf16;
using fused::red_release_add;
using fused::sts128;
using fused::wait_counter;

constexpr int kM = 128;
constexpr int kKB = 64;
constexpr int kUmmaK = 16;
constexpr int kW = 64;             // node feature width served by this kernel
// TWO independent in-order item streams share one CTA (twelve warps):
//   D stream  warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue   (items of kind D)
//   X stream  warp 6 TMA producer, warp 7 MMA issuer, warps 8-11 epilogue  (items of kind X)
// Every warp walks the CTA's whole queue and skips the other kind.  With ONE producer and ONE
// issuer per CTA (the first version) a consumer item that was not ready held up the producer items
// queued behind it together with their HBM loads, and every node block cost a pipeline drain:
// producer items alone 0.76 ms, consumer items alone 0.51 ms, both 1.00 ms.  Now the only coupling
// between the streams is the two counters per node block.
constexpr int kThreads = 12 * 32;
// D stream: five 24 KB stages (one k-block each: two 8 KB boxes of hs + one of gs).  X stream: two
// 24 KB stages = one 128-column pair in flight (each stage one 64-column half: dU box 16 KB + W3k
// box 8 KB; the weight-gradient product addresses both dU boxes through one descriptor).
constexpr int kStageBytes = 24 * 1024;
constexpr int kStages = 5;    // D ring
constexpr int kXStages = 2;   // X ring
constexpr int kXbBytes = kM * kW * 2;          // 16 KB: x tile of a node block
constexpr int kDKbBytes = 24 * 1024;           // D k-block: hs 2 x 8 KB + gs 8 KB
static_assert(kDKbBytes == kStageBytes, "see the stage layout above");
constexpr int kMnBlockBytes = 64 * kKB * 2;    // 8 KB: MN-major box of 64 K-rows x 64 elements
constexpr int kDuBoxBytes = kM * kKB * 2;      // 16 KB: 128 node rows x 64 columns of dU
constexpr int kMaxPairs = 4;                   // 128-column pairs of dU per CTA
constexpr int kPairCols = 128;
// TMEM columns: [0, 256) dW3k^T accumulators (4 pairs x 64), [256, 320) dx, [320, 512) three D slots
constexpr uint32_t kTmemDw = 0, kTmemDx = 256, kTmemD = 320;
constexpr int kDSlots = 3;
constexpr int kSmemBytes =
    (kStages + kXStages) * kStageBytes + kXbBytes + 1024 + 256 + 2 * 4 * 4096;
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory per CTA");

enum { IT_D = 0, IT_X = 1, IT_NOP = 2 };

// Item fields (fused::Item):
//   D: r0 = node n, r1 = chunk c (128 k-rows of dU[n]), r2 = row of the dU buffer for node n
//      wait (ring, before the stores): flags[wait_idx] >= wait_target;  signal: flags[signal_idx]
//      (the host puts the wait on the first and the signal on the last D item a CTA has in a node
//      block: a release covers every earlier store of the warp that issues it)
//   X: one 128-column pair of the CTA's slice for one node block.  r0 = first node of the block
//      (row of x and of the dx partials), r1 = first row of the block in the dU buffer, r2 = valid
//      rows (<= 128), r3 = pair index q | first << 8 | last << 9.  The first pair of a block waits
//      (before the loads of dU), the last one signals "block consumed" and hands the dx tile to
//      the epilogue.  The pairs of a block are spread between the CTA's D items, so that their L2
//      traffic overlaps the D items' HBM stream.
struct BwdParams {
    const Item* __restrict__ items;
    int64_t num_items;
    int32_t* flags;
    const int32_t* __restrict__ pptr;   // padded group pointers [N+1]
    __nv_bfloat16* __restrict__ du;     // [du_rows, w*K]   (k, o) order
    float* __restrict__ dx_part;        // [parts, N, w] or NULL
    float* __restrict__ dwT;            // [w*K, w] fp32 or NULL
    int64_t N;
    int32_t K;
    int32_t accumulate_dw;              // dwT += (1) or = (0)
    int32_t tune;                       // GKN_FUSED_TUNE bit 0: plain (unhinted) loads of hs
};

struct DUnit {
    int32_t row0, nkb, last_steps;
};

__device__ __forceinline__ DUnit make_dunit(int32_t b, int32_t len) {  // len: multiple of 16
    DUnit u;
    u.row0 = b;
    u.nkb = (len + kKB - 1) / kKB;
    u.last_steps = (len - (u.nkb - 1) * kKB) / kUmmaK;
    return u;
}

// tmap_hs : hs bf16 [E_pad, K]          box 64 x 64    (A of D, MN-major)
// tmap_gs : gs bf16 [E_pad, w]          box 64 x 64    (B of D, MN-major)
// tmap_du : dU bf16 [du_rows, w*K]      box 64 x 128   (A of both X products)
// tmap_w3k: W3k bf16 [w, w*K]           box 64 x 64    (B of dx, K-major)
// tmap_x  : x bf16 [N, w]               box 64 x 128   (B of dW3k^T, MN-major)
__global__ void __launch_bounds__(kThreads, 1)
node_fused_bwd_kernel(const __grid_constant__ CUtensorMap tmap_hs,
                      const __grid_constant__ CUtensorMap tmap_gs,
                      const __grid_constant__ CUtensorMap tmap_du,
                      const __grid_constant__ CUtensorMap tmap_w3k,
                      const __grid_constant__ CUtensorMap tmap_x, const BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t xs_base = smem_base + kStages * kStageBytes;          // X ring
    const uint32_t xb_base = xs_base + kXStages * kStageBytes;           // x tile of a node block
    const uint32_t bar_base = xb_base + kXbBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto dfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto dempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + kDSlots + s); };
    const uint32_t xfull_bar = bar_base + 8u * (2 * kStages + 2 * kDSlots);       // x tile
    const uint32_t xempty_bar = xfull_bar + 8u;
    const uint32_t gfull_bar = xfull_bar + 16u;                                   // dx accumulator
    const uint32_t gempty_bar = xfull_bar + 24u;
    const uint32_t wdone_bar = xfull_bar + 32u;
    auto pfull_bar = [&](int s) { return xfull_bar + 40u + 8u * s; };             // X ring
    auto pempty_bar = [&](int s) { return xfull_bar + 40u + 8u * (kXStages + s); };
    const uint32_t tmem_ptr_smem = xfull_bar + 40u + 8u * 2 * kXStages;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t v0 = blockIdx.x, v1 = p.num_items, vstep = gridDim.x;
    const bool do_dx = p.dx_part != nullptr, do_dw = p.dwT != nullptr;

    // this CTA's slice of the 128-column pairs of dU
    const int pairs = kW * p.K / kPairCols;
    const int pbase = pairs / int(gridDim.x), prem = pairs % int(gridDim.x);
    const int np = pbase + (int(blockIdx.x) < prem ? 1 : 0);
    const int p0 = int(blockIdx.x) * pbase + (int(blockIdx.x) < prem ? int(blockIdx.x) : prem);

    // the padded row range of a D item's node: two loads from a table that consecutive items of a
    // CTA walk almost sequentially (L1 hits after the first touch of a line)
    // only the two loads: the arithmetic on them (make_dunit) is done where the unit is used, one
    // iteration later - a dependent instruction right here would stall on the loads at once
    auto decode_d = [&](int32_t n) { return make_int2(__ldg(p.pptr + n), __ldg(p.pptr + n + 1)); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < kDSlots; ++s) {
            mbar_init(dfull_bar(s), 1);
            mbar_init(dempty_bar(s), 4);
        }
        mbar_init(xfull_bar, 1);
        mbar_init(xempty_bar, 1);
        mbar_init(gfull_bar, 1);
        mbar_init(gempty_bar, 4);
        mbar_init(wdone_bar, 1);
        for (int s = 0; s < kXStages; ++s) {
            mbar_init(pfull_bar(s), 1);
            mbar_init(pempty_bar(s), 1);
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_hs);
        tma_prefetch_desc(&tmap_gs);
        tma_prefetch_desc(&tmap_du);
        tma_prefetch_desc(&tmap_w3k);
        tma_prefetch_desc(&tmap_x);
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

    // staging tiles of the two epilogue groups and a helper they share
    auto store_tile_f32 = [&](uint32_t stage_smem, uint32_t tcol, uint8_t* gbase, int rows_here,
                              bool accumulate) {
        // 128 x 64 fp32 accumulator -> global rows of 256 B, through the per-warp swizzled tile
#pragma unroll
        for (int piece = 0; piece < 2; ++piece) {
            float f[32];
            tmem_ld32(tcol + piece * 32, f);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                sts128(stage_smem + lane * 128 + ((c ^ (lane & 7)) << 4),
                       make_uint4(__float_as_uint(f[4 * c]), __float_as_uint(f[4 * c + 1]),
                                  __float_as_uint(f[4 * c + 2]), __float_as_uint(f[4 * c + 3])));
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = j * 4 + (lane >> 3), c = lane & 7;
                uint4 t = lds128(stage_smem + r * 128 + ((c ^ (r & 7)) << 4));
                if (r < rows_here) {
                    float4* gp = reinterpret_cast<float4*>(gbase + r * (kW * 4) + piece * 128 + c * 16);
                    if (accumulate) {
                        const float4 o = *gp;
                        t.x = __float_as_uint(__uint_as_float(t.x) + o.x);
                        t.y = __float_as_uint(__uint_as_float(t.y) + o.y);
                        t.z = __float_as_uint(__uint_as_float(t.z) + o.z);
                        t.w = __float_as_uint(__uint_as_float(t.w) + o.w);
                    }
                    *reinterpret_cast<uint4*>(gp) = t;
                }
            }
            __syncwarp();
        }
    };

    if (warp == 0) {
        // ===================== D stream: TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint64_t pol_stream = l2_policy_evict_first();
            ItemPipe pipe;
            pipe.init(p.items, v0, v1, vstep);
            int2 un = pipe.peek().a.x == IT_D ? decode_d(pipe.peek().a.y) : make_int2(0, 0);
            for (int64_t v = v0; v < v1; v += vstep) {
                const DUnit u = make_dunit(un.x, un.y - un.x);
                const Item it = pipe.pop();
                // row pointers of the NEXT item (its record arrived iterations ago): in flight
                // while this item is processed
                if (pipe.peek().a.x == IT_D) un = decode_d(pipe.peek().a.y);
                if (it.a.x != IT_D) continue;
                for (int kb = 0; kb < u.nkb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t st = smem_base + stage * kStageBytes;
                    mbar_arrive_expect_tx(full_bar(stage), kDKbBytes);
                    const int32_t r = u.row0 + kb * kKB;
                    if (p.tune & 1) {
                        tma_load_2d(st, &tmap_hs, full_bar(stage), it.a.z * kM, r);
                        tma_load_2d(st + kMnBlockBytes, &tmap_hs, full_bar(stage), it.a.z * kM + 64, r);
                    } else {
                        tma_load_2d_hint(st, &tmap_hs, full_bar(stage), it.a.z * kM, r, pol_stream);
                        tma_load_2d_hint(st + kMnBlockBytes, &tmap_hs, full_bar(stage),
                                         it.a.z * kM + 64, r, pol_stream);
                    }
                    tma_load_2d(st + 2 * kMnBlockBytes, &tmap_gs, full_bar(stage), 0, r);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== D stream: MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_mn = umma_idesc_bf16(kM, kW, 1, 1);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t n_d = 0;
            ItemPipe pipe;
            pipe.init(p.items, v0, v1, vstep);
            int2 un = pipe.peek().a.x == IT_D ? decode_d(pipe.peek().a.y) : make_int2(0, 0);
            for (int64_t v = v0; v < v1; v += vstep) {
                const DUnit u = make_dunit(un.x, un.y - un.x);
                const Item it = pipe.pop();
                if (pipe.peek().a.x == IT_D) un = decode_d(pipe.peek().a.y);
                if (it.a.x != IT_D) continue;
                // every D item takes a turn on the accumulator slots - also the one of an empty
                // group (no MMAs: the commit fires at once and the epilogue stores zeros)
                const uint32_t ds = n_d % kDSlots, dphase = (n_d / kDSlots) & 1u;
                ++n_d;
                mbar_wait(dempty_bar(ds), dphase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + kTmemD + ds * kW;
                for (int kb = 0; kb < u.nkb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t st = smem_base + stage * kStageBytes;
                    const uint64_t adesc = umma_smem_desc(st, kMnBlockBytes, 1024, kSwizzle128B);
                    const uint64_t bdesc =
                        umma_smem_desc(st + 2 * kMnBlockBytes, kMnBlockBytes, 1024, kSwizzle128B);
                    const int steps = kb == u.nkb - 1 ? u.last_steps : kKB / kUmmaK;
                    for (int k = 0; k < steps; ++k) {
                        // MN-major operands: a K-step of 16 rows advances the start by 2 KB
                        umma_bf16_ss(tmem_d, adesc + 128ull * k, bdesc + 128ull * k, idesc_mn,
                                     (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(dfull_bar(ds));
            }
        }
    } else if (warp < 6) {
        // ===================== D stream: epilogue (warps 2..5) =====================
        const int q4 = warp & 3;           // TMEM lane quarter this warp may access
        const uint32_t lane_addr = uint32_t(q4 * 32) << 16;
        const uint32_t stage_smem = bar_base + 256 + q4 * 4096;
        const int64_t du_pitch = int64_t(kW) * p.K;  // elements per row of the dU buffer
        uint32_t n_d = 0;
        // The counter update of a D item (a release, then the add) is deferred until this warp next
        // has to wait or is about to store the next item: a release right behind the stores waits
        // for their whole round trip.  It is never held across a blocking wait.
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
        int2 un = pipe.peek().a.x == IT_D ? decode_d(pipe.peek().a.y) : make_int2(0, 0);
        for (int64_t v = v0; v < v1; v += vstep) {
            const DUnit u = make_dunit(un.x, un.y - un.x);
            const Item it = pipe.pop();
            if (pipe.peek().a.x == IT_D) un = decode_d(pipe.peek().a.y);
            if (it.a.x != IT_D) continue;
            const uint32_t ds = n_d % kDSlots, dphase = (n_d / kDSlots) & 1u;
            ++n_d;
            __nv_bfloat16* out = p.du + int64_t(it.a.w) * du_pitch + int64_t(it.a.z) * kM * kW;
            if (it.b.x >= 0) {  // ring slot free?
                flush();
                if (lane == 0) wait_counter(p.flags + it.b.x, it.b.y);
                __syncwarp();
            }
            uint32_t ready = 0;
            if (lane == 0) ready = mbar_try_wait(dfull_bar(ds), dphase);
            ready = __shfl_sync(0xffffffffu, ready, 0);
            if (!ready) {
                flush();
                if (lane == 0) mbar_wait_backoff(dfull_bar(ds), dphase, 32);
            }
            __syncwarp();
            tc_fence_after();
            uint32_t w[32];
            if (u.nkb == 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) w[j] = 0u;
            } else {
                const uint32_t tcol = tmem_base + lane_addr + kTmemD + ds * kW;
                float f0[32], f1[32];
                tmem_ld32(tcol, f0);
                tmem_ld32(tcol + 32, f1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    w[j] = pack_bf16(f0[2 * j], f0[2 * j + 1]);
                    w[16 + j] = pack_bf16(f1[2 * j], f1[2 * j + 1]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dempty_bar(ds));
            flush();  // the previous item's counter, before this item's stores
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                sts128(stage_smem + lane * 128 + ((c ^ (lane & 7)) << 4),
                       make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
            }
            __syncwarp();
            // rows of dU[n][chunk] are 128 B and consecutive: 512 contiguous bytes per store
            uint8_t* gbase = reinterpret_cast<uint8_t*>(out + q4 * 32 * kW);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = j * 4 + (lane >> 3), c = lane & 7;
                const uint4 t = lds128(stage_smem + r * 128 + ((c ^ (r & 7)) << 4));
                *reinterpret_cast<uint4*>(gbase + r * 128 + c * 16) = t;
            }
            __syncwarp();
            pending = it.b.z;
        }
        flush();
    } else if (warp == 6) {
        // ===================== X stream: TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t n_x = 0;  // node blocks started so far
            ItemPipe pipe;
            pipe.init(p.items, v0, v1, vstep);
            for (int64_t v = v0; v < v1; v += vstep) {
                const Item it = pipe.pop();
                if (it.a.x != IT_X) continue;
                const int q = it.b.w & 0xff;
                const bool first = (it.b.w >> 8) & 1;
                if (first && do_dw) {
                    mbar_wait(xempty_bar, (n_x & 1u) ^ 1u);
                    mbar_arrive_expect_tx(xfull_bar, kXbBytes);
                    tma_load_2d(xb_base, &tmap_x, xfull_bar, 0, it.a.y);
                }
                if (first) ++n_x;
                if (it.b.x >= 0) {
                    wait_counter(p.flags + it.b.x, it.b.y);
                    fence_proxy_async_global();
                }
                const int32_t c0 = (p0 + q) * kPairCols;
#pragma unroll
                for (int j = 0; j < 2; ++j) {  // the X ring has exactly two stages: j is the stage
                    mbar_wait(pempty_bar(stage), phase ^ 1u);
                    const uint32_t st = xs_base + stage * kStageBytes;
                    mbar_arrive_expect_tx(pfull_bar(stage), kDuBoxBytes + kMnBlockBytes);
                    tma_load_2d(st, &tmap_du, pfull_bar(stage), c0 + j * 64, it.a.z);
                    tma_load_2d(st + kDuBoxBytes, &tmap_w3k, pfull_bar(stage), c0 + j * 64, 0);
                    if (++stage == kXStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 7) {
        // ===================== X stream: MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_mn = umma_idesc_bf16(kM, kW, 1, 1);  // dW3k^T product
            constexpr uint32_t idesc_k = umma_idesc_bf16(kM, kW, 0, 0);   // dx product
            uint32_t phase = 0;
            uint32_t n_x = 0;
            ItemPipe pipe;
            pipe.init(p.items, v0, v1, vstep);
            for (int64_t v = v0; v < v1; v += vstep) {
                const Item it = pipe.pop();
                if (it.a.x != IT_X) continue;
                const int q = it.b.w & 0xff;
                const bool first = (it.b.w >> 8) & 1, last = (it.b.w >> 9) & 1;
                const uint32_t xphase = n_x & 1u;
                const bool first_block = n_x == 0;
                if (first) {
                    mbar_wait(gempty_bar, xphase ^ 1u);  // the dx accumulator has been drained
                    if (do_dw) mbar_wait(xfull_bar, xphase);
                    tc_fence_after();
                }
                const uint32_t tmem_x = tmem_base + kTmemDx;
                const uint64_t xdesc = umma_smem_desc(xb_base, kMnBlockBytes, 1024, kSwizzle128B);
                const uint32_t st = xs_base;  // a pair fills the whole X ring: stage 0 then stage 1
                mbar_wait(pfull_bar(0), phase);
                mbar_wait(pfull_bar(1), phase);
                tc_fence_after();
                if (do_dw) {
                    // dW3k^T[128 cols x 64 i] += dU_tile^T . x_tile  (reduction: 128 nodes); the
                    // two 64-column boxes of dU are one stage (24 KB) apart
                    const uint64_t adesc = umma_smem_desc(st, kStageBytes, 1024, kSwizzle128B);
#pragma unroll
                    for (int k = 0; k < kM / kUmmaK; ++k) {
                        umma_bf16_ss(tmem_base + kTmemDw + q * kW, adesc + 128ull * k,
                                     xdesc + 128ull * k, idesc_mn, (!first_block || k != 0) ? 1u : 0u);
                    }
                }
                if (do_dx) {
                    // dx[128 nodes x 64 i] += dU_tile . W3k_tile^T  (reduction: 128 columns)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint64_t adesc = umma_smem_desc(st + j * kStageBytes, 16, 1024, kSwizzle128B);
                        const uint64_t bdesc =
                            umma_smem_desc(st + j * kStageBytes + kDuBoxBytes, 16, 1024, kSwizzle128B);
#pragma unroll
                        for (int k = 0; k < kKB / kUmmaK; ++k) {
                            umma_bf16_ss(tmem_x, adesc + 2ull * k, bdesc + 2ull * k, idesc_k,
                                         (!first || (j | k) != 0) ? 1u : 0u);
                        }
                    }
                }
                umma_commit(pempty_bar(0));
                umma_commit(pempty_bar(1));
                phase ^= 1u;
                if (last) {
                    if (do_dw) umma_commit(xempty_bar);
                    umma_commit(gfull_bar);
                    ++n_x;
                }
            }
            umma_commit(wdone_bar);
        }
    } else {
        // ===================== X stream: epilogue (warps 8..11) =====================
        const int q4 = warp & 3;
        const uint32_t lane_addr = uint32_t(q4 * 32) << 16;
        const uint32_t stage_smem = bar_base + 256 + (4 + q4) * 4096;
        uint32_t n_x = 0;
        ItemPipe pipe;
        pipe.init(p.items, v0, v1, vstep);
        for (int64_t v = v0; v < v1; v += vstep) {
            const Item it = pipe.pop();
            if (it.a.x != IT_X || !((it.b.w >> 9) & 1)) continue;  // only a block's last pair
            const uint32_t xphase = n_x & 1u;
            ++n_x;
            if (lane == 0) mbar_wait_backoff(gfull_bar, xphase, 32);
            __syncwarp();
            tc_fence_after();
            // every MMA of the block has completed: this CTA is done with the block's dU
            if (it.b.z >= 0 && lane == 0) atomicAdd(p.flags + it.b.z, 1);
            if (do_dx) {
                store_tile_f32(stage_smem, tmem_base + lane_addr + kTmemDx,
                               reinterpret_cast<uint8_t*>(
                                   p.dx_part + (int64_t(blockIdx.x) * p.N + it.a.y + q4 * 32) * kW),
                               it.a.w - q4 * 32, false);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(gempty_bar);
        }
        // ---- the weight-gradient accumulators of this CTA's column slice leave TMEM once
        if (do_dw && np > 0) {
            if (lane == 0) mbar_wait(wdone_bar, 0);
            __syncwarp();
            tc_fence_after();
            for (int q = 0; q < np; ++q) {
                store_tile_f32(stage_smem, tmem_base + lane_addr + kTmemDw + q * kW,
                               reinterpret_cast<uint8_t*>(
                                   p.dwT + (int64_t(p0 + q) * kPairCols + q4 * 32) * kW),
                               32, p.accumulate_dw != 0);
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

// out[j] = (addend ? addend[j] : 0) + sum_{c < parts} part[c][j]   (fixed order)
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float4* __restrict__ part, int32_t parts, int64_t n4,
                    const float4* __restrict__ addend, float4* __restrict__ out) {
    for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n4;
         j += int64_t(gridDim.x) * blockDim.x) {
        float4 acc = addend != nullptr ? __ldg(addend + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        int c = 0;
        for (; c + 4 <= parts; c += 4) {
            const float4 a = __ldcs(part + (c + 0) * n4 + j), b = __ldcs(part + (c + 1) * n4 + j);
            const float4 d = __ldcs(part + (c + 2) * n4 + j), e = __ldcs(part + (c + 3) * n4 + j);
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
            acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
            acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
        }
        for (; c < parts; ++c) {
            const float4 a = __ldcs(part + c * n4 + j);
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        }
        out[j] = acc;
    }
}

}  // namespace

}  // namespace gkn

extern "C" int gkn_node_fused_bwd_grid(int32_t K, int32_t W, int32_t* grid, int32_t* parts) {
    using namespace gkn;
    if (W != kW || K <= 0 || K % 256 != 0) {
        return set_error("gkn_node_fused_bwd_grid: W must be 64 and K a positive multiple of 256");
    }
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    const int pairs = kW * K / kPairCols;
    if ((pairs + sms - 1) / sms > kMaxPairs) {
        return set_error("gkn_node_fused_bwd_grid: too few SMs for this kernel width "
                         "(more than 4 column pairs per CTA)");
    }
    if (grid) *grid = sms;
    if (parts) *parts = pairs < sms ? pairs : sms;  // CTAs that own at least one column pair
    return 0;
}

extern "C" int gkn_node_fused_bwd_bf16(const void* hs_bf16, const void* gs_bf16, const int32_t* pptr,
                                       const void* x_bf16, const void* w3k_bf16, void* du_bf16,
                                       const int32_t* items, int64_t num_items, int32_t* flags,
                                       float* dx_part, float* dwT, int32_t accumulate_dw, int64_t N,
                                       int64_t E_pad, int64_t du_rows, int32_t K, int32_t W,
                                       void* stream) {
    using namespace gkn;
    if (num_items <= 0) return 0;
    if (hs_bf16 == nullptr || gs_bf16 == nullptr || pptr == nullptr || x_bf16 == nullptr ||
        w3k_bf16 == nullptr || du_bf16 == nullptr || items == nullptr || flags == nullptr) {
        return set_error("gkn_node_fused_bwd_bf16: NULL pointer argument");
    }
    int grid = 0;
    int rc = gkn_node_fused_bwd_grid(K, W, &grid, nullptr);
    if (rc) return rc;
    if (E_pad <= 0 || E_pad % 16 != 0 || E_pad >= (int64_t(1) << 31) - 256 || N <= 0 ||
        du_rows <= 0 || du_rows >= (int64_t(1) << 31)) {
        return set_error("gkn_node_fused_bwd_bf16: bad sizes");
    }
    CUtensorMap th, tg, td, tw, tx;
    rc = make_tensor_map_bf16_2d(&th, hs_bf16, uint64_t(K), uint64_t(E_pad), 64, kKB);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tg, gs_bf16, uint64_t(kW), uint64_t(E_pad), 64, kKB);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&td, du_bf16, uint64_t(kW) * K, uint64_t(du_rows), 64, kM);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tw, w3k_bf16, uint64_t(kW) * K, uint64_t(kW), 64, kKB);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tx, x_bf16, uint64_t(kW), uint64_t(N), 64, kM);
    if (rc) return rc;
    BwdParams p{};
    p.items = reinterpret_cast<const Item*>(items);
    p.num_items = num_items;
    p.flags = flags;
    p.pptr = pptr;
    p.du = static_cast<__nv_bfloat16*>(du_bf16);
    p.dx_part = dx_part;
    p.dwT = dwT;
    p.N = N;
    p.K = K;
    p.accumulate_dw = accumulate_dw;
    {
        static int tune = -1;
        if (tune < 0) {
            const char* e = getenv("GKN_FUSED_TUNE");
            tune = e != nullptr ? atoi(e) : 0;
        }
        p.tune = tune;
    }
    cudaError_t err = cudaFuncSetAttribute(node_fused_bwd_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(node_fused_bwd)", err);
    // the X items are laid out for exactly one CTA per SM (item index = CTA index mod grid)
    node_fused_bwd_kernel<<<dim3(unsigned(grid)), dim3(kThreads), kSmemBytes,
                            static_cast<cudaStream_t>(stream)>>>(th, tg, td, tw, tx, p);
    GKN_CHECK_LAUNCH("node_fused_bwd launch");
    return 0;
}

extern "C" int gkn_sum_partials(const float* part, int32_t parts, int64_t n, const float* addend,
                                float* out, void* stream) {
    using namespace gkn;
    if (n <= 0) return 0;
    if (part == nullptr || out == nullptr || parts < 0) {
        return set_error("gkn_sum_partials: NULL pointer argument");
    }
    if (n % 4 != 0) return set_error("gkn_sum_partials: n must be a multiple of 4");
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > int64_t(sms) * 8) blocks = int64_t(sms) * 8;
    sum_partials_kernel<<<dim3(unsigned(blocks)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(part), parts, n4, reinterpret_cast<const float4*>(addend),
        reinterpret_cast<float4*>(out));
    GKN_CHECK_LAUNCH("sum_partials launch");
    return 0;
}
