// K4 — fused gradient of the edge hidden state on tcgen05 tensor cores (sm_100a).
//
// The per-edge kernel of NNConv_old.message (reference graph_kernel.py:200-202) is shared by all
// A = 2*depth applications (graph_kernel.py:271-273, :299-302), so the loss gradient with respect
// to it is summed over applications before it is contracted with the last edge-MLP layer:
//     S[e, i*w + o] = sum_a xs[a, src e, i] * gs[a, dst e, o]             (never stored)
//     dh[e, k]      = sum_n S[e, n] * W3[n, k]                            (n = i*w + o)
// This kernel is the GEMM  dh = S @ W3  whose A operand S is SYNTHESISED tile by tile by the tensor
// core itself, one 64-wide k-block ahead of its use: S exists only as 128 x 64 tiles in TMEM (fp32)
// and shared memory (bf16).  (Legacy mma.sync was measured for the synthesis first: on this part
// an m16n8k16 occupies the tensor pipe ~45 cycles per SM sub-partition, 11x slower per MAC than
// tcgen05, and it serialises with the main MMA - the synthesis cost 1.4x the GEMM itself.)
//
// Synthesis as an MMA.  Edges are destination-sorted, so the 128 rows of a tile fall into D runs of
// equal destination ("slots", D <= 8).  For k-block i (one input channel, all 64 output channels)
//     S_tile[128 e, 64 o] = A'[128 e, 16 D] @ B'[16 D, 64 o]
//     A'[e, 16 j + a] = xs[a, src e, i] if slot(e) == j else 0        (block diagonal, per k-block)
//     B'[16 j + a, o] = gs[a, dst_j, o]                               (per tile)
// i.e. D tcgen05.mma of M = 128, N = 64, K = 16 (32 tensor cycles each) per 512-cycle main k-block.
//
// Structure (one persistent CTA per SM, 576 threads):
//   warp 0      TMA producer of the main B operand: W3^T tiles [256 k-rows x 64 n].
//   warp 1      TMEM allocator; one thread issues BOTH MMA streams, the synthesis two k-blocks
//               ahead of the main MMA (M = 128, N = 256); tcgen05.commit hands stages back.
//   warps 2..5  epilogue: main accumulator (256 TMEM columns) -> bf16 -> dh.
//   warps 6..9  one thread per tile row: writes its A' row (own slot from a float4 gather of xs
//               along i, zeros elsewhere) in the K-major 128-byte-swizzled layout; builds B' and the
//               slot table when the tile changes.
//   warps 10..  converters, two groups of four warps taking k-blocks in turn: synthesised fp32
//               tile (tcgen05.ld) -> bf16 -> swizzled A operand of the main MMA.
// TMEM: columns [0, 256) main accumulator, [256, 512) four 64-column synthesis stages.
// Work unit = (128-edge tile, 256-column chunk of K); the synthesis is repeated per chunk.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "gkn_b200.h"
#include "gkn_internal.h"
#include "gkn_ptx.cuh"

namespace gkn {

namespace {

constexpr int kM = 128;    // edges per tile
constexpr int kN = 256;    // output columns (of K) per accumulator
constexpr int kKB = 64;    // reduction (n) elements per k-block
constexpr int kConvGroups = 2;  // converter warp groups, taking k-blocks round-robin
constexpr int kThreads = (10 + 4 * kConvGroups) * 32;
constexpr int kBStages = 3, kAStages = 3, kPStages = 2, kSStages = 4;
constexpr int kBBytes = kN * kKB * 2;        // 32 KB  main B tile (W3^T)
constexpr int kABytes = kM * kKB * 2;        // 16 KB  main A tile (S, bf16)
constexpr int kMaxSlots = 8;                 // distinct destinations per 128-edge tile
constexpr int kPBytes = 2 * kM * 128;        // 32 KB  A' (two 64-wide K-major tiles = 8 slots)
constexpr int kGBytes = 2 * 64 * 128;        // 16 KB  B' (64 rows o, 8 slots)
constexpr int kSmemBytes =
    kBStages * kBBytes + kAStages * kABytes + kPStages * kPBytes + kGBytes + 1024 + 512;
constexpr int kPad = 16;  // applications padded to the MMA reduction depth
constexpr int kLook = 2;  // synthesis runs this many k-blocks ahead of the main MMA

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// gt[node][o][a] (bf16, a padded to 16 with zeros) from gs[a][node][o] (fp32)
__global__ void build_gt_kernel(const float* __restrict__ gs, __nv_bfloat16* __restrict__ gt,
                                int64_t N, int A, int W) {
    const int64_t total = N * W;
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += int64_t(gridDim.x) * blockDim.x) {
        uint32_t out[kPad / 2];
#pragma unroll
        for (int a = 0; a < kPad; a += 2) {
            const float v0 = a < A ? gs[int64_t(a) * total + t] : 0.f;
            const float v1 = a + 1 < A ? gs[int64_t(a + 1) * total + t] : 0.f;
            out[a / 2] = pack_bf16(v0, v1);
        }
        uint4* dst = reinterpret_cast<uint4*>(gt + t * kPad);
        dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
        dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
    }
}

// status[0] = max over 128-edge tiles of the number of distinct destinations in the tile
__global__ void max_slots_kernel(const int32_t* __restrict__ dst, int64_t E,
                                 int32_t* __restrict__ status) {
    const int64_t tiles = (E + kM - 1) / kM;
    int worst = 0;
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < tiles;
         t += int64_t(gridDim.x) * blockDim.x) {
        const int64_t e0 = t * kM, e1 = e0 + kM < E ? e0 + kM : E;
        int runs = 1;
        for (int64_t e = e0 + 1; e < e1; ++e) runs += dst[e] != dst[e - 1];
        worst = max(worst, runs);
    }
    atomicMax(status, worst);
}

struct DhParams {
    const float* __restrict__ xs;          // [A, N, W]
    const __nv_bfloat16* __restrict__ gt;  // [N, W, 16]
    const int32_t* __restrict__ src;       // [E] sorted order
    const int32_t* __restrict__ dst;       // [E] sorted order (non-decreasing)
    __nv_bfloat16* __restrict__ dh;        // [E, K]
    int64_t E, N;
    int32_t A, K;
    int64_t num_tiles;
};

template <int W>
__global__ void __launch_bounds__(kThreads, 1)
edge_grad_dh_kernel(const __grid_constant__ CUtensorMap tmap_w3t, const DhParams p) {
    static_assert(W == 64, "k-block = one i value: width 64");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base + kBStages * kBBytes;   // S tiles (bf16)
    const uint32_t p_base = a_base + kAStages * kABytes;      // A' buffers
    const uint32_t g_base = p_base + kPStages * kPBytes;      // B'
    const uint32_t bar_base = g_base + kGBytes;
    // barriers: bfull, bempty [kBStages]; afull, aempty [kAStages]; pfull, pempty [kPStages];
    //           sfull, sempty [kSStages]; tfull, tempty
    constexpr int oB = 0, oA = oB + 2 * kBStages, oP = oA + 2 * kAStages, oS = oP + 2 * kPStages,
                  oT = oS + 2 * kSStages, oEnd = oT + 2;
    auto bfull = [&](int s) { return bar_base + 8u * (oB + s); };
    auto bempty = [&](int s) { return bar_base + 8u * (oB + kBStages + s); };
    auto afull = [&](int s) { return bar_base + 8u * (oA + s); };
    auto aempty = [&](int s) { return bar_base + 8u * (oA + kAStages + s); };
    auto pfull = [&](int s) { return bar_base + 8u * (oP + s); };
    auto pempty = [&](int s) { return bar_base + 8u * (oP + kPStages + s); };
    auto sfull = [&](int s) { return bar_base + 8u * (oS + s); };
    auto sempty = [&](int s) { return bar_base + 8u * (oS + kSStages + s); };
    const uint32_t tfull = bar_base + 8u * oT, tempty = bar_base + 8u * (oT + 1);
    const uint32_t tmem_ptr_smem = bar_base + 8u * oEnd;
    const uint32_t slots_smem = tmem_ptr_smem + 16;  // int D; int dst_of_slot[8]
    // scratch for the row-slot prefix of the synthesis warps: 4 ints after the slot table
    const uint32_t wsum_smem = slots_smem + 4 * (1 + kMaxSlots);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int chunks = p.K / kN;
    constexpr int NK = (W * W) / kKB;  // 64 k-blocks per unit
    const int64_t total_units = p.num_tiles * chunks;
    const int64_t u0 = (total_units * blockIdx.x) / gridDim.x;
    const int64_t u1 = (total_units * (blockIdx.x + 1)) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kBStages; ++s) {
            mbar_init(bfull(s), 1);
            mbar_init(bempty(s), 1);
        }
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(afull(s), 4);
            mbar_init(aempty(s), 1);
        }
        for (int s = 0; s < kPStages; ++s) {
            mbar_init(pfull(s), 4);
            mbar_init(pempty(s), 1);
        }
        for (int s = 0; s < kSStages; ++s) {
            mbar_init(sfull(s), 1);
            mbar_init(sempty(s), 4);
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, 4);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_w3t);
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

    if (warp == 0) {
        // ===================== TMA producer: W3^T tiles =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = u0; u < u1; ++u) {
                const int32_t chunk = int32_t(u % chunks);
                for (int kb = 0; kb < NK; ++kb) {
                    mbar_wait(bempty(stage), phase ^ 1u);
                    mbar_arrive_expect_tx(bfull(stage), kBBytes);
                    tma_load_2d(smem_base + stage * kBBytes, &tmap_w3t, bfull(stage), kb * kKB,
                                chunk * kN);
                    if (++stage == kBStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: synthesis stream + main stream =====================
        if (lane == 0) {
            constexpr uint32_t idesc_main = umma_idesc_bf16(kM, kN, 0, 0);
            constexpr uint32_t idesc_syn = umma_idesc_bf16(kM, 64, 0, 0);
            int bs = 0, as = 0;
            uint32_t bphase = 0, aphase = 0;
            uint32_t pq = 0;  // synthesis sequence number (A' buffer pq % 2, S stage pq % 4)
            uint32_t it = 0;
            auto issue_synth = [&]() {
                const uint32_t ps = pq % kPStages, ss = pq % kSStages;
                mbar_wait(pfull(ps), (pq / kPStages) & 1u);
                mbar_wait(sempty(ss), ((pq / kSStages) & 1u) ^ 1u);
                tc_fence_after();
                int D;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(D) : "r"(slots_smem));
                const uint32_t tmem_s = tmem_base + kN + ss * 64;
                for (int j = 0; j < D; ++j) {
                    // slot j: tile j / 4 of the buffer, 32 bytes (16 bf16) at K offset 16 (j % 4)
                    const uint64_t adesc = umma_smem_desc(
                        p_base + ps * kPBytes + (j >> 2) * (kM * 128), 16, 1024, kSwizzle128B);
                    const uint64_t bdesc =
                        umma_smem_desc(g_base + (j >> 2) * (64 * 128), 16, 1024, kSwizzle128B);
                    umma_bf16_ss(tmem_s, adesc + uint64_t(2 * (j & 3)), bdesc + uint64_t(2 * (j & 3)),
                                 idesc_syn, j != 0 ? 1u : 0u);
                }
                umma_commit(pempty(ps));
                umma_commit(sfull(ss));
                ++pq;
            };
            for (int64_t u = u0; u < u1; ++u, ++it) {
                for (int kb = 0; kb < kLook; ++kb) issue_synth();
                mbar_wait(tempty, (it & 1u) ^ 1u);  // epilogue has drained the accumulator
                tc_fence_after();
                for (int kb = 0; kb < NK; ++kb) {
                    if (kb + kLook < NK) issue_synth();
                    mbar_wait(bfull(bs), bphase);
                    mbar_wait(afull(as), aphase);
                    tc_fence_after();
                    const uint64_t adesc = umma_smem_desc(a_base + as * kABytes, 16, 1024, kSwizzle128B);
                    const uint64_t bdesc =
                        umma_smem_desc(smem_base + bs * kBBytes, 16, 1024, kSwizzle128B);
#pragma unroll
                    for (int k = 0; k < kKB / 16; ++k) {
                        umma_bf16_ss(tmem_base, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k),
                                     idesc_main, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(bempty(bs));
                    umma_commit(aempty(as));
                    if (++bs == kBStages) {
                        bs = 0;
                        bphase ^= 1u;
                    }
                    if (++as == kAStages) {
                        as = 0;
                        aphase ^= 1u;
                    }
                }
                umma_commit(tfull);
            }
        }
    } else if (warp < 6) {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const uint32_t lane_addr = uint32_t(q * 32) << 16;
        uint32_t it = 0;
        for (int64_t u = u0; u < u1; ++u, ++it) {
            const int64_t tile = u / chunks;
            const int32_t chunk = int32_t(u % chunks);
            const int64_t e = tile * kM + row;
            if (lane == 0) mbar_wait_backoff(tfull, it & 1u, 128);
            __syncwarp();
            tc_fence_after();
            const uint32_t tcol = tmem_base + lane_addr;
            __nv_bfloat16* out = p.dh + e * p.K + chunk * kN;
#pragma unroll
            for (int piece = 0; piece < kN / 32; ++piece) {
                float v[32];
                tmem_ld32(tcol + piece * 32, v);
                tmem_ld_wait();
                if (e < p.E) {
                    uint4* o4 = reinterpret_cast<uint4*>(out + piece * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        o4[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]),
                                           pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                           pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                                           pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
        }
    } else if (warp < 10) {
        // ===================== A' writers (warps 6..9): thread = tile row =====================
        const int q = warp & 3;  // row quarter (warps 6..9 -> quarters 2, 3, 0, 1)
        const int row = q * 32 + lane;
        const int tid = threadIdx.x - 6 * 32;  // 0..127 (cooperative work, not the row id)
        const uint32_t swz = uint32_t(row & 7);
        const uint32_t row_off = uint32_t(row >> 3) * 1024u + swz * 128u;
        uint32_t pq = 0;  // A' sequence number (counts k-blocks across units)
        int64_t cur_tile = -1;
        int myslot = 0, D = 1;
        int32_t s_node = 0;
        bool valid = false;
        uint32_t xq[4][8];  // bf16 pairs (a, a+1) of 4 consecutive k-blocks
        int xgrp = -1;

        auto load_x = [&](int grp) {
            float4 v[kPad];
#pragma unroll
            for (int a = 0; a < kPad; ++a) {
                v[a] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid && a < p.A) {
                    v[a] = __ldg(reinterpret_cast<const float4*>(
                        p.xs + (int64_t(a) * p.N + s_node) * W + 4 * grp));
                }
            }
#pragma unroll
            for (int a = 0; a < kPad; a += 2) {
                xq[0][a / 2] = pack_bf16(v[a].x, v[a + 1].x);
                xq[1][a / 2] = pack_bf16(v[a].y, v[a + 1].y);
                xq[2][a / 2] = pack_bf16(v[a].z, v[a + 1].z);
                xq[3][a / 2] = pack_bf16(v[a].w, v[a + 1].w);
            }
            xgrp = grp;
        };

        for (int64_t u = u0; u < u1; ++u) {
            const int64_t tile = u / chunks;
            if (tile != cur_tile) {
                cur_tile = tile;
                const int64_t e = tile * kM + row;
                valid = e < p.E;
                const int64_t ec = valid ? e : p.E - 1;
                s_node = p.src[ec];
                const int32_t d = p.dst[ec];
                const bool flag = valid && (row == 0 || d != p.dst[ec - 1]);
                const unsigned fm = __ballot_sync(0xffffffffu, flag);
                // B' and the slot table are rewritten only after every synthesis MMA of the previous
                // tile has completed: both A' buffers have been handed back (pempty)
                if (pq >= kPStages) {
#pragma unroll
                    for (int s2 = 0; s2 < kPStages; ++s2) {
                        const uint32_t last = pq - 1 - ((pq - 1 - s2) % kPStages);  // last use of buffer s2
                        if (lane == 0) mbar_wait_backoff(pempty(s2), (last / kPStages) & 1u, 32);
                    }
                }
                __syncwarp();
                named_bar_sync(1, 128);
                if (lane == 0) {
                    asm volatile("st.shared.s32 [%0], %1;" ::"r"(wsum_smem + 4u * q), "r"(__popc(fm))
                                 : "memory");
                }
                named_bar_sync(1, 128);
                int before = 0, total = 0;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    int c;
                    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(c) : "r"(wsum_smem + 4u * w));
                    if (w < q) before += c;
                    total += c;
                }
                D = total < 1 ? 1 : (total > kMaxSlots ? kMaxSlots : total);  // host checks <= 8
                myslot = before + __popc(fm & ((2u << lane) - 1u)) - 1;
                if (myslot < 0) myslot = 0;
                if (flag && myslot < kMaxSlots) {
                    asm volatile("st.shared.s32 [%0], %1;" ::"r"(slots_smem + 4u * (1 + myslot)), "r"(d)
                                 : "memory");
                }
                if (tid == 0) {
                    asm volatile("st.shared.s32 [%0], %1;" ::"r"(slots_smem), "r"(D) : "memory");
                }
                named_bar_sync(1, 128);
                // B'[o][16 j + a] = gt[dst_j][o][a]: 64 rows x D slots, 32 bytes each
                for (int t = tid; t < 64 * D; t += 128) {
                    const int j = t >> 6, o = t & 63;
                    int dj;
                    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(dj) : "r"(slots_smem + 4u * (1 + j)));
                    const uint4* gp =
                        reinterpret_cast<const uint4*>(p.gt + (int64_t(dj) * W + o) * kPad);
                    const uint4 v0 = __ldg(gp), v1 = __ldg(gp + 1);
                    const uint32_t rbase = g_base + uint32_t(j >> 2) * (64 * 128) +
                                           uint32_t(o >> 3) * 1024u + uint32_t(o & 7) * 128u;
                    const uint32_t c0 = uint32_t(2 * (j & 3));
                    sts128(rbase + ((c0 ^ uint32_t(o & 7)) << 4), v0);
                    sts128(rbase + (((c0 + 1) ^ uint32_t(o & 7)) << 4), v1);
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                xgrp = -1;
            }
            for (int kb = 0; kb < NK; ++kb) {
                // A' row of k-block kb: own slot = 16 bf16 (applications, zero padded), zeros elsewhere
                const uint32_t ps = pq % kPStages;
                if ((kb >> 2) != xgrp) load_x(kb >> 2);
                const int kk = kb & 3;
                uint4 lo = make_uint4(xq[0][0], xq[0][1], xq[0][2], xq[0][3]);
                uint4 hi = make_uint4(xq[0][4], xq[0][5], xq[0][6], xq[0][7]);
#pragma unroll
                for (int c = 1; c < 4; ++c) {  // static register indexing
                    if (kk == c) {
                        lo = make_uint4(xq[c][0], xq[c][1], xq[c][2], xq[c][3]);
                        hi = make_uint4(xq[c][4], xq[c][5], xq[c][6], xq[c][7]);
                    }
                }
                if (lane == 0) mbar_wait_backoff(pempty(ps), ((pq / kPStages) & 1u) ^ 1u, 32);
                __syncwarp();
                const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
                const uint32_t base = p_base + ps * kPBytes + row_off;
                for (int j = 0; j < D; ++j) {
                    const uint32_t tile_off = uint32_t(j >> 2) * (kM * 128);
                    const uint32_t c0 = uint32_t(2 * (j & 3));
                    const bool mine = j == myslot;
                    sts128(base + tile_off + ((c0 ^ swz) << 4), mine ? lo : zero);
                    sts128(base + tile_off + (((c0 + 1) ^ swz) << 4), mine ? hi : zero);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(pfull(ps));
                ++pq;
            }
        }
    } else {
        // ===================== converters (warps 10..): synthesised tile -> bf16 A operand ======
        // group g takes the k-blocks with (global k-block number) % kConvGroups == g
        const int grp = (warp - 10) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = uint32_t(q * 32) << 16;
        const uint32_t swz = uint32_t(row & 7);
        const uint32_t row_off = uint32_t(row >> 3) * 1024u + swz * 128u;
        const uint64_t nkb = uint64_t(u1 - u0) * NK;
        for (uint64_t kq = grp; kq < nkb; kq += kConvGroups) {
            const uint32_t ss = uint32_t(kq % kSStages), as = uint32_t(kq % kAStages);
            if (lane == 0) mbar_wait_backoff(sfull(ss), uint32_t(kq / kSStages) & 1u, 32);
            __syncwarp();
            tc_fence_after();
            uint32_t packed[32];
            {
                float v0[32], v1[32];
                const uint32_t tcol = tmem_base + lane_addr + kN + ss * 64;
                tmem_ld32(tcol, v0);
                tmem_ld32(tcol + 32, v1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    packed[j] = pack_bf16(v0[2 * j], v0[2 * j + 1]);
                    packed[16 + j] = pack_bf16(v1[2 * j], v1[2 * j + 1]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(sempty(ss));
                mbar_wait_backoff(aempty(as), (uint32_t(kq / kAStages) & 1u) ^ 1u, 32);
            }
            __syncwarp();
            const uint32_t abase = a_base + as * kABytes + row_off;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                sts128(abase + ((uint32_t(c) ^ swz) << 4),
                       make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2],
                                  packed[4 * c + 3]));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(afull(as));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

}  // namespace gkn

extern "C" int64_t gkn_edge_grad_workspace_bytes(int64_t N, int32_t W) {
    if (N < 0 || W <= 0) return 0;
    return N * int64_t(W) * gkn::kPad * 2;
}

extern "C" int gkn_edge_grad_max_slots(const int32_t* dst_s, int64_t E, int32_t* status,
                                       void* stream) {
    using namespace gkn;
    if (E < 0 || status == nullptr || (E > 0 && dst_s == nullptr)) {
        return set_error("gkn_edge_grad_max_slots: bad argument");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t err = cudaMemsetAsync(status, 0, 4, st);
    if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(status)", err);
    if (E == 0) return 0;
    const int64_t tiles = (E + kM - 1) / kM;
    int64_t blocks = (tiles + 127) / 128;
    if (blocks > 1184) blocks = 1184;
    max_slots_kernel<<<int(blocks), 128, 0, st>>>(dst_s, E, status);
    GKN_CHECK_LAUNCH("max_slots launch");
    return 0;
}

extern "C" int gkn_edge_grad_dh_bf16(const float* xs, const float* gs, const int32_t* src_s,
                                     const int32_t* dst_s, const void* w3t_bf16, void* dh_bf16,
                                     int64_t E, int64_t N, int32_t A, int32_t K, int32_t W,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace gkn;
    if (E == 0) return 0;
    if (E < 0 || N <= 0) return set_error("gkn_edge_grad_dh_bf16: bad size");
    if (xs == nullptr || gs == nullptr || src_s == nullptr || dst_s == nullptr ||
        w3t_bf16 == nullptr || dh_bf16 == nullptr || workspace == nullptr) {
        return set_error("gkn_edge_grad_dh_bf16: NULL pointer argument");
    }
    if (W != 64) return set_error("gkn_edge_grad_dh_bf16: width must be 64");
    if (A < 1 || A > kPad) return set_error("gkn_edge_grad_dh_bf16: 1 <= applications <= 16");
    if (K <= 0 || K % kN != 0) {
        return set_error("gkn_edge_grad_dh_bf16: kernel_width K must be a positive multiple of 256");
    }
    if (workspace_bytes < gkn_edge_grad_workspace_bytes(N, W)) {
        return set_error("gkn_edge_grad_dh_bf16: workspace too small");
    }
    if (E >= (int64_t(1) << 31) - 2 * kM) {
        return set_error("gkn_edge_grad_dh_bf16: too many edges for one launch (>= 2^31)");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __nv_bfloat16* gt = static_cast<__nv_bfloat16*>(workspace);
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    {
        int64_t blocks = (N * W + 255) / 256;
        if (blocks > int64_t(sms) * 8) blocks = int64_t(sms) * 8;
        build_gt_kernel<<<int(blocks), 256, 0, st>>>(gs, gt, N, A, W);
        GKN_CHECK_LAUNCH("build_gt launch");
    }
    CUtensorMap tmap;
    rc = make_tensor_map_bf16_2d(&tmap, w3t_bf16, uint64_t(W) * uint64_t(W), uint64_t(K), kKB, kN);
    if (rc) return rc;
    DhParams p;
    p.xs = xs;
    p.gt = gt;
    p.src = src_s;
    p.dst = dst_s;
    p.dh = static_cast<__nv_bfloat16*>(dh_bf16);
    p.E = E;
    p.N = N;
    p.A = A;
    p.K = K;
    p.num_tiles = (E + kM - 1) / kM;
    auto kern = edge_grad_dh_kernel<64>;
    cudaError_t err =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(edge_grad_dh)", err);
    int64_t grid = p.num_tiles * (K / kN);
    if (grid > sms) grid = sms;
    kern<<<int(grid), kThreads, kSmemBytes, st>>>(tmap, p);
    GKN_CHECK_LAUNCH("edge_grad_dh launch");
    return 0;
}
