// Node-first factorisation of the graph-kernel contraction (tcgen05 grouped GEMMs, sm_100a).
//
// NNConv_old.message (reference graph_kernel.py:200-202) is the trilinear form
//     msg[e, o] = sum_{i,k} x[src e, i] * h[e, k] * W3[i*w + o, k]            (+ the b3 term)
// The edge-first kernels (gkn_contract_tc.cu) contract h with W3 per edge: 2*E*K*w^2 FLOP.  But
// x[src e] is a NODE quantity, so W3 can be contracted with x once per node instead,
//     U[n][o, k] = sum_i x[n, i] * W3[i*w + o, k]        (plain dense GEMM [N,w] x [w, w*K])
//     msg[e, o]  = sum_k h[e, k] * U[src e][o, k]         (GEMM grouped by source node)
// which is N*2*K*w^2 + E*2*K*w FLOP: w/(1 + w*N/E) times fewer (about 40x for the BBA all-atom
// graph) and no longer tensor-bound - every kernel here streams the edge hidden state once and
// is bounded by HBM.  The backward follows from the same two stages:
//     dU[n][k, o] = sum_{e: src e = n} h[e, k] * gm[e, o]            (gm[e] = gn[dst e])
//     dx = dU W3^T,  dW3 += x^T dU                                   (plain dense GEMMs)
//     dh[e, k]    = sum_a sum_o gm_a[e, o] * U_a[src e][o, k]        (all applications at once)
//
// Edge layout: "padded source-sorted": the edges of source node n occupy rows
// [pptr[n], pptr[n] + deg[n]) and every group is padded to a multiple of 16 rows (pad rows carry
// zero gradients), so a group is a whole number of UMMA K-steps when the edge index is the
// reduction dimension (MODE_DU) and 128-row M tiles start at a group boundary (MODE_MSG/DH; rows
// of the tile that belong to the next group are computed and discarded).
//
// One persistent kernel template, three modes; all operands K-major bf16 staged by TMA with the
// 128-byte swizzle, fp32 accumulators in TMEM, warp roles as in gkn_contract_tc.cu:
//   MODE_MSG  unit = 128-row tile t of group n:  D[128 x w]   = Hs[tile, :] . U[n]^T      (K = kernel width)
//   MODE_DU   unit = (group n, 128 k-rows c):    D[128 x w]   = HsT[c, grp] . GsT[:, grp]^T (K = padded degree)
//   MODE_DH   unit = (tile, 256-column chunk c): D[128 x 256] = sum_a Gs_a[tile, :] . Ut_a[n][c, :]^T (K = A*w)
#include <stdio.h>
#include <stdlib.h>

#include "gkn_internal.h"
#include "gkn_ptx.cuh"

#include <cuda_bf16.h>

namespace gkn {

namespace {

constexpr int kM = 128;      // accumulator rows per unit
constexpr int kKB = 64;      // reduction elements per k-block (one 128-byte swizzle row)
constexpr int kUmmaK = 16;
// epilogue warp groups (4 warps each) taking units alternately.  Measured with 2: node_du 0.582 ->
// 0.569 ms, node_msg and node_dh 1-2 % slower, step time unchanged - the epilogue is not what
// paces these kernels, so one group (and 16 KB less shared memory) it is.
constexpr int kEpiGroups = 1;
constexpr int kThreads = (2 + 4 * kEpiGroups) * 32;
constexpr int kABytes = kM * kKB * 2;  // 16 KB
constexpr int kW = 64;       // node feature width of the K-major backward kernels (MODE_DU / MODE_DH)

// *_MN: the same products with operands read straight from their natural row-major buffers as
// MN-major UMMA operands (no transposed copies of h / gm, no [k,o]-ordered copy of U):
//   MODE_DU_MN  A = hs[grp, c] (M = k contiguous), B = gs[grp, :] (N = o contiguous), K = edges
//   MODE_DH_MN  A = Gs_a[tile, :] (K-major),        B = U_a[n][:, c] (N = k contiguous), K = o
enum { MODE_MSG = 0, MODE_DU = 1, MODE_DH = 2, MODE_DU_MN = 3, MODE_DH_MN = 4 };

__host__ __device__ constexpr bool mode_is_dh(int m) { return m == MODE_DH || m == MODE_DH_MN; }
__host__ __device__ constexpr bool mode_is_du(int m) { return m == MODE_DU || m == MODE_DU_MN; }
constexpr int kMnBlockBytes = 64 * kKB * 2;  // one MN-major TMA box: 64 K-rows x 64 MN elements
constexpr int kMaskMaxK = 1024;              // widest edge MLP the fused-mask dh epilogue serves

// Row length of the padded-order gradient buffer gs: 128-byte rows at least, so that one row is a
// whole swizzle row of an MN-major box.  At w = 32 the upper 32 columns are zeros (written by
// gkn_node_gather_bf16); the dU product then runs 64 accumulator columns wide and stores 32.
__host__ __device__ constexpr int gs_width(int W) { return W < 64 ? 64 : W; }

template <int MODE, int W>
struct NCfg {
    // accumulator columns
    static constexpr int kBN = mode_is_dh(MODE) ? 256 : (MODE == MODE_DU_MN ? gs_width(W) : W);
    static constexpr int kBBytes = kBN * kKB * 2;               // 4 .. 32 KB
    static constexpr int kStageBytes = kABytes + kBBytes;       // 20 .. 48 KB
    static constexpr int kStages = 192 * 1024 / kStageBytes > 8 ? 8 : 192 * 1024 / kStageBytes;
    static constexpr int kAcc = mode_is_dh(MODE) ? 2 : 4;       // TMEM accumulator ring
    static constexpr bool kAMn = MODE == MODE_DU_MN;            // A operand MN-major
    static constexpr bool kBMn = MODE == MODE_DU_MN || MODE == MODE_DH_MN;
    static constexpr int kABoxes = kAMn ? kM / 64 : 1;          // TMA boxes per k-block
    static constexpr int kBBoxes = kBMn ? kBN / 64 : 1;
    static constexpr int kTmemCols = kAcc * kBN;                // 512 / 256
    // MODE_DH_MN: reduction = the w output channels of every application, in k-blocks of 64 (two
    // per application at w = 128; at w = 32 one k-block of two UMMA steps whose B boxes are 32 rows)
    static constexpr int kKbPerApp = MODE == MODE_DH_MN && W > 64 ? W / 64 : 1;
    static constexpr int kFullSteps = MODE == MODE_DH_MN && W < 64 ? W / kUmmaK : kKB / kUmmaK;
    static constexpr int kBBoxRows = MODE == MODE_DH_MN && W < 64 ? W : 64;
    static constexpr int kTxBytes =
        kBMn ? kABytes + kBBoxes * kBBoxRows * 128 : kStageBytes;  // bytes TMA delivers per stage
    // MODE_DH_MN with the ReLU mask of DenseNet layers.2 fused (NodeParams::act): per-warp column
    // sums of the masked output (the bias gradient), kMaskMaxK columns x 4 epilogue warps
    static constexpr int kColsumBytes = MODE == MODE_DH_MN ? 4 * kMaskMaxK * 4 : 0;
    static constexpr int kSmemBytes =
        kStages * kStageBytes + 1024 + 256 + kEpiGroups * 4 * 4096 /*output staging*/ + kColsumBytes;
    static_assert(kAcc % kEpiGroups == 0, "accumulator stages are owned by epilogue groups");
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory per CTA");
};

struct NodeParams {
    const int4* __restrict__ units;     // MSG: (group, first padded row, rows to store, -), one per
                                        // tile; DH: the same per (tile, column chunk) with the chunk
                                        // in .w, ordered (group, chunk, tile)
    const int32_t* __restrict__ pptr;   // DU: padded group pointers [N+1]
    void* __restrict__ out;
    int64_t v_begin, v_end;             // unit range of this launch
    int32_t K;                          // kernel width (columns of h)
    int32_t A;                          // DH: applications
    int32_t a_rows_per_app;             // DH: rows of Gs per application (= padded edge count)
    int32_t b_rows_per_app;             // DH: rows of Ut per application (= chunk nodes * K)
    int32_t node_base;                  // first node of the U / Ut / dUt chunk
    int32_t hints;                      // bit 0: A operand loaded with L2 evict_first; bit 1:
                                        // streaming (.cs) output stores  (GKN_NODE_HINTS)
    const uint32_t* __restrict__ act_bits;  // DH_MN: [E_pad, K/32] bit c%32 of word c/32 = (hs[q,c] > 0);
                                        // the output is zeroed where the bit is clear (or NULL)
    float* __restrict__ colsum_part;    // DH_MN with act: per-CTA column sums [gridDim.x, K]
};

struct Unit {
    int32_t a_c0, a_c1, b_c0, b_c1;     // TMA coordinates of k-block 0
    int32_t da_c0, da_c1, db_c0, db_c1; // per-k-block increments
    int32_t nkb;                        // k-blocks
    int32_t last_steps;                 // UMMA K-steps in the last k-block (1..4)
    int32_t row0, rows, group, chunk;   // epilogue
};

template <int MODE, int W>
__device__ __forceinline__ Unit decode_unit(const NodeParams& p, int64_t v) {
    Unit u;
    if constexpr (MODE == MODE_MSG) {
        const int4 t = __ldg(p.units + v);
        u.group = t.x; u.row0 = t.y; u.rows = t.z; u.chunk = 0;
        u.a_c0 = 0; u.a_c1 = t.y; u.da_c0 = kKB; u.da_c1 = 0;
        u.b_c0 = 0; u.b_c1 = (t.x - p.node_base) * W; u.db_c0 = kKB; u.db_c1 = 0;
        u.nkb = p.K / kKB; u.last_steps = kKB / kUmmaK;
    } else if constexpr (MODE == MODE_DH_MN) {
        const int4 t = __ldg(p.units + v);  // (group, first row, rows, column chunk)
        u.group = t.x; u.row0 = t.y; u.rows = t.z; u.chunk = t.w;
        u.a_c0 = 0; u.a_c1 = t.y; u.da_c0 = 0; u.da_c1 = p.a_rows_per_app;
        u.b_c0 = u.chunk * 256; u.b_c1 = (t.x - p.node_base) * W;
        u.db_c0 = 0; u.db_c1 = p.b_rows_per_app;
        u.nkb = p.A * NCfg<MODE, W>::kKbPerApp; u.last_steps = NCfg<MODE, W>::kFullSteps;
    } else if constexpr (MODE == MODE_DU_MN) {
        const int32_t chunks = p.K / kM;
        const int32_t g = p.node_base + int32_t(v / chunks);
        const int32_t b = __ldg(p.pptr + g), e = __ldg(p.pptr + g + 1);
        const int32_t len = e - b;  // multiple of 16
        u.group = g; u.chunk = int32_t(v % chunks); u.row0 = b; u.rows = kM;
        u.a_c0 = u.chunk * kM; u.a_c1 = b; u.da_c0 = 0; u.da_c1 = kKB;
        u.b_c0 = 0; u.b_c1 = b; u.db_c0 = 0; u.db_c1 = kKB;
        u.nkb = (len + kKB - 1) / kKB;
        u.last_steps = (len - (u.nkb - 1) * kKB) / kUmmaK;
    } else if constexpr (MODE == MODE_DH) {
        const int4 t = __ldg(p.units + v);  // (group, first row, rows, column chunk)
        u.group = t.x; u.row0 = t.y; u.rows = t.z; u.chunk = t.w;
        u.a_c0 = 0; u.a_c1 = t.y; u.da_c0 = 0; u.da_c1 = p.a_rows_per_app;
        u.b_c0 = 0; u.b_c1 = (t.x - p.node_base) * p.K + u.chunk * 256;
        u.db_c0 = 0; u.db_c1 = p.b_rows_per_app;
        u.nkb = p.A; u.last_steps = kKB / kUmmaK;
    } else {
        const int32_t chunks = p.K / kM;
        const int32_t g = p.node_base + int32_t(v / chunks);
        const int32_t b = __ldg(p.pptr + g), e = __ldg(p.pptr + g + 1);
        const int32_t len = e - b;  // multiple of 16
        u.group = g; u.chunk = int32_t(v % chunks); u.row0 = b; u.rows = kM;
        u.a_c0 = b; u.a_c1 = u.chunk * kM; u.da_c0 = kKB; u.da_c1 = 0;
        u.b_c0 = b; u.b_c1 = 0; u.db_c0 = kKB; u.db_c1 = 0;
        u.nkb = (len + kKB - 1) / kKB;
        u.last_steps = (len - (u.nkb - 1) * kKB) / kUmmaK;
    }
    return u;
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

template <int MODE, int W>
__global__ void __launch_bounds__(kThreads, 1)
node_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, const NodeParams p) {
    static_assert(MODE == MODE_MSG || MODE == MODE_DU_MN || MODE == MODE_DH_MN || W == kW,
                  "the K-major backward kernels serve w = 64");
    using C = NCfg<MODE, W>;
    constexpr int kStages = C::kStages;
    constexpr int kStageBytes = C::kStageBytes;
    constexpr int kBN = C::kBN;
    constexpr int kAcc = C::kAcc;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages * kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + kAcc + s); };
    const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 2 * kAcc);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // Units are dealt round-robin: consecutive units (the tiles of one group, the column chunks of
    // one tile) share operands, and dealt this way they run at the same time on neighbouring CTAs,
    // so the second reader finds the operand in L2.  (With a contiguous range per CTA the re-read
    // comes one whole unit later, after ~148 x 0.5 MB of other traffic: measured 11.2 GB of DRAM
    // reads for 7.1 GB of operands in the dh pass.)
    const int64_t v0 = p.v_begin + blockIdx.x;
    const int64_t v1 = p.v_end;
    const int64_t vstep = gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < kAcc; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, C::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint64_t pol_stream = l2_policy_evict_first();
            // the next unit's table entry / row pointers are fetched one unit ahead so that their
            // load latency never sits between two units (MODE_DU* units are only ~2 k-blocks long)
            Unit un = v0 < v1 ? decode_unit<MODE, W>(p, v0) : Unit{};
            for (int64_t v = v0; v < v1; v += vstep) {
                const Unit u = un;
                if (v + vstep < v1) un = decode_unit<MODE, W>(p, v + vstep);
                for (int kb = 0; kb < u.nkb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    mbar_arrive_expect_tx(full_bar(stage), C::kTxBytes);
                    // k-block -> (application / plain index, 64-channel half inside it)
                    const int ka = kb / C::kKbPerApp, half = (kb % C::kKbPerApp) * kKB;
                    const int a_c0 = u.a_c0 + ka * u.da_c0 + half, a_c1 = u.a_c1 + ka * u.da_c1;
                    const int b_c0 = u.b_c0 + ka * u.db_c0, b_c1 = u.b_c1 + ka * u.db_c1 + half;
#pragma unroll
                    for (int j = 0; j < C::kABoxes; ++j) {
                        if (p.hints & 1) {
                            tma_load_2d_hint(sa + j * kMnBlockBytes, &tmap_a, full_bar(stage),
                                             a_c0 + j * 64, a_c1, pol_stream);
                        } else {
                            tma_load_2d(sa + j * kMnBlockBytes, &tmap_a, full_bar(stage),
                                        a_c0 + j * 64, a_c1);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < C::kBBoxes; ++j) {
                        tma_load_2d(sa + kABytes + j * kMnBlockBytes, &tmap_b, full_bar(stage),
                                    b_c0 + j * 64, b_c1);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kM, kBN, C::kAMn ? 1 : 0, C::kBMn ? 1 : 0);
            // K-major: 16 K elements = 32 B inside the swizzle row.  MN-major: a K element is a
            // 128-byte smem row, 64-element MN blocks are kMnBlockBytes apart (LBO), 8-row groups
            // 1024 B apart (SBO); a K-step of 16 rows advances the start address by 2048 B.
            constexpr uint32_t a_lbo = C::kAMn ? kMnBlockBytes : 16, b_lbo = C::kBMn ? kMnBlockBytes : 16;
            constexpr uint64_t a_step = C::kAMn ? (16 * 128) >> 4 : 2, b_step = C::kBMn ? (16 * 128) >> 4 : 2;
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            Unit un = v0 < v1 ? decode_unit<MODE, W>(p, v0) : Unit{};
            for (int64_t v = v0; v < v1; v += vstep) {
                const Unit u = un;
                if (v + vstep < v1) un = decode_unit<MODE, W>(p, v + vstep);
                if (u.nkb == 0) continue;  // empty group: the epilogue writes zeros on its own
                const uint32_t as = it % kAcc;
                const uint32_t aphase = (it / kAcc) & 1u;
                ++it;
                mbar_wait(tempty_bar(as), aphase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * kBN;
                for (int kb = 0; kb < u.nkb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    const uint64_t adesc = umma_smem_desc(sa, a_lbo, 1024, kSwizzle128B);
                    const uint64_t bdesc = umma_smem_desc(sa + kABytes, b_lbo, 1024, kSwizzle128B);
                    const int steps = kb == u.nkb - 1 ? u.last_steps : C::kFullSteps;
                    for (int k = 0; k < steps; ++k) {
                        umma_bf16_ss(tmem_d, adesc + a_step * uint64_t(k), bdesc + b_step * uint64_t(k),
                                     idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull_bar(as));
            }
        }
    } else {
        // ===================== epilogue (warps 2..): thread = accumulator row =====================
        // kEpiGroups groups of four warps take the units alternately (accumulator stage parity).
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;      // epilogue group
        const int row = q * 32 + lane;
        const uint32_t lane_addr = uint32_t(q * 32) << 16;
        const uint32_t stage_smem = bar_base + 256 + (grp * 4 + q) * 4096;  // per-warp staging tile
        // fused ReLU mask (MODE_DH_MN): this warp's column sums of what it stores
        bool masked = false;
        float* csum_smem = nullptr;
        if constexpr (MODE == MODE_DH_MN) {
            masked = p.act_bits != nullptr;
            csum_smem = reinterpret_cast<float*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 256 +
                                                 kEpiGroups * 4 * 4096) + q * kMaskMaxK;
            if (masked) {
                for (int c = lane; c < p.K; c += 32) csum_smem[c] = 0.f;
                __syncwarp();
            }
        }
        uint32_t it = 0;
        Unit un = v0 < v1 ? decode_unit<MODE, W>(p, v0) : Unit{};
        [[maybe_unused]] uint32_t mb_next[8];
        [[maybe_unused]] auto load_bits = [&](const Unit& t, uint32_t (&m)[8]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = 0u;
            if (row < t.rows) {
                const uint4* bp = reinterpret_cast<const uint4*>(
                    p.act_bits + int64_t(t.row0 + row) * (p.K / 32) + t.chunk * 8);
                const uint4 lo = __ldg(bp), hi = __ldg(bp + 1);
                m[0] = lo.x; m[1] = lo.y; m[2] = lo.z; m[3] = lo.w;
                m[4] = hi.x; m[5] = hi.y; m[6] = hi.z; m[7] = hi.w;
            }
        };
        if constexpr (MODE == MODE_DH_MN) {
            if (masked && v0 < v1) load_bits(un, mb_next);
        }
        for (int64_t v = v0; v < v1; v += vstep) {
            const Unit u = un;
            if (v + vstep < v1) un = decode_unit<MODE, W>(p, v + vstep);
            if (u.nkb == 0) {
                if (mode_is_du(MODE) && grp == 0) {
                    uint4* dst = reinterpret_cast<uint4*>(
                        static_cast<__nv_bfloat16*>(p.out) +
                        (int64_t(u.group - p.node_base) * p.K + u.chunk * kM + row) * W);
#pragma unroll
                    for (int j = 0; j < W / 8; ++j) dst[j] = make_uint4(0u, 0u, 0u, 0u);
                }
                continue;
            }
            const uint32_t as = it % kAcc;
            const uint32_t aphase = (it / kAcc) & 1u;
            ++it;
            if (int(as % kEpiGroups) != grp) continue;  // the other group's unit (kAcc is even)
            // fused ReLU backward: this thread's row has 256 mask bits per unit (8 words), fetched
            // one unit ahead - under a saturated HBM a load issued at its point of use would sit
            // several microseconds inside the store loop (measured: 2.4 -> 4.2 ms per launch with
            // the bf16 activations themselves loaded here, even with L2 prefetch)
            [[maybe_unused]] uint32_t mb[8];
            if constexpr (MODE == MODE_DH_MN) {
                if (masked) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) mb[j] = mb_next[j];
                    if (v + vstep < v1) load_bits(un, mb_next);
                }
            }
            if (lane == 0) mbar_wait(tfull_bar(as), aphase);
            __syncwarp();
            tc_fence_after();
            const uint32_t tcol = tmem_base + lane_addr + as * kBN;
            // Output rows leave in 128-byte pieces through a per-warp staging tile (32 rows x 128 B,
            // 16-byte chunks XOR-swizzled by the row): the accumulator is read row-per-lane, the
            // global stores go out with 8 consecutive lanes on one row's 128 bytes (whole lines,
            // and for MODE_DU* 512 contiguous bytes per instruction).
            const int rows_here = u.rows - q * 32;  // valid rows among this warp's 32
            uint8_t* gbase;                         // piece 0 of this warp's row 0
            int64_t pitch;                          // bytes between consecutive output rows
            if constexpr (MODE == MODE_MSG) {
                gbase = reinterpret_cast<uint8_t*>(static_cast<float*>(p.out) +
                                                   int64_t(u.row0 + q * 32) * W);
                pitch = W * 4;
            } else if constexpr (mode_is_dh(MODE)) {
                gbase = reinterpret_cast<uint8_t*>(static_cast<__nv_bfloat16*>(p.out) +
                                                   int64_t(u.row0 + q * 32) * p.K + u.chunk * 256);
                pitch = int64_t(p.K) * 2;
            } else {
                gbase = reinterpret_cast<uint8_t*>(
                    static_cast<__nv_bfloat16*>(p.out) +
                    (int64_t(u.group - p.node_base) * p.K + u.chunk * kM + q * 32) * W);
                pitch = W * 2;
            }
            constexpr int kColsPerPiece = MODE == MODE_MSG ? 32 : 64;  // 128 bytes of output
#pragma unroll
            for (int piece = 0; piece < kBN / kColsPerPiece; ++piece) {
                uint32_t w[32];
                if constexpr (MODE == MODE_MSG) {
                    float f[32];
                    tmem_ld32(tcol + piece * 32, f);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(f[j]);
                } else {
                    float f0[32], f1[32];
                    tmem_ld32(tcol + piece * 64, f0);
                    tmem_ld32(tcol + piece * 64 + 32, f1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        w[j] = pack_bf16(f0[2 * j], f0[2 * j + 1]);
                        w[16 + j] = pack_bf16(f1[2 * j], f1[2 * j + 1]);
                    }
                    if constexpr (MODE == MODE_DH_MN) {
                        if (masked) {
                            // word j of the piece = columns 2j, 2j+1: bits (2j, 2j+1) of the
                            // piece's two mask words widen to halfword masks
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const uint32_t two = (mb[piece * 2 + (j >> 4)] >> (2 * (j & 15))) & 3u;
                                w[j] &= (two & 1u) * 0xffffu + (two >> 1) * 0xffff0000u;
                            }
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    sts128(stage_smem + lane * 128 + ((c ^ (lane & 7)) << 4),
                           make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
                }
                __syncwarp();
                float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = j * 4 + (lane >> 3), c = lane & 7;
                    const uint4 t = lds128(stage_smem + r * 128 + ((c ^ (r & 7)) << 4));
                    // (MODE_DU_MN at w = 32: the accumulator is 64 columns wide, 32 are output)
                    if (r < rows_here && (!mode_is_du(MODE) || W >= 64 || c * 8 < W)) {
                        if constexpr (MODE == MODE_DH_MN) {
                            if (masked) {
                                // column sums of the masked values (the bias gradient)
                                cs[0] += __uint_as_float(t.x << 16);
                                cs[1] += __uint_as_float(t.x & 0xffff0000u);
                                cs[2] += __uint_as_float(t.y << 16);
                                cs[3] += __uint_as_float(t.y & 0xffff0000u);
                                cs[4] += __uint_as_float(t.z << 16);
                                cs[5] += __uint_as_float(t.z & 0xffff0000u);
                                cs[6] += __uint_as_float(t.w << 16);
                                cs[7] += __uint_as_float(t.w & 0xffff0000u);
                            }
                        }
                        uint4* gp = reinterpret_cast<uint4*>(gbase + r * pitch + piece * 128 + c * 16);
                        if (p.hints & 2) {
                            __stcs(gp, t);
                        } else {
                            *gp = t;
                        }
                    }
                }
                if constexpr (MODE == MODE_DH_MN) {
                    if (masked) {
                        // lanes l, l+8, l+16, l+24 hold the same 8 columns: fixed-order fold
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 8);
                            cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
                        }
                        if (lane < 8) {
                            float* dst = csum_smem + u.chunk * 256 + piece * 64 + lane * 8;
#pragma unroll
                            for (int i = 0; i < 8; ++i) dst[i] += cs[i];
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        if constexpr (MODE == MODE_DH_MN) {
            if (masked) {
                // the four warps' sums, folded in warp order: one row of the per-CTA partials
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const float* all = csum_smem - q * kMaskMaxK;
                float* dst = p.colsum_part + int64_t(blockIdx.x) * p.K;
                for (int c = threadIdx.x - 64; c < p.K; c += 128) {
                    dst[c] = ((all[c] + all[kMaskMaxK + c]) + all[2 * kMaskMaxK + c]) +
                             all[3 * kMaskMaxK + c];
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

// cache hints per mode: GKN_NODE_HINTS = "msg,du,dh" (three integers, bit 0 = evict_first loads of
// the A operand, bit 1 = streaming output stores); defaults chosen by measurement
int node_hints(int mode) {
    static int cached[3] = {-1, -1, -1};
    if (cached[0] < 0) {
        int v[3] = {0, 0, 0};
        const char* e = getenv("GKN_NODE_HINTS");
        if (e != nullptr) sscanf(e, "%d,%d,%d", &v[0], &v[1], &v[2]);
        cached[1] = v[1];
        cached[2] = v[2];
        cached[0] = v[0];
    }
    return mode == MODE_MSG ? cached[0] : (mode_is_du(mode) ? cached[1] : cached[2]);
}

template <int MODE, int W = kW>
int launch_node_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const NodeParams& p,
                     cudaStream_t stream) {
    auto kern = node_gemm_kernel<MODE, W>;
    constexpr int smem = NCfg<MODE, W>::kSmemBytes;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(node_gemm)", err);
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    int64_t grid = p.v_end - p.v_begin;
    if (grid > sms) grid = sms;
    if (grid < 1) return 0;
    NodeParams q = p;
    q.hints = node_hints(MODE);
    kern<<<dim3(unsigned(grid)), dim3(kThreads), smem, stream>>>(ta, tb, q);
    GKN_CHECK_LAUNCH("node_gemm launch");
    return 0;
}

int check_common(const char* name, int32_t K, int32_t W, int64_t E_pad, bool any_width = false) {
    static thread_local char buf[160];
    if (any_width ? (W != 32 && W != 64 && W != 128) : W != kW) {
        snprintf(buf, sizeof(buf), "%s: width must be %s", name, any_width ? "32, 64 or 128" : "64");
        return set_error(buf);
    }
    if (K <= 0 || K % 256 != 0) {
        snprintf(buf, sizeof(buf), "%s: kernel width K must be a positive multiple of 256", name);
        return set_error(buf);
    }
    if (E_pad < 0 || E_pad % 16 != 0 || E_pad >= (int64_t(1) << 31) - 256) {
        snprintf(buf, sizeof(buf), "%s: padded edge count must be a multiple of 16 below 2^31", name);
        return set_error(buf);
    }
    return 0;
}

// Gs[q, :] = bf16(g[idx[q], :]) (zero row where idx[q] < 0; rows of gs_width(W) values, columns
// >= W zero) and, at w = 64, its transpose GsT[:, q]
template <int W>
__global__ void __launch_bounds__(256)
node_gather_kernel(const float* __restrict__ g, const int32_t* __restrict__ idx, int64_t E_pad,
                   __nv_bfloat16* __restrict__ gs, __nv_bfloat16* __restrict__ gsT) {
    constexpr int Wp = gs_width(W);
    constexpr int kTpr = Wp / 8;          // threads per row (8 columns each: one 16-byte store)
    constexpr int kRpp = 256 / kTpr;      // rows per pass
    __shared__ __align__(16) __nv_bfloat16 tile[W == kW ? 64 : 1][Wp + 8];
    const int t = threadIdx.x;
    const bool transpose = W == kW && gsT != nullptr;  // (uniform) the smem tile and its barriers
    for (int64_t q0 = int64_t(blockIdx.x) * 64; q0 < E_pad; q0 += int64_t(gridDim.x) * 64) {
        if (transpose) __syncthreads();
#pragma unroll
        for (int pass = 0; pass < 64 / kRpp; ++pass) {
            const int r = pass * kRpp + t / kTpr;
            const int c = (t % kTpr) * 8;
            const int64_t qq = q0 + r;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (qq < E_pad && c < W) {
                const int32_t n = __ldg(idx + qq);
                if (n >= 0) {
                    const float4* src = reinterpret_cast<const float4*>(g + int64_t(n) * W + c);
                    v0 = __ldg(src);
                    v1 = __ldg(src + 1);
                }
            }
            const uint4 pk = make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w),
                                        pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
            if (transpose) *reinterpret_cast<uint4*>(&tile[W == kW ? r : 0][c]) = pk;
            if (gs != nullptr && qq < E_pad) *reinterpret_cast<uint4*>(gs + qq * Wp + c) = pk;
        }
        if constexpr (W == kW) {
            if (transpose) {
                __syncthreads();
                // thread -> channel o = t / 4, 16 consecutive positions (E_pad is a multiple of 16)
                const int o = t >> 2;
                const int qb = (t & 3) * 16;
                if (q0 + qb < E_pad) {
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint16_t lo = *reinterpret_cast<const uint16_t*>(&tile[qb + 2 * j][o]);
                        const uint16_t hi = *reinterpret_cast<const uint16_t*>(&tile[qb + 2 * j + 1][o]);
                        w[j] = uint32_t(lo) | (uint32_t(hi) << 16);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(gsT + int64_t(o) * E_pad + q0 + qb);
                    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                }
            }
        }
    }
}

// bits[q, c / 32] bit (c % 32) = act[q, c] > 0 (bf16 pattern positive as a signed halfword: +0, -0
// and negatives clear, +inf set).  One 32-bit word = 64 contiguous bytes of act per thread.
__global__ void __launch_bounds__(256)
relu_bits_kernel(const uint4* __restrict__ act, int64_t words, uint32_t* __restrict__ bits) {
    for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < words;
         i += int64_t(gridDim.x) * 256) {
        uint32_t out = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 v = __ldcs(act + i * 4 + j);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t lo = int16_t(w[k] & 0xffffu) > 0, hi = int16_t(w[k] >> 16) > 0;
                out |= (lo | (hi << 1)) << (j * 8 + k * 2);
            }
        }
        bits[i] = out;
    }
}

// colsum[c] (+)= sum_b part[b, c] in block order (the per-CTA partials of the masked dh epilogue)
__global__ void __launch_bounds__(256)
dh_colsum_fold_kernel(const float* __restrict__ part, int nblocks, int K, float* __restrict__ colsum,
                      int accumulate) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= K) return;
    float s = accumulate ? colsum[c] : 0.f;
    for (int b = 0; b < nblocks; ++b) s += part[int64_t(b) * K + c];
    colsum[c] = s;
}

int node_dh_mn_impl(const char* name, const void* gs_bf16, const void* u_bf16, const int32_t* units,
                    int64_t unit_begin, int64_t unit_end, int64_t E_pad, int64_t chunk_nodes,
                    int32_t node_base, int32_t A, const uint32_t* act_bits, void* dh_bf16,
                    float* colsum, int32_t accumulate, void* workspace, int64_t workspace_bytes,
                    int32_t K, int32_t W, cudaStream_t st) {
    static thread_local char buf[200];
    auto fail = [&](const char* what) {
        snprintf(buf, sizeof(buf), "%s: %s", name, what);
        return set_error(buf);
    };
    if (unit_end <= unit_begin) return 0;
    if (gs_bf16 == nullptr || u_bf16 == nullptr || units == nullptr || dh_bf16 == nullptr) {
        return fail("NULL pointer argument");
    }
    int rc = check_common(name, K, W, E_pad, true);
    if (rc) return rc;
    if (A <= 0) return fail("A must be positive");
    const int64_t a_rows = int64_t(A) * E_pad, b_rows = int64_t(A) * chunk_nodes * W;
    if (chunk_nodes <= 0 || a_rows >= (int64_t(1) << 31) || b_rows >= (int64_t(1) << 31)) {
        return fail("operand too large for one launch (use smaller node chunks)");
    }
    int sms = 0;
    rc = device_sm_count(&sms);
    if (rc) return rc;
    NodeParams p{};
    p.v_begin = unit_begin * (K / 256);
    p.v_end = unit_end * (K / 256);
    const int64_t grid = p.v_end - p.v_begin < sms ? p.v_end - p.v_begin : sms;
    if (act_bits != nullptr) {
        if (K > kMaskMaxK) return fail("the fused activation mask serves K <= 1024");
        if (colsum == nullptr || workspace == nullptr) return fail("NULL colsum / workspace");
        if (workspace_bytes < grid * K * int64_t(sizeof(float))) return fail("workspace too small");
    }
    CUtensorMap ta, tb;
    rc = make_tensor_map_bf16_2d(&ta, gs_bf16, uint64_t(gs_width(W)), uint64_t(a_rows), kKB, kM);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tb, u_bf16, uint64_t(K), uint64_t(b_rows), 64, W < 64 ? W : kKB);
    if (rc) return rc;
    p.units = reinterpret_cast<const int4*>(units);
    p.out = dh_bf16;
    p.K = K;
    p.A = A;
    p.a_rows_per_app = int32_t(E_pad);
    p.b_rows_per_app = int32_t(chunk_nodes * W);
    p.node_base = node_base;
    p.act_bits = act_bits;
    p.colsum_part = static_cast<float*>(workspace);
    switch (W) {
        case 32: rc = launch_node_gemm<MODE_DH_MN, 32>(ta, tb, p, st); break;
        case 64: rc = launch_node_gemm<MODE_DH_MN, 64>(ta, tb, p, st); break;
        default: rc = launch_node_gemm<MODE_DH_MN, 128>(ta, tb, p, st); break;
    }
    if (rc || act_bits == nullptr) return rc;
    dh_colsum_fold_kernel<<<(K + 255) / 256, 256, 0, st>>>(p.colsum_part, int(grid), K, colsum,
                                                          accumulate);
    GKN_CHECK_LAUNCH("dh_colsum_fold launch");
    return 0;
}

}  // namespace

}  // namespace gkn

extern "C" int gkn_node_msg_bf16(const void* hs_bf16, const void* u_bf16, const int32_t* units,
                                 int64_t unit_begin, int64_t unit_end, int64_t E_pad,
                                 int64_t chunk_nodes, int32_t node_base, float* msg, int32_t K,
                                 int32_t W, void* stream) {
    using namespace gkn;
    if (unit_end <= unit_begin) return 0;
    if (hs_bf16 == nullptr || u_bf16 == nullptr || units == nullptr || msg == nullptr) {
        return set_error("gkn_node_msg_bf16: NULL pointer argument");
    }
    int rc = check_common("gkn_node_msg_bf16", K, W, E_pad, true);
    if (rc) return rc;
    if (chunk_nodes <= 0 || chunk_nodes * W >= (int64_t(1) << 31)) {
        return set_error("gkn_node_msg_bf16: bad node chunk");
    }
    CUtensorMap ta, tb;
    rc = make_tensor_map_bf16_2d(&ta, hs_bf16, uint64_t(K), uint64_t(E_pad), kKB, kM);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tb, u_bf16, uint64_t(K), uint64_t(chunk_nodes) * W, kKB, W);
    if (rc) return rc;
    NodeParams p{};
    p.units = reinterpret_cast<const int4*>(units);
    p.out = msg;
    p.v_begin = unit_begin;
    p.v_end = unit_end;
    p.K = K;
    p.node_base = node_base;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (W) {
        case 32: return launch_node_gemm<MODE_MSG, 32>(ta, tb, p, st);
        case 64: return launch_node_gemm<MODE_MSG, 64>(ta, tb, p, st);
        default: return launch_node_gemm<MODE_MSG, 128>(ta, tb, p, st);
    }
}

extern "C" int gkn_node_du_bf16(const void* hsT_bf16, const void* gsT_bf16, const int32_t* pptr,
                                int64_t node_begin, int64_t node_end, int64_t E_pad,
                                void* dut_bf16, int32_t K, int32_t W, void* stream) {
    using namespace gkn;
    if (node_end <= node_begin) return 0;
    if (hsT_bf16 == nullptr || gsT_bf16 == nullptr || pptr == nullptr || dut_bf16 == nullptr) {
        return set_error("gkn_node_du_bf16: NULL pointer argument");
    }
    int rc = check_common("gkn_node_du_bf16", K, W, E_pad);
    if (rc) return rc;
    if (E_pad == 0) return set_error("gkn_node_du_bf16: empty edge set");
    CUtensorMap ta, tb;
    rc = make_tensor_map_bf16_2d(&ta, hsT_bf16, uint64_t(E_pad), uint64_t(K), kKB, kM);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tb, gsT_bf16, uint64_t(E_pad), uint64_t(kW), kKB, kW);
    if (rc) return rc;
    NodeParams p{};
    p.pptr = pptr;
    p.out = dut_bf16;
    p.v_begin = 0;
    p.v_end = (node_end - node_begin) * (K / kM);
    p.K = K;
    p.node_base = int32_t(node_begin);
    return launch_node_gemm<MODE_DU>(ta, tb, p, static_cast<cudaStream_t>(stream));
}

extern "C" int gkn_node_dh_bf16(const void* gs_bf16, const void* ut_bf16, const int32_t* units,
                                int64_t unit_begin, int64_t unit_end, int64_t E_pad,
                                int64_t chunk_nodes, int32_t node_base, int32_t A, void* dh_bf16,
                                int32_t K, int32_t W, void* stream) {
    using namespace gkn;
    if (unit_end <= unit_begin) return 0;
    if (gs_bf16 == nullptr || ut_bf16 == nullptr || units == nullptr || dh_bf16 == nullptr) {
        return set_error("gkn_node_dh_bf16: NULL pointer argument");
    }
    int rc = check_common("gkn_node_dh_bf16", K, W, E_pad);
    if (rc) return rc;
    if (A <= 0) return set_error("gkn_node_dh_bf16: A must be positive");
    const int64_t a_rows = int64_t(A) * E_pad, b_rows = int64_t(A) * chunk_nodes * K;
    if (chunk_nodes <= 0 || a_rows >= (int64_t(1) << 31) || b_rows >= (int64_t(1) << 31)) {
        return set_error("gkn_node_dh_bf16: operand too large for one launch (use smaller node chunks)");
    }
    CUtensorMap ta, tb;
    rc = make_tensor_map_bf16_2d(&ta, gs_bf16, uint64_t(kW), uint64_t(a_rows), kKB, kM);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tb, ut_bf16, uint64_t(kW), uint64_t(b_rows), kKB, 256);
    if (rc) return rc;
    NodeParams p{};
    p.units = reinterpret_cast<const int4*>(units);
    p.out = dh_bf16;
    p.v_begin = unit_begin * (K / 256);
    p.v_end = unit_end * (K / 256);
    p.K = K;
    p.A = A;
    p.a_rows_per_app = int32_t(E_pad);
    p.b_rows_per_app = int32_t(chunk_nodes * K);
    p.node_base = node_base;
    return launch_node_gemm<MODE_DH>(ta, tb, p, static_cast<cudaStream_t>(stream));
}

extern "C" int gkn_node_du_mn_bf16(const void* hs_bf16, const void* gs_bf16, const int32_t* pptr,
                                   int64_t node_begin, int64_t node_end, int64_t E_pad,
                                   void* dut_bf16, int32_t K, int32_t W, void* stream) {
    using namespace gkn;
    if (node_end <= node_begin) return 0;
    if (hs_bf16 == nullptr || gs_bf16 == nullptr || pptr == nullptr || dut_bf16 == nullptr) {
        return set_error("gkn_node_du_mn_bf16: NULL pointer argument");
    }
    int rc = check_common("gkn_node_du_mn_bf16", K, W, E_pad, true);
    if (rc) return rc;
    if (E_pad == 0) return set_error("gkn_node_du_mn_bf16: empty edge set");
    CUtensorMap ta, tb;
    rc = make_tensor_map_bf16_2d(&ta, hs_bf16, uint64_t(K), uint64_t(E_pad), 64, kKB);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tb, gs_bf16, uint64_t(gs_width(W)), uint64_t(E_pad), 64, kKB);
    if (rc) return rc;
    NodeParams p{};
    p.pptr = pptr;
    p.out = dut_bf16;
    p.v_begin = 0;
    p.v_end = (node_end - node_begin) * (K / kM);
    p.K = K;
    p.node_base = int32_t(node_begin);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (W) {
        case 32: return launch_node_gemm<MODE_DU_MN, 32>(ta, tb, p, st);
        case 64: return launch_node_gemm<MODE_DU_MN, 64>(ta, tb, p, st);
        default: return launch_node_gemm<MODE_DU_MN, 128>(ta, tb, p, st);
    }
}

extern "C" int gkn_node_dh_mn_bf16(const void* gs_bf16, const void* u_bf16, const int32_t* units,
                                   int64_t unit_begin, int64_t unit_end, int64_t E_pad,
                                   int64_t chunk_nodes, int32_t node_base, int32_t A,
                                   void* dh_bf16, int32_t K, int32_t W, void* stream) {
    return gkn::node_dh_mn_impl("gkn_node_dh_mn_bf16", gs_bf16, u_bf16, units, unit_begin, unit_end,
                                E_pad, chunk_nodes, node_base, A, nullptr, dh_bf16, nullptr, 0,
                                nullptr, 0, K, W, static_cast<cudaStream_t>(stream));
}

extern "C" int gkn_relu_bits_bf16(const void* act_bf16, int64_t rows, int32_t K, uint32_t* bits,
                                  void* stream) {
    using namespace gkn;
    if (rows == 0) return 0;
    if (act_bf16 == nullptr || bits == nullptr) return set_error("gkn_relu_bits_bf16: NULL pointer");
    if (rows < 0 || K <= 0 || K % 32 != 0) {
        return set_error("gkn_relu_bits_bf16: K must be a positive multiple of 32");
    }
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    const int64_t words = rows * (K / 32);
    int64_t blocks = (words + 255) / 256;
    if (blocks > int64_t(sms) * 16) blocks = int64_t(sms) * 16;
    relu_bits_kernel<<<dim3(unsigned(blocks)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(act_bf16), words, bits);
    GKN_CHECK_LAUNCH("relu_bits launch");
    return 0;
}

extern "C" int gkn_node_dh_mn_relu_bf16(const void* gs_bf16, const void* u_bf16,
                                        const int32_t* units, int64_t unit_begin, int64_t unit_end,
                                        int64_t E_pad, int64_t chunk_nodes, int32_t node_base,
                                        int32_t A, const uint32_t* act_bits, void* dz_bf16,
                                        float* colsum, int32_t accumulate, void* workspace,
                                        int64_t workspace_bytes, int32_t K, int32_t W,
                                        void* stream) {
    if (act_bits == nullptr) {
        return gkn::set_error("gkn_node_dh_mn_relu_bf16: NULL activation mask");
    }
    return gkn::node_dh_mn_impl("gkn_node_dh_mn_relu_bf16", gs_bf16, u_bf16, units, unit_begin,
                                unit_end, E_pad, chunk_nodes, node_base, A, act_bits, dz_bf16,
                                colsum, accumulate, workspace, workspace_bytes, K, W,
                                static_cast<cudaStream_t>(stream));
}

extern "C" int gkn_node_gather_bf16(const float* g, const int32_t* idx, int64_t E_pad,
                                    void* gs_bf16, void* gsT_bf16, int32_t W, void* stream) {
    using namespace gkn;
    if (E_pad == 0) return 0;
    if (g == nullptr || idx == nullptr || (gs_bf16 == nullptr && gsT_bf16 == nullptr)) {
        return set_error("gkn_node_gather_bf16: NULL pointer argument");
    }
    int rc = check_common("gkn_node_gather_bf16", 256, W, E_pad, true);
    if (rc) return rc;
    if (W != kW && gsT_bf16 != nullptr) {
        return set_error("gkn_node_gather_bf16: the transposed copy exists at w = 64 only");
    }
    int sms = 0;
    rc = device_sm_count(&sms);
    if (rc) return rc;
    int64_t blocks = (E_pad + 63) / 64;
    if (blocks > int64_t(sms) * 8) blocks = int64_t(sms) * 8;
    const dim3 grid{unsigned(blocks)};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto* gs = static_cast<__nv_bfloat16*>(gs_bf16);
    auto* gsT = static_cast<__nv_bfloat16*>(gsT_bf16);
    switch (W) {
        case 32: node_gather_kernel<32><<<grid, 256, 0, st>>>(g, idx, E_pad, gs, gsT); break;
        case 64: node_gather_kernel<64><<<grid, 256, 0, st>>>(g, idx, E_pad, gs, gsT); break;
        default: node_gather_kernel<128><<<grid, 256, 0, st>>>(g, idx, E_pad, gs, gsT); break;
    }
    GKN_CHECK_LAUNCH("node_gather launch");
    return 0;
}
