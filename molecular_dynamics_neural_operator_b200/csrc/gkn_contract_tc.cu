// K1 / K3 — fused bilinear edge contraction on tcgen05 tensor cores (sm_100a).
//
// Replaces NNConv_old.message (reference graph_kernel.py:200-202):
//     weight = net(pseudo).view(E, w, w);  msg = (x_j[:,None,:] @ weight)[:,0,:]
// The reference materialises the [E, w, w] per-edge kernel in HBM.  Here the last edge-MLP layer
//     P[e, n] = sum_k h_e[e, k] * W3[n, k]          (n = i*w + o, W3 = net.layers.4.weight [w*w, K])
// runs as a bf16 tcgen05 GEMM whose 128 x 256 fp32 accumulator chunk lives only in TMEM; the
// epilogue warps read it back with tcgen05.ld and contract it immediately against the neighbour
// features (forward) or the message gradient (backward-dx):
//     MODE_FWD : msg[e, o] = sum_i x[src[e], i] * P[e, i*w + o]
//     MODE_DX  : dxj[e, i] = sum_o P[e, i*w + o] * g[dst[e], o]
// so P never reaches HBM (flash-style recompute in backward).  The b3 bias term of layers.4 is a
// node-level product and is applied by the scatter kernels, not here.
//
// Structure (one persistent CTA per SM, 192 threads; CTAs run as PAIRS, cluster of 2):
//   warp 0      TMA producer: per k-block its own h_e tile [128 x 64] and HALF of the W3 chunk tile
//               ([128 x 64] of the 256 rows), 6 stages of 32 KB, 128-byte swizzle; both CTAs'
//               bytes complete on the leader's full[] mbarrier.
//   warp 1      TMEM allocator; in the leader CTA, the single-thread issuer of
//               tcgen05.mma.cta_group::2 (M = 256 edges across the pair, N = 256, K = 16 per MMA);
//               multicast tcgen05.commit releases smem stages (empty[]) and publishes
//               accumulators (tfull[]) in both CTAs.
//   warps 2..5  epilogue: 128 threads = this CTA's 128 accumulator rows (edges); two TMEM
//               accumulator stages (2 x 256 columns) so the contraction of chunk c overlaps the
//               MMAs of chunk c+1; the peer's warps release a stage by a remote mbarrier arrive.
// Pairing cuts operand bytes per flop by a third (measured: tensor pipe 91 % of peak cycles,
// 1.6 PFLOP/s isolated, 2.0 PFLOP/s inside the training step where the lower power draw buys
// clock).  GKN_CONTRACT_CTA_GROUP=1 selects the single-CTA variant (4 stages of 48 KB).
// Work unit = (tile group, column chunk).  Units are split evenly over CTA pairs (contiguous
// ranges), so there is no tile-quantisation tail; a tile shared by two pairs is combined with fp32
// reductions onto the zero-initialised output (at most two addends per element =>
// order-independent).
#include <stdlib.h>

#include "gkn_internal.h"
#include "gkn_ptx.cuh"

namespace gkn {

namespace {

constexpr int kBlockM = 128;  // accumulator rows (edges) per CTA
constexpr int kBlockN = 256;  // accumulator columns per MMA
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kTmemCols = 512;

// CG = 1: one CTA per 128-edge tile, B tile = 256 rows.  CG = 2: a CTA pair (cluster of 2) shares one
// tcgen05.mma.cta_group::2 over 256 edges x 256 columns: each CTA stages its own 128 edge rows
// and HALF of the W3 chunk (128 rows), so operand bytes per flop drop by a third and the freed
// smem deepens the pipeline (6 stages of 32 KB instead of 4 of 48 KB).
template <int CG>
struct Cfg {
    static constexpr int kBRows = kBlockN / CG;
    static constexpr int kBBytes = kBRows * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = CG == 1 ? 4 : 6;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

enum { MODE_FWD = 0, MODE_DX = 1 };

struct ContractParams {
    const int32_t* __restrict__ idx;  // FWD: src[e]; DX: dst[e]  (row of `node` to gather)
    const float* __restrict__ node;   // FWD: x [N, W]; DX: g [N, W]
    float* __restrict__ out;          // FWD: msg [E, W]; DX: dxj [E, W]
    int64_t E;
    int32_t num_k_blocks;  // K / 64
    int64_t num_tiles;     // ceil(E / 128)
};

template <int W, int MODE, int CG>
__global__ void __launch_bounds__(kThreads, 1)
contract_tc_kernel(const __grid_constant__ CUtensorMap tmap_h,
                   const __grid_constant__ CUtensorMap tmap_w, const ContractParams p) {
    using C = Cfg<CG>;
    constexpr int kStages = C::kStages;
    constexpr int kStageBytes = C::kStageBytes;
    constexpr int kChunks = (W * W) / kBlockN;  // column chunks per edge tile
    constexpr int kIPerChunk = kBlockN / W;      // i-values covered by one chunk
    static_assert((W * W) % kBlockN == 0 && kBlockN % W == 0 && W % 32 == 0, "unsupported width");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages * kStageBytes;
    // barrier layout: full[kStages], empty[kStages], tfull[2], tempty[2], tmem_ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
    const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;  // 0 = leader of the pair
    const int64_t group = blockIdx.x / CG;                   // CTA (pair) index
    const int64_t ngroups = gridDim.x / CG;

    // Contiguous unit range of this CTA (pair); unit = (tile group, column chunk), a tile group
    // being CG consecutive 128-edge tiles.  The forward's ranges are always tile-group aligned: a
    // message row is then summed over ALL its column chunks by one epilogue thread and stored once
    // - no partial sums combined through global atomics, no zero-fill of msg beforehand, and the
    // result does not depend on the grid (quantisation tail <= one tile group per CTA: 2 % at the
    // bench shape).  With many tile groups per CTA the same alignment also keeps all CTAs walking
    // the W3 column chunks in phase, so a chunk fetched by one CTA is an L2 hit for the others
    // (matters when W3 is large, e.g. 32 MB at w=128); otherwise the dx units, whose outputs are
    // disjoint per chunk, are split evenly.
    const int64_t num_tgroups = (p.num_tiles + CG - 1) / CG;
    const int64_t total_units = num_tgroups * kChunks;
    int64_t u0, u1;
    if (MODE == MODE_FWD || num_tgroups >= ngroups * 64) {
        u0 = ((num_tgroups * group) / ngroups) * kChunks;
        u1 = ((num_tgroups * (group + 1)) / ngroups) * kChunks;
    } else {
        u0 = (total_units * group) / ngroups;
        u1 = (total_units * (group + 1)) / ngroups;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4 * CG);  // one arrive per epilogue warp of every CTA
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_h);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1) {
        if constexpr (CG == 2) {
            tmem_alloc_pair(tmem_ptr_smem, kTmemCols);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_ptr_smem, kTmemCols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) {
        cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA signal
    } else {
        __syncthreads();
    }
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

    if (warp == 0) {
        // ===================== TMA producer (every CTA loads its own operands) ==============
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = u0; u < u1; ++u) {
                const int32_t tile = int32_t(u / kChunks) * CG + int32_t(rank);
                const int32_t chunk = int32_t(u % kChunks);
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    if constexpr (CG == 2) {
                        // both CTAs' bytes complete on the leader's barrier; only it arms it
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * kStageBytes);
                        tma_load_2d_pair(sa, &tmap_h, full_bar(stage), kb * kBlockK,
                                         tile * kBlockM);
                        tma_load_2d_pair(sa + kABytes, &tmap_w, full_bar(stage), kb * kBlockK,
                                         chunk * kBlockN + int32_t(rank) * C::kBRows);
                    } else {
                        mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
                        tma_load_2d(sa, &tmap_h, full_bar(stage), kb * kBlockK, tile * kBlockM);
                        tma_load_2d(sa + kABytes, &tmap_w, full_bar(stage), kb * kBlockK,
                                    chunk * kBlockN);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only when paired) =====================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * CG, kBlockN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int64_t u = u0; u < u1; ++u, ++it) {
                const uint32_t as = it & 1u;
                const uint32_t aphase = (it >> 1) & 1u;
                mbar_wait(tempty_bar(as), aphase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * kBlockN;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    const uint64_t adesc = umma_smem_desc(sa, 16, 1024, kSwizzle128B);
                    const uint64_t bdesc = umma_smem_desc(sa + kABytes, 16, 1024, kSwizzle128B);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        // advance 16 bf16 (= 32 B) along K inside the 128-byte swizzle atom
                        const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                        if constexpr (CG == 2) {
                            umma_bf16_ss_pair(tmem_d, adesc + uint64_t(2 * k),
                                              bdesc + uint64_t(2 * k), idesc, acc);
                        } else {
                            umma_bf16_ss(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k),
                                         idesc, acc);
                        }
                    }
                    // smem stage reusable (in both CTAs) once these MMAs retire
                    if constexpr (CG == 2) {
                        umma_commit_pair(empty_bar(stage), 3);
                    } else {
                        umma_commit(empty_bar(stage));
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                // accumulator chunk complete (each CTA's epilogue reads its own 128 rows)
                if constexpr (CG == 2) {
                    umma_commit_pair(tfull_bar(as), 3);
                } else {
                    umma_commit(tfull_bar(as));
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const uint32_t lane_addr = uint32_t(q * 32) << 16;

        float acc[MODE == MODE_FWD ? W : 1];  // FWD: msg row accumulator
        float gvec[MODE == MODE_DX ? W : 1];  // DX : message-gradient row
        int32_t cur_tile = -1;
        int64_t e = 0;
        int32_t nidx = 0;
        bool valid = false;

        auto flush_fwd = [&]() {
            if constexpr (MODE == MODE_FWD) {
                if (cur_tile < 0 || !valid) return;
                float* dst = p.out + e * W;  // all kChunks chunks of the tile were summed here
#pragma unroll
                for (int o = 0; o < W; o += 4) {
                    *reinterpret_cast<float4*>(dst + o) =
                        make_float4(acc[o], acc[o + 1], acc[o + 2], acc[o + 3]);
                }
            }
        };

        uint32_t it = 0;
        for (int64_t u = u0; u < u1; ++u, ++it) {
            const int32_t tile = int32_t(u / kChunks) * CG + int32_t(rank);
            const int32_t chunk = int32_t(u % kChunks);
            if (tile != cur_tile) {
                flush_fwd();
                cur_tile = tile;
                e = int64_t(tile) * kBlockM + row;
                valid = e < p.E;
                nidx = valid ? p.idx[e] : 0;
                if constexpr (MODE == MODE_FWD) {
#pragma unroll
                    for (int o = 0; o < W; ++o) acc[o] = 0.f;
                } else {
                    const float4* gp = reinterpret_cast<const float4*>(p.node + int64_t(nidx) * W);
#pragma unroll
                    for (int o = 0; o < W; o += 4) {
                        const float4 t = __ldg(gp + (o >> 2));
                        gvec[o] = t.x;
                        gvec[o + 1] = t.y;
                        gvec[o + 2] = t.z;
                        gvec[o + 3] = t.w;
                    }
                }
            }
            // FWD: neighbour features for the i-values of this chunk, fetched before the wait.
            // DX : one dot product per i-value of this chunk.
            float xv[kIPerChunk];
#pragma unroll
            for (int i = 0; i < kIPerChunk; ++i) xv[i] = 0.f;
            if constexpr (MODE == MODE_FWD) {
                const float* xp = p.node + int64_t(nidx) * W + chunk * kIPerChunk;
#pragma unroll
                for (int i = 0; i < kIPerChunk; ++i) xv[i] = __ldg(xp + i);
            }

            const uint32_t as = it & 1u;
            const uint32_t aphase = (it >> 1) & 1u;
            // one lane polls; the other 31 park at the warp barrier instead of hammering smem
            if (lane == 0) mbar_wait(tfull_bar(as), aphase);
            __syncwarp();
            tc_fence_after();
            const uint32_t tcol = tmem_base + lane_addr + as * kBlockN;
#pragma unroll
            for (int piece = 0; piece < kBlockN / 32; ++piece) {
                float v[32];
                tmem_ld32(tcol + piece * 32, v);
                tmem_ld_wait();
                // column n = piece*32 + j  <->  (i_local, o) = (n / W, n % W)
                constexpr int kPiecesPerI = W / 32;
                const int il = piece / kPiecesPerI;
                const int ob = (piece % kPiecesPerI) * 32;
                if constexpr (MODE == MODE_FWD) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[ob + j] = fmaf(xv[il], v[j], acc[ob + j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) xv[il] = fmaf(v[j], gvec[ob + j], xv[il]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                // the accumulator stage is free once every epilogue warp of the pair has read it
                if (CG == 2 && rank != 0) {
                    mbar_arrive_remote(tempty_bar(as), 0);
                } else {
                    mbar_arrive(tempty_bar(as));
                }
            }

            if constexpr (MODE == MODE_DX) {
                if (valid) {
                    float* dst = p.out + e * W + chunk * kIPerChunk;
                    if constexpr (kIPerChunk % 4 == 0) {
#pragma unroll
                        for (int i = 0; i < kIPerChunk; i += 4) {
                            *reinterpret_cast<float4*>(dst + i) =
                                make_float4(xv[i], xv[i + 1], xv[i + 2], xv[i + 3]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < kIPerChunk; ++i) dst[i] = xv[i];
                    }
                }
            }
        }
        flush_fwd();
    }

    tc_fence_before();
    if constexpr (CG == 2) {
        cluster_sync_all();  // the peer may still signal this CTA's barriers / read its smem
    } else {
        __syncthreads();
    }
    if (warp == 1) {
        tc_fence_after();
        if constexpr (CG == 2) {
            tmem_dealloc_pair(tmem_base, kTmemCols);
        } else {
            tmem_dealloc(tmem_base, kTmemCols);
        }
    }
}

template <int W, int MODE, int CG>
int launch_contract(const CUtensorMap& tmap_h, const CUtensorMap& tmap_w, const ContractParams& p,
                    int max_ctas, cudaStream_t stream) {
    auto kern = contract_tc_kernel<W, MODE, CG>;
    constexpr int smem = Cfg<CG>::kSmemBytes;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(contract_tc)", err);
    // grid <= number of tile groups: every CTA (pair) owns at least one tile group
    const int64_t num_tgroups = (p.num_tiles + CG - 1) / CG;
    int64_t groups = max_ctas / CG;
    if (groups > num_tgroups) groups = num_tgroups;
    if (groups < 1) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(groups * CG));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg, kern, tmap_h, tmap_w, p);
    if (err != cudaSuccess) return set_cuda_error("contract_tc launch", err);
    count_launch();
    return 0;
}

// CTA pairing: on by default; GKN_CONTRACT_CTA_GROUP=1 selects the single-CTA kernel
int contract_cta_group() {
    static int cached = 0;
    if (cached == 0) {
        const char* v = getenv("GKN_CONTRACT_CTA_GROUP");
        cached = (v != nullptr && v[0] == '1') ? 1 : 2;
    }
    return cached;
}

template <int MODE>
int contract_dispatch(const void* h_bf16, const void* w3_bf16, const int32_t* idx,
                      const float* node, float* out, int64_t E, int32_t K, int32_t W,
                      cudaStream_t stream) {
    if (E == 0) return 0;
    if (E < 0) return set_error("gkn_contract_*_bf16: negative edge count");
    if (h_bf16 == nullptr || w3_bf16 == nullptr || idx == nullptr || node == nullptr ||
        out == nullptr) {
        return set_error("gkn_contract_*_bf16: NULL pointer argument");
    }
    if (K <= 0 || K % kBlockK != 0) {
        return set_error("gkn_contract_*_bf16: kernel_width K must be a positive multiple of 64");
    }
    if (W != 32 && W != 64 && W != 128) {
        return set_error("gkn_contract_*_bf16: width must be 32, 64 or 128");
    }
    if (E >= (int64_t(1) << 31) - 2 * kBlockM) {
        return set_error("gkn_contract_*_bf16: too many edges for one launch (>= 2^31)");
    }
    const int cg = contract_cta_group();
    CUtensorMap tmap_h, tmap_w;
    int rc = make_tensor_map_bf16_2d(&tmap_h, h_bf16, uint64_t(K), uint64_t(E), kBlockK, kBlockM);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tmap_w, w3_bf16, uint64_t(K), uint64_t(W) * uint64_t(W), kBlockK,
                                 kBlockN / cg);
    if (rc) return rc;
    ContractParams p;
    p.idx = idx;
    p.node = node;
    p.out = out;
    p.E = E;
    p.num_k_blocks = K / kBlockK;
    p.num_tiles = (E + kBlockM - 1) / kBlockM;
    int sms = 0;
    rc = device_sm_count(&sms);
    if (rc) return rc;
#define GKN_LAUNCH(WW)                                                                  \
    return cg == 2 ? launch_contract<WW, MODE, 2>(tmap_h, tmap_w, p, sms, stream)       \
                   : launch_contract<WW, MODE, 1>(tmap_h, tmap_w, p, sms, stream)
    switch (W) {
        case 32: GKN_LAUNCH(32);
        case 64: GKN_LAUNCH(64);
        default: GKN_LAUNCH(128);
    }
#undef GKN_LAUNCH
}

}  // namespace

}  // namespace gkn

extern "C" int gkn_contract_fwd_bf16(const void* h_bf16, const void* w3_bf16, const int32_t* src,
                                     const float* x, float* msg, int64_t E, int32_t K, int32_t W,
                                     void* stream) {
    return gkn::contract_dispatch<gkn::MODE_FWD>(h_bf16, w3_bf16, src, x, msg, E, K, W,
                                                 static_cast<cudaStream_t>(stream));
}

extern "C" int gkn_contract_dx_bf16(const void* h_bf16, const void* w3_bf16, const int32_t* dst,
                                    const float* g, float* dxj, int64_t E, int32_t K, int32_t W,
                                    void* stream) {
    return gkn::contract_dispatch<gkn::MODE_DX>(h_bf16, w3_bf16, dst, g, dxj, E, K, W,
                                                static_cast<cudaStream_t>(stream));
}
