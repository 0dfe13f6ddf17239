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
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: h_e tile [128 x 64] + W3 chunk tile [256 x 64] per k-block, 4 stages,
//               128-byte swizzle, completion on full[] mbarriers.
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=256, K=16 per MMA);
//               tcgen05.commit releases smem stages (empty[]) and publishes accumulators (tfull[]).
//   warps 2..5  epilogue: 128 threads = 128 accumulator rows (edges); two TMEM accumulator stages
//               (2 x 256 columns) so the contraction of chunk c overlaps the MMAs of chunk c+1.
// Work unit = (edge tile, column chunk).  Units are split evenly over CTAs (contiguous ranges), so
// there is no tile-quantisation tail; a tile shared by two CTAs is combined with fp32 reductions
// onto the zero-initialised output (at most two addends per element => order-independent).
#include "gkn_internal.h"
#include "gkn_ptx.cuh"

namespace gkn {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kStages = 4;
constexpr int kThreads = 192;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kBBytes = kBlockN * kBlockK * 2;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kTmemCols = 512;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

enum { MODE_FWD = 0, MODE_DX = 1 };

struct ContractParams {
    const int32_t* __restrict__ idx;  // FWD: src[e]; DX: dst[e]  (row of `node` to gather)
    const float* __restrict__ node;   // FWD: x [N, W]; DX: g [N, W]
    float* __restrict__ out;          // FWD: msg [E, W]; DX: dxj [E, W]
    int64_t E;
    int32_t num_k_blocks;  // K / 64
    int64_t num_tiles;     // ceil(E / 128)
};

template <int W, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
contract_tc_kernel(const __grid_constant__ CUtensorMap tmap_h,
                   const __grid_constant__ CUtensorMap tmap_w, const ContractParams p) {
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

    // Contiguous unit range of this CTA.  With many tiles per CTA the ranges are tile-aligned: all
    // CTAs then walk the W3 column chunks in phase, so a chunk fetched by one CTA is an L2 hit for
    // the others (matters when W3 is large, e.g. 32 MB at w=128), at a tail cost < 1/64.  With few
    // tiles per CTA, units are split evenly instead (no tile-quantisation tail).
    const int64_t total_units = p.num_tiles * kChunks;
    int64_t u0, u1;
    if (p.num_tiles >= int64_t(gridDim.x) * 64) {
        u0 = ((p.num_tiles * blockIdx.x) / gridDim.x) * kChunks;
        u1 = ((p.num_tiles * (blockIdx.x + 1)) / gridDim.x) * kChunks;
    } else {
        u0 = (total_units * blockIdx.x) / gridDim.x;
        u1 = (total_units * (blockIdx.x + 1)) / gridDim.x;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_h);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, kTmemCols);
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
            for (int64_t u = u0; u < u1; ++u) {
                const int32_t tile = int32_t(u / kChunks);
                const int32_t chunk = int32_t(u % kChunks);
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    tma_load_2d(sa, &tmap_h, full_bar(stage), kb * kBlockK, tile * kBlockM);
                    tma_load_2d(sa + kABytes, &tmap_w, full_bar(stage), kb * kBlockK,
                                chunk * kBlockN);
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
            constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kBlockN, 0, 0);
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
                        umma_bf16_ss(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k),
                                     idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // smem stage reusable once these MMAs retire
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull_bar(as));  // accumulator chunk complete
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
        int chunks_in_tile = 0;
        int64_t e = 0;
        int32_t nidx = 0;
        bool valid = false;

        auto flush_fwd = [&]() {
            if constexpr (MODE == MODE_FWD) {
                if (cur_tile < 0 || !valid) return;
                float* dst = p.out + e * W;
                if (chunks_in_tile == kChunks) {
#pragma unroll
                    for (int o = 0; o < W; o += 4) {
                        *reinterpret_cast<float4*>(dst + o) =
                            make_float4(acc[o], acc[o + 1], acc[o + 2], acc[o + 3]);
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < W; ++o) atomicAdd(dst + o, acc[o]);
                }
            }
        };

        uint32_t it = 0;
        for (int64_t u = u0; u < u1; ++u, ++it) {
            const int32_t tile = int32_t(u / kChunks);
            const int32_t chunk = int32_t(u % kChunks);
            if (tile != cur_tile) {
                flush_fwd();
                cur_tile = tile;
                chunks_in_tile = 0;
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
            ++chunks_in_tile;

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
            if (lane == 0) mbar_arrive(tempty_bar(as));

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
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int W, int MODE>
int launch_contract(const CUtensorMap& tmap_h, const CUtensorMap& tmap_w, const ContractParams& p,
                    int max_ctas, cudaStream_t stream) {
    auto kern = contract_tc_kernel<W, MODE>;
    cudaError_t err =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(contract_tc)", err);
    // grid <= num_tiles: every CTA then owns at least one tile's worth of units, so a tile is
    // shared by at most two CTAs and its two partial sums combine order-independently.
    int grid = int(p.num_tiles < max_ctas ? p.num_tiles : max_ctas);
    if (grid < 1) return 0;
    kern<<<grid, kThreads, kSmemBytes, stream>>>(tmap_h, tmap_w, p);
    err = cudaGetLastError();
    if (err != cudaSuccess) return set_cuda_error("contract_tc launch", err);
    count_launch();
    return 0;
}

template <int MODE>
int contract_dispatch(const void* h_bf16, const void* w3_bf16, const int32_t* idx,
                      const float* node, float* out, int64_t E, int32_t K, int32_t W,
                      cudaStream_t stream) {
    if (E == 0) return 0;
    if (K <= 0 || K % kBlockK != 0) {
        return set_error("gkn_contract_*_bf16: kernel_width K must be a positive multiple of 64");
    }
    if (W != 32 && W != 64 && W != 128) {
        return set_error("gkn_contract_*_bf16: width must be 32, 64 or 128");
    }
    if (E >= (int64_t(1) << 31) - kBlockM) {
        return set_error("gkn_contract_*_bf16: too many edges for one launch (>= 2^31)");
    }
    CUtensorMap tmap_h, tmap_w;
    int rc = make_tensor_map_bf16_2d(&tmap_h, h_bf16, uint64_t(K), uint64_t(E), kBlockK, kBlockM);
    if (rc) return rc;
    rc = make_tensor_map_bf16_2d(&tmap_w, w3_bf16, uint64_t(K), uint64_t(W) * uint64_t(W), kBlockK,
                                 kBlockN);
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
    if (MODE == MODE_FWD) {
        // tiles split across two CTAs are combined with reductions onto zero
        cudaError_t err = cudaMemsetAsync(out, 0, size_t(E) * W * sizeof(float), stream);
        if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(msg)", err);
    }
    switch (W) {
        case 32: return launch_contract<32, MODE>(tmap_h, tmap_w, p, sms, stream);
        case 64: return launch_contract<64, MODE>(tmap_h, tmap_w, p, sms, stream);
        default: return launch_contract<128, MODE>(tmap_h, tmap_w, p, sms, stream);
    }
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
