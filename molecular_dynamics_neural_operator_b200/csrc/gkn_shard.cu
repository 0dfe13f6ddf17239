// The one exchange step of the sharded large-frame rollout (sharded.py), fused with
// the node update it feeds and with the broadcast that follows it, over NVLink peer memory.
//
// Source-range sharding of the node-first path leaves every rank with PARTIAL sums of messages (and
// of x_j, for the layers.4 bias term) at all N destinations: S_p [2, N, w] fp32 on rank p.  The
// reference semantics (NNConv_old.update after PyG's mean, graph_kernel.py:194-209) need
//     x'[n] = act( inv_cnt[n] * sum_p S_p[0, n] + (inv_cnt[n] * sum_p S_p[1, n]) @ B3 + x[n] @ root + bias )
// Round 1 did that as ncclAllReduce(S) (2 x 7/8 x 102 MB per rank over the wire at 8 ranks) followed
// by a node-update kernel on every rank.  Here rank r owns the nodes [n0, n1) of the OUTPUT: one
// kernel
//   - loads the partial rows of its nodes straight from every peer's buffer (ld.global on mapped
//     peer pointers; the peers are summed in rank order, so the result does not depend on timing),
//   - applies the node update (two w x w products per node, operands parked in shared memory, four
//     nodes per warp round so matrix elements are reused from registers),
//   - stores the new feature row into the x' buffer of EVERY rank (st.global on peer pointers).
// A reduce-scatter, the compute, and an all-gather in one pass: per rank 2 N w 4 bytes in and
// N w 4 bytes out cross NVLink once, nothing is staged, and no rank computes a node twice.
// With the multicast addresses of the two symmetric buffers (NVLS: NVSwitch multicast objects) the
// sum over the ranks is done BY THE SWITCH - one multimem.ld_reduce per 16 bytes instead of one
// load per peer - and the new row is written once with multimem.st and replicated by the switch:
// the bytes a rank moves over its links drop from (world-1)/world x 3 N w 4 to 3 N w 4 / world.
// The caller brackets the launch with two barriers of the symmetric-memory handle (all partials
// written before / all x' rows landed after).
#include "gkn_internal.h"

#include <stdint.h>

namespace gkn {

namespace {

constexpr int kWarps = 8;
constexpr int kMaxWorld = 16;
constexpr int kNB = 4;  // nodes per warp round

struct PeerPtrs {
    const float* part[kMaxWorld];  // S_p [2, N, w]
    float* xout[kMaxWorld];        // x' [N, w] on rank p
    const float* mc_part;          // multicast address of the S buffers (or NULL: peer loads)
    float* mc_xout;                // multicast address of the x' buffers (or NULL: peer stores)
};

__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

template <int CPL>
__device__ __forceinline__ void ldg_vec(const float* p, float (&v)[CPL]) {
    if constexpr (CPL == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (CPL == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = __ldg(p);
    }
}

template <int CPL>
__device__ __forceinline__ void st_vec(float* p, const float (&v)[CPL]) {
    if constexpr (CPL == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (CPL == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
        p[0] = v[0];
    }
}

// peer loads must not be served from a stale L1 line of an earlier application: volatile = .cv-like
__device__ __forceinline__ float4 ld_peer(const float* p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}

// sum over all ranks' copies of 16 bytes, computed in the NVSwitch
__device__ __forceinline__ float4 ld_reduce_mc(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}

// one store, replicated into every rank's copy by the NVSwitch
template <int CPL>
__device__ __forceinline__ void st_mc(float* p, const float (&v)[CPL]) {
    if constexpr (CPL == 4) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p),
                     "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                     : "memory");
    } else if constexpr (CPL == 2) {
        asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]),
                     "f"(v[1])
                     : "memory");
    } else {
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory");
    }
}

template <int W>
__global__ void __launch_bounds__(kWarps * 32)
shard_reduce_update_kernel(const PeerPtrs pp, int world, const float* __restrict__ x,
                           const float* __restrict__ inv_cnt, const float* __restrict__ B3,
                           const float* __restrict__ root, const float* __restrict__ bias,
                           int64_t n0, int64_t n1, int64_t N, int relu) {
    constexpr int LPR = W / 4;                 // lanes per row of w floats
    constexpr int RPW = 32 / LPR;              // rows (= nodes) a warp loads side by side
    constexpr int CPL = W >= 32 ? W / 32 : 1;  // output columns per lane in the product phase
    __shared__ __align__(16) float s_r[kWarps][kNB][W];
    __shared__ __align__(16) float s_g[kWarps][kNB][W];
    __shared__ __align__(16) float s_n[kWarps][kNB][W];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int sub = lane / LPR;
    const int col = (lane % LPR) * 4;
    const int64_t warp0 = int64_t(blockIdx.x) * kWarps + wib;
    const int64_t nwarps = int64_t(gridDim.x) * kWarps;
    const int64_t rounds = (n1 - n0 + kNB - 1) / kNB;
    for (int64_t rd = warp0; rd < rounds; rd += nwarps) {
        const int64_t base = n0 + rd * kNB;
        // ---- reduce over the peers: lanes [sub*LPR, (sub+1)*LPR) take node base + sub (+ RPW)
#pragma unroll
        for (int nb = sub; nb < kNB; nb += RPW) {
            const int64_t n = base + nb;
            if (n >= n1) continue;
            float4 ar = make_float4(0.f, 0.f, 0.f, 0.f), ag = ar;
            if (pp.mc_part != nullptr) {  // the switch adds the ranks' rows
                ar = ld_reduce_mc(pp.mc_part + n * W + col);
                ag = ld_reduce_mc(pp.mc_part + (N + n) * W + col);
            } else {
                for (int p = 0; p < world; ++p) {  // rank order: a fixed summation order
                    ar = f4add(ar, ld_peer(pp.part[p] + n * W + col));
                    ag = f4add(ag, ld_peer(pp.part[p] + (N + n) * W + col));
                }
            }
            const float sc = inv_cnt[n];
            *reinterpret_cast<float4*>(&s_r[wib][nb][col]) = make_float4(ar.x * sc, ar.y * sc, ar.z * sc, ar.w * sc);
            *reinterpret_cast<float4*>(&s_g[wib][nb][col]) = make_float4(ag.x * sc, ag.y * sc, ag.z * sc, ag.w * sc);
            *reinterpret_cast<float4*>(&s_n[wib][nb][col]) =
                __ldg(reinterpret_cast<const float4*>(x + n * W + col));
        }
        __syncwarp();
        // ---- node update for the kNB nodes together: every matrix element fetched is used kNB times
        const int c0 = lane * CPL;
        if (c0 < W) {
            const int nvalid = int(n1 - base < kNB ? n1 - base : kNB);
            float o[kNB][CPL], bv[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) bv[k] = __ldg(bias + c0 + k);
#pragma unroll
            for (int nb = 0; nb < kNB; ++nb)
#pragma unroll
                for (int k = 0; k < CPL; ++k) o[nb][k] = nb < nvalid ? s_r[wib][nb][c0 + k] + bv[k] : 0.f;
#pragma unroll 4
            for (int i = 0; i < W; ++i) {
                float m1[CPL], m2[CPL];
                ldg_vec<CPL>(B3 + i * W + c0, m1);
                ldg_vec<CPL>(root + i * W + c0, m2);
#pragma unroll
                for (int nb = 0; nb < kNB; ++nb) {
                    const float a = s_g[wib][nb][i], b = s_n[wib][nb][i];
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        o[nb][k] = fmaf(a, m1[k], o[nb][k]);
                        o[nb][k] = fmaf(b, m2[k], o[nb][k]);
                    }
                }
            }
            // ---- broadcast: the new row goes to every rank's x' buffer
#pragma unroll
            for (int nb = 0; nb < kNB; ++nb) {
                if (nb >= nvalid) continue;
                float v[CPL];
#pragma unroll
                for (int k = 0; k < CPL; ++k) v[k] = relu ? fmaxf(o[nb][k], 0.f) : o[nb][k];
                if (pp.mc_xout != nullptr) {
                    st_mc<CPL>(pp.mc_xout + (base + nb) * W + c0, v);
                } else {
                    for (int p = 0; p < world; ++p) st_vec<CPL>(pp.xout[p] + (base + nb) * W + c0, v);
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace

}  // namespace gkn

extern "C" int gkn_shard_reduce_update(const uint64_t* partial_ptrs, const uint64_t* xout_ptrs,
                                       uint64_t mc_partial, uint64_t mc_xout, int32_t world,
                                       const float* x, const float* inv_cnt,
                                       const float* B3, const float* root, const float* bias,
                                       int64_t n0, int64_t n1, int64_t N, int32_t W, int32_t relu,
                                       void* stream) {
    using namespace gkn;
    if (n1 <= n0) return 0;
    if (partial_ptrs == nullptr || xout_ptrs == nullptr || x == nullptr || inv_cnt == nullptr ||
        B3 == nullptr || root == nullptr || bias == nullptr) {
        return set_error("gkn_shard_reduce_update: NULL pointer argument");
    }
    if (world < 1 || world > kMaxWorld) {
        return set_error("gkn_shard_reduce_update: world size must be in [1, 16]");
    }
    if (n0 < 0 || n1 > N) return set_error("gkn_shard_reduce_update: node range outside [0, N)");
    PeerPtrs pp{};
    pp.mc_part = reinterpret_cast<const float*>(mc_partial);
    pp.mc_xout = reinterpret_cast<float*>(mc_xout);
    for (int p = 0; p < world; ++p) {
        if (partial_ptrs[p] == 0 || xout_ptrs[p] == 0) {
            return set_error("gkn_shard_reduce_update: NULL peer buffer");
        }
        pp.part[p] = reinterpret_cast<const float*>(partial_ptrs[p]);
        pp.xout[p] = reinterpret_cast<float*>(xout_ptrs[p]);
    }
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    const int64_t rounds = (n1 - n0 + kNB - 1) / kNB;
    int64_t blocks = (rounds + kWarps - 1) / kWarps;
    if (blocks > int64_t(sms) * 8) blocks = int64_t(sms) * 8;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define GKN_SRU(WW)                                                                               \
    shard_reduce_update_kernel<WW><<<dim3(unsigned(blocks)), kWarps * 32, 0, st>>>(               \
        pp, world, x, inv_cnt, B3, root, bias, n0, n1, N, relu)
    switch (W) {
        case 32: GKN_SRU(32); break;
        case 64: GKN_SRU(64); break;
        case 128: GKN_SRU(128); break;
        default: return set_error("gkn_shard_reduce_update: width must be 32, 64 or 128");
    }
#undef GKN_SRU
    GKN_CHECK_LAUNCH("shard_reduce_update launch");
    return 0;
}
