// fp32 CUDA-core variant of the fused bilinear edge contraction (precision="fp32" mode).
//
// Same math as gkn_contract_tc.cu (reference graph_kernel.py:200-202) but with fp32 operands and
// fp32 FMA accumulation, so it meets the 1e-5 relative parity bar against the fp64 oracle that
// bf16 tensor-core operands cannot.  It is also the on-device cross-check for the tcgen05 path.
// The [E, w, w] kernel tensor is still never materialised: a 64-edge x 64-column slab of P for one
// input channel i lives in registers, is contracted, and is discarded.
#include "gkn_internal.h"

namespace gkn {

namespace {

constexpr int kTE = 64;  // edges per block
constexpr int kKS = 16;  // k-slab per smem stage
enum { MODE_FWD = 0, MODE_DX = 1 };

template <int W, int MODE>
__global__ void __launch_bounds__(256)
contract_simt_kernel(const float* __restrict__ h, const float* __restrict__ w3,
                     const int32_t* __restrict__ idx, const float* __restrict__ node,
                     float* __restrict__ out, int64_t E, int K) {
    constexpr int OT = W < 64 ? W : 64;  // output-channel tile
    constexpr int TO = OT / 16;          // per-thread columns
    constexpr int NOT = W / OT;          // o-tiles
    __shared__ float hs[kKS][kTE + 4];
    __shared__ float ws[kKS][OT + 4];
    __shared__ float ns[kTE][W + 1];  // gathered node rows: x[src] (FWD) or g[dst] (DX)

    const int tx = threadIdx.x & 15;
    const int ty = threadIdx.x >> 4;
    const int64_t e0 = int64_t(blockIdx.x) * kTE;

    for (int t = threadIdx.x; t < kTE * W; t += 256) {
        const int r = t / W, c = t % W;
        const int64_t e = e0 + r;
        ns[r][c] = e < E ? node[int64_t(idx[e]) * W + c] : 0.f;
    }
    __syncthreads();

    // computes Pacc[r][c] = sum_k h[e0+ty*4+r, k] * w3[i*W + ot*OT + tx*TO + c, k]
    auto p_tile = [&](int i, int ot, float (&Pacc)[4][TO]) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < TO; ++c) Pacc[r][c] = 0.f;
        for (int kk = 0; kk < K; kk += kKS) {
            // h slab: 64 edges x 16 k  (each thread loads 4 values; k fastest for coalescing)
#pragma unroll
            for (int t = threadIdx.x; t < kTE * kKS; t += 256) {
                const int r = t / kKS, k = t % kKS;
                const int64_t e = e0 + r;
                hs[k][r] = e < E ? h[e * K + kk + k] : 0.f;
            }
#pragma unroll
            for (int t = threadIdx.x; t < OT * kKS; t += 256) {
                const int c = t / kKS, k = t % kKS;
                ws[k][c] = w3[(int64_t(i) * W + ot * OT + c) * K + kk + k];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kKS; ++k) {
                float a[4], b[TO];
#pragma unroll
                for (int r = 0; r < 4; ++r) a[r] = hs[k][ty * 4 + r];
#pragma unroll
                for (int c = 0; c < TO; ++c) b[c] = ws[k][tx * TO + c];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < TO; ++c) Pacc[r][c] = fmaf(a[r], b[c], Pacc[r][c]);
            }
            __syncthreads();
        }
    };

    if constexpr (MODE == MODE_FWD) {
        for (int ot = 0; ot < NOT; ++ot) {
            float macc[4][TO];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < TO; ++c) macc[r][c] = 0.f;
            for (int i = 0; i < W; ++i) {
                float Pacc[4][TO];
                p_tile(i, ot, Pacc);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float xv = ns[ty * 4 + r][i];
#pragma unroll
                    for (int c = 0; c < TO; ++c) macc[r][c] = fmaf(xv, Pacc[r][c], macc[r][c]);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int64_t e = e0 + ty * 4 + r;
                if (e < E) {
#pragma unroll
                    for (int c = 0; c < TO; ++c) out[e * W + ot * OT + tx * TO + c] = macc[r][c];
                }
            }
        }
    } else {
        for (int i = 0; i < W; ++i) {
            float part[4] = {0.f, 0.f, 0.f, 0.f};
            for (int ot = 0; ot < NOT; ++ot) {
                float Pacc[4][TO];
                p_tile(i, ot, Pacc);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < TO; ++c)
                        part[r] = fmaf(Pacc[r][c], ns[ty * 4 + r][ot * OT + tx * TO + c], part[r]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float v = part[r];
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                const int64_t e = e0 + ty * 4 + r;
                if (tx == 0 && e < E) out[e * W + i] = v;
            }
        }
    }
}

template <int MODE>
int simt_dispatch(const float* h, const float* w3, const int32_t* idx, const float* node,
                  float* out, int64_t E, int32_t K, int32_t W, cudaStream_t stream) {
    if (E == 0) return 0;
    if (E < 0 || h == nullptr || w3 == nullptr || idx == nullptr || node == nullptr || out == nullptr) {
        return set_error("gkn_contract_*_f32: bad size or NULL pointer argument");
    }
    if (K <= 0 || K % kKS != 0) return set_error("gkn_contract_*_f32: K must be a multiple of 16");
    const int64_t blocks = (E + kTE - 1) / kTE;
    if (blocks > 0x7fffffff) return set_error("gkn_contract_*_f32: too many edges");
    const dim3 grid{unsigned(blocks)};
    switch (W) {
        case 16: contract_simt_kernel<16, MODE><<<grid, 256, 0, stream>>>(h, w3, idx, node, out, E, K); break;
        case 32: contract_simt_kernel<32, MODE><<<grid, 256, 0, stream>>>(h, w3, idx, node, out, E, K); break;
        case 64: contract_simt_kernel<64, MODE><<<grid, 256, 0, stream>>>(h, w3, idx, node, out, E, K); break;
        case 128: contract_simt_kernel<128, MODE><<<grid, 256, 0, stream>>>(h, w3, idx, node, out, E, K); break;
        default: return set_error("gkn_contract_*_f32: width must be 16, 32, 64 or 128");
    }
    GKN_CHECK_LAUNCH("contract_simt launch");
    return 0;
}

}  // namespace
}  // namespace gkn

extern "C" int gkn_contract_fwd_f32(const float* h, const float* w3, const int32_t* src,
                                    const float* x, float* msg, int64_t E, int32_t K, int32_t W,
                                    void* stream) {
    return gkn::simt_dispatch<gkn::MODE_FWD>(h, w3, src, x, msg, E, K, W,
                                             static_cast<cudaStream_t>(stream));
}

extern "C" int gkn_contract_dx_f32(const float* h, const float* w3, const int32_t* dst,
                                   const float* g, float* dxj, int64_t E, int32_t K, int32_t W,
                                   void* stream) {
    return gkn::simt_dispatch<gkn::MODE_DX>(h, w3, dst, g, dxj, E, K, W,
                                            static_cast<cudaStream_t>(stream));
}
