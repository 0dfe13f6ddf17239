// K2 / K2b / K6 — segmented reductions over destination-sorted edges, no global atomics.
//
// Replaces PyG's mean aggregation + NNConv_old.update (reference graph_kernel.py:198, :204-209:
// scatter-add + count + clamp + divide, then "+ x @ root + bias") and the autograd of the gather
// x_j = x[edge_index[0]].  Because gkn_graph_prepare orders edges by destination, every node owns
// a contiguous run of message rows: one warp walks the run with float4 lanes across channels and
// sub-warp groups across edges, then folds the groups with warp shuffles.  The summation order is
// fixed (strided by group, then a shuffle tree), so results are run-to-run deterministic.
#include <cuda_bf16.h>

#include "gkn_internal.h"

namespace gkn {

namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

__device__ __forceinline__ float4 f4shfl_xor(float4 a, int off) {
    return make_float4(__shfl_xor_sync(0xffffffffu, a.x, off), __shfl_xor_sync(0xffffffffu, a.y, off),
                       __shfl_xor_sync(0xffffffffu, a.z, off), __shfl_xor_sync(0xffffffffu, a.w, off));
}

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

// streaming (read-once) load that does not pollute L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// out[n] = act(inv_cnt[n] * sum_{p in seg(n)} (msg[p] + xb3[src_s[p]]) + xroot[n] + bias)
template <int W>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
scatter_mean_fwd_kernel(const float* __restrict__ msg, const int32_t* __restrict__ src_s,
                        const int32_t* __restrict__ dst_ptr, const float* __restrict__ inv_cnt,
                        const float* __restrict__ xb3, const float* __restrict__ xroot,
                        const float* __restrict__ bias, float* __restrict__ out, int64_t N,
                        int relu) {
    constexpr int LPR = W / 4;     // lanes per row (float4 each)
    constexpr int RPW = 32 / LPR;  // rows walked in parallel by one warp
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR;
    const int col = (lane % LPR) * 4;
    const int64_t warp0 = int64_t(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = int64_t(gridDim.x) * kWarpsPerBlock;
    for (int64_t n = warp0; n < N; n += nwarps) {
        const int beg = dst_ptr[n], end = dst_ptr[n + 1];
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        int p = beg + sub;
        for (; p + RPW < end; p += 2 * RPW) {
            float4 m0 = ldg4_stream(msg + int64_t(p) * W + col);
            float4 m1 = ldg4_stream(msg + int64_t(p + RPW) * W + col);
            if (xb3 != nullptr) {
                m0 = f4add(m0, ldg4(xb3 + int64_t(src_s[p]) * W + col));
                m1 = f4add(m1, ldg4(xb3 + int64_t(src_s[p + RPW]) * W + col));
            }
            acc0 = f4add(acc0, m0);
            acc1 = f4add(acc1, m1);
        }
        if (p < end) {
            float4 m0 = ldg4_stream(msg + int64_t(p) * W + col);
            if (xb3 != nullptr) m0 = f4add(m0, ldg4(xb3 + int64_t(src_s[p]) * W + col));
            acc0 = f4add(acc0, m0);
        }
        float4 acc = f4add(acc0, acc1);
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) acc = f4add(acc, f4shfl_xor(acc, off));
        if (sub == 0) {
            const float s = inv_cnt[n];
            float4 r = make_float4(acc.x * s, acc.y * s, acc.z * s, acc.w * s);
            if (xroot != nullptr) r = f4add(r, ldg4(xroot + n * W + col));
            if (bias != nullptr) r = f4add(r, ldg4(bias + col));
            if (relu) {
                r.x = fmaxf(r.x, 0.f);
                r.y = fmaxf(r.y, 0.f);
                r.z = fmaxf(r.z, 0.f);
                r.w = fmaxf(r.w, 0.f);
            }
            *reinterpret_cast<float4*>(out + n * W + col) = r;
        }
    }
}

// dx[n] = addend[n] + sum_{q in srcseg(n)} (dxj[p] + gb3[dst_s[p]]),  p = src_perm[q]
template <int W>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
scatter_src_sum_kernel(const float* __restrict__ dxj, const int32_t* __restrict__ src_ptr,
                       const int32_t* __restrict__ src_perm, const int32_t* __restrict__ dst_s,
                       const float* __restrict__ gb3, const float* __restrict__ addend,
                       float* __restrict__ dx, int64_t N) {
    constexpr int LPR = W / 4;
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR;
    const int col = (lane % LPR) * 4;
    const int64_t warp0 = int64_t(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = int64_t(gridDim.x) * kWarpsPerBlock;
    for (int64_t n = warp0; n < N; n += nwarps) {
        const int beg = src_ptr[n], end = src_ptr[n + 1];
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        int q = beg + sub;
        for (; q + RPW < end; q += 2 * RPW) {
            const int p0 = src_perm[q], p1 = src_perm[q + RPW];
            float4 m0 = ldg4_stream(dxj + int64_t(p0) * W + col);
            float4 m1 = ldg4_stream(dxj + int64_t(p1) * W + col);
            if (gb3 != nullptr) {
                m0 = f4add(m0, ldg4(gb3 + int64_t(dst_s[p0]) * W + col));
                m1 = f4add(m1, ldg4(gb3 + int64_t(dst_s[p1]) * W + col));
            }
            acc0 = f4add(acc0, m0);
            acc1 = f4add(acc1, m1);
        }
        if (q < end) {
            const int p0 = src_perm[q];
            float4 m0 = ldg4_stream(dxj + int64_t(p0) * W + col);
            if (gb3 != nullptr) m0 = f4add(m0, ldg4(gb3 + int64_t(dst_s[p0]) * W + col));
            acc0 = f4add(acc0, m0);
        }
        float4 acc = f4add(acc0, acc1);
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) acc = f4add(acc, f4shfl_xor(acc, off));
        if (sub == 0) {
            if (addend != nullptr) acc = f4add(acc, ldg4(addend + n * W + col));
            *reinterpret_cast<float4*>(dx + n * W + col) = acc;
        }
    }
}

// Fused node update (forward AND backward use the same kernel):
//   r[n]   = scale[n] * sum_{q in seg(n)} rows[p]                    p = perm ? perm[q] : q
//   agg[n] = scale[n] * sum_{q in seg(n)} gnode[gidx[p]]             (optionally written out)
//   out[n] = act( r[n] + agg[n] @ M1 + nin[n] @ M2 + bias )
// forward : rows=msg, gidx=src, gnode=x, scale=1/deg, M1=B3 (b3 viewed [w,w]), nin=x, M2=root
//           => NNConv_old.update with the layers.4 bias folded in (graph_kernel.py:200-209)
// backward: rows=dxj, perm=src_perm, gidx=dst, gnode=gn, M1=B3^T, nin=g, M2=root^T, no scale/bias
// The two w x w products are done per node inside the warp (operands from a per-warp smem row,
// matrices through the read-only path), so no separate tiny GEMM launches are needed.
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

// A warp handles NB nodes per round: their segment sums are parked in per-warp smem rows, then the
// two w x w products are done for all NB nodes together so every matrix element fetched (vector
// loads through the read-only path) is reused NB times from registers.
template <int W, int NB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
segment_matvec_kernel(const float* __restrict__ rows, const int32_t* __restrict__ perm,
                      const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ gidx,
                      const float* __restrict__ gnode, const float* __restrict__ scale,
                      const float* __restrict__ M1, const float* __restrict__ nin,
                      const float* __restrict__ M2, const float* __restrict__ bias,
                      float* __restrict__ out, float* __restrict__ agg_out, int64_t N, int relu) {
    constexpr int LPR = W / 4;
    constexpr int RPW = 32 / LPR;
    constexpr int CPL = W >= 32 ? W / 32 : 1;  // output columns per lane in the matvec phase
    __shared__ __align__(16) float s_r[kWarpsPerBlock][NB][W];
    __shared__ __align__(16) float s_g[kWarpsPerBlock][NB][W];
    __shared__ __align__(16) float s_n[kWarpsPerBlock][NB][W];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int sub = lane / LPR;
    const int col = (lane % LPR) * 4;
    const int64_t warp0 = int64_t(blockIdx.x) * kWarpsPerBlock + wib;
    const int64_t nwarps = int64_t(gridDim.x) * kWarpsPerBlock;
    const int64_t rounds = (N + NB - 1) / NB;
    for (int64_t rd = warp0; rd < rounds; rd += nwarps) {
        const int64_t n0 = rd * NB;
#pragma unroll 1
        for (int nb = 0; nb < NB; ++nb) {
            const int64_t n = n0 + nb;
            if (n >= N) break;
            const int beg = seg_ptr[n], end = seg_ptr[n + 1];
            float4 ar0 = make_float4(0.f, 0.f, 0.f, 0.f), ar1 = ar0, ag0 = ar0, ag1 = ar0;
            int q = beg + sub;
            // IF rows in flight per lane: indices first, then the 2*IF dependent row loads.  The
            // walk is a chain of three dependent loads per round and one warp per node is latency-
            // bound, so every round - the last, partial one included - issues all its loads at
            // once (slots past the end are predicated off and add zeros).
            constexpr int IF = W <= 64 ? 8 : 4;
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (; q < end; q += IF * RPW) {
                int pp[IF], ii[IF];
#pragma unroll
                for (int u = 0; u < IF; ++u) {
                    const int qq = q + u * RPW;
                    pp[u] = qq < end ? (perm ? perm[qq] : qq) : -1;
                }
#pragma unroll
                for (int u = 0; u < IF; ++u) ii[u] = pp[u] >= 0 ? gidx[pp[u]] : 0;
                float4 rr[IF], gg[IF];
#pragma unroll
                for (int u = 0; u < IF; ++u) {
                    const bool ok = pp[u] >= 0;
                    rr[u] = ok && rows ? ldg4_stream(rows + int64_t(pp[u]) * W + col) : zero4;
                    gg[u] = ok ? ldg4(gnode + int64_t(ii[u]) * W + col) : zero4;
                }
                // fixed pairwise order: even slots into the *0 accumulators, odd into the *1
#pragma unroll
                for (int u = 0; u < IF; u += 4) {
                    ar0 = f4add(ar0, f4add(rr[u], rr[u + 2]));
                    ar1 = f4add(ar1, f4add(rr[u + 1], rr[u + 3]));
                    ag0 = f4add(ag0, f4add(gg[u], gg[u + 2]));
                    ag1 = f4add(ag1, f4add(gg[u + 1], gg[u + 3]));
                }
            }
            float4 ar = f4add(ar0, ar1), ag = f4add(ag0, ag1);
#pragma unroll
            for (int off = LPR; off < 32; off <<= 1) {
                ar = f4add(ar, f4shfl_xor(ar, off));
                ag = f4add(ag, f4shfl_xor(ag, off));
            }
            if (sub == 0) {
                const float sc = scale ? scale[n] : 1.f;
                ar = make_float4(ar.x * sc, ar.y * sc, ar.z * sc, ar.w * sc);
                ag = make_float4(ag.x * sc, ag.y * sc, ag.z * sc, ag.w * sc);
                *reinterpret_cast<float4*>(&s_r[wib][nb][col]) = ar;
                *reinterpret_cast<float4*>(&s_g[wib][nb][col]) = ag;
                if (nin) *reinterpret_cast<float4*>(&s_n[wib][nb][col]) = ldg4(nin + n * W + col);
                if (agg_out) *reinterpret_cast<float4*>(agg_out + n * W + col) = ag;
            }
        }
        __syncwarp();
        const int c0 = lane * CPL;
        if (c0 < W) {
            const int nvalid = int(N - n0 < NB ? N - n0 : NB);
            float o[NB][CPL];
            float bv[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) bv[k] = bias ? __ldg(bias + c0 + k) : 0.f;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int k = 0; k < CPL; ++k) o[nb][k] = nb < nvalid ? s_r[wib][nb][c0 + k] + bv[k] : 0.f;
            // M1 == NULL (then M2 is NULL too): plain segment sums, out = r (+ bias), agg_out = agg
#pragma unroll 4
            for (int i = 0; M1 != nullptr && i < W; ++i) {
                float m1[CPL], m2[CPL];
                ldg_vec<CPL>(M1 + i * W + c0, m1);
                ldg_vec<CPL>(M2 + i * W + c0, m2);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const float a = s_g[wib][nb][i], b = s_n[wib][nb][i];
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        o[nb][k] = fmaf(a, m1[k], o[nb][k]);
                        o[nb][k] = fmaf(b, m2[k], o[nb][k]);
                    }
                }
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                if (nb < nvalid) {
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        out[(n0 + nb) * W + c0 + k] = relu ? fmaxf(o[nb][k], 0.f) : o[nb][k];
                    }
                }
            }
        }
        __syncwarp();
    }
}

__global__ void relu_grad_scale_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                       const float* __restrict__ inv_cnt, float* __restrict__ g,
                                       float* __restrict__ gn, int64_t total4, int W4, int relu) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total4; t += stride) {
        float4 v = ldg4(gy + t * 4);
        if (relu) {
            const float4 yy = ldg4(y + t * 4);
            v.x = yy.x > 0.f ? v.x : 0.f;
            v.y = yy.y > 0.f ? v.y : 0.f;
            v.z = yy.z > 0.f ? v.z : 0.f;
            v.w = yy.w > 0.f ? v.w : 0.f;
        }
        const float s = inv_cnt[t / W4];
        *reinterpret_cast<float4*>(g + t * 4) = v;
        *reinterpret_cast<float4*>(gn + t * 4) = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
    }
}

// S[e - e0, i*W + o] = sum_a xs[a, src_s[e], i] * gs[a, dst_s[e], o]; one block per edge (looped),
// each thread a 4 x 4 patch of the W x W outer-product sum.
template <int W, bool BF16>
__global__ void __launch_bounds__(256)
edge_outer_sum_kernel(const float* __restrict__ xs, const float* __restrict__ gs,
                      const int32_t* __restrict__ src_s, const int32_t* __restrict__ dst_s,
                      void* __restrict__ S, int64_t e0, int64_t e1, int64_t N, int A) {
    extern __shared__ float sm[];  // X[A][W] then G[A][W]
    float* X = sm;
    float* G = sm + A * W;
    constexpr int B4 = W / 4;
    for (int64_t e = e0 + blockIdx.x; e < e1; e += gridDim.x) {
        const int64_t s = src_s[e], d = dst_s[e];
        __syncthreads();
        for (int t = threadIdx.x; t < A * W; t += 256) {
            const int a = t / W, c = t % W;
            X[t] = xs[(int64_t(a) * N + s) * W + c];
            G[t] = gs[(int64_t(a) * N + d) * W + c];
        }
        __syncthreads();
        for (int item = threadIdx.x; item < B4 * B4; item += 256) {
            const int ib = (item / B4) * 4, ob = (item % B4) * 4;
            float acc[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
            for (int a = 0; a < A; ++a) {
                const float4 xv = *reinterpret_cast<const float4*>(X + a * W + ib);
                const float4 gv = *reinterpret_cast<const float4*>(G + a * W + ob);
                const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
                const float gc[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(xr[r], gc[c], acc[r][c]);
            }
            const int64_t base = (e - e0) * int64_t(W) * W;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int64_t off = base + int64_t(ib + r) * W + ob;
                if constexpr (BF16) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(acc[r][0], acc[r][1]);
                    __nv_bfloat162 hi = __floats2bfloat162_rn(acc[r][2], acc[r][3]);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(S) + off) = pk;
                } else {
                    *reinterpret_cast<float4*>(static_cast<float*>(S) + off) =
                        make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                }
            }
        }
    }
}

// bf16 variant of edge_outer_sum on the (legacy) warp-level tensor-core path: per edge
// S_e = X_e^T G_e is a 64x16 @ 16x64 product (the A <= 16 applications are the reduction dim),
// i.e. 4 x 8 mma.sync.m16n8k16 tiles.  One warp per edge; X_e / G_e rows are gathered, rounded to
// bf16 and parked transposed in a per-warp smem patch ([channel][application], 32 B rows) so every
// fragment register is one 32-bit load.  The n-tiles take G columns in the permuted order
// o(nt, n) = 16*(n/2) + 2*nt + n%2, which makes each lane own 16 consecutive output columns of a
// row: results leave as two 16-byte stores per row, full 128-byte lines per 4 lanes.
// W = 64 only (the widths 32 / 128 keep the CUDA-core kernel above).
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4],
                                               const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
        "{%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256, 3)
edge_outer_sum_mma64_kernel(const float* __restrict__ xs, const float* __restrict__ gs,
                            const int32_t* __restrict__ src_s, const int32_t* __restrict__ dst_s,
                            __nv_bfloat16* __restrict__ S, int64_t e0, int64_t e1, int64_t N,
                            int A) {
    constexpr int W = 64, KP = 16, LD = KP + 2;  // 36-byte rows (9 words, odd): the paired stores of
                                                 // rows 2l / 2l+1 spread over 16 banks, loads over 32
    __shared__ __align__(16) __nv_bfloat16 sx[8][W][LD];
    __shared__ __align__(16) __nv_bfloat16 sg[8][W][LD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __nv_bfloat16(*X)[LD] = sx[warp];
    __nv_bfloat16(*G)[LD] = sg[warp];
    // zero the padded application slots once (A..15); they are never written again
    for (int t = lane; t < W * KP; t += 32) {
        const int ch = t / KP, a = t % KP;
        if (a >= A) {
            X[ch][a] = __float2bfloat16(0.f);
            G[ch][a] = __float2bfloat16(0.f);
        }
    }
    const int r = lane >> 2, q = lane & 3;
    const int64_t wstride = int64_t(gridDim.x) * 8;
    for (int64_t e = e0 + int64_t(blockIdx.x) * 8 + warp; e < e1; e += wstride) {
        const int64_t s = src_s[e], d = dst_s[e];
        __syncwarp();
        // gather in groups of 4 applications (8 independent row loads in flight per lane); two
        // applications are packed per 32-bit smem store
        for (int a0 = 0; a0 < A; a0 += 4) {
            float2 xv[4], gv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int a = a0 + u < A ? a0 + u : A - 1;
                xv[u] = __ldg(reinterpret_cast<const float2*>(xs + (int64_t(a) * N + s) * W) + lane);
                gv[u] = __ldg(reinterpret_cast<const float2*>(gs + (int64_t(a) * N + d) * W) + lane);
            }
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
                const int a = a0 + u;
                if (a < A) {
                    const bool two = a + 1 < A;  // odd A: the pad slot stays zero
                    __nv_bfloat162 t;
                    t = __floats2bfloat162_rn(xv[u].x, two ? xv[u + 1].x : 0.f);
                    *reinterpret_cast<__nv_bfloat162*>(&X[2 * lane][a]) = t;
                    t = __floats2bfloat162_rn(xv[u].y, two ? xv[u + 1].y : 0.f);
                    *reinterpret_cast<__nv_bfloat162*>(&X[2 * lane + 1][a]) = t;
                    t = __floats2bfloat162_rn(gv[u].x, two ? gv[u + 1].x : 0.f);
                    *reinterpret_cast<__nv_bfloat162*>(&G[2 * lane][a]) = t;
                    t = __floats2bfloat162_rn(gv[u].y, two ? gv[u + 1].y : 0.f);
                    *reinterpret_cast<__nv_bfloat162*>(&G[2 * lane + 1][a]) = t;
                }
            }
        }
        __syncwarp();
        // B fragments for the 8 n-tiles (reused by all 4 m-tiles)
        uint32_t bf[8][2];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int o = 16 * (r >> 1) + 2 * nt + (r & 1);
            bf[nt][0] = *reinterpret_cast<const uint32_t*>(&G[o][2 * q]);
            bf[nt][1] = *reinterpret_cast<const uint32_t*>(&G[o][2 * q + 8]);
        }
        __nv_bfloat16* out = S + (e - e0) * int64_t(W) * W;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            uint32_t af[4];
            af[0] = *reinterpret_cast<const uint32_t*>(&X[16 * mt + r][2 * q]);
            af[1] = *reinterpret_cast<const uint32_t*>(&X[16 * mt + r + 8][2 * q]);
            af[2] = *reinterpret_cast<const uint32_t*>(&X[16 * mt + r][2 * q + 8]);
            af[3] = *reinterpret_cast<const uint32_t*>(&X[16 * mt + r + 8][2 * q + 8]);
            float acc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
                mma_bf16_16816(acc[nt], af, bf[nt]);
            }
            // lane (r, q) holds, for rows i = 16mt + r and i + 8, the 16 columns [16q, 16q+16):
            // tile nt contributes columns 16q + 2nt + {0,1}
            uint32_t lo[8], hi[8];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                __nv_bfloat162 t0 = __floats2bfloat162_rn(acc[nt][0], acc[nt][1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(acc[nt][2], acc[nt][3]);
                lo[nt] = *reinterpret_cast<uint32_t*>(&t0);
                hi[nt] = *reinterpret_cast<uint32_t*>(&t1);
            }
            uint4* p0 = reinterpret_cast<uint4*>(out + (16 * mt + r) * W + 16 * q);
            uint4* p1 = reinterpret_cast<uint4*>(out + (16 * mt + r + 8) * W + 16 * q);
            p0[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            p0[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            p1[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            p1[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        }
    }
}

__global__ void bias_act_cast_bf16_kernel(const float* __restrict__ in,
                                          const float* __restrict__ bias,
                                          __nv_bfloat16* __restrict__ out, int64_t total4,
                                          int64_t cols, int relu) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total4; t += stride) {
        float4 v = ldg4_stream(in + t * 4);
        if (bias != nullptr) {
            const float4 b = ldg4(bias + (t * 4) % cols);
            v = f4add(v, b);
        }
        if (relu) {
            v.x = fmaxf(v.x, 0.f);
            v.y = fmaxf(v.y, 0.f);
            v.z = fmaxf(v.z, 0.f);
            v.w = fmaxf(v.w, 0.f);
        }
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(out + t * 4) = pk;
    }
}

// dz_out[r,c] = dz[r,c] * (act[r,c] > 0)  (cast to OUT type), and per-block column partial sums
// partial[b, c] = sum over the block's rows; a second kernel folds the partials in a fixed order
// (deterministic bias gradient without atomics).  One thread owns 4 adjacent columns.
template <typename ACT>
__device__ __forceinline__ void load_act4(const ACT* p, float (&a)[4]) {
    if constexpr (sizeof(ACT) == 2) {
        const uint2 raw = *reinterpret_cast<const uint2*>(p);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
        a[0] = __low2float(lo);
        a[1] = __high2float(lo);
        a[2] = __low2float(hi);
        a[3] = __high2float(hi);
    } else {
        const float4 t = *reinterpret_cast<const float4*>(p);
        a[0] = t.x;
        a[1] = t.y;
        a[2] = t.z;
        a[3] = t.w;
    }
}

// `out` may alias `dz` (fp32 in place): every element is read and written by the same thread.
// NSIDE > 0 additionally accumulates partial[b, 1+j, c] = sum_r out[r,c] * side[r,j] (j < NSIDE):
// with side = edge_attr this is the layer-0 weight gradient, produced in the same pass.
template <typename DZ, typename ACT, typename OUT, int NSIDE>
__global__ void __launch_bounds__(256)
mask_cast_colsum_kernel(const DZ* dz, const ACT* __restrict__ act, OUT* out,
                        const float* __restrict__ side, float* __restrict__ partial, int64_t rows,
                        int cols) {
    constexpr int U = 4;  // rows in flight per thread
    constexpr int kMaxSideRows = 256;
    __shared__ float s_side[NSIDE > 0 ? kMaxSideRows * NSIDE : 1];
    const int64_t rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float* pbase = partial + int64_t(blockIdx.x) * (1 + NSIDE) * cols;
    const bool side_in_smem = NSIDE > 0 && (r1 - r0) <= kMaxSideRows;
    if (NSIDE > 0) {
        if (side_in_smem) {
            for (int t = threadIdx.x; t < int(r1 - r0) * NSIDE; t += 256) {
                s_side[t] = side[r0 * NSIDE + t];
            }
        }
        __syncthreads();
    }
    for (int c = threadIdx.x * 4; c < cols; c += 256 * 4) {
        float4 acc[1 + NSIDE];
#pragma unroll
        for (int j = 0; j <= NSIDE; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t rb = r0; rb < r1; rb += U) {
            float v4[U][4];
            float a[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t r = rb + u;
                if (r < r1) {
                    load_act4(dz + r * cols + c, v4[u]);
                    load_act4(act + r * cols + c, a[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t r = rb + u;
                if (r >= r1) break;
                float4 w = make_float4(v4[u][0], v4[u][1], v4[u][2], v4[u][3]);
                w.x = a[u][0] > 0.f ? w.x : 0.f;
                w.y = a[u][1] > 0.f ? w.y : 0.f;
                w.z = a[u][2] > 0.f ? w.z : 0.f;
                w.w = a[u][3] > 0.f ? w.w : 0.f;
                acc[0] = f4add(acc[0], w);
#pragma unroll
                for (int j = 0; j < NSIDE; ++j) {
                    const float sj = side_in_smem ? s_side[int(r - r0) * NSIDE + j]
                                                  : __ldg(side + r * NSIDE + j);
                    acc[1 + j].x = fmaf(w.x, sj, acc[1 + j].x);
                    acc[1 + j].y = fmaf(w.y, sj, acc[1 + j].y);
                    acc[1 + j].z = fmaf(w.z, sj, acc[1 + j].z);
                    acc[1 + j].w = fmaf(w.w, sj, acc[1 + j].w);
                }
                const int64_t off = r * cols + c;
                if (out == nullptr) {
                    // sums only (the masked values themselves are not needed downstream)
                } else if constexpr (sizeof(OUT) == 2) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(w.x, w.y);
                    __nv_bfloat162 hi = __floats2bfloat162_rn(w.z, w.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(out + off) = pk;
                } else {
                    *reinterpret_cast<float4*>(out + off) = w;
                }
            }
        }
#pragma unroll
        for (int j = 0; j <= NSIDE; ++j) {
            *reinterpret_cast<float4*>(pbase + int64_t(j) * cols + c) = acc[j];
        }
    }
}

// colsum[j, c] (+)= sum_b partial[b, j, c]; 8 columns (one 32-byte sector per partial row) x 32
// block-groups per CTA, fixed fold order.  (Round 1 folded 32 columns x 8 groups per CTA: 32 CTAs
// and 74 dependent loads per thread at K = 1024, 18 us per call, 32 calls per training step.)
__global__ void __launch_bounds__(256)
colsum_fold_kernel(const float* __restrict__ partial, float* __restrict__ colsum, int nblocks,
                   int planes, int cols, int accumulate) {
    __shared__ float sm[32][9];
    const int cx = threadIdx.x & 7, by = threadIdx.x >> 3;
    const int c = blockIdx.x * 8 + cx;
    const int j = blockIdx.y;
    float s = 0.f;
    if (c < cols) {
        for (int b = by; b < nblocks; b += 32) s += partial[(int64_t(b) * planes + j) * cols + c];
    }
    sm[by][cx] = s;
    __syncthreads();
    if (by == 0 && c < cols) {
        float t = accumulate ? colsum[int64_t(j) * cols + c] : 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += sm[k][cx];
        colsum[int64_t(j) * cols + c] = t;
    }
}

// h1[e,k] = relu(sum_j ea[e,j] * w1[k,j] + b1[k])   (DenseNet layers.0 + ReLU, fan-in KIN <= 8;
// reference graph_kernel.py:239-242), written straight in the dtype layer .2 consumes.
template <typename OUT, int KIN>
__global__ void __launch_bounds__(256)
edge_layer0_kernel(const float* __restrict__ ea, const float* __restrict__ w1,
                   const float* __restrict__ b1, OUT* __restrict__ out, int64_t E, int K) {
    const int64_t stride = gridDim.x;
    const bool vec_in = (reinterpret_cast<uintptr_t>(ea) & 15) == 0;
    for (int c = threadIdx.x * 4; c < K; c += 256 * 4) {
        float w[4][KIN], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            b[k] = b1[c + k];
#pragma unroll
            for (int j = 0; j < KIN; ++j) w[k][j] = w1[(c + k) * KIN + j];
        }
        // four edges per trip: their 4 * KIN inputs arrive as float4 loads (the same addresses
        // for every thread of the block: broadcast), 16 * KIN independent FMAs, four row stores
        constexpr int kEdges = 4;
        for (int64_t e0 = int64_t(blockIdx.x) * kEdges; e0 < E; e0 += stride * kEdges) {
            float x[kEdges][KIN];
            if (vec_in && e0 + kEdges <= E && (KIN * kEdges) % 4 == 0) {
                const float4* xp = reinterpret_cast<const float4*>(ea + e0 * KIN);
                float flat[kEdges * KIN];
#pragma unroll
                for (int q = 0; q < kEdges * KIN / 4; ++q) {
                    const float4 v = __ldg(xp + q);
                    flat[4 * q] = v.x; flat[4 * q + 1] = v.y; flat[4 * q + 2] = v.z; flat[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < kEdges; ++i)
#pragma unroll
                    for (int j = 0; j < KIN; ++j) x[i][j] = flat[i * KIN + j];
            } else {
#pragma unroll
                for (int i = 0; i < kEdges; ++i)
#pragma unroll
                    for (int j = 0; j < KIN; ++j) x[i][j] = e0 + i < E ? __ldg(ea + (e0 + i) * KIN + j) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < kEdges; ++i) {
                if (e0 + i >= E) break;
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float acc = b[k];
#pragma unroll
                    for (int j = 0; j < KIN; ++j) acc = fmaf(x[i][j], w[k][j], acc);
                    o[k] = fmaxf(acc, 0.f);
                }
                OUT* dst = out + (e0 + i) * K + c;
                if constexpr (sizeof(OUT) == 2) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]);
                    __nv_bfloat162 hi = __floats2bfloat162_rn(o[2], o[3]);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(dst) = pk;
                } else {
                    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }
}

int node_grid(int64_t N) {
    int sms = 148;
    device_sm_count(&sms);
    int64_t blocks = (N + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int64_t cap = int64_t(sms) * 8;  // 8 resident blocks of 8 warps per SM
    if (blocks > cap) blocks = cap;
    return int(blocks < 1 ? 1 : blocks);
}

int elem_grid(int64_t total, int threads) {
    int sms = 148;
    device_sm_count(&sms);
    int64_t blocks = (total + threads - 1) / threads;
    const int64_t cap = int64_t(sms) * 8;
    if (blocks > cap) blocks = cap;
    return int(blocks < 1 ? 1 : blocks);
}

// out[i, o] = sum_r x[r, i] * g[r, o] over many rows (the root-weight gradient X^T G of all
// applications at once, and db3 = agg^T G; reference graph_kernel.py:204-209 and :200-202 under
// autograd) and, optionally, gsum[o] = sum_r g[r, o] (the conv bias gradient).  A [W, W] result from
// tens of thousands of rows: cuBLAS runs the batched form on 24 CTAs (0.18 ms, three times per
// training step); here every CTA takes a contiguous row range (TILE x TILE accumulators per thread,
// rows staged through shared memory), writes a partial [W*W + W], and a second launch folds the
// partials in CTA order - deterministic, no atomics.
template <int W>
__global__ void __launch_bounds__(256)
xty_partial_kernel(const float* __restrict__ x, const float* __restrict__ g, int64_t rows,
                   int64_t rows_per_cta, float* __restrict__ part) {
    constexpr int T = W / 16;      // thread tile edge (16 x 16 threads cover W x W)
    constexpr int kChunk = 32;     // rows staged at a time
    __shared__ __align__(16) float xs[kChunk][W];
    __shared__ __align__(16) float gsm[kChunk][W];
    const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
    float acc[T][T];
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b) acc[a][b] = 0.f;
    float gcol = 0.f;  // thread t < W sums column t of g
    const int64_t r0 = int64_t(blockIdx.x) * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    for (int64_t rb = r0; rb < r1; rb += kChunk) {
        const int n = int(r1 - rb < kChunk ? r1 - rb : kChunk);
        __syncthreads();
        for (int i = threadIdx.x; i < kChunk * W / 4; i += 256) {
            const int r = i / (W / 4), c4 = i % (W / 4);
            float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vg = vx;
            if (r < n) {
                vx = __ldg(reinterpret_cast<const float4*>(x + (rb + r) * W) + c4);
                vg = __ldg(reinterpret_cast<const float4*>(g + (rb + r) * W) + c4);
            }
            *reinterpret_cast<float4*>(&xs[r][c4 * 4]) = vx;
            *reinterpret_cast<float4*>(&gsm[r][c4 * 4]) = vg;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < kChunk; ++r) {
            float xv[T], gv[T];
#pragma unroll
            for (int a = 0; a < T; ++a) xv[a] = xs[r][ti * T + a];
#pragma unroll
            for (int b = 0; b < T; ++b) gv[b] = gsm[r][tj * T + b];
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) acc[a][b] = fmaf(xv[a], gv[b], acc[a][b]);
        }
        if (threadIdx.x < W) {
#pragma unroll 8
            for (int r = 0; r < kChunk; ++r) gcol += gsm[r][threadIdx.x];
        }
    }
    float* dst = part + int64_t(blockIdx.x) * (W * W + W);
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b) dst[(ti * T + a) * W + tj * T + b] = acc[a][b];
    if (threadIdx.x < W) dst[W * W + threadIdx.x] = gcol;
}

// out[i] = sum_b part[b, i] (i < n_out), gsum[j] = sum_b part[b, n_out + j], in block order
__global__ void __launch_bounds__(256)
xty_fold_kernel(const float* __restrict__ part, int nblocks, int n_out, int n_sum,
                float* __restrict__ out, float* __restrict__ gsum) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_out + n_sum) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += part[int64_t(b) * (n_out + n_sum) + i];
    if (i < n_out) {
        out[i] = s;
    } else if (gsum != nullptr) {
        gsum[i - n_out] = s;
    }
}

}  // namespace
}  // namespace gkn

extern "C" int gkn_scatter_mean_fwd(const float* msg, const int32_t* src_s, const int32_t* dst_ptr,
                                    const float* inv_cnt, const float* xb3, const float* xroot,
                                    const float* bias, float* out, int64_t N, int64_t E, int32_t W,
                                    int32_t relu, void* stream) {
    using namespace gkn;
    (void)E;
    if (N == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = node_grid(N);
    const int threads = kWarpsPerBlock * 32;
    switch (W) {
        case 16: scatter_mean_fwd_kernel<16><<<grid, threads, 0, st>>>(msg, src_s, dst_ptr, inv_cnt, xb3, xroot, bias, out, N, relu); break;
        case 32: scatter_mean_fwd_kernel<32><<<grid, threads, 0, st>>>(msg, src_s, dst_ptr, inv_cnt, xb3, xroot, bias, out, N, relu); break;
        case 64: scatter_mean_fwd_kernel<64><<<grid, threads, 0, st>>>(msg, src_s, dst_ptr, inv_cnt, xb3, xroot, bias, out, N, relu); break;
        case 128: scatter_mean_fwd_kernel<128><<<grid, threads, 0, st>>>(msg, src_s, dst_ptr, inv_cnt, xb3, xroot, bias, out, N, relu); break;
        default: return set_error("gkn_scatter_mean_fwd: width must be 16, 32, 64 or 128");
    }
    GKN_CHECK_LAUNCH("scatter_mean_fwd launch");
    return 0;
}

extern "C" int gkn_scatter_src_sum(const float* dxj, const int32_t* src_ptr,
                                   const int32_t* src_perm, const int32_t* dst_s, const float* gb3,
                                   const float* addend, float* dx, int64_t N, int64_t E, int32_t W,
                                   void* stream) {
    using namespace gkn;
    (void)E;
    if (N == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = node_grid(N);
    const int threads = kWarpsPerBlock * 32;
    switch (W) {
        case 16: scatter_src_sum_kernel<16><<<grid, threads, 0, st>>>(dxj, src_ptr, src_perm, dst_s, gb3, addend, dx, N); break;
        case 32: scatter_src_sum_kernel<32><<<grid, threads, 0, st>>>(dxj, src_ptr, src_perm, dst_s, gb3, addend, dx, N); break;
        case 64: scatter_src_sum_kernel<64><<<grid, threads, 0, st>>>(dxj, src_ptr, src_perm, dst_s, gb3, addend, dx, N); break;
        case 128: scatter_src_sum_kernel<128><<<grid, threads, 0, st>>>(dxj, src_ptr, src_perm, dst_s, gb3, addend, dx, N); break;
        default: return set_error("gkn_scatter_src_sum: width must be 16, 32, 64 or 128");
    }
    GKN_CHECK_LAUNCH("scatter_src_sum launch");
    return 0;
}

extern "C" int64_t gkn_xty_workspace_bytes(int32_t W) {
    int sms = 148;
    gkn::device_sm_count(&sms);
    return int64_t(sms) * 2 * (int64_t(W) * W + W) * int64_t(sizeof(float));
}

extern "C" int gkn_xty_f32(const float* x, const float* g, int64_t rows, int32_t W, float* out,
                           float* gsum, void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace gkn;
    if (W != 16 && W != 32 && W != 64 && W != 128) {
        return set_error("gkn_xty_f32: width must be 16, 32, 64 or 128");
    }
    if (rows < 0 || out == nullptr || workspace == nullptr ||
        (rows > 0 && (x == nullptr || g == nullptr))) {
        return set_error("gkn_xty_f32: NULL pointer argument");
    }
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    if (workspace_bytes < int64_t(sms) * 2 * (int64_t(W) * W + W) * int64_t(sizeof(float))) {
        return set_error("gkn_xty_f32: workspace too small");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // row ranges in multiples of 32 rows, at most two CTAs per SM
    const int64_t chunks = (rows + 31) / 32;
    int64_t blocks = chunks < int64_t(sms) * 2 ? chunks : int64_t(sms) * 2;
    if (blocks < 1) blocks = 1;
    const int64_t rows_per = ((chunks + blocks - 1) / blocks) * 32;
    blocks = rows > 0 ? (rows + rows_per - 1) / rows_per : 1;
    float* part = static_cast<float*>(workspace);
    const dim3 grid{unsigned(blocks)};
    switch (W) {
        case 16: xty_partial_kernel<16><<<grid, 256, 0, st>>>(x, g, rows, rows_per, part); break;
        case 32: xty_partial_kernel<32><<<grid, 256, 0, st>>>(x, g, rows, rows_per, part); break;
        case 64: xty_partial_kernel<64><<<grid, 256, 0, st>>>(x, g, rows, rows_per, part); break;
        default: xty_partial_kernel<128><<<grid, 256, 0, st>>>(x, g, rows, rows_per, part); break;
    }
    GKN_CHECK_LAUNCH("xty_partial launch");
    const int n_out = W * W;
    xty_fold_kernel<<<(n_out + W + 255) / 256, 256, 0, st>>>(part, int(blocks), n_out, W, out, gsum);
    GKN_CHECK_LAUNCH("xty_fold launch");
    return 0;
}

extern "C" int gkn_relu_grad_scale(const float* gy, const float* y, const float* inv_cnt, float* g,
                                   float* gn, int64_t N, int32_t W, int32_t relu, void* stream) {
    using namespace gkn;
    if (N == 0) return 0;
    if (W % 4 != 0) return set_error("gkn_relu_grad_scale: width must be a multiple of 4");
    if (relu && y == nullptr) return set_error("gkn_relu_grad_scale: relu needs the layer output");
    const int64_t total4 = N * W / 4;
    relu_grad_scale_kernel<<<elem_grid(total4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        gy, y, inv_cnt, g, gn, total4, W / 4, relu);
    GKN_CHECK_LAUNCH("relu_grad_scale launch");
    return 0;
}

extern "C" int gkn_edge_outer_sum(const float* xs, const float* gs, const int32_t* src_s,
                                  const int32_t* dst_s, void* S, int32_t out_is_bf16, int64_t e0,
                                  int64_t e1, int64_t N, int32_t A, int32_t W, void* stream) {
    using namespace gkn;
    if (e1 <= e0) return 0;
    if (A < 1 || A > 64) return set_error("gkn_edge_outer_sum: A must be in [1, 64]");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int sms = 148;
    device_sm_count(&sms);
    int64_t blocks = e1 - e0;
    if (blocks > int64_t(sms) * 8) blocks = int64_t(sms) * 8;
    if (out_is_bf16 && W == 64 && A <= 16) {
        // tensor-core path: one warp per edge
        int64_t b8 = (e1 - e0 + 7) / 8;
        if (b8 > int64_t(sms) * 8) b8 = int64_t(sms) * 8;
        edge_outer_sum_mma64_kernel<<<int(b8), 256, 0, st>>>(
            xs, gs, src_s, dst_s, static_cast<__nv_bfloat16*>(S), e0, e1, N, A);
        GKN_CHECK_LAUNCH("edge_outer_sum_mma launch");
        return 0;
    }
    const size_t smem = size_t(2) * A * W * sizeof(float);
    const int grid = int(blocks);
#define GKN_EOS(WW)                                                                              \
    if (out_is_bf16)                                                                             \
        edge_outer_sum_kernel<WW, true><<<grid, 256, smem, st>>>(xs, gs, src_s, dst_s, S, e0, e1, N, A); \
    else                                                                                         \
        edge_outer_sum_kernel<WW, false><<<grid, 256, smem, st>>>(xs, gs, src_s, dst_s, S, e0, e1, N, A);
    switch (W) {
        case 16: GKN_EOS(16) break;
        case 32: GKN_EOS(32) break;
        case 64: GKN_EOS(64) break;
        case 128: {
            // 2 * 64 * 128 * 4 = 64 KB > 48 KB default for large A
            if (smem > 48 * 1024) {
                cudaFuncSetAttribute(edge_outer_sum_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
                cudaFuncSetAttribute(edge_outer_sum_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            }
            GKN_EOS(128)
            break;
        }
        default: return set_error("gkn_edge_outer_sum: width must be 16, 32, 64 or 128");
    }
#undef GKN_EOS
    GKN_CHECK_LAUNCH("edge_outer_sum launch");
    return 0;
}

extern "C" int gkn_bias_act_cast_bf16(const float* in, const float* bias, void* out_bf16,
                                      int64_t rows, int64_t cols, int32_t relu, void* stream) {
    using namespace gkn;
    if (rows == 0 || cols == 0) return 0;
    if (cols % 4 != 0) return set_error("gkn_bias_act_cast_bf16: cols must be a multiple of 4");
    const int64_t total4 = rows * cols / 4;
    bias_act_cast_bf16_kernel<<<elem_grid(total4, 256), 256, 0,
                                static_cast<cudaStream_t>(stream)>>>(
        in, bias, static_cast<__nv_bfloat16*>(out_bf16), total4, cols, relu);
    GKN_CHECK_LAUNCH("bias_act_cast_bf16 launch");
    return 0;
}

extern "C" int64_t gkn_mask_cast_colsum_workspace_bytes(int64_t cols, int32_t nside) {
    int sms = 148;
    gkn::device_sm_count(&sms);
    return int64_t(sms) * 4 * (1 + nside) * cols * int64_t(sizeof(float));
}

namespace gkn {
namespace {
template <typename DZ, typename ACT, typename OUT>
void launch_mask(int nside, int grid, cudaStream_t st, const DZ* dz, const ACT* act, OUT* out,
                 const float* side, float* partial, int64_t rows, int cols) {
    if (nside == 0) {
        mask_cast_colsum_kernel<DZ, ACT, OUT, 0><<<grid, 256, 0, st>>>(dz, act, out, side, partial, rows, cols);
    } else {
        mask_cast_colsum_kernel<DZ, ACT, OUT, 6><<<grid, 256, 0, st>>>(dz, act, out, side, partial, rows, cols);
    }
}
}  // namespace
}  // namespace gkn

extern "C" int gkn_mask_cast_colsum(const void* dz, int32_t dz_is_bf16, const void* act,
                                    int32_t act_is_bf16, void* out, int32_t out_is_bf16,
                                    const float* side, int32_t nside, float* colsum,
                                    int32_t accumulate, int64_t rows, int64_t cols,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace gkn;
    if (rows == 0 || cols == 0) return 0;
    if (cols % 4 != 0 || cols > (1 << 20)) {
        return set_error("gkn_mask_cast_colsum: cols must be a multiple of 4");
    }
    if (nside != 0 && nside != 6) {
        return set_error("gkn_mask_cast_colsum: side width must be 0 or 6 (edge_features)");
    }
    if (nside != 0 && side == nullptr) return set_error("gkn_mask_cast_colsum: side is NULL");
    if (dz_is_bf16 && !(act_is_bf16 && out_is_bf16)) {
        return set_error("gkn_mask_cast_colsum: bf16 dz requires bf16 act and bf16 out");
    }
    int sms = 148;
    device_sm_count(&sms);
    int64_t nblocks = int64_t(sms) * 4;
    if (nblocks > rows) nblocks = rows;
    if (workspace_bytes < nblocks * (1 + nside) * cols * int64_t(sizeof(float))) {
        return set_error("gkn_mask_cast_colsum: workspace too small");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* partial = static_cast<float*>(workspace);
    const int grid = int(nblocks);
    const int c = int(cols);
    using bf = __nv_bfloat16;
    const float* dzf = static_cast<const float*>(dz);
    if (dz_is_bf16) {
        launch_mask<bf, bf, bf>(nside, grid, st, static_cast<const bf*>(dz), static_cast<const bf*>(act), static_cast<bf*>(out), side, partial, rows, c);
    } else if (act_is_bf16 && out_is_bf16) {
        launch_mask<float, bf, bf>(nside, grid, st, dzf, static_cast<const bf*>(act), static_cast<bf*>(out), side, partial, rows, c);
    } else if (act_is_bf16) {
        launch_mask<float, bf, float>(nside, grid, st, dzf, static_cast<const bf*>(act), static_cast<float*>(out), side, partial, rows, c);
    } else if (out_is_bf16) {
        launch_mask<float, float, bf>(nside, grid, st, dzf, static_cast<const float*>(act), static_cast<bf*>(out), side, partial, rows, c);
    } else {
        launch_mask<float, float, float>(nside, grid, st, dzf, static_cast<const float*>(act), static_cast<float*>(out), side, partial, rows, c);
    }
    GKN_CHECK_LAUNCH("mask_cast_colsum launch");
    const dim3 fgrid((c + 7) / 8, 1 + nside);
    colsum_fold_kernel<<<fgrid, 256, 0, st>>>(partial, colsum, grid, 1 + nside, c, accumulate);
    GKN_CHECK_LAUNCH("colsum_fold launch");
    return 0;
}

extern "C" int gkn_edge_layer0(const float* ea, const float* w1, const float* b1, void* out,
                               int32_t out_is_bf16, int64_t E, int32_t K, int32_t kin,
                               void* stream) {
    using namespace gkn;
    if (E == 0) return 0;
    if (kin != 6) return set_error("gkn_edge_layer0: edge feature width must be 6");
    if (K % 4 != 0) return set_error("gkn_edge_layer0: K must be a multiple of 4");
    int sms = 148;
    device_sm_count(&sms);
    int64_t blocks = int64_t(sms) * 8;
    if (blocks > E) blocks = E;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (out_is_bf16) {
        edge_layer0_kernel<__nv_bfloat16, 6><<<int(blocks), 256, 0, st>>>(
            ea, w1, b1, static_cast<__nv_bfloat16*>(out), E, K);
    } else {
        edge_layer0_kernel<float, 6><<<int(blocks), 256, 0, st>>>(ea, w1, b1, static_cast<float*>(out), E, K);
    }
    GKN_CHECK_LAUNCH("edge_layer0 launch");
    return 0;
}

extern "C" int gkn_segment_matvec(const float* rows, const int32_t* perm, const int32_t* seg_ptr,
                                  const int32_t* gidx, const float* gnode, const float* scale,
                                  const float* M1, const float* nin, const float* M2,
                                  const float* bias, float* out, float* agg_out, int64_t N,
                                  int32_t W, int32_t relu, void* stream) {
    using namespace gkn;
    if (N == 0) return 0;
    // rows may be NULL: r[n] = 0 (the node-first backward has no per-edge rows to sum);
    // gidx may be NULL for a graph without edges (never dereferenced: all segments empty)
    // M1 and M2 may both be NULL: no matrix products (plain segment sums; nin is then unused)
    if (N < 0 || seg_ptr == nullptr || gnode == nullptr || out == nullptr ||
        (M1 == nullptr) != (M2 == nullptr) || (M1 != nullptr && nin == nullptr)) {
        return set_error("gkn_segment_matvec: bad size or NULL pointer argument");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = node_grid(N);
    const int threads = kWarpsPerBlock * 32;
    // many nodes: 4 (2 at w=128) nodes per warp round so matrix loads are reused; few nodes:
    // one node per warp keeps every SM busy with the (latency-bound) segment walks
    const bool batch_nodes = N >= int64_t(grid) * kWarpsPerBlock * 8;
#define GKN_SMV(WW, NBB)                                                                         \
    segment_matvec_kernel<WW, NBB><<<grid, threads, 0, st>>>(rows, perm, seg_ptr, gidx, gnode,   \
                                                             scale, M1, nin, M2, bias, out,      \
                                                             agg_out, N, relu)
    switch (W) {
        case 16: if (batch_nodes) GKN_SMV(16, 4); else GKN_SMV(16, 1); break;
        case 32: if (batch_nodes) GKN_SMV(32, 4); else GKN_SMV(32, 1); break;
        case 64: if (batch_nodes) GKN_SMV(64, 4); else GKN_SMV(64, 1); break;
        case 128: if (batch_nodes) GKN_SMV(128, 4); else GKN_SMV(128, 1); break;
        default: return set_error("gkn_segment_matvec: width must be 16, 32, 64 or 128");
    }
#undef GKN_SMV
    GKN_CHECK_LAUNCH("segment_matvec launch");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// edge_attr[e] = [pos[row e], pos[col e]]   (dataset feeder: reference dataset.py:194-201 builds this
// with a Python loop per edge).  One thread per edge: two gathered 12-byte reads from the (small,
// cache-resident) coordinate table, three 8-byte stores into the 24-byte record.
// ------------------------------------------------------------------------------------------------
namespace gkn {
namespace {
__global__ void __launch_bounds__(256)
edge_attr_gather_kernel(const float* __restrict__ pos, const int64_t* __restrict__ edge_index,
                        int64_t E, int64_t M, float* __restrict__ edge_attr,
                        int32_t* __restrict__ status) {
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < E;
         e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = __ldg(edge_index + e), j = __ldg(edge_index + E + e);
        float2* out = reinterpret_cast<float2*>(edge_attr + e * 6);
        if (i < 0 || i >= M || j < 0 || j >= M) {
            if (status != nullptr) atomicOr(status, 1);  // out-of-range node id: record, write zeros
            out[0] = out[1] = out[2] = make_float2(0.f, 0.f);
            continue;
        }
        const float ax = __ldg(pos + 3 * i), ay = __ldg(pos + 3 * i + 1), az = __ldg(pos + 3 * i + 2);
        const float bx = __ldg(pos + 3 * j), by = __ldg(pos + 3 * j + 1), bz = __ldg(pos + 3 * j + 2);
        out[0] = make_float2(ax, ay);
        out[1] = make_float2(az, bx);
        out[2] = make_float2(by, bz);
    }
}
}  // namespace
}  // namespace gkn

extern "C" int gkn_edge_attr_gather(const float* pos, const int64_t* edge_index, int64_t E, int64_t M,
                                    float* edge_attr, int32_t* status, void* stream) {
    using namespace gkn;
    if (E == 0) return 0;
    if (E < 0 || M <= 0 || pos == nullptr || edge_index == nullptr || edge_attr == nullptr) {
        return set_error("gkn_edge_attr_gather: bad size or NULL pointer argument");
    }
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    int64_t blocks = (E + 255) / 256;
    if (blocks > int64_t(sms) * 16) blocks = int64_t(sms) * 16;
    edge_attr_gather_kernel<<<dim3(unsigned(blocks)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pos, edge_index, E, M, edge_attr, status);
    GKN_CHECK_LAUNCH("edge_attr_gather launch");
    return 0;
}
