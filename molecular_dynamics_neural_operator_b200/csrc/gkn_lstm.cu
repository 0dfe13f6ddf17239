// Node encoder's LSTM (reference graph_kernel.py:278-285: nn.LSTM(3, 3) driven window step by
// window step with the state carried over = ONE sequence of T = window * batch steps over
// R = num_residues independent rows).  Hidden size 3: there is no matrix product worth a tensor
// core here, only a long dependent chain per row, so the kernels keep one row per thread with all
// weights in registers and walk the sequence; cuDNN's generic path spends ~1.1 ms per training
// step on this (0.9 ms in one element-wise backward kernel), these two launches ~0.1 ms.
// fp32 throughout (expf / tanhf, no fast-math): the same numbers as the native torch kernels to
// rounding, in both precision modes.
#include "gkn_internal.h"

#include <stdio.h>

#include "../../include/gkn_b200.h"

namespace gkn {

namespace {

constexpr int kLstmThreads = 128;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// gate pre-activations a[j] = b_ih[j] + b_hh[j] + W_ih[j,:] x + W_hh[j,:] h, PyTorch row order
// (input, forget, cell, output), D rows each
template <int D>
struct LstmWeights {
    float wi[4 * D][D], wh[4 * D][D], b[4 * D];
    __device__ __forceinline__ void load(const float* w_ih, const float* w_hh, const float* b_ih,
                                         const float* b_hh) {
#pragma unroll
        for (int j = 0; j < 4 * D; ++j) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                wi[j][k] = __ldg(w_ih + j * D + k);
                wh[j][k] = __ldg(w_hh + j * D + k);
            }
            b[j] = __ldg(b_ih + j) + __ldg(b_hh + j);
        }
    }
};

// out[t,r,:] = h_t; gates[t,r,:] = activated (i, f, g, o); cells[t,r,:] = c_t (both optional,
// saved for the backward)
template <int D>
__global__ void __launch_bounds__(kLstmThreads)
lstm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w_ih,
                const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                const float* __restrict__ b_hh, float* __restrict__ out,
                float* __restrict__ gates, float* __restrict__ cells, int64_t T, int64_t R) {
    const int64_t r = int64_t(blockIdx.x) * kLstmThreads + threadIdx.x;
    if (r >= R) return;
    LstmWeights<D> w;
    w.load(w_ih, w_hh, b_ih, b_hh);
    float h[D], c[D], xt[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        h[k] = 0.f;
        c[k] = 0.f;
        xt[k] = T > 0 ? __ldg(x + r * D + k) : 0.f;
    }
    for (int64_t t = 0; t < T; ++t) {
        float xn[D];  // next step's input, in flight while this step's chain runs
#pragma unroll
        for (int k = 0; k < D; ++k) xn[k] = t + 1 < T ? __ldg(x + ((t + 1) * R + r) * D + k) : 0.f;
        float a[4 * D];
#pragma unroll
        for (int j = 0; j < 4 * D; ++j) {
            float s = w.b[j];
#pragma unroll
            for (int k = 0; k < D; ++k) s = fmaf(w.wi[j][k], xt[k], s);
#pragma unroll
            for (int k = 0; k < D; ++k) s = fmaf(w.wh[j][k], h[k], s);
            a[j] = s;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            a[k] = sigmoidf_(a[k]);
            a[D + k] = sigmoidf_(a[D + k]);
            a[2 * D + k] = tanhf(a[2 * D + k]);
            a[3 * D + k] = sigmoidf_(a[3 * D + k]);
            c[k] = fmaf(a[D + k], c[k], a[k] * a[2 * D + k]);
            h[k] = a[3 * D + k] * tanhf(c[k]);
        }
        const int64_t q = t * R + r;
#pragma unroll
        for (int k = 0; k < D; ++k) out[q * D + k] = h[k];
        if (gates != nullptr) {
#pragma unroll
            for (int j = 0; j < 4 * D; ++j) gates[q * 4 * D + j] = a[j];
#pragma unroll
            for (int k = 0; k < D; ++k) cells[q * D + k] = c[k];
        }
#pragma unroll
        for (int k = 0; k < D; ++k) xt[k] = xn[k];
    }
}

// Reverse walk.  Per thread: dh, dc carried; the weight / bias gradients of its row accumulate in
// registers and are folded per block at the end (fixed order) into dw_part[block, :] =
// (dW_ih [4D,D], dW_hh [4D,D], db [4D]); dx optional.
template <int D>
__global__ void __launch_bounds__(kLstmThreads)
lstm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w_ih,
                const float* __restrict__ w_hh, const float* __restrict__ out,
                const float* __restrict__ gates, const float* __restrict__ cells,
                const float* __restrict__ dout, float* __restrict__ dx,
                float* __restrict__ dw_part, int64_t T, int64_t R) {
    constexpr int G = 4 * D;
    constexpr int kAcc = 2 * G * D + G;
    const int64_t r = int64_t(blockIdx.x) * kLstmThreads + threadIdx.x;
    const bool live = r < R;
    float wi[G][D], wh[G][D];
#pragma unroll
    for (int j = 0; j < G; ++j) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            wi[j][k] = __ldg(w_ih + j * D + k);
            wh[j][k] = __ldg(w_hh + j * D + k);
        }
    }
    float acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.f;
    float dh[D], dc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) dh[k] = dc[k] = 0.f;

    if (live) {
        // the saved tensors of step t - 1 are fetched while step t's chain runs
        float a[G], c[D], cp[D], hp[D], xt[D], go[D];
        auto fetch = [&](int64_t t, float (&fa)[G], float (&fcp)[D], float (&fhp)[D],
                         float (&fx)[D], float (&fg)[D]) {
            const int64_t q = t * R + r;
#pragma unroll
            for (int j = 0; j < G; ++j) fa[j] = __ldg(gates + q * G + j);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                fcp[k] = t > 0 ? __ldg(cells + (q - R) * D + k) : 0.f;
                fhp[k] = t > 0 ? __ldg(out + (q - R) * D + k) : 0.f;
                fx[k] = __ldg(x + q * D + k);
                fg[k] = __ldg(dout + q * D + k);
            }
        };
        if (T > 0) {
            fetch(T - 1, a, cp, hp, xt, go);
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = __ldg(cells + ((T - 1) * R + r) * D + k);
        }
        for (int64_t t = T - 1; t >= 0; --t) {
            const int64_t q = t * R + r;
            float na[G], ncp[D], nhp[D], nx[D], ngo[D];
            if (t > 0) fetch(t - 1, na, ncp, nhp, nx, ngo);
            float da[G];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float i = a[k], f = a[D + k], g = a[2 * D + k], o = a[3 * D + k];
                const float tc = tanhf(c[k]);
                const float dht = go[k] + dh[k];
                const float dct = fmaf(dht * o, 1.f - tc * tc, dc[k]);
                da[k] = dct * g * i * (1.f - i);
                da[D + k] = dct * cp[k] * f * (1.f - f);
                da[2 * D + k] = dct * i * (1.f - g * g);
                da[3 * D + k] = dht * tc * o * (1.f - o);
                dc[k] = dct * f;
            }
#pragma unroll
            for (int j = 0; j < G; ++j) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    acc[j * D + k] = fmaf(da[j], xt[k], acc[j * D + k]);
                    acc[G * D + j * D + k] = fmaf(da[j], hp[k], acc[G * D + j * D + k]);
                }
                acc[2 * G * D + j] += da[j];
            }
#pragma unroll
            for (int k = 0; k < D; ++k) {
                float s = 0.f, sx = 0.f;
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    s = fmaf(wh[j][k], da[j], s);
                    sx = fmaf(wi[j][k], da[j], sx);
                }
                dh[k] = s;
                if (dx != nullptr) dx[q * D + k] = sx;
            }
            if (t > 0) {
#pragma unroll
                for (int j = 0; j < G; ++j) a[j] = na[j];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    c[k] = cp[k];  // c_{t-1} was this step's previous cell state
                    cp[k] = ncp[k];
                    hp[k] = nhp[k];
                    xt[k] = nx[k];
                    go[k] = ngo[k];
                }
            }
        }
    }

    // block fold: lanes by shuffle, warps through shared memory in warp order
    __shared__ float sm[kLstmThreads / 32][kAcc];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kAcc; ++i) {
        float v = acc[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kAcc; i += kLstmThreads) {
        float v = 0.f;
#pragma unroll
        for (int wq = 0; wq < kLstmThreads / 32; ++wq) v += sm[wq][i];
        dw_part[int64_t(blockIdx.x) * kAcc + i] = v;
    }
}

int check_lstm(const char* name, int64_t T, int64_t R, int32_t D) {
    static thread_local char buf[160];
    if (D != 3) {
        snprintf(buf, sizeof(buf), "%s: input and hidden size must be 3 (graph_kernel.py:278)", name);
        return set_error(buf);
    }
    if (T < 0 || R < 0 || R > (int64_t(1) << 31) / kLstmThreads) {
        snprintf(buf, sizeof(buf), "%s: bad sequence length / row count", name);
        return set_error(buf);
    }
    return 0;
}

}  // namespace

}  // namespace gkn

extern "C" int gkn_lstm_blocks(int64_t R) {
    return int((R + gkn::kLstmThreads - 1) / gkn::kLstmThreads);
}

extern "C" int gkn_lstm_fwd(const float* x, const float* w_ih, const float* w_hh,
                            const float* b_ih, const float* b_hh, float* out, float* gates,
                            float* cells, int64_t T, int64_t R, int32_t D, void* stream) {
    using namespace gkn;
    int rc = check_lstm("gkn_lstm_fwd", T, R, D);
    if (rc) return rc;
    if (T == 0 || R == 0) return 0;
    if (x == nullptr || w_ih == nullptr || w_hh == nullptr || b_ih == nullptr || b_hh == nullptr ||
        out == nullptr || (gates == nullptr) != (cells == nullptr)) {
        return set_error("gkn_lstm_fwd: NULL pointer argument");
    }
    lstm_fwd_kernel<3><<<gkn_lstm_blocks(R), kLstmThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        x, w_ih, w_hh, b_ih, b_hh, out, gates, cells, T, R);
    GKN_CHECK_LAUNCH("lstm_fwd launch");
    return 0;
}

extern "C" int gkn_lstm_bwd(const float* x, const float* w_ih, const float* w_hh, const float* out,
                            const float* gates, const float* cells, const float* dout, float* dx,
                            float* dw_part, int64_t T, int64_t R, int32_t D, void* stream) {
    using namespace gkn;
    int rc = check_lstm("gkn_lstm_bwd", T, R, D);
    if (rc) return rc;
    if (R == 0) return 0;
    if (x == nullptr || w_ih == nullptr || w_hh == nullptr || out == nullptr || gates == nullptr ||
        cells == nullptr || dout == nullptr || dw_part == nullptr) {
        return set_error("gkn_lstm_bwd: NULL pointer argument");
    }
    lstm_bwd_kernel<3><<<gkn_lstm_blocks(R), kLstmThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        x, w_ih, w_hh, out, gates, cells, dout, dx, dw_part, T, R);
    GKN_CHECK_LAUNCH("lstm_bwd launch");
    return 0;
}
