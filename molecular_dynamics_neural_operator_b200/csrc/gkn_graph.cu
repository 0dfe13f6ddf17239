// K0 — radius-graph build, and the destination-sorted CSR metadata every other kernel consumes.
//
// Radius graph: replaces construct_pairdata (reference graph_kernel.py:362-393), i.e.
//     cm = scipy.spatial.distance_matrix(p, p) < threshold;  coo_matrix(cm)  ->  (row, col)
//     edge_attr[e] = concat(p[row_e], p[col_e])              (a per-edge Python loop, :372-379)
// The reference is O(N^2) in float64.  Here atoms are binned into a uniform cell list (cell edge
// just above the cutoff), re-ordered by cell into a float4 array so candidate scans are
// coalesced, and each warp tests one atom against its 27 neighbouring cells.  The predicate is
// evaluated exactly as SciPy does after promoting float32 coordinates to float64:
//     sqrt(((dx*dx + dy*dy) + dz*dz)) < threshold        (no FMA contraction)
// so the edge set - and, after the per-row sort, the (row, col) order with self loops included -
// is bit-identical to the reference's.  Two passes (count -> scan -> fill) size the output.
//
// Graph prepare: replaces the bookkeeping hidden in PyG's propagate (gather by edge_index[0],
// reduce at edge_index[1]; reference graph_kernel.py:198).  Atomics appear only in the counting
// sort's histogram/placement; every run is then put into a canonical (ascending) order so
// results never depend on atomic arrival order.
#include <cmath>
#include <limits>

#include "gkn_internal.h"

namespace gkn {

namespace {

constexpr int kScanThreads = 1024;

// ------------------------------------------------------------------------------------------------
// single-block exclusive scan (metadata sizes only; not on the per-application hot loop)
// ------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kScanThreads)
exclusive_scan_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, int64_t n,
                      int64_t in_stride, int64_t out_stride) {
    // one block per segment (blockIdx.x), segment length n; writes n+1 values; 4 items / thread
    constexpr int IT = 4;
    in += int64_t(blockIdx.x) * in_stride;
    out += int64_t(blockIdx.x) * out_stride;
    __shared__ TOut warp_tot[32];
    __shared__ TOut carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += kScanThreads * IT) {
        const int64_t i0 = base + int64_t(threadIdx.x) * IT;
        TOut v[IT], tsum = 0;
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            v[k] = i0 + k < n ? TOut(in[i0 + k]) : TOut(0);
            tsum += v[k];
        }
        TOut incl = tsum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const TOut t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            TOut w = warp_tot[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const TOut t = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w += t;
            }
            warp_tot[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const TOut carry = carry_s;
        TOut run = carry + (warp == 0 ? TOut(0) : warp_tot[warp - 1]) + incl - tsum;
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            if (i0 + k < n) out[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

// ------------------------------------------------------------------------------------------------
// graph prepare
// ------------------------------------------------------------------------------------------------
__global__ void prep_count_kernel(const int64_t* __restrict__ ei, int64_t N, int64_t E,
                                  int32_t* __restrict__ cnt_dst, int32_t* __restrict__ cnt_src,
                                  int32_t* __restrict__ status) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < E; e += stride) {
        const int64_t s = ei[e], d = ei[E + e];
        if (s < 0 || s >= N || d < 0 || d >= N) {
            atomicOr(status, 1);
            continue;
        }
        atomicAdd(cnt_dst + d, 1);
        atomicAdd(cnt_src + s, 1);
    }
}

__global__ void prep_inv_cnt_kernel(const int32_t* __restrict__ dst_ptr, float* __restrict__ inv,
                                    int64_t N) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t n = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; n < N; n += stride) {
        const int c = dst_ptr[n + 1] - dst_ptr[n];
        inv[n] = 1.0f / float(c < 1 ? 1 : c);
    }
}

__global__ void prep_place_dst_kernel(const int64_t* __restrict__ ei, int64_t N, int64_t E,
                                      const int32_t* __restrict__ dst_ptr,
                                      int32_t* __restrict__ cursor, int32_t* __restrict__ tmp) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < E; e += stride) {
        const int64_t s = ei[e], d = ei[E + e];
        if (s < 0 || s >= N || d < 0 || d >= N) continue;
        const int pos = atomicAdd(cursor + d, 1);
        tmp[dst_ptr[d] + pos] = int32_t(e);
    }
}

// one warp per node: put the run tmp[beg,end) into ascending order by rank counting
// (values are distinct), emitting `vals[beg + rank] = v` and optional side outputs.
__global__ void prep_canon_dst_kernel(const int64_t* __restrict__ ei, int64_t N, int64_t E,
                                      const int32_t* __restrict__ dst_ptr,
                                      const int32_t* __restrict__ tmp, int32_t* __restrict__ perm,
                                      int32_t* __restrict__ src_s, int32_t* __restrict__ dst_s) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t n = warp0; n < N; n += nwarps) {
        const int beg = dst_ptr[n], end = dst_ptr[n + 1];
        for (int j = beg + lane; j < end; j += 32) {
            const int32_t v = tmp[j];
            int rank = 0;
            for (int k = beg; k < end; ++k) rank += (tmp[k] < v) ? 1 : 0;
            perm[beg + rank] = v;
            src_s[beg + rank] = int32_t(ei[v]);
            dst_s[beg + rank] = int32_t(n);
        }
    }
}

__global__ void prep_place_src_kernel(const int32_t* __restrict__ src_s, int64_t E,
                                      const int32_t* __restrict__ src_ptr,
                                      int32_t* __restrict__ cursor, int32_t* __restrict__ tmp) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < E; p += stride) {
        const int s = src_s[p];
        const int pos = atomicAdd(cursor + s, 1);
        tmp[src_ptr[s] + pos] = int32_t(p);
    }
}

__global__ void prep_canon_src_kernel(int64_t N, const int32_t* __restrict__ src_ptr,
                                      const int32_t* __restrict__ tmp,
                                      int32_t* __restrict__ src_perm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t n = warp0; n < N; n += nwarps) {
        const int beg = src_ptr[n], end = src_ptr[n + 1];
        for (int j = beg + lane; j < end; j += 32) {
            const int32_t v = tmp[j];
            int rank = 0;
            for (int k = beg; k < end; ++k) rank += (tmp[k] < v) ? 1 : 0;
            src_perm[beg + rank] = v;
        }
    }
}

// Symmetric, (row, col)-sorted graphs (what the radius-graph kernel emits) need no sorting at all:
// the destination-sorted order has the same CSR structure, position p <-> (dst = row[p],
// src = col[p]), and the only non-trivial array is the position of every edge's reverse edge.
__global__ void plan_from_csr_kernel(const int64_t* __restrict__ ei, const int64_t* __restrict__ row_ptr,
                                     int64_t N, int64_t E, int32_t* __restrict__ rev,
                                     int32_t* __restrict__ src_s, int32_t* __restrict__ dst_s,
                                     int32_t* __restrict__ ptr32, float* __restrict__ inv_cnt,
                                     int32_t* __restrict__ status) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    const int64_t t0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (int64_t n = t0; n <= N; n += stride) {
        ptr32[n] = int32_t(row_ptr[n]);
        if (n < N) {
            const int64_t c = row_ptr[n + 1] - row_ptr[n];
            inv_cnt[n] = 1.0f / float(c < 1 ? 1 : c);
        }
    }
    for (int64_t e = t0; e < E; e += stride) {
        const int64_t r = ei[e], c = ei[E + e];
        src_s[e] = int32_t(c);
        dst_s[e] = int32_t(r);
        int64_t lo = row_ptr[c], hi = row_ptr[c + 1];  // find r among the columns of row c
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (ei[E + mid] < r) lo = mid + 1; else hi = mid;
        }
        if (lo < row_ptr[c + 1] && ei[E + lo] == r) {
            rev[e] = int32_t(lo);
        } else {
            rev[e] = 0;
            atomicOr(status, 1);  // not symmetric / not sorted
        }
    }
}

// ------------------------------------------------------------------------------------------------
// radius graph
// ------------------------------------------------------------------------------------------------
struct FrameMeta {
    double minx, miny, minz;
    double inv_cell;
    int32_t nx, ny, nz, ncell;
};

__host__ __device__ inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct RadiusWs {
    FrameMeta* meta;       // [F]
    int32_t* cell_count;   // [F, C+1]  (histogram, then reused as placement cursor)
    int32_t* cell_start;   // [F, C+1]
    int32_t* atom_cell;    // [F*N]
    float4* sorted_pos;    // [F*N]  xyz + original atom index (bit-cast) in cell order
    int64_t* block_sums;   // [nblk]     row-pointer scan: per-block sums
    int64_t* block_offs;   // [nblk + 1] and their exclusive scan
    float4* nbr;           // large frames: [F*N, kListCap] sorted neighbour lists of pass 1
                           // (x, y, z, atom index): pass 2 needs no gather
    int32_t* list_ok;      // large frames: [F*N] 1 = row has its list, 0 = pass 2 rebuilds it
    int32_t* crowded_count;  // large frames: cells left to rg_crowded_kernel
    int2* crowded_cells;     // [F * C] (frame, cell)
    uint32_t* hit_mask;    // small frames: [F*N, ceil(N/32)] hit bit matrix
    int64_t* cta_edges;    // small frames: edge total of every pass-1 CTA
    int64_t nblk;
    int64_t C;             // cells reserved per frame
    int64_t bytes;
};

inline RadiusWs carve_radius_ws(void* base, int64_t F, int64_t N) {
    RadiusWs w;
    w.C = N + 64;
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        uint8_t* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    w.meta = reinterpret_cast<FrameMeta*>(take(F * int64_t(sizeof(FrameMeta))));
    w.cell_count = reinterpret_cast<int32_t*>(take(F * (w.C + 1) * 4));
    w.cell_start = reinterpret_cast<int32_t*>(take(F * (w.C + 1) * 4));
    w.atom_cell = reinterpret_cast<int32_t*>(take(F * N * 4));
    w.sorted_pos = reinterpret_cast<float4*>(take(F * N * 16));
    w.nblk = (F * N + 4095) / 4096;
    w.block_sums = reinterpret_cast<int64_t*>(take(w.nblk * 8));
    w.block_offs = reinterpret_cast<int64_t*>(take((w.nblk + 1) * 8));
    const bool small = N <= 2048;  // kRgCap
    w.nbr = reinterpret_cast<float4*>(take(small ? 0 : F * N * 64 * 16));  // kListCap
    w.list_ok = reinterpret_cast<int32_t*>(take(small ? 0 : F * N * 4));
    w.crowded_count = reinterpret_cast<int32_t*>(take(small ? 0 : 4));
    w.crowded_cells = reinterpret_cast<int2*>(take(small ? 0 : F * w.C * 8));
    w.hit_mask = reinterpret_cast<uint32_t*>(take(small ? F * N * ((N + 31) / 32) * 4 : 0));
    w.cta_edges = reinterpret_cast<int64_t*>(take(small ? F * (N / 8 + 1) * 8 : 0));
    w.bytes = off;
    return w;
}

// one block per frame: bounding box, then pick a cell edge >= cutoff whose grid fits in C cells
__global__ void __launch_bounds__(1024)
rg_frame_setup_kernel(const float* __restrict__ pos, int64_t N, double threshold, int64_t C,
                      FrameMeta* __restrict__ meta, int32_t* __restrict__ cell_count,
                      int32_t* __restrict__ crowded_count) {
    const int64_t f = blockIdx.x;
    // this frame's cell histogram and the crowded-cell worklist start from zero
    for (int64_t t = threadIdx.x; t <= C; t += blockDim.x) cell_count[f * (C + 1) + t] = 0;
    if (f == 0 && threadIdx.x == 0) *crowded_count = 0;
    const float* p = pos + f * N * 3;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float v = p[i * 3 + d];
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
    }
    __shared__ float smn[3][1024], smx[3][1024];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        smn[d][threadIdx.x] = mn[d];
        smx[d][threadIdx.x] = mx[d];
    }
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                smn[d][threadIdx.x] = fminf(smn[d][threadIdx.x], smn[d][threadIdx.x + s]);
                smx[d][threadIdx.x] = fmaxf(smx[d][threadIdx.x], smx[d][threadIdx.x + s]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        FrameMeta m;
        m.minx = smn[0][0];
        m.miny = smn[1][0];
        m.minz = smn[2][0];
        const double ext[3] = {double(smx[0][0]) - m.minx, double(smx[1][0]) - m.miny,
                               double(smx[2][0]) - m.minz};
        // cell edge strictly above the cutoff so that |d| < threshold implies adjacent cells;
        // grown until the grid fits the C reserved cells.  All in double so that huge or
        // non-finite coordinates (e.g. a diverged rollout) cannot overflow the cell counts:
        // non-finite extents fall back to a single cell (the predicate then rejects NaN/Inf).
        double cell = fmax(threshold, 1e-30) * 1.0001;
        int64_t nx = 1, ny = 1, nz = 1;
        const bool finite = isfinite(ext[0]) && isfinite(ext[1]) && isfinite(ext[2]) &&
                            isfinite(cell) && N > 0;
        if (finite) {
            for (;;) {
                const double dx = floor(ext[0] / cell) + 1.0, dy = floor(ext[1] / cell) + 1.0,
                             dz = floor(ext[2] / cell) + 1.0;
                if (dx * dy * dz <= double(C)) {
                    nx = int64_t(dx);
                    ny = int64_t(dy);
                    nz = int64_t(dz);
                    break;
                }
                cell *= 1.26;
            }
        } else {
            cell = 1.0;
            m.minx = m.miny = m.minz = 0.0;
        }
        m.inv_cell = 1.0 / cell;
        m.nx = int32_t(nx);
        m.ny = int32_t(ny);
        m.nz = int32_t(nz);
        m.ncell = int32_t(nx * ny * nz);
        meta[f] = m;
    }
}

__device__ __forceinline__ void cell_of(const FrameMeta& m, float x, float y, float z, int& cx,
                                        int& cy, int& cz) {
    cx = int(floor((double(x) - m.minx) * m.inv_cell));
    cy = int(floor((double(y) - m.miny) * m.inv_cell));
    cz = int(floor((double(z) - m.minz) * m.inv_cell));
    cx = min(max(cx, 0), m.nx - 1);
    cy = min(max(cy, 0), m.ny - 1);
    cz = min(max(cz, 0), m.nz - 1);
}

__global__ void rg_assign_kernel(const float* __restrict__ pos, int64_t F, int64_t N, int64_t C,
                                 const FrameMeta* __restrict__ meta,
                                 int32_t* __restrict__ atom_cell,
                                 int32_t* __restrict__ cell_count) {
    const int64_t total = F * N;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < total; g += stride) {
        const int64_t f = g / N;
        const FrameMeta m = meta[f];
        int cx, cy, cz;
        cell_of(m, pos[g * 3], pos[g * 3 + 1], pos[g * 3 + 2], cx, cy, cz);
        const int c = (cx * m.ny + cy) * m.nz + cz;
        atom_cell[g] = c;
        atomicAdd(cell_count + f * (C + 1) + c, 1);
    }
}

__global__ void rg_place_kernel(const float* __restrict__ pos, int64_t F, int64_t N, int64_t C,
                                const int32_t* __restrict__ atom_cell,
                                const int32_t* __restrict__ cell_start,
                                int32_t* __restrict__ cursor, float4* __restrict__ sorted_pos) {
    const int64_t total = F * N;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < total; g += stride) {
        const int64_t f = g / N;
        const int c = atom_cell[g];
        // the histogram doubles as the placement cursor, counted down to zero
        const int slot = cell_start[f * (C + 1) + c] + atomicSub(cursor + f * (C + 1) + c, 1) - 1;
        sorted_pos[f * N + slot] =
            make_float4(pos[g * 3], pos[g * 3 + 1], pos[g * 3 + 2], __int_as_float(int(g - f * N)));
    }
}

// Exact reference predicate without the square root.  SciPy evaluates, after promoting float32
// coordinates to float64,  fl(sqrt(fl(fl(dx*dx + dy*dy) + dz*dz))) < threshold.  Correctly
// rounded sqrt is monotone, so this equals  d2 <= d2_max  with d2_max the LARGEST double whose
// rounded square root is still below the threshold; exact_d2_bound() finds it on the host by
// stepping ulps around threshold^2.  d2 itself is formed with the same operations in the same
// order and no FMA contraction.
struct D2Bound {
    double d2_max;  // exact: hit  <=>  d2 (float64, reference operation order) <= d2_max
    float lo, hi;   // fp32 screen: d2f < lo => certainly a hit, d2f > hi => certainly a miss
};

__device__ __forceinline__ bool within(double xi, double yi, double zi, float4 pj, const D2Bound& b) {
    // fp32 screen first: the float32 squared distance is within a few ulps (< 2^-21 relative) of
    // the float64 one, so only pairs inside a 4e-6-wide band around the cutoff need float64.
    // NaNs fail both comparisons and fall through to the exact test, which rejects them.
    const float fx = pj.x - float(xi), fy = pj.y - float(yi), fz = pj.z - float(zi);
    const float d2f = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
    if (d2f < b.lo) return true;
    if (d2f > b.hi) return false;
    const double dx = __dsub_rn(double(pj.x), xi);
    const double dy = __dsub_rn(double(pj.y), yi);
    const double dz = __dsub_rn(double(pj.z), zi);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return d2 <= b.d2_max;
}

__device__ __forceinline__ bool within_exact(const float4& pi, const float4& pj, const D2Bound& b) {
    const double dx = __dsub_rn(double(pj.x), double(pi.x));
    const double dy = __dsub_rn(double(pj.y), double(pi.y));
    const double dz = __dsub_rn(double(pj.z), double(pi.z));
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return d2 <= b.d2_max;
}

// loads that must be ISSUED where they are written (software prefetch of the next row): volatile
// asm keeps nvcc from sinking them to their first use
__device__ __forceinline__ unsigned ld_nc_u32_here(const uint32_t* p) {
    unsigned v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_nc_f4_here(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ int64_t ld_nc_s64_here(const int64_t* p) {
    long long v;
    asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p));
    return int64_t(v);
}

// One warp handles one atom straight from global memory (sparse cells, and the general path for
// any cell occupancy).  The nine (x, y) neighbour columns are contiguous z-runs of the cell-sorted
// array; lanes 0..8 fetch their bounds in parallel and the runs are then walked as ONE flat
// candidate range, two independent coalesced chunks of 32 per step.
// FILL=false: count neighbours.  FILL=true: hits are parked in a per-warp smem buffer, ranked by
// atom index and written as the sorted row (coordinates of both endpoints come from the buffer);
// rows longer than the buffer use the row's own output slots as scratch instead.
constexpr int kHitCap = 256;  // per-warp hit buffer (float4)

template <bool FILL>
__device__ __forceinline__ void rg_atom_warp(float4 pi, int64_t node0, int cx, int cy, int cz,
                                             const FrameMeta& m, int lane,
                                             float4* __restrict__ hitbuf,
                                             const float* __restrict__ pos, const D2Bound& d2_max,
                                             const int32_t* __restrict__ cs,
                                             const float4* __restrict__ sp,
                                             int32_t* __restrict__ deg,
                                             const int64_t* __restrict__ row_ptr, int64_t E,
                                             int64_t* __restrict__ edge_index,
                                             float* __restrict__ edge_attr) {
    // the atom's cell (cx, cy, cz) is the centre cell of the calling CTA: nothing to recompute
    const int64_t g = node0 + __float_as_int(pi.w);
    const float px = pi.x, py = pi.y, pz = pi.z;
    const double xi = px, yi = py, zi = pz;
    int jb = 0, je = 0;
    if (lane < 9) {
        const int ax = cx - 1 + lane / 3, ay = cy - 1 + lane % 3;
        if (ax >= 0 && ax < m.nx && ay >= 0 && ay < m.ny) {
            const int z0 = max(cz - 1, 0), z1 = min(cz + 1, m.nz - 1);
            jb = cs[(ax * m.ny + ay) * m.nz + z0];
            je = cs[(ax * m.ny + ay) * m.nz + z1 + 1];
        }
    }
    const int len = je - jb;
    int incl = len;
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 15);
    const int excl = incl - len;
    int rb[9], ro[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        rb[r] = __shfl_sync(0xffffffffu, jb, r);
        ro[r] = __shfl_sync(0xffffffffu, excl, r);
    }
    auto candidate = [&](int t) -> float4 {
        int bsel = rb[0], osel = ro[0];
#pragma unroll
        for (int r = 1; r < 9; ++r) {
            if (t >= ro[r]) {
                bsel = rb[r];
                osel = ro[r];
            }
        }
        return sp[bsel + (t - osel)];
    };
    const int64_t beg = FILL ? row_ptr[g] : 0;
    // rows are sized by the count pass: deg is exact, so the buffer decision is warp-uniform
    const bool in_smem = FILL && hitbuf != nullptr && (row_ptr[g + 1] - beg) <= kHitCap;
    int64_t* scratch = FILL ? edge_index + beg : nullptr;  // row-id slots double as scratch
    int count = 0;
    for (int base = 0; base < total; base += 64) {  // warp-uniform trip count
        const int t0 = base + lane, t1 = base + 32 + lane;
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (t0 < total) p0 = candidate(t0);
        if (t1 < total) p1 = candidate(t1);
        const bool h0 = t0 < total && within(xi, yi, zi, p0, d2_max);
        const bool h1 = t1 < total && within(xi, yi, zi, p1, d2_max);
        const unsigned m0 = __ballot_sync(0xffffffffu, h0);
        const unsigned m1 = __ballot_sync(0xffffffffu, h1);
        if (FILL) {
            const unsigned below = (1u << lane) - 1u;
            const int s0 = count + __popc(m0 & below);
            const int s1 = count + __popc(m0) + __popc(m1 & below);
            if (in_smem) {
                if (h0) hitbuf[s0] = p0;
                if (h1) hitbuf[s1] = p1;
            } else {
                if (h0) scratch[s0] = int64_t(__float_as_int(p0.w));
                if (h1) scratch[s1] = int64_t(__float_as_int(p1.w));
            }
        }
        count += __popc(m0) + __popc(m1);
    }
    if (!FILL) {
        if (lane == 0) deg[g] = count;
        return;
    }
    __syncwarp();
    if (in_smem) {
        for (int j = lane; j < count; j += 32) {
            const float4 pj = hitbuf[j];
            const int v = __float_as_int(pj.w);
            int rank = 0;
            for (int k = 0; k < count; ++k) rank += (__float_as_int(hitbuf[k].w) < v) ? 1 : 0;
            const int64_t slot = beg + rank;
            edge_index[slot] = g;
            edge_index[E + slot] = node0 + v;
            float2* ea = reinterpret_cast<float2*>(edge_attr + slot * 6);
            ea[0] = make_float2(px, py);
            ea[1] = make_float2(pz, pj.x);
            ea[2] = make_float2(pj.y, pj.z);
        }
        __syncwarp();
        return;
    }
    // long row: rank-count sort of the (distinct) neighbour ids through the output slots
    int64_t* cols = edge_index + E + beg;
    for (int j = lane; j < count; j += 32) {
        const int64_t v = scratch[j];
        int rank = 0;
        for (int k = 0; k < count; ++k) rank += (scratch[k] < v) ? 1 : 0;
        cols[rank] = node0 + v;
    }
    __syncwarp();
    for (int j = lane; j < count; j += 32) {
        scratch[j] = g;  // row id (already offset by f*N)
        const int64_t c = cols[j];
        float* ea = edge_attr + (beg + j) * 6;
        ea[0] = px;
        ea[1] = py;
        ea[2] = pz;
        ea[3] = pos[c * 3];
        ea[4] = pos[c * 3 + 1];
        ea[5] = pos[c * 3 + 2];
    }
    __syncwarp();
}

// One CTA per (frame, cell): the atoms of the 27 surrounding cells are staged in shared memory
// once (coalesced float4 loads of the cell-sorted array) and shared by every atom of the centre
// cell.  For FILL the staged list is first sorted by atom index (bitonic, in smem), so each atom's
// hits come out in ascending column order and are written directly - no per-row sort, and both
// endpoints' coordinates for edge_attr come from smem.  Cells whose neighbourhood exceeds the
// staging capacity fall back to the per-atom global path.
constexpr int kRgCap = 2048;  // staged candidates (float4) = 32 KB
constexpr int kRgThreads = 256;
constexpr int kRgDenseMin = 24;  // centre-cell atoms from which staging + sorting pays off

// The CTA-cooperative treatment of ONE cell (see rg_cell_kernel): must be called by all threads of
// the CTA with the same arguments.
template <bool FILL>
__device__ __forceinline__ void rg_cell_cta(int c, const FrameMeta& m, int64_t node0, int64_t N,
                                            const D2Bound& d2_max, const float* __restrict__ pos,
                                            const int32_t* __restrict__ cs,
                                            const float4* __restrict__ sp, float4* s_cand,
                                            int* s_run_beg, int* s_run_off,
                                            int32_t* __restrict__ deg,
                                            const int64_t* __restrict__ row_ptr, int64_t E,
                                            int64_t* __restrict__ edge_index,
                                            float* __restrict__ edge_attr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ib = cs[c], ie = cs[c + 1];
    if (ib == ie) return;  // empty centre cell (block-uniform)
    const int cz = c % m.nz, cy = (c / m.nz) % m.ny, cx = c / (m.nz * m.ny);
    if (ie - ib < kRgDenseMin) {
        // sparse centre cell: staging + sorting would cost more than it saves
        __syncthreads();  // s_cand is about to be reused as per-warp hit buffers
        for (int i = ib + warp; i < ie; i += kRgThreads / 32) {
            rg_atom_warp<FILL>(sp[i], node0, cx, cy, cz, m, lane, s_cand + warp * kHitCap, pos,
                               d2_max, cs, sp, deg, row_ptr, E, edge_index, edge_attr);
        }
        return;
    }
    __syncthreads();  // previous item's smem no longer in use
    if (threadIdx.x < 9) {
        const int r = threadIdx.x;
        const int ax = cx - 1 + r / 3, ay = cy - 1 + r % 3;
        int jb = 0, je = 0;
        if (ax >= 0 && ax < m.nx && ay >= 0 && ay < m.ny) {
            const int z0 = max(cz - 1, 0), z1 = min(cz + 1, m.nz - 1);
            jb = cs[(ax * m.ny + ay) * m.nz + z0];
            je = cs[(ax * m.ny + ay) * m.nz + z1 + 1];
        }
        s_run_beg[r] = jb;
        s_run_off[r + 1] = je - jb;  // lengths first, prefix-summed below
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int off = 0;
        for (int r = 0; r < 9; ++r) {
            const int len = s_run_off[r + 1];
            s_run_off[r] = off;
            off += len;
        }
        s_run_off[9] = off;
    }
    __syncthreads();
    const int ncand = s_run_off[9];
    if (ncand > kRgCap) {
        // oversize neighbourhood: per-atom path straight from global memory
        __syncthreads();
        for (int i = ib + warp; i < ie; i += kRgThreads / 32) {
            rg_atom_warp<FILL>(sp[i], node0, cx, cy, cz, m, lane, s_cand + warp * kHitCap, pos,
                               d2_max, cs, sp, deg, row_ptr, E, edge_index, edge_attr);
        }
        return;
    }
    for (int r = 0; r < 9; ++r) {
        const int len = s_run_off[r + 1] - s_run_off[r];
        for (int t = threadIdx.x; t < len; t += kRgThreads) {
            s_cand[s_run_off[r] + t] = sp[s_run_beg[r] + t];
        }
    }
    int P = ncand;
    if (FILL) {
        P = 32;
        while (P < ncand) P <<= 1;
        for (int t = ncand + threadIdx.x; t < P; t += kRgThreads) {
            s_cand[t] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff));
        }
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < P; t += kRgThreads) {
                    const int u = t ^ j;
                    if (u > t) {
                        const float4 a = s_cand[t], b = s_cand[u];
                        const bool asc = (t & k) == 0;
                        if ((__float_as_int(a.w) > __float_as_int(b.w)) == asc) {
                            s_cand[t] = b;
                            s_cand[u] = a;
                        }
                    }
                }
                __syncthreads();
            }
        }
    } else {
        __syncthreads();
    }
    for (int i = ib + warp; i < ie; i += kRgThreads / 32) {
        const float4 pi = sp[i];
        const double xi = pi.x, yi = pi.y, zi = pi.z;
        const int64_t g = node0 + __float_as_int(pi.w);
        const int64_t beg = FILL ? row_ptr[g] : 0;
        int count = 0;
        for (int base = 0; base < ncand; base += 32) {
            const int j = base + lane;
            bool hit = false;
            float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < ncand) {
                pj = s_cand[j];
                hit = within(xi, yi, zi, pj, d2_max);
            }
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (FILL && hit) {
                const int64_t slot = beg + count + __popc(mask & ((1u << lane) - 1u));
                edge_index[slot] = g;
                edge_index[E + slot] = node0 + __float_as_int(pj.w);
                float2* ea = reinterpret_cast<float2*>(edge_attr + slot * 6);
                ea[0] = make_float2(pi.x, pi.y);
                ea[1] = make_float2(pi.z, pj.x);
                ea[2] = make_float2(pj.y, pj.z);
            }
            count += __popc(mask);
        }
        if (!FILL && lane == 0) deg[g] = count;
    }
}

template <bool FILL>
__global__ void __launch_bounds__(kRgThreads)
rg_cell_kernel(const float* __restrict__ pos, int64_t F, int64_t N, int64_t C, D2Bound d2_max,
               const FrameMeta* __restrict__ meta, const int32_t* __restrict__ cell_start,
               const float4* __restrict__ sorted_pos, int32_t* __restrict__ deg,
               const int64_t* __restrict__ row_ptr, int64_t E, int64_t* __restrict__ edge_index,
               float* __restrict__ edge_attr) {
    __shared__ float4 s_cand[kRgCap];
    __shared__ int s_run_beg[9], s_run_off[10];
    const int64_t f = blockIdx.y;
    const FrameMeta m = meta[f];
    const int32_t* cs = cell_start + f * (C + 1);
    const float4* sp = sorted_pos + f * N;
    for (int c = blockIdx.x; c < m.ncell; c += gridDim.x) {
        rg_cell_cta<FILL>(c, m, f * N, N, d2_max, pos, cs, sp, s_cand, s_run_beg, s_run_off, deg,
                          row_ptr, E, edge_index, edge_attr);
    }
}

// ------------------------------------------------------------------------------------------------
// Large frames, list variant: pass 1 does ALL the geometry and leaves, per atom, its neighbour
// list already sorted by atom index (<= kListCap entries of (x, y, z, index)) in the workspace; pass 2 is a pure
// streaming write of the 40 bytes per edge.
//   rg_cell_list_kernel  one WARP per cell.  The <= 256 atoms of the 27 surrounding cells are read
//                        once (coalesced float4 loads of the cell-sorted array) into registers,
//                        8 per lane; each atom of the centre cell is then tested against them
//                        (fp32 screen, float64 only inside the band), its hits are compacted into
//                        a per-warp smem buffer and rank-sorted from there.
//                        Cells whose neighbourhood exceeds 256 atoms and rows longer than kListCap
//                        only get their degree here and are rebuilt by pass 2 (rg_atom_warp).
//   rg_list_fill_kernel  one warp per row, rows in atom order so the output is one sequential
//                        stream; column ids and neighbour coordinates come from the saved list (no
//                        gathers), edge_attr through a 768-byte smem staging buffer so that
//                        consecutive lanes store consecutive float2.
// ------------------------------------------------------------------------------------------------
constexpr int kListCap = 64;
constexpr int kCellRegs = 8;  // candidates per lane held in registers

constexpr int kHitBuf = 32 * kCellRegs + 4;  // per-warp hit buffer: every candidate + sort padding

__device__ __forceinline__ void sts_u32(uint32_t addr, int v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ void sts_f4(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// One warp per cell, no CTA-level synchronisation: cells are handed out warp by warp.  Cells whose
// neighbourhood does not fit the registers go to a worklist for rg_crowded_kernel.
__global__ void __launch_bounds__(kRgThreads, 3)
rg_cell_list_kernel(int64_t F, int64_t N, int64_t C, D2Bound b, const FrameMeta* __restrict__ meta,
                    const int32_t* __restrict__ cell_start, const float4* __restrict__ sorted_pos,
                    int32_t* __restrict__ deg, float4* __restrict__ nbr,
                    int32_t* __restrict__ list_ok, int32_t* __restrict__ crowded_count,
                    int2* __restrict__ crowded_cells) {
    __shared__ __align__(16) int s_hit[kRgThreads / 32][kHitBuf];
    __shared__ float4 s_hitp[kRgThreads / 32][kListCap];  // coordinates of the first 64 hits
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t f = blockIdx.y;
    const FrameMeta m = meta[f];
    const int32_t* cs = cell_start + f * (C + 1);
    const float4* sp = sorted_pos + f * N;
    const int64_t node0 = f * N;
    const unsigned below = (1u << lane) - 1u;
    const int* hb = s_hit[warp];
    const uint32_t hb_addr = uint32_t(__cvta_generic_to_shared(s_hit[warp]));
    const float4* hp = s_hitp[warp];
    const uint32_t hp_addr = uint32_t(__cvta_generic_to_shared(s_hitp[warp]));
    constexpr int kWarps = kRgThreads / 32;
    for (int c = blockIdx.x * kWarps + warp; c < m.ncell; c += gridDim.x * kWarps) {
        const int ib = cs[c], ie = cs[c + 1];
        if (ib == ie) continue;  // warp-uniform
        const int cz = c % m.nz, cy = (c / m.nz) % m.ny, cx = c / (m.nz * m.ny);
        int jb = 0, je = 0;
        if (lane < 9) {
            const int ax = cx - 1 + lane / 3, ay = cy - 1 + lane % 3;
            if (ax >= 0 && ax < m.nx && ay >= 0 && ay < m.ny) {
                const int z0 = max(cz - 1, 0), z1 = min(cz + 1, m.nz - 1);
                jb = cs[(ax * m.ny + ay) * m.nz + z0];
                je = cs[(ax * m.ny + ay) * m.nz + z1 + 1];
            }
        }
        const int len = je - jb;
        int incl = len;
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const int T = __shfl_sync(0xffffffffu, incl, 15);
        if (T > 32 * kCellRegs) {
            if (lane == 0) crowded_cells[atomicAdd(crowded_count, 1)] = make_int2(int(f), c);
            continue;
        }
        const int excl = incl - len;
        float4 cand[kCellRegs];
        {
            int rb[9], ro[9];
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                rb[r] = __shfl_sync(0xffffffffu, jb, r);
                ro[r] = __shfl_sync(0xffffffffu, excl, r);
            }
#pragma unroll
            for (int q = 0; q < kCellRegs; ++q) {
                const int t = q * 32 + lane;
                int bsel = rb[0], osel = ro[0];
#pragma unroll
                for (int r = 1; r < 9; ++r) {
                    if (t >= ro[r]) {
                        bsel = rb[r];
                        osel = ro[r];
                    }
                }
                cand[q] = t < T ? sp[bsel + (t - osel)] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int a = ib; a < ie; ++a) {
            const float4 pi = sp[a];  // same address in every lane
            const int64_t g = node0 + __float_as_int(pi.w);
            int total = 0;
#pragma unroll
            for (int q = 0; q < kCellRegs; ++q) {
                if (q * 32 < T) {  // warp-uniform
                    const bool valid = lane < T - q * 32;
                    const float fx = cand[q].x - pi.x, fy = cand[q].y - pi.y, fz = cand[q].z - pi.z;
                    const float d2f = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
                    bool hit = valid && d2f < b.lo;
                    unsigned hm = __ballot_sync(0xffffffffu, hit);
                    const unsigned maybe = __ballot_sync(0xffffffffu, valid && !(d2f > b.hi));
                    if (hm != maybe) {  // warp-uniform, rare: some pair sits in the fp32 band
                        hit = valid && within_exact(pi, cand[q], b);
                        hm = __ballot_sync(0xffffffffu, hit);
                    }
                    if (hit) {
                        const int at = total + __popc(hm & below);
                        sts_u32(hb_addr + 4u * unsigned(at), __float_as_int(cand[q].w));
                        if (at < kListCap) sts_f4(hp_addr + 16u * unsigned(at), cand[q]);
                    }
                    total += __popc(hm);
                }
            }
            if (lane == 0) {
                deg[g] = total;
                list_ok[g] = total <= kListCap ? 1 : 0;
            }
            if (total == 0 || total > kListCap) continue;  // (total == 0: NaN coordinates)
            if (lane < 3) sts_u32(hb_addr + 4u * unsigned(total + lane), 0x7fffffff);  // padding
            __syncwarp();
            // rank sort straight from smem: independent compares, no shuffle dependency chains
            float4* out = nbr + g * kListCap;
            const int4* hb4 = reinterpret_cast<const int4*>(hb);
            const int v0 = hb[min(lane, total - 1)];
            const float4 p0 = hp[min(lane, total - 1)];
            int r0 = 0;
            if (total <= 32) {
                for (int k = 0; k < total; k += 4) {
                    const int4 h = hb4[k >> 2];
                    r0 += (h.x < v0) + (h.y < v0) + (h.z < v0) + (h.w < v0);
                }
                if (lane < total) out[r0] = p0;
            } else {
                const int v1 = hb[min(lane + 32, total - 1)];
                const float4 p1 = hp[min(lane + 32, total - 1)];
                int r1 = 0;
                for (int k = 0; k < total; k += 4) {
                    const int4 h = hb4[k >> 2];
                    r0 += (h.x < v0) + (h.y < v0) + (h.z < v0) + (h.w < v0);
                    r1 += (h.x < v1) + (h.y < v1) + (h.z < v1) + (h.w < v1);
                }
                out[r0] = p0;
                if (lane + 32 < total) out[r1] = p1;
            }
            __syncwarp();
        }
    }
}

// Crowded cells parked by rg_cell_list_kernel: the whole CTA works on one cell from shared memory
// (degrees only - pass 2 rebuilds these rows).
__global__ void __launch_bounds__(kRgThreads)
rg_crowded_kernel(const float* __restrict__ pos, int64_t N, int64_t C, D2Bound b,
                  const FrameMeta* __restrict__ meta, const int32_t* __restrict__ cell_start,
                  const float4* __restrict__ sorted_pos, int32_t* __restrict__ deg,
                  int32_t* __restrict__ list_ok, const int32_t* __restrict__ crowded_count,
                  const int2* __restrict__ crowded_cells) {
    __shared__ float4 s_cand[kRgCap];
    __shared__ int s_run_beg[9], s_run_off[10];
    const int n = *crowded_count;
    for (int w = blockIdx.x; w < n; w += gridDim.x) {
        const int2 fc = crowded_cells[w];
        const int64_t f = fc.x;
        const FrameMeta m = meta[f];
        const int32_t* cs = cell_start + f * (C + 1);
        const float4* sp = sorted_pos + f * N;
        rg_cell_cta<false>(fc.y, m, f * N, N, b, pos, cs, sp, s_cand, s_run_beg, s_run_off, deg,
                           nullptr, 0, nullptr, nullptr);
        for (int a = cs[fc.y] + threadIdx.x; a < cs[fc.y + 1]; a += kRgThreads) {
            list_ok[f * N + __float_as_int(sp[a].w)] = 0;
        }
    }
}

__global__ void __launch_bounds__(kRgThreads, 4)
rg_list_fill_kernel(const float* __restrict__ pos, int64_t F, int64_t N, int64_t C, D2Bound b,
                    const FrameMeta* __restrict__ meta, const int32_t* __restrict__ cell_start,
                    const float4* __restrict__ sorted_pos, const float4* __restrict__ nbr,
                    const int32_t* __restrict__ list_ok, const int64_t* __restrict__ row_ptr,
                    int64_t E, int64_t* __restrict__ edge_index, float* __restrict__ edge_attr) {
    __shared__ float2 s_stage[kRgThreads / 32][96];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* st = s_stage[warp];
    const int64_t total = F * N;
    const int64_t wstride = int64_t(gridDim.x) * (kRgThreads / 32);
    // software pipeline: everything the next row needs from global memory (row pointers, its saved
    // list, its own coordinates) is requested before the current row is written
    struct RowIn {
        int64_t beg, end;
        int ok;
        float4 c0, c1;  // neighbours lane and lane + 32 of the row: (x, y, z, atom index)
        float px, py, pz;
    };
    auto fetch = [&](int64_t g) {
        RowIn r;
        r.beg = ld_nc_s64_here(row_ptr + g);
        r.end = ld_nc_s64_here(row_ptr + g + 1);
        r.ok = int(ld_nc_u32_here(reinterpret_cast<const uint32_t*>(list_ok + g)));
        r.c0 = ld_nc_f4_here(nbr + g * kListCap + lane);
        r.c1 = ld_nc_f4_here(nbr + g * kListCap + 32 + lane);
        r.px = pos[g * 3];
        r.py = pos[g * 3 + 1];
        r.pz = pos[g * 3 + 2];
        return r;
    };
    int64_t g = int64_t(blockIdx.x) * (kRgThreads / 32) + warp;
    RowIn nxt;
    if (g < total) nxt = fetch(g);
    for (; g < total; g += wstride) {
        const RowIn cur = nxt;
        if (g + wstride < total) nxt = fetch(g + wstride);
        const int64_t f = g / N;
        const int64_t node0 = f * N;
        const int64_t beg = cur.beg;
        const int d = int(cur.end - beg);
        const float px = cur.px, py = cur.py, pz = cur.pz;
        if (!cur.ok) {
            // rebuild this row from the cell list (crowded cell or more than kListCap neighbours);
            // the row's own output slots serve as scratch
            const FrameMeta m = meta[f];
            int cx, cy, cz;
            cell_of(m, px, py, pz, cx, cy, cz);
            rg_atom_warp<true>(make_float4(px, py, pz, __int_as_float(int(g - node0))), node0, cx, cy,
                               cz, m, lane, nullptr, pos, b, cell_start + f * (C + 1),
                               sorted_pos + f * N, nullptr, row_ptr, E, edge_index, edge_attr);
            __syncwarp();
            continue;
        }
        int64_t* er = edge_index + beg;
        int64_t* ec = edge_index + E + beg;
        float2* ea = reinterpret_cast<float2*>(edge_attr + beg * 6);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int e0 = 32 * h;
            if (e0 >= d) break;  // warp-uniform
            const int e = e0 + lane;
            if (e < d) {
                const float4 pj = h == 0 ? cur.c0 : cur.c1;
                er[e] = g;
                ec[e] = node0 + __float_as_int(pj.w);
                st[3 * lane] = make_float2(px, py);
                st[3 * lane + 1] = make_float2(pz, pj.x);
                st[3 * lane + 2] = make_float2(pj.y, pj.z);
            }
            __syncwarp();
            const int n2 = 3 * min(32, d - e0);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int u = 32 * k + lane;
                if (u < n2) ea[96 * h + u] = st[u];
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Small frames (N <= kRgCap atoms, e.g. the 28-residue and 504-atom BBA graphs): no cell list.
// Pass 1 evaluates every pair ONCE and keeps the answer as a bit matrix (1 bit per pair, N^2/8
// bytes per frame - 2 MB for 64 BBA frames, L2 resident); pass 2 is a pure streaming write of the
// 40 bytes per edge, fully coalesced, driven by that bit matrix.
//   rg_small_mask_kernel   CTA = (row slice, frame): the frame's coordinates are staged in smem as
//                          float4; a warp keeps 32 candidate atoms in registers and sweeps the CTA's
//                          rows (one broadcast LDS.128 + ~14 instructions per 32 pairs); the hit
//                          words go to smem, then to `mask` (coalesced) with the row degrees and
//                          the CTA's edge total.
//   rg_small_rowptr_kernel same decomposition: CTA b adds the totals of CTAs < b and scans its own
//                          <= 256 degrees -> row_ptr (int64).  No chained scan, no spinning.
//   rg_small_fill_kernel   a warp expands a row's bit words into a column list in smem and streams
//                          row ids, column ids and edge_attr with consecutive lanes on consecutive
//                          addresses.  Columns come out ascending by construction.
// ------------------------------------------------------------------------------------------------
constexpr int kSmThreads = 256;
constexpr int kSmWarps = kSmThreads / 32;
constexpr int kSmMaskWords = 4096;  // smem hit words per CTA in pass 1 (16 KB)
constexpr int kSmFixCap = 256;      // hit words per CTA that may need the float64 predicate

__device__ __forceinline__ void stage_frame_float4(const float* __restrict__ p, int n,
                                                   float4* __restrict__ s_p) {
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        s_p[t] = make_float4(p[t * 3], p[t * 3 + 1], p[t * 3 + 2], 0.f);
    }
}

__global__ void __launch_bounds__(kSmThreads)
rg_small_mask_kernel(const float* __restrict__ pos, int N, int rows_per, D2Bound b,
                     int32_t* __restrict__ deg, uint32_t* __restrict__ mask,
                     int64_t* __restrict__ partial) {
    extern __shared__ float4 s_dyn[];
    float4* s_p = s_dyn;                                        // [N]
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_p + N);    // [rows_per, nws]
    __shared__ int s_wsum[kSmWarps];
    __shared__ int s_nfix;
    __shared__ uint32_t s_fix[kSmFixCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = (N + 31) >> 5;
    const int nws = (nw + 2 * kSmWarps - 1) / (2 * kSmWarps) * (2 * kSmWarps);  // padded row stride
    const unsigned tail = (N & 31) ? (1u << (N & 31)) - 1u : 0xffffffffu;
    const int64_t f = blockIdx.y;
    const int r0 = blockIdx.x * rows_per, r1 = min(N, r0 + rows_per);
    if (threadIdx.x == 0) s_nfix = 0;
    stage_frame_float4(pos + f * N * 3, N, s_p);
    __syncthreads();
    // Two column chunks (64 candidate atoms in registers) per sweep over the CTA's rows.  Lanes
    // past the last atom re-test atom N-1 (their bits are cleared when the words leave smem), and
    // all lanes store the warp-uniform hit word, so the sweep has no divergent code at all.
    for (int c = warp; c < nw; c += 2 * kSmWarps) {
        const int c2 = c + kSmWarps;  // < nws: lands in the row's padding when >= nw
        const float4 pj = s_p[min(c * 32 + lane, N - 1)];
        const float4 pk = s_p[min(c2 * 32 + lane, N - 1)];
        uint32_t* wrow = s_mask + c;
#pragma unroll 2
        for (int i = r0; i < r1; ++i, wrow += nws) {
            const float4 pi = s_p[i];
            const float fx = pj.x - pi.x, fy = pj.y - pi.y, fz = pj.z - pi.z;
            const float gx = pk.x - pi.x, gy = pk.y - pi.y, gz = pk.z - pi.z;
            const float d2f = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
            const float d2g = fmaf(gz, gz, fmaf(gy, gy, gx * gx));
            // fp32 screen: below `lo` certainly a hit, above `hi` certainly a miss; words with a
            // pair in between (or a NaN, which fails both comparisons) are redone exactly below
            const unsigned sure = __ballot_sync(0xffffffffu, d2f < b.lo);
            const unsigned maybe = __ballot_sync(0xffffffffu, !(d2f > b.hi));
            const unsigned sure2 = __ballot_sync(0xffffffffu, d2g < b.lo);
            const unsigned maybe2 = __ballot_sync(0xffffffffu, !(d2g > b.hi));
            wrow[0] = sure;
            wrow[kSmWarps] = sure2;
            if ((sure ^ maybe) | (sure2 ^ maybe2)) {  // warp-uniform, rare
                if (lane == 0) {
                    if (sure != maybe) {
                        const int k = atomicAdd(&s_nfix, 1);
                        if (k < kSmFixCap) s_fix[k] = uint32_t(i - r0) << 16 | uint32_t(c);
                    }
                    if (sure2 != maybe2 && c2 < nw) {
                        const int k = atomicAdd(&s_nfix, 1);
                        if (k < kSmFixCap) s_fix[k] = uint32_t(i - r0) << 16 | uint32_t(c2);
                    }
                }
            }
        }
    }
    __syncthreads();
    const int nfix = s_nfix;
    if (nfix > kSmFixCap) {
        // pathological input (e.g. a lattice at exactly the cutoff spacing): everything in float64
        for (int c = warp; c < nw; c += kSmWarps) {
            const int j = c * 32 + lane;
            const float4 pj = s_p[min(j, N - 1)];
            for (int i = r0; i < r1; ++i) {
                const unsigned m = __ballot_sync(0xffffffffu, j < N && within_exact(s_p[i], pj, b));
                if (lane == 0) s_mask[(i - r0) * nws + c] = m;
            }
        }
    } else {
        for (int k = warp; k < nfix; k += kSmWarps) {
            const int ir = int(s_fix[k] >> 16), c = int(s_fix[k] & 0xffffu);
            const int j = c * 32 + lane;
            const float4 pj = s_p[min(j, N - 1)];
            const unsigned m = __ballot_sync(0xffffffffu, j < N && within_exact(s_p[r0 + ir], pj, b));
            if (lane == 0) s_mask[ir * nws + c] = m;
        }
    }
    __syncthreads();
    int wsum = 0;
    for (int i = r0 + warp; i < r1; i += kSmWarps) {
        const int64_t g = f * N + i;
        int cnt = 0;
        for (int c = lane; c < nw; c += 32) {
            const unsigned m = s_mask[(i - r0) * nws + c] & (c == nw - 1 ? tail : 0xffffffffu);
            mask[g * nw + c] = m;
            cnt += __popc(m);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) deg[g] = cnt;
        wsum += cnt;
    }
    if (lane == 0) s_wsum[warp] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kSmWarps; ++w) tot += s_wsum[w];
        partial[int64_t(blockIdx.y) * gridDim.x + blockIdx.x] = tot;
    }
}

__global__ void __launch_bounds__(kSmThreads)
rg_small_rowptr_kernel(const int32_t* __restrict__ deg, const int64_t* __restrict__ partial, int N,
                       int rows_per, int64_t* __restrict__ row_ptr) {
    __shared__ int64_t s_red[kSmWarps];
    __shared__ int s_wtot[kSmWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t bid = int64_t(blockIdx.y) * gridDim.x + blockIdx.x;
    int64_t v = 0;
    for (int64_t t = threadIdx.x; t < bid; t += kSmThreads) v += partial[t];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) s_red[warp] = v;
    const int r0 = blockIdx.x * rows_per, r1 = min(N, r0 + rows_per);
    const int i = r0 + threadIdx.x;  // rows_per <= kSmThreads
    const int64_t g = int64_t(blockIdx.y) * N + i;
    const int d = i < r1 ? deg[g] : 0;
    int incl = d;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    int64_t base = 0;
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kSmWarps; ++w) {
        base += s_red[w];
        if (w < warp) before += s_wtot[w];
        total += s_wtot[w];
    }
    if (i < r1) row_ptr[g] = base + before + incl - d;
    if (bid == int64_t(gridDim.x) * gridDim.y - 1 && threadIdx.x == 0) {
        row_ptr[int64_t(gridDim.y) * N] = base + total;
    }
}

__global__ void __launch_bounds__(kSmThreads)
rg_small_fill_kernel(const float* __restrict__ pos, int N, int rows_per, int col_cap,
                     const uint32_t* __restrict__ mask, const int64_t* __restrict__ row_ptr,
                     int64_t E, int64_t* __restrict__ edge_index, float* __restrict__ edge_attr) {
    extern __shared__ float4 s_dyn[];
    float4* s_p = s_dyn;  // [N]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* s_stage = reinterpret_cast<float2*>(s_p + N) + warp * 96;  // 32 edges x 6 floats
    unsigned char* s_cols =
        reinterpret_cast<unsigned char*>(reinterpret_cast<float2*>(s_p + N) + kSmWarps * 96) +
        size_t(warp) * col_cap * 2;  // uint16 column list of the warp's current row
    const int nw = (N + 31) >> 5;
    const int64_t f = blockIdx.y;
    const int64_t node0 = f * N;
    const int r0 = blockIdx.x * rows_per, r1 = min(N, r0 + rows_per);
    stage_frame_float4(pos + node0 * 3, N, s_p);
    __syncthreads();
    const unsigned below = (1u << lane) - 1u, mybit = 1u << lane;
    // software pipeline: the next row's first 32 hit words and its row pointer are in flight
    // while the current row is expanded and written
    int i = r0 + warp;
    unsigned wnext = 0u;
    int64_t bnext = 0;
    if (i < r1) {
        if (lane < nw) wnext = mask[(node0 + i) * nw + lane];
        bnext = row_ptr[node0 + i];
    }
    for (; i < r1; i += kSmWarps) {
        const int64_t g = node0 + i;
        const uint32_t* mrow = mask + g * nw;
        unsigned wreg = wnext;
        const int64_t beg = bnext;
        if (i + kSmWarps < r1) {
            if (lane < nw) wnext = ld_nc_u32_here(mask + (g + kSmWarps) * nw + lane);
            bnext = ld_nc_s64_here(row_ptr + g + kSmWarps);
        }
        // 1. expand the row's hit words into an ascending column list (uint16) in smem
        int d = 0;
        for (int c0 = 0; c0 < nw; c0 += 32) {
            if (c0 > 0) wreg = c0 + lane < nw ? mrow[c0 + lane] : 0u;
            const int pc = __popc(wreg);
            int incl = pc;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            const int excl2 = 2 * (d + incl - pc);  // byte offset of this word's first column
            const int nn = min(32, nw - c0);
            int col = c0 * 32 + lane;
#pragma unroll 4
            for (int cc = 0; cc < nn; ++cc, col += 32) {
                const unsigned m = __shfl_sync(0xffffffffu, wreg, cc);
                const int o = __shfl_sync(0xffffffffu, excl2, cc);
                if (m & mybit) {
                    *reinterpret_cast<uint16_t*>(s_cols + o + 2 * __popc(m & below)) = uint16_t(col);
                }
            }
            d += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        // 2. stream the row: 32 edges per round; edge_attr goes through a 768-byte smem staging
        //    buffer so that consecutive lanes store consecutive float2
        int64_t* er = edge_index + beg;
        int64_t* ec = edge_index + E + beg;
        const float4 pi = s_p[i];
        float2* ea = reinterpret_cast<float2*>(edge_attr + beg * 6);  // beg*24 bytes: 8-aligned
        for (int e0 = 0; e0 < d; e0 += 32, ea += 96) {
            const int e = e0 + lane;
            if (e < d) {
                const int col = reinterpret_cast<const uint16_t*>(s_cols)[e];
                const float4 pj = s_p[col];
                er[e] = g;
                ec[e] = node0 + col;
                s_stage[3 * lane] = make_float2(pi.x, pi.y);
                s_stage[3 * lane + 1] = make_float2(pi.z, pj.x);
                s_stage[3 * lane + 2] = make_float2(pj.y, pj.z);
            }
            __syncwarp();
            const int n2 = 3 * min(32, d - e0);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int u = 32 * k + lane;
                if (u < n2) ea[u] = s_stage[u];
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// multi-block exclusive scan (int32 -> int64) for the row pointers: block sums, scan of the sums
// (single block), then per-block scan + offset
// ------------------------------------------------------------------------------------------------
constexpr int kScanItems = 4;
constexpr int kScanChunk = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads)
scan_block_sums_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ sums) {
    __shared__ int64_t sm[32];
    const int64_t base = int64_t(blockIdx.x) * kScanChunk;
    int64_t v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + int64_t(threadIdx.x) * kScanItems + k;
        if (i < n) v += in[i];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int64_t t = sm[threadIdx.x];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (threadIdx.x == 0) sums[blockIdx.x] = t;
    }
}

template <typename TOut>
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ offs,
                  TOut* __restrict__ out) {
    __shared__ int64_t warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = int64_t(blockIdx.x) * kScanChunk + int64_t(threadIdx.x) * kScanItems;
    int64_t v[kScanItems], tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        tsum += v[k];
    }
    int64_t incl = tsum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int64_t wv = warp_tot[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, wv, off);
            if (lane >= off) wv += t;
        }
        warp_tot[lane] = wv;
    }
    __syncthreads();
    int64_t run = offs[blockIdx.x] + (warp == 0 ? 0 : warp_tot[warp - 1]) + incl - tsum;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = TOut(run);
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = TOut(offs[gridDim.x]);
}

// Row slices of the small-frame kernels: pass 1 (mask + row pointers) and pass 2 (fill).
struct SmallSplit {
    int rows_per, gx;
};
SmallSplit small_split(int64_t F, int64_t N, int sms, int ctas_per_sm, int min_rows, int max_rows) {
    int64_t gx_want = (int64_t(sms) * ctas_per_sm + F - 1) / F;
    if (gx_want < 1) gx_want = 1;
    int64_t rows = (N + gx_want - 1) / gx_want;
    if (rows < min_rows) rows = min_rows;
    if (rows > max_rows) rows = max_rows;
    if (rows > N) rows = N;
    if (rows < 1) rows = 1;
    SmallSplit sp;
    sp.rows_per = int(rows);
    sp.gx = int((N + rows - 1) / rows);
    return sp;
}
SmallSplit small_split_pass1(int64_t F, int64_t N, int sms) {
    const int nws = int(((N + 31) / 32 + 2 * kSmWarps - 1) / (2 * kSmWarps) * (2 * kSmWarps));
    int max_rows = kSmMaskWords / (nws < 1 ? 1 : nws);
    if (max_rows > kSmThreads) max_rows = kSmThreads;
    return small_split(F, N, sms, 4, 8, max_rows);  // gx <= N/8 + 1 (cta_edges capacity)
}

// largest double d2 with fl(sqrt(d2)) < thr  (see `within`); -1 if no distance qualifies; plus the
// conservative float32 screening band around it
D2Bound exact_d2_bound(double thr) {
    D2Bound b;
    b.d2_max = -1.0;
    if (thr > 0.0) {  // false for NaN too
        if (std::isinf(thr)) {
            b.d2_max = std::numeric_limits<double>::max();
        } else {
            double t = thr * thr;
            if (std::isinf(t)) t = std::numeric_limits<double>::max();
            while (t > 0.0 && std::sqrt(t) >= thr) t = std::nextafter(t, -1.0);
            while (std::sqrt(std::nextafter(t, std::numeric_limits<double>::infinity())) < thr) {
                t = std::nextafter(t, std::numeric_limits<double>::infinity());
            }
            b.d2_max = std::sqrt(t) < thr ? t : -1.0;
        }
    }
    if (b.d2_max < 0.0) {
        b.lo = -1.0f;  // d2f >= 0 is never below it
        b.hi = -1.0f;  // everything finite is above it => miss (NaN falls to the exact test)
    } else {
        const double lo = b.d2_max * (1.0 - 2e-6), hi = b.d2_max * (1.0 + 2e-6);
        b.lo = std::nextafterf(float(lo), -1.0f);
        if (double(b.lo) > lo) b.lo = std::nextafterf(b.lo, -1.0f);
        b.hi = hi >= double(std::numeric_limits<float>::max())
                   ? std::numeric_limits<float>::infinity()
                   : std::nextafterf(float(hi), std::numeric_limits<float>::infinity());
    }
    return b;
}

int grid_for(int64_t total, int threads, int per_sm) {
    int sms = 148;
    device_sm_count(&sms);
    int64_t blocks = (total + threads - 1) / threads;
    const int64_t cap = int64_t(sms) * per_sm;
    if (blocks > cap) blocks = cap;
    return int(blocks < 1 ? 1 : blocks);
}

}  // namespace
}  // namespace gkn

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int64_t gkn_graph_prepare_workspace_bytes(int64_t N, int64_t E) {
    const int64_t nblk = (N + 4095) / 4096;  // row-pointer scans: block sums + their offsets
    return (2 * N + 2 * E + 64) * int64_t(sizeof(int32_t)) + (2 * nblk + 4) * 8;
}

extern "C" int gkn_graph_prepare(const int64_t* edge_index, int64_t N, int64_t E, int32_t* perm,
                                 int32_t* src_s, int32_t* dst_s, int32_t* dst_ptr, float* inv_cnt,
                                 int32_t* src_ptr, int32_t* src_perm, int32_t* status,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace gkn;
    if (N < 0 || E < 0) return set_error("gkn_graph_prepare: negative size");
    if (N >= (int64_t(1) << 31) || E >= (int64_t(1) << 31)) {
        return set_error("gkn_graph_prepare: N and E must fit in int32");
    }
    if (workspace_bytes < gkn_graph_prepare_workspace_bytes(N, E)) {
        return set_error("gkn_graph_prepare: workspace too small");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int32_t* cnt_dst = static_cast<int32_t*>(workspace);
    int32_t* cnt_src = cnt_dst + N;
    int32_t* tmp = cnt_src + N;
    int32_t* tmp2 = tmp + E;
    const int64_t nblk = (N + 4095) / 4096;
    int64_t* blk_sums = reinterpret_cast<int64_t*>(
        static_cast<uint8_t*>(workspace) + (2 * N + 2 * E + 64) / 2 * 2 * int64_t(sizeof(int32_t)));
    int64_t* blk_offs = blk_sums + nblk + 1;
    auto scan_rows = [&](const int32_t* cnt, int32_t* ptr) -> int {
        if (nblk <= 1) {
            exclusive_scan_kernel<int32_t, int32_t><<<1, kScanThreads, 0, st>>>(cnt, ptr, N, 0, 0);
            GKN_CHECK_LAUNCH("prep_scan");
            return 0;
        }
        scan_block_sums_kernel<<<unsigned(nblk), kScanThreads, 0, st>>>(cnt, N, blk_sums);
        GKN_CHECK_LAUNCH("prep_scan_sums");
        exclusive_scan_kernel<int64_t, int64_t><<<1, kScanThreads, 0, st>>>(blk_sums, blk_offs, nblk, 0, 0);
        GKN_CHECK_LAUNCH("prep_scan_offs");
        scan_apply_kernel<int32_t><<<unsigned(nblk), kScanThreads, 0, st>>>(cnt, N, blk_offs, ptr);
        GKN_CHECK_LAUNCH("prep_scan_rows");
        return 0;
    };
    cudaError_t err = cudaMemsetAsync(cnt_dst, 0, size_t(2 * N) * 4, st);
    if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(prepare)", err);
    err = cudaMemsetAsync(status, 0, 4, st);
    if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(status)", err);
    const int eg = grid_for(E, 256, 8);
    const int ng = grid_for(N * 32, 256, 8);
    prep_count_kernel<<<eg, 256, 0, st>>>(edge_index, N, E, cnt_dst, cnt_src, status);
    GKN_CHECK_LAUNCH("prep_count");
    if (int rc = scan_rows(cnt_dst, dst_ptr)) return rc;
    if (int rc = scan_rows(cnt_src, src_ptr)) return rc;
    prep_inv_cnt_kernel<<<grid_for(N, 256, 8), 256, 0, st>>>(dst_ptr, inv_cnt, N);
    GKN_CHECK_LAUNCH("prep_inv_cnt");
    err = cudaMemsetAsync(cnt_dst, 0, size_t(2 * N) * 4, st);
    if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(prepare cursors)", err);
    if (E > 0) {
        prep_place_dst_kernel<<<eg, 256, 0, st>>>(edge_index, N, E, dst_ptr, cnt_dst, tmp);
        GKN_CHECK_LAUNCH("prep_place_dst");
        prep_canon_dst_kernel<<<ng, 256, 0, st>>>(edge_index, N, E, dst_ptr, tmp, perm, src_s,
                                                  dst_s);
        GKN_CHECK_LAUNCH("prep_canon_dst");
        prep_place_src_kernel<<<eg, 256, 0, st>>>(src_s, E, src_ptr, cnt_src, tmp2);
        GKN_CHECK_LAUNCH("prep_place_src");
        prep_canon_src_kernel<<<ng, 256, 0, st>>>(N, src_ptr, tmp2, src_perm);
        GKN_CHECK_LAUNCH("prep_canon_src");
    }
    return 0;
}

extern "C" int64_t gkn_radius_graph_workspace_bytes(int64_t F, int64_t N) {
    if (F < 0 || N < 0) return 0;
    return gkn::carve_radius_ws(nullptr, F, N).bytes;
}

extern "C" int gkn_radius_graph_count(const float* pos, int64_t F, int64_t N, double threshold,
                                      int32_t* deg, int64_t* row_ptr, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
    using namespace gkn;
    if (F < 0 || N < 0) return set_error("gkn_radius_graph_count: negative size");
    if (pos == nullptr || deg == nullptr || row_ptr == nullptr || workspace == nullptr) {
        return set_error("gkn_radius_graph_count: NULL pointer argument");
    }
    if (N >= (int64_t(1) << 30)) return set_error("gkn_radius_graph_count: N too large");
    if (F > 65535) return set_error("gkn_radius_graph_count: at most 65535 frames per call");
    RadiusWs w = carve_radius_ws(workspace, F, N);
    if (workspace_bytes < w.bytes) return set_error("gkn_radius_graph_count: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t total = F * N;
    if (total == 0) {
        cudaError_t err = cudaMemsetAsync(row_ptr, 0, 8, st);
        if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(row_ptr)", err);
        return 0;
    }
    int sms = 148;
    device_sm_count(&sms);
    if (N <= kRgCap) {
        // small frames: all-pairs from shared memory, no cell list; hits kept as a bit matrix
        const SmallSplit sp = small_split_pass1(F, N, sms);
        const int nws = int(((N + 31) / 32 + 2 * kSmWarps - 1) / (2 * kSmWarps) * (2 * kSmWarps));
        const dim3 sg(unsigned(sp.gx), unsigned(F));
        const size_t smem = size_t(N) * 16 + size_t(sp.rows_per) * nws * 4;
        rg_small_mask_kernel<<<sg, kSmThreads, smem, st>>>(pos, int(N), sp.rows_per,
                                                           exact_d2_bound(threshold), deg,
                                                           w.hit_mask, w.cta_edges);
        GKN_CHECK_LAUNCH("rg_small_mask");
        if (int64_t(sp.gx) * F <= 4096) {
            rg_small_rowptr_kernel<<<sg, kSmThreads, 0, st>>>(deg, w.cta_edges, int(N), sp.rows_per,
                                                              row_ptr);
            GKN_CHECK_LAUNCH("rg_small_rowptr");
            return 0;
        }
    } else {
        rg_frame_setup_kernel<<<unsigned(F), 1024, 0, st>>>(pos, N, threshold, w.C, w.meta,
                                                            w.cell_count, w.crowded_count);
        GKN_CHECK_LAUNCH("rg_frame_setup");
        const int ag = grid_for(total, 256, 8);
        rg_assign_kernel<<<ag, 256, 0, st>>>(pos, F, N, w.C, w.meta, w.atom_cell, w.cell_count);
        GKN_CHECK_LAUNCH("rg_assign");
        exclusive_scan_kernel<int32_t, int32_t><<<unsigned(F), kScanThreads, 0, st>>>(
            w.cell_count, w.cell_start, w.C, w.C + 1, w.C + 1);
        GKN_CHECK_LAUNCH("rg_scan_cells");
        rg_place_kernel<<<ag, 256, 0, st>>>(pos, F, N, w.C, w.atom_cell, w.cell_start,
                                            w.cell_count, w.sorted_pos);
        GKN_CHECK_LAUNCH("rg_place");
        // one warp per cell; cells the registers cannot hold are parked for the CTA-wide kernel
        int64_t cgx = (int64_t(sms) * 16 + F - 1) / F;
        if (cgx > (w.C + 7) / 8) cgx = (w.C + 7) / 8;
        const dim3 cg(unsigned(cgx < 1 ? 1 : cgx), unsigned(F));
        const D2Bound bound = exact_d2_bound(threshold);
        rg_cell_list_kernel<<<cg, kRgThreads, 0, st>>>(F, N, w.C, bound, w.meta, w.cell_start,
                                                      w.sorted_pos, deg, w.nbr, w.list_ok,
                                                      w.crowded_count, w.crowded_cells);
        GKN_CHECK_LAUNCH("rg_cell_list");
        rg_crowded_kernel<<<unsigned(sms) * 4, kRgThreads, 0, st>>>(
            pos, N, w.C, bound, w.meta, w.cell_start, w.sorted_pos, deg, w.list_ok,
            w.crowded_count, w.crowded_cells);
        GKN_CHECK_LAUNCH("rg_crowded");
    }
    scan_block_sums_kernel<<<unsigned(w.nblk), kScanThreads, 0, st>>>(deg, total, w.block_sums);
    GKN_CHECK_LAUNCH("rg_scan_sums");
    exclusive_scan_kernel<int64_t, int64_t>
        <<<1, kScanThreads, 0, st>>>(w.block_sums, w.block_offs, w.nblk, 0, 0);
    GKN_CHECK_LAUNCH("rg_scan_offs");
    scan_apply_kernel<int64_t><<<unsigned(w.nblk), kScanThreads, 0, st>>>(deg, total, w.block_offs,
                                                                        row_ptr);
    GKN_CHECK_LAUNCH("rg_scan_rows");
    return 0;
}

extern "C" int gkn_radius_graph_fill(const float* pos, int64_t F, int64_t N, double threshold,
                                     const int64_t* row_ptr, int64_t E, int64_t* edge_index,
                                     float* edge_attr, void* workspace, int64_t workspace_bytes,
                                     void* stream) {
    using namespace gkn;
    if (F < 0 || N < 0 || E < 0) return set_error("gkn_radius_graph_fill: negative size");
    if (pos == nullptr || row_ptr == nullptr || workspace == nullptr ||
        (E > 0 && (edge_index == nullptr || edge_attr == nullptr))) {
        return set_error("gkn_radius_graph_fill: NULL pointer argument");
    }
    RadiusWs w = carve_radius_ws(workspace, F, N);
    if (workspace_bytes < w.bytes) return set_error("gkn_radius_graph_fill: workspace too small");
    const int64_t total = F * N;
    if (total == 0 || E == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int sms = 148;
    device_sm_count(&sms);
    if (N <= kRgCap) {
        const SmallSplit sp = small_split(F, N, sms, 6, 8, int(N));
        const int col_cap = int((N + 7) / 8 * 8);
        const size_t smem = size_t(N) * 16 + size_t(kSmWarps) * (96 * 8 + col_cap * 2);
        if (smem > 48 * 1024) {
            cudaError_t err = cudaFuncSetAttribute(
                rg_small_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(rg_small_fill)", err);
        }
        const dim3 sg(unsigned(sp.gx), unsigned(F));
        rg_small_fill_kernel<<<sg, kSmThreads, smem, st>>>(pos, int(N), sp.rows_per, col_cap,
                                                           w.hit_mask, row_ptr, E, edge_index,
                                                           edge_attr);
    } else {
        int64_t cgx = (int64_t(sms) * 16 + F - 1) / F;
        if (cgx > w.C) cgx = w.C;
        const dim3 cg(unsigned(cgx < 1 ? 1 : cgx), unsigned(F));
        if (E > total * 40) {
            // dense graphs (most rows longer than the saved lists): rebuild per cell from smem
            rg_cell_kernel<true><<<cg, kRgThreads, 0, st>>>(pos, F, N, w.C,
                                                           exact_d2_bound(threshold), w.meta,
                                                           w.cell_start, w.sorted_pos, nullptr,
                                                           row_ptr, E, edge_index, edge_attr);
        } else {
            const int fg = grid_for(total * 32, kRgThreads, 8);
            rg_list_fill_kernel<<<fg, kRgThreads, 0, st>>>(pos, F, N, w.C, exact_d2_bound(threshold),
                                                          w.meta, w.cell_start, w.sorted_pos, w.nbr,
                                                          w.list_ok, row_ptr, E, edge_index,
                                                          edge_attr);
        }
    }
    GKN_CHECK_LAUNCH("rg_fill");
    return 0;
}

extern "C" int gkn_plan_from_csr(const int64_t* edge_index, const int64_t* row_ptr, int64_t N,
                                 int64_t E, int32_t* rev, int32_t* src_s, int32_t* dst_s,
                                 int32_t* ptr32, float* inv_cnt, int32_t* status, void* stream) {
    using namespace gkn;
    if (N < 0 || E < 0) return set_error("gkn_plan_from_csr: negative size");
    if (N >= (int64_t(1) << 31) || E >= (int64_t(1) << 31)) {
        return set_error("gkn_plan_from_csr: N and E must fit in int32");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t err = cudaMemsetAsync(status, 0, 4, st);
    if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(status)", err);
    const int64_t work = E > N + 1 ? E : N + 1;
    plan_from_csr_kernel<<<grid_for(work, 256, 8), 256, 0, st>>>(edge_index, row_ptr, N, E, rev,
                                                                src_s, dst_s, ptr32, inv_cnt,
                                                                status);
    GKN_CHECK_LAUNCH("plan_from_csr");
    return 0;
}
