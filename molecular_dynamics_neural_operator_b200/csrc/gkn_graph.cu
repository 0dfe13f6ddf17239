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
    // one block per segment (blockIdx.x), segment length n; writes n+1 values
    in += int64_t(blockIdx.x) * in_stride;
    out += int64_t(blockIdx.x) * out_stride;
    __shared__ TOut warp_tot[32];
    __shared__ TOut carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += kScanThreads) {
        const int64_t i = base + threadIdx.x;
        const TOut v = i < n ? TOut(in[i]) : TOut(0);
        TOut incl = v;
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
        const TOut warp_excl = warp == 0 ? TOut(0) : warp_tot[warp - 1];
        if (i < n) out[i] = carry + warp_excl + incl - v;
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
    w.bytes = off;
    return w;
}

// one block per frame: bounding box, then pick a cell edge >= cutoff whose grid fits in C cells
__global__ void __launch_bounds__(256)
rg_frame_setup_kernel(const float* __restrict__ pos, int64_t N, double threshold, int64_t C,
                      FrameMeta* __restrict__ meta) {
    const int64_t f = blockIdx.x;
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
    __shared__ float smn[3][256], smx[3][256];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        smn[d][threadIdx.x] = mn[d];
        smx[d][threadIdx.x] = mx[d];
    }
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
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
        // cell edge strictly above the cutoff so that |d| < threshold implies adjacent cells
        double cell = fmax(threshold, 1e-30) * 1.0001;
        int64_t nx, ny, nz;
        for (;;) {
            nx = int64_t(floor(ext[0] / cell)) + 1;
            ny = int64_t(floor(ext[1] / cell)) + 1;
            nz = int64_t(floor(ext[2] / cell)) + 1;
            if (nx * ny * nz <= C) break;
            cell *= 1.26;
        }
        if (N == 0) nx = ny = nz = 1;
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
        const int slot = cell_start[f * (C + 1) + c] + atomicAdd(cursor + f * (C + 1) + c, 1);
        sorted_pos[f * N + slot] =
            make_float4(pos[g * 3], pos[g * 3 + 1], pos[g * 3 + 2], __int_as_float(int(g - f * N)));
    }
}

// exact reference predicate: float32 -> float64 promotion, ((dx^2 + dy^2) + dz^2), sqrt, strict <
__device__ __forceinline__ bool within(double xi, double yi, double zi, float4 pj, double thr) {
    const double dx = __dsub_rn(double(pj.x), xi);
    const double dy = __dsub_rn(double(pj.y), yi);
    const double dz = __dsub_rn(double(pj.z), zi);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return __dsqrt_rn(d2) < thr;
}

// One warp per atom.  FILL=false: count neighbours.  FILL=true: write the sorted row.
template <bool FILL>
__global__ void __launch_bounds__(256)
rg_neighbor_kernel(const float* __restrict__ pos, int64_t F, int64_t N, int64_t C, double thr,
                   const FrameMeta* __restrict__ meta, const int32_t* __restrict__ cell_start,
                   const float4* __restrict__ sorted_pos, int32_t* __restrict__ deg,
                   const int64_t* __restrict__ row_ptr, int64_t E, int64_t* __restrict__ edge_index,
                   float* __restrict__ edge_attr) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t total = F * N;
    for (int64_t g = warp0; g < total; g += nwarps) {
        const int64_t f = g / N;
        const FrameMeta m = meta[f];
        const float px = pos[g * 3], py = pos[g * 3 + 1], pz = pos[g * 3 + 2];
        const double xi = px, yi = py, zi = pz;
        int cx, cy, cz;
        cell_of(m, px, py, pz, cx, cy, cz);
        const int32_t* cs = cell_start + f * (C + 1);
        const float4* sp = sorted_pos + f * N;
        const int64_t beg = FILL ? row_ptr[g] : 0;
        int64_t* scratch = FILL ? edge_index + beg : nullptr;  // row-id slots double as scratch
        int count = 0;
        for (int ax = max(cx - 1, 0); ax <= min(cx + 1, m.nx - 1); ++ax) {
            for (int ay = max(cy - 1, 0); ay <= min(cy + 1, m.ny - 1); ++ay) {
                // cells (ax, ay, z0..z1) are contiguous in the cell-sorted array
                const int z0 = max(cz - 1, 0), z1 = min(cz + 1, m.nz - 1);
                const int c0 = (ax * m.ny + ay) * m.nz + z0;
                const int c1 = (ax * m.ny + ay) * m.nz + z1;
                const int jb = cs[c0], je = cs[c1 + 1];
                for (int base = jb; base < je; base += 32) {  // warp-uniform trip count
                    const int j = base + lane;
                    bool hit = false;
                    float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < je) {
                        pj = sp[j];
                        hit = within(xi, yi, zi, pj, thr);
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, hit);
                    if (FILL && hit) {
                        const int slot = count + __popc(mask & ((1u << lane) - 1u));
                        scratch[slot] = int64_t(__float_as_int(pj.w));
                    }
                    count += __popc(mask);
                }
            }
        }
        if (!FILL) {
            if (lane == 0) deg[g] = count;
        } else {
            __syncwarp();
            // rank-count sort of the (distinct) neighbour ids into the column slots
            int64_t* cols = edge_index + E + beg;
            const int64_t node0 = f * N;
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
    }
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
    return (2 * N + 2 * E + 64) * int64_t(sizeof(int32_t));
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
    cudaError_t err = cudaMemsetAsync(cnt_dst, 0, size_t(2 * N) * 4, st);
    if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(prepare)", err);
    err = cudaMemsetAsync(status, 0, 4, st);
    if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(status)", err);
    const int eg = grid_for(E, 256, 8);
    const int ng = grid_for(N * 32, 256, 8);
    prep_count_kernel<<<eg, 256, 0, st>>>(edge_index, N, E, cnt_dst, cnt_src, status);
    GKN_CHECK_LAUNCH("prep_count");
    exclusive_scan_kernel<int32_t, int32_t><<<1, kScanThreads, 0, st>>>(cnt_dst, dst_ptr, N, 0, 0);
    GKN_CHECK_LAUNCH("prep_scan_dst");
    exclusive_scan_kernel<int32_t, int32_t><<<1, kScanThreads, 0, st>>>(cnt_src, src_ptr, N, 0, 0);
    GKN_CHECK_LAUNCH("prep_scan_src");
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
    if (N >= (int64_t(1) << 30)) return set_error("gkn_radius_graph_count: N too large");
    RadiusWs w = carve_radius_ws(workspace, F, N);
    if (workspace_bytes < w.bytes) return set_error("gkn_radius_graph_count: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t total = F * N;
    if (total == 0) {
        cudaError_t err = cudaMemsetAsync(row_ptr, 0, 8, st);
        if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(row_ptr)", err);
        return 0;
    }
    cudaError_t err = cudaMemsetAsync(w.cell_count, 0, size_t(F) * (w.C + 1) * 4, st);
    if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(cell_count)", err);
    rg_frame_setup_kernel<<<unsigned(F), 256, 0, st>>>(pos, N, threshold, w.C, w.meta);
    GKN_CHECK_LAUNCH("rg_frame_setup");
    const int ag = grid_for(total, 256, 8);
    rg_assign_kernel<<<ag, 256, 0, st>>>(pos, F, N, w.C, w.meta, w.atom_cell, w.cell_count);
    GKN_CHECK_LAUNCH("rg_assign");
    exclusive_scan_kernel<int32_t, int32_t>
        <<<unsigned(F), kScanThreads, 0, st>>>(w.cell_count, w.cell_start, w.C, w.C + 1, w.C + 1);
    GKN_CHECK_LAUNCH("rg_scan_cells");
    err = cudaMemsetAsync(w.cell_count, 0, size_t(F) * (w.C + 1) * 4, st);
    if (err != cudaSuccess) return set_cuda_error("cudaMemsetAsync(cell cursor)", err);
    rg_place_kernel<<<ag, 256, 0, st>>>(pos, F, N, w.C, w.atom_cell, w.cell_start, w.cell_count,
                                        w.sorted_pos);
    GKN_CHECK_LAUNCH("rg_place");
    const int wg = grid_for(total * 32, 256, 8);
    rg_neighbor_kernel<false><<<wg, 256, 0, st>>>(pos, F, N, w.C, threshold, w.meta, w.cell_start,
                                                  w.sorted_pos, deg, nullptr, 0, nullptr, nullptr);
    GKN_CHECK_LAUNCH("rg_count");
    exclusive_scan_kernel<int32_t, int64_t><<<1, kScanThreads, 0, st>>>(deg, row_ptr, total, 0, 0);
    GKN_CHECK_LAUNCH("rg_scan_rows");
    return 0;
}

extern "C" int gkn_radius_graph_fill(const float* pos, int64_t F, int64_t N, double threshold,
                                     const int64_t* row_ptr, int64_t E, int64_t* edge_index,
                                     float* edge_attr, void* workspace, int64_t workspace_bytes,
                                     void* stream) {
    using namespace gkn;
    if (F < 0 || N < 0 || E < 0) return set_error("gkn_radius_graph_fill: negative size");
    RadiusWs w = carve_radius_ws(workspace, F, N);
    if (workspace_bytes < w.bytes) return set_error("gkn_radius_graph_fill: workspace too small");
    const int64_t total = F * N;
    if (total == 0 || E == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int wg = grid_for(total * 32, 256, 8);
    rg_neighbor_kernel<true><<<wg, 256, 0, st>>>(pos, F, N, w.C, threshold, w.meta, w.cell_start,
                                                 w.sorted_pos, nullptr, row_ptr, E, edge_index,
                                                 edge_attr);
    GKN_CHECK_LAUNCH("rg_fill");
    return 0;
}
