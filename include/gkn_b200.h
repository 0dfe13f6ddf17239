/*
 * gkn_b200.h — C ABI of libgkn_b200.so, the B200 (sm_100a) implementation of the graph-kernel
 * neural-operator hot path of ramanathanlab/molecular_dynamics_neural_operator.
 *
 * The reference has no FFI of its own: its boundary is the PyTorch nn.Module surface
 * (NNConv_old.forward graph_kernel.py:194-198, KernelNN.forward :277-309, construct_pairdata
 * :362-393).  These entry points are what a binding for that path calls underneath that surface;
 * INTEGRATION.md shows the ctypes stub.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name says host;
 *   - all tensors are dense row-major; "bf16" buffers hold raw 16-bit bfloat16 values;
 *   - no allocation inside: the caller owns inputs, outputs and workspaces;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); calls are asynchronous,
 *     re-entrant and keep no global mutable state besides caches keyed by device ordinal;
 *   - return value 0 = success; non-zero = error, message via gkn_last_error() (thread-local);
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Edge layout ("destination-sorted CSR"): gkn_graph_prepare turns an arbitrary edge_index[2,E]
 * into positions p = 0..E-1 ordered by destination node (ties: original edge order), with
 * dst_ptr[N+1] row pointers.  All per-edge buffers below (h, msg, dxj, S) use that order.
 */
#ifndef GKN_B200_H_
#define GKN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GKN_B200_VERSION 100 /* major*100 + minor */
#define GKN_ERR_INVALID 1
#define GKN_ERR_CUDA 2

/* ---- housekeeping ---------------------------------------------------------------------- */
int gkn_version(void);
const char* gkn_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long gkn_launch_count(void);
int gkn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- graph metadata (replaces PyG propagate's implicit gather/scatter bookkeeping;
 *      reference graph_kernel.py:198, dataset.py:41-45) ---------------------------------- */
/* bytes of scratch gkn_graph_prepare needs for (N, E) */
int64_t gkn_graph_prepare_workspace_bytes(int64_t N, int64_t E);
/* edge_index: int64 [2,E] (row 0 = source j, row 1 = destination i, PyG source_to_target flow).
 * Outputs (all int32 unless noted): perm[E] position -> original edge id; src_s[E], dst_s[E]
 * endpoints in sorted order; dst_ptr[N+1]; inv_cnt float[N] = 1/max(in-degree,1);
 * src_ptr[N+1] and src_perm[E]: sorted positions grouped by source node (ascending p).
 * status: int32[1], set to non-zero on the device if an index is outside [0,N). */
int gkn_graph_prepare(const int64_t* edge_index, int64_t N, int64_t E, int32_t* perm,
                      int32_t* src_s, int32_t* dst_s, int32_t* dst_ptr, float* inv_cnt,
                      int32_t* src_ptr, int32_t* src_perm, int32_t* status, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* Fast path for symmetric graphs already in (row, col) order with CSR row pointers - exactly what
 * gkn_radius_graph_fill emits.  No sorting: position p <-> (dst = row[p], src = col[p]);
 * rev[p] = position of the reverse edge, which is at once `perm` and `src_perm` of the general
 * plan; ptr32 = row_ptr as int32 (= dst_ptr = src_ptr).  status != 0: graph was not symmetric. */
int gkn_plan_from_csr(const int64_t* edge_index, const int64_t* row_ptr, int64_t N, int64_t E,
                      int32_t* rev, int32_t* src_s, int32_t* dst_s, int32_t* ptr32, float* inv_cnt,
                      int32_t* status, void* stream);

/* ---- radius graph (replaces construct_pairdata, reference graph_kernel.py:362-393:
 *      scipy distance_matrix(float64) < threshold -> coo_matrix -> per-edge python loop) -- */
int64_t gkn_radius_graph_workspace_bytes(int64_t F, int64_t N);
/* Pass 1: pos float32 [F,N,3]; threshold as the float64 the reference compares against.
 * Writes deg int32 [F*N] (neighbours of each atom inside its own frame, self included) and
 * row_ptr int64 [F*N+1] (exclusive scan of deg); total edges = row_ptr[F*N]. */
int gkn_radius_graph_count(const float* pos, int64_t F, int64_t N, double threshold, int32_t* deg,
                           int64_t* row_ptr, void* workspace, int64_t workspace_bytes,
                           void* stream);
/* Pass 2 (same workspace, untouched since pass 1): edge_index int64 [2,E] with node ids offset by
 * f*N (PyG batching, dataset.py:41-45), lexicographic (row, col) order; edge_attr float32 [E,6] =
 * [pos[row], pos[col]].  E must equal row_ptr[F*N]. */
int gkn_radius_graph_fill(const float* pos, int64_t F, int64_t N, double threshold,
                          const int64_t* row_ptr, int64_t E, int64_t* edge_index, float* edge_attr,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* ---- fused bilinear edge contraction (replaces NNConv_old.message, graph_kernel.py:200-202)
 *   P[e, i*W+o] = sum_k h[e,k] * w3[i*W+o, k]  is never written to memory.
 *   fwd: msg[e,o] = sum_i x[src[e], i] * P[e, i*W+o]
 *   dx : dxj[e,i] = sum_o P[e, i*W+o] * g[dst[e], o]
 * h [E,K], w3 [W*W,K] (= net.layers.4.weight); x, g [N,W] float32; msg, dxj [E,W] float32.
 * bf16: tcgen05/TMEM tensor-core path (bf16 operands, fp32 accumulate), K % 64 == 0,
 *       W in {32,64,128}.  f32: CUDA-core path with fp32 operands, any W % 16 == 0, K % 16 == 0. */
int gkn_contract_fwd_bf16(const void* h_bf16, const void* w3_bf16, const int32_t* src,
                          const float* x, float* msg, int64_t E, int32_t K, int32_t W,
                          void* stream);
int gkn_contract_dx_bf16(const void* h_bf16, const void* w3_bf16, const int32_t* dst,
                         const float* g, float* dxj, int64_t E, int32_t K, int32_t W,
                         void* stream);
int gkn_contract_fwd_f32(const float* h, const float* w3, const int32_t* src, const float* x,
                         float* msg, int64_t E, int32_t K, int32_t W, void* stream);
int gkn_contract_dx_f32(const float* h, const float* w3, const int32_t* dst, const float* g,
                        float* dxj, int64_t E, int32_t K, int32_t W, void* stream);

/* ---- node-first factorisation of the same contraction (tcgen05 grouped GEMMs) --------------
 * NNConv_old.message (graph_kernel.py:200-202) is trilinear in (x[src e], h[e], W3); because x is a
 * node quantity, W3 can be contracted with it once per NODE (a plain dense GEMM done by the
 * caller: U[n][o,k] = sum_i x[n,i] w3[i*W+o,k]) and the per-edge work drops from 2*K*W*W to 2*K*W
 * FLOP.  Edge buffers use the "padded source-sorted" layout: the edges of source node n occupy
 * rows [pptr[n], pptr[n]+deg[n]), groups padded to a multiple of 16 rows (E_pad rows in total).
 * `units`: int32 [U,4] = (source node, first row, rows to store (<= 128), 0), one per 128-row
 * tile of a group, ordered by node.  K % 256 == 0; W in {32, 64, 128} for msg, gather and the
 * *_mn forms of du / dh, W = 64 for gkn_node_du_bf16 / gkn_node_dh_bf16 and the transposed gather
 * output.  Node-indexed operands may be passed in
 * chunks of nodes [node_base, node_base + chunk_nodes) covering the units of the launch.
 *   msg : msg[q,o] = sum_k hs[q,k] * u[src q][o,k]          hs bf16 [E_pad,K], u bf16 [chunk,W,K],
 *         msg float32 [E_pad,W] (forward message, before the mean)
 *   du  : dut[n][k,o] = sum_{q in group n} hsT[k,q] * gsT[o,q]      hsT bf16 [K,E_pad] (= hs^T),
 *         gsT bf16 [W,E_pad] (message gradient per edge, transposed, zero in pad rows),
 *         dut bf16 [node_end-node_begin, K, W] (gradient w.r.t. U, [k,o] order)
 *   dh  : dh[q,k] = sum_a sum_o gs[a][q,o] * ut[a][src q][k,o]    gs bf16 [A,E_pad,W],
 *         ut bf16 [A,chunk,K,W], dh bf16 [E_pad,K] (gradient of the edge hidden state).  Its
 *         `units` table is int32 [U*(K/256),4] = (source node, first row, rows, 256-column chunk)
 *         in (node, chunk, tile) order, so the tiles of a group that share a chunk of the node's U
 *         are adjacent (L2 reuse); unit_begin/unit_end count tiles as in msg.
 *   gather: gs[q,:W] = bf16(g[idx[q],:]) (zero where idx[q] < 0), gsT its transpose; either output
 *         may be NULL.  gs rows hold max(W,64) values (one 128-byte row at least: at W = 32 the
 *         upper 32 columns are written as zeros); the *_mn kernels read gs with that row length. */
int gkn_node_msg_bf16(const void* hs_bf16, const void* u_bf16, const int32_t* units,
                      int64_t unit_begin, int64_t unit_end, int64_t E_pad, int64_t chunk_nodes,
                      int32_t node_base, float* msg, int32_t K, int32_t W, void* stream);
int gkn_node_du_bf16(const void* hsT_bf16, const void* gsT_bf16, const int32_t* pptr,
                     int64_t node_begin, int64_t node_end, int64_t E_pad, void* dut_bf16,
                     int32_t K, int32_t W, void* stream);
int gkn_node_dh_bf16(const void* gs_bf16, const void* ut_bf16, const int32_t* units,
                     int64_t unit_begin, int64_t unit_end, int64_t E_pad, int64_t chunk_nodes,
                     int32_t node_base, int32_t A, void* dh_bf16, int32_t K, int32_t W,
                     void* stream);
/* The same two products with operands read in place as MN-major tensor-core operands: du from
 * hs [E_pad,K] and gs [E_pad,max(W,64)] (no transposed copies; dut [nodes,K,W]), dh from
 * gs [A,E_pad,max(W,64)] and u [A,chunk,W,K] (the forward's [o,k]-ordered U, so it can be saved
 * instead of recomputed in [k,o] order).  W in {32, 64, 128}: this is the backward of
 * graph_kernel.py:200-202 for every width the reference's configurations use. */
int gkn_node_du_mn_bf16(const void* hs_bf16, const void* gs_bf16, const int32_t* pptr,
                        int64_t node_begin, int64_t node_end, int64_t E_pad, void* dut_bf16,
                        int32_t K, int32_t W, void* stream);
int gkn_node_dh_mn_bf16(const void* gs_bf16, const void* u_bf16, const int32_t* units,
                        int64_t unit_begin, int64_t unit_end, int64_t E_pad, int64_t chunk_nodes,
                        int32_t node_base, int32_t A, void* dh_bf16, int32_t K, int32_t W,
                        void* stream);
/* gkn_node_dh_mn_bf16 with the backward of DenseNet layers.2's ReLU (graph_kernel.py:239-242) fused
 * into its epilogue: dz[q,k] = dh[q,k] where the layer's output hs[q,k] is positive, zero
 * elsewhere, and colsum[k] (+)= sum_q dz[q,k] over the rows of the launch (the gradient of
 * layers.2.bias; accumulate != 0 adds to colsum, which is how node-range launches are chained).
 * The activation pattern comes as one bit per element: act_bits uint32 [E_pad, K/32], bit k%32 of
 * word k/32 set where hs[q,k] > 0 - written once per step by gkn_relu_bits_bf16 (hs bf16
 * [rows, K], K % 32 == 0), so the epilogue reads 32 bytes per row instead of 512.  The sums are
 * taken per CTA in a fixed order and folded in CTA order (no atomics; deterministic).
 * workspace: float [SM count, K] scratch.  K <= 1024. */
int gkn_relu_bits_bf16(const void* act_bf16, int64_t rows, int32_t K, uint32_t* bits, void* stream);
int gkn_node_dh_mn_relu_bf16(const void* gs_bf16, const void* u_bf16, const int32_t* units,
                             int64_t unit_begin, int64_t unit_end, int64_t E_pad,
                             int64_t chunk_nodes, int32_t node_base, int32_t A,
                             const uint32_t* act_bits, void* dz_bf16, float* colsum,
                             int32_t accumulate, void* workspace, int64_t workspace_bytes,
                             int32_t K, int32_t W, void* stream);
int gkn_node_gather_bf16(const float* g, const int32_t* idx, int64_t E_pad, void* gs_bf16,
                         void* gsT_bf16, int32_t W, void* stream);

/* ---- node-first factorisation, fused producer/consumer (tcgen05) ------------------------------
 * One persistent kernel per NNConv application that BOTH produces U = x @ W3r with the tensor
 * core and consumes it in the grouped message GEMM (NNConv_old.message, graph_kernel.py:200-202)
 * while it is still in the L2 - U never takes the HBM round trip of the two-kernel form above.
 * The kernel executes a host-ordered list of work items: item v belongs to CTA v % grid (grid <=
 * the SM count; 0 = one CTA per SM), every CTA runs its items in list order; an item is 8 int32:
 *     (kind, r0, r1, r2,  wait_idx, wait_target, signal_idx, r3)
 *   kind 0 "P": U[r2 + i, r1 .. r1+256) = x[r0 + i, :] @ W3r[:, r1 .. r1+256) for i < r3 (<= 128)
 *   kind 1 "C": msg[r0 + i, :] = hs[r0 + i, :] @ U[r2]^T for i < r1 (<= 128)   (U[r2]: [W, K])
 *   kind 2    : no-op
 *   wait_idx >= 0: the item waits until flags[wait_idx] >= wait_target (C: before it reads U;
 *   P: before it overwrites its rows of U - the "ring" use, where U holds a few node blocks only);
 *   signal_idx >= 0: each of the 4 epilogue warps adds 1 to flags[signal_idx] when the item's
 *   output is in memory.  The waits must be acyclic together with the per-CTA order (e.g. every
 *   wait refers to items with a smaller time key and every CTA's items are in key order); all
 *   CTAs are co-resident, so a waiting CTA only ever waits for CTAs that are running.
 * x bf16 [N, max(W,64)] (zero padded), w3r bf16 [max(W,64), W*K] = layers.4.weight viewed
 * [i, (o,k)] (zero rows beyond W), hs bf16 [E_pad, K], u bf16 [u_rows, W*K], msg fp32 [E_pad, W],
 * flags int32 (zeroed by the caller before the launch).  hs and msg may be NULL when E_pad = 0
 * (P items only: a plain dense GEMM).  W in {32, 64, 128}, K % 256 == 0. */
int gkn_node_fused_fwd_bf16(const void* x_bf16, const void* w3r_bf16, const void* hs_bf16,
                            void* u_bf16, const int32_t* items, int64_t num_items, int32_t* flags,
                            float* msg, int64_t N, int64_t E_pad, int64_t u_rows, int32_t K,
                            int32_t W, int32_t grid, void* stream);

/* Fused backward of one application (w = 64): the gradient w.r.t. U is produced by "D" items
 * (dU[n][k,o] = sum_{e: src e = n} hs[e,k] * gs[e,o], as gkn_node_du_mn_bf16) into a ring of node
 * blocks and consumed from the L2 by "X" items that feed BOTH
 *     dwT[(k,o), i] (+)= sum_n dU[n][(k,o)] * x[n, i]           (gradient of layers.4.weight,
 *                                                                 accumulated in TMEM for the launch)
 *     dx_part[c][n, i] = sum_{(k,o) in slice(c)} dU[n][(k,o)] * w3k[i, (k,o)]
 * from one read of dU; CTA c owns a fixed slice of the w*K columns, gkn_sum_partials adds the
 * `parts` partial dx tensors in a fixed order.  Items (8 int32, as above):
 *   kind 0 "D": (0, node, chunk of 128 k-rows, row of the dU buffer, wait_idx, target, signal_idx, 0)
 *   kind 1 "X": (1, first node of the block, first dU row of the block, valid rows <= 128, ...)
 *               must sit at list positions == CTA index (mod grid), one per CTA and block
 * gkn_node_fused_bwd_grid returns the grid size the X items must be laid out for (one CTA per SM)
 * and the number of dx partials.  hs bf16 [E_pad,K]; gs bf16 [E_pad,64] (zero pad rows); pptr int32
 * [N+1]; x bf16 [N,64]; w3k bf16 [64, K*64] = layers.4.weight as [i, (k,o)]; du bf16
 * [du_rows, K*64] (zero-initialised once by the caller); dx_part fp32 [parts,N,64] or NULL (dx not
 * needed); dwT fp32 [K*64, 64] or NULL; accumulate_dw: add to dwT instead of overwriting it. */
int gkn_node_fused_bwd_grid(int32_t K, int32_t W, int32_t* grid, int32_t* parts);
int gkn_node_fused_bwd_bf16(const void* hs_bf16, const void* gs_bf16, const int32_t* pptr,
                            const void* x_bf16, const void* w3k_bf16, void* du_bf16,
                            const int32_t* items, int64_t num_items, int32_t* flags,
                            float* dx_part, float* dwT, int32_t accumulate_dw, int64_t N,
                            int64_t E_pad, int64_t du_rows, int32_t K, int32_t W, void* stream);
/* out[j] = (addend ? addend[j] : 0) + sum_{c < parts} part[c*n + j], summed in order of c
 * (n % 4 == 0; out may alias addend). */
int gkn_sum_partials(const float* part, int32_t parts, int64_t n, const float* addend, float* out,
                     void* stream);

/* ---- segmented mean / sums (replaces PyG aggregate(aggr="mean") + NNConv_old.update,
 *      graph_kernel.py:204-209, and their autograd) --------------------------------------- */
/* out[n,:] = act( inv_cnt[n] * sum_{p in [dst_ptr[n],dst_ptr[n+1])} (msg[p,:] + xb3[src_s[p],:])
 *                 + xroot[n,:] + bias[:] )     xb3, xroot, bias may be NULL; relu != 0 => ReLU. */
int gkn_scatter_mean_fwd(const float* msg, const int32_t* src_s, const int32_t* dst_ptr,
                         const float* inv_cnt, const float* xb3, const float* xroot,
                         const float* bias, float* out, int64_t N, int64_t E, int32_t W,
                         int32_t relu, void* stream);
/* dx[n,:] = addend[n,:] + sum_{q in [src_ptr[n],src_ptr[n+1])} (dxj[p,:] + gb3[dst_s[p],:]),
 * p = src_perm[q].  gb3, addend may be NULL. */
int gkn_scatter_src_sum(const float* dxj, const int32_t* src_ptr, const int32_t* src_perm,
                        const int32_t* dst_s, const float* gb3, const float* addend, float* dx,
                        int64_t N, int64_t E, int32_t W, void* stream);
/* Fused node update, one kernel for NNConv_old.update (graph_kernel.py:204-209, forward) and for
 * its autograd (backward):
 *   r[n]   = scale[n] * sum_{q in [seg_ptr[n],seg_ptr[n+1])} rows[p,:]      p = perm ? perm[q] : q
 *   agg[n] = scale[n] * sum_q gnode[gidx[p],:]                  (written to agg_out if not NULL)
 *   out[n] = act( r[n] + agg[n] @ M1 + nin[n] @ M2 + bias )     M1, M2: [W,W] row-major
 *   rows may be NULL (r = 0).  M1 = M2 = NULL (nin unused): plain segment sums, out = act(r + bias)
 *   - the partial sums of a source-range shard before its all-reduce (sharded.py).
 * forward : rows=msg, perm=NULL, seg_ptr=dst_ptr, gidx=src_s, gnode=x, scale=inv_cnt,
 *           M1 = layers.4.bias viewed [W,W], nin=x, M2=root, bias, relu
 * backward: rows=dxj, perm=src_perm, seg_ptr=src_ptr, gidx=dst_s, gnode=gn, scale=NULL,
 *           M1 = (b3 view)^T, nin=g, M2=root^T, bias=NULL, relu=0 */
int gkn_segment_matvec(const float* rows, const int32_t* perm, const int32_t* seg_ptr,
                       const int32_t* gidx, const float* gnode, const float* scale,
                       const float* M1, const float* nin, const float* M2, const float* bias,
                       float* out, float* agg_out, int64_t N, int32_t W, int32_t relu,
                       void* stream);
/* ---- sharded large-frame rollout: the exchange step fused with the node update ----------------
 * Source-range sharding (sharded.py) leaves rank p with partial sums S_p [2, N, W] fp32 (messages
 * and x_j) at every destination.  Rank r computes, for ITS nodes [n0, n1),
 *     x'[n] = act( inv_cnt[n] * sum_p S_p[0,n] + (inv_cnt[n] * sum_p S_p[1,n]) @ B3 + x[n] @ root + bias )
 * (PyG mean + NNConv_old.update, graph_kernel.py:194-209) reading the rows of every peer's S_p
 * through mapped peer pointers and writing x'[n] into every rank's x' buffer: reduce-scatter,
 * compute and all-gather in one kernel over NVLink.  partial_ptrs / xout_ptrs: HOST arrays of
 * `world` device addresses (this rank's own buffers included, in rank order; e.g. the buffer_ptrs
 * of a torch symmetric-memory handle).  mc_partial / mc_xout: the MULTICAST addresses of the two
 * symmetric buffers (NVLS), or 0: with them the sum over the ranks is one multimem.ld_reduce per 16
 * bytes, computed in the NVSwitch, and the new row is one multimem.st replicated by the switch.
 * The caller synchronises the ranks before (all S_p written) and after (all x' rows landed) the
 * launch.  W in {32, 64, 128}, world <= 16. */
int gkn_shard_reduce_update(const uint64_t* partial_ptrs, const uint64_t* xout_ptrs,
                            uint64_t mc_partial, uint64_t mc_xout, int32_t world, const float* x, const float* inv_cnt, const float* B3,
                            const float* root, const float* bias, int64_t n0, int64_t n1,
                            int64_t N, int32_t W, int32_t relu, void* stream);

/* Node encoder's LSTM (graph_kernel.py:278-285: nn.LSTM(3, 3) called window step by window step
 * with the state carried over, i.e. one sequence of T steps over R independent rows, zero initial
 * state).  fp32; PyTorch parameter layout: w_ih [4D, D], w_hh [4D, D], b_ih, b_hh [4D], gate rows in
 * (input, forget, cell, output) order; D = 3.
 *   fwd: x [T,R,D] -> out [T,R,D] (h_t); gates [T,R,4D] (activated) and cells [T,R,D] (c_t) are
 *        saved for the backward (both NULL: inference).
 *   bwd: dout [T,R,D] -> dx [T,R,D] (may be NULL) and dw_part [gkn_lstm_blocks(R), 4D*D*2 + 4D]:
 *        per-block sums of (dW_ih, dW_hh, db) - the caller adds the blocks; db is the gradient of
 *        both bias vectors.  One thread per row, fixed summation order, no atomics. */
int gkn_lstm_blocks(int64_t R);
int gkn_lstm_fwd(const float* x, const float* w_ih, const float* w_hh, const float* b_ih,
                 const float* b_hh, float* out, float* gates, float* cells, int64_t T, int64_t R,
                 int32_t D, void* stream);
int gkn_lstm_bwd(const float* x, const float* w_ih, const float* w_hh, const float* out,
                 const float* gates, const float* cells, const float* dout, float* dx,
                 float* dw_part, int64_t T, int64_t R, int32_t D, void* stream);

/* Dataset feeder (reference dataset.py:194-201, a Python loop per edge there):
 * edge_attr[e, :] = [pos[edge_index[0,e], :], pos[edge_index[1,e], :]] for a batch of graphs whose
 * node ids index the rows of pos (float32 [M,3]; edge_index int64 [2,E]; edge_attr float32 [E,6]).
 * status (int32[1], may be NULL) gets bit 0 set if a node id lies outside [0, M). */
int gkn_edge_attr_gather(const float* pos, const int64_t* edge_index, int64_t E, int64_t M,
                         float* edge_attr, int32_t* status, void* stream);
/* out[i,o] = sum_r x[r,i] * g[r,o] (float32, [W,W] from rows x W operands) and, when gsum != NULL,
 * gsum[o] = sum_r g[r,o]: the gradients of NNConv_old.root / .bias (graph_kernel.py:204-209) over
 * the node rows of all applications stacked, and of layers.4.bias (agg^T G).  Two launches, per-CTA
 * partial sums folded in CTA order (deterministic).  W in {16, 32, 64, 128};
 * workspace >= gkn_xty_workspace_bytes(W). */
int64_t gkn_xty_workspace_bytes(int32_t W);
int gkn_xty_f32(const float* x, const float* g, int64_t rows, int32_t W, float* out, float* gsum,
                void* workspace, int64_t workspace_bytes, void* stream);
/* g[n,:] = gy[n,:] * (relu ? (y[n,:] > 0) : 1);  gn[n,:] = g[n,:] * inv_cnt[n] */
int gkn_relu_grad_scale(const float* gy, const float* y, const float* inv_cnt, float* g, float* gn,
                        int64_t N, int32_t W, int32_t relu, void* stream);

/* ---- deferred weight-gradient operand --------------------------------------------------
 * S[e-e0, i*W+o] = sum_a xs[a, src_s[e], i] * gs[a, dst_s[e], o]   for e in [e0, e1)
 * (the gradient of the loss w.r.t. the per-edge kernel, summed over the A applications that share
 * it, graph_kernel.py:271-273, :299-302).  xs, gs: float32 [A,N,W]; S: bf16 or float32
 * [e1-e0, W*W].  Used by the staged (chunked) backward; the fused backward never stores S. */
int gkn_edge_outer_sum(const float* xs, const float* gs, const int32_t* src_s,
                       const int32_t* dst_s, void* S, int32_t out_is_bf16, int64_t e0, int64_t e1,
                       int64_t N, int32_t A, int32_t W, void* stream);

/* ---- fused gradient of the edge hidden state (tcgen05) ------------------------------------
 * dh[e, k] = sum_n S[e, n] * w3[n, k]  with S as above, synthesised tile by tile in shared memory
 * while the tensor core runs: S never reaches HBM (the autograd of NNConv_old.message,
 * graph_kernel.py:200-202, through the shared last edge-MLP layer, :271-273).
 * xs, gs: float32 [A, N, W] (A <= 16); src_s, dst_s: sorted-order endpoints (dst_s non-decreasing);
 * w3t: bf16 [K, W*W] = net.layers.4.weight transposed; dh: bf16 [E, K] (fp32 accumulation, one
 * rounding).  W = 64, K % 256 == 0.  workspace: gkn_edge_grad_workspace_bytes(N, W) bytes. */
int64_t gkn_edge_grad_workspace_bytes(int64_t N, int32_t W);
/* status[0] = max over 128-edge tiles of the number of distinct destinations in the tile; the fused
 * kernels below require it to be <= 8 (graphs of lower degree use the staged path). */
int gkn_edge_grad_max_slots(const int32_t* dst_s, int64_t E, int32_t* status, void* stream);
int gkn_edge_grad_dh_bf16(const float* xs, const float* gs, const int32_t* src_s,
                          const int32_t* dst_s, const void* w3t_bf16, void* dh_bf16, int64_t E,
                          int64_t N, int32_t A, int32_t K, int32_t W, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* ---- casts ------------------------------------------------------------------------------ */
/* out_bf16[i] = bf16_rn( relu ? max(in[i] + (bias ? bias[i % cols] : 0), 0) : in[i] + bias ) */
int gkn_bias_act_cast_bf16(const float* in, const float* bias, void* out_bf16, int64_t rows,
                           int64_t cols, int32_t relu, void* stream);

/* ---- edge-MLP layer .0 and the backward of "bias + ReLU" ----------------------------------
 * gkn_edge_layer0: out[e,k] = relu(sum_j ea[e,j]*w1[k,j] + b1[k]), bf16 or fp32 (DenseNet
 * layers.0 + ReLU, graph_kernel.py:239-242; kin must be 6 = edge_features, :328).
 * gkn_mask_cast_colsum: out[r,c] = dz[r,c] * (act[r,c] > 0); dz fp32 or bf16 (bf16 dz needs bf16
 * act/out), out cast to bf16 or kept fp32 (may alias dz when dtypes match); colsum[0,c] (+)= sum_r out[r,c] (bias gradient) and, when side != NULL
 * (nside = 6), colsum[1+j,c] (+)= sum_r out[r,c]*side[r,j] (layer-.0 weight gradient, transposed).
 * out may be NULL when only the sums are needed.
 * colsum is [(1+nside), cols]; deterministic two-pass reduction, no atomics.  act is the layer's
 * saved post-ReLU activation (bf16 or fp32). */
int gkn_edge_layer0(const float* ea, const float* w1, const float* b1, void* out,
                    int32_t out_is_bf16, int64_t E, int32_t K, int32_t kin, void* stream);
int64_t gkn_mask_cast_colsum_workspace_bytes(int64_t cols, int32_t nside);
int gkn_mask_cast_colsum(const void* dz, int32_t dz_is_bf16, const void* act, int32_t act_is_bf16,
                         void* out, int32_t out_is_bf16, const float* side, int32_t nside,
                         float* colsum, int32_t accumulate, int64_t rows, int64_t cols,
                         void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* GKN_B200_H_ */
