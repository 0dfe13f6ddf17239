"""Autograd glue for the graph-kernel layer: ONE ``torch.autograd.Function`` for the edge MLP plus
the whole stack of NNConv applications that share it (reference graph_kernel.py:271-273, :299-302).

Two implementations of the same function sit behind ``gkn_stack``:
``_GKNStackNode`` (default in bf16 mode at w = 32, 64 and 128) contracts
W3 with the NODE features first and does the per-edge work as tcgen05 grouped GEMMs - see the
section "Node-first factorisation" below and DESIGN.md section 1b; ``_GKNStack`` is the edge-first
fused contraction described next (``GKN_FACTORIZATION=edge``, fp32 mode, other shapes).

Why one Function and not one per layer: the edge hidden state ``h_e`` (output of net.layers.0-.3),
the last-layer weight W3/b3 and the graph are identical for all 2*depth applications, so

* forward computes ``h_e`` once; each application is then two kernels: the fused tcgen05
  contraction (``msg``) and the segmented mean (+ root, bias, ReLU);
* backward recomputes the per-edge kernel flash-style for ``dx`` (one contraction per application)
  and DEFERS the gradients of the shared operands: with
      S_e[i,o] = sum_a x^(a-1)[src e, i] * gn^(a)[dst e, o]
  (the loss gradient w.r.t. the per-edge kernel, summed over applications) a single pass gives
      dW3 = S^T h_e,   db3 = sum_e S_e,   dh_e = S W3
  i.e. 2 GEMM units instead of 2 per application, and ``dh_e`` is consumed chunk by chunk by the
  backward of layers .0-.3 without ever being stored for all edges.

Everything E-sized that is kept for backward is ``h_e`` and the layer-0 activation (bf16 in bf16
mode); per application only the node tensors ``x^(a)`` ([N, w]) are saved.

Precision modes
  "bf16": layer 0 in fp32 (fan-in 6); layer 2, the fused contraction and the weight-gradient
          GEMMs take bf16 operands and accumulate in fp32 (fp32 outputs, ``out_dtype``).
  "fp32": everything in fp32 on CUDA cores / cuBLAS SGEMM (parity mode, 1e-5).
The dense GEMMs of layers .0-.3 and of the staged weight gradients are plain library GEMMs
(cuBLAS through torch); the graph-kernel contraction itself is libgkn_b200.so.
"""
from __future__ import annotations

import torch

from . import ops
from .ops import GraphPlan

import os

# Edges per chunk of the staged weight-gradient pass (see _edge_grads): S for one chunk is
# chunk * w*w bf16 values (134 MB at 16384 edges, w = 64).  Measured on B200 (bench workload,
# ms/step): 16384 -> 72.9, 8192 -> 88.2, 4096 -> 84.3, 2048 -> 91.3; smaller chunks lose more in
# GEMM efficiency and dW3 accumulate traffic than they gain in L2 residency.
_S_CHUNK_EDGES = int(os.environ.get("GKN_S_CHUNK_EDGES", "16384"))

# Experimental: dh_e from the fused tcgen05 kernel (gkn_edge_grad_dh_bf16: S synthesised by the
# tensor core, never stored) instead of the staged S @ W3 library GEMM.  Bit-for-bit the same
# operands and accumulation precision, but as measured on B200 it is 2-2.6x SLOWER than the staged
# GEMM (hand-off latency between its five pipeline roles), so it is off by
# default; dW3 still needs the staged S either way.
FUSED_DH = os.environ.get("GKN_FUSED_DH", "0") == "1"


def _mm32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a @ b with an fp32 result (fp32 accumulation for bf16 operands)."""
    if a.dtype == torch.bfloat16:
        return torch.mm(a, b, out_dtype=torch.float32)
    return torch.mm(a, b)


def _addmm32(acc: torch.Tensor, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """acc + a @ b with fp32 accumulation/result; the addition rides in the GEMM epilogue."""
    if a.dtype == torch.bfloat16:
        return torch.addmm(acc, a, b, out_dtype=torch.float32)
    return acc.addmm_(a, b)


def _edge_hidden(ea, w1, b1, w2, b2, precision):
    """h1 = relu(ea W1^T + b1), h = relu(h1 W2^T + b2)   (reference graph_kernel.py:239-242).
    Layer .0 (fan-in 6) is one streaming kernel that writes the operand dtype of layer .2
    directly; layer .2 is a library GEMM with the bias + ReLU epilogue fused (cuBLASLt)."""
    if precision == "bf16":
        h1 = ops.edge_layer0(ea, w1.contiguous(), b1, torch.bfloat16)
        h = torch._addmm_activation(b2.to(torch.bfloat16), h1, w2.to(torch.bfloat16).t())
    else:
        h1 = ops.edge_layer0(ea, w1.contiguous(), b1, torch.float32)
        h = torch._addmm_activation(b2, h1, w2.t())
    return h1, h.contiguous()


def _edge_grads(xs, gns, plan: GraphPlan, ea, h1, h, w1, w2, w3q, need_ea: bool, precision: str):
    """Gradients of everything edge-sized from the per-application node tensors: W3 and, through
    dh_e, layers .0-.3 (W2, b2, W1, b1, optionally edge_attr).  (b3 is a node-level product.)

    Staged implementation: S is produced for a bounded chunk of edges at a time by
    gkn_edge_outer_sum and multiplied with library GEMMs; dh_e only ever exists for one chunk.
    """
    E, K = h.shape
    WW = w3q.shape[0]
    kin = ea.size(1)
    dev = h.device
    f32 = dict(dtype=torch.float32, device=dev)
    dw3 = torch.zeros((WW, K), **f32)
    dw2 = torch.zeros((K, K), **f32)
    db2 = torch.zeros(K, **f32)
    cs1 = torch.zeros((1 + kin, K), **f32)  # row 0: db1, rows 1..: dW1^T
    dea = torch.empty_like(ea) if need_ea else None
    bf16 = precision == "bf16"
    s_dtype = torch.bfloat16 if bf16 else torch.float32
    w2q = w2.to(torch.bfloat16) if bf16 else w2
    dh_all = None
    if (FUSED_DH and bf16 and xs.size(2) == 64 and xs.size(0) <= 16 and K % 256 == 0
            and ops.edge_grad_slots(plan) <= ops.EDGE_GRAD_MAX_SLOTS):
        dh_all = ops.edge_grad_dh(xs, gns, plan, w3q.t().contiguous())
    for e0 in range(0, E, _S_CHUNK_EDGES):
        e1 = min(E, e0 + _S_CHUNK_EDGES)
        hc, h1c = h[e0:e1], h1[e0:e1]
        S = ops.edge_outer_sum(xs, gns, plan, e0, e1, s_dtype)
        dw3 = _addmm32(dw3, S.t(), hc)
        # dh_e for this chunk.  bf16 mode: the GEMM rounds its fp32 accumulators to bf16 itself
        # (rounding before or after the 0/1 ReLU mask is the same), halving the traffic of the
        # mask pass below.
        dz2 = torch.mm(S, w3q) if dh_all is None else dh_all[e0:e1]
        del S
        # backward of "+ b2, ReLU": mask, bias gradient and operand hand-off in one pass
        dz2q = ops.mask_cast_colsum(dz2, hc, db2, s_dtype)
        dw2 = _addmm32(dw2, dz2q.t(), h1c)
        dz1 = _mm32(dz2q, w2q)
        # backward of "+ b1, ReLU" and of layer .0 itself: db1 and dW1 = dz1^T edge_attr
        dz1 = ops.mask_cast_colsum(dz1, h1c, cs1, torch.float32, side=ea[e0:e1],
                                   sums_only=not need_ea)
        if need_ea:
            torch.mm(dz1, w1, out=dea[e0:e1])
    return dw3, dw2, db2, cs1[1:].t().contiguous(), cs1[0].contiguous(), dea


class _GKNStack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, ea, w1, b1, w2, b2, w3, b3, root1, bias1, root2, bias2, plan, schedule,
                relu, precision):
        W = x0.size(1)
        det = lambda t: None if t is None else t.detach()  # noqa: E731
        x0, ea, w1, b1, w2, b2, w3, b3 = map(det, (x0, ea, w1, b1, w2, b2, w3, b3))
        h1, h = _edge_hidden(ea.contiguous(), w1, b1.contiguous(), w2, b2.contiguous(), precision)
        w3q = (w3.to(torch.bfloat16) if precision == "bf16" else w3).contiguous()
        B3 = b3.view(W, W)
        roots = (det(root1), det(root2))
        biases = (det(bias1), det(bias2))
        x = x0.contiguous()
        xs, aggs = [x], []
        want_agg = any(ctx.needs_input_grad)
        B3c = B3.contiguous()
        for c in schedule:
            msg = ops.contract_fwd(h, w3q, plan.src, x, precision)
            # mean over incoming edges + (mean x_j) @ B3 + x @ root + bias (+ ReLU), one kernel
            x_new, agg = ops.segment_matvec(msg, None, plan.dst_ptr, plan.src, x, plan.inv_cnt, B3c,
                                            x, roots[c].contiguous(), biases[c].contiguous(), relu,
                                            want_agg=want_agg)
            if want_agg:
                aggs.append(agg)
            x = x_new
            xs.append(x)
        ctx.plan, ctx.schedule, ctx.relu, ctx.precision = plan, schedule, relu, precision
        ctx.has2 = root2 is not None
        ctx.n_apps = len(schedule)
        ctx.save_for_backward(ea, h1, h, w1, w2, w3q, B3c, roots[0],
                              roots[1] if ctx.has2 else roots[0], *xs, *aggs)
        return x

    @staticmethod
    def backward(ctx, gout):
        ea, h1, h, w1, w2, w3q, B3, r1, r2, *rest = ctx.saved_tensors
        plan, schedule, relu, precision = ctx.plan, ctx.schedule, ctx.relu, ctx.precision
        A = ctx.n_apps
        xs, aggs = rest[:A + 1], rest[A + 1:]
        roots = (r1, r2)
        rootsT = (r1.t().contiguous(), r2.t().contiguous())
        B3T = B3.t().contiguous()
        W = xs[0].size(1)
        need = ctx.needs_input_grad
        need_x0, need_ea, need_edge = need[0], need[1], any(need[1:8])
        gs, gns = [None] * A, [None] * A
        gx = gout.contiguous().float()
        for a in range(A - 1, -1, -1):
            c = schedule[a]
            g, gn = ops.relu_grad_scale(gx, xs[a + 1], plan.inv_cnt, relu)
            gs[a], gns[a] = g, gn
            if a > 0 or need_x0:
                dxj = ops.contract_dx(h, w3q, plan.dst, gn, precision)
                # scatter dx_j onto sources + (sum gn[dst]) @ B3^T + g @ root^T, one kernel
                gx, _ = ops.segment_matvec(dxj, plan.src_perm, plan.src_ptr, plan.dst, gn, None,
                                           B3T, g, rootsT[c], None, False)
            else:
                gx = None
        # node-level parameter gradients, batched over applications
        G = torch.stack(gs)  # [A, N, W]
        X = torch.stack(xs[:A])
        droot, dbias = [None, None], [None, None]
        for c in (0, 1):
            idx = [a for a in range(A) if schedule[a] == c]
            if idx:
                droot[c], dbias[c] = ops.xty(X[idx].reshape(-1, W), G[idx].reshape(-1, W), True)
            else:
                droot[c] = torch.zeros_like(roots[c])
                dbias[c] = torch.zeros(W, dtype=torch.float32, device=gout.device)
        dw3 = db3 = dw2 = db2 = dw1 = db1 = dea = None
        if need_edge:
            # db3[i*W+o] = sum_e x_j[e,i] gn[dst e,o] = sum_n (mean-scaled agg)[n,i] g[n,o]
            db3 = ops.xty(torch.stack(list(aggs)).reshape(-1, W), G.reshape(-1, W))[0].reshape(-1)
            dw3, dw2, db2, dw1, db1, dea = _edge_grads(
                X, torch.stack(gns), plan, ea, h1, h, w1, w2, w3q, need_ea, precision)
        return (gx if need_x0 else None, dea, dw1, db1, dw2, db2, dw3, db3, droot[0], dbias[0],
                droot[1] if ctx.has2 else None, dbias[1] if ctx.has2 else None,
                None, None, None, None)


# ------------------------------------------------------------------------------------------------
# Node-first factorisation (default in bf16 mode): the same trilinear form, with W3
# contracted against the NODE features first (csrc/gkn_node_gemm.cu).  Every application is then
#   U = x W3r (dense GEMM, [N, w] x [w, w*K])  ->  msg = grouped GEMM(h, U[src])  ->  node update
# and its backward is  dU = grouped GEMM(h^T, gm)  ->  dx = dU W3^T,  dW3 += x^T dU  (dense GEMMs);
# dh_e is one grouped GEMM over all applications.  About w/(1 + w N/E) times fewer FLOPs than the
# edge-first contraction; all kernels stream h once and are HBM-bound.
# ------------------------------------------------------------------------------------------------
FACTORIZATION = os.environ.get("GKN_FACTORIZATION", "node")
# node-indexed temporaries (U, dU: w*K bf16 values per node) are produced for node ranges of at
# most this many bytes
_NODE_CHUNK_BYTES = int(os.environ.get("GKN_NODE_CHUNK_MB", "4096")) << 20
# ... and, when U is recomputed for the dh pass (A * w*K values per node), of this many
_NODE_DH_CHUNK_BYTES = int(os.environ.get("GKN_NODE_DH_CHUNK_MB", "8192")) << 20
# rows per chunk of the edge-MLP backward (bounds its dz1 temporary: 512 MB at K = 1024).  Measured
# at the bench shape (1.03 M padded edges): 32768 rows 34.06, 65536 33.33, 131072 33.03, 262144
# 32.75 ms per step - a chunk is far larger than the L2 either way, fewer and larger GEMMs win
_MLP_BWD_CHUNK_ROWS = int(os.environ.get("GKN_MLP_BWD_CHUNK_ROWS", "262144"))
# MN-major operand variants of the dU / dh kernels: h, gm and U are read in place (no transposed
# copy of h, no [k,o]-ordered recomputation of U).  GKN_NODE_MN=0 selects the K-major kernels.
NODE_MN = os.environ.get("GKN_NODE_MN", "1") == "1"
# the dh kernel applies the ReLU backward of layers.2 in its epilogue (gkn_node_dh_mn_relu_bf16)
# from a one-bit-per-element copy of the activation pattern (gkn_relu_bits_bf16): 2 GB read once
# instead of a 6 GB elementwise pass per step; GKN_DH_FUSED_MASK=0 keeps the separate mask pass
DH_FUSED_MASK = os.environ.get("GKN_DH_FUSED_MASK", "1") == "1"
_MLP_SIDE_GEMM = os.environ.get("GKN_MLP_SIDE_GEMM", "1") == "1"
# U of every application ([A, N, w*K] bf16) is kept for the dh pass when it fits this budget
# (6.3 GB for 8 BBA frames); otherwise it is recomputed node range by node range in backward
_SAVE_U_BYTES = int(os.environ.get("GKN_NODE_SAVE_U_MB", "24576")) << 20
# How U = x W3r reaches the grouped message kernel (csrc/gkn_node_fused.cu), GKN_NODE_FUSED.
# Measured at the bench shape (16 BBA frames, one B200):
#   "0" (default) library GEMM (torch.mm, 0.185 ms) + gkn_node_msg_bf16 (0.51 ms): 36.4 ms/step;
#   "gemm"  U written by the producer half of the fused kernel run on its own - a hand-written
#           tcgen05 GEMM, 0.205 ms isolated - then gkn_node_msg_bf16: 37.3 ms/step;
#   "1"     one launch per application produces AND consumes U, which crosses the L2 only
#           (-17 % HBM traffic by ncu), but its in-order item pipelines leave it latency-bound:
#           0.75 ms per application in the step, 37.7 ms/step.
# The fastest form is the default; the other two are the same function (tests compare all three).
_NODE_FUSED_MODE = os.environ.get("GKN_NODE_FUSED", "0")
NODE_FUSED = _NODE_FUSED_MODE == "1"
NODE_U_GEMM = _NODE_FUSED_MODE == "gemm"
# node blocks of the U ring when U is not kept for the backward (forward-only calls)
_FUSED_RING = int(os.environ.get("GKN_NODE_FUSED_RING", "4"))
_FUSED_SLACK = float(os.environ.get("GKN_NODE_FUSED_SLACK", "1.5"))
# Fused backward (csrc/gkn_node_fused_bwd.cu, default; GKN_NODE_FUSED_BWD=0 selects the three-kernel
# form): dU is produced into a ring of node blocks and consumed from the L2 by the weight-gradient
# / dx products of the same launch - two decoupled item streams per CTA, the weight gradient of
# layers.4 accumulating in TMEM, no library GEMM, 41 % less HBM traffic per application (ncu).
# 0.77 ms per application in the training step against 0.94 ms for gkn_node_du_mn_bf16 + two
# library GEMMs (bench shape, one B200): 35.8 vs 37.1 ms per step.
NODE_FUSED_BWD = os.environ.get("GKN_NODE_FUSED_BWD", "1") == "1"
_FUSED_BWD_RING = int(os.environ.get("GKN_NODE_FUSED_BWD_RING", "3"))
_FUSED_BWD_MIN_DEGREE = int(os.environ.get("GKN_NODE_FUSED_BWD_MIN_DEGREE", "64"))


# Training at w = 32 on graphs of low degree stays on the edge-first kernels: there the per-edge
# contraction is only 2*K*32*32 FLOP (1.4 ns per edge on the tensor cores) while the node-first
# path moves U = N*32*K bf16 per application through HBM - 0.3 + 20/degree ns per edge.  Measured
# on config 1 (28 residues x batch 128, 14 edges per node): 6.2 ms per step edge-first, 9.4 ms
# node-first; the node-first kernels win from about 20 edges per node (model) upwards.
_NODE_MIN_DEGREE_W32 = float(os.environ.get("GKN_NODE_MIN_DEGREE_W32", "24"))


def node_path_supported(W: int, K: int, E: int, precision: str, needs_grad: bool = True,
                        num_nodes: int | None = None) -> bool:
    """w in {32, 64, 128} (the K-major backward kernels of GKN_NODE_MN=0 serve w = 64 only);
    ``num_nodes`` lets the w = 32 training case choose by degree (see above)."""
    if precision != "bf16" or K % 256 != 0 or E <= 0:
        return False
    if W == 64:
        return True
    if W not in (32, 128) or (needs_grad and not NODE_MN):
        return False
    if W == 32 and needs_grad and num_nodes and E < _NODE_MIN_DEGREE_W32 * num_nodes:
        return False
    return True


def _node_chunks(N: int, per_node_bytes: int, budget: int | None = None):
    nc = max(1, min(N, (_NODE_CHUNK_BYTES if budget is None else budget) // per_node_bytes))
    return [(n0, min(N, n0 + nc)) for n0 in range(0, N, nc)]


class _GKNStackNode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, ea, w1, b1, w2, b2, w3, b3, root1, bias1, root2, bias2, plan, schedule,
                relu, precision):
        W = x0.size(1)
        det = lambda t: None if t is None else t.detach()  # noqa: E731
        x0, ea, w1, b1, w2, b2, w3, b3 = map(det, (x0, ea, w1, b1, w2, b2, w3, b3))
        lay = ops.node_layout(plan)
        N, K = plan.num_nodes, w3.size(1)
        bf = torch.bfloat16
        # edge attributes and the edge MLP in the padded source-sorted layout (pad rows: zeros)
        ea_s = torch.zeros((lay.e_pad, ea.size(1)), dtype=torch.float32, device=ea.device)
        ea_s.index_copy_(0, lay.qpos_long, ea.contiguous())
        h1, hs = _edge_hidden(ea_s, w1, b1.contiguous(), w2, b2.contiguous(), "bf16")
        w3q = w3.to(bf).contiguous()           # [w*w, K] = [i, o, k]
        w3r = w3q.view(W, W * K)               # [i, (o, k)]
        B3c = b3.view(W, W).contiguous()
        roots = (det(root1), det(root2))
        biases = (det(bias1), det(bias2))
        x = x0.contiguous()
        xs, aggs = [x], []
        want_agg = any(ctx.needs_input_grad)
        msg = torch.empty((lay.e_pad, W), dtype=torch.float32, device=x.device)
        chunks = _node_chunks(N, W * K * 2)
        A = len(schedule)
        u_all = None
        if NODE_MN and want_agg and A * N * W * K * 2 <= _SAVE_U_BYTES:
            u_all = torch.empty((A, N, W * K), dtype=bf, device=x.device)
        fused = None
        if NODE_U_GEMM:
            w3g = w3r if W >= 64 else torch.cat([w3r, torch.zeros_like(w3r)], 0)
        if NODE_FUSED and (u_all is not None or not want_agg):
            # one launch per application: U tiles are produced and consumed inside it.  With a
            # backward to come U is written to the kept buffer; forward-only calls use a small ring
            ring = 0 if u_all is not None else _FUSED_RING
            fused = ops.fused_fwd_schedule(lay, W, K, ring=ring, slack=_FUSED_SLACK)
            flags = torch.zeros((A, fused.num_flags), dtype=torch.int32, device=x.device)
            u_ring = None if u_all is not None else torch.empty((fused.u_rows, W * K), dtype=bf,
                                                                device=x.device)
            w3p = w3r if W >= 64 else torch.cat([w3r, torch.zeros_like(w3r)], 0)
        for a, c in enumerate(schedule):
            xb = x.to(bf)
            if fused is not None:
                xp = xb if W >= 64 else torch.cat([xb, torch.zeros_like(xb)], 1)
                ops.node_fused_fwd(xp, w3p, hs, u_all[a] if u_all is not None else u_ring, fused,
                                   flags[a], msg, N, lay.e_pad, K, W)
            elif u_all is not None:
                if NODE_U_GEMM:
                    xp = xb if W >= 64 else torch.cat([xb, torch.zeros_like(xb)], 1)
                    ops.node_u_gemm(xp, w3g, u_all[a], N, K, W)   # [N, (o, k)], kept for the dh pass
                else:
                    torch.mm(xb, w3r, out=u_all[a])
                ops.node_msg(hs, u_all[a].view(N, W, K), lay, 0, lay.num_units, 0, msg)
            else:
                for n0, n1 in chunks:
                    if NODE_U_GEMM:
                        xp = xb[n0:n1] if W >= 64 else torch.cat([xb[n0:n1], torch.zeros_like(xb[n0:n1])], 1)
                        U = torch.empty((n1 - n0, W * K), dtype=bf, device=x.device)
                        ops.node_u_gemm(xp.contiguous(), w3g, U, n1 - n0, K, W)
                    else:
                        U = torch.mm(xb[n0:n1], w3r)   # [nc, (o, k)]
                    ops.node_msg(hs, U.view(n1 - n0, W, K), lay, lay.unit_first[n0],
                                 lay.unit_first[n1], n0, msg)
                    del U
            x_new, agg = ops.segment_matvec(msg, lay.qpos, plan.dst_ptr, lay.src_pad, x,
                                            plan.inv_cnt, B3c, x, roots[c].contiguous(),
                                            biases[c].contiguous(), relu, want_agg=want_agg)
            if want_agg:
                aggs.append(agg)
            x = x_new
            xs.append(x)
        ctx.plan, ctx.schedule, ctx.relu = plan, schedule, relu
        ctx.has2 = root2 is not None
        ctx.n_apps = len(schedule)
        ctx.has_u = u_all is not None
        ctx.save_for_backward(ea_s, h1, hs, w1, w2, w3q, B3c, roots[0],
                              roots[1] if ctx.has2 else roots[0], *xs, *aggs,
                              *([u_all] if u_all is not None else []))
        return x

    @staticmethod
    def backward(ctx, gout):
        ea_s, h1, hs, w1, w2, w3q, B3, r1, r2, *rest = ctx.saved_tensors
        plan, schedule, relu = ctx.plan, ctx.schedule, ctx.relu
        lay = ops.node_layout(plan)
        A = ctx.n_apps
        u_all = rest.pop() if ctx.has_u else None
        xs, aggs = rest[:A + 1], rest[A + 1:]
        roots = (r1, r2)
        rootsT = (r1.t().contiguous(), r2.t().contiguous())
        B3T = B3.t().contiguous()
        N, W = xs[0].shape
        K = hs.size(1)
        bf = torch.bfloat16
        dev = hs.device
        need = ctx.needs_input_grad
        need_x0, need_ea, need_edge = need[0], need[1], any(need[1:8])
        # W3 as [i, (k, o)]: the operand of dx = dU W3^T, dW3 = x^T dU and of U^T for dh
        w3kq = w3q.view(W, W, K).permute(0, 2, 1).contiguous().view(W, K * W)
        w3r = w3q.view(W, W * K)
        hsT = gsT = gs_one = None
        if not NODE_MN:
            hsT = hs.t().contiguous()  # [K, E_pad]: edges become the reduction dimension of dU
            gsT = torch.empty((W, lay.e_pad), dtype=bf, device=dev)
        elif not need_edge:
            gs_one = torch.empty((lay.e_pad, ops.gs_width(W)), dtype=bf, device=dev)
        gs_all = (torch.empty((A, lay.e_pad, ops.gs_width(W)), dtype=bf, device=dev)
                  if need_edge else None)
        # the fused kernel pays 148 dx partials (38 KB) per node and one work item per (node, 128
        # k-rows): it wins on graphs of BBA-like degree (119 edges per node: 0.80 vs 0.94 ms) and
        # loses a few per cent at ~30 edges per node (10k-atom frames: 125 vs 120 ms per step)
        fused_bwd = (NODE_FUSED_BWD and NODE_MN and W == 64
                     and lay.e_pad >= _FUSED_BWD_MIN_DEGREE * N)
        if fused_bwd:
            try:
                bsched = ops.fused_bwd_schedule(N, K, W, dev, ring=_FUSED_BWD_RING)
            except RuntimeError:  # fewer SMs than the column slicing of this kernel width needs
                fused_bwd = False
        dw3k = dwT = None
        if need_edge and not fused_bwd:
            dw3k = torch.zeros((W, K * W), dtype=torch.float32, device=dev)
        if fused_bwd:
            _, parts = ops.fused_bwd_grid(K, W)
            du_ring = torch.zeros((bsched.u_rows, K * W), dtype=bf, device=dev)
            bflags = torch.zeros((A, bsched.num_flags), dtype=torch.int32, device=dev)
            dx_part = torch.empty((parts, N, W), dtype=torch.float32, device=dev)
            dwT = torch.empty((K * W, W), dtype=torch.float32, device=dev) if need_edge else None
            first_dw = True
        chunks = _node_chunks(N, W * K * 2)
        gs = [None] * A
        xbs = [None] * A
        gx = gout.contiguous().float()
        for a in range(A - 1, -1, -1):
            c = schedule[a]
            g, gn = ops.relu_grad_scale(gx, xs[a + 1], plan.inv_cnt, relu)
            gs[a] = g
            need_dx = a > 0 or need_x0
            if not (need_dx or need_edge):
                gx = None
                continue
            gs_a = gs_all[a] if need_edge else gs_one
            ops.node_gather(gn, lay, gs_a, gsT)
            xbs[a] = xs[a].to(bf)
            if fused_bwd:
                ops.node_fused_bwd(hs, gs_a, lay, xbs[a], w3kq, du_ring, bsched, bflags[a],
                                   dx_part if need_dx else None, dwT, not first_dw)
                first_dw = False
                if need_dx:
                    rest_g, _ = ops.segment_matvec(None, plan.src_perm, plan.src_ptr, plan.dst, gn,
                                                   None, B3T, g, rootsT[c], None, False)
                    gx = ops.sum_partials(dx_part, rest_g, rest_g)
                else:
                    gx = None
                continue
            dxc = torch.empty((N, W), dtype=torch.float32, device=dev) if need_dx else None
            for n0, n1 in chunks:
                dut = torch.empty((n1 - n0, K * W), dtype=bf, device=dev)
                if NODE_MN:
                    ops.node_du_mn(hs, gs_a, lay, n0, n1, dut)
                else:
                    ops.node_du(hsT, gsT, lay, n0, n1, dut)
                if need_dx:
                    dxc[n0:n1] = _mm32(dut, w3kq.t())
                if need_edge:
                    dw3k = _addmm32(dw3k, xbs[a][n0:n1].t(), dut)
                del dut
            if need_dx:
                # (sum over outgoing edges of gn[dst]) @ B3^T + g @ root^T, then + dU W3^T
                rest_g, _ = ops.segment_matvec(None, plan.src_perm, plan.src_ptr, plan.dst, gn, None,
                                               B3T, g, rootsT[c], None, False)
                gx = rest_g.add_(dxc)
            else:
                gx = None
        if fused_bwd:
            del du_ring, dx_part
            if need_edge:
                dw3k = dwT.t()  # [i, (k, o)]
        del hsT, gsT, gs_one
        G = torch.stack(gs)
        X = torch.stack(xs[:A])
        droot, dbias = [None, None], [None, None]
        for c in (0, 1):
            idx = [a for a in range(A) if schedule[a] == c]
            if idx:
                droot[c], dbias[c] = ops.xty(X[idx].reshape(-1, W), G[idx].reshape(-1, W), True)
            else:
                droot[c] = torch.zeros_like(roots[c])
                dbias[c] = torch.zeros(W, dtype=torch.float32, device=dev)
        dw3 = db3 = dw2 = db2 = dw1 = db1 = dea = None
        if need_edge:
            db3 = ops.xty(torch.stack(list(aggs)).reshape(-1, W), G.reshape(-1, W))[0].reshape(-1)
            dw3 = dw3k.reshape(W, K, W).permute(0, 2, 1).reshape(W * W, K)
            # dh_e for all applications at once, node range by node range (U^T recomputed)
            dh = torch.empty((lay.e_pad, K), dtype=bf, device=dev)
            # the ReLU backward of layers.2 and its bias gradient ride in the dh epilogue
            mask_kw, db2_fused = {}, None
            if NODE_MN and DH_FUSED_MASK and K <= 1024:
                db2_fused = torch.zeros(K, dtype=torch.float32, device=dev)
                mask_kw = dict(act_bits=ops.relu_bits(hs), colsum=db2_fused)
            if u_all is not None:
                ops.node_dh_mn(gs_all, u_all.view(A, N, W, K), lay, 0, lay.num_units, 0, dh,
                               **mask_kw)
            else:
                for n0, n1 in _node_chunks(N, A * W * K * 2, _NODE_DH_CHUNK_BYTES):
                    ut = torch.empty((A, n1 - n0, K * W), dtype=bf, device=dev)
                    for a in range(A):
                        torch.mm(xbs[a][n0:n1], w3r if NODE_MN else w3kq, out=ut[a])
                    if NODE_MN:
                        ops.node_dh_mn(gs_all, ut.view(A, n1 - n0, W, K), lay, lay.unit_first[n0],
                                       lay.unit_first[n1], n0, dh, accumulate=True, **mask_kw)
                    else:
                        ops.node_dh(gs_all, ut.view(A, n1 - n0, K, W), lay, lay.unit_first[n0],
                                    lay.unit_first[n1], n0, dh)
                    del ut
            del gs_all, u_all
            dw2, db2, dw1, db1, dea_s = _mlp_grads(dh, ea_s, h1, hs, w1, w2, need_ea,
                                                   db2_masked=db2_fused)
            if need_ea:
                dea = dea_s.index_select(0, lay.qpos_long)
        return (gx if need_x0 else None, dea, dw1, db1, dw2, db2, dw3, db3, droot[0], dbias[0],
                droot[1] if ctx.has2 else None, dbias[1] if ctx.has2 else None,
                None, None, None, None)


def _mlp_grads(dh, ea, h1, h, w1, w2, need_ea: bool, db2_masked=None):
    """Backward of edge-MLP layers .0-.3 from dh_e (bf16 [rows, K]); rows in bounded chunks.
    ``db2_masked``: dh already is dz2 (masked by layers.2's activation, ops.node_dh_mn with
    ``act``) and this is its column sum."""
    E, K = h.shape
    kin = ea.size(1)
    f32 = dict(dtype=torch.float32, device=h.device)
    dw2 = torch.zeros((K, K), **f32)
    db2 = torch.zeros(K, **f32) if db2_masked is None else db2_masked
    cs1 = torch.zeros((1 + kin, K), **f32)
    dea = torch.empty_like(ea) if need_ea else None
    w2q = w2.to(torch.bfloat16)
    side_gemm = _MLP_SIDE_GEMM and not need_ea
    if side_gemm:
        # dW1 = dz1^T edge_attr as a tensor-core GEMM: edge_attr split into bf16 hi + lo parts
        # (16 columns: hi, lo, zero pad) keeps it exact to 2^-17
        hi = ea.to(torch.bfloat16)
        lo = (ea - hi.float()).to(torch.bfloat16)
        ea_ext = torch.cat([hi, lo, torch.zeros((E, 16 - 2 * kin), dtype=torch.bfloat16,
                                                device=h.device)], 1)
        side_acc = torch.zeros((K, 16), **f32)
    for e0 in range(0, E, _MLP_BWD_CHUNK_ROWS):
        e1 = min(E, e0 + _MLP_BWD_CHUNK_ROWS)
        if db2_masked is None:
            dz2q = ops.mask_cast_colsum(dh[e0:e1], h[e0:e1], db2, torch.bfloat16)
        else:
            dz2q = dh[e0:e1]
        dw2 = _addmm32(dw2, dz2q.t(), h1[e0:e1])
        if need_ea:
            dz1 = _mm32(dz2q, w2q)
            dz1 = ops.mask_cast_colsum(dz1, h1[e0:e1], cs1, torch.float32, side=ea[e0:e1])
            torch.mm(dz1, w1, out=dea[e0:e1])
        elif side_gemm:
            dz1 = torch.mm(dz2q, w2q)  # bf16: the GEMM rounds its fp32 accumulators itself
            dz1 = ops.mask_cast_colsum(dz1, h1[e0:e1], cs1[0], torch.bfloat16)  # in place, + db1
            side_acc = _addmm32(side_acc, dz1.t(), ea_ext[e0:e1])
        else:
            # only the column sums (db1, dW1) are needed
            dz1 = torch.mm(dz2q, w2q)
            ops.mask_cast_colsum(dz1, h1[e0:e1], cs1, torch.bfloat16, side=ea[e0:e1],
                                 sums_only=True)
    if side_gemm:
        cs1[1:] = (side_acc[:, :kin] + side_acc[:, kin:2 * kin]).t()
    return dw2, db2, cs1[1:].t().contiguous(), cs1[0].contiguous(), dea


def gkn_stack(x0, edge_attr_sorted, net, root1, bias1, root2, bias2, plan: GraphPlan, schedule,
              relu: bool, precision: str = "bf16") -> torch.Tensor:
    """Edge MLP + ``len(schedule)`` NNConv applications sharing it.

    ``edge_attr_sorted`` is edge_attr in the plan's destination-sorted order; ``net`` is the
    DenseNet([ker_in, K, K, w*w], ReLU); ``schedule[a]`` picks the (root, bias) pair (0 or 1) of
    application a; ``relu`` applies F.relu after every application (graph_kernel.py:299-302)."""
    if x0.dtype != torch.float32 or edge_attr_sorted.dtype != torch.float32:
        raise TypeError("node features and edge attributes must be float32")
    l0, l2, l4 = net.layers[0], net.layers[2], net.layers[4]
    fn = _GKNStack
    tensors = (x0, edge_attr_sorted, l0.weight, l0.bias, l2.weight, l2.bias, l4.weight, l4.bias,
               root1, bias1, root2, bias2)
    needs_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)
    if FACTORIZATION == "node" and node_path_supported(x0.size(1), l4.weight.size(1),
                                                       plan.num_edges, precision, needs_grad,
                                                       plan.num_nodes):
        fn = _GKNStackNode
    return fn.apply(x0, edge_attr_sorted, l0.weight, l0.bias, l2.weight, l2.bias, l4.weight,
                           l4.bias, root1, bias1, root2, bias2, plan, tuple(schedule), bool(relu),
                           precision)


# ------------------------------------------------------------------------------------------------
# Node encoder's LSTM (reference graph_kernel.py:278-285) - csrc/gkn_lstm.cu
# ------------------------------------------------------------------------------------------------
# GKN_LSTM=0 leaves nn.LSTM to cuDNN / the native torch kernels
OWN_LSTM = os.environ.get("GKN_LSTM", "1") == "1"


class _LSTMSeq(torch.autograd.Function):
    """``nn.LSTM(3, 3)`` over one [T, R, 3] sequence with zero initial state: all h_t."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh):
        x, w_ih, w_hh, b_ih, b_hh = (t.detach().contiguous() for t in (x, w_ih, w_hh, b_ih, b_hh))
        save = any(ctx.needs_input_grad)
        out, gates, cells = ops.lstm_fwd(x, w_ih, w_hh, b_ih, b_hh, save)
        if save:
            ctx.save_for_backward(x, w_ih, w_hh, out, gates, cells)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w_ih, w_hh, out, gates, cells = ctx.saved_tensors
        dx, dw_ih, dw_hh, db = ops.lstm_bwd(x, w_ih, w_hh, out, gates, cells,
                                            dout.contiguous().float(), ctx.needs_input_grad[0])
        return dx, dw_ih, dw_hh, db, db.clone()


def lstm_supported(lstm: torch.nn.LSTM, x: torch.Tensor) -> bool:
    return (OWN_LSTM and x.is_cuda and x.dtype == torch.float32 and lstm.input_size == 3
            and lstm.hidden_size == 3 and lstm.num_layers == 1 and not lstm.bidirectional
            and lstm.bias and not lstm.batch_first and lstm.proj_size == 0)


def lstm_sequence(lstm: torch.nn.LSTM, x: torch.Tensor) -> torch.Tensor:
    """All hidden states of ``lstm`` over x [T, R, 3] from a zero initial state."""
    return _LSTMSeq.apply(x, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0,
                          lstm.bias_hh_l0)

