"""Autograd glue for the graph-kernel layer: one ``torch.autograd.Function`` for the whole stack of
NNConv applications that share an edge MLP (reference graph_kernel.py:271-273, :299-302).

Why one Function for the stack and not one per layer: the edge hidden state ``h_e`` (output of
net.layers.0-.3), the last-layer weight W3/b3 and the graph are identical for all 2*depth
applications, so

* forward computes ``h_e`` once and each application is two kernels: the fused tcgen05 contraction
  (``msg``) and the segmented mean (+ root, bias, ReLU);
* backward recomputes the per-edge kernel flash-style for ``dx`` (one contraction per application)
  and DEFERS the gradients of the shared operands: with
      S_e[i,o] = sum_a x^(a-1)[src e, i] * gn^(a)[dst e, o]
  (the loss gradient w.r.t. the per-edge kernel summed over applications) a single pass gives
      dW3 = S^T h_e,   db3 = sum_e S_e,   dh_e = S W3
  i.e. 2 GEMM units instead of 2 per application.

Everything E-sized that is saved for backward is ``h_e`` itself; per application only the node
tensors ``x^(a)`` ([N, w]) are kept.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from .ops import GraphPlan

# rows of S materialised at a time by the staged weight-gradient path (see _deferred_grads)
_S_CHUNK_EDGES = 16384


def edge_hidden(net, edge_attr_sorted: torch.Tensor, precision: str) -> torch.Tensor:
    """h_e = relu(layers.2(relu(layers.0(edge_attr))))  (reference graph_kernel.py:239-242).

    These two layers are ordinary dense GEMMs executed once per forward (cuBLAS through torch).
    bf16 mode: layer 0 in fp32 (fan-in 6), layer 2 with bf16 operands / fp32 accumulation, result
    kept in bf16 — exactly the A operand the tensor-core contraction consumes.
    """
    l0, l2 = net.layers[0], net.layers[2]
    h1 = F.relu(F.linear(edge_attr_sorted, l0.weight, l0.bias))
    if precision == "bf16":
        h = F.relu(F.linear(h1.to(torch.bfloat16), l2.weight.to(torch.bfloat16),
                            l2.bias.to(torch.bfloat16)))
    else:
        h = F.relu(F.linear(h1, l2.weight, l2.bias))
    return h.contiguous()


def _deferred_grads(xs, gns, plan: GraphPlan, h, w3q, need_h: bool, need_w: bool, precision: str):
    """dh_e, dW3, db3 from the per-application node tensors (see module docstring).

    Staged implementation: S is produced for a bounded chunk of edges at a time by
    gkn_edge_outer_sum and multiplied with library GEMMs.  (The fused tcgen05 kernels that never
    store S replace this path; it is kept as the fp32 route and cross-check.)
    """
    E, K = h.shape
    WW = w3q.shape[0]
    dev = h.device
    dh = torch.empty_like(h) if need_h else None
    dw3 = torch.zeros((WW, K), dtype=torch.float32, device=dev) if need_w else None
    db3 = torch.zeros(WW, dtype=torch.float32, device=dev) if need_w else None
    s_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
    for e0 in range(0, E, _S_CHUNK_EDGES):
        e1 = min(E, e0 + _S_CHUNK_EDGES)
        S = ops.edge_outer_sum(xs, gns, plan, e0, e1, s_dtype)
        if need_h:
            torch.mm(S, w3q, out=dh[e0:e1])
        if need_w:
            dw3 += torch.mm(S.t(), h[e0:e1]).float()
            db3 += S.sum(0, dtype=torch.float32)
    return dh, dw3, db3


class _ConvStack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, h, w3, b3, root1, bias1, root2, bias2, plan, schedule, relu, precision):
        W = x0.size(1)
        if precision == "bf16":
            w3q = w3.detach().to(torch.bfloat16).contiguous()
        else:
            w3q = w3.detach().contiguous()
        B3 = b3.detach().view(W, W)
        roots = (root1.detach(), root2.detach() if root2 is not None else None)
        biases = (bias1.detach(), bias2.detach() if bias2 is not None else None)
        x = x0.detach().contiguous()
        hq = h.detach()
        xs = [x]
        for c in schedule:
            xroot = torch.mm(x, roots[c])
            xb3 = torch.mm(x, B3)
            msg = ops.contract_fwd(hq, w3q, plan.src, x, precision)
            x = ops.scatter_mean(msg, plan, xb3=xb3, xroot=xroot, bias=biases[c].contiguous(),
                                 relu=relu)
            xs.append(x)
        ctx.plan = plan
        ctx.schedule = schedule
        ctx.relu = relu
        ctx.precision = precision
        ctx.has2 = root2 is not None
        ctx.save_for_backward(hq, w3q, B3, roots[0], roots[1] if ctx.has2 else roots[0], *xs)
        return x

    @staticmethod
    def backward(ctx, gout):
        hq, w3q, B3, r1, r2, *xs = ctx.saved_tensors
        plan, schedule, relu, precision = ctx.plan, ctx.schedule, ctx.relu, ctx.precision
        roots = (r1, r2)
        A = len(schedule)
        W = xs[0].size(1)
        need_x0, need_h, need_w3 = ctx.needs_input_grad[0], ctx.needs_input_grad[1], (
            ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        droot = [torch.zeros_like(r1), torch.zeros_like(r2)]
        dbias = [torch.zeros(W, dtype=torch.float32, device=gout.device) for _ in range(2)]
        gns = [None] * A
        gx = gout.contiguous().float()
        for a in range(A - 1, -1, -1):
            c = schedule[a]
            xin, y = xs[a], xs[a + 1]
            g, gn = ops.relu_grad_scale(gx, y, plan.inv_cnt, relu)
            droot[c].addmm_(xin.t(), g)
            dbias[c] += g.sum(0)
            gns[a] = gn
            if a > 0 or need_x0:
                dxj = ops.contract_dx(hq, w3q, plan.dst, gn, precision)
                gb3 = torch.mm(gn, B3.t())
                addend = torch.mm(g, roots[c].t())
                gx = ops.scatter_src_sum(dxj, plan, gb3=gb3, addend=addend)
            else:
                gx = None
        dh = dw3 = db3 = None
        if need_h or need_w3:
            dh, dw3, db3 = _deferred_grads(torch.stack(xs[:A]), torch.stack(gns), plan, hq, w3q,
                                           need_h, need_w3, precision)
        return (gx if need_x0 else None, dh, dw3, db3, droot[0], dbias[0],
                droot[1] if ctx.has2 else None, dbias[1] if ctx.has2 else None,
                None, None, None, None)


def conv_stack(x0, h, w3, b3, root1, bias1, root2, bias2, plan: GraphPlan, schedule, relu: bool,
               precision: str = "bf16") -> torch.Tensor:
    """Apply ``len(schedule)`` NNConv layers that share (h, w3, b3); ``schedule[a]`` picks the
    (root, bias) pair (0 or 1) of application a; ``relu`` applies F.relu after every application
    (reference graph_kernel.py:299-302)."""
    if x0.dtype != torch.float32:
        raise TypeError(f"node features must be float32, got {x0.dtype}")
    return _ConvStack.apply(x0, h, w3, b3, root1, bias1, root2, bias2, plan, tuple(schedule),
                            bool(relu), precision)
