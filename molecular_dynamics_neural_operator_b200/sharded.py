"""Intra-frame sharding of the conv stack for frames too large for one GPU to turn around quickly
(BASELINE config 5: 100k atoms, 3 M edges, 16 applications per rollout step).

The path has exactly one natural exchange: each rank owns a contiguous range of
DESTINATION nodes together with their incoming edges — contiguous in the destination-sorted edge
layout — so the per-edge work (edge MLP, fused contraction) and the segmented mean are purely
local, and after every application the ranks all-gather the new node features ``x`` (N x w fp32:
51 MB at w = 128, about 70 us on NVLink 5 against ~9 ms of contraction per rank).  Inference only
(forward); training shards over frames instead (train.py).

``emulate_world`` runs all ranks of the exchange in one process on one GPU (the all-gather becomes
a concatenation): that is how the sharding arithmetic is tested without a multi-GPU box.
"""
from __future__ import annotations

from dataclasses import dataclass
import torch
import torch.distributed as dist

from . import functional as GF
from . import ops
from .ops import GraphPlan


@dataclass
class FrameShard:
    """Rank-local slice of a destination-sorted graph: nodes [n0, n1), their incoming edges."""

    n0: int
    n1: int
    rows: int  # padded rows every rank contributes to the all-gather (ceil(N / world))
    src: torch.Tensor  # int32 [E_loc]  GLOBAL source ids of the local edges
    dst_ptr: torch.Tensor  # int32 [n1-n0+1]  local CSR pointers (offset removed)
    inv_cnt: torch.Tensor  # float32 [n1-n0]
    edge_attr: torch.Tensor  # float32 [E_loc, 6]  in destination-sorted order


def shard_frame(plan: GraphPlan, edge_attr_sorted: torch.Tensor, rank: int, world: int) -> FrameShard:
    """Equal node ranges (ceil(N / world) rows per rank, the last one possibly shorter)."""
    N = plan.num_nodes
    rows = (N + world - 1) // world
    n0, n1 = min(N, rank * rows), min(N, (rank + 1) * rows)
    bounds = plan.dst_ptr[[n0, n1]].tolist()  # two ints: one small D2H read per graph
    e0, e1 = int(bounds[0]), int(bounds[1])
    return FrameShard(
        n0=n0, n1=n1, rows=rows,
        src=plan.src[e0:e1].contiguous(),
        dst_ptr=(plan.dst_ptr[n0:n1 + 1] - e0).contiguous(),
        inv_cnt=plan.inv_cnt[n0:n1].contiguous(),
        edge_attr=edge_attr_sorted[e0:e1].contiguous(),
    )


def _gather_rows(x_loc: torch.Tensor, shard: FrameShard, num_nodes: int, group) -> torch.Tensor:
    """All-gather the per-rank node rows into the full [N, w] tensor (ranks padded to equal rows)."""
    world = dist.get_world_size(group)
    W = x_loc.size(1)
    buf = torch.empty((world * shard.rows, W), dtype=x_loc.dtype, device=x_loc.device)
    if x_loc.size(0) != shard.rows:
        pad = torch.zeros((shard.rows, W), dtype=x_loc.dtype, device=x_loc.device)
        pad[: x_loc.size(0)] = x_loc
        x_loc = pad
    dist.all_gather_into_tensor(buf, x_loc.contiguous(), group=group)
    return buf[:num_nodes]


@torch.no_grad()
def sharded_stack_forward(x0: torch.Tensor, shard: FrameShard, net, root1, bias1, root2, bias2,
                          schedule, relu: bool, precision: str = "bf16", group=None) -> torch.Tensor:
    """Forward of the conv stack with destination-range sharding.  Every rank passes the same full
    ``x0`` [N, w] and gets the full result back."""
    N, W = x0.shape
    l0, l2, l4 = net.layers[0], net.layers[2], net.layers[4]
    _, h = GF._edge_hidden(shard.edge_attr, l0.weight, l0.bias.contiguous(), l2.weight,
                           l2.bias.contiguous(), precision)
    w3q = (l4.weight.to(torch.bfloat16) if precision == "bf16" else l4.weight).contiguous()
    B3 = l4.bias.view(W, W).contiguous()
    roots, biases = (root1, root2), (bias1, bias2)
    x = x0.contiguous()
    for c in schedule:
        msg = ops.contract_fwd(h, w3q, shard.src, x, precision)
        x_loc, _ = ops.segment_matvec(msg, None, shard.dst_ptr, shard.src, x, shard.inv_cnt, B3,
                                      x[shard.n0:shard.n1], roots[c].contiguous(),
                                      biases[c].contiguous(), relu)
        x = _gather_rows(x_loc, shard, N, group)
    return x


@torch.no_grad()
def emulate_world(x0, plan: GraphPlan, edge_attr_sorted, net, root1, bias1, root2, bias2, schedule,
                  relu: bool, world: int, precision: str = "bf16") -> torch.Tensor:
    """All `world` ranks of sharded_stack_forward executed in turn on one GPU; the all-gather is a
    concatenation.  Same kernels, same per-rank arguments as the distributed run."""
    N, W = x0.shape
    l0, l2, l4 = net.layers[0], net.layers[2], net.layers[4]
    shards = [shard_frame(plan, edge_attr_sorted, r, world) for r in range(world)]
    hs = [GF._edge_hidden(s.edge_attr, l0.weight, l0.bias.contiguous(), l2.weight,
                          l2.bias.contiguous(), precision)[1] for s in shards]
    w3q = (l4.weight.to(torch.bfloat16) if precision == "bf16" else l4.weight).contiguous()
    B3 = l4.bias.view(W, W).contiguous()
    roots, biases = (root1, root2), (bias1, bias2)
    x = x0.contiguous()
    for c in schedule:
        parts = []
        for s, h in zip(shards, hs):
            if s.n1 == s.n0:
                continue
            msg = ops.contract_fwd(h, w3q, s.src, x, precision)
            x_loc, _ = ops.segment_matvec(msg, None, s.dst_ptr, s.src, x, s.inv_cnt, B3,
                                          x[s.n0:s.n1], roots[c].contiguous(),
                                          biases[c].contiguous(), relu)
            parts.append(x_loc)
        x = torch.cat(parts, dim=0)
    return x


@torch.no_grad()
def kernelnn_forward_sharded(model, data, group=None) -> torch.Tensor:
    """``KernelNN.forward`` for one large frame with the conv stack sharded over the ranks of
    ``group`` (node encoder and fc2 are tiny and replicated).  Every rank returns the full output."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    x = model.node_encoder(data)
    plan = getattr(data, "_gkn_plan", None)
    if plan is None:
        plan = ops.prepare_graph(data.edge_index, x.size(0))
    ea = data.edge_attr.index_select(0, plan.perm_long)
    shard = shard_frame(plan, ea, rank, world)
    schedule = (0,) * model.depth + (1,) * model.depth
    x = sharded_stack_forward(x.contiguous(), shard, model.conv1.net, model.conv1.root,
                              model.conv1.bias, model.conv2.root, model.conv2.bias, schedule,
                              relu=True, precision=model.precision, group=group)
    return model.fc2(x)
