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

Node-first factorisation (functional._GKNStackNode, the default where it applies): the per-edge
work is grouped by SOURCE node, so a rank owns a contiguous range of source nodes and their
outgoing edges.  ``U = x W3r`` is then needed only for the owned nodes (no exchange before the
message kernel), every rank produces PARTIAL sums of messages at all destinations, and the one
exchange per application is an all-reduce (sum) of those partial node sums (2 x N x w fp32) after
which every rank applies the (tiny, replicated) node update.  ``SourceShard`` /
``sharded_stack_forward_node`` / ``emulate_world_node``.
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


# ------------------------------------------------------------------------------------------------
# Source-range sharding of the node-first path
# ------------------------------------------------------------------------------------------------
@dataclass
class SourceShard:
    """Rank-local slice for the node-first kernels: source nodes [n0, n1) and their outgoing edges
    (as a plan over all N nodes that holds only those edges, with its padded layout)."""

    n0: int
    n1: int
    plan: GraphPlan
    lay: ops.NodeLayout
    edge_attr: torch.Tensor  # float32 [E_loc, 6] in the local plan's destination-sorted order
    inv_cnt: torch.Tensor  # float32 [N]  1 / max(GLOBAL in-degree, 1)


def shard_frame_by_source(edge_index: torch.Tensor, edge_attr: torch.Tensor, num_nodes: int,
                          rank: int, world: int, row_ptr: torch.Tensor | None = None,
                          inv_cnt: torch.Tensor | None = None) -> SourceShard:
    """Equal source-node ranges.  With ``row_ptr`` (edge_index sorted by source, as the radius graph
    emits it) the local edges are one contiguous slice; otherwise they are selected by a mask.
    ``inv_cnt``: 1 / max(global in-degree, 1) if the caller already has it (a full-graph plan)."""
    N = int(num_nodes)
    rows = (N + world - 1) // world
    n0, n1 = min(N, rank * rows), min(N, (rank + 1) * rows)
    if row_ptr is not None:
        e0, e1 = (int(v) for v in row_ptr[[n0, n1]].tolist())
        ei, ea = edge_index[:, e0:e1], edge_attr[e0:e1]
    else:
        sel = ((edge_index[0] >= n0) & (edge_index[0] < n1)).nonzero().squeeze(1)
        ei, ea = edge_index.index_select(1, sel), edge_attr.index_select(0, sel)
    if inv_cnt is None:
        cnt = torch.bincount(edge_index[1], minlength=N)
        inv_cnt = 1.0 / cnt.clamp(min=1).float()
    if ei.size(1) == 0:  # this rank owns no edges: it only takes part in the all-reduce
        return SourceShard(n0, n1, None, None, ea, inv_cnt)
    plan = ops.prepare_graph(ei.contiguous(), N, validate=False)
    return SourceShard(n0, n1, plan, ops.node_layout(plan), ea.index_select(0, plan.perm_long),
                       inv_cnt)


def _node_partial_sums(x, shard: SourceShard, hs, w3r, msg, out=None):
    """One application's local work: U for the owned nodes, messages of their outgoing edges, and
    the partial sums of those messages (and of x_j, for the b3 term) at every destination.
    Returns [2, N, w] fp32 (written into ``out`` when given, e.g. a symmetric-memory buffer)."""
    N, W = x.shape
    K = hs.size(1)
    lay, plan = shard.lay, shard.plan
    xb = x[shard.n0:shard.n1].to(torch.bfloat16)
    for c0, c1 in GF._node_chunks(shard.n1 - shard.n0, W * K * 2):
        U = torch.mm(xb[c0:c1], w3r)
        ops.node_msg(hs, U.view(c1 - c0, W, K), lay, lay.unit_first[shard.n0 + c0],
                     lay.unit_first[shard.n0 + c1], shard.n0 + c0, msg)
        del U
    if out is not None:
        ops.segment_matvec(msg, lay.qpos, plan.dst_ptr, lay.src_pad, x, None, None, None, None,
                           None, False, out=out[0], agg_out=out[1])
        return out
    r, agg = ops.segment_matvec(msg, lay.qpos, plan.dst_ptr, lay.src_pad, x, None, None, None, None,
                                None, False, want_agg=True)
    return torch.stack([r, agg])


def _node_update(x, sums, inv_cnt, B3, root, bias, relu: bool):
    """x' = act(mean msg + (mean x_j) @ B3 + x @ root + bias) from the all-reduced sums: the fused
    node-update kernel with one-element segments (node n "reduces" row n of the summed messages
    and of the summed x_j), i.e. one pass over the four [N, w] tensors instead of six."""
    N = x.size(0)
    ident = _identity_segments(N, x.device)
    out, _ = ops.segment_matvec(sums[0], None, ident[0], ident[1], sums[1], inv_cnt, B3, x,
                                root.contiguous(), bias.contiguous(), relu)
    return out


_IDENT_CACHE: dict = {}


def _identity_segments(N: int, device):
    key = (N, str(device))
    if key not in _IDENT_CACHE:
        _IDENT_CACHE.clear()  # one graph size at a time
        _IDENT_CACHE[key] = (torch.arange(N + 1, dtype=torch.int32, device=device),
                             torch.arange(N, dtype=torch.int32, device=device))
    return _IDENT_CACHE[key]


def _node_shard_setup(shard: SourceShard, net, W: int):
    l0, l2, l4 = net.layers[0], net.layers[2], net.layers[4]
    K = l4.weight.size(1)
    if not GF.node_path_supported(W, K, 1, "bf16", needs_grad=False):
        raise ValueError(f"node-first sharding needs bf16, w in (32, 64, 128), K % 256 == 0 "
                         f"(got w = {W}, K = {K})")
    w3r = l4.weight.to(torch.bfloat16).contiguous().view(W, W * K)
    B3 = l4.bias.view(W, W).contiguous()
    if shard.lay is None:  # this rank owns no edges
        return None, w3r, B3, None
    lay = shard.lay
    ea_s = torch.zeros((lay.e_pad, shard.edge_attr.size(1)), dtype=torch.float32,
                       device=shard.edge_attr.device)
    ea_s.index_copy_(0, lay.qpos_long, shard.edge_attr.contiguous())
    _, hs = GF._edge_hidden(ea_s, l0.weight, l0.bias.contiguous(), l2.weight, l2.bias.contiguous(),
                            "bf16")
    msg = torch.empty((lay.e_pad, W), dtype=torch.float32, device=hs.device)
    return hs, w3r, B3, msg


# The exchange of the node-first sharding (GKN_SHARD_EXCHANGE).  "multimem" (default): as "peer" below
# but through the NVLS multicast addresses of the symmetric buffers - the NVSwitch adds the ranks'
# rows (multimem.ld_reduce) and replicates the new rows (multimem.st); falls back to "peer" when the
# fabric offers no multicast.  "peer" (when torch's symmetric memory can be set up for the group): partial sums live in symmetric-memory buffers and ONE kernel per rank reduces
# the rows of its node range straight out of every peer's buffer, applies the node update and
# stores the new rows into every peer's x' buffer (gkn_shard_reduce_update) - reduce-scatter,
# compute and all-gather fused over NVLink.  "nccl": ncclAllReduce of the partial sums, then the
# node update on every rank (round 1).  GKN_SHARD_EXCHANGE=nccl selects the latter.
import os as _os

_PEER_CACHE: dict = {}


class PeerExchange:
    """Symmetric-memory buffers of one (N, W, group): partial sums [2, N, W] and two x' buffers
    (double buffered: a fast rank may already be storing x of application a+1 into its peers while
    a slow one still reads x of application a)."""

    def __init__(self, N: int, W: int, device, group):
        import torch.distributed._symmetric_memory as symm

        g = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(g), dist.get_world_size(g)
        self.N, self.W = N, W
        self.part = symm.empty((2, N, W), dtype=torch.float32, device=device)
        self.h_part = symm.rendezvous(self.part, g)
        self.xbuf, self.h_x = [], []
        for _ in range(2):
            t = symm.empty((N, W), dtype=torch.float32, device=device)
            self.xbuf.append(t)
            self.h_x.append(symm.rendezvous(t, g))
        rows = (N + self.world - 1) // self.world
        self.n0, self.n1 = min(N, self.rank * rows), min(N, (self.rank + 1) * rows)
        self.part_ptrs = list(self.h_part.buffer_ptrs)
        self.x_ptrs = [list(h.buffer_ptrs) for h in self.h_x]
        # NVLS: multicast addresses (0 when the fabric offers none).  GKN_SHARD_EXCHANGE=peer keeps
        # the explicit per-peer loads / stores
        use_mc = _os.environ.get("GKN_SHARD_EXCHANGE", "multimem") == "multimem"
        self.mc_part = int(getattr(self.h_part, "multicast_ptr", 0) or 0) if use_mc else 0
        self.mc_x = [int(getattr(h, "multicast_ptr", 0) or 0) if use_mc else 0 for h in self.h_x]
        if not (self.mc_part and all(self.mc_x)):
            self.mc_part, self.mc_x = 0, [0, 0]

    def barrier(self):
        self.h_part.barrier()


def peer_exchange(N: int, W: int, device, group):
    """The cached PeerExchange of (N, W, group), or None when symmetric memory is unavailable (then
    the NCCL all-reduce is used).  Collective: every rank of the group must call it."""
    if _os.environ.get("GKN_SHARD_EXCHANGE", "multimem") not in ("peer", "multimem"):
        return None
    key = (N, W, str(device), id(group))
    if key not in _PEER_CACHE:
        try:
            _PEER_CACHE.clear()  # one frame size at a time
            _PEER_CACHE[key] = PeerExchange(N, W, device, group)
        except Exception as exc:  # noqa: BLE001  (no symmetric memory on this system / backend)
            import warnings

            warnings.warn(f"peer-memory exchange unavailable ({type(exc).__name__}: {exc}); "
                          "using the NCCL all-reduce")
            _PEER_CACHE[key] = None
    return _PEER_CACHE[key]


def exchange_in_use() -> str:
    """Which exchange the last sharded node-first forward used on this rank."""
    live = [v for v in _PEER_CACHE.values() if v is not None]
    if not live:
        return "ncclAllReduce of the partial sums + replicated node update"
    if live[0].mc_part:
        return ("fused kernel over NVLS multicast (gkn_shard_reduce_update: multimem.ld_reduce / "
                "multimem.st on symmetric memory)")
    return "fused peer-memory kernel (gkn_shard_reduce_update over symmetric memory)"


@torch.no_grad()
def sharded_stack_forward_node(x0, shard: SourceShard, net, root1, bias1, root2, bias2, schedule,
                               relu: bool, group=None) -> torch.Tensor:
    """Forward of the conv stack, node-first kernels, source-range sharding.  Every rank passes the
    same full ``x0`` [N, w] and gets the full result back."""
    N, W = x0.shape
    hs, w3r, B3, msg = _node_shard_setup(shard, net, W)
    roots, biases = (root1, root2), (bias1, bias2)
    x = x0.contiguous()
    px = peer_exchange(N, W, x.device, group) if dist.get_world_size(group) > 1 else None
    for a, c in enumerate(schedule):
        # with the peer exchange the partial rows are written where the peers can read them
        if hs is not None:
            sums = _node_partial_sums(x, shard, hs, w3r, msg, out=px.part if px is not None else None)
        elif px is not None:
            sums = px.part.zero_()
        else:
            sums = torch.zeros((2, N, W), dtype=torch.float32, device=x.device)
        if px is None:
            dist.all_reduce(sums, group=group)
            x = _node_update(x, sums, shard.inv_cnt, B3, roots[c], biases[c], relu)
            continue
        px.barrier()                 # every rank's partial sums are in place
        ops.shard_reduce_update(px.part_ptrs, px.x_ptrs[a & 1], x, shard.inv_cnt, B3,
                                roots[c].contiguous(), biases[c].contiguous(), px.n0, px.n1, relu,
                                mc_partial=px.mc_part, mc_xout=px.mc_x[a & 1])
        px.barrier()                 # every row of x' has landed everywhere
        x = px.xbuf[a & 1]
    return x.clone() if px is not None else x


@torch.no_grad()
def emulate_world_node(x0, edge_index, edge_attr, net, root1, bias1, root2, bias2, schedule,
                       relu: bool, world: int, row_ptr=None, fused_exchange: bool = False) -> torch.Tensor:
    """All ranks of sharded_stack_forward_node in turn on one GPU; the all-reduce is a sum, or
    (``fused_exchange``) the peer-memory kernel run once per emulated rank."""
    N, W = x0.shape
    shards = [shard_frame_by_source(edge_index, edge_attr, N, r, world, row_ptr)
              for r in range(world)]
    setups = [_node_shard_setup(s, net, W) for s in shards]
    roots, biases = (root1, root2), (bias1, bias2)
    x = x0.contiguous()
    rows = (N + world - 1) // world
    for c in schedule:
        parts = [(_node_partial_sums(x, s, hs, w3r, msg) if hs is not None else
                  torch.zeros((2, N, W), dtype=torch.float32, device=x.device))
                 for s, (hs, w3r, B3, msg) in zip(shards, setups)]
        if fused_exchange:
            # what every rank's gkn_shard_reduce_update launch does, one emulated rank after the
            # other: "peer" buffers are the other ranks' tensors on this same GPU
            xouts = [torch.empty_like(x) for _ in range(world)]
            for r in range(world):
                ops.shard_reduce_update([p.data_ptr() for p in parts], [t.data_ptr() for t in xouts],
                                        x, shards[0].inv_cnt, setups[0][2], roots[c].contiguous(),
                                        biases[c].contiguous(), min(N, r * rows),
                                        min(N, (r + 1) * rows), relu)
            for t in xouts[1:]:
                assert torch.equal(t, xouts[0])  # every rank ends up with the same x'
            x = xouts[0]
        else:
            sums = torch.stack(parts).sum(0)
            x = _node_update(x, sums, shards[0].inv_cnt, setups[0][2], roots[c], biases[c], relu)
    return x


@torch.no_grad()
def kernelnn_forward_sharded(model, data, group=None) -> torch.Tensor:
    """``KernelNN.forward`` for one large frame with the conv stack sharded over the ranks of
    ``group`` (node encoder and fc2 are tiny and replicated).  Every rank returns the full output."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    x = model.node_encoder(data)
    schedule = (0,) * model.depth + (1,) * model.depth
    K = model.conv1.net.layers[4].weight.size(1)
    if (GF.FACTORIZATION == "node" and data.edge_index.size(1) > 0
            and GF.node_path_supported(x.size(1), K, 1, model.precision, needs_grad=False)):
        # node-first kernels: source ranges, one all-reduce of the partial node sums per application
        full = getattr(data, "_gkn_plan", None)
        shard = shard_frame_by_source(data.edge_index, data.edge_attr, x.size(0), rank, world,
                                      getattr(data, "_gkn_row_ptr", None),
                                      None if full is None else full.inv_cnt)
        x = sharded_stack_forward_node(x.contiguous(), shard, model.conv1.net, model.conv1.root,
                                       model.conv1.bias, model.conv2.root, model.conv2.bias,
                                       schedule, relu=True, group=group)
        return model.fc2(x)
    plan = getattr(data, "_gkn_plan", None)
    if plan is None:
        plan = ops.prepare_graph(data.edge_index, x.size(0))
    ea = data.edge_attr.index_select(0, plan.perm_long)
    shard = shard_frame(plan, ea, rank, world)
    x = sharded_stack_forward(x.contiguous(), shard, model.conv1.net, model.conv1.root,
                              model.conv1.bias, model.conv2.root, model.conv2.bias, schedule,
                              relu=True, precision=model.precision, group=group)
    return model.fc2(x)
