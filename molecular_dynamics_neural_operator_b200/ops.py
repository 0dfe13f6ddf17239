"""Tensor-level wrappers over the C ABI (``include/gkn_b200.h``).

PyTorch is used here only for device memory and streams: every function checks its arguments,
allocates outputs with ``torch.empty`` and enqueues the CUDA kernels of libgkn_b200.so on the
current stream.  CPU tensors are rejected — there is no fallback path.
"""
from __future__ import annotations

from ctypes import c_void_p
from dataclasses import dataclass

import torch

from . import _lib


class _KernelTimer:
    """Optional CUDA-event timing of the contraction launches on their own stream (bench.py's
    live roofline measurement).  Disabled by default: zero overhead on the normal path."""

    def __init__(self):
        self.enabled = False
        self.events: list = []
        self.tagged: dict = {}

    def enable(self, tag: str = "contract"):
        self.enabled = True
        self.events = []
        self.tagged = {}

    def disable(self):
        self.enabled = False

    def collect(self) -> dict:
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self.events)
        out = {"count": len(self.events), "ms": ms}
        for tag, evs in self.tagged.items():
            out[tag] = {"count": len(evs), "ms": sum(a.elapsed_time(b) for a, b in evs)}
        return out


KERNEL_TIMER = _KernelTimer()


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: the GKN B200 kernels need CUDA tensors (no CPU fallback)")


def _p(t: torch.Tensor | None, dtype: torch.dtype | None = None, name: str = "tensor") -> c_void_p:
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError(f"{name}: the GKN B200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return c_void_p(t.data_ptr())


@dataclass
class GraphPlan:
    """Destination-sorted CSR view of an edge list (see gkn_graph_prepare in gkn_b200.h)."""

    num_nodes: int
    num_edges: int
    perm: torch.Tensor  # int32 [E]  sorted position -> original edge id
    src: torch.Tensor  # int32 [E]  source node per sorted position
    dst: torch.Tensor  # int32 [E]  destination node per sorted position
    dst_ptr: torch.Tensor  # int32 [N+1]
    inv_cnt: torch.Tensor  # float32 [N]  1 / max(in-degree, 1)
    src_ptr: torch.Tensor  # int32 [N+1]
    src_perm: torch.Tensor  # int32 [E]  sorted positions grouped by source node

    @property
    def perm_long(self) -> torch.Tensor:
        if not hasattr(self, "_perm_long"):
            self._perm_long = self.perm.long()
        return self._perm_long


def prepare_graph(edge_index: torch.Tensor, num_nodes: int, validate: bool = True) -> GraphPlan:
    """Build the destination-sorted CSR metadata for ``edge_index`` (int64 [2, E])."""
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError(f"edge_index must have shape [2, E], got {tuple(edge_index.shape)}")
    if edge_index.dtype != torch.int64:
        raise TypeError(f"edge_index must be int64 (torch.long), got {edge_index.dtype}")
    _require_cuda(edge_index, "edge_index")
    lib = _lib.load()
    ei = edge_index.contiguous()
    dev = ei.device
    N, E = int(num_nodes), int(ei.size(1))
    i32 = dict(dtype=torch.int32, device=dev)
    perm = torch.empty(E, **i32)
    src = torch.empty(E, **i32)
    dst = torch.empty(E, **i32)
    dst_ptr = torch.empty(N + 1, **i32)
    src_ptr = torch.empty(N + 1, **i32)
    src_perm = torch.empty(E, **i32)
    inv_cnt = torch.empty(N, dtype=torch.float32, device=dev)
    status = torch.zeros(1, **i32)
    ws_bytes = int(lib.gkn_graph_prepare_workspace_bytes(N, E))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.gkn_graph_prepare(
                _p(ei, torch.int64, "edge_index"), N, E, _p(perm), _p(src), _p(dst), _p(dst_ptr),
                _p(inv_cnt), _p(src_ptr), _p(src_perm), _p(status), _p(ws), ws_bytes, _stream(),
            ),
            "gkn_graph_prepare",
        )
    if validate and int(status.item()) != 0:
        raise ValueError(f"edge_index contains node ids outside [0, {N})")
    return GraphPlan(N, E, perm, src, dst, dst_ptr, inv_cnt, src_ptr, src_perm)


def plan_from_csr(edge_index: torch.Tensor, row_ptr: torch.Tensor, num_nodes: int,
                  validate: bool = True) -> GraphPlan:
    """GraphPlan of a symmetric (row, col)-sorted graph with int64 row pointers — the output of
    ``radius_graph(..., return_row_ptr=True)`` — without any sorting (gkn_plan_from_csr)."""
    _require_cuda(edge_index, "edge_index")
    lib = _lib.load()
    ei = edge_index.contiguous()
    dev = ei.device
    N, E = int(num_nodes), int(ei.size(1))
    i32 = dict(dtype=torch.int32, device=dev)
    rev, src, dst = torch.empty(E, **i32), torch.empty(E, **i32), torch.empty(E, **i32)
    ptr32 = torch.empty(N + 1, **i32)
    inv_cnt = torch.empty(N, dtype=torch.float32, device=dev)
    status = torch.zeros(1, **i32)
    with torch.cuda.device(dev):
        _lib.check(
            lib.gkn_plan_from_csr(_p(ei, torch.int64, "edge_index"), _p(row_ptr, torch.int64, "row_ptr"),
                                  N, E, _p(rev), _p(src), _p(dst), _p(ptr32), _p(inv_cnt),
                                  _p(status), _stream()),
            "gkn_plan_from_csr",
        )
    if validate and int(status.item()) != 0:
        raise ValueError("plan_from_csr: edge_index is not a symmetric (row, col)-sorted graph")
    return GraphPlan(N, E, rev, src, dst, ptr32, inv_cnt, ptr32, rev)


def radius_graph(pos: torch.Tensor, threshold: float, return_row_ptr: bool = False):
    """Batched radius graph.  pos float32 [F, N, 3] -> (edge_index int64 [2, E], edge_attr [E, 6])
    (and the int64 CSR row pointers [F*N + 1] with ``return_row_ptr``).

    Node ids of frame f are offset by f*N; edges are in (row, col) lexicographic order with
    self loops, identical to scipy distance_matrix(float64) < threshold -> coo_matrix.
    """
    if pos.dim() == 2:
        pos = pos.unsqueeze(0)
    if pos.dim() != 3 or pos.size(-1) != 3:
        raise ValueError(f"pos must have shape [F, N, 3], got {tuple(pos.shape)}")
    _require_cuda(pos, "pos")
    lib = _lib.load()
    pos = pos.contiguous()
    dev = pos.device
    F, N = int(pos.size(0)), int(pos.size(1))
    ws_bytes = int(lib.gkn_radius_graph_workspace_bytes(F, N))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    deg = torch.empty(F * N, dtype=torch.int32, device=dev)
    row_ptr = torch.empty(F * N + 1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.gkn_radius_graph_count(
                _p(pos, torch.float32, "pos"), F, N, float(threshold), _p(deg), _p(row_ptr),
                _p(ws), ws_bytes, _stream(),
            ),
            "gkn_radius_graph_count",
        )
        E = int(row_ptr[-1].item())  # output size is data dependent: one sync per build
        edge_index = torch.empty((2, E), dtype=torch.int64, device=dev)
        edge_attr = torch.empty((E, 6), dtype=torch.float32, device=dev)
        _lib.check(
            lib.gkn_radius_graph_fill(
                _p(pos), F, N, float(threshold), _p(row_ptr), E, _p(edge_index), _p(edge_attr),
                _p(ws), ws_bytes, _stream(),
            ),
            "gkn_radius_graph_fill",
        )
    if return_row_ptr:
        return edge_index, edge_attr, row_ptr
    return edge_index, edge_attr


def edge_attr_gather(pos: torch.Tensor, edge_index: torch.Tensor, validate: bool = True) -> torch.Tensor:
    """edge_attr [E, 6] = [pos[row], pos[col]] for node coordinates pos float32 [M, 3] (a batch of
    graphs: M = total nodes, edge_index already offset) - gkn_edge_attr_gather, the feeder's
    replacement for the reference's per-edge Python loop (dataset.py:194-201)."""
    lib = _lib.load()
    _require_cuda(pos, "pos")
    pos = pos.reshape(-1, 3).contiguous()
    ei = edge_index.contiguous()
    if ei.dim() != 2 or ei.size(0) != 2 or ei.dtype != torch.int64:
        raise ValueError("edge_index must be int64 [2, E]")
    E, M = int(ei.size(1)), int(pos.size(0))
    out = torch.empty((E, 6), dtype=torch.float32, device=pos.device)
    status = torch.zeros(1, dtype=torch.int32, device=pos.device)
    with torch.cuda.device(pos.device):
        _lib.check(lib.gkn_edge_attr_gather(_p(pos, torch.float32, "pos"), _p(ei, torch.int64, "edge_index"),
                                            E, M, _p(out), _p(status), _stream()),
                   "gkn_edge_attr_gather")
    if validate and int(status.item()) != 0:
        raise ValueError(f"edge_index contains node ids outside [0, {M})")
    return out


def _contract(kind: str, h, w3, idx, node, precision: str) -> torch.Tensor:
    _require_cuda(h, "h")
    lib = _lib.load()
    E, K = int(h.size(0)), int(h.size(1))
    W = int(node.size(1))
    if w3.shape != (W * W, K):
        raise ValueError(f"w3 must be [{W * W}, {K}], got {tuple(w3.shape)}")
    out = torch.empty((E, W), dtype=torch.float32, device=h.device)
    if precision == "bf16":
        fn = lib.gkn_contract_fwd_bf16 if kind == "fwd" else lib.gkn_contract_dx_bf16
        dt = torch.bfloat16
    elif precision == "fp32":
        fn = lib.gkn_contract_fwd_f32 if kind == "fwd" else lib.gkn_contract_dx_f32
        dt = torch.float32
    else:
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
    with torch.cuda.device(h.device):
        if KERNEL_TIMER.enabled:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        _lib.check(
            fn(_p(h, dt, "h"), _p(w3, dt, "w3"), _p(idx, torch.int32, "idx"),
               _p(node, torch.float32, "node"), _p(out), E, K, W, _stream()),
            f"gkn_contract_{kind}_{precision}",
        )
        if KERNEL_TIMER.enabled:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            KERNEL_TIMER.events.append((ev0, ev1))
    return out


def contract_fwd(h, w3, src, x, precision: str = "bf16") -> torch.Tensor:
    """msg[e,o] = sum_i x[src[e],i] * sum_k h[e,k] w3[i*W+o,k]   (per-edge kernel never stored)."""
    return _contract("fwd", h, w3, src, x, precision)


def contract_dx(h, w3, dst, g, precision: str = "bf16") -> torch.Tensor:
    """dxj[e,i] = sum_o (sum_k h[e,k] w3[i*W+o,k]) * g[dst[e],o]."""
    return _contract("dx", h, w3, dst, g, precision)


def scatter_mean(msg, plan: GraphPlan, xb3=None, xroot=None, bias=None, relu=False) -> torch.Tensor:
    lib = _lib.load()
    W = int(msg.size(1))
    out = torch.empty((plan.num_nodes, W), dtype=torch.float32, device=msg.device)
    f32 = torch.float32
    with torch.cuda.device(msg.device):
        _lib.check(
            lib.gkn_scatter_mean_fwd(
                _p(msg, f32, "msg"), _p(plan.src), _p(plan.dst_ptr), _p(plan.inv_cnt),
                _p(xb3, f32, "xb3"), _p(xroot, f32, "xroot"), _p(bias, f32, "bias"), _p(out),
                plan.num_nodes, plan.num_edges, W, int(bool(relu)), _stream(),
            ),
            "gkn_scatter_mean_fwd",
        )
    return out


def scatter_src_sum(dxj, plan: GraphPlan, gb3=None, addend=None) -> torch.Tensor:
    lib = _lib.load()
    W = int(dxj.size(1))
    out = torch.empty((plan.num_nodes, W), dtype=torch.float32, device=dxj.device)
    f32 = torch.float32
    with torch.cuda.device(dxj.device):
        _lib.check(
            lib.gkn_scatter_src_sum(
                _p(dxj, f32, "dxj"), _p(plan.src_ptr), _p(plan.src_perm), _p(plan.dst),
                _p(gb3, f32, "gb3"), _p(addend, f32, "addend"), _p(out),
                plan.num_nodes, plan.num_edges, W, _stream(),
            ),
            "gkn_scatter_src_sum",
        )
    return out


def relu_grad_scale(gy, y, inv_cnt, relu: bool) -> tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    N, W = int(gy.size(0)), int(gy.size(1))
    g = torch.empty_like(gy)
    gn = torch.empty_like(gy)
    f32 = torch.float32
    with torch.cuda.device(gy.device):
        _lib.check(
            lib.gkn_relu_grad_scale(
                _p(gy, f32, "gy"), _p(y if relu else None, f32, "y"), _p(inv_cnt, f32, "inv_cnt"),
                _p(g), _p(gn), N, W, int(bool(relu)), _stream(),
            ),
            "gkn_relu_grad_scale",
        )
    return g, gn


def edge_outer_sum(xs, gs, plan: GraphPlan, e0: int, e1: int, out_dtype=torch.bfloat16) -> torch.Tensor:
    """S[e-e0, i*W+o] = sum_a xs[a, src[e], i] * gs[a, dst[e], o] for sorted positions [e0, e1)."""
    lib = _lib.load()
    A, N, W = (int(s) for s in xs.shape)
    if gs.shape != xs.shape:
        raise ValueError("xs and gs must have the same [A, N, W] shape")
    if out_dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("out_dtype must be bfloat16 or float32")
    S = torch.empty((e1 - e0, W * W), dtype=out_dtype, device=xs.device)
    with torch.cuda.device(xs.device):
        _lib.check(
            lib.gkn_edge_outer_sum(
                _p(xs, torch.float32, "xs"), _p(gs, torch.float32, "gs"), _p(plan.src),
                _p(plan.dst), _p(S), int(out_dtype == torch.bfloat16), e0, e1, N, A, W, _stream(),
            ),
            "gkn_edge_outer_sum",
        )
    return S


EDGE_GRAD_MAX_SLOTS = 8


def edge_grad_slots(plan: GraphPlan) -> int:
    """Largest number of distinct destinations inside any 128-edge tile of the plan (cached on the
    plan; one small kernel and a 4-byte read).  The fused edge-gradient kernels need <= 8."""
    cached = getattr(plan, "_edge_grad_slots", None)
    if cached is not None:
        return cached
    lib = _lib.load()
    status = torch.zeros(1, dtype=torch.int32, device=plan.dst.device)
    with torch.cuda.device(plan.dst.device):
        _lib.check(lib.gkn_edge_grad_max_slots(_p(plan.dst), plan.num_edges, _p(status), _stream()),
                   "gkn_edge_grad_max_slots")
    val = int(status.item())
    object.__setattr__(plan, "_edge_grad_slots", val)
    return val


def edge_grad_dh(xs, gs, plan: GraphPlan, w3t: torch.Tensor) -> torch.Tensor:
    """dh[e, k] = sum_n S[e, n] * W3[n, k] (bf16 [E, K]) with S = sum_a xs[a, src] (x) gs[a, dst]
    synthesised inside the tcgen05 kernel - S is never stored.  w3t: bf16 [K, W*W] (W3 transposed)."""
    lib = _lib.load()
    A, N, W = (int(s) for s in xs.shape)
    if gs.shape != xs.shape:
        raise ValueError("xs and gs must have the same [A, N, W] shape")
    K = int(w3t.size(0))
    if w3t.shape != (K, W * W):
        raise ValueError(f"w3t must be [K, {W * W}], got {tuple(w3t.shape)}")
    E = plan.num_edges
    if edge_grad_slots(plan) > EDGE_GRAD_MAX_SLOTS:
        raise ValueError("edge_grad_dh: a 128-edge tile spans more than 8 destination nodes "
                         "(use the staged path for graphs of such low degree)")
    dh = torch.empty((E, K), dtype=torch.bfloat16, device=xs.device)
    ws_bytes = int(lib.gkn_edge_grad_workspace_bytes(N, W))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=xs.device)
    with torch.cuda.device(xs.device):
        _lib.check(
            lib.gkn_edge_grad_dh_bf16(
                _p(xs, torch.float32, "xs"), _p(gs, torch.float32, "gs"), _p(plan.src), _p(plan.dst),
                _p(w3t, torch.bfloat16, "w3t"), _p(dh), E, N, A, K, W, _p(ws), ws_bytes, _stream(),
            ),
            "gkn_edge_grad_dh_bf16",
        )
    return dh


def bias_act_cast_bf16(x, bias=None, relu=False) -> torch.Tensor:
    lib = _lib.load()
    rows, cols = int(x.size(0)), int(x.size(1))
    out = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(
            lib.gkn_bias_act_cast_bf16(
                _p(x, torch.float32, "x"), _p(bias, torch.float32, "bias"), _p(out), rows, cols,
                int(bool(relu)), _stream(),
            ),
            "gkn_bias_act_cast_bf16",
        )
    return out


def mask_cast_colsum(dz, act, colsum, out_dtype, side=None, accumulate: bool = True,
                     sums_only: bool = False) -> torch.Tensor:
    """out = dz * (act > 0) in ``out_dtype`` (bf16 / fp32); colsum[0] (+)= out.sum(0) and, with
    ``side`` [rows, 6], colsum[1+j] (+)= (out * side[:, j:j+1]).sum(0).  ``colsum`` is
    [(1 + nside), cols] (or [cols] without side).  One pass over dz, deterministic reduction.
    ``sums_only``: do not write ``out`` (returns None)."""
    lib = _lib.load()
    rows, cols = int(dz.size(0)), int(dz.size(1))
    nside = 0 if side is None else int(side.size(1))
    if colsum.numel() != (1 + nside) * cols:
        raise ValueError("colsum must have (1 + nside) * cols elements")
    # per-call scratch from the caching allocator (stream-ordered, so safe on any stream)
    ws = torch.empty(int(lib.gkn_mask_cast_colsum_workspace_bytes(cols, nside)), dtype=torch.uint8,
                     device=dz.device)
    if sums_only:
        out = None  # the kernel skips the stores
    else:
        out = dz if out_dtype == dz.dtype else torch.empty((rows, cols), dtype=out_dtype,
                                                          device=dz.device)
    with torch.cuda.device(dz.device):
        _lib.check(
            lib.gkn_mask_cast_colsum(
                _p(dz, None, "dz"), int(dz.dtype == torch.bfloat16), _p(act, None, "act"),
                int(act.dtype == torch.bfloat16), _p(out), int(out_dtype == torch.bfloat16),
                _p(side, torch.float32, "side"), nside, _p(colsum, torch.float32, "colsum"),
                int(bool(accumulate)), rows, cols, _p(ws), ws.numel(), _stream(),
            ),
            "gkn_mask_cast_colsum",
        )
    return out


def edge_layer0(ea, w1, b1, out_dtype) -> torch.Tensor:
    """relu(ea @ w1.T + b1) written directly as bf16 / fp32 [E, K] (gkn_edge_layer0)."""
    lib = _lib.load()
    _require_cuda(ea, "edge_attr")
    E, kin = int(ea.size(0)), int(ea.size(1))
    K = int(w1.size(0))
    out = torch.empty((E, K), dtype=out_dtype, device=ea.device)
    with torch.cuda.device(ea.device):
        _lib.check(
            lib.gkn_edge_layer0(_p(ea, torch.float32, "edge_attr"), _p(w1, torch.float32, "w1"),
                                _p(b1, torch.float32, "b1"), _p(out),
                                int(out_dtype == torch.bfloat16), E, K, kin, _stream()),
            "gkn_edge_layer0",
        )
    return out


def segment_matvec(rows, perm, seg_ptr, gidx, gnode, scale, M1, nin, M2, bias, relu: bool,
                   want_agg: bool = False, out=None, agg_out=None):
    """out[n] = act(scale*sum rows[p] + (scale*sum gnode[gidx[p]]) @ M1 + nin[n] @ M2 + bias).
    See gkn_segment_matvec in gkn_b200.h.  Returns (out, agg or None); ``out`` / ``agg_out``:
    caller-provided fp32 [N, W] buffers to write into."""
    lib = _lib.load()
    ref = nin if nin is not None else gnode  # M1 = M2 = nin = None: plain segment sums
    N, W = int(ref.size(0)), int(ref.size(1))
    f32, i32 = torch.float32, torch.int32
    if out is None:
        out = torch.empty((N, W), dtype=f32, device=ref.device)
    agg = agg_out if agg_out is not None else (
        torch.empty((N, W), dtype=f32, device=ref.device) if want_agg else None)
    with torch.cuda.device(ref.device):
        _lib.check(
            lib.gkn_segment_matvec(
                _p(rows, f32, "rows"), _p(perm, i32, "perm"), _p(seg_ptr, i32, "seg_ptr"),
                _p(gidx, i32, "gidx"), _p(gnode, f32, "gnode"), _p(scale, f32, "scale"),
                _p(M1, f32, "M1"), _p(nin, f32, "nin"), _p(M2, f32, "M2"), _p(bias, f32, "bias"),
                _p(out), _p(agg), N, W, int(bool(relu)), _stream(),
            ),
            "gkn_segment_matvec",
        )
    return out, agg


# ------------------------------------------------------------------------------------------------
# Node-first factorisation (csrc/gkn_node_gemm.cu): padded source-sorted edge layout + grouped GEMMs
# ------------------------------------------------------------------------------------------------
@dataclass
class NodeLayout:
    """Padded source-sorted edge layout of a GraphPlan (see gkn_node_msg_bf16 in gkn_b200.h):
    the edges of source node n occupy rows [pptr[n], pptr[n] + deg[n]); every group is padded to a
    multiple of 16 rows."""

    num_nodes: int
    e_pad: int
    num_units: int
    pptr: torch.Tensor  # int32 [N+1]
    units: torch.Tensor  # int32 [U, 4]  (source node, first row, rows to store, 0)
    qpos: torch.Tensor  # int32 [E]   destination-sorted position -> padded row
    src_pad: torch.Tensor  # int32 [E_pad]  source node of every padded row
    dst_pad: torch.Tensor  # int32 [E_pad]  destination node, -1 in pad rows
    unit_first: list  # host list [N+1]: first unit of every node (for node-range chunking)
    unit_first_dev: torch.Tensor = None  # the same on the device (int64 [N+1])

    @property
    def qpos_long(self) -> torch.Tensor:
        if not hasattr(self, "_qpos_long"):
            self._qpos_long = self.qpos.long()
        return self._qpos_long

    def tensors(self):
        extra = [self.unit_first_dev] if self.unit_first_dev is not None else []
        return [self.pptr, self.units, self.qpos, self.src_pad, self.dst_pad] + extra + \
            list(getattr(self, "_dh_units", {}).values()) + \
            [s.items for s in getattr(self, "_fused_sched", {}).values()]

    def dh_units(self, chunks: int) -> torch.Tensor:
        """int32 [U*chunks, 4] = (source node, first row, rows, column chunk) in (node, chunk,
        tile) order: the tiles of a group that share one chunk of the node's U are adjacent, so the
        second tile finds it in L2.  Node-major, so node ranges are contiguous unit ranges."""
        cache = self.__dict__.setdefault("_dh_units", {})
        if chunks not in cache:
            u = self.units.long()
            U = u.size(0)
            uf = self.unit_first_dev
            first = uf[u[:, 0]]  # first tile of the group
            t = torch.arange(U, device=u.device) - first
            c = torch.arange(chunks, device=u.device)
            tiles = uf.diff()[u[:, 0]]
            # destination slot of (tile u, chunk c): group's first slot + c * tiles(group) + t
            slot = (first * chunks)[:, None] + c[None, :] * tiles[:, None] + t[:, None]
            rec = torch.stack([u[:, 0:1].expand(U, chunks), u[:, 1:2].expand(U, chunks),
                               u[:, 2:3].expand(U, chunks), c[None, :].expand(U, chunks)], 2)
            out = torch.empty((U * chunks, 4), dtype=torch.int32, device=u.device)
            out[slot.reshape(-1)] = rec.reshape(-1, 4).int()
            cache[chunks] = out
        return cache[chunks]


def node_layout(plan: GraphPlan, dh_chunks: int | None = None) -> NodeLayout:
    """Build (once per plan, cached on it) the padded source-sorted layout.  Index bookkeeping
    only (torch ops on the device, two small host reads for the data-dependent sizes).
    ``dh_chunks`` (= kernel_width // 256) also prebuilds the unit table of the dh pass."""
    cached = getattr(plan, "_node_layout", None)
    if cached is not None:
        if dh_chunks:
            cached.dh_units(dh_chunks)
        return cached
    dev = plan.src.device
    N, E = plan.num_nodes, plan.num_edges
    sp = plan.src_ptr.long()
    deg = sp[1:] - sp[:-1]
    degp = (deg + 15) // 16 * 16
    tiles = (degp + 127) // 128
    zero = torch.zeros(1, dtype=torch.long, device=dev)
    pptr = torch.cat([zero, torch.cumsum(degp, 0)])
    ufirst = torch.cat([zero, torch.cumsum(tiles, 0)])
    # one host read for everything data dependent: the unit offsets and the padded edge count
    host = torch.cat([ufirst, pptr[-1:]]).tolist()
    unit_first, e_pad = host[:-1], host[-1]
    U = unit_first[-1]
    nodes = torch.arange(N, device=dev)
    grp = torch.repeat_interleave(nodes, deg, output_size=E)
    q_of_j = torch.arange(E, device=dev) + (pptr[:-1] - sp[:-1])[grp]
    p_of_j = plan.src_perm.long()
    qpos = torch.empty(E, dtype=torch.long, device=dev)
    qpos[p_of_j] = q_of_j
    src_pad = torch.repeat_interleave(nodes, degp, output_size=e_pad).int()
    dst_pad = torch.full((e_pad,), -1, dtype=torch.int32, device=dev)
    dst_pad[q_of_j] = plan.dst[p_of_j]
    ug = torch.repeat_interleave(nodes, tiles, output_size=U)
    t = torch.arange(U, device=dev) - ufirst[:-1][ug]
    row0 = pptr[:-1][ug] + 128 * t
    rows = torch.clamp(degp[ug] - 128 * t, max=128)
    units = torch.stack([ug, row0, rows, torch.zeros_like(ug)], 1).int().contiguous()
    lay = NodeLayout(N, e_pad, U, pptr.int(), units, qpos.int(), src_pad, dst_pad, unit_first,
                     ufirst)
    object.__setattr__(plan, "_node_layout", lay)
    if dh_chunks:
        lay.dh_units(dh_chunks)
    return lay


def _timed(tag: str):
    """Context manager recording CUDA events around a launch when KERNEL_TIMER is enabled."""
    class _T:
        def __enter__(self_inner):
            if KERNEL_TIMER.enabled:
                self_inner.ev0 = torch.cuda.Event(enable_timing=True)
                self_inner.ev0.record()

        def __exit__(self_inner, *exc):
            if KERNEL_TIMER.enabled:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record()
                KERNEL_TIMER.tagged.setdefault(tag, []).append((self_inner.ev0, ev1))
            return False

    return _T()


def node_msg(hs, u, lay: NodeLayout, unit_begin: int, unit_end: int, node_base: int,
             msg: torch.Tensor) -> None:
    """msg[q, :] = hs[q, :] @ u[src q]^T for the units [unit_begin, unit_end) (gkn_node_msg_bf16).
    hs bf16 [E_pad, K]; u bf16 [chunk_nodes, W, K] for nodes [node_base, ...); msg fp32 [E_pad, W]."""
    lib = _lib.load()
    K = int(hs.size(1))
    W = int(msg.size(1))
    with torch.cuda.device(hs.device), _timed("node_msg"):
        _lib.check(
            lib.gkn_node_msg_bf16(_p(hs, torch.bfloat16, "hs"), _p(u, torch.bfloat16, "u"),
                                  _p(lay.units, torch.int32, "units"), unit_begin, unit_end,
                                  lay.e_pad, int(u.size(0)), node_base,
                                  _p(msg, torch.float32, "msg"), K, W, _stream()),
            "gkn_node_msg_bf16",
        )


def node_du(hsT, gsT, lay: NodeLayout, node_begin: int, node_end: int, dut: torch.Tensor) -> None:
    """dut[n - node_begin][k, o] = sum_{q in group n} hsT[k, q] * gsT[o, q]  (gkn_node_du_bf16)."""
    lib = _lib.load()
    K = int(hsT.size(0))
    W = int(gsT.size(0))
    with torch.cuda.device(hsT.device), _timed("node_du"):
        _lib.check(
            lib.gkn_node_du_bf16(_p(hsT, torch.bfloat16, "hsT"), _p(gsT, torch.bfloat16, "gsT"),
                                 _p(lay.pptr, torch.int32, "pptr"), node_begin, node_end, lay.e_pad,
                                 _p(dut, torch.bfloat16, "dut"), K, W, _stream()),
            "gkn_node_du_bf16",
        )


def node_dh(gs, ut, lay: NodeLayout, unit_begin: int, unit_end: int, node_base: int,
            dh: torch.Tensor) -> None:
    """dh[q, k] = sum_a gs[a][q, :] . ut[a][src q][k, :]  (gkn_node_dh_bf16).
    gs bf16 [A, E_pad, W]; ut bf16 [A, chunk_nodes, K, W]; dh bf16 [E_pad, K]."""
    lib = _lib.load()
    A, W = int(gs.size(0)), int(gs.size(2))
    K = int(dh.size(1))
    with torch.cuda.device(gs.device), _timed("node_dh"):
        _lib.check(
            lib.gkn_node_dh_bf16(_p(gs, torch.bfloat16, "gs"), _p(ut, torch.bfloat16, "ut"),
                                 _p(lay.dh_units(K // 256), torch.int32, "units"), unit_begin, unit_end,
                                 lay.e_pad, int(ut.size(1)), node_base, A,
                                 _p(dh, torch.bfloat16, "dh"), K, W, _stream()),
            "gkn_node_dh_bf16",
        )


def gs_width(W: int) -> int:
    """Row length of the padded-order bf16 gradient buffer gs: at least one 128-byte swizzle row
    (w = 32: 64 values, the upper half zeros - written by ``node_gather``)."""
    return max(int(W), 64)


def node_du_mn(hs, gs, lay: NodeLayout, node_begin: int, node_end: int, dut: torch.Tensor) -> None:
    """node_du from the row-major hs [E_pad, K] and gs [E_pad, gs_width(W)] into dut
    [nodes, K * W] (gkn_node_du_mn_bf16; w in {32, 64, 128})."""
    lib = _lib.load()
    K = int(hs.size(1))
    W = int(dut.size(1)) // K
    if gs.size(1) != gs_width(W):
        raise ValueError(f"gs must have {gs_width(W)} columns at w = {W}")
    with torch.cuda.device(hs.device), _timed("node_du"):
        _lib.check(
            lib.gkn_node_du_mn_bf16(_p(hs, torch.bfloat16, "hs"), _p(gs, torch.bfloat16, "gs"),
                                    _p(lay.pptr, torch.int32, "pptr"), node_begin, node_end,
                                    lay.e_pad, _p(dut, torch.bfloat16, "dut"), K, W, _stream()),
            "gkn_node_du_mn_bf16",
        )


def relu_bits(act: torch.Tensor) -> torch.Tensor:
    """One bit per element of a bf16 activation matrix [rows, K]: int32 [rows, K // 32], bit k % 32
    of word k // 32 set where act[q, k] > 0 (gkn_relu_bits_bf16)."""
    lib = _lib.load()
    rows, K = int(act.size(0)), int(act.size(1))
    bits = torch.empty((rows, K // 32), dtype=torch.int32, device=act.device)
    with torch.cuda.device(act.device), _timed("relu_bits"):
        _lib.check(lib.gkn_relu_bits_bf16(_p(act, torch.bfloat16, "act"), rows, K,
                                          _p(bits, torch.int32, "bits"), _stream()),
                   "gkn_relu_bits_bf16")
    return bits


def node_dh_mn(gs, u, lay: NodeLayout, unit_begin: int, unit_end: int, node_base: int,
               dh: torch.Tensor, act_bits: torch.Tensor | None = None,
               colsum: torch.Tensor | None = None, accumulate: bool = False) -> None:
    """node_dh from gs [A, E_pad, gs_width(W)] and u bf16 [A, chunk_nodes, W, K] (the forward's U)
    (gkn_node_dh_mn_bf16; w in {32, 64, 128}).

    With ``act_bits`` (``relu_bits(hs)``: int32 [E_pad, K // 32]) the ReLU backward of DenseNet
    layers.2 is applied in the epilogue: ``dh`` receives dz2 = dh * (hs > 0) and ``colsum`` (fp32
    [K]) its column sums over the launch's rows (gkn_node_dh_mn_relu_bf16; ``accumulate`` adds to
    ``colsum``)."""
    lib = _lib.load()
    A, W = int(gs.size(0)), int(u.size(2))
    K = int(dh.size(1))
    if gs.size(2) != gs_width(W):
        raise ValueError(f"gs must have {gs_width(W)} columns at w = {W}")
    common = (_p(gs, torch.bfloat16, "gs"), _p(u, torch.bfloat16, "u"),
              _p(lay.dh_units(K // 256), torch.int32, "units"), unit_begin, unit_end, lay.e_pad,
              int(u.size(1)), node_base, A)
    with torch.cuda.device(gs.device), _timed("node_dh"):
        if act_bits is None:
            _lib.check(lib.gkn_node_dh_mn_bf16(*common, _p(dh, torch.bfloat16, "dh"), K, W,
                                               _stream()), "gkn_node_dh_mn_bf16")
            return
        if tuple(act_bits.shape) != (dh.size(0), K // 32) or colsum is None or colsum.numel() != K:
            raise ValueError("act_bits must be [E_pad, K // 32] and colsum have K entries")
        ws = torch.empty((device_sm_count(), K), dtype=torch.float32, device=gs.device)
        _lib.check(
            lib.gkn_node_dh_mn_relu_bf16(*common, _p(act_bits, torch.int32, "act_bits"),
                                         _p(dh, torch.bfloat16, "dh"),
                                         _p(colsum, torch.float32, "colsum"), int(accumulate),
                                         _p(ws), ws.numel() * 4, K, W, _stream()),
            "gkn_node_dh_mn_relu_bf16",
        )


def xty(x: torch.Tensor, g: torch.Tensor, colsum: bool = False):
    """(x^T g [W, W], column sums of g [W] or None) for float32 x, g [rows, W] (gkn_xty_f32)."""
    lib = _lib.load()
    rows, W = int(x.size(0)), int(x.size(1))
    if g.shape != x.shape:
        raise ValueError("x and g must have the same shape")
    _require_cuda(x, "x")
    if W not in (16, 32, 64, 128):  # widths the reference never instantiates: library product
        return x.t() @ g, (g.sum(0) if colsum else None)
    f32 = dict(dtype=torch.float32, device=x.device)
    out = torch.empty((W, W), **f32)
    gsum = torch.empty(W, **f32) if colsum else None
    with torch.cuda.device(x.device), _timed("xty"):
        nbytes = int(lib.gkn_xty_workspace_bytes(W))
        ws = torch.empty(nbytes // 4, **f32)
        _lib.check(lib.gkn_xty_f32(_p(x, torch.float32, "x"), _p(g, torch.float32, "g"), rows, W,
                                   _p(out), _p(gsum), _p(ws), nbytes, _stream()), "gkn_xty_f32")
    return out, gsum


def node_gather(g, lay: NodeLayout, gs: torch.Tensor | None, gsT: torch.Tensor | None) -> None:
    """gs[q, :W] = bf16(g[dst_pad[q], :]) (zero in pad rows and in columns >= W; gs has
    gs_width(W) columns) and/or, at w = 64, its transpose gsT [W, E_pad]."""
    lib = _lib.load()
    W = int(g.size(1))
    if gs is not None and gs.size(1) != gs_width(W):
        raise ValueError(f"gs must have {gs_width(W)} columns at w = {W}")
    with torch.cuda.device(g.device):
        _lib.check(
            lib.gkn_node_gather_bf16(_p(g, torch.float32, "g"), _p(lay.dst_pad, torch.int32, "idx"),
                                     lay.e_pad, _p(gs, torch.bfloat16, "gs"),
                                     _p(gsT, torch.bfloat16, "gsT"), W, _stream()),
            "gkn_node_gather_bf16",
        )


# ------------------------------------------------------------------------------------------------
# Node encoder's LSTM (csrc/gkn_lstm.cu; reference graph_kernel.py:278-285)
# ------------------------------------------------------------------------------------------------
def lstm_fwd(x, w_ih, w_hh, b_ih, b_hh, save: bool):
    """h_t of nn.LSTM(D, D) (zero initial state) over x [T, R, D]; with ``save`` also the activated
    gates [T, R, 4D] and cell states [T, R, D] the backward needs (gkn_lstm_fwd)."""
    lib = _lib.load()
    T, R, D = (int(v) for v in x.shape)
    f32 = dict(dtype=torch.float32, device=x.device)
    out = torch.empty((T, R, D), **f32)
    gates = torch.empty((T, R, 4 * D), **f32) if save else None
    cells = torch.empty((T, R, D), **f32) if save else None
    with torch.cuda.device(x.device), _timed("lstm"):
        _lib.check(
            lib.gkn_lstm_fwd(_p(x, torch.float32, "x"), _p(w_ih, torch.float32, "w_ih"),
                             _p(w_hh, torch.float32, "w_hh"), _p(b_ih, torch.float32, "b_ih"),
                             _p(b_hh, torch.float32, "b_hh"), _p(out), _p(gates), _p(cells), T, R,
                             D, _stream()),
            "gkn_lstm_fwd",
        )
    return out, gates, cells


def lstm_bwd(x, w_ih, w_hh, out, gates, cells, dout, need_dx: bool):
    """Gradients of ``lstm_fwd``: (dx or None, dW_ih, dW_hh, db) with db the gradient of both bias
    vectors (gkn_lstm_bwd; per-block partial sums folded here)."""
    lib = _lib.load()
    T, R, D = (int(v) for v in x.shape)
    f32 = dict(dtype=torch.float32, device=x.device)
    blocks = int(lib.gkn_lstm_blocks(R))
    part = torch.empty((blocks, 8 * D * D + 4 * D), **f32)
    dx = torch.empty((T, R, D), **f32) if need_dx else None
    with torch.cuda.device(x.device), _timed("lstm"):
        _lib.check(
            lib.gkn_lstm_bwd(_p(x, torch.float32, "x"), _p(w_ih, torch.float32, "w_ih"),
                             _p(w_hh, torch.float32, "w_hh"), _p(out, torch.float32, "out"),
                             _p(gates, torch.float32, "gates"), _p(cells, torch.float32, "cells"),
                             _p(dout, torch.float32, "dout"), _p(dx), _p(part), T, R, D, _stream()),
            "gkn_lstm_bwd",
        )
    tot = part.sum(0)
    g = 4 * D * D
    return dx, tot[:g].view(4 * D, D), tot[g:2 * g].view(4 * D, D), tot[2 * g:]


# ------------------------------------------------------------------------------------------------
# Fused producer/consumer forward of the node-first path (csrc/gkn_node_fused.cu)
# ------------------------------------------------------------------------------------------------
FUSED_BLOCK = 128     # nodes per block of U (the M tile of the producer GEMM)
FUSED_TILE_COLS = 256  # columns of U per producer tile
IT_P, IT_C, IT_NOP = 0, 1, 2


@dataclass
class FusedSchedule:
    """Work-item list of a fused producer/consumer kernel (see gkn_b200.h).  Item v is executed by
    CTA ``v % grid``; every CTA runs its items in list order."""

    items: torch.Tensor   # int32 [num_items, 8]
    num_items: int
    num_flags: int
    u_rows: int           # rows of the U / dU buffer the items address
    ring: int
    grid: int = 0         # CTAs the list was dealt for (0: any)


def device_sm_count() -> int:
    import ctypes

    lib = _lib.load()
    n = ctypes.c_int(0)
    _lib.check(lib.gkn_device_info(ctypes.byref(n), None, None), "gkn_device_info")
    return int(n.value)


def fused_fwd_schedule(lay: NodeLayout, W: int, K: int, ring: int = 0, slack: float = 1.0,
                       grid: int | None = None) -> FusedSchedule:
    """Item list of gkn_node_fused_fwd_bf16 for one padded layout (cached on it).  Device-side
    index arithmetic only.

    Every item gets a time key in units of "one node block's worth of work": the producer tiles of
    block b are spread over [b-1, b), the message units of block b over [b+slack, b+1+slack) - a
    block of U is complete ``slack`` blocks before its first consumer asks for it and lives for
    about 2 + slack blocks (16 MB each at w = 64), inside the L2.  Tiles and units are dealt
    round-robin over the ``grid`` CTAs SEPARATELY (every CTA gets the same number of each kind) and
    each CTA's queue is sorted by key; queue slot t of CTA c is list position t*grid + c, unused
    slots are no-ops.  Waits only ever refer to items with a smaller key, and every queue is in
    key order, so the waits cannot form a cycle.  ``ring`` > 0: U is a ring of that many node
    blocks; a producer tile of block b then waits for the consumers of block b - ring."""
    cache = lay.__dict__.setdefault("_fused_sched", {})
    G = int(grid) if grid else device_sm_count()
    key = (W, K, ring, slack, G)
    if key in cache:
        return cache[key]
    if ring and ring < 2 + slack:
        raise ValueError("fused_fwd_schedule: the ring must hold at least 2 + slack node blocks "
                         "(a tile may only wait for consumers with a smaller time key)")
    dev = lay.units.device
    N, U = lay.num_nodes, lay.num_units
    NB = (N + FUSED_BLOCK - 1) // FUSED_BLOCK
    J = W * K // FUSED_TILE_COLS
    i64 = dict(dtype=torch.long, device=dev)
    # units per node block, from the per-node unit offsets (no host read)
    edges = torch.arange(0, NB + 1, **i64).mul_(FUSED_BLOCK).clamp_(max=N)
    ufirst_blk = lay.unit_first_dev[edges]           # [NB+1]
    T = ufirst_blk.diff()                             # C items per block
    # ---- producer tiles
    b = torch.arange(NB, **i64).repeat_interleave(J)
    j = torch.arange(J, **i64).repeat(NB)
    p_key = (b - 1).double() + (j.double() + 0.5) / J
    neg = torch.full_like(b, -1)
    if ring:
        slot = b % ring
        p_wait = torch.where(b >= ring, NB + b - ring, neg)
        p_target = torch.where(b >= ring, 4 * T[(b - ring).clamp_(min=0)], torch.zeros_like(b))
    else:
        slot, p_wait, p_target = b, neg, torch.zeros_like(b)
    p_items = torch.stack([torch.full_like(b, IT_P), b * FUSED_BLOCK, j * FUSED_TILE_COLS,
                           slot * FUSED_BLOCK, p_wait, p_target, b,
                           (N - b * FUSED_BLOCK).clamp_(max=FUSED_BLOCK)], 1)
    # ---- message units
    u = lay.units.long()
    grp, row0, rows = u[:, 0], u[:, 1], u[:, 2]
    blk = grp // FUSED_BLOCK
    t = torch.arange(U, **i64) - ufirst_blk[blk]
    c_key = blk.double() + slack + (t.double() + 0.5) / T[blk].clamp(min=1).double()
    urow = (blk % ring) * FUSED_BLOCK + grp % FUSED_BLOCK if ring else grp
    c_items = torch.stack([torch.full_like(grp, IT_C), row0, rows, urow, blk,
                           torch.full_like(grp, 4 * J), NB + blk if ring else torch.full_like(grp, -1),
                           torch.zeros_like(grp)], 1)
    # ---- deal both kinds round-robin, sort every CTA's queue by key, interleave the queues
    nP = NB * J
    cta = torch.cat([torch.arange(nP, **i64) % G, torch.arange(U, **i64) % G])
    keys = torch.cat([p_key, c_key])
    by_key = torch.argsort(keys, stable=True)
    order = by_key[torch.argsort(cta[by_key], stable=True)]     # sorted by (cta, key)
    cta_s = cta[order]
    cnt = torch.bincount(cta_s, minlength=G)
    first = torch.cumsum(cnt, 0) - cnt
    rank = torch.arange(nP + U, **i64) - first[cta_s]
    qlen = (nP + G - 1) // G + (U + G - 1) // G      # longest queue (host arithmetic, no sync)
    items = torch.zeros((qlen * G, 8), **i64)
    items[:, 0], items[:, 4], items[:, 6] = IT_NOP, -1, -1
    items[rank * G + cta_s] = torch.cat([p_items, c_items])[order]
    sched = FusedSchedule(items.int().contiguous(), qlen * G, 2 * NB,
                          ring * FUSED_BLOCK if ring else N, ring, G)
    cache[key] = sched
    return sched


_GEMM_SCHED_CACHE: dict = {}


def node_u_gemm(xb, w3r, u, N: int, K: int, W: int) -> None:
    """U = x @ W3r on the producer half of the fused kernel alone (gkn_node_fused_fwd_bf16 with
    "P" items only, no counters): the dense node-level GEMM of the node-first forward as a
    hand-written tcgen05 kernel - A = x tile, B = W3 read in place as an MN-major operand, bf16
    output through the swizzled staging tile.  xb bf16 [N, max(W,64)], w3r bf16 [max(W,64), W*K],
    u bf16 [N, W*K]."""
    lib = _lib.load()
    Wp = max(W, 64)
    if xb.shape != (N, Wp) or w3r.shape != (Wp, W * K) or u.numel() != N * W * K:
        raise ValueError("node_u_gemm: shape mismatch")
    G = device_sm_count()
    key = (N, W, K, G, str(xb.device))
    if key not in _GEMM_SCHED_CACHE:
        import numpy as np

        NB, J = (N + FUSED_BLOCK - 1) // FUSED_BLOCK, W * K // FUSED_TILE_COLS
        b = np.repeat(np.arange(NB, dtype=np.int64), J)
        j = np.tile(np.arange(J, dtype=np.int64), NB)
        it = np.zeros((NB * J, 8), dtype=np.int64)
        it[:, 0], it[:, 1], it[:, 2], it[:, 3] = IT_P, b * FUSED_BLOCK, j * FUSED_TILE_COLS, b * FUSED_BLOCK
        it[:, 4], it[:, 6] = -1, -1
        it[:, 7] = np.minimum(FUSED_BLOCK, N - b * FUSED_BLOCK)
        _GEMM_SCHED_CACHE[key] = torch.from_numpy(it.astype(np.int32)).to(xb.device)
    items = _GEMM_SCHED_CACHE[key]
    with torch.cuda.device(xb.device), _timed("node_u_gemm"):
        _lib.check(
            lib.gkn_node_fused_fwd_bf16(
                _p(xb, torch.bfloat16, "xb"), _p(w3r, torch.bfloat16, "w3r"), c_void_p(0),
                _p(u, torch.bfloat16, "u"), _p(items, torch.int32, "items"), int(items.size(0)),
                _p(items, torch.int32, "flags"), c_void_p(0), N, 0, N, K, W, G, _stream()),
            "gkn_node_fused_fwd_bf16",
        )


def node_fused_fwd(xb, w3r, hs, u, sched: FusedSchedule, flags, msg, N: int, e_pad: int, K: int,
                   W: int) -> None:
    """One NNConv application's U production + message GEMM in one launch
    (gkn_node_fused_fwd_bf16).  xb bf16 [N, max(W,64)]; w3r bf16 [max(W,64), W*K]; hs bf16
    [e_pad, K]; u bf16 [sched.u_rows, W*K] (written, and read back through the L2); flags int32
    [sched.num_flags], zero on entry; msg fp32 [e_pad, W]."""
    lib = _lib.load()
    Wp = max(W, 64)
    if xb.shape != (N, Wp) or w3r.shape != (Wp, W * K):
        raise ValueError(f"node_fused_fwd: xb must be [{N}, {Wp}] and w3r [{Wp}, {W * K}]")
    if u.numel() != sched.u_rows * W * K:
        raise ValueError("node_fused_fwd: the U buffer does not match the schedule")
    if flags.numel() < sched.num_flags:
        raise ValueError("node_fused_fwd: flags too small")
    with torch.cuda.device(xb.device), _timed("node_fused_fwd"):
        _lib.check(
            lib.gkn_node_fused_fwd_bf16(
                _p(xb, torch.bfloat16, "xb"), _p(w3r, torch.bfloat16, "w3r"),
                _p(hs, torch.bfloat16, "hs"), _p(u, torch.bfloat16, "u"),
                _p(sched.items, torch.int32, "items"), sched.num_items,
                _p(flags, torch.int32, "flags"), _p(msg, torch.float32, "msg"), N, e_pad,
                sched.u_rows, K, W, sched.grid, _stream()),
            "gkn_node_fused_fwd_bf16",
        )


# ------------------------------------------------------------------------------------------------
# Fused backward of one application (csrc/gkn_node_fused_bwd.cu)
# ------------------------------------------------------------------------------------------------
IT_D, IT_X = 0, 1
_BWD_SCHED_CACHE: dict = {}


def fused_bwd_grid(K: int, W: int) -> tuple[int, int]:
    """(grid size the X items are laid out for, number of dx partials) on the current device."""
    import ctypes

    lib = _lib.load()
    g, parts = ctypes.c_int32(0), ctypes.c_int32(0)
    _lib.check(lib.gkn_node_fused_bwd_grid(K, W, ctypes.byref(g), ctypes.byref(parts)),
               "gkn_node_fused_bwd_grid")
    return int(g.value), int(parts.value)


def fused_bwd_items(N: int, K: int, W: int, grid: int, ring: int, slack: float = 1.0,
                    groups: int = 1, spread: float = 0.9):
    """Host-side construction of the item list of gkn_node_fused_bwd_bf16 (numpy int32 [n, 8]).
    It depends only on (N, K, grid, ring, slack) - every node has K/128 "D" items whether or not it
    has edges - so it is built once and cached.

    Time keys as in the forward: the D items of block b are spread over [b-1, b); CTA c's X items
    of block b (one per 128-column pair of its slice) over [b+slack, b+slack+0.9), i.e. in between
    the CTA's D items of later blocks.  D items are dealt round-robin; every CTA's queue is sorted
    by key and queue slot t of CTA c is list position t*grid + c (unused slots are no-ops).  Inside
    a queue the D items alternate between the ``groups`` epilogue groups of the kernel (one, as
    built); the first D item a (CTA, group) has in a block carries the ring wait (block b-ring
    consumed), the last one the signal - a release covers all earlier stores of its warp.  Waits
    only refer to items with smaller keys: X(b) >= b+slack > D(b); D(b+ring) >= b+ring-1 >=
    b+slack+0.9 > X(b)."""
    import numpy as np

    if ring < 1 + spread + slack:
        raise ValueError("fused_bwd_items: the dU ring must hold at least 1 + spread + slack node blocks")
    C = K // 128
    NB = (N + FUSED_BLOCK - 1) // FUSED_BLOCK
    pairs = W * K // 128
    base, rem = divmod(pairs, grid)
    parts = min(grid, pairs)
    # ---- D items: global order (node, chunk), round-robin over the CTAs
    nD = N * C
    d = np.arange(nD, dtype=np.int64)
    node, chunk = d // C, d % C
    blk = node // FUSED_BLOCK
    t_blk = np.minimum(FUSED_BLOCK, N - blk * FUSED_BLOCK) * C
    idx_in_blk = d - blk * FUSED_BLOCK * C
    d_key = blk - 1 + (idx_in_blk + 0.5) / t_blk
    d_cta = d % grid
    D = np.zeros((nD, 8), dtype=np.int64)
    D[:, 0], D[:, 1], D[:, 2] = IT_D, node, chunk
    D[:, 3] = (blk % ring) * FUSED_BLOCK + node % FUSED_BLOCK
    D[:, 4], D[:, 6] = -1, -1
    # ---- X items: per CTA with a non-empty slice, per block, per pair
    npairs = base + (np.arange(grid) < rem)
    xc = np.repeat(np.arange(grid), npairs)                    # CTA of every (cta, q)
    xq = np.concatenate([np.arange(n) for n in npairs]) if pairs else np.zeros(0, dtype=np.int64)
    xn = npairs[xc]
    nXb = xc.size                                               # X items per block
    xb = np.repeat(np.arange(NB), nXb)
    xc, xq, xn = np.tile(xc, NB), np.tile(xq, NB), np.tile(xn, NB)
    x_key = xb + slack + spread * (xq + 0.5) / xn
    X = np.zeros((xb.size, 8), dtype=np.int64)
    X[:, 0], X[:, 1], X[:, 2] = IT_X, FUSED_BLOCK * xb, (xb % ring) * FUSED_BLOCK
    X[:, 3] = np.minimum(FUSED_BLOCK, N - FUSED_BLOCK * xb)
    first, last = xq == 0, xq == xn - 1
    X[:, 4] = np.where(first, xb, -1)
    X[:, 6] = np.where(last, NB + xb, -1)
    X[:, 7] = xq | (first.astype(np.int64) << 8) | (last.astype(np.int64) << 9)
    # ---- per-CTA queues in key order
    items = np.concatenate([D, X])
    cta = np.concatenate([d_cta, xc])
    key = np.concatenate([d_key, x_key])
    kind_blk = np.concatenate([blk, xb])
    order = np.lexsort((key, cta))                              # by cta, then key (stable)
    items, cta, kind_blk = items[order], cta[order], kind_blk[order]
    cnt = np.bincount(cta, minlength=grid)
    start = np.cumsum(cnt) - cnt
    rank = np.arange(items.shape[0]) - start[cta]
    # D items alternate between the epilogue groups in queue order
    is_d = items[:, 0] == IT_D
    d_rank = np.cumsum(is_d) - 1
    d_first_of_cta = np.zeros(grid, dtype=np.int64)
    d_cnt = np.bincount(cta[is_d], minlength=grid)
    d_first_of_cta[1:] = np.cumsum(d_cnt)[:-1]
    grp = (d_rank - d_first_of_cta[cta]) % groups
    # first / last D item of every (cta, group, block)
    di = np.nonzero(is_d)[0]
    gid = (cta[di] * groups + grp[di]) * NB + kind_blk[di]           # non-decreasing runs per (cta, grp)?
    o = np.argsort(gid, kind="stable")
    gs, dis = gid[o], di[o]
    run_first = np.ones(gs.size, dtype=bool)
    run_first[1:] = gs[1:] != gs[:-1]
    run_last = np.ones(gs.size, dtype=bool)
    run_last[:-1] = gs[1:] != gs[:-1]
    firsts, lasts = dis[run_first], dis[run_last]
    bl = kind_blk[lasts]
    items[lasts, 6] = bl
    signals = np.bincount(bl, minlength=NB)                     # releases per block
    bf = kind_blk[firsts]
    w = bf >= ring
    items[firsts[w], 4] = NB + bf[w] - ring
    items[firsts[w], 5] = 4 * parts
    # X items wait for every release of their block
    xi = np.nonzero((items[:, 0] == IT_X) & (items[:, 4] >= 0))[0]
    items[xi, 5] = 4 * signals[items[xi, 4]]
    qlen = int(cnt.max()) if cnt.size else 0
    out = np.zeros((qlen * grid, 8), dtype=np.int64)
    out[:, 0], out[:, 4], out[:, 6] = IT_NOP, -1, -1
    out[rank * grid + cta] = items
    return out.astype(np.int32), 2 * NB


def fused_bwd_schedule(N: int, K: int, W: int, device, ring: int = 3,
                       slack: float = 1.0, spread: float = 0.9) -> FusedSchedule:
    grid, _ = fused_bwd_grid(K, W)
    key = (N, K, W, grid, ring, slack, spread, str(device))
    if key not in _BWD_SCHED_CACHE:
        items, nflags = fused_bwd_items(N, K, W, grid, ring, slack, spread=spread)
        t = torch.from_numpy(items).to(device)
        _BWD_SCHED_CACHE[key] = FusedSchedule(t, int(t.size(0)), nflags, ring * FUSED_BLOCK, ring,
                                              grid)
    return _BWD_SCHED_CACHE[key]


def node_fused_bwd(hs, gs, lay: NodeLayout, xb, w3k, du, sched: FusedSchedule, flags, dx_part, dwT,
                   accumulate_dw: bool) -> None:
    """dU produced and consumed in one launch: dwT (+)= dU^T-by-x products, dx partials per CTA
    (gkn_node_fused_bwd_bf16).  du: the zero-initialised ring [sched.u_rows, W*K] bf16."""
    lib = _lib.load()
    K, W = int(hs.size(1)), int(gs.size(1))
    N = lay.num_nodes
    if du.numel() != sched.u_rows * W * K:
        raise ValueError("node_fused_bwd: the dU ring does not match the schedule")
    with torch.cuda.device(hs.device), _timed("node_fused_bwd"):
        _lib.check(
            lib.gkn_node_fused_bwd_bf16(
                _p(hs, torch.bfloat16, "hs"), _p(gs, torch.bfloat16, "gs"),
                _p(lay.pptr, torch.int32, "pptr"), _p(xb, torch.bfloat16, "xb"),
                _p(w3k, torch.bfloat16, "w3k"), _p(du, torch.bfloat16, "du"),
                _p(sched.items, torch.int32, "items"), sched.num_items,
                _p(flags, torch.int32, "flags"), _p(dx_part, torch.float32, "dx_part"),
                _p(dwT, torch.float32, "dwT"), int(bool(accumulate_dw)), N, lay.e_pad, sched.u_rows,
                K, W, _stream()),
            "gkn_node_fused_bwd_bf16",
        )


def sum_partials(part: torch.Tensor, addend: torch.Tensor | None, out: torch.Tensor) -> torch.Tensor:
    """out = addend + part.sum(0) with a fixed summation order (gkn_sum_partials)."""
    lib = _lib.load()
    parts = int(part.size(0))
    n = part.numel() // max(parts, 1)
    with torch.cuda.device(part.device):
        _lib.check(lib.gkn_sum_partials(_p(part, torch.float32, "part"), parts, n,
                                        _p(addend, torch.float32, "addend"),
                                        _p(out, torch.float32, "out"), _stream()),
                   "gkn_sum_partials")
    return out


# ------------------------------------------------------------------------------------------------
# Sharded rollout: exchange + node update + broadcast in one kernel over peer memory
# ------------------------------------------------------------------------------------------------
def shard_reduce_update(partial_ptrs, xout_ptrs, x, inv_cnt, B3, root, bias, n0: int, n1: int,
                        relu: bool, mc_partial: int = 0, mc_xout: int = 0) -> None:
    """For the nodes [n0, n1) this rank owns: sum the partial rows of every rank (``partial_ptrs``:
    device addresses of the ranks' [2, N, W] fp32 buffers, in rank order), apply the node update
    and store the new rows into every rank's x' buffer (``xout_ptrs``) - gkn_shard_reduce_update.
    ``mc_partial`` / ``mc_xout``: multicast addresses of the two symmetric buffers (NVLS) - the
    NVSwitch then does the sum and replicates the stores.  The caller synchronises the ranks
    before and after."""
    import ctypes

    lib = _lib.load()
    world = len(partial_ptrs)
    if len(xout_ptrs) != world:
        raise ValueError("shard_reduce_update: one partial buffer and one x' buffer per rank")
    N, W = int(x.size(0)), int(x.size(1))
    arr = ctypes.c_uint64 * world
    pa, xa = arr(*[int(v) for v in partial_ptrs]), arr(*[int(v) for v in xout_ptrs])
    f32 = torch.float32
    with torch.cuda.device(x.device):
        _lib.check(
            lib.gkn_shard_reduce_update(
                ctypes.cast(pa, c_void_p), ctypes.cast(xa, c_void_p), int(mc_partial), int(mc_xout),
                world, _p(x, f32, "x"),
                _p(inv_cnt, f32, "inv_cnt"), _p(B3, f32, "B3"), _p(root, f32, "root"),
                _p(bias, f32, "bias"), int(n0), int(n1), N, W, int(bool(relu)), _stream()),
            "gkn_shard_reduce_update",
        )
