"""CPU: the node-first factorisation's host logic (padded source-sorted layout, unit table, the
autograd algebra of ``functional._GKNStackNode``) checked against the oracle.

The CUDA entry points are replaced here by short torch restatements of what each kernel is
documented to compute in include/gkn_b200.h (same arguments, same layouts), so everything
ABOVE the C ABI runs unchanged on the CPU; the kernels themselves are checked on the GPU
(tests/test_node_path_gpu.py).  The emulation is test infrastructure only.
"""
import numpy as np
import pytest
import torch

from molecular_dynamics_neural_operator_b200 import functional as GF
from molecular_dynamics_neural_operator_b200 import ops
from oracle import gkn_oracle as orc


def cpu_plan(edge_index: torch.Tensor, N: int) -> ops.GraphPlan:
    """What gkn_graph_prepare produces (destination-sorted positions, ties in edge order)."""
    src, dst = edge_index[0], edge_index[1]
    E = src.numel()
    perm = torch.sort(dst * E + torch.arange(E), stable=True).indices
    src_s, dst_s = src[perm], dst[perm]
    cnt = torch.bincount(dst_s, minlength=N)
    dst_ptr = torch.cat([torch.zeros(1, dtype=torch.long), cnt.cumsum(0)])
    src_perm = torch.sort(src_s * E + torch.arange(E), stable=True).indices
    scnt = torch.bincount(src_s, minlength=N)
    src_ptr = torch.cat([torch.zeros(1, dtype=torch.long), scnt.cumsum(0)])
    inv = 1.0 / cnt.clamp(min=1).float()
    return ops.GraphPlan(N, E, perm.int(), src_s.int(), dst_s.int(), dst_ptr.int(), inv,
                         src_ptr.int(), src_perm.int())


def random_graph(N, E, seed, isolated=()):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, N, (2, E), generator=g)
    for n in isolated:  # nodes with no outgoing and no incoming edges
        ei[ei == n] = (n + 1) % N
    return ei


# ---- torch restatements of the C entry points ----------------------------------------------
def emu_edge_layer0(ea, w1, b1, out_dtype):
    return torch.relu(ea @ w1.t() + b1).to(out_dtype)


def emu_segment_matvec(rows, perm, seg_ptr, gidx, gnode, scale, M1, nin, M2, bias, relu,
                       want_agg=False):
    N, W = (nin if nin is not None else gnode).shape
    r = torch.zeros(N, W)
    agg = torch.zeros(N, W)
    for n in range(N):
        q = torch.arange(int(seg_ptr[n]), int(seg_ptr[n + 1]))
        p = perm[q].long() if perm is not None else q
        if rows is not None:
            r[n] = rows[p].sum(0)
        agg[n] = gnode[gidx[p].long()].sum(0)
    if scale is not None:
        r, agg = r * scale[:, None], agg * scale[:, None]
    out = r if M1 is None else r + agg @ M1 + nin @ M2  # M1 = M2 = None: plain segment sums
    if bias is not None:
        out = out + bias
    if relu:
        out = torch.relu(out)
    return out, (agg if want_agg else None)


def emu_relu_grad_scale(gy, y, inv_cnt, relu):
    g = gy * (y > 0) if relu else gy.clone()
    return g, g * inv_cnt[:, None]


def emu_mask_cast_colsum(dz, act, colsum, out_dtype, side=None, accumulate=True, sums_only=False):
    out = (dz.float() * (act.float() > 0)).to(out_dtype)
    o = out.float()
    cs = colsum.view(-1, dz.size(1))
    cs[0] += o.sum(0)
    if side is not None:
        cs[1:] += side.t() @ o
    return None if sums_only else out


def emu_node_msg(hs, u, lay, u0, u1, node_base, msg):
    for g, row0, rows, _ in lay.units[u0:u1].tolist():
        msg[row0:row0 + rows] = hs[row0:row0 + rows].float() @ u[g - node_base].float().t()


def emu_node_du(hsT, gsT, lay, n0, n1, dut):
    K, W = hsT.size(0), gsT.size(0)
    pp = lay.pptr.tolist()
    d = dut.view(n1 - n0, K, W)
    for n in range(n0, n1):
        sl = slice(pp[n], pp[n + 1])
        d[n - n0] = (hsT[:, sl].float() @ gsT[:, sl].float().t()).to(dut.dtype)


def emu_node_dh(gs, ut, lay, u0, u1, node_base, dh):
    A = gs.size(0)
    for g, row0, rows, _ in lay.units[u0:u1].tolist():
        acc = torch.zeros(rows, dh.size(1))
        for a in range(A):
            acc += gs[a, row0:row0 + rows].float() @ ut[a, g - node_base].float().t()
        dh[row0:row0 + rows] = acc.to(dh.dtype)


def emu_node_du_mn(hs, gs, lay, n0, n1, dut):
    W = dut.size(1) // hs.size(1)  # gs rows are ops.gs_width(W) long, columns >= W are zeros
    assert gs.size(1) == ops.gs_width(W)
    emu_node_du(hs.t(), gs[:, :W].t(), lay, n0, n1, dut)


def emu_relu_bits(act):
    pos = (act > 0).view(act.size(0), -1, 32).long()
    words = (pos << torch.arange(32)).sum(-1)
    return torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)


def emu_node_dh_mn(gs, u, lay, u0, u1, node_base, dh, act_bits=None, colsum=None, accumulate=False):
    W = u.size(2)
    assert gs.size(2) == ops.gs_width(W)
    emu_node_dh(gs[:, :, :W], u.transpose(2, 3), lay, u0, u1, node_base, dh)
    if act_bits is not None:  # ReLU backward of layers.2 + bias gradient over this call's rows
        keep = ((act_bits.long()[:, :, None] >> torch.arange(32)) & 1).bool().view(dh.shape)
        tot = torch.zeros(dh.size(1))
        for _, row0, rows, _ in lay.units[u0:u1].tolist():
            sl = slice(row0, row0 + rows)
            dh[sl] = torch.where(keep[sl], dh[sl], torch.zeros((), dtype=dh.dtype))
            tot += dh[sl].float().sum(0)
        colsum.copy_(colsum + tot if accumulate else tot)


def emu_node_gather(g, lay, gs, gsT):
    idx = lay.dst_pad.long()
    rows = torch.where((idx >= 0)[:, None], g[idx.clamp(min=0)], torch.zeros(1)).to(torch.bfloat16)
    if gs is not None:
        gs.zero_()
        gs[:, :g.size(1)].copy_(rows)
    if gsT is not None:
        gsT.copy_(rows.t())


def emu_node_u_gemm(xb, w3r, u, N, K, W):
    u.view(N, W * K).copy_((xb.float() @ w3r.float()).to(u.dtype))


def emu_node_fused_fwd(xb, w3r, hs, u, sched, flags, msg, N, e_pad, K, W, seed=0):
    """Executes the item list the way the kernel does: ``sched.grid`` CTAs, item v belongs to CTA
    v % grid, every CTA runs its items strictly in order, an item whose counter has not reached
    its target blocks its CTA.  CTAs are advanced in a random order; a state in which no CTA can
    advance is a deadlock (AssertionError).  A producer tile additionally refuses to overwrite a
    U row that a not-yet-executed consumer still needs only through the counters - exactly the
    kernel's protection - so a wrong schedule shows up as wrong numbers."""
    items = sched.items.tolist()
    num_ctas = sched.grid
    queues = [items[c::num_ctas] for c in range(num_ctas)]
    pos = [0] * num_ctas
    rng = np.random.default_rng(seed)
    uf = u.view(sched.u_rows, W * K)
    left = len(items)
    while left:
        progressed = False
        for c in rng.permutation(num_ctas):
            while pos[c] < len(queues[c]):
                kind, r0, r1, r2, widx, wtarget, sidx, r3 = queues[c][pos[c]]
                if widx >= 0 and int(flags[widx]) < wtarget:
                    break
                if kind == ops.IT_P:
                    tile = xb[r0:r0 + r3].float() @ w3r[:, r1:r1 + ops.FUSED_TILE_COLS].float()
                    uf[r2:r2 + r3, r1:r1 + ops.FUSED_TILE_COLS] = tile.to(u.dtype)
                elif kind == ops.IT_C:
                    msg[r0:r0 + r1] = hs[r0:r0 + r1].float() @ uf[r2].view(W, K).float().t()
                if sidx >= 0 and kind != ops.IT_NOP:
                    flags[sidx] += 4  # one per epilogue warp
                pos[c] += 1
                left -= 1
                progressed = True
                if rng.random() < 0.3:
                    break  # let another CTA run: interleavings differ from run to run
        assert progressed, "fused schedule deadlocked"


EMU_GRID = 7  # CTAs of the emulated fused backward


def emu_node_fused_bwd(hs, gs, lay, xb, w3k, du, sched, flags, dx_part, dwT, accumulate_dw, seed=0):
    """In-order multi-CTA execution of the backward item list (see emu_node_fused_fwd): D items
    write dU tiles into the ring (alternating between the CTA's two epilogue groups; a signal
    releases every earlier store of its group, which is why only the last D item a group has in a
    block carries one), X items consume one 128-column pair of the CTA's slice of a block."""
    K, W = hs.size(1), gs.size(1)
    grid = sched.grid
    items = sched.items.tolist()
    queues = [items[c::grid] for c in range(grid)]
    pos = [0] * grid
    pairs = W * K // 128
    base, rem = divmod(pairs, grid)
    p0 = [c * base + min(c, rem) for c in range(grid)]
    npair = [base + (c < rem) for c in range(grid)]
    pp = lay.pptr.tolist()
    duf = du.view(sched.u_rows, K * W)
    acc = [torch.zeros(npair[c] * 128, W) for c in range(grid)]
    dxa = [None] * grid
    # stores of a group become visible only when that group signals (models the deferred release)
    staged = [[[], []] for _ in range(grid)]
    dcount = [0] * grid
    rng = np.random.default_rng(seed)
    left = len(items)
    while left:
        progressed = False
        for c in rng.permutation(grid):
            while pos[c] < len(queues[c]):
                kind, r0, r1, r2, widx, wtarget, sidx, r3 = queues[c][pos[c]]
                if widx >= 0 and int(flags[widx]) < wtarget:
                    break
                if kind == ops.IT_D:
                    g = 0  # one epilogue group
                    dcount[c] += 1
                    rows = slice(pp[r0], pp[r0 + 1])
                    tile = hs[rows, r1 * 128:(r1 + 1) * 128].float().t() @ gs[rows].float()
                    staged[c][g].append((r2, r1, tile.to(du.dtype)))
                    if sidx >= 0:
                        for row, ch, t in staged[c][g]:
                            duf[row].view(K, W)[ch * 128:(ch + 1) * 128] = t
                        staged[c][g] = []
                        flags[sidx] += 4
                elif kind == ops.IT_X:
                    q, first, last = r3 & 0xff, (r3 >> 8) & 1, (r3 >> 9) & 1
                    cols = slice((p0[c] + q) * 128, (p0[c] + q + 1) * 128)
                    blk = duf[r1:r1 + r2][:, cols].float()
                    acc[c][q * 128:(q + 1) * 128] += blk.t() @ xb[r0:r0 + r2].float()
                    part = blk @ w3k[:, cols].float().t()
                    dxa[c] = part if first else dxa[c] + part
                    if last:
                        if dx_part is not None:
                            dx_part[c, r0:r0 + r2] = dxa[c]
                        flags[sidx] += 4
                pos[c] += 1
                left -= 1
                progressed = True
                if rng.random() < 0.3:
                    break
        assert progressed, "fused backward schedule deadlocked"
    assert all(not g for s in staged for g in s), "a D item was never released"
    if dwT is not None:
        for c in range(grid):
            sl = slice(p0[c] * 128, (p0[c] + npair[c]) * 128)
            if accumulate_dw:
                dwT[sl] += acc[c]
            else:
                dwT[sl] = acc[c]


def emu_xty(x, g, colsum=False):
    return x.t() @ g, (g.sum(0) if colsum else None)


def emu_sum_partials(part, addend, out):
    tot = part.sum(0).reshape(out.shape)
    out.copy_(tot if addend is None else addend + tot)
    return out


def install_emulation(setattr_fn):
    """Replace every CUDA entry point the node-first host code calls by its torch restatement
    (``setattr_fn(obj, name, value)``: monkeypatch.setattr, or plain setattr in a spawned worker)."""
    for name, fn in [("edge_layer0", emu_edge_layer0), ("segment_matvec", emu_segment_matvec),
                     ("relu_grad_scale", emu_relu_grad_scale),
                     ("mask_cast_colsum", emu_mask_cast_colsum), ("node_msg", emu_node_msg),
                     ("node_du", emu_node_du), ("node_dh", emu_node_dh),
                     ("node_du_mn", emu_node_du_mn), ("node_dh_mn", emu_node_dh_mn),
                     ("relu_bits", emu_relu_bits), ("xty", emu_xty),
                     ("node_gather", emu_node_gather), ("node_fused_fwd", emu_node_fused_fwd),
                     ("node_u_gemm", emu_node_u_gemm),
                     ("node_fused_bwd", emu_node_fused_bwd), ("sum_partials", emu_sum_partials),
                     ("fused_bwd_grid", lambda K, W: (EMU_GRID, EMU_GRID)),
                     ("device_sm_count", lambda: EMU_GRID)]:
        setattr_fn(ops, name, fn)
    setattr_fn(ops, "prepare_graph", lambda ei, N, validate=True: cpu_plan(ei, N))
    setattr_fn(GF, "_mm32", lambda a, b: a.float() @ b.float())
    setattr_fn(GF, "_addmm32", lambda acc, a, b: acc + a.float() @ b.float())
    setattr_fn(torch, "_addmm_activation",
               lambda b, a, w, **kw: torch.relu(a.float() @ w.float() + b.float()).to(a.dtype))


@pytest.fixture
def emulated(monkeypatch):
    install_emulation(monkeypatch.setattr)


# ---- layout ----------------------------------------------------------------------------------
@pytest.mark.parametrize("N,E,isolated", [(7, 40, ()), (5, 700, ()), (9, 60, (3, 8)), (4, 1, ())])
def test_padded_source_sorted_layout_invariants(N, E, isolated):
    ei = random_graph(N, E, seed=N + E, isolated=isolated)
    plan = cpu_plan(ei, N)
    lay = ops.node_layout(plan)
    pp = lay.pptr.long()
    deg = torch.bincount(ei[0], minlength=N)
    assert lay.e_pad % 16 == 0 and int(pp[-1]) == lay.e_pad
    assert torch.equal((pp[1:] - pp[:-1]), (deg + 15) // 16 * 16)
    # every edge sits in the group of its source, at a distinct row; pad rows are marked -1
    q = lay.qpos.long()
    assert q.unique().numel() == E
    assert torch.equal(lay.src_pad.long()[q], plan.src.long())
    assert torch.equal(lay.dst_pad.long()[q], plan.dst.long())
    assert int((lay.dst_pad >= 0).sum()) == E
    grp_rows = q - pp[plan.src.long()]
    assert bool((grp_rows >= 0).all()) and bool((grp_rows < deg[plan.src.long()]).all())
    # the unit table tiles every group exactly once, in node order
    covered = torch.zeros(lay.e_pad, dtype=torch.int32)
    for g, row0, rows, _ in lay.units.tolist():
        assert 0 < rows <= 128 and pp[g] <= row0 and row0 + rows <= pp[g + 1]
        covered[row0:row0 + rows] += 1
    assert bool((covered == 1).all())
    assert lay.unit_first[0] == 0 and lay.unit_first[-1] == lay.num_units
    assert torch.equal(lay.units[:, 0].long(), lay.units[:, 0].long().sort().values)
    assert ops.node_layout(plan) is lay  # cached on the plan
    # unit table of the dh pass: every (tile, chunk) once, node-major, chunk-major inside a node
    d = lay.dh_units(4)
    assert d.shape == (lay.num_units * 4, 4)
    assert len(set(map(tuple, d.tolist()))) == lay.num_units * 4
    for n in range(N):
        seg = d[lay.unit_first[n] * 4:lay.unit_first[n + 1] * 4]
        assert bool((seg[:, 0] == n).all())
        assert seg[:, 3].tolist() == sorted(seg[:, 3].tolist())


# ---- autograd algebra -------------------------------------------------------------------------
def _stack_case(N, E, K, depth, seed, isolated=(), W=64):
    ei = random_graph(N, E, seed, isolated)
    g = torch.Generator().manual_seed(seed + 1)
    net = orc.make_net_params(6, K, W, seed=seed, dtype=torch.float32)
    p = {
        "x": torch.randn(N, W, generator=g), "ea": torch.randn(E, 6, generator=g),
        "root1": torch.randn(W, W, generator=g) / 8, "bias1": torch.randn(W, generator=g) / 8,
        "root2": torch.randn(W, W, generator=g) / 8, "bias2": torch.randn(W, generator=g) / 8,
    }
    for k, v in net.items():
        p["net." + k] = v
    return ei, p, depth


def _run_node_stack(ei, p, depth, N, relu=True, ea_grad=True):
    plan = cpu_plan(ei, N)
    t = {k: v.clone().requires_grad_(ea_grad or k != "ea") for k, v in p.items()}
    ea_sorted = t["ea"].index_select(0, plan.perm.long())
    out = GF._GKNStackNode.apply(
        t["x"], ea_sorted, t["net.layers.0.weight"], t["net.layers.0.bias"],
        t["net.layers.2.weight"], t["net.layers.2.bias"], t["net.layers.4.weight"],
        t["net.layers.4.bias"], t["root1"], t["bias1"], t["root2"], t["bias2"], plan,
        (0,) * depth + (1,) * depth, relu, "bf16")
    return out, t


def _run_oracle_stack(ei, p, depth, relu=True):
    t = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    net = {k[4:]: v for k, v in t.items() if k.startswith("net.")}
    if relu:
        out = orc.conv_stack(t["x"], ei, t["ea"], net, t["root1"], t["bias1"], t["root2"],
                             t["bias2"], depth)
    else:  # the same stack without the activation (smooth in the operands: tight tolerances)
        out = t["x"]
        for c in (1,) * depth + (2,) * depth:
            out = orc.nnconv_forward(out, ei, t["ea"], net, t[f"root{c}"], t[f"bias{c}"])
    return out, t


def fro(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


@pytest.mark.parametrize("mn,save_u", [(True, True), (True, False), (False, False)],
                         ids=["mn_saved_u", "mn_recomputed_u", "kmajor"])
@pytest.mark.parametrize("N,E,K,depth,isolated,chunk_mb", [
    (6, 90, 256, 1, (), 1024),      # one application, single node chunk
    (5, 700, 256, 2, (), 1024),     # groups longer than one 128-row tile
    (9, 120, 256, 2, (2, 7), 0),    # isolated nodes, node-range chunking of U / dU / U^T (1 node)
])
@pytest.mark.parametrize("fwd_mode,fused_bwd", [("gemm", False), ("1", True), ("0", False)],
                         ids=["own_gemm", "fused_fwd_bwd", "library_gemm"])
def test_node_first_stack_matches_oracle(emulated, monkeypatch, N, E, K, depth, isolated, chunk_mb,
                                         mn, save_u, fwd_mode, fused_bwd):
    """Without the activation between applications the stack is smooth in its operands, so bf16
    rounding is the only difference from the fp64 oracle: every gradient within 2^-5 except those
    upstream of the edge MLP's own ReLUs (mask flips, see tests/test_layer_gpu.py)."""
    monkeypatch.setattr(GF, "_NODE_CHUNK_BYTES", chunk_mb << 20)
    monkeypatch.setattr(GF, "_NODE_DH_CHUNK_BYTES", chunk_mb << 20)
    monkeypatch.setattr(GF, "_MLP_BWD_CHUNK_ROWS", 64)
    monkeypatch.setattr(GF, "NODE_MN", mn)
    monkeypatch.setattr(GF, "_SAVE_U_BYTES", (1 << 40) if save_u else 0)
    monkeypatch.setattr(GF, "NODE_FUSED", fwd_mode == "1")
    monkeypatch.setattr(GF, "NODE_U_GEMM", fwd_mode == "gemm")
    monkeypatch.setattr(GF, "NODE_FUSED_BWD", fused_bwd)
    monkeypatch.setattr(GF, "_FUSED_BWD_MIN_DEGREE", 0)  # take it whatever the degree
    ei, p, depth = _stack_case(N, E, K, depth, seed=11 * N + depth, isolated=isolated)
    out, t = _run_node_stack(ei, p, depth, N, relu=False)
    ref, tr = _run_oracle_stack(ei, p, depth, relu=False)
    assert fro(out, ref) < 2.0 ** -5
    gen = torch.Generator().manual_seed(5)
    gout = torch.randn(out.shape, generator=gen)
    out.backward(gout)
    ref.backward(gout.double())
    for k in p:
        tol = 2.0 ** -2 if k.startswith(("net.layers.0", "net.layers.2", "ea")) else 2.0 ** -5
        assert t[k].grad is not None, k
        assert fro(t[k].grad, tr[k].grad) < tol, k


@pytest.mark.parametrize("W", [32, 128])
@pytest.mark.parametrize("save_u", [True, False], ids=["saved_u", "recomputed_u"])
def test_node_first_stack_at_other_widths_matches_oracle(emulated, monkeypatch, W, save_u):
    """Training on the node-first path at w = 32 (config 1) and w = 128: the padded gs buffer
    (ops.gs_width) and the per-width dU / dh kernels' host contract, against the fp64 oracle."""
    monkeypatch.setattr(GF, "_MLP_BWD_CHUNK_ROWS", 64)
    monkeypatch.setattr(GF, "_SAVE_U_BYTES", (1 << 40) if save_u else 0)
    monkeypatch.setattr(GF, "_NODE_CHUNK_BYTES", W * 256 * 2 * 3)  # node ranges of 3 nodes
    N, E, K, depth = 7, 260, 256, 2
    assert GF.node_path_supported(W, K, E, "bf16")
    ei, p, depth = _stack_case(N, E, K, depth, seed=W, W=W)
    out, t = _run_node_stack(ei, p, depth, N, relu=False)
    ref, tr = _run_oracle_stack(ei, p, depth, relu=False)
    assert fro(out, ref) < 2.0 ** -5
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    out.backward(gout)
    ref.backward(gout.double())
    for k in p:
        tol = 2.0 ** -2 if k.startswith(("net.layers.0", "net.layers.2", "ea")) else 2.0 ** -5
        assert fro(t[k].grad, tr[k].grad) < tol, k


@pytest.mark.parametrize("save_u", [True, False], ids=["saved_u", "recomputed_u"])
def test_relu_backward_fused_into_the_dh_pass_changes_nothing(emulated, monkeypatch, save_u):
    """GKN_DH_FUSED_MASK: dz2 and the layers.2 bias gradient from the dh epilogue (node ranges
    chained through ``accumulate``) against the separate mask pass."""
    monkeypatch.setattr(GF, "_SAVE_U_BYTES", (1 << 40) if save_u else 0)
    monkeypatch.setattr(GF, "_NODE_DH_CHUNK_BYTES", 2 * 2 * 64 * 256 * 2 * 3)  # 3 nodes per range
    N, E, K, depth = 8, 300, 256, 2
    ei, p, depth = _stack_case(N, E, K, depth, seed=21)
    grads = []
    for fused in (True, False):
        monkeypatch.setattr(GF, "DH_FUSED_MASK", fused)
        out, t = _run_node_stack(ei, p, depth, N, relu=True)
        out.backward(torch.ones_like(out))
        grads.append(t)
    for k in p:
        assert fro(grads[0][k].grad, grads[1][k].grad) < 1e-6, k


def test_node_first_stack_with_relu_matches_oracle(emulated):
    """With F.relu after every application (graph_kernel.py:299-302): outputs within the stack
    bound; gradients are compared where the bf16 and fp64 activation masks agree."""
    N, E, K, depth = 8, 300, 256, 2
    ei, p, depth = _stack_case(N, E, K, depth, seed=3)
    out, t = _run_node_stack(ei, p, depth, N, relu=True)
    ref, tr = _run_oracle_stack(ei, p, depth, relu=True)
    assert fro(out, ref) < 2.0 ** -5
    gen = torch.Generator().manual_seed(5)
    gout = torch.randn(out.shape, generator=gen)
    out.backward(gout)
    ref.backward(gout.double())
    for k in ("x", "root1", "bias2", "net.layers.4.weight", "net.layers.4.bias"):
        a, b = t[k].grad.double().flatten(), tr[k].grad.flatten()
        assert float(a.dot(b) / (a.norm() * b.norm())) > 0.98, k


def test_layer0_gradients_agree_with_and_without_edge_attr_gradient(emulated):
    """Without an edge_attr gradient, dW1 comes from the hi/lo-split bf16 GEMM and db1 from the
    in-place mask pass; with it, from the fp32 side sums.  Same numbers up to bf16 rounding of dz1."""
    N, E, K, depth = 7, 200, 256, 1
    ei, p, depth = _stack_case(N, E, K, depth, seed=4)
    grads = []
    for ea_grad in (True, False):
        out, t = _run_node_stack(ei, p, depth, N, relu=False, ea_grad=ea_grad)
        out.backward(torch.ones_like(out))
        grads.append(t)
    for k in ("net.layers.0.weight", "net.layers.0.bias", "net.layers.2.weight"):
        assert fro(grads[1][k].grad, grads[0][k].grad) < 2.0 ** -7, k
    assert grads[1]["ea"].grad is None


def test_only_the_node_features_need_a_gradient(emulated):
    """Frozen edge MLP / roots (e.g. a rollout differentiated w.r.t. its input): no dh pass, no
    message-gradient stack, dx still through dU."""
    N, E, K, depth = 6, 150, 256, 1
    ei, p, depth = _stack_case(N, E, K, depth, seed=8)
    plan = cpu_plan(ei, N)
    t = {k: v.clone() for k, v in p.items()}
    t["x"].requires_grad_(True)
    out = GF._GKNStackNode.apply(
        t["x"], t["ea"].index_select(0, plan.perm.long()), t["net.layers.0.weight"],
        t["net.layers.0.bias"], t["net.layers.2.weight"], t["net.layers.2.bias"],
        t["net.layers.4.weight"], t["net.layers.4.bias"], t["root1"], t["bias1"], t["root2"],
        t["bias2"], plan, (0, 1), False, "bf16")
    ref, tr = _run_oracle_stack(ei, p, depth, relu=False)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    out.backward(gout)
    ref.backward(gout.double())
    assert fro(t["x"].grad, tr["x"].grad) < 2.0 ** -5


def test_node_path_selection():
    assert GF.node_path_supported(64, 1024, 10, "bf16")
    assert GF.node_path_supported(32, 1024, 10, "bf16")
    # w = 32 training chooses by degree: config 1 (14 edges per node) stays edge-first, denser
    # graphs and forward-only calls go node-first
    assert not GF.node_path_supported(32, 1024, 50776, "bf16", num_nodes=3584)
    assert GF.node_path_supported(32, 1024, 50776, "bf16", needs_grad=False, num_nodes=3584)
    assert GF.node_path_supported(32, 1024, 60000, "bf16", num_nodes=504)
    assert GF.node_path_supported(128, 1024, 50776, "bf16", num_nodes=3584)
    assert GF.node_path_supported(128, 1024, 10, "bf16")
    assert GF.node_path_supported(32, 1024, 10, "bf16", needs_grad=False)
    assert GF.node_path_supported(128, 1024, 10, "bf16", needs_grad=False)
    assert not GF.node_path_supported(16, 1024, 10, "bf16")
    assert not GF.node_path_supported(16, 1024, 10, "bf16", needs_grad=False)
    assert not GF.node_path_supported(64, 1024, 10, "fp32")
    assert not GF.node_path_supported(64, 1000, 10, "bf16")
    assert not GF.node_path_supported(64, 1024, 0, "bf16")


# ---- fused producer/consumer forward: schedule ------------------------------------------------
@pytest.mark.parametrize("N,E,W,ring,grid", [(300, 5000, 64, 0, 7), (300, 5000, 64, 3, 148),
                                             (700, 3000, 32, 4, 5), (130, 400, 128, 3, 16),
                                             (40, 90, 64, 0, 148)])
def test_fused_forward_schedule_is_balanced_ordered_and_complete(N, E, W, ring, grid):
    K = 256
    plan = cpu_plan(random_graph(N, E, seed=N), N)
    lay = ops.node_layout(plan)
    sched = ops.fused_fwd_schedule(lay, W, K, ring=ring, grid=grid)
    it = sched.items.long()
    NB, J = -(-N // 128), W * K // 256
    assert sched.grid == grid and sched.num_items % grid == 0 and sched.num_flags == 2 * NB
    P, C = it[it[:, 0] == ops.IT_P], it[it[:, 0] == ops.IT_C]
    # every tile of U once, every message unit once
    assert len(P) == NB * J and len({(int(a), int(b)) for a, b in P[:, 1:3]}) == NB * J
    assert sorted(C[:, 1].tolist()) == sorted(lay.units[:, 1].tolist())
    # both kinds are spread evenly: every CTA gets the same number of each, +-1
    cta = torch.arange(sched.num_items) % grid
    for kind in (ops.IT_P, ops.IT_C):
        per = torch.bincount(cta[it[:, 0] == kind], minlength=grid)
        assert int(per.max() - per.min()) <= 1
    # time keys: (block of the item, phase) reconstructed from the fields; every CTA's queue is in
    # key order and a wait only refers to a block whose items all have smaller keys
    cnt = torch.bincount(C[:, 4], minlength=NB)
    for c in range(grid):
        q = it[c::grid]
        last = -10.0
        for row in q.tolist():
            if row[0] == ops.IT_P:
                key = row[1] // 128 - 1 + (row[2] // 256 + 0.5) / J
            elif row[0] == ops.IT_C:
                key = row[4] + 1.0  # lower bound of [b + slack, b + 1 + slack), slack = 1
            else:
                continue
            assert key >= last - 1.0 + 1e-9 if row[0] == ops.IT_C else key >= last - 1e-9
            last = max(last, key)
    for row in C.tolist():
        assert row[5] == 4 * J and 0 <= row[4] < NB
    for row in P.tolist():
        b = row[1] // 128
        if ring and b >= ring:
            assert row[4] == NB + b - ring and row[5] == 4 * int(cnt[b - ring])
            assert row[3] == (b % ring) * 128
        else:
            assert row[4] == -1
    assert ops.fused_fwd_schedule(lay, W, K, ring=ring, grid=grid) is sched  # cached on the layout


@pytest.mark.parametrize("ring", [0, 3])
@pytest.mark.parametrize("num_ctas", [1, 3, 148])
def test_fused_forward_emulated_execution_matches_two_kernel_form(ring, num_ctas):
    N, E, W, K = 400, 6000, 64, 256
    plan = cpu_plan(random_graph(N, E, seed=1), N)
    lay = ops.node_layout(plan)
    g = torch.Generator().manual_seed(0)
    xb = torch.randn(N, W, generator=g).bfloat16()
    w3r = (torch.randn(W, W * K, generator=g) / 8).bfloat16()
    hs = torch.randn(lay.e_pad, K, generator=g).bfloat16()
    sched = ops.fused_fwd_schedule(lay, W, K, ring=ring, grid=num_ctas)
    u = torch.zeros(sched.u_rows, W * K, dtype=torch.bfloat16)
    msg = torch.zeros(lay.e_pad, W)
    emu_node_fused_fwd(xb, w3r, hs, u, sched, torch.zeros(sched.num_flags, dtype=torch.int32), msg,
                       N, lay.e_pad, K, W, seed=num_ctas)
    U = (xb.float() @ w3r.float()).bfloat16()
    ref = torch.zeros_like(msg)
    emu_node_msg(hs, U.view(N, W, K), lay, 0, lay.num_units, 0, ref)
    assert torch.equal(msg, ref)
    if not ring:
        assert torch.equal(u, U)


def test_fused_forward_ring_must_cover_the_slack():
    plan = cpu_plan(random_graph(300, 2000, seed=2), 300)
    with pytest.raises(ValueError, match="ring"):
        ops.fused_fwd_schedule(ops.node_layout(plan), 64, 256, ring=2, slack=0.5, grid=8)


# ---- fused backward: schedule -------------------------------------------------------------------
@pytest.mark.parametrize("N,K,grid,ring,slack", [(300, 256, 7, 3, 1.0), (128, 512, 148, 3, 0.5),
                                                 (1000, 1024, 148, 3, 1.0), (5, 256, 4, 4, 2.0),
                                                 (700, 256, 160, 3, 1.0)])
def test_fused_backward_schedule_layout(N, K, grid, ring, slack):
    items, nflags = ops.fused_bwd_items(N, K, 64, grid, ring, slack)
    NB, C = -(-N // 128), K // 128
    pairs = 64 * K // 128
    parts = min(grid, pairs)
    assert nflags == 2 * NB and items.shape[0] % grid == 0
    kinds = items[:, 0]
    D, X = items[kinds == ops.IT_D], items[kinds == ops.IT_X]
    # every (node, chunk) once; every (block, pair) once
    assert len({(int(a), int(b)) for a, b in D[:, 1:3]}) == N * C == len(D)
    assert len(X) == NB * pairs
    base, rem = divmod(pairs, grid)
    for c in range(grid):
        q = items[c::grid]
        npair = base + (c < rem)
        xq = q[q[:, 0] == ops.IT_X]
        assert len(xq) == NB * npair
        # pairs of a block in order, first / last flags, the wait on the first, the signal on the last
        for b in range(NB if npair else 0):
            rows = xq[b * npair:(b + 1) * npair]
            assert (rows[:, 1] == 128 * b).all() and (rows[:, 7] & 0xff).tolist() == list(range(npair))
            assert ((rows[:, 7] >> 8) & 1).tolist() == [1] + [0] * (npair - 1)
            assert ((rows[:, 7] >> 9) & 1).tolist() == [0] * (npair - 1) + [1]
            assert rows[0, 4] == b and (rows[1:, 4] == -1).all()
            assert rows[-1, 6] == NB + b and (rows[:-1, 6] == -1).all()
        # per block one wait (if the slot is being reused) on the CTA's first D item and one signal
        # on its last (one epilogue group; with two the D items would alternate between them)
        dq = q[q[:, 0] == ops.IT_D]
        for g in (0,):
            mine = dq
            blk = mine[:, 1] // 128
            for b in np.unique(blk):
                rows = mine[blk == b]
                assert rows[-1, 6] == b and (rows[:-1, 6] == -1).all()
                if b >= ring:
                    assert rows[0, 4] == NB + b - ring and rows[0, 5] == 4 * parts
                    assert (rows[1:, 4] == -1).all()
                else:
                    assert (rows[:, 4] == -1).all()
                assert (rows[:, 3] // 128 == b % ring).all()
    # the X items of a block wait for every release of that block
    for b in range(NB):
        sig = int(((kinds == ops.IT_D) & (items[:, 6] == b)).sum())
        firsts = X[(X[:, 1] == 128 * b) & (X[:, 4] >= 0)]
        assert (firsts[:, 5] == 4 * sig).all() and len(firsts) == parts
    with pytest.raises(ValueError):
        ops.fused_bwd_items(N, K, 64, grid, 2, 0.5)


def test_fused_backward_emulated_execution_matches_dense_products():
    N, E, K, W = 300, 5000, 256, 64
    plan = cpu_plan(random_graph(N, E, seed=3, isolated=(5, 299)), N)
    lay = ops.node_layout(plan)
    g = torch.Generator().manual_seed(0)
    hs = torch.randn(lay.e_pad, K, generator=g).bfloat16()
    gs = torch.randn(lay.e_pad, W, generator=g).bfloat16()
    gs[lay.dst_pad < 0] = 0
    xb = torch.randn(N, W, generator=g).bfloat16()
    w3k = (torch.randn(W, K * W, generator=g) / 8).bfloat16()
    items, nflags = ops.fused_bwd_items(N, K, W, EMU_GRID, 3)
    sched = ops.FusedSchedule(torch.from_numpy(items), len(items), nflags, 3 * 128, 3, EMU_GRID)
    du = torch.zeros(3 * 128, K * W, dtype=torch.bfloat16)
    dx_part = torch.zeros(EMU_GRID, N, W)
    dwT = torch.full((K * W, W), 7.0)
    assert min(EMU_GRID, W * K // 128) == EMU_GRID
    emu_node_fused_bwd(hs, gs, lay, xb, w3k, du, sched, torch.zeros(nflags, dtype=torch.int32),
                       dx_part, dwT, False)
    dut = torch.zeros(N, K * W, dtype=torch.bfloat16)
    emu_node_du_mn(hs, gs, lay, 0, N, dut)
    assert fro(dx_part.sum(0), dut.float() @ w3k.float().t()) < 1e-6
    assert fro(dwT, dut.float().t() @ xb.float()) < 1e-6


# ---- randomised schedules: no deadlock, same numbers ----------------------------------------------
def test_random_fused_schedules_execute_without_deadlock():
    """Seeded sweep over graph size, grid size, ring and slack for both fused kernels: the in-order
    multi-CTA emulator must drain every list (it raises on a wait that can never be satisfied)
    and reproduce the unfused numbers."""
    rng = np.random.default_rng(7)
    for trial in range(6):
        N = int(rng.integers(3, 420))
        E = int(rng.integers(N, 12 * N))
        grid = int(rng.choice([1, 2, 5, 13, 148]))
        K, W = 256, 64
        plan = cpu_plan(random_graph(N, E, seed=trial), N)
        lay = ops.node_layout(plan)
        g = torch.Generator().manual_seed(trial)
        hs = torch.randn(lay.e_pad, K, generator=g).bfloat16()
        xb = torch.randn(N, W, generator=g).bfloat16()
        # forward
        ring = int(rng.choice([0, 3, 4]))
        slack = float(rng.choice([0.25, 1.0]))
        w3r = (torch.randn(W, W * K, generator=g) / 8).bfloat16()
        sched = ops.fused_fwd_schedule(lay, W, K, ring=ring, slack=slack, grid=grid)
        u = torch.zeros(sched.u_rows, W * K, dtype=torch.bfloat16)
        msg = torch.zeros(lay.e_pad, W)
        emu_node_fused_fwd(xb, w3r, hs, u, sched, torch.zeros(sched.num_flags, dtype=torch.int32),
                           msg, N, lay.e_pad, K, W, seed=trial)
        ref = torch.zeros_like(msg)
        emu_node_msg(hs, (xb.float() @ w3r.float()).bfloat16().view(N, W, K), lay, 0, lay.num_units,
                     0, ref)
        assert torch.equal(msg, ref), (trial, "forward")
        # backward
        bring = int(rng.choice([3, 4]))
        items, nflags = ops.fused_bwd_items(N, K, W, grid, bring, 1.0)
        bs = ops.FusedSchedule(torch.from_numpy(items), len(items), nflags, bring * 128, bring, grid)
        gs = torch.randn(lay.e_pad, W, generator=g).bfloat16()
        gs[lay.dst_pad < 0] = 0
        w3k = (torch.randn(W, K * W, generator=g) / 8).bfloat16()
        parts = min(grid, W * K // 128)
        dx_part = torch.zeros(grid, N, W)
        dwT = torch.zeros(K * W, W)
        emu_node_fused_bwd(hs, gs, lay, xb, w3k, torch.zeros(bring * 128, K * W, dtype=torch.bfloat16),
                           bs, torch.zeros(nflags, dtype=torch.int32), dx_part, dwT, False, seed=trial)
        dut = torch.zeros(N, K * W, dtype=torch.bfloat16)
        emu_node_du_mn(hs, gs, lay, 0, N, dut)
        assert fro(dx_part[:parts].sum(0), dut.float() @ w3k.float().t()) < 1e-6, (trial, "dx")
        assert fro(dwT, dut.float().t() @ xb.float()) < 1e-6, (trial, "dw")
