"""GPU parity of the fused producer/consumer forward (csrc/gkn_node_fused.cu,
gkn_node_fused_fwd_bf16) through the C ABI: U tiles and messages against a torch fp32 restatement
on identical bf16 operands, kept-U and ring modes, widths 32 / 64 / 128, graphs large enough that
every CTA of the persistent grid takes part and the ring wraps several times; then the whole
stack with the fused forward against the two-kernel form (reference graph_kernel.py:200-202).

Tolerances as in tests/test_node_path_gpu.py: fp32 accumulation of bf16 products, so only the
summation order differs from an fp32 matmul of the same operands - one bf16 ulp for U (bf16
output), 1e-5 relative (max-norm) for messages computed from the SAME U.
"""
import pytest
import torch

from tests.helpers import rel_err
from tests.test_node_path_cpu import cpu_plan, emu_node_msg, fro, random_graph

pytestmark = pytest.mark.gpu


def _case(N, E, W, K, seed):
    from molecular_dynamics_neural_operator_b200 import ops

    ei = random_graph(N, E, seed)
    plan = ops.prepare_graph(ei.cuda(), N)
    lay = ops.node_layout(plan)
    g = torch.Generator().manual_seed(seed)
    Wp = max(W, 64)
    xb = torch.zeros(N, Wp)
    xb[:, :W] = torch.randn(N, W, generator=g)
    w3r = torch.zeros(Wp, W * K)
    w3r[:W] = torch.randn(W, W * K, generator=g) / 8
    hs = torch.randn(lay.e_pad, K, generator=g)
    return ei, lay, xb.bfloat16(), w3r.bfloat16(), hs.bfloat16()


@pytest.mark.parametrize("N,E,W,K,ring", [
    (300, 9000, 64, 256, 0),       # 3 node blocks, kept U
    (2000, 60000, 64, 1024, 0),    # every CTA busy, K = 1024 (the model's kernel width)
    (2000, 60000, 64, 1024, 4),    # ring of 4 blocks wraps 4 times
    (1500, 20000, 64, 256, 3),     # smallest legal ring
    (700, 30000, 32, 256, 0),      # w = 32: x and W3r zero padded to 64
    (700, 30000, 32, 1024, 3),
    (600, 20000, 128, 256, 0),     # w = 128: two reduction blocks per tile
    (600, 20000, 128, 1024, 3),
    (100, 300, 64, 256, 3),        # a single partial node block, ring larger than the graph
    (129, 3000, 64, 256, 0),       # second block holds one node
])
def test_fused_forward_matches_fp32_restatement(N, E, W, K, ring):
    from molecular_dynamics_neural_operator_b200 import ops

    ei, lay, xb, w3r, hs = _case(N, E, W, K, seed=N + W + ring)
    sched = ops.fused_fwd_schedule(lay, W, K, ring=ring)
    u = torch.zeros(sched.u_rows, W * K, dtype=torch.bfloat16, device="cuda")
    msg = torch.full((lay.e_pad, W), float("nan"), device="cuda")
    flags = torch.zeros(sched.num_flags, dtype=torch.int32, device="cuda")
    ops.node_fused_fwd(xb.cuda(), w3r.cuda(), hs.cuda(), u, sched, flags, msg, N, lay.e_pad, K, W)
    torch.cuda.synchronize()
    U32 = xb.float() @ w3r.float()
    NB, J = -(-N // 128), W * K // 256
    if not ring:
        # one bf16 rounding of the same fp32 sums (up to summation order)
        assert rel_err(u.float(), U32) < 2.0 ** -7
        assert float((u.cpu().float() - U32).abs().max()) <= 2.0 ** -8 * float(U32.abs().max()) * 1.01
        u_ref = u.cpu()
    else:
        u_ref = U32.bfloat16()
    # messages: fp32 restatement on the U the kernel itself produced (kept mode) / on bf16(U32)
    cl = ops.node_layout(cpu_plan(ei, N))
    ref = torch.zeros(lay.e_pad, W)
    emu_node_msg(hs, u_ref.view(N, W, K), cl, 0, cl.num_units, 0, ref)
    rows = torch.zeros(lay.e_pad, dtype=torch.bool)
    for g_, row0, r, _ in cl.units.tolist():
        rows[row0:row0 + r] = True
    got = msg.cpu()
    assert bool(torch.isfinite(got[rows]).all())
    tol = 1e-5 if not ring else 2.0 ** -7  # ring mode: U itself may differ by one bf16 ulp
    assert rel_err(got[rows], ref[rows]) < tol
    # counters: every tile and (ring mode) every consumer signalled through its 4 epilogue warps
    f = flags.cpu()
    assert f[:NB].tolist() == [4 * J] * NB
    if ring:
        cnt = torch.bincount(cl.units[:, 0].long() // 128, minlength=NB)
        assert f[NB:2 * NB].tolist() == (4 * cnt).tolist()


def test_fused_forward_is_deterministic_and_reentrant_on_flags():
    """Two launches on fresh counters give bit-identical U and messages."""
    from molecular_dynamics_neural_operator_b200 import ops

    N, E, W, K = 1000, 40000, 64, 512
    ei, lay, xb, w3r, hs = _case(N, E, W, K, seed=9)
    sched = ops.fused_fwd_schedule(lay, W, K, ring=0)
    outs = []
    for _ in range(2):
        u = torch.zeros(sched.u_rows, W * K, dtype=torch.bfloat16, device="cuda")
        msg = torch.zeros(lay.e_pad, W, device="cuda")
        flags = torch.zeros(sched.num_flags, dtype=torch.int32, device="cuda")
        ops.node_fused_fwd(xb.cuda(), w3r.cuda(), hs.cuda(), u, sched, flags, msg, N, lay.e_pad, K, W)
        outs.append((u.cpu(), msg.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("N,W,K", [(300, 64, 256), (2000, 64, 1024), (129, 32, 256), (500, 128, 512)])
def test_own_u_gemm_matches_fp32_restatement(N, W, K):
    """U = x @ W3r from the producer items alone (the default forward's dense GEMM)."""
    from molecular_dynamics_neural_operator_b200 import ops

    g = torch.Generator().manual_seed(N + W)
    Wp = max(W, 64)
    xb = torch.zeros(N, Wp)
    xb[:, :W] = torch.randn(N, W, generator=g)
    w3r = torch.zeros(Wp, W * K)
    w3r[:W] = torch.randn(W, W * K, generator=g) / 8
    xb, w3r = xb.bfloat16(), w3r.bfloat16()
    u = torch.full((N, W * K), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.node_u_gemm(xb.cuda(), w3r.cuda(), u, N, K, W)
    ref = xb.float() @ w3r.float()
    assert rel_err(u.float(), ref) < 2.0 ** -7


@pytest.mark.parametrize("fwd_mode,fused_bwd", [("gemm", False), ("1", False), ("1", True), ("gemm", True)],
                         ids=["own_gemm", "fused_fwd", "fused_fwd_bwd", "own_gemm_fused_bwd"])
@pytest.mark.parametrize("needs_grad", [True, False], ids=["train_kept_u", "inference"])
def test_stack_variants_match_library_gemm_form(monkeypatch, needs_grad, fwd_mode, fused_bwd):
    from molecular_dynamics_neural_operator_b200 import functional as GF
    from molecular_dynamics_neural_operator_b200 import ops
    from oracle import gkn_oracle as orc

    N, E, W, K, depth = 700, 40000, 64, 256, 2
    ei = random_graph(N, E, 4)
    plan = ops.prepare_graph(ei.cuda(), N)
    g = torch.Generator().manual_seed(3)
    net = orc.make_net_params(6, K, W, seed=5, dtype=torch.float32)
    base = {"x": torch.randn(N, W, generator=g), "ea": torch.randn(E, 6, generator=g),
            "root1": torch.randn(W, W, generator=g) / 8, "bias1": torch.randn(W, generator=g) / 8,
            "root2": torch.randn(W, W, generator=g) / 8, "bias2": torch.randn(W, generator=g) / 8}
    base.update({"net." + k: v for k, v in net.items()})
    results = []
    for variant in (True, False):
        monkeypatch.setattr(GF, "NODE_FUSED", variant and fwd_mode == "1")
        monkeypatch.setattr(GF, "NODE_U_GEMM", variant and fwd_mode == "gemm")
        monkeypatch.setattr(GF, "NODE_FUSED_BWD", variant and fused_bwd)
        monkeypatch.setattr(GF, "_FUSED_BWD_MIN_DEGREE", 0)  # take it whatever the degree
        t = {k: v.cuda().requires_grad_(needs_grad) for k, v in base.items()}
        with torch.set_grad_enabled(needs_grad):
            out = GF._GKNStackNode.apply(
                t["x"], t["ea"].index_select(0, plan.perm_long), t["net.layers.0.weight"],
                t["net.layers.0.bias"], t["net.layers.2.weight"], t["net.layers.2.bias"],
                t["net.layers.4.weight"], t["net.layers.4.bias"], t["root1"], t["bias1"],
                t["root2"], t["bias2"], plan, (0,) * depth + (1,) * depth, True, "bf16")
        if needs_grad:
            out.backward(torch.ones_like(out))
        results.append((out.detach().cpu(), {k: v.grad.cpu() for k, v in t.items()} if needs_grad else {}))
    # same operands, same roundings except the summation order inside U / dU: well inside one
    # layer's bound
    assert fro(results[0][0], results[1][0]) < 2.0 ** -7
    for k in results[0][1]:
        assert fro(results[0][1][k], results[1][1][k]) < 2.0 ** -5, k


# ---- fused backward (csrc/gkn_node_fused_bwd.cu) --------------------------------------------------
@pytest.mark.parametrize("N,E,K,ring,with_dx,with_dw", [
    (300, 9000, 256, 3, True, True),      # 3 node blocks, partial last block
    (2000, 60000, 1024, 3, True, True),   # the model's kernel width: every CTA owns 3-4 column pairs
    (2000, 60000, 1024, 4, True, True),   # larger ring
    (1000, 30000, 512, 3, True, False),   # dx only (frozen edge MLP)
    (1000, 30000, 512, 3, False, True),   # weight gradient only (first application, x0 needs no grad)
    (130, 3000, 256, 4, True, True),      # second block holds two nodes; isolated nodes below
])
def test_fused_backward_matches_separate_products(N, E, K, ring, with_dx, with_dw):
    """dU from the grouped kernel (gkn_node_du_mn_bf16, itself checked against fp32 in
    tests/test_node_path_gpu.py) multiplied out in fp32 by torch is the reference: the fused kernel
    rounds dU to bf16 identically and accumulates the two products in fp32 (reduction lengths w*K
    and N: 1e-4 relative on the max-norm scale covers the summation order)."""
    from molecular_dynamics_neural_operator_b200 import ops

    W = 64
    ei = random_graph(N, E, seed=N + K, isolated=(3, N - 1))
    plan = ops.prepare_graph(ei.cuda(), N)
    lay = ops.node_layout(plan)
    g = torch.Generator().manual_seed(K + ring)
    bf = torch.bfloat16
    hs = torch.randn(lay.e_pad, K, generator=g).to(bf).cuda()
    gs = torch.randn(lay.e_pad, W, generator=g).to(bf).cuda()
    gs[lay.dst_pad < 0] = 0
    xb = torch.randn(N, W, generator=g).to(bf).cuda()
    w3k = (torch.randn(W, K * W, generator=g) / 8).to(bf).cuda()
    sched = ops.fused_bwd_schedule(N, K, W, hs.device, ring=ring)
    grid, parts = ops.fused_bwd_grid(K, W)
    du = torch.zeros(sched.u_rows, K * W, dtype=bf, device="cuda")
    dx_part = torch.full((parts, N, W), float("nan"), device="cuda") if with_dx else None
    dwT = torch.full((K * W, W), float("nan"), device="cuda") if with_dw else None
    flags = torch.zeros(sched.num_flags, dtype=torch.int32, device="cuda")
    ops.node_fused_bwd(hs, gs, lay, xb, w3k, du, sched, flags, dx_part, dwT, False)
    torch.cuda.synchronize()
    dut = torch.empty(N, K * W, dtype=bf, device="cuda")
    ops.node_du_mn(hs, gs, lay, 0, N, dut)
    NB = -(-N // 128)
    f = flags.cpu()
    # counters: four warps per release (the last D item a CTA has in a block) and per consumed block
    it = sched.items.cpu()
    rel = torch.bincount(it[(it[:, 0] == ops.IT_D) & (it[:, 6] >= 0), 6].long(), minlength=NB)
    assert f[:NB].tolist() == (4 * rel).tolist()
    assert f[NB:].tolist() == [4 * parts] * NB
    if with_dx:
        addend = torch.randn(N, W, generator=g).cuda()
        dx = ops.sum_partials(dx_part, addend, torch.empty_like(addend))
        ref = addend + dut.float() @ w3k.float().t()
        assert rel_err(dx, ref) < 1e-4
    if with_dw:
        ref = dut.float().t() @ xb.float()
        assert rel_err(dwT, ref) < 1e-4
        # a second launch accumulates
        flags.zero_()
        ops.node_fused_bwd(hs, gs, lay, xb, w3k, du, sched, flags, dx_part, dwT, True)
        assert rel_err(dwT, 2 * ref) < 1e-4


def test_fused_backward_is_deterministic():
    from molecular_dynamics_neural_operator_b200 import ops

    N, E, K, W = 1500, 50000, 512, 64
    ei = random_graph(N, E, seed=21)
    plan = ops.prepare_graph(ei.cuda(), N)
    lay = ops.node_layout(plan)
    g = torch.Generator().manual_seed(2)
    bf = torch.bfloat16
    hs = torch.randn(lay.e_pad, K, generator=g).to(bf).cuda()
    gs = torch.randn(lay.e_pad, W, generator=g).to(bf).cuda()
    gs[lay.dst_pad < 0] = 0
    xb = torch.randn(N, W, generator=g).to(bf).cuda()
    w3k = (torch.randn(W, K * W, generator=g) / 8).to(bf).cuda()
    sched = ops.fused_bwd_schedule(N, K, W, hs.device)
    _, parts = ops.fused_bwd_grid(K, W)
    outs = []
    for _ in range(2):
        du = torch.zeros(sched.u_rows, K * W, dtype=bf, device="cuda")
        dx_part = torch.empty(parts, N, W, device="cuda")
        dwT = torch.empty(K * W, W, device="cuda")
        flags = torch.zeros(sched.num_flags, dtype=torch.int32, device="cuda")
        ops.node_fused_bwd(hs, gs, lay, xb, w3k, du, sched, flags, dx_part, dwT, False)
        dx = ops.sum_partials(dx_part, None, torch.empty(N, W, device="cuda"))
        outs.append((dx.cpu(), dwT.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
