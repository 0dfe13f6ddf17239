"""GPU parity of the node-first factorisation (csrc/gkn_node_gemm.cu) through the C ABI:
each grouped-GEMM kernel against a torch fp32 restatement on identical bf16 operands, the whole
stack against the edge-first tensor-core path and the fp64 oracle.

Tolerances: the kernels accumulate bf16 products in fp32, so against an fp32 matmul of the SAME
bf16 operands only the summation order differs: 1e-5 relative (max-norm) for fp32 outputs, one
bf16 rounding (one ulp = 2^-7 relative at most, on the max-norm scale) for bf16 outputs.
Stack level: bounds of tests/test_layer_gpu.py.
"""
import pytest
import torch

from tests.helpers import rel_err
from tests.test_node_path_cpu import (_run_oracle_stack, _stack_case, cpu_plan, emu_node_dh,
                                      emu_node_dh_mn, emu_node_du, emu_node_du_mn,
                                      emu_node_gather, emu_node_msg, emu_relu_bits, fro,
                                      random_graph)

pytestmark = pytest.mark.gpu

BF16_OUT_TOL = 2.0 ** -7
W = 64


def _plan_on_gpu(ei, N):
    from molecular_dynamics_neural_operator_b200 import ops

    return ops.prepare_graph(ei.cuda(), N)


def _graphs():
    # (name, edge_index, N): ragged degrees, groups longer than one and two tiles, isolated nodes
    g = torch.Generator().manual_seed(0)
    yield "ragged", random_graph(40, 3000, 1), 40
    yield "long_groups", random_graph(6, 2000, 2), 6
    yield "isolated", random_graph(50, 700, 3, isolated=(0, 17, 49)), 50
    src = torch.randint(0, 300, (20000,), generator=g)
    dst = torch.randint(0, 300, (20000,), generator=g)
    yield "mid", torch.stack([src, dst]), 300


@pytest.mark.parametrize("case", list(_graphs()), ids=lambda c: c[0])
@pytest.mark.parametrize("K", [256, 1024])
def test_grouped_gemm_kernels_match_fp32_restatement(case, K):
    from molecular_dynamics_neural_operator_b200 import ops

    _, ei, N = case
    plan = _plan_on_gpu(ei, N)
    lay = ops.node_layout(plan)
    cl = ops.node_layout(cpu_plan(ei, N))  # the same layout from the CPU restatement of the plan
    assert lay.e_pad == cl.e_pad and torch.equal(lay.units.cpu(), cl.units)
    assert torch.equal(lay.pptr.cpu(), cl.pptr)
    gen = torch.Generator().manual_seed(K + N)
    bf = torch.bfloat16
    E_pad = lay.e_pad
    hs = torch.randn(E_pad, K, generator=gen).to(bf)
    u = (torch.randn(N, W, K, generator=gen) / 8).to(bf)
    gn = torch.randn(N, W, generator=gen)
    A = 3
    gs_in = torch.randn(A, N, W, generator=gen)
    ut = (torch.randn(A, N, K, W, generator=gen) / 8).to(bf)

    # ---- msg ----
    msg = torch.full((E_pad, W), float("nan"), device="cuda")
    ops.node_msg(hs.cuda(), u.cuda(), lay, 0, lay.num_units, 0, msg)
    ref = torch.zeros(E_pad, W)
    emu_node_msg(hs, u, cl, 0, cl.num_units, 0, ref)
    assert rel_err(msg, ref) <= 1e-5

    # ---- gather (+ transpose) ----
    gs = torch.empty(E_pad, W, dtype=bf, device="cuda")
    gsT = torch.empty(W, E_pad, dtype=bf, device="cuda")
    ops.node_gather(gn.cuda(), lay, gs, gsT)
    rgs, rgsT = torch.empty(E_pad, W, dtype=bf), torch.empty(W, E_pad, dtype=bf)
    # the CPU layout's dst_pad follows the CPU plan's tie order; use the GPU layout's indices
    import copy
    cl2 = copy.copy(cl)
    cl2.dst_pad = lay.dst_pad.cpu()
    emu_node_gather(gn, cl2, rgs, rgsT)
    assert torch.equal(gs.cpu(), rgs) and torch.equal(gsT.cpu(), rgsT)

    # ---- dU ----
    hsT = hs.t().contiguous()
    dut = torch.full((N, K * W), float("nan"), dtype=bf, device="cuda")
    ops.node_du(hsT.cuda(), gsT, lay, 0, N, dut)
    rdut = torch.zeros(N, K * W, dtype=bf)
    emu_node_du(hsT, rgsT, cl, 0, N, rdut)
    assert rel_err(dut.float(), rdut.float()) <= BF16_OUT_TOL

    # ---- dh (all applications at once) ----
    gs_all = torch.empty(A, E_pad, W, dtype=bf, device="cuda")
    rgs_all = torch.empty(A, E_pad, W, dtype=bf)
    for a in range(A):
        ops.node_gather(gs_in[a].cuda(), lay, gs_all[a], None)
        emu_node_gather(gs_in[a], cl2, rgs_all[a], None)
    dh = torch.full((E_pad, K), float("nan"), dtype=bf, device="cuda")
    ops.node_dh(gs_all, ut.cuda(), lay, 0, lay.num_units, 0, dh)
    rdh = torch.zeros(E_pad, K, dtype=bf)
    emu_node_dh(rgs_all, ut, cl, 0, cl.num_units, 0, rdh)
    assert rel_err(dh.float(), rdh.float()) <= BF16_OUT_TOL
    pad = (lay.dst_pad < 0).cpu()
    assert float(dh.float().cpu()[pad].abs().max() if pad.any() else 0.0) == 0.0

    # ---- the MN-major operand variants give the same numbers from the untransposed buffers ----
    dut2 = torch.full((N, K * W), float("nan"), dtype=bf, device="cuda")
    ops.node_du_mn(hs.cuda(), gs, lay, 0, N, dut2)
    assert rel_err(dut2.float(), dut.float()) <= BF16_OUT_TOL
    dh2 = torch.full((E_pad, K), float("nan"), dtype=bf, device="cuda")
    ops.node_dh_mn(gs_all, ut.transpose(2, 3).contiguous().cuda(), lay, 0, lay.num_units, 0, dh2)
    assert rel_err(dh2.float(), dh.float()) <= BF16_OUT_TOL


@pytest.mark.parametrize("case", list(_graphs()), ids=lambda c: c[0])
@pytest.mark.parametrize("width,K", [(32, 256), (32, 1024), (128, 256), (128, 1024)])
def test_backward_kernels_at_other_widths_match_fp32_restatement(case, width, K):
    """gkn_node_gather_bf16 / gkn_node_du_mn_bf16 / gkn_node_dh_mn_bf16 at w = 32 (gs rows padded
    to 64 values, dU accumulated 64 wide and stored 32 wide, two UMMA K-steps per application in
    the dh pass) and w = 128 (two k-blocks per application in the dh pass)."""
    import copy

    from molecular_dynamics_neural_operator_b200 import ops

    _, ei, N = case
    plan = _plan_on_gpu(ei, N)
    lay = ops.node_layout(plan)
    cl = copy.copy(ops.node_layout(cpu_plan(ei, N)))
    cl.dst_pad = lay.dst_pad.cpu()
    gen = torch.Generator().manual_seed(K + N + width)
    bf = torch.bfloat16
    E_pad, Wp, A = lay.e_pad, ops.gs_width(width), 3
    hs = torch.randn(E_pad, K, generator=gen).to(bf)
    gs_in = torch.randn(A, N, width, generator=gen)
    u = (torch.randn(A, N, width, K, generator=gen) / 8).to(bf)

    gs_all = torch.full((A, E_pad, Wp), float("nan"), dtype=bf, device="cuda")
    rgs_all = torch.empty(A, E_pad, Wp, dtype=bf)
    for a in range(A):
        ops.node_gather(gs_in[a].cuda(), lay, gs_all[a], None)
        emu_node_gather(gs_in[a], cl, rgs_all[a], None)
    assert torch.equal(gs_all.cpu(), rgs_all)
    with pytest.raises(RuntimeError, match="w = 64 only"):
        ops.node_gather(gs_in[0].cuda(), lay, gs_all[0],
                        torch.empty(width, E_pad, dtype=bf, device="cuda"))

    dut = torch.full((N, K * width), float("nan"), dtype=bf, device="cuda")
    ops.node_du_mn(hs.cuda(), gs_all[0], lay, 0, N, dut)
    rdut = torch.zeros(N, K * width, dtype=bf)
    emu_node_du_mn(hs, rgs_all[0], cl, 0, N, rdut)
    assert rel_err(dut.float(), rdut.float()) <= BF16_OUT_TOL
    # node ranges give the same rows
    parts = torch.full((N, K * width), float("nan"), dtype=bf, device="cuda")
    for n0 in range(0, N, 17):
        n1 = min(N, n0 + 17)
        ops.node_du_mn(hs.cuda(), gs_all[0], lay, n0, n1, parts[n0:n1])
    assert torch.equal(parts, dut)

    dh = torch.full((E_pad, K), float("nan"), dtype=bf, device="cuda")
    ops.node_dh_mn(gs_all, u.cuda(), lay, 0, lay.num_units, 0, dh)
    rdh = torch.zeros(E_pad, K, dtype=bf)
    emu_node_dh_mn(rgs_all, u, cl, 0, cl.num_units, 0, rdh)
    assert rel_err(dh.float(), rdh.float()) <= BF16_OUT_TOL
    pad = (lay.dst_pad < 0).cpu()
    assert float(dh.float().cpu()[pad].abs().max() if pad.any() else 0.0) == 0.0
    # ... and in node ranges (the last range ends at the end of the U buffer: w = 32 boxes there)
    dh_parts = torch.full((E_pad, K), float("nan"), dtype=bf, device="cuda")
    for n0 in range(0, N, 17):
        n1 = min(N, n0 + 17)
        ops.node_dh_mn(gs_all, u[:, n0:n1].contiguous().cuda(), lay, lay.unit_first[n0],
                       lay.unit_first[n1], n0, dh_parts)
    assert torch.equal(dh_parts, dh)


@pytest.mark.parametrize("case", list(_graphs()), ids=lambda c: c[0])
@pytest.mark.parametrize("width,K", [(64, 1024), (64, 256), (32, 512), (128, 1024)])
def test_dh_pass_with_fused_relu_backward(case, width, K):
    """gkn_relu_bits_bf16 + gkn_node_dh_mn_relu_bf16: the bit pattern of act > 0 (bit-exact
    against the CPU restatement), the same dh values where the activation is positive, zeros
    elsewhere (bit-exact against masking the plain kernel's output), column sums within fp32
    summation-order error of a float64 sum, node-range launches chained through ``accumulate``,
    and run-to-run identical (no atomics)."""
    from molecular_dynamics_neural_operator_b200 import ops

    _, ei, N = case
    plan = _plan_on_gpu(ei, N)
    lay = ops.node_layout(plan)
    gen = torch.Generator().manual_seed(K + N + width)
    bf = torch.bfloat16
    E_pad, Wp, A = lay.e_pad, ops.gs_width(width), 2
    gs_all = torch.zeros(A, E_pad, Wp, dtype=bf, device="cuda")
    for a in range(A):
        ops.node_gather(torch.randn(N, width, generator=gen).cuda(), lay, gs_all[a], None)
    u = (torch.randn(A, N, width, K, generator=gen) / 8).to(bf).cuda()
    act = torch.relu(torch.randn(E_pad, K, generator=gen)).to(bf).cuda()  # about half zeros
    act[0, :8] = torch.tensor([0.0, -0.0, 1e-30, -1.0, float("inf"), 2.0, 0.0, 3.0]).to(bf)

    bits = ops.relu_bits(act)
    assert torch.equal(bits.cpu(), emu_relu_bits(act.cpu()))
    plain = torch.empty(E_pad, K, dtype=bf, device="cuda")
    ops.node_dh_mn(gs_all, u, lay, 0, lay.num_units, 0, plain)
    want = torch.where(act > 0, plain, torch.zeros((), dtype=bf, device="cuda"))
    got = torch.full((E_pad, K), float("nan"), dtype=bf, device="cuda")
    colsum = torch.full((K,), float("nan"), device="cuda")
    ops.node_dh_mn(gs_all, u, lay, 0, lay.num_units, 0, got, act_bits=bits, colsum=colsum)
    assert torch.equal(got, want)
    ref = want.double().sum(0)
    scale = want.double().abs().sum(0).max()
    assert float((colsum.double() - ref).abs().max()) <= 1e-5 * float(scale)

    again = torch.empty_like(got)
    colsum2 = torch.empty_like(colsum)
    ops.node_dh_mn(gs_all, u, lay, 0, lay.num_units, 0, again, act_bits=bits, colsum=colsum2)
    assert torch.equal(again, got) and torch.equal(colsum2, colsum)

    parts = torch.full((E_pad, K), float("nan"), dtype=bf, device="cuda")
    csum = torch.zeros(K, device="cuda")
    for n0 in range(0, N, 17):
        n1 = min(N, n0 + 17)
        ops.node_dh_mn(gs_all, u[:, n0:n1].contiguous(), lay, lay.unit_first[n0],
                       lay.unit_first[n1], n0, parts, act_bits=bits, colsum=csum, accumulate=True)
    assert torch.equal(parts, got)
    assert float((csum.double() - ref).abs().max()) <= 1e-5 * float(scale)


def test_grouped_gemm_node_range_chunks_cover_the_same_result():
    from molecular_dynamics_neural_operator_b200 import ops

    ei, N, K = random_graph(64, 6000, 9), 64, 256
    plan = _plan_on_gpu(ei, N)
    lay = ops.node_layout(plan)
    gen = torch.Generator().manual_seed(1)
    bf = torch.bfloat16
    hs = torch.randn(lay.e_pad, K, generator=gen).to(bf).cuda()
    u = (torch.randn(N, W, K, generator=gen) / 8).to(bf).cuda()
    whole = torch.zeros(lay.e_pad, W, device="cuda")
    ops.node_msg(hs, u, lay, 0, lay.num_units, 0, whole)
    parts = torch.zeros(lay.e_pad, W, device="cuda")
    for n0 in range(0, N, 24):
        n1 = min(N, n0 + 24)
        ops.node_msg(hs, u[n0:n1].contiguous(), lay, lay.unit_first[n0], lay.unit_first[n1], n0,
                     parts)
    assert torch.equal(whole, parts)
    gsT = torch.randn(W, lay.e_pad, generator=gen).to(bf).cuda()
    gsT[:, lay.dst_pad < 0] = 0
    hsT = hs.t().contiguous()
    d_whole = torch.empty(N, K * W, dtype=bf, device="cuda")
    ops.node_du(hsT, gsT, lay, 0, N, d_whole)
    d_parts = torch.empty(N, K * W, dtype=bf, device="cuda")
    for n0 in range(0, N, 24):
        n1 = min(N, n0 + 24)
        ops.node_du(hsT, gsT, lay, n0, n1, d_parts[n0:n1])
    assert torch.equal(d_whole, d_parts)


def _gpu_stack(fn, ei, p, depth, N, relu):
    from molecular_dynamics_neural_operator_b200 import ops

    plan = ops.prepare_graph(ei.cuda(), N)
    t = {k: v.clone().cuda().requires_grad_(True) for k, v in p.items()}
    ea_sorted = t["ea"].index_select(0, plan.perm_long)
    out = fn.apply(t["x"], ea_sorted, t["net.layers.0.weight"], t["net.layers.0.bias"],
                   t["net.layers.2.weight"], t["net.layers.2.bias"], t["net.layers.4.weight"],
                   t["net.layers.4.bias"], t["root1"], t["bias1"], t["root2"], t["bias2"], plan,
                   (0,) * depth + (1,) * depth, relu, "bf16")
    return out, t


@pytest.mark.parametrize("mn,save_u", [(True, True), (True, False), (False, False)],
                         ids=["mn_saved_u", "mn_recomputed_u", "kmajor"])
@pytest.mark.parametrize("N,E,K,depth", [(20, 3000, 256, 1), (120, 8000, 1024, 2)])
def test_node_first_stack_matches_edge_first_stack_and_oracle(monkeypatch, N, E, K, depth, mn, save_u):
    """No activation between applications (smooth: only bf16 rounding differs)."""
    from molecular_dynamics_neural_operator_b200 import functional as GF

    monkeypatch.setattr(GF, "NODE_MN", mn)
    monkeypatch.setattr(GF, "_SAVE_U_BYTES", (1 << 40) if save_u else 0)
    monkeypatch.setattr(GF, "_NODE_CHUNK_BYTES", 64 * K * 2 * 7)  # node ranges of 7 nodes
    monkeypatch.setattr(GF, "_NODE_DH_CHUNK_BYTES", 2 * depth * 64 * K * 2 * 5)  # and of 5

    ei, p, depth = _stack_case(N, E, K, depth, seed=N)
    out_n, tn = _gpu_stack(GF._GKNStackNode, ei, p, depth, N, relu=False)
    out_e, te = _gpu_stack(GF._GKNStack, ei, p, depth, N, relu=False)
    ref, tr = _run_oracle_stack(ei, p, depth, relu=False)
    assert fro(out_n.cpu(), ref) < 2.0 ** -5
    assert fro(out_n.cpu(), out_e.cpu()) < 2.0 ** -5
    gout = torch.randn(out_n.shape, generator=torch.Generator().manual_seed(5))
    out_n.backward(gout.cuda())
    out_e.backward(gout.cuda())
    ref.backward(gout.double())
    for k in p:
        # upstream of the edge MLP's own ReLUs the error is mask flips: 2^-4 on these graphs of
        # thousands of edges (tests/test_layer_gpu.py header; 2^-2 is kept for < 300 edges only)
        tol = 2.0 ** -4 if k.startswith(("net.layers.0", "net.layers.2", "ea")) else 2.0 ** -5
        assert fro(tn[k].grad.cpu(), tr[k].grad) < tol, k
        assert fro(tn[k].grad.cpu(), te[k].grad.cpu()) < tol, k


@pytest.mark.parametrize("width", [32, 128])
@pytest.mark.parametrize("save_u", [True, False], ids=["saved_u", "recomputed_u"])
def test_node_first_training_at_other_widths_matches_edge_first_and_oracle(monkeypatch, width, save_u):
    """Config 1 (w = 32) and w = 128 train on the node-first path: outputs and every gradient
    against the fp64 oracle and the edge-first stack (no activation between applications)."""
    from molecular_dynamics_neural_operator_b200 import functional as GF
    from molecular_dynamics_neural_operator_b200 import ops

    N, E, K, depth = 120, 8000, 1024, 2
    monkeypatch.setattr(GF, "_SAVE_U_BYTES", (1 << 40) if save_u else 0)
    monkeypatch.setattr(GF, "_NODE_CHUNK_BYTES", width * K * 2 * 7)
    monkeypatch.setattr(GF, "_NODE_DH_CHUNK_BYTES", 2 * depth * width * K * 2 * 5)
    assert GF.node_path_supported(width, K, E, "bf16", needs_grad=True)
    ei, p, depth = _stack_case(N, E, K, depth, seed=width, W=width)
    ops.KERNEL_TIMER.enable()
    try:
        out_n, tn = _gpu_stack(GF._GKNStackNode, ei, p, depth, N, relu=False)
        gout = torch.randn(out_n.shape, generator=torch.Generator().manual_seed(5))
        out_n.backward(gout.cuda())
        got = ops.KERNEL_TIMER.collect()
    finally:
        ops.KERNEL_TIMER.disable()
    assert got["node_du"]["count"] >= 2 * depth and got["node_dh"]["count"] >= 1
    out_e, te = _gpu_stack(GF._GKNStack, ei, p, depth, N, relu=False)
    out_e.backward(gout.cuda())
    ref, tr = _run_oracle_stack(ei, p, depth, relu=False)
    ref.backward(gout.double())
    assert fro(out_n.cpu(), ref) < 2.0 ** -5
    for k in p:
        tol = 2.0 ** -4 if k.startswith(("net.layers.0", "net.layers.2", "ea")) else 2.0 ** -5
        assert fro(tn[k].grad.cpu(), tr[k].grad) < tol, k
        assert fro(tn[k].grad.cpu(), te[k].grad.cpu()) < tol, k


@pytest.mark.parametrize("width", [32, 128])
def test_forward_only_node_first_at_other_widths(width):
    """Rollout (no gradients) takes the node-first message kernel at w = 32 and 128 as well."""
    from molecular_dynamics_neural_operator_b200 import functional as GF
    from molecular_dynamics_neural_operator_b200 import ops
    from oracle import gkn_oracle as orc

    N, E, K = 90, 5000, 256
    ei = random_graph(N, E, seed=width)
    gen = torch.Generator().manual_seed(width)
    net = orc.make_net_params(6, K, width, seed=1, dtype=torch.float32)
    x = torch.randn(N, width, generator=gen)
    ea = torch.randn(E, 6, generator=gen)
    root = torch.randn(width, width, generator=gen) / 8
    bias = torch.randn(width, generator=gen) / 8
    ref = orc.nnconv_forward(x.double(), ei, ea.double(), {k: v.double() for k, v in net.items()},
                             root.double(), bias.double())
    plan = ops.prepare_graph(ei.cuda(), N)
    c = lambda t: t.cuda()  # noqa: E731
    args = (c(x), c(ea).index_select(0, plan.perm_long), c(net["layers.0.weight"]),
            c(net["layers.0.bias"]), c(net["layers.2.weight"]), c(net["layers.2.bias"]),
            c(net["layers.4.weight"]), c(net["layers.4.bias"]), c(root), c(bias), None, None, plan,
            (0,), False, "bf16")
    assert GF.node_path_supported(width, K, E, "bf16", needs_grad=False)
    with torch.no_grad():
        out_n = GF._GKNStackNode.apply(*args)
        out_e = GF._GKNStack.apply(*args)
    assert fro(out_n.cpu(), ref) < 2.0 ** -7
    assert fro(out_n.cpu(), out_e.cpu()) < 2.0 ** -7


def test_node_first_is_the_default_for_the_bench_shape():
    """KernelNN at w = 64 / K = 1024 in bf16 runs the grouped-GEMM kernels (launch accounting)."""
    from molecular_dynamics_neural_operator_b200 import functional as GF
    from molecular_dynamics_neural_operator_b200 import ops

    assert GF.FACTORIZATION == "node"
    ei, p, depth = _stack_case(30, 2000, 1024, 1, seed=2)
    ops.KERNEL_TIMER.enable()
    try:
        out, t = _gpu_stack(GF._GKNStackNode, ei, p, depth, 30, relu=True)
        out.sum().backward()
        got = ops.KERNEL_TIMER.collect()
    finally:
        ops.KERNEL_TIMER.disable()
    fwd = got.get("node_fused_fwd", got.get("node_msg"))
    assert fwd["count"] == 2 and got["node_dh"]["count"] == 1
    # 67 edges per node: the fused backward (one launch per application, no library GEMM)
    assert GF.NODE_FUSED_BWD and got["node_fused_bwd"]["count"] == 2 and "node_du" not in got


def test_low_degree_graphs_keep_the_three_kernel_backward():
    """Below ~64 edges per node the fused backward's per-node overheads outweigh its savings
    (DESIGN.md section 6): such graphs take gkn_node_du_mn_bf16 + the two dense products."""
    from molecular_dynamics_neural_operator_b200 import functional as GF
    from molecular_dynamics_neural_operator_b200 import ops

    ei, p, depth = _stack_case(200, 3000, 256, 1, seed=4)  # 15 edges per node
    ops.KERNEL_TIMER.enable()
    try:
        out, t = _gpu_stack(GF._GKNStackNode, ei, p, depth, 200, relu=True)
        out.sum().backward()
        got = ops.KERNEL_TIMER.collect()
    finally:
        ops.KERNEL_TIMER.disable()
    assert got["node_du"]["count"] == 2 and "node_fused_bwd" not in got
