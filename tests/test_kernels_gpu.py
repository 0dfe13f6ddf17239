"""GPU: each C-ABI kernel against a plain float64 torch/numpy restatement on seeded inputs,
including ragged / empty / partial-tile shapes."""
import numpy as np
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rand_graph(N, E, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)])


def contract_ref(h, w3, idx, node, W, kind):
    P = (h.double() @ w3.double().t()).view(-1, W, W)
    nd = node.double()[idx.long()]
    return torch.einsum("ei,eio->eo", nd, P) if kind == "fwd" else torch.einsum("eio,eo->ei", P, nd)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("E,K,W,N", [(1, 64, 64, 3), (127, 64, 32, 9), (300, 64, 64, 40), (640, 128, 64, 77),
                                     (1000, 1024, 64, 504), (257, 256, 32, 28), (200, 128, 128, 50),
                                     (20000, 1024, 64, 504), (19000, 64, 32, 100)])
def test_fused_contraction_fwd_and_dx(precision, E, K, W, N):
    """(19000, 64, 32): > 148 tiles of few chunks => unit ranges split tiles across CTAs."""
    from molecular_dynamics_neural_operator_b200 import ops

    g = torch.Generator().manual_seed(E + K + W)
    h = torch.relu(torch.randn(E, K, generator=g))
    w3 = torch.randn(W * W, K, generator=g) / K ** 0.5
    idx = torch.randint(0, N, (E,), generator=g, dtype=torch.int32)
    node = torch.randn(N, W, generator=g)
    if precision == "bf16":
        hq, wq = h.bfloat16(), w3.bfloat16()
        h, w3 = hq.float(), wq.float()  # reference sees the same rounded operands
    else:
        hq, wq = h, w3
    for kind, fn in (("fwd", ops.contract_fwd), ("dx", ops.contract_dx)):
        out = fn(hq.to(DEV), wq.to(DEV), idx.to(DEV), node.to(DEV), precision)
        ref = contract_ref(h, w3, idx, node, W, kind)
        err = (out.cpu().double() - ref).abs().max().item()
        assert err <= 1e-5 * max(ref.abs().max().item(), 1.0) * 4, (kind, err)


def test_contraction_empty_edge_set_and_bad_shapes():
    from molecular_dynamics_neural_operator_b200 import ops

    h = torch.zeros(0, 64, device=DEV, dtype=torch.bfloat16)
    w3 = torch.zeros(32 * 32, 64, device=DEV, dtype=torch.bfloat16)
    out = ops.contract_fwd(h, w3, torch.zeros(0, dtype=torch.int32, device=DEV), torch.zeros(4, 32, device=DEV))
    assert out.shape == (0, 32)
    with pytest.raises(RuntimeError, match="multiple of 64"):
        ops.contract_fwd(torch.zeros(8, 48, device=DEV, dtype=torch.bfloat16),
                         torch.zeros(32 * 32, 48, device=DEV, dtype=torch.bfloat16),
                         torch.zeros(8, dtype=torch.int32, device=DEV), torch.zeros(4, 32, device=DEV))
    with pytest.raises(RuntimeError, match="width"):
        ops.contract_fwd(torch.zeros(8, 64, device=DEV, dtype=torch.bfloat16),
                         torch.zeros(48 * 48, 64, device=DEV, dtype=torch.bfloat16),
                         torch.zeros(8, dtype=torch.int32, device=DEV), torch.zeros(4, 48, device=DEV))


@pytest.mark.parametrize("N,E,W", [(50, 700, 64), (504, 20000, 64), (300, 5000, 32), (100, 3000, 128),
                                   (40, 100, 16), (10, 0, 64)])
def test_segmented_reductions(N, E, W):
    from molecular_dynamics_neural_operator_b200 import ops
    from oracle import gkn_oracle as orc

    ei = rand_graph(N, E, 7 * N + E)
    plan = ops.prepare_graph(ei.to(DEV), N)
    g = torch.Generator().manual_seed(1)
    msg, dxj = torch.randn(E, W, generator=g), torch.randn(E, W, generator=g)
    xb3, xroot, gb3, add = (torch.randn(N, W, generator=g) for _ in range(4))
    bias = torch.randn(W, generator=g)
    order = plan.perm.cpu().long()
    src_s, dst_s = ei[0][order], ei[1][order]
    ref = orc.scatter_mean((msg + xb3[src_s]).double(), dst_s, N) + xroot.double() + bias.double()
    out = ops.scatter_mean(msg.to(DEV), plan, xb3.to(DEV), xroot.to(DEV), bias.to(DEV), relu=False)
    assert (out.cpu().double() - ref).abs().max().item() <= 1e-5
    out = ops.scatter_mean(msg.to(DEV), plan, None, None, None, relu=True)
    ref = torch.relu(orc.scatter_mean(msg.double(), dst_s, N))  # isolated nodes -> 0
    assert (out.cpu().double() - ref).abs().max().item() <= 1e-5
    ref = torch.zeros(N, W, dtype=torch.float64).index_add_(0, src_s, (dxj + gb3[dst_s]).double()) + add.double()
    out = ops.scatter_src_sum(dxj.to(DEV), plan, gb3.to(DEV), add.to(DEV))
    assert (out.cpu().double() - ref).abs().max().item() <= 1e-4
    gy, y = torch.randn(N, W, generator=g), torch.randn(N, W, generator=g)
    gg, gn = ops.relu_grad_scale(gy.to(DEV), y.to(DEV), plan.inv_cnt, True)
    refg = gy * (y > 0)
    assert torch.equal(gg.cpu(), refg)
    assert torch.allclose(gn.cpu(), refg * plan.inv_cnt.cpu()[:, None])
    if E:
        A = 3
        xs, gs = torch.randn(A, N, W, generator=g), torch.randn(A, N, W, generator=g)
        e0, e1 = E // 3, min(E, E // 3 + 257)
        refS = torch.einsum("aei,aeo->eio", xs[:, src_s[e0:e1]].double(), gs[:, dst_s[e0:e1]].double())
        S = ops.edge_outer_sum(xs.to(DEV), gs.to(DEV), plan, e0, e1, torch.float32)
        assert (S.cpu().double() - refS.reshape(e1 - e0, -1)).abs().max().item() <= 1e-4
        Sb = ops.edge_outer_sum(xs.to(DEV), gs.to(DEV), plan, e0, e1, torch.bfloat16)
        if W == 64:
            # tensor-core path: operands rounded to bf16, fp32 accumulation, one output rounding
            xq, gq = xs.bfloat16().double(), gs.bfloat16().double()
            refq = torch.einsum("aei,aeo->eio", xq[:, src_s[e0:e1]], gq[:, dst_s[e0:e1]]).reshape(e1 - e0, -1)
            err = (Sb.cpu().double() - refq).abs().max().item()
            assert err <= 2.0 ** -8 * max(refq.abs().max().item(), 1.0)
        else:
            assert torch.equal(Sb.cpu(), S.cpu().bfloat16())  # same fp32 sum, one rounding


def test_bias_act_cast_bf16_is_one_rounding():
    from molecular_dynamics_neural_operator_b200 import ops

    x, b = torch.randn(333, 64), torch.randn(64)
    out = ops.bias_act_cast_bf16(x.to(DEV), b.to(DEV), relu=True)
    assert torch.equal(out.cpu(), torch.relu(x + b).bfloat16())
    out = ops.bias_act_cast_bf16(x.to(DEV), None, relu=False)
    assert torch.equal(out.cpu(), x.bfloat16())


def test_launch_counter_counts_library_kernels():
    from molecular_dynamics_neural_operator_b200 import _lib, ops

    n0 = _lib.launch_count()
    ops.bias_act_cast_bf16(torch.randn(8, 64, device=DEV))
    assert _lib.launch_count() == n0 + 1


@pytest.mark.parametrize("act_dt,out_dt", [(torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32),
                                           (torch.float32, torch.float32), (torch.float32, torch.bfloat16)])
def test_mask_cast_colsum(act_dt, out_dt):
    from molecular_dynamics_neural_operator_b200 import ops

    g = torch.Generator().manual_seed(3)
    for rows, cols in [(1, 64), (37, 128), (5000, 1024)]:
        dz = torch.randn(rows, cols, generator=g)
        act = torch.relu(torch.randn(rows, cols, generator=g)).to(act_dt)
        colsum = torch.full((cols,), 0.5)
        ref = dz * (act.float() > 0)
        cs = colsum.to(DEV)
        out = ops.mask_cast_colsum(dz.to(DEV), act.to(DEV), cs, out_dt, accumulate=True)
        assert torch.equal(out.cpu(), ref.to(out_dt))
        assert torch.allclose(cs.cpu(), 0.5 + ref.sum(0), rtol=1e-5, atol=1e-4)
        cs2 = torch.full((cols,), 7.0, device=DEV)
        ops.mask_cast_colsum(dz.to(DEV), act.to(DEV), cs2, out_dt, accumulate=False)
        assert torch.allclose(cs2.cpu(), ref.sum(0), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("N,E,W", [(50, 700, 64), (504, 20000, 64), (300, 5000, 32), (100, 3000, 128),
                                   (40, 100, 16), (10, 0, 64),
                                   (80001, 200000, 64), (80001, 150000, 128), (80001, 100000, 32)])  # batched-node path
def test_segment_matvec_forward_and_backward_forms(N, E, W):
    from molecular_dynamics_neural_operator_b200 import ops

    ei = rand_graph(N, E, 11 * N + E)
    plan = ops.prepare_graph(ei.to(DEV), N)
    g = torch.Generator().manual_seed(5)
    rows = torch.randn(E, W, generator=g)
    x, nin = torch.randn(N, W, generator=g), torch.randn(N, W, generator=g)
    M1, M2 = torch.randn(W, W, generator=g) / W ** 0.5, torch.randn(W, W, generator=g) / W ** 0.5
    bias = torch.randn(W, generator=g)
    order = plan.perm.cpu().long()
    src_s, dst_s = ei[0][order], ei[1][order]
    inv = plan.inv_cnt.cpu().double()[:, None]
    # forward form
    r = torch.zeros(N, W, dtype=torch.float64).index_add_(0, dst_s, rows.double()) * inv
    agg = torch.zeros(N, W, dtype=torch.float64).index_add_(0, dst_s, x.double()[src_s]) * inv
    ref = torch.relu(r + agg @ M1.double() + nin.double() @ M2.double() + bias.double())
    out, agg_out = ops.segment_matvec(rows.to(DEV), None, plan.dst_ptr, plan.src, x.to(DEV), plan.inv_cnt,
                                      M1.to(DEV), nin.to(DEV), M2.to(DEV), bias.to(DEV), True, want_agg=True)
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-5
    assert (agg_out.cpu().double() - agg).abs().max().item() <= 2e-5
    # backward form: rows are in sorted-position order, summed per SOURCE node through src_perm
    r = torch.zeros(N, W, dtype=torch.float64).index_add_(0, src_s, rows.double())
    agg = torch.zeros(N, W, dtype=torch.float64).index_add_(0, src_s, x.double()[dst_s])
    ref = r + agg @ M1.double() + nin.double() @ M2.double()
    out, none = ops.segment_matvec(rows.to(DEV), plan.src_perm, plan.src_ptr, plan.dst, x.to(DEV), None,
                                   M1.to(DEV), nin.to(DEV), M2.to(DEV), None, False)
    assert none is None
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-4


def test_mask_cast_colsum_side_sums_give_layer0_weight_gradient():
    from molecular_dynamics_neural_operator_b200 import ops

    g = torch.Generator().manual_seed(4)
    rows, cols = 3000, 256
    dz = torch.randn(rows, cols, generator=g)
    act = torch.relu(torch.randn(rows, cols, generator=g)).bfloat16()
    side = torch.randn(rows, 6, generator=g) * 10
    cs = torch.zeros(7, cols, device=DEV)
    out = ops.mask_cast_colsum(dz.to(DEV), act.to(DEV), cs, torch.float32, side=side.to(DEV))
    ref = (dz * (act.float() > 0)).double()
    assert torch.equal(out.cpu().double(), ref)
    assert torch.allclose(cs[0].cpu().double(), ref.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(cs[1:].cpu().double(), side.double().t() @ ref, rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_edge_layer0(dt):
    from molecular_dynamics_neural_operator_b200 import ops

    g = torch.Generator().manual_seed(6)
    for E, K in [(1, 64), (2, 64), (3, 64), (777, 128), (5000, 1024)]:
        ea = torch.randn(E + 1, 6, generator=g) * 5
        w1, b1 = torch.randn(K, 6, generator=g), torch.randn(K, generator=g)
        dev = ea.to(DEV)
        # rows [0, E): 16-byte aligned input (four edges per float4 group); rows [1, E + 1): not
        for view, rows in ((dev[:E], ea[:E]), (dev[1:], ea[1:])):
            out = ops.edge_layer0(view, w1.to(DEV), b1.to(DEV), dt)
            ref = torch.relu(rows.double() @ w1.double().t() + b1.double())
            tol = 1e-5 if dt == torch.float32 else 2.0 ** -8
            assert ((out.cpu().double() - ref).abs() / (ref.abs() + 1)).max().item() <= tol


def test_mask_cast_colsum_bf16_input_in_place():
    from molecular_dynamics_neural_operator_b200 import ops

    g = torch.Generator().manual_seed(8)
    dz = torch.randn(4000, 512, generator=g).bfloat16()
    act = torch.relu(torch.randn(4000, 512, generator=g)).bfloat16()
    cs = torch.zeros(512, device=DEV)
    d = dz.to(DEV)
    out = ops.mask_cast_colsum(d, act.to(DEV), cs, torch.bfloat16)
    ref = dz.float() * (act.float() > 0)
    assert out.data_ptr() == d.data_ptr()  # same dtype: masked in place
    assert torch.equal(out.cpu().float(), ref)
    assert torch.allclose(cs.cpu(), ref.sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("N,E,A,K", [(40, 3000, 12, 256), (504, 20000, 12, 1024), (100, 5000, 3, 256),
                                     (300, 12000, 16, 512), (7, 130, 1, 256), (64, 128 * 40, 8, 256),
                                     (5, 129, 5, 256)])
def test_fused_edge_hidden_gradient(N, E, A, K):
    """gkn_edge_grad_dh_bf16 (S synthesised by the tensor core, never stored) against the staged
    restatement  S = sum_a xs[a, src] (x) gs[a, dst];  dh = S @ W3  in float64 on bf16-rounded
    operands.  Random graphs: Poisson in-degrees, tiles spanning 2..8 destinations, ragged last tile."""
    from molecular_dynamics_neural_operator_b200 import ops

    W = 64
    g = torch.Generator().manual_seed(N + E + A)
    plan = ops.prepare_graph(rand_graph(N, E, N + E).to(DEV), N)
    assert ops.edge_grad_slots(plan) <= ops.EDGE_GRAD_MAX_SLOTS
    xs = torch.randn(A, N, W, generator=g).to(DEV)
    gs = torch.randn(A, N, W, generator=g).to(DEV)
    w3 = (torch.randn(W * W, K, generator=g) / W).bfloat16().to(DEV)
    dh = ops.edge_grad_dh(xs, gs, plan, w3.t().contiguous())
    assert dh.shape == (E, K) and dh.dtype == torch.bfloat16
    xq = xs.bfloat16().double()[:, plan.src.long()]  # [A, E, W]
    gq = gs.bfloat16().double()[:, plan.dst.long()]
    S = torch.einsum("aei,aeo->eio", xq, gq).reshape(E, W * W)
    Sq = S.float().bfloat16().double()  # the kernel rounds S to bf16 once (its MMA operand)
    ref = Sq @ w3.double()
    err = (dh.double() - ref).norm() / ref.norm()
    assert err < 2 ** -7, err
    # and against the existing staged kernel (bf16 S through HBM, library GEMM)
    S2 = ops.edge_outer_sum(xs, gs, plan, 0, E, torch.bfloat16)
    ref2 = (S2.double() @ w3.double())
    assert (dh.double() - ref2).norm() / ref2.norm() < 2 ** -7


def test_fused_edge_gradient_rejects_low_degree_graphs():
    from molecular_dynamics_neural_operator_b200 import ops

    N, E, W, K = 2000, 9000, 64, 256
    plan = ops.prepare_graph(rand_graph(N, E, 5).to(DEV), N)
    assert ops.edge_grad_slots(plan) > ops.EDGE_GRAD_MAX_SLOTS
    xs = torch.zeros(2, N, W, device=DEV)
    with pytest.raises(ValueError, match="more than 8 destination"):
        ops.edge_grad_dh(xs, xs, plan, torch.zeros(K, W * W, dtype=torch.bfloat16, device=DEV))


@pytest.mark.parametrize("rows,W", [(1, 64), (31, 32), (8064 * 6, 64), (5000, 128), (100000, 32), (700, 16), (90, 48)])
def test_xty_matches_float64_product(rows, W):
    """gkn_xty_f32 (root / bias / b3 gradients): fp32 accumulation in a fixed order against a
    float64 product, 1e-5 of the largest entry; run-to-run identical."""
    from molecular_dynamics_neural_operator_b200 import ops

    gen = torch.Generator().manual_seed(rows + W)
    x = torch.randn(rows, W, generator=gen).cuda()
    g = torch.randn(rows, W, generator=gen).cuda()
    out, gsum = ops.xty(x, g, colsum=True)
    ref = x.double().t() @ g.double()
    assert rel_err(out, ref) <= 1e-5
    assert rel_err(gsum, g.double().sum(0)) <= 1e-5
    out2, none = ops.xty(x, g)
    assert none is None and torch.equal(out2, out)
