"""GPU: each C-ABI kernel against a plain float64 torch/numpy restatement on seeded inputs,
including ragged / empty / partial-tile shapes."""
import numpy as np
import pytest
import torch

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
