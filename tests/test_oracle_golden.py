"""CPU: pin the oracle (oracle/gkn_oracle.py) against vectors produced by the reference's own
source (tests/golden/make_golden.py).  fp64 on both sides => tolerances are rounding-level."""
import numpy as np
import pytest
import torch

from oracle import gkn_oracle as orc
from tests.helpers import grads_of, kernelnn_batch, load_golden, state_dict_of

NNCONV_CASES = ["nnconv_w16_k64", "nnconv_w32_k64", "nnconv_w64_k64_isolated", "nnconv_w64_k256"]


@pytest.mark.parametrize("name", NNCONV_CASES)
def test_nnconv_forward_backward_matches_reference(name):
    g = load_golden(name)
    sd = {k: v.requires_grad_(True) for k, v in state_dict_of(g).items()}
    net = {k[4:]: v for k, v in sd.items() if k.startswith("net.")}
    x = torch.from_numpy(g["x"]).double().requires_grad_(True)
    ei = torch.from_numpy(g["edge_index"])
    ea = torch.from_numpy(g["edge_attr"]).double()
    out = orc.nnconv_forward(x, ei, ea, net, sd["root"], sd["bias"])
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=1e-12, atol=1e-12)
    out.backward(torch.from_numpy(g["gout"]).double())
    np.testing.assert_allclose(x.grad.numpy(), g["grad_x"], rtol=1e-11, atol=1e-12)
    stride = int(g["grad_row_stride"])
    for k, ref in grads_of(g).items():
        got = sd[k].grad
        if k == "net.layers.4.weight" and stride > 1:
            got = got[::stride]
        np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-10, atol=1e-12, err_msg=k)
    # chunked evaluation (used for graphs whose [E,w,w] tensor does not fit) is the same function
    out2 = orc.nnconv_forward_chunked(x.detach(), ei, ea, {k: v.detach() for k, v in net.items()},
                                      sd["root"].detach(), sd["bias"].detach(), chunk=7)
    np.testing.assert_allclose(out2.numpy(), g["out"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["kernelnn_w16_k64_d2", "kernelnn_w32_k64_d1"])
def test_kernelnn_forward_loss_backward_matches_reference(name):
    g = load_golden(name)
    sd = {k: v.requires_grad_(True) for k, v in state_dict_of(g).items()}
    # conv2.net.* aliases conv1.net.* in the reference (graph_kernel.py:271-273)
    for k in list(sd):
        if k.startswith("conv2.net."):
            sd[k] = sd["conv1.net." + k[len("conv2.net."):]]
    data = kernelnn_batch(g)
    out, latent = orc.kernelnn_forward(sd, data, int(g["depth"]), int(g["window"]),
                                       int(g["num_residues"]), return_latent=True)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(latent.detach().numpy(), g["latent"], rtol=1e-10, atol=1e-12)
    B = int(g["batch"])
    loss = orc.lp_loss_rel(out.view(B, -1), data.y.view(B, -1), size_average=False)
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=1e-12)
    loss.backward()
    for k, ref in grads_of(g).items():
        np.testing.assert_allclose(sd[k].grad.numpy(), ref.numpy(), rtol=1e-8, atol=1e-11, err_msg=k)


def test_kernelnn_config1_shape_matches_reference():
    """BASELINE configs[0] shape (width 32, kernel_width 1024, depth 4, 28 residues, window 10):
    weights rebuilt from the seed the generator used, output / loss / gradients produced by the
    reference's own source."""
    from tests.helpers import seeded_state_dict

    g = load_golden("kernelnn_cfg1_w32_k1024_d4")
    w, K, R = int(g["width"]), int(g["ker_width"]), int(g["num_residues"])
    shapes = {"lstm.weight_ih_l0": (12, 3), "lstm.weight_hh_l0": (12, 3), "lstm.bias_ih_l0": (12,),
              "lstm.bias_hh_l0": (12,), "lstm_fc.weight": (3, 3), "lstm_fc.bias": (3,),
              "emb.weight": (20, 4), "fc1.weight": (w, 7), "fc1.bias": (w,), "fc2.weight": (3, w),
              "fc2.bias": (3,)}
    for c in ("conv1", "conv2"):
        shapes.update({f"{c}.root": (w, w), f"{c}.bias": (w,),
                       f"{c}.net.layers.0.weight": (K, 6), f"{c}.net.layers.0.bias": (K,),
                       f"{c}.net.layers.2.weight": (K, K), f"{c}.net.layers.2.bias": (K,),
                       f"{c}.net.layers.4.weight": (w * w, K), f"{c}.net.layers.4.bias": (w * w,)})
    template = {k: torch.empty(v) for k, v in shapes.items()}
    sd32 = seeded_state_dict(template, int(g["weight_seed"]))
    sd = {}
    for k, v in sd32.items():
        sd[k] = sd["conv1.net." + k[10:]] if k.startswith("conv2.net.") else v.double().requires_grad_(True)
    data = kernelnn_batch(g)
    out = orc.kernelnn_forward(sd, data, int(g["depth"]), int(g["window"]), R)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=1e-10, atol=1e-12)
    B = int(g["batch"])
    loss = orc.lp_loss_rel(out.view(B, -1), data.y.view(B, -1), size_average=False)
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=1e-12)
    loss.backward()
    stride = int(g["grad_row_stride"])
    for k, ref in grads_of(g).items():
        got = sd[k].grad
        if got.dim() == 2 and got.size(0) >= 256:
            got = got[::stride]
        np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-8, atol=1e-11, err_msg=k)


@pytest.mark.parametrize("name", ["pairdata_n28", "pairdata_n300"])
def test_construct_pairdata_matches_reference(name):
    g = load_golden(name)
    ei, ea = orc.construct_pairdata_arrays(g["x_position"], float(g["threshold"]))
    assert ei.dtype == np.int64 and ea.dtype == np.float32
    assert np.array_equal(ei, g["edge_index"])
    assert np.array_equal(ea, g["edge_attr"])
    # scalable kd-tree variant used for large N is the same function
    ei2, ea2 = orc.radius_graph_kdtree(g["x_position"][-1], float(g["threshold"]))
    assert np.array_equal(ei2, g["edge_index"]) and np.array_equal(ea2, g["edge_attr"])


def test_lploss_matches_reference():
    g = load_golden("lploss")
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    np.testing.assert_allclose(orc.lp_loss_rel(x, y, False).numpy(), g["rel_sum"], rtol=1e-14)
    np.testing.assert_allclose(orc.lp_loss_rel(x, y, True).numpy(), g["rel_mean"], rtol=1e-14)


def test_oracle_three_formulations_agree():
    """literal view(E,w,w)+matmul == fused bilinear form == hand-derived deferred backward."""
    torch.manual_seed(0)
    N, E, w, K = 30, 200, 8, 16
    net = orc.make_net_params(6, K, w, seed=3)
    ei = torch.stack([torch.randint(0, N, (E,)), torch.randint(0, N, (E,))])
    ea = torch.randn(E, 6, dtype=torch.float64)
    x = torch.randn(N, w, dtype=torch.float64)
    h = orc.edge_hidden(ea, net)
    W3, b3 = net["layers.4.weight"], net["layers.4.bias"]
    lit = torch.matmul(x[ei[0]].unsqueeze(1), orc.dense_net(ea, net).view(E, w, w)).squeeze(1)
    P = h @ W3.t()
    fused = torch.einsum("ei,eio->eo", x[ei[0]], P.view(E, w, w)) + x[ei[0]] @ b3.view(w, w)
    assert torch.allclose(lit, fused, rtol=1e-12, atol=1e-12)
    # deferred weight gradient: dW3 = S^T h with S_e = x_j (outer) g_e
    g = torch.randn(E, w, dtype=torch.float64)
    W3r = W3.clone().requires_grad_(True)
    msg = torch.einsum("ei,eio->eo", x[ei[0]], (h @ W3r.t()).view(E, w, w))
    (msg * g).sum().backward()
    S = torch.einsum("ei,eo->eio", x[ei[0]], g).reshape(E, w * w)
    assert torch.allclose(W3r.grad, S.t() @ h, rtol=1e-12, atol=1e-12)
