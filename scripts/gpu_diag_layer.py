"""Print per-tensor errors of the module path vs the golden vectors (bring-up aid)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tests.helpers import grads_of, kernelnn_batch, load_golden, rel_err  # noqa: E402
from tests.test_layer_gpu import build_conv, build_model, fro_err  # noqa: E402
from molecular_dynamics_neural_operator_b200.graph_kernel import LpLoss  # noqa: E402

for prec in ("fp32", "bf16"):
    for name in ["nnconv_w16_k64", "nnconv_w32_k64", "nnconv_w64_k64_isolated"]:
        if prec == "bf16" and "w16" in name:
            continue
        g = load_golden(name)
        conv = build_conv(g, prec)
        x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
        out = conv(x, torch.from_numpy(g["edge_index"]).cuda(), torch.from_numpy(g["edge_attr"]).cuda())
        out.backward(torch.from_numpy(g["gout"]).cuda())
        print(f"{prec} {name}: out rel={rel_err(out, torch.from_numpy(g['out'])):.2e} fro={fro_err(out, torch.from_numpy(g['out'])):.2e}"
              f" | dx rel={rel_err(x.grad, torch.from_numpy(g['grad_x'])):.2e} fro={fro_err(x.grad, torch.from_numpy(g['grad_x'])):.2e}")
        stride = int(g["grad_row_stride"])
        params = dict(conv.named_parameters())
        for k, ref in grads_of(g).items():
            got = params[k].grad
            if k == "net.layers.4.weight" and stride > 1:
                got = got[::stride]
            print(f"     {k:28s} rel={rel_err(got, ref):.2e} fro={fro_err(got, ref):.2e}")
    for name in ["kernelnn_w16_k64_d2", "kernelnn_w32_k64_d1"]:
        if prec == "bf16" and "w16" in name:
            continue
        g = load_golden(name)
        model = build_model(g, prec)
        data = kernelnn_batch(g, torch.float32, "cuda")
        out, latent = model(data, return_latent=True)
        B = int(g["batch"])
        loss = LpLoss(size_average=False)(out.view(B, -1), data.y.view(B, -1))
        loss.backward()
        print(f"{prec} {name}: out rel={rel_err(out, torch.from_numpy(g['out'])):.2e} latent rel={rel_err(latent, torch.from_numpy(g['latent'])):.2e}"
              f" loss={loss.item():.6f} vs {float(g['loss']):.6f}")
        params = dict(model.named_parameters())
        for k, ref in grads_of(g).items():
            print(f"     {k:34s} rel={rel_err(params[k].grad, ref):.2e} fro={fro_err(params[k].grad, ref):.2e}")
