"""Launch gkn_edge_outer_sum (tensor-core variant) a few times at the bench chunk size (ncu target)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
N, W, A, E = 4032, 64, 12, 479872
g = torch.Generator(device=dev).manual_seed(0)
ei = torch.stack([torch.randint(0, N, (E,), device=dev, generator=g), torch.randint(0, N, (E,), device=dev, generator=g)])
plan = ops.prepare_graph(ei, N)
xs = torch.randn(A, N, W, device=dev, generator=g)
gs = torch.randn(A, N, W, device=dev, generator=g)
for i in range(4):
    S = ops.edge_outer_sum(xs, gs, plan, i * 16384, (i + 1) * 16384, torch.bfloat16)
torch.cuda.synchronize()
print("ok")
