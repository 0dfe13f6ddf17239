"""Launch the fused contraction (fwd and dx) a few times at bench size — the command profiled by
ncu for profiles/ (Measurement)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import ops  # noqa: E402

E, K, W, N = 8 * 59984, 1024, 64, 8 * 504
if len(sys.argv) > 1:
    E = int(sys.argv[1])
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
h = torch.relu(torch.randn(E, K, device=dev, generator=g)).bfloat16()
w3 = (torch.randn(W * W, K, device=dev, generator=g) / 32).bfloat16()
src = torch.randint(0, N, (E,), device=dev, generator=g, dtype=torch.int32)
x = torch.randn(N, W, device=dev, generator=g)
for _ in range(4):
    ops.contract_fwd(h, w3, src, x, "bf16")
    ops.contract_dx(h, w3, src, x, "bf16")
torch.cuda.synchronize()
print("ok")
