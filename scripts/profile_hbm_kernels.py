"""The HBM-bound kernels at the BASELINE shapes, a few launches each, for
`ncu --set full -k regex:"segment_matvec|scatter_mean|rg_"`:
segmented node update (forward / backward form) at w = 64 (config 2, 64 frames) and w = 128
(config 5, one frame), the plain segmented mean, and both passes of the radius graph."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, frames in (("bba_allatom", 64), ("synthetic_100k", 1)):
    wl = syn.WORKLOADS[name]
    pos = torch.from_numpy(syn.trajectory(wl, frames, seed=0)).to(dev)
    for _ in range(2):
        flush.zero_()
        ei, ea, rp = ops.radius_graph(pos, wl.cutoff, return_row_ptr=True)
    N, E, W = wl.num_atoms * frames, ei.size(1), wl.width
    plan = ops.plan_from_csr(ei, rp, N)
    msg = torch.randn(E, W, device=dev, generator=g)
    x = torch.randn(N, W, device=dev, generator=g)
    M = torch.randn(W, W, device=dev, generator=g)
    bias = torch.randn(W, device=dev, generator=g)
    for _ in range(2):
        flush.zero_()
        ops.segment_matvec(msg, None, plan.dst_ptr, plan.src, x, plan.inv_cnt, M, x, M, bias, True)
        flush.zero_()
        ops.segment_matvec(msg, plan.src_perm, plan.src_ptr, plan.dst, x, None, M, x, M, None, False)
        flush.zero_()
        ops.scatter_mean(msg, plan, None, None, bias, True)
    del msg, x, ei, ea
torch.cuda.synchronize()
print("done")
