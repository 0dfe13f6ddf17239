"""One radius-graph build (for `ncu --set full -k regex:rg_`): python scripts/profile_graph_ncu.py WORKLOAD FRAMES"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import ops, synthetic as syn  # noqa: E402

name, frames = sys.argv[1], int(sys.argv[2])
wl = syn.WORKLOADS[name]
pos = torch.from_numpy(syn.trajectory(wl, frames, seed=0)).cuda()
for _ in range(2):
    ei, ea = ops.radius_graph(pos, wl.cutoff)
torch.cuda.synchronize()
print(name, frames, "E =", ei.size(1))
