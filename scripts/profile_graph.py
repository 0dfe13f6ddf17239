"""Per-kernel device times of the radius-graph build and CSR prepare (torch.profiler)."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
for name, frames in (("bba_allatom", 64), ("synthetic_10k", 16), ("synthetic_100k", 1)):
    wl = syn.WORKLOADS[name]
    pos = torch.from_numpy(syn.trajectory(wl, frames, seed=0)).to(dev)
    for _ in range(2):
        ei, ea = ops.radius_graph(pos, wl.cutoff)
        ops.prepare_graph(ei, wl.num_atoms * frames)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        ei, ea = ops.radius_graph(pos, wl.cutoff)
        ops.prepare_graph(ei, wl.num_atoms * frames)
        torch.cuda.synchronize()
    print(f"== {name} x{frames}: E={ei.size(1)}")
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", 0)
        if t and str(e.device_type).endswith("CUDA"):
            rows.append((t, e.count, e.key[:90]))
    for t, c, k in sorted(rows, reverse=True)[:14]:
        print(f"   {t:9.1f} us  n={c:3d}  {k}")
