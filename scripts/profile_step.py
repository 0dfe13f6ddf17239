"""Per-kernel device-time breakdown of the bench training step with torch.profiler (CUPTI).
Iteration aid; the committed evidence is the ncu launch list under profiles/."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import synthetic as syn  # noqa: E402
from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN  # noqa: E402
from molecular_dynamics_neural_operator_b200.train import Trainer  # noqa: E402

wl = syn.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "bba_allatom"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3, window_size=wl.window,
                 num_residues=wl.num_atoms, precision="bf16").to(dev)
trainer = Trainer(model)
traj = syn.trajectory(wl, wl.window + B + 1, seed=0)
xw, y = syn.window_batch(traj, list(range(B)), wl.window)
xw, y = torch.from_numpy(xw).to(dev), torch.from_numpy(y).to(dev)
aa = torch.from_numpy(syn.residue_ids(wl)).to(dev)


def step():
    batch = syn.make_batch(xw, aa, y, wl.cutoff)
    return trainer.step(batch, B)


for _ in range(3):
    step()
torch.cuda.synchronize()
steps = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0)
    if t and e.device_type is not None and str(e.device_type).endswith("CUDA"):
        rows.append((t / 1e3 / steps, e.count // steps, e.key[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(10):
    step()
torch.cuda.synchronize()
print(f"wall per step: {(time.perf_counter() - t0) * 100:.2f} ms")
print(f"device time per step: {tot:.2f} ms")
for t, c, k in rows[:45]:
    print(f"{t:9.3f} ms {100 * t / tot:5.1f}%  n={c:4d}  {k}")
