"""Two training steps of the bench shape (for `ncu --set full -k regex:node_gemm_kernel`):
launches per step: 12 x node_msg (forward), 12 x node_du (backward), 1 x node_dh."""
import sys
from pathlib import Path

import torch

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
for _ in range(2):
    batch = syn.make_batch(xw, aa, y, wl.cutoff, prepare=True)
    trainer.step(batch, B)
torch.cuda.synchronize()
print("done")
