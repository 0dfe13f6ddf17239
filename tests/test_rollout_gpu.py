"""GPU: device-resident rollout (recursive_propagation, reference graph_kernel.py:396-413) against
the oracle's CPU rollout (model forward -> slide window -> SciPy graph rebuild)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_recursive_propagation_matches_oracle_rollout():
    from molecular_dynamics_neural_operator_b200 import synthetic as syn
    from molecular_dynamics_neural_operator_b200.dataset import PairData
    from molecular_dynamics_neural_operator_b200.graph_kernel import (KernelNN, construct_pairdata,
                                                                      recursive_propagation)
    from oracle import gkn_oracle as orc

    wl = syn.Workload("r", 40, 32, 64, 1, 7.0, 3, 12.0)
    traj = syn.trajectory(wl, 8, seed=5)
    aa = torch.from_numpy(syn.residue_ids(wl))
    torch.manual_seed(1)
    model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3,
                     window_size=wl.window, num_residues=wl.num_atoms, precision="fp32").cuda()
    with torch.no_grad():  # make the model a near-identity on the last frame so the rollout stays physical
        model.fc2.weight.mul_(0.01)

    class DS:
        def __getitem__(self, i):
            return construct_pairdata(traj[i:i + wl.window], aa, threshold=wl.cutoff)

    steps = 4
    fc = recursive_propagation(model, DS(), torch.device("cuda"), num_steps=steps, starting_points=[0, 2],
                               threshold=wl.cutoff)
    assert len(fc) == 2 * steps and all(isinstance(f, PairData) for f in fc)

    sd = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}
    k = 0
    for start in (0, 2):
        win = traj[start:start + wl.window].astype(np.float64)
        for _ in range(steps):
            ei, ea = orc.construct_pairdata_arrays(win.astype(np.float32), wl.cutoff)
            data = SimpleNamespace(x_aminoacid=aa, x_position=torch.from_numpy(win.astype(np.float32)).double(),
                                   edge_index=torch.from_numpy(ei), edge_attr=torch.from_numpy(ea).double())
            out = orc.kernelnn_forward(sd, data, wl.depth, wl.window, wl.num_atoms)
            win = np.concatenate([win[1:], out.numpy()[None]], axis=0)
            got = fc[k]
            k += 1
            assert np.allclose(got.x_position.cpu().numpy(), win, rtol=1e-4, atol=1e-4)
            ei2, ea2 = orc.construct_pairdata_arrays(got.x_position.cpu().numpy(), wl.cutoff)
            assert np.array_equal(got.edge_index.cpu().numpy(), ei2)  # graph of the new last frame
            assert np.array_equal(got.edge_attr.cpu().numpy(), ea2)
