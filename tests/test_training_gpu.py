"""GPU: the training step on the bench workload shape (small K) — loss decreases, results are
deterministic run to run, and a two-process NCCL run (when 2 GPUs are visible) matches."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_step(seed=0, frames=4):
    from molecular_dynamics_neural_operator_b200 import synthetic as syn
    from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN
    from molecular_dynamics_neural_operator_b200.train import Trainer

    wl = syn.Workload("t", 120, 32, 64, 2, 5.0, 4, 14.0)
    traj = syn.trajectory(wl, wl.window + frames + 1, seed=seed)
    xw, y = syn.window_batch(traj, list(range(frames)), wl.window)
    torch.manual_seed(seed)
    model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3,
                     window_size=wl.window, num_residues=wl.num_atoms, precision="bf16").cuda()
    trainer = Trainer(model, lr=1e-3, weight_decay=0.0)
    aa = torch.from_numpy(syn.residue_ids(wl)).cuda()
    xw, y = torch.from_numpy(xw).cuda(), torch.from_numpy(y).cuda()

    def step():
        batch = syn.make_batch(xw, aa, y, wl.cutoff)
        return trainer.step(batch, frames).item()

    return model, step


def test_training_reduces_the_loss_and_is_deterministic():
    _, step = make_step()
    losses = [step() for _ in range(12)]
    assert losses[-1] < 0.9 * losses[0]
    _, step2 = make_step()
    losses2 = [step2() for _ in range(12)]
    assert losses == losses2  # no atomics on the hot loop => bitwise reproducible


def test_cuda_extension_is_loaded_in_tree():
    """The driver checks that GPU tests really ran native code: the in-tree .so must be mapped."""
    from molecular_dynamics_neural_operator_b200 import _lib

    _lib.load()
    maps = open("/proc/self/maps").read()
    assert "libgkn_b200.so" in maps
