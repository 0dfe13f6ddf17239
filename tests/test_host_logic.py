"""CPU: host-side logic that needs no GPU — PairData batching, synthetic workloads, module
construction / state-dict layout / argument validation."""
import numpy as np
import pytest
import torch

from molecular_dynamics_neural_operator_b200 import synthetic as syn
from molecular_dynamics_neural_operator_b200.dataset import PairData, collate
from molecular_dynamics_neural_operator_b200.graph_kernel import DenseNet, KernelNN, LpLoss, NNConv_old
from tests.helpers import load_golden


def test_pairdata_collate_offsets_edge_index_by_num_nodes():
    def sample(n, e, seed):
        g = torch.Generator().manual_seed(seed)
        return PairData(x_aminoacid=torch.randint(0, 20, (n,), generator=g),
                        x_position=torch.randn(3, n, 3, generator=g), y=torch.randn(n, 3, generator=g),
                        edge_attr=torch.randn(e, 6, generator=g),
                        edge_index=torch.randint(0, n, (2, e), generator=g))

    a, b, c = sample(4, 5, 0), sample(4, 7, 1), sample(4, 3, 2)
    batch = collate([a, b, c])
    assert batch.num_nodes == 12
    assert batch.x_position.shape == (9, 4, 3)  # [B*W, R, 3] as KernelNN.forward expects
    assert torch.equal(batch.edge_index[:, :5], a.edge_index)
    assert torch.equal(batch.edge_index[:, 5:12], b.edge_index + 4)
    assert torch.equal(batch.edge_index[:, 12:], c.edge_index + 8)
    assert a.__inc__("edge_index") == 4 and a.__inc__("edge_attr") == 0


def test_state_dict_keys_match_the_reference_checkpoint_layout():
    g = load_golden("kernelnn_w16_k64_d2")
    ref_keys = sorted(k[3:] for k in g if k.startswith("sd."))
    model = KernelNN(16, 64, 2, 6, in_width=7, out_width=3, window_size=3, num_residues=12)
    assert sorted(model.state_dict().keys()) == ref_keys
    for k in ref_keys:
        assert tuple(model.state_dict()[k].shape) == g["sd." + k].shape, k
    assert model.conv1.net is model.conv2.net
    # a DataParallel-style checkpoint ("module." prefix, graph_kernel.py:633-639) loads after stripping
    sd = {"module." + k: torch.from_numpy(g["sd." + k]) for k in ref_keys}
    model.load_state_dict({k[len("module."):]: v for k, v in sd.items()})


def test_unsupported_configurations_raise_instead_of_falling_back():
    net = DenseNet([6, 32, 32, 16 * 16], torch.nn.ReLU)
    with pytest.raises(ValueError):
        NNConv_old(16, 16, net, aggr="add")
    with pytest.raises(ValueError):
        NNConv_old(16, 8, DenseNet([6, 32, 32, 16 * 8], torch.nn.ReLU), aggr="mean")
    with pytest.raises(ValueError):
        NNConv_old(16, 16, DenseNet([6, 32, 16 * 16], torch.nn.ReLU), aggr="mean")
    with pytest.raises(ValueError):
        NNConv_old(16, 16, net, aggr="mean", root_weight=False)
    conv = NNConv_old(16, 16, net, aggr="mean")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv(torch.zeros(4, 16), torch.zeros(2, 3, dtype=torch.long), torch.zeros(3, 6))


def test_lploss_matches_golden():
    g = load_golden("lploss")
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    assert np.isclose(LpLoss(size_average=False)(x, y).item(), float(g["rel_sum"]), rtol=1e-12)
    assert np.isclose(LpLoss(size_average=True)(x, y).item(), float(g["rel_mean"]), rtol=1e-12)


def test_synthetic_workloads_have_the_survey_sizes():
    from oracle import gkn_oracle as orc

    wl = syn.WORKLOADS["bba_allatom"]
    traj = syn.trajectory(wl, 3, seed=0)
    assert traj.shape == (3, 504, 3) and traj.dtype == np.float32
    ei, _ = orc.construct_pairdata_arrays(traj[:1], wl.cutoff)
    assert 50000 < ei.shape[1] < 70000  # ~57-60k edges / frame
    ca = syn.WORKLOADS["bba_ca"]
    t2 = syn.trajectory(ca, 1, seed=0)
    bonds = np.linalg.norm(np.diff(t2[0], axis=0), axis=1)
    assert np.all(np.abs(bonds - 3.8) < 1.5)
    xs, ys = syn.window_batch(traj[:, :10], [0], 2)
    assert xs.shape == (1, 2, 10, 3) and ys.shape == (1, 10, 3)
    assert np.array_equal(ys[0], traj[2, :10])
