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


def _write_npz(path, frames=9, R=7, seed=0):
    rng = np.random.default_rng(seed)
    pc = rng.normal(size=(frames, 3, R)).astype(np.float32)  # [F, 3, R] as in the reference files
    cms = np.empty(frames, dtype=object)
    for f in range(frames):
        e = int(rng.integers(3, 12))
        cms[f] = np.concatenate([rng.integers(0, R, e), rng.integers(0, R, e)]).astype(np.int64)
    np.savez(path, contact_map=cms, point_cloud=pc, rmsd=rng.random(frames).astype(np.float32),
             amino_acids=rng.integers(0, 20, R).astype(np.int64))
    return pc, cms


def test_contact_map_dataset_follows_the_reference_getitem(tmp_path):
    from molecular_dynamics_neural_operator_b200.dataset import ContactMapDataset

    pc, cms = _write_npz(tmp_path / "traj.npz")
    W, H = 3, 2
    ds = ContactMapDataset(str(tmp_path / "traj.npz"), window_size=W, horizon=H)
    assert len(ds) == len(cms) - W - H + 1
    pos = np.transpose(pc, [0, 2, 1])
    for idx in (0, 2, len(ds) - 1):
        d = ds[idx]
        ei = cms[idx].reshape(2, -1)
        # literal restatement of dataset.py:180-204 (per-edge python loop)
        ea = np.array([np.concatenate((pos[idx, i, :], pos[idx, j, :])).flatten()
                       for i, j in zip(ei[0], ei[1])])
        assert torch.equal(d.edge_index, torch.from_numpy(ei))
        assert np.array_equal(d.edge_attr.numpy(), ea.astype(np.float32))
        assert np.array_equal(d.x_position.numpy(), pos[idx:idx + W])
        assert np.array_equal(d.y.numpy(), pos[idx + W + H - 1])
        assert d.x_aminoacid.dtype == torch.long and d.num_nodes == 7
    with pytest.raises(ValueError):
        ContactMapDataset(str(tmp_path / "traj.npz"), window_size=20, horizon=1)
    # a directory of files is concatenated in sorted order (dataset.py:133-143)
    _write_npz(tmp_path / "traj2.npz", seed=1)
    assert len(ContactMapDataset(str(tmp_path), window_size=1, horizon=1)) == 18 - 1


def test_checkpoint_roundtrip_uses_the_reference_layout(tmp_path):
    from molecular_dynamics_neural_operator_b200.train import Trainer, load_checkpoint, save_checkpoint

    model = KernelNN(16, 64, 1, 6, in_width=7, out_width=3, window_size=2, num_residues=5)
    trainer = Trainer(model)
    save_checkpoint(tmp_path / "best.pt", 7, trainer)
    ckpt = torch.load(tmp_path / "best.pt", weights_only=False)
    assert set(ckpt) == {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict"}
    assert all(k.startswith("module.") for k in ckpt["model_state_dict"])  # graph_kernel.py:528, :635
    assert "module.conv2.net.layers.4.weight" in ckpt["model_state_dict"]
    model2 = KernelNN(16, 64, 1, 6, in_width=7, out_width=3, window_size=2, num_residues=5)
    assert load_checkpoint(tmp_path / "best.pt", model2, Trainer(model2)) == 7
    for a, b in zip(model.parameters(), model2.parameters()):
        assert torch.equal(a, b)


def test_clock_sampler_reports_clocks_and_power_per_timed_loop():
    """bench.py's `clocks` object: overall median / min / max, throttle reasons, and per timed loop
    (marks) the median clock and power - parsed from nvidia-smi's csv lines (no GPU needed)."""
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("bench_mod", Path(__file__).resolve().parents[1] / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class _Proc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    s = bench.ClockSampler(0)
    s.proc = _Proc()
    row = "{}, 1965, Not Active, Not Active, Not Active, {}, {}"
    s.mark("value")
    s.lines += [row.format(1900, "Active", 760.0), row.format(1880, "Active", 770.0)]
    s.mark("e2e")
    s.lines += [row.format(1700, "Active", 950.0), "garbage", row.format(1740, "Not Active", "[N/A]")]
    out = s.stop()
    assert out["sm_mhz"] == 1810.0 and out["sm_min_mhz"] == 1700.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["samples"] == 4
    assert out["by_loop"]["value"] == {"sm_mhz": 1890.0, "sm_min_mhz": 1880.0, "power_w": 765.0,
                                       "samples": 2}
    assert out["by_loop"]["e2e"]["sm_mhz"] == 1720.0 and out["by_loop"]["e2e"]["power_w"] == 950.0
