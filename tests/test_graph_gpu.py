"""GPU parity of the radius-graph build and the CSR prepare: bit-exact against the reference's
golden vectors and the SciPy oracle; structural properties at full size."""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["pairdata_n28", "pairdata_n300"])
def test_construct_pairdata_bit_exact_vs_reference_golden(name):
    from molecular_dynamics_neural_operator_b200.graph_kernel import construct_pairdata

    g = load_golden(name)
    n = g["x_position"].shape[1]
    data = construct_pairdata(g["x_position"], torch.zeros(n, dtype=torch.long), float(g["threshold"]))
    assert data.edge_index.dtype == torch.int64 and data.edge_attr.dtype == torch.float32
    assert np.array_equal(data.edge_index.cpu().numpy(), g["edge_index"])
    assert np.array_equal(data.edge_attr.cpu().numpy(), g["edge_attr"])
    assert data.x_position.shape == g["x_position"].shape


@pytest.mark.parametrize("F,N,L,thr", [(1, 1, 1.0, 8.0), (1, 28, 20.0, 8.0), (5, 100, 12.0, 4.2),
                                        (2, 504, 17.1, 8.0), (3, 64, 300.0, 4.2), (1, 3000, 31.0, 4.2),
                                        (2, 50, 5.0, 0.5), (1, 40, 10.0, 1000.0),
                                        (1, 2300, 5.0, 9.0),   # > 2048 staged candidates: per-atom fallback
                                        (3, 700, 9.0, 3.0), (2, 2048, 30.0, 4.2), (7, 33, 9.0, 4.2),
                                        (300, 28, 14.0, 8.0), (1, 1999, 12.0, 8.0),
                                        (5000, 12, 9.0, 4.2)])  # > 4096 pass-1 CTAs: generic row-pointer scan
def test_radius_graph_bit_exact_vs_scipy_oracle(F, N, L, thr):
    from molecular_dynamics_neural_operator_b200 import ops
    from oracle import gkn_oracle as orc

    rng = np.random.default_rng(F * 1000 + N)
    pos = (rng.random((F, N, 3)) * L - L / 3).astype(np.float32)
    ei_ref, ea_ref = orc.radius_graph_batched(pos, thr)
    ei, ea = ops.radius_graph(torch.from_numpy(pos).cuda(), thr)
    assert np.array_equal(ei.cpu().numpy(), ei_ref)
    assert np.array_equal(ea.cpu().numpy(), ea_ref)


def test_radius_graph_boundary_pairs_use_the_float64_predicate():
    """Pairs sitting exactly at / one ulp around the cutoff, and duplicated atoms."""
    from molecular_dynamics_neural_operator_b200 import ops
    from oracle import gkn_oracle as orc

    thr = 4.2
    base = np.zeros((8, 3), dtype=np.float32)
    base[1] = [4.2, 0, 0]
    base[2] = [np.nextafter(np.float32(4.2), np.float32(0)), 0, 0]
    base[3] = [np.nextafter(np.float32(4.2), np.float32(9)), 0, 0]
    base[4] = [3.0, 2.9393876, 0]  # |d| ~ 4.2
    base[5] = [0, 0, 0]  # duplicate of atom 0
    base[6] = [2.4248712, 2.4248712, 2.4248712]
    base[7] = [-4.2, 0, 0]
    ei_ref, ea_ref = orc.radius_graph_batched(base[None], thr)
    ei, ea = ops.radius_graph(torch.from_numpy(base[None]).cuda(), thr)
    assert np.array_equal(ei.cpu().numpy(), ei_ref) and np.array_equal(ea.cpu().numpy(), ea_ref)


@pytest.mark.parametrize("n_side,thr", [(6, 1.0), (8, 2.0), (13, 1.0)])
def test_radius_graph_lattice_at_exactly_the_cutoff_spacing(n_side, thr):
    """Integer lattice with neighbours at exactly the cutoff: nearly every hit word is inside the
    fp32 screening band, so the small-frame kernel's float64 fix-up list overflows and its
    all-float64 path runs; the strict `<` must still exclude the distance == cutoff pairs."""
    from molecular_dynamics_neural_operator_b200 import ops
    from oracle import gkn_oracle as orc

    g = np.arange(n_side, dtype=np.float32)
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(1, -1, 3)
    pos = np.concatenate([pos, pos[:, ::-1] + np.float32(100.0)], 0)  # second frame, other order
    ei_ref, ea_ref = orc.radius_graph_batched(pos, thr)
    ei, ea = ops.radius_graph(torch.from_numpy(np.ascontiguousarray(pos)).cuda(), thr)
    assert np.array_equal(ei.cpu().numpy(), ei_ref) and np.array_equal(ea.cpu().numpy(), ea_ref)


def test_radius_graph_mixed_density_large_frames():
    """Large-frame path with every row flavour in one frame: rows served from the saved neighbour
    lists, rows longer than the list capacity (rebuilt by pass 2) and cells too crowded for the
    one-warp-per-cell kernel (CTA-wide worklist kernel)."""
    from molecular_dynamics_neural_operator_b200 import ops
    from oracle import gkn_oracle as orc

    rng = np.random.default_rng(7)
    frames = []
    for f in range(2):
        bulk = rng.random((4000, 3)) * 50.0
        blob_a = 10.0 + rng.random((100, 3)) * 2.5      # 100 neighbours > 64, neighbourhood <= 256
        blob_b = 30.0 + rng.random((280, 3)) * 3.0      # > 256 atoms around: crowded cells
        p = np.concatenate([bulk, blob_a, blob_b]).astype(np.float32)
        frames.append(p[rng.permutation(len(p))])
    pos = np.stack(frames)
    ei_ref, ea_ref = orc.radius_graph_batched(pos, 4.2)
    assert ei_ref.shape[1] < 40 * pos.shape[0] * pos.shape[1]  # stays on the list-driven fill
    ei, ea = ops.radius_graph(torch.from_numpy(pos).cuda(), 4.2)
    assert np.array_equal(ei.cpu().numpy(), ei_ref) and np.array_equal(ea.cpu().numpy(), ea_ref)


def test_radius_graph_full_size_properties_and_kdtree_oracle():
    """BASELINE config 3 (10k atoms, 4.2 A): equals the kd-tree oracle; symmetric, sorted, self loops."""
    from molecular_dynamics_neural_operator_b200 import ops
    from molecular_dynamics_neural_operator_b200 import synthetic as syn
    from oracle import gkn_oracle as orc

    wl = syn.WORKLOADS["synthetic_10k"]
    pos = syn.trajectory(wl, 2, seed=1)
    ei, ea = ops.radius_graph(torch.from_numpy(pos).cuda(), wl.cutoff)
    ei_np = ei.cpu().numpy()
    N = wl.num_atoms
    refs = [orc.radius_graph_kdtree(pos[f], wl.cutoff) for f in range(2)]
    ei_ref = np.concatenate([r[0] + f * N for f, r in enumerate(refs)], axis=1)
    ea_ref = np.concatenate([r[1] for r in refs], axis=0)
    assert np.array_equal(ei_np, ei_ref) and np.array_equal(ea.cpu().numpy(), ea_ref)
    key = ei_np[0] * (2 * N) + ei_np[1]
    assert np.all(np.diff(key) > 0)  # strictly increasing (row, col): sorted, no duplicates
    rev = np.sort(ei_np[1] * (2 * N) + ei_np[0])
    assert np.array_equal(rev, key)  # symmetric
    assert np.count_nonzero(ei_np[0] == ei_np[1]) == 2 * N  # every self loop present
    assert 250000 < ei_np.shape[1] // 2 < 340000  # ~291k edges / frame


def test_radius_graph_100k_atoms_matches_kdtree_oracle():
    """BASELINE config 5's frame (100k atoms, 100 A box, 4.2 A): bit-exact against the kd-tree
    oracle (the dense SciPy path would need an 80 GB distance matrix)."""
    from molecular_dynamics_neural_operator_b200 import ops
    from molecular_dynamics_neural_operator_b200 import synthetic as syn
    from oracle import gkn_oracle as orc

    wl = syn.WORKLOADS["synthetic_100k"]
    pos = syn.trajectory(wl, 1, seed=3)
    ei, ea = ops.radius_graph(torch.from_numpy(pos).cuda(), wl.cutoff)
    ei_ref, ea_ref = orc.radius_graph_kdtree(pos[0], wl.cutoff)
    assert 2_900_000 < ei_ref.shape[1] < 3_200_000  # ~3.06 M edges
    assert np.array_equal(ei.cpu().numpy(), ei_ref)
    assert np.array_equal(ea.cpu().numpy(), ea_ref)


def test_prepare_graph_against_numpy_stable_sort():
    from molecular_dynamics_neural_operator_b200 import ops

    g = torch.Generator().manual_seed(0)
    for N, E in [(1, 0), (7, 1), (50, 700), (504, 57000), (4000, 100000)]:
        ei = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)])
        plan = ops.prepare_graph(ei.cuda(), N)
        dst = ei[1].numpy()
        order = np.argsort(dst, kind="stable")
        assert np.array_equal(plan.perm.cpu().numpy(), order)
        assert np.array_equal(plan.src.cpu().numpy(), ei[0].numpy()[order])
        assert np.array_equal(plan.dst.cpu().numpy(), dst[order])
        cnt = np.bincount(dst, minlength=N)
        assert np.array_equal(plan.dst_ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(cnt)]))
        assert np.array_equal(plan.inv_cnt.cpu().numpy(), (1.0 / np.maximum(cnt, 1)).astype(np.float32))
        s_sorted = ei[0].numpy()[order]
        assert np.array_equal(plan.src_perm.cpu().numpy(), np.argsort(s_sorted, kind="stable"))


def test_prepare_graph_rejects_out_of_range_ids():
    from molecular_dynamics_neural_operator_b200 import ops

    ei = torch.tensor([[0, 1, 5], [1, 0, 1]], dtype=torch.long).cuda()
    with pytest.raises(ValueError, match="outside"):
        ops.prepare_graph(ei, 3)


def test_plan_from_csr_equals_general_prepare_on_radius_graphs():
    from molecular_dynamics_neural_operator_b200 import ops

    rng = np.random.default_rng(3)
    for F, N, L, thr in [(1, 1, 1.0, 2.0), (3, 200, 14.0, 4.2), (2, 504, 17.1, 8.0)]:
        pos = torch.from_numpy((rng.random((F, N, 3)) * L).astype(np.float32)).cuda()
        ei, ea, row_ptr = ops.radius_graph(pos, thr, return_row_ptr=True)
        fast = ops.plan_from_csr(ei, row_ptr, F * N)
        ref = ops.prepare_graph(ei, F * N)
        for name in ("perm", "src", "dst", "dst_ptr", "inv_cnt", "src_ptr", "src_perm"):
            assert torch.equal(getattr(fast, name), getattr(ref, name)), name
    bad = torch.tensor([[0, 1, 1], [1, 0, 2]], dtype=torch.long).cuda()  # 1->2 has no reverse
    with pytest.raises(ValueError, match="symmetric"):
        ops.plan_from_csr(bad, torch.tensor([0, 1, 3, 3], dtype=torch.long).cuda(), 3)


def test_radius_graph_survives_nan_inf_and_huge_coordinates():
    """A diverged rollout can feed garbage positions: the build must not crash and must still
    agree with SciPy (NaN / Inf distances fail the strict comparison, so those atoms are isolated
    - not even a self loop)."""
    import warnings

    from molecular_dynamics_neural_operator_b200 import ops
    from oracle import gkn_oracle as orc

    rng = np.random.default_rng(9)
    pos = (rng.random((2, 60, 3)) * 10).astype(np.float32)
    pos[0, 3] = np.nan
    pos[0, 7, 1] = np.inf
    pos[1, 5] = [1e30, -1e30, 1e30]
    pos[1, 6] = [3e38, 0, 0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ei_ref, ea_ref = orc.radius_graph_batched(pos, 4.2)
    ei, ea = ops.radius_graph(torch.from_numpy(pos).cuda(), 4.2)
    assert np.array_equal(ei.cpu().numpy(), ei_ref)
    assert np.array_equal(ea.cpu().numpy(), ea_ref, equal_nan=True)
    allnan = np.full((1, 10, 3), np.nan, dtype=np.float32)
    ei, ea = ops.radius_graph(torch.from_numpy(allnan).cuda(), 4.2)
    assert ei.shape == (2, 0)
