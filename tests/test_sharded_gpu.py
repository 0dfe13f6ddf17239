"""GPU: intra-frame destination-range sharding (sharded.py) gives the single-GPU result.
Emulated ranks on one GPU always; a real 2-process NCCL run when two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_case(N=300, W=64, K=128, depth=2, seed=0):
    from molecular_dynamics_neural_operator_b200 import ops
    from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN

    rng = np.random.default_rng(seed)
    pos = torch.from_numpy((rng.random((1, N, 3)) * 14).astype(np.float32)).cuda()
    ei, ea, row_ptr = ops.radius_graph(pos, 4.2, return_row_ptr=True)
    plan = ops.plan_from_csr(ei, row_ptr, N)
    torch.manual_seed(seed)
    model = KernelNN(W, K, depth, 6, in_width=7, out_width=3, window_size=1, num_residues=N,
                     precision="bf16").cuda().eval()
    x0 = torch.randn(N, W, device="cuda")
    return model, plan, ea.index_select(0, plan.perm_long), x0


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_emulated_ranks_match_the_unsharded_stack(world):
    from molecular_dynamics_neural_operator_b200 import functional as GF
    from molecular_dynamics_neural_operator_b200.sharded import emulate_world

    model, plan, ea, x0 = make_case()
    sched = (0,) * model.depth + (1,) * model.depth
    with torch.no_grad():
        ref = GF.gkn_stack(x0, ea, model.conv1.net, model.conv1.root, model.conv1.bias,
                           model.conv2.root, model.conv2.bias, plan, sched, relu=True, precision="bf16")
    out = emulate_world(x0, plan, ea, model.conv1.net, model.conv1.root, model.conv1.bias,
                        model.conv2.root, model.conv2.bias, sched, True, world, "bf16")
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6)


def _worker(rank, world, port, out):
    from types import SimpleNamespace

    import torch.distributed as dist

    from molecular_dynamics_neural_operator_b200 import ops
    from molecular_dynamics_neural_operator_b200.sharded import kernelnn_forward_sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        model, plan, ea_sorted, _ = make_case()
        model = model.to(f"cuda:{rank}")
        rng = np.random.default_rng(0)
        N = plan.num_nodes
        pos = torch.from_numpy((rng.random((1, N, 3)) * 14).astype(np.float32)).cuda()
        ei, ea = ops.radius_graph(pos, 4.2)
        data = SimpleNamespace(x_position=pos, x_aminoacid=torch.zeros(N, dtype=torch.long, device="cuda"),
                               edge_index=ei, edge_attr=ea)
        y = kernelnn_forward_sharded(model, data)
        with torch.no_grad():
            ref = model(data)
        out[rank] = float((y - ref).abs().max() / ref.abs().max())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_process_nccl_run_matches_single_gpu():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] <= 1e-6 and out[1] <= 1e-6
