"""CPU, world_size 2, gloo: (1) the frame-sharded data-parallel path (GradReducer / Trainer)
produces the single-process gradient and parameter update; (2) the source-range sharded rollout of
the node-first path (sharded.py: partial node sums + all-reduce) gives the unsharded result, with
the CUDA entry points replaced by their torch restatements (tests/test_node_path_cpu.py).  The GKN kernels themselves need a GPU, so a
small CPU module with the same structural quirk (one sub-network shared by two layers, reference
graph_kernel.py:271-273) stands in for the model; the reducer never looks inside the model."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from molecular_dynamics_neural_operator_b200.train import (GradReducer, Trainer, broadcast_parameters,
                                                            shard_frames)


class SharedNetModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = torch.nn.Linear(3, 8)
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 8))
        self.conv1 = torch.nn.ModuleDict({"net": net})
        self.conv2 = torch.nn.ModuleDict({"net": net})  # aliased parameters
        self.big = torch.nn.Linear(8, 600)  # forces several buckets with a small bucket size
        self.fc2 = torch.nn.Linear(600, 3)

    def forward(self, batch):
        x = torch.relu(self.fc1(batch.x))
        x = torch.relu(self.conv1["net"](x))
        x = torch.relu(self.conv2["net"](x))
        return self.fc2(torch.relu(self.big(x)))


class Batch:
    def __init__(self, x, y):
        self.x, self.y = x, y


def make_data(frames=6, nodes=5):
    g = torch.Generator().manual_seed(0)
    return torch.randn(frames, nodes, 3, generator=g), torch.randn(frames, nodes, 3, generator=g)


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
        model = SharedNetModel()
        broadcast_parameters(model)
        trainer = Trainer(model, lr=0.01, weight_decay=5e-4, bucket_bytes=4096)
        assert len(trainer.reducer.buckets) >= 3
        x, y = make_data()
        idx = list(shard_frames(x.size(0), rank, world))
        batch = Batch(x[idx].reshape(-1, 3), y[idx].reshape(-1, 3))
        loss = trainer.step(batch, len(idx))
        grads = {k: p.grad.clone() for k, p in model.named_parameters()}
        params = {k: p.detach().clone() for k, p in model.named_parameters()}
        out[rank] = (float(loss), grads, params)
    finally:
        dist.destroy_process_group()


def test_shard_frames_partitions_the_batch():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_frames(64, r, world))
        assert seen == list(range(64))
        sizes = {len(shard_frames(64, r, world)) for r in range(world)}
        assert sizes == {64 // world}


@pytest.mark.timeout(120)
def test_two_rank_gradients_and_update_match_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(world, free_port(), out), nprocs=world, join=True)
    # single-process reference on the full batch, starting from rank 0's init
    torch.manual_seed(100)
    model = SharedNetModel()
    trainer = Trainer(model, lr=0.01, weight_decay=5e-4, bucket_bytes=4096)
    x, y = make_data()
    loss = trainer.step(Batch(x.reshape(-1, 3), y.reshape(-1, 3)), x.size(0))
    ref_grads = {k: p.grad for k, p in model.named_parameters()}
    ref_params = {k: p.detach() for k, p in model.named_parameters()}
    assert abs(out[0][0] + out[1][0] - float(loss)) < 1e-5  # loss is a sum over frames
    for r in range(world):
        _, grads, params = out[r]
        for k in ref_grads:
            assert torch.allclose(grads[k], ref_grads[k], rtol=1e-5, atol=1e-6), (r, k)
            assert torch.allclose(params[k], ref_params[k], rtol=1e-5, atol=1e-6), (r, k)


def test_reducer_is_a_noop_without_process_group():
    model = SharedNetModel()
    red = GradReducer(model.parameters(), bucket_bytes=1024)
    assert not red.enabled
    x, y = make_data(2)
    red.zero_grad()
    ((model(Batch(x.reshape(-1, 3), y)) - y.reshape(-1, 3)) ** 2).sum().backward()
    red.finish()
    # gradients live inside the flat bucket buffers (no pack/unpack copies)
    for p in model.parameters():
        b = red.bucket_of[id(p)]
        lo = red.flat[b].data_ptr()
        hi = lo + red.flat[b].numel() * 4
        assert lo <= p.grad.data_ptr() < hi
    red.close()


# ------------------------------------------------------------------------------------------------
# source-range sharding of the node-first rollout (sharded.sharded_stack_forward_node)
# ------------------------------------------------------------------------------------------------
def _rollout_case():
    from oracle import gkn_oracle as orc
    from tests.test_node_path_cpu import random_graph

    N, E, W, K = 11, 400, 64, 256
    ei = random_graph(N, E, seed=5, isolated=(4,))
    g = torch.Generator().manual_seed(6)
    net_p = orc.make_net_params(6, K, W, seed=7, dtype=torch.float32)
    net = torch.nn.Module()
    net.layers = torch.nn.ModuleList([torch.nn.Linear(6, K), torch.nn.ReLU(), torch.nn.Linear(K, K),
                                      torch.nn.ReLU(), torch.nn.Linear(K, W * W)])
    with torch.no_grad():
        for i in (0, 2, 4):
            net.layers[i].weight.copy_(net_p[f"layers.{i}.weight"])
            net.layers[i].bias.copy_(net_p[f"layers.{i}.bias"])
    x0 = torch.randn(N, W, generator=g)
    ea = torch.randn(E, 6, generator=g)
    roots = [torch.randn(W, W, generator=g) / 8 for _ in range(2)]
    biases = [torch.randn(W, generator=g) / 8 for _ in range(2)]
    return ei, ea, x0, net, roots, biases


def rollout_worker(rank, world, port, out):
    from molecular_dynamics_neural_operator_b200 import sharded
    from tests.test_node_path_cpu import install_emulation

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        install_emulation(setattr)
        ei, ea, x0, net, roots, biases = _rollout_case()
        shard = sharded.shard_frame_by_source(ei, ea, x0.size(0), rank, world)
        y = sharded.sharded_stack_forward_node(x0, shard, net, roots[0], biases[0], roots[1],
                                               biases[1], (0, 1, 1), relu=True)
        out[rank] = y
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_source_sharded_rollout_matches_unsharded_stack(monkeypatch):
    from molecular_dynamics_neural_operator_b200 import functional as GF
    from tests.test_node_path_cpu import cpu_plan, install_emulation

    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(rollout_worker, args=(world, free_port(), out), nprocs=world, join=True)
    install_emulation(monkeypatch.setattr)
    ei, ea, x0, net, roots, biases = _rollout_case()
    plan = cpu_plan(ei, x0.size(0))
    with torch.no_grad():
        ref = GF.gkn_stack(x0, ea.index_select(0, plan.perm.long()), net, roots[0], biases[0],
                           roots[1], biases[1], plan, (0, 1, 1), relu=True, precision="bf16")
    assert torch.equal(out[0], out[1])  # every rank returns the full, identical result
    err = float((out[0] - ref).abs().max() / ref.abs().max())
    assert err <= 2.0 ** -9  # fp32 re-association, amplified at most to a few bf16 ulps
