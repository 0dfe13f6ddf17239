"""CPU: the training-driver pieces that mirror the reference's `main` loop
(reference graph_kernel.py:541-547, :583-639): StepLR stepped once per epoch, best-validation-loss
checkpoint in the reference's dictionary layout, one JSON record per epoch; and the robustness of
the bucketed gradient reducer against `zero_grad(set_to_none=True)`.  A small CPU module stands in
for the model (the trainer never looks inside it)."""
import json

import pytest
import torch

from molecular_dynamics_neural_operator_b200.train import (Trainer, fit, load_checkpoint,
                                                            shard_frames_by_edges, train_epoch,
                                                            validate)


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(3, 8)
        self.b = torch.nn.Linear(8, 3)

    def forward(self, batch):
        return self.b(torch.relu(self.a(batch.x)))


class Batch:
    def __init__(self, x, y):
        self.x, self.y = x, y


def batches(n=3, frames=4, nodes=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [Batch(torch.randn(frames * nodes, 3, generator=g), torch.randn(frames * nodes, 3, generator=g))
            for _ in range(n)]


def test_scheduler_steps_once_per_epoch_and_decays():
    torch.manual_seed(0)
    tr = Trainer(Tiny(), lr=0.01, scheduler_step=2, scheduler_gamma=0.5)
    lrs = []
    for _ in range(5):
        train_epoch(tr, batches(), 4)
        lrs.append(tr.end_epoch())
    # StepLR(step_size=2, gamma=0.5) as at reference graph_kernel.py:544-546, stepped at :622
    assert lrs == pytest.approx([0.01, 0.005, 0.005, 0.0025, 0.0025])
    assert tr.scheduler.state_dict()["last_epoch"] == 5


def test_fit_writes_metrics_and_best_checkpoint(tmp_path):
    torch.manual_seed(0)
    model = Tiny()
    tr = Trainer(model, lr=0.01, scheduler_step=1, scheduler_gamma=0.9)
    tb, vb = batches(4, seed=1), batches(2, seed=2)
    hist = fit(tr, lambda: tb, lambda: vb, 4, epochs=4, run_path=tmp_path)
    lines = [json.loads(l) for l in (tmp_path / "metrics.jsonl").read_text().splitlines()]
    assert len(lines) == 4 and [r["epoch"] for r in lines] == [0, 1, 2, 3]
    assert set(lines[0]) >= {"avg_train_loss", "avg_valid_loss", "avg_train_mse", "avg_valid_mse", "lr"}
    assert lines[0]["lr"] == pytest.approx(0.01) and lines[3]["lr"] == pytest.approx(0.01 * 0.9 ** 3)
    best_epoch = min(range(4), key=lambda e: (hist[e]["avg_valid_loss"], e))
    ckpt = torch.load(tmp_path / "best.pt", weights_only=False)
    # the reference's dictionary (graph_kernel.py:633-638), DataParallel-prefixed keys
    assert set(ckpt) == {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict"}
    assert ckpt["epoch"] == best_epoch
    assert all(k.startswith("module.") for k in ckpt["model_state_dict"])
    assert ckpt["scheduler_state_dict"]["last_epoch"] == best_epoch + 1
    # best flag: only strict improvements
    running = float("inf")
    for r in hist:
        assert r["best"] == (r["avg_valid_loss"] < running)
        running = min(running, r["avg_valid_loss"])
    # round trip into a fresh trainer resumes the schedule
    m2 = Tiny()
    t2 = Trainer(m2, lr=0.01, scheduler_step=1, scheduler_gamma=0.9)
    assert load_checkpoint(tmp_path / "best.pt", m2, t2) == best_epoch
    assert t2.scheduler.state_dict()["last_epoch"] == best_epoch + 1
    loss_a, _ = validate(m2, vb, 4)
    assert loss_a == pytest.approx(hist[best_epoch]["avg_valid_loss"], rel=1e-6)


def test_reducer_survives_set_to_none_zero_grad():
    torch.manual_seed(0)
    model = Tiny()
    tr = Trainer(model, lr=0.0, weight_decay=0.0)
    b = batches(1)[0]
    tr.step(b, 4)
    ref = [p.grad.clone() for p in model.parameters()]
    # a caller zeroes gradients the PyTorch-default way: the views into the buckets are dropped
    tr.optimizer.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in model.parameters())
    out = model(b)
    tr.loss_fn(out.view(4, -1), b.y.view(4, -1)).backward()
    tr.reducer.finish()
    for p, g in zip(model.parameters(), ref):
        slot = tr.reducer._slot(p)
        assert p.grad.data_ptr() == slot.data_ptr()  # re-attached to the reduced buffer
        assert torch.allclose(p.grad, g)


def test_reducer_rejects_second_backward_without_finish():
    torch.manual_seed(0)
    model = Tiny()
    tr = Trainer(model)
    b = batches(1)[0]
    tr.reducer.zero_grad()
    tr.loss_fn(model(b).view(4, -1), b.y.view(4, -1)).backward()
    with pytest.raises(RuntimeError, match="second backward"):
        tr.loss_fn(model(b).view(4, -1), b.y.view(4, -1)).backward()


def test_frames_are_dealt_by_edge_count():
    counts = [60000, 100, 59000, 61000, 200, 58000, 300, 60500]
    parts = shard_frames_by_edges(counts, 4)
    assert sorted(i for p in parts for i in p) == list(range(8))      # every frame exactly once
    loads = [sum(counts[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(counts)                     # longest-first bound
    # round-robin (frames r, r + world, ...) would put 60000 + 200 next to 59000 + 300 ... and
    # 61000 + 60500 on one rank; by edges the two extra large frames are the only imbalance left
    even = [57000, 61000, 59500, 60500, 58000, 62000, 60000, 59000]
    loads = [sum(even[i] for i in p) for p in shard_frames_by_edges(even, 4)]
    assert max(loads) - min(loads) <= 1000 and all(len(p) == 2 for p in shard_frames_by_edges(even, 4))
    assert shard_frames_by_edges(counts, 4) == parts                  # deterministic
    assert shard_frames_by_edges([5, 5, 5, 5], 2) == [[0, 2], [1, 3]]
    assert shard_frames_by_edges([], 3) == [[], [], []]
