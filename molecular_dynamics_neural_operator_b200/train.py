"""Training step and frame-parallel data parallelism for the GKN model.

The reference trains through ``torch_geometric.nn.DataParallel`` (reference graph_kernel.py:528,
:454): one process, parameters re-broadcast every iteration, outputs gathered to GPU 0.  Here the
path shards the way it naturally does — trajectory frames are independent graphs — with one
process per GPU: every rank holds the parameters, runs its own frames and the only exchange is a
gradient all-reduce.  The loss is a SUM over frames (``LpLoss(size_average=False)``, :547), so
summing gradients across ranks reproduces the single-process gradient with no rescaling.

``GradReducer`` buckets that all-reduce and launches each bucket as soon as autograd has produced
all of its gradients (post-accumulate hooks), on a side stream, so communication overlaps the rest
of backward.  Buckets follow gradient-ready order: ``fc2`` first, then the conv stack's
``root/bias`` and the shared ``W3/b3`` (final when the stack's backward returns), then the edge
MLP's ``W2, b2, W1, b1`` and finally ``fc1/emb/lstm``.  Works with NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist

from .graph_kernel import LpLoss


class GradReducer:
    """Bucketed, backward-overlapped gradient all-reduce (sum) over flat contiguous buffers."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 8 << 20,
                 process_group=None):
        self.group = process_group
        self.params = [p for p in params if p.requires_grad]
        # reverse registration order ~ order in which autograd finishes gradients
        order = list(reversed(self.params))
        self.buckets: list[list[torch.nn.Parameter]] = []
        cur, cur_bytes = [], 0
        for p in order:
            nbytes = p.numel() * p.element_size()
            if cur and cur_bytes + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self.flat: list[torch.Tensor] = []
        self.bucket_of: dict[int, int] = {}
        self.offset_of: dict[int, int] = {}
        for b, ps in enumerate(self.buckets):
            n = sum(p.numel() for p in ps)
            flat = torch.zeros(n, dtype=ps[0].dtype, device=ps[0].device)
            off = 0
            for p in ps:
                # gradients live inside the bucket buffer (no pack/unpack copies)
                p.grad = flat[off:off + p.numel()].view_as(p)
                self.offset_of[id(p)] = off
                off += p.numel()
                self.bucket_of[id(p)] = b
            self.flat.append(flat)
        self.pending = [0] * len(self.buckets)
        self.works: list = []
        self.cuda = self.params[0].is_cuda if self.params else False
        self.comm_stream = torch.cuda.Stream() if self.cuda else None
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        self.reset()

    def reset(self):
        self.pending = [len(ps) for ps in self.buckets]
        self.works = []

    def zero_grad(self):
        """Zero the bucket buffers (use this, not optimizer.zero_grad(): the default
        set_to_none=True would detach the gradients from the buffers that are reduced; the hook
        re-attaches them if that happens anyway)."""
        for flat in self.flat:
            flat.zero_()
        for p in self.params:
            slot = self._slot(p)
            if p.grad is None or p.grad.data_ptr() != slot.data_ptr():
                p.grad = slot.view_as(p)
        self.reset()

    def _launch(self, b: int):
        if not self.enabled:
            return
        flat = self.flat[b]
        if self.cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                w = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            w = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.works.append(w)

    def _slot(self, p: torch.nn.Parameter) -> torch.Tensor:
        b, off = self.bucket_of[id(p)], self.offset_of[id(p)]
        return self.flat[b][off:off + p.numel()]

    def _hook(self, p: torch.nn.Parameter):
        b = self.bucket_of[id(p)]
        slot = self._slot(p)
        if p.grad is None or p.grad.data_ptr() != slot.data_ptr():
            # someone (optimizer.zero_grad(set_to_none=True), model.zero_grad()) replaced the view
            # into the bucket: move the gradient back so that what is reduced is what Adam reads
            if p.grad is not None:
                slot.copy_(p.grad.reshape(-1))
            p.grad = slot.view_as(p)
        if self.pending[b] <= 0:
            raise RuntimeError("GradReducer: a second backward() ran before finish(); call "
                               "Trainer.step / reducer.finish() once per backward")
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(b)

    def finish(self):
        """Call after backward: flush buckets whose hooks did not all fire (unused parameters)
        and wait for the reductions."""
        for b, left in enumerate(self.pending):
            if left > 0:
                self._launch(b)
        for w in self.works:
            w.wait()
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.reset()

    def close(self):
        for h in self.handles:
            h.remove()


class Trainer:
    """One optimisation step of the reference's training loop (graph_kernel.py:445-474):
    zero_grad -> model(batch) -> LpLoss.rel (sum over frames) -> backward -> Adam step,
    with the gradient all-reduce when torch.distributed is initialised."""

    def __init__(self, model: torch.nn.Module, lr: float = 0.01, weight_decay: float = 5e-4,
                 scheduler_step: int = 50, scheduler_gamma: float = 0.8,
                 bucket_bytes: int = 8 << 20, process_group=None):
        self.model = model
        # nn.Module.parameters() de-duplicates the shared edge MLP (conv2.net is conv1.net, :271-273)
        params = list(model.parameters())
        self.reducer = GradReducer(params, bucket_bytes, process_group)
        # reference: Adam(lr, weight_decay) + StepLR (graph_kernel.py:541-546)
        self.optimizer = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay,
                                          fused=params[0].is_cuda)
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=scheduler_step,
                                                         gamma=scheduler_gamma)
        self.loss_fn = LpLoss(size_average=False)
        self.last_mse: Optional[torch.Tensor] = None

    def step(self, batch, num_frames: int, with_mse: bool = False) -> torch.Tensor:
        self.reducer.zero_grad()
        out = self.model(batch)
        loss = self.loss_fn(out.view(num_frames, -1), batch.y.view(num_frames, -1))
        loss.backward()
        self.reducer.finish()
        self.optimizer.step()
        if with_mse:  # the reference logs it beside the loss (graph_kernel.py:465)
            self.last_mse = torch.nn.functional.mse_loss(out.detach(), batch.y)
        return loss.detach()

    def end_epoch(self) -> float:
        """`scheduler.step()` once per epoch (reference graph_kernel.py:622); returns the new lr."""
        self.scheduler.step()
        return float(self.scheduler.get_last_lr()[0])


def shard_frames(num_frames: int, rank: int, world: int) -> range:
    """Frames r, r+world, r+2*world, ... of a global batch go to rank r."""
    return range(rank, num_frames, world)


def shard_frames_by_edges(edge_counts, world: int) -> list[list[int]]:
    """Frames of one global batch dealt to ``world`` ranks so that the per-rank EDGE totals are as
    equal as a greedy longest-first assignment makes them (the cost of a frame is linear in its
    edge count; PyG's DataParallel balances by node count).  Returns the frame indices per rank,
    each list in increasing order.  Deterministic: ties go to the lower frame index / rank."""
    order = sorted(range(len(edge_counts)), key=lambda i: (-int(edge_counts[i]), i))
    loads = [0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], len(out[k]), k))
        out[r].append(i)
        loads[r] += int(edge_counts[i])
    return [sorted(v) for v in out]


def broadcast_parameters(model: torch.nn.Module, src: int = 0, process_group=None) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=src, group=process_group)
        for b in model.buffers():
            dist.broadcast(b.data, src=src, group=process_group)


# --------------------------------------------------------------------------------------------
# Epoch loops and checkpoints with the reference's layout (graph_kernel.py:445-493, :631-639)
# --------------------------------------------------------------------------------------------
def train_epoch(trainer: Trainer, batches, frames_per_batch: int, with_mse: bool = False):
    """`train` of the reference (graph_kernel.py:445-474) over an iterable of batched PairData:
    mean loss per batch (and mean MSE with ``with_mse``).  The losses stay on the device until the
    epoch ends, so the loop never waits for the GPU."""
    trainer.model.train()
    losses, mses = [], []
    for batch in batches:
        losses.append(trainer.step(batch, frames_per_batch, with_mse=with_mse))
        if with_mse:
            mses.append(trainer.last_mse)
    n = max(len(losses), 1)
    loss = float(torch.stack(losses).sum()) / n if losses else 0.0
    if with_mse:
        return loss, (float(torch.stack(mses).sum()) / n if mses else 0.0)
    return loss


@torch.no_grad()
def validate(model: torch.nn.Module, batches, frames_per_batch: int):
    """`validate` of the reference (graph_kernel.py:477-493): mean LpLoss and MSE."""
    model.eval()
    loss_fn = LpLoss(size_average=False)
    mse_fn = torch.nn.MSELoss()
    losses, mses = [], []
    for batch in batches:
        out = model(batch)
        losses.append(loss_fn(out.view(frames_per_batch, -1), batch.y.view(frames_per_batch, -1)))
        mses.append(mse_fn(out, batch.y))
    n = max(len(losses), 1)
    if not losses:
        return 0.0, 0.0
    return float(torch.stack(losses).sum()) / n, float(torch.stack(mses).sum()) / n


def save_checkpoint(path, epoch: int, trainer: Trainer, dataparallel_prefix: bool = True) -> None:
    """Same dictionary the reference writes (graph_kernel.py:631-639).  The reference saves a
    DataParallel-wrapped model, so its keys carry a ``module.`` prefix; keep that by default so
    the reference's notebook can load the file."""
    sd = trainer.model.state_dict()
    if dataparallel_prefix:
        sd = {"module." + k: v for k, v in sd.items()}
    torch.save({"epoch": epoch, "model_state_dict": sd,
                "optimizer_state_dict": trainer.optimizer.state_dict(),
                "scheduler_state_dict": trainer.scheduler.state_dict()}, path)


def load_checkpoint(path, model: torch.nn.Module, trainer: Optional[Trainer] = None,
                    map_location=None) -> int:
    """Load a checkpoint written by the reference or by save_checkpoint (with or without the
    ``module.`` prefix).  Returns the stored epoch."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = {(k[len("module."):] if k.startswith("module.") else k): v
          for k, v in ckpt["model_state_dict"].items()}
    model.load_state_dict(sd)
    if trainer is not None:
        if "optimizer_state_dict" in ckpt:
            trainer.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        if "scheduler_state_dict" in ckpt:
            trainer.scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    return int(ckpt.get("epoch", 0))


def fit(trainer: Trainer, train_batches, valid_batches, frames_per_batch: int, epochs: int,
        run_path=None, metrics_file: str = "metrics.jsonl", start_epoch: int = 0,
        best_loss: float = float("inf"), log=None) -> list[dict]:
    """The epoch loop of the reference's ``main`` (graph_kernel.py:583-639) without W&B, movies
    and plots: train, validate, one JSON line of metrics per epoch (in place of ``wandb.log``,
    :618-621), ``scheduler.step()`` (:622) and ``best.pt`` whenever the validation loss improves
    (:631-639, same dictionary, ``module.``-prefixed keys).  ``train_batches`` / ``valid_batches``
    are callables returning a fresh iterable per epoch, or re-iterable containers.  Under
    torch.distributed only rank 0 writes files.  Returns the list of per-epoch records."""
    import json
    import os
    from timeit import default_timer

    rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    if run_path is not None and rank0:
        os.makedirs(run_path, exist_ok=True)
    history = []
    for epoch in range(start_epoch, start_epoch + epochs):
        t0 = default_timer()
        tb = train_batches() if callable(train_batches) else train_batches
        vb = valid_batches() if callable(valid_batches) else valid_batches
        train_loss, train_mse = train_epoch(trainer, tb, frames_per_batch, with_mse=True)
        valid_loss, valid_mse = validate(trainer.model, vb, frames_per_batch)
        lr_used = float(trainer.optimizer.param_groups[0]["lr"])
        trainer.end_epoch()
        rec = {"epoch": epoch, "avg_train_loss": train_loss, "avg_valid_loss": valid_loss,
               "avg_train_mse": train_mse, "avg_valid_mse": valid_mse, "lr": lr_used,
               "time_s": default_timer() - t0, "best": valid_loss < best_loss}
        history.append(rec)
        if log is not None:
            log(rec)
        if run_path is not None and rank0:
            with open(os.path.join(run_path, metrics_file), "a") as f:
                f.write(json.dumps(rec) + "\n")
        if valid_loss < best_loss:
            best_loss = valid_loss
            if run_path is not None and rank0:
                save_checkpoint(os.path.join(run_path, "best.pt"), epoch, trainer)
    return history
