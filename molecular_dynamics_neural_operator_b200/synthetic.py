"""Seeded synthetic MD trajectories of the shapes BASELINE.json names, and the
device-side feeder that turns a window of coordinates into the ``PairData`` batch the model
consumes (the role of ContactMapDataset.__getitem__ + DataListLoader + Batch, reference
dataset.py:180-227, with the per-frame radius graph built on the GPU instead of read from HDF5).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import ops
from .dataset import PairData


@dataclass(frozen=True)
class Workload:
    name: str
    num_atoms: int
    width: int
    ker_width: int
    depth: int
    cutoff: float
    window: int
    box: float  # cube edge (uniform workloads) or 0 for the chain


# the configs of BASELINE.json / SURVEY.md §8
WORKLOADS = {
    "bba_ca": Workload("bba_ca", 28, 32, 1024, 4, 8.0, 10, 0.0),
    "bba_allatom": Workload("bba_allatom", 504, 64, 1024, 6, 8.0, 10, 17.1),
    "synthetic_10k": Workload("synthetic_10k", 10000, 64, 1024, 6, 4.2, 10, 46.4),
    "synthetic_100k": Workload("synthetic_100k", 100000, 128, 1024, 8, 4.2, 10, 100.0),
}


def chain_positions(rng: np.random.Generator, n: int, bond: float = 3.8) -> np.ndarray:
    """Compact persistent random walk standing in for a C-alpha trace (bond 3.8 A)."""
    pos = np.zeros((n, 3))
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    for i in range(1, n):
        # new direction: previous direction bent by ~70 degrees, biased back toward the centroid
        r = rng.normal(size=3)
        r -= r.dot(d) * d
        r /= np.linalg.norm(r)
        ang = np.deg2rad(rng.uniform(55, 85))
        nd = np.cos(ang) * d + np.sin(ang) * r
        back = pos[:i].mean(0) - pos[i - 1]
        nb = np.linalg.norm(back)
        if nb > 1e-9:
            nd = nd + 0.35 * back / nb
        d = nd / np.linalg.norm(nd)
        pos[i] = pos[i - 1] + bond * d
    return pos


def trajectory(workload: Workload, num_frames: int, seed: int = 0, jitter: float = 0.2) -> np.ndarray:
    """float32 [num_frames, N, 3]: a fixed base configuration plus per-frame Gaussian jitter
    (MD-like temporal coherence), ``numpy.random.default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    n = workload.num_atoms
    if workload.box > 0:
        base = rng.random((n, 3)) * workload.box
    else:
        base = chain_positions(rng, n)
    frames = base[None] + jitter * rng.normal(size=(num_frames, n, 3))
    return frames.astype(np.float32)


def residue_ids(workload: Workload, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed + 1)
    return rng.integers(0, 20, size=workload.num_atoms).astype(np.int64)


def window_batch(traj: np.ndarray, starts, window: int, horizon: int = 1):
    """Host-side slicing only (what the reference Dataset does, dataset.py:180-204):
    returns x_position [B, W, N, 3] and y [B, N, 3] as float32 numpy arrays."""
    xs = np.stack([traj[s:s + window] for s in starts])
    ys = np.stack([traj[s + window + horizon - 1] for s in starts])
    return xs, ys


def make_batch(x_window: torch.Tensor, aminoacid: torch.Tensor, y: torch.Tensor | None,
               cutoff: float, graph_frame: int = 0, prepare: bool = False,
               ker_width: int | None = None) -> PairData:
    """Device-side feeder.  x_window [B, W, N, 3] (device), aminoacid [N] int64, y [B, N, 3].

    Graph and edge attributes come from window frame ``graph_frame`` of every sample: 0 (the
    first frame) at training time as in dataset.py:189-201, -1 (the last) in rollout as in
    graph_kernel.py:363-379.  Equivalent to collate([PairData(...) for each sample]).
    ``prepare=True`` also builds the destination-sorted CSR plan and attaches it to the batch
    (``KernelNN.forward`` picks it up), so both happen off the training stream when prefetched;
    ``ker_width`` (the model's kernel width) lets the unit table of the dh pass be prebuilt too."""
    B, W, N, _ = x_window.shape
    frames = x_window[:, graph_frame].contiguous()
    edge_index, edge_attr, row_ptr = ops.radius_graph(frames, cutoff, return_row_ptr=True)
    data = PairData(
        x_aminoacid=aminoacid.repeat(B),
        x_position=x_window.reshape(B * W, N, 3),
        y=None if y is None else y.reshape(B * N, 3),
        edge_attr=edge_attr,
        edge_index=edge_index,
    )
    if prepare:
        # radius graphs are symmetric and (row, col)-sorted: no sorting needed for the CSR plan
        data._gkn_plan = ops.plan_from_csr(edge_index, row_ptr, B * N, validate=False)
        from . import functional as GF

        if GF.FACTORIZATION == "node" and edge_index.size(1) > 0:
            # padded source-sorted layout of the node-first kernels (two small host reads): built
            # here so that it, too, happens off the training stream when prefetched
            ops.node_layout(data._gkn_plan, dh_chunks=ker_width // 256 if ker_width else None)
    return data


class BatchPrefetcher:
    """Builds the NEXT step's batch on a side CUDA stream while the current step computes.

    The radius-graph build has a data-dependent output size (one host read of the edge count);
    done on the training stream that read would drain the whole queue every step.  Here the
    optional H2D copy, graph build and CSR prepare run on their own stream and the training
    stream only waits on an event."""

    def __init__(self, device):
        self.device = device
        # high priority: the build is a chain of small kernels with host reads in between (edge
        # count, layout sizes); they must not queue behind the training stream's long kernels
        self.stream = torch.cuda.Stream(device=device, priority=-1)
        self._pending = None
        self._event = None

    def submit(self, build_fn) -> None:
        main = torch.cuda.current_stream(self.device)
        # inputs of build_fn may have been produced on the training stream (a rollout output fed
        # back in, a trajectory tensor just created): order the side stream after it
        self.stream.wait_stream(main)
        with torch.cuda.stream(self.stream):
            batch = build_fn()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        # tensors were allocated on the side stream but will be consumed on the training stream
        tensors = [getattr(batch, f) for f in PairData._FIELDS]
        plan = getattr(batch, "_gkn_plan", None)
        if plan is not None:
            tensors += [plan.perm, plan.src, plan.dst, plan.dst_ptr, plan.inv_cnt, plan.src_ptr,
                        plan.src_perm]
            lay = getattr(plan, "_node_layout", None)
            if lay is not None:
                tensors += lay.tensors() + [lay.qpos_long]
        for t in tensors:
            if t is not None and t.is_cuda:
                t.record_stream(main)
        self._pending, self._event = batch, ev

    def get(self) -> PairData:
        if self._pending is None:
            raise RuntimeError("BatchPrefetcher.get() called before submit()")
        torch.cuda.current_stream(self.device).wait_event(self._event)
        batch, self._pending = self._pending, None
        return batch
