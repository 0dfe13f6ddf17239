"""Shared helpers for the test-suite (test infrastructure; may import the oracle)."""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def state_dict_of(gold: dict, dtype=torch.float64) -> dict:
    return {k[3:]: torch.from_numpy(v).to(dtype) for k, v in gold.items() if k.startswith("sd.")}


def grads_of(gold: dict) -> dict:
    return {k[5:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("grad.")}


def kernelnn_batch(gold: dict, dtype=torch.float64, device="cpu"):
    """Rebuild the PairData-like batch a kernelnn_* golden file was generated from."""
    B, W, R = int(gold["batch"]), int(gold["window"]), int(gold["num_residues"])
    traj = torch.from_numpy(gold["traj"])
    xp = torch.cat([traj[b:b + W] for b in range(B)]).to(dtype)
    y = torch.cat([traj[b + W] for b in range(B)]).to(dtype)
    aa = torch.from_numpy(gold["aminoacid"]).repeat(B)
    return SimpleNamespace(
        x_aminoacid=aa.to(device), x_position=xp.to(device), y=y.to(device),
        edge_attr=torch.from_numpy(gold["edge_attr"]).to(dtype).to(device),
        edge_index=torch.from_numpy(gold["edge_index"]).to(device),
    )


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (the relative measure the parity tolerances are stated in)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def seeded_state_dict(template: dict, seed: int) -> dict:
    """float32 weights drawn from numpy's PCG64 stream (stable across versions and machines) for
    every tensor of ``template`` (name -> tensor, only the shapes are used), names in sorted
    order, U(-1/sqrt(fan_in), 1/sqrt(fan_in)).  ``conv2.net.*`` aliases ``conv1.net.*``
    (reference graph_kernel.py:271-273).  Lets a large model be pinned by a golden file that holds
    only outputs and gradients: generator and test rebuild the same weights from the seed."""
    rng = np.random.default_rng(seed)
    out = {}
    for k in sorted(template):
        if k.startswith("conv2.net."):
            continue
        shape = tuple(template[k].shape)
        fan_in = shape[-1] if len(shape) > 1 else shape[0]
        b = 1.0 / np.sqrt(max(fan_in, 1))
        out[k] = torch.from_numpy(rng.uniform(-b, b, size=shape).astype(np.float32))
    for k in template:
        if k.startswith("conv2.net."):
            out[k] = out["conv1.net." + k[len("conv2.net."):]]
    return out
