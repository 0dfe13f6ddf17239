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
