"""``PairData`` boundary type and batching, mirroring reference dataset.py:21-53.

The reference's ``PairData`` is a ``torch_geometric.data.Data`` subclass whose only behaviour on
this path is (a) carrying five tensors, (b) ``num_nodes = x_aminoacid.size(0)`` and (c) offsetting
``edge_index`` by ``num_nodes`` when graphs are batched (``__inc__``, dataset.py:41-45).  PyG is not
a dependency here; this stand-in exposes the same attributes and the same batching rule.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch


class PairData:
    """Same five attributes as reference dataset.py:21-35."""

    _FIELDS = ("x_aminoacid", "x_position", "y", "edge_attr", "edge_index")

    def __init__(
        self,
        x_aminoacid: Optional[torch.Tensor] = None,
        x_position: Optional[torch.Tensor] = None,
        y: Optional[torch.Tensor] = None,
        edge_attr: Optional[torch.Tensor] = None,
        edge_index: Optional[torch.Tensor] = None,
    ) -> None:
        self.x_aminoacid = x_aminoacid
        self.x_position = x_position
        self.y = y
        self.edge_attr = edge_attr
        self.edge_index = edge_index

    @property
    def num_nodes(self) -> int:  # dataset.py:37-39
        return self.x_aminoacid.size(0)

    def __inc__(self, key, value=None, *args, **kwargs):  # dataset.py:41-45
        return self.num_nodes if key == "edge_index" else 0

    def to(self, device, non_blocking: bool = False) -> "PairData":
        kw = {f: (getattr(self, f).to(device, non_blocking=non_blocking)
                  if getattr(self, f) is not None else None) for f in self._FIELDS}
        return PairData(**kw)

    def pin_memory(self) -> "PairData":  # dataset.py:47-53
        for f in self._FIELDS:
            t = getattr(self, f)
            if t is not None:
                setattr(self, f, t.pin_memory())
        return self

    def __repr__(self) -> str:
        parts = [f"{f}={list(getattr(self, f).shape)}" for f in self._FIELDS
                 if getattr(self, f) is not None]
        return f"PairData({', '.join(parts)})"


def collate(batch: Sequence[PairData]) -> PairData:
    """``Batch.from_data_list`` semantics for PairData: tensors concatenated along dim 0
    (``edge_index`` along dim 1, each graph's node ids shifted by the nodes before it)."""
    offs, n = [], 0
    for d in batch:
        offs.append(n)
        n += d.num_nodes

    def cat(name, dim=0):
        ts = [getattr(d, name) for d in batch]
        return None if any(t is None for t in ts) else torch.cat(ts, dim=dim)

    edge_index = torch.cat([d.edge_index + o for d, o in zip(batch, offs)], dim=1)
    return PairData(x_aminoacid=cat("x_aminoacid"), x_position=cat("x_position"), y=cat("y"),
                    edge_attr=cat("edge_attr"), edge_index=edge_index)
