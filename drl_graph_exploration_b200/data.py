"""Minimal stand-ins for ``torch_geometric.data.Data`` / ``DataLoader`` (PyG is not a dependency):
exactly the surface ``policy.py`` / ``test.py`` use -- ``Data(x=, edge_index=, edge_attr=)``,
``.to(device)``, ``DataLoader(list, batch_size=)`` yielding a concatenated batch with ``.batch``."""
from __future__ import annotations

import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, batch=None):
        self.x, self.edge_index, self.edge_attr, self.batch = x, edge_index, edge_attr, batch

    @property
    def num_nodes(self):
        return self.x.size(0)

    def to(self, device):
        """Like PyG's ``Data.to``; data that already lives on ``device`` is returned as it is (with the CSR / normalisation a
        forward pass cached on it, ``_dge_structure``), so the target-net forward, the policy forward and its backward of one
        gradient step share one preprocessing."""
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        ts = (self.x, self.edge_index, self.edge_attr, self.batch)
        if all(t is None or (t.device.type == dev.type and (dev.index is None or t.device.index == dev.index)) for t in ts):
            return self
        out = type(self)(*(None if t is None else t.to(device) for t in ts))
        if hasattr(self, "num_graphs"):
            out.num_graphs = self.num_graphs
        return out

    def __repr__(self):
        return f"Data(x={tuple(self.x.shape)}, edge_index={tuple(self.edge_index.shape)}, edge_attr={tuple(self.edge_attr.shape)})"


class Batch(Data):
    @staticmethod
    def from_data_list(items):
        xs, eis, eas, bs = [], [], [], []
        off = 0
        for g, d in enumerate(items):
            xs.append(d.x); eis.append(d.edge_index + off); eas.append(d.edge_attr)
            bs.append(torch.full((d.x.size(0),), g, dtype=torch.long, device=d.x.device))
            off += d.x.size(0)
        b = Batch(torch.cat(xs), torch.cat(eis, dim=1), torch.cat(eas), torch.cat(bs))
        b.num_graphs = len(items)
        return b


class DataLoader:
    def __init__(self, dataset, batch_size=1, shuffle=False):
        self.dataset, self.batch_size = list(dataset), batch_size
        assert not shuffle

    def __iter__(self):
        for i in range(0, len(self.dataset), self.batch_size):
            yield Batch.from_data_list(self.dataset[i:i + self.batch_size])

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size
