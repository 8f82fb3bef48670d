"""Drop-in for the hot-path part of the reference's utils.py: calculate_f1 (utils.py:163-169),
consistency_loss (utils.py:187-211), fix_seeds (82-89) and GpuMemoryProfiler (13-80).
Plotting helpers are out of scope."""
from __future__ import annotations

import random

import numpy as np
import torch

from . import ops


def fix_seeds(seed=42):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    ops.reset_seed_counter()


def calculate_f1(logits, labels, mask):
    """Micro-F1 of argmax(logits[mask]) -- equal to accuracy for single-label classification, so it
    is counted on the device (one 64-byte D2H) instead of sklearn on the host."""
    acc = ops.loss_forward(logits.detach(), labels, mask.view(torch.uint8)).cpu()
    n = float(acc[1])
    return float(acc[2]) / n if n > 0 else 0.0


def consistency_loss(edge_probs, edge_indices, node_embeddings):
    """MSE(edge_probs, cos(emb[src], emb[dst])); gradients flow to both arguments."""
    n = node_embeddings.size(0)
    g = ops.graph_of(edge_indices, n)
    dev = node_embeddings.device
    y = torch.zeros(n, dtype=torch.int64, device=dev)
    tm = torch.zeros(n, dtype=torch.uint8, device=dev)
    return ops.fused_loss(node_embeddings, y, tm, edge_probs, g, c1=0.0, c2=1.0, reg1=False, reg2=True, c0=0.0)


class GpuMemoryProfiler:
    """Per-segment peak-memory deltas with the reference's segment names (edge_mlp_pre,
    edge_score, gnn_forward, backward)."""

    def __init__(self, enabled=False, device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self.enabled = bool(enabled) and torch.cuda.is_available() and self.device.type == "cuda"
        self._epoch = None
        self._stats = {}
        self._open = {}

    def start_epoch(self, epoch):
        if self.enabled:
            self._epoch = epoch
            self._stats.setdefault(epoch, {})

    def begin(self, name):
        if not self.enabled or self._epoch is None:
            return
        torch.cuda.synchronize(self.device)
        self._open[name] = (torch.cuda.max_memory_allocated(self.device), torch.cuda.memory_allocated(self.device))

    def end(self, name):
        if not self.enabled or self._epoch is None:
            return 0, 0
        start = self._open.pop(name, None)
        if start is None:
            return 0, 0
        torch.cuda.synchronize(self.device)
        peak = torch.cuda.max_memory_allocated(self.device)
        alloc = torch.cuda.memory_allocated(self.device)
        row = (max(0, peak - start[0]), alloc, alloc - start[1])
        self._stats.setdefault(self._epoch, {}).setdefault(name, []).append(row)
        return row[0], alloc

    def summarize_epoch(self, epoch):
        if not self.enabled:
            return {}
        mb = 1024 ** 2
        out = {}
        for name, rows in self._stats.get(epoch, {}).items():
            if not rows:
                continue
            pk, al, inc = zip(*rows)
            out[name] = {
                "max_peak_inc_bytes": max(pk), "max_peak_inc_mb": max(pk) / mb,
                "mean_peak_inc_mb": sum(pk) / len(pk) / mb,
                "max_alloc_after_bytes": max(al), "max_alloc_after_mb": max(al) / mb,
                "mean_alloc_after_mb": sum(al) / len(al) / mb,
                "max_alloc_inc_bytes": max(inc), "max_alloc_inc_mb": max(inc) / mb,
                "mean_alloc_inc_mb": sum(inc) / len(inc) / mb, "calls": len(rows),
            }
        return out

    def end_epoch(self):
        if self.enabled:
            self._epoch = None
            self._open.clear()
