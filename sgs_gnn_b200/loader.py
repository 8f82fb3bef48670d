"""Host -> device batch pipeline for the training loops (training_hybrid.py:30,42 `batch = batch.to(device)`).

The reference uploads every cluster batch synchronously at the top of its step.  `prefetch(loader, device)` keeps
that contract -- every step's inputs cross PCIe from (pinned) host memory, every step -- but issues the copy of
batch k+1 on a separate CUDA stream while step k computes, into freshly allocated device tensors (double buffering by
construction: batch k stays alive until its step drops it).  Batches that already live on the device pass through.

Host batches may carry `edge_index` as int32 [2, E] (`Batch.compact()`): node ids fit 31 bits, so the int64 form of
the reference only doubles the bytes on the host link; ops.Graph takes the int32 rows as they are.  A source-sorted edge
list additionally sends its source row as the CSR row pointer (N + 1 ints), rebuilt on the device during the upload.
"""
from __future__ import annotations

import os

import torch

_copy_streams = {}


def _copy_stream(dev):
    key = (dev.type, dev.index)
    s = _copy_streams.get(key)
    if s is None:
        s = _copy_streams[key] = torch.cuda.Stream(device=dev)
    return s


def _async_capable(batch, dev):
    return (dev.type == "cuda" and hasattr(batch, "upload_async") and getattr(batch, "x", None) is not None
            and batch.x.device.type == "cpu" and not os.environ.get("SGS_NO_PREFETCH"))


def prefetch(loader, device):
    """Generator over `loader` whose host batches arrive on `device`; the upload of the next batch overlaps the
    consumer's work on the current one.  Yields (batch_on_device_or_as_is)."""
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    it = iter(loader)

    def start(b):
        if not _async_capable(b, dev):
            return b, None
        main = torch.cuda.current_stream(dev)
        cs = _copy_stream(dev)
        cs.wait_stream(main)          # the destination blocks may have been freed by work still queued on `main`
        out = b.upload_async(dev, cs)
        ev = torch.cuda.Event()
        ev.record(cs)
        return out, ev

    try:
        nxt = start(next(it))
    except StopIteration:
        return
    while nxt is not None:
        cur, ev = nxt
        if ev is not None:
            torch.cuda.current_stream(dev).wait_event(ev)
            if hasattr(cur, "finish_upload"):
                cur = cur.finish_upload()      # e.g. the NVLink all-gather of a row-sliced replicated tensor
        try:
            nb = next(it)
        except StopIteration:
            nb = None
        nxt = start(nb) if nb is not None else None
        yield cur


def copy_fields_async(fields, dev, stream):
    """{name: host tensor or None} -> {name: device tensor}: destinations allocated on the current stream, copies
    enqueued on `stream` (non-blocking from pinned memory)."""
    out = {}
    for k, v in fields.items():
        out[k] = None if v is None else torch.empty(v.shape, dtype=v.dtype, device=dev)
    with torch.cuda.stream(stream):
        for k, v in fields.items():
            if v is not None:
                out[k].copy_(v, non_blocking=True)
    return out


def bind_host_to_gpu(device=None):
    """Pin the calling thread (and the threads it starts later) to the CPUs NVML reports as nearest to `device`, so
    that pinned host batches allocated afterwards are first-touched on the GPU's NUMA node and the launching thread
    does not migrate across sockets.  (The one-GPU box of this pool is a single-node 16-vCPU VM, where this is a
    no-op; its upload rate still varies 18-36 GB/s run to run.)  Returns the CPU set applied, or None when NVML / the
    affinity call is not available (never an error: this is a placement hint)."""
    try:
        import pynvml
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(idx).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        near = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus = near & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
