"""Host <-> device copies of large NumPy arrays for the drop-in signatures.

The reference's functions take and return NumPy arrays, so the literal drop-ins (``build_dictionary``,
``compute_derivatives``, ``fit_from_field`` on a host stack ...) are bounded by the copy of pageable memory, which
the driver stages through a small internal buffer at ~10 GB/s.  Here arrays of at least ``MIN_BYTES`` go through two
pinned staging buffers: a multi-threaded host copy (torch's CPU copy kernel, GIL released) fills one while the copy
engine drains the other at PCIe rate.  Smaller arrays take the plain path.  Nothing here touches values.
"""

from __future__ import annotations

import threading
import warnings

import numpy as np

from . import _lib as L

MIN_BYTES = 16 << 20        # below this the plain pageable copy is as fast
STAGE_BYTES = 32 << 20      # per staging buffer (two of them, allocated on first use)

_stage = {}                 # device index -> (buffers, events)
_lock = threading.Lock()    # the two staging buffers of a device serve one transfer at a time


def _staging(torch):
    dev = torch.cuda.current_device()
    if dev not in _stage:
        bufs = [torch.empty(STAGE_BYTES, dtype=torch.uint8).pin_memory() for _ in range(2)]
        _stage[dev] = (bufs, [None, None])
    return _stage[dev]


def bind_host_to_gpu(device=None) -> dict:
    """Restrict this process to the CPUs NVML names as local to the GPU (its NUMA node), so that the pinned buffers
    allocated afterwards are first-touched next to the GPU's PCIe root and host -> device copies do not cross the
    socket interconnect.  One process per GPU calls it once, before allocating pinned memory.  Returns what it did
    (``{"bound": bool, "cpus": n, "numa_node": k or None, "why": "..."}``); never raises: a box without NVML, without
    NUMA information (a VM) or with a cpuset that excludes the local CPUs is left as it is."""
    import os
    info = {"bound": False, "cpus": None, "numa_node": None, "why": ""}
    try:
        torch = L.torch_cuda()
        dev = torch.cuda.current_device() if device is None else int(device)
        import pynvml as n
        n.nvmlInit()
        pr = torch.cuda.get_device_properties(dev)
        try:
            h = n.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        except Exception:
            h = n.nvmlDeviceGetHandleByIndex(dev)
        try:
            info["numa_node"] = int(n.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        have = os.sched_getaffinity(0)
        words = n.nvmlDeviceGetCpuAffinity(h, (max(have | {os.cpu_count() or 1}) + 64) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        want = local & have
        if not want:
            info["why"] = "the GPU's local CPUs are outside this process's cpuset"
        elif want == have:
            info["why"] = "already local (or the box exposes one NUMA node)"
            info["cpus"] = len(have)
        else:
            os.sched_setaffinity(0, want)
            info.update(bound=True, cpus=len(want))
    except Exception as e:                                   # noqa: BLE001 - diagnostic only
        info["why"] = f"{type(e).__name__}: {e}"[:120]
    return info


def _as_tensor(torch, a):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)     # read-only arrays: we only read
        return torch.from_numpy(a)


def to_device(a: np.ndarray, out=None):
    """NumPy array -> CUDA tensor of the same dtype and shape on the current device / stream (``out``: a contiguous
    CUDA tensor of that dtype and size to fill instead of a new one)."""
    torch = L.torch_cuda()
    a = np.ascontiguousarray(a)
    src = _as_tensor(torch, a)
    if out is not None and (not out.is_contiguous() or out.dtype != src.dtype or out.numel() != src.numel()):
        raise ValueError("out must be a contiguous CUDA tensor of the array's dtype and size")
    if a.nbytes < MIN_BYTES:
        return src.cuda() if out is None else out.view(-1).copy_(src.reshape(-1), non_blocking=True).view(out.shape)
    flat = src.reshape(-1)
    dst = torch.empty(a.shape, dtype=src.dtype, device="cuda") if out is None else out
    dflat = dst.view(-1)
    per = STAGE_BYTES // a.itemsize
    stream = torch.cuda.current_stream()
    with _lock:
        bufs, evs = _staging(torch)
        for k, lo in enumerate(range(0, flat.numel(), per)):
            hi = min(flat.numel(), lo + per)
            b = k & 1
            if evs[b] is not None:
                evs[b].synchronize()                         # the copy engine has drained this buffer
            pb = bufs[b][: (hi - lo) * a.itemsize].view(src.dtype)
            pb.copy_(flat[lo:hi])                            # multi-threaded host copy into pinned memory
            dflat[lo:hi].copy_(pb, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            evs[b] = ev
    return dst


def to_host(t) -> np.ndarray:
    """CUDA tensor -> new NumPy array (synchronises the current stream)."""
    torch = L.torch_cuda()
    t = t.detach()
    if not t.is_cuda or t.numel() * t.element_size() < MIN_BYTES:
        return t.cpu().numpy()
    t = t.contiguous()
    out = np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
    oflat = _as_tensor(torch, out).reshape(-1)
    tflat = t.view(-1)
    per = STAGE_BYTES // t.element_size()
    stream = torch.cuda.current_stream()
    with _lock:
        bufs, evs = _staging(torch)
        pending = None                                   # (pinned chunk, lo, hi, event) whose device -> pinned copy is in flight
        for k, lo in enumerate(range(0, tflat.numel(), per)):
            hi = min(tflat.numel(), lo + per)
            b = k & 1
            if evs[b] is not None:
                evs[b].synchronize()
            pb = bufs[b][: (hi - lo) * t.element_size()].view(t.dtype)
            pb.copy_(tflat[lo:hi], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            evs[b] = ev
            if pending is not None:                      # drain the previous chunk while this one is in flight
                pbuf, plo, phi, pev = pending
                pev.synchronize()
                oflat[plo:phi].copy_(pbuf)
            pending = (pb, lo, hi, ev)
        pbuf, plo, phi, pev = pending
        pev.synchronize()
        oflat[plo:phi].copy_(pbuf)
    return out
