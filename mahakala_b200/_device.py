"""Device-memory plumbing: torch owns HBM allocations and streams; kernels come from the C ABI.

The reference returns immutable ``jax.Array`` objects that live on the accelerator and convert to NumPy
on demand (``np.asarray``, ``np.allclose`` ...; used at tests/test_shadows.py:45 and geodesics.py:428 of
the reference).  ``DeviceArray`` is the equivalent here: a ``torch.Tensor`` subclass resident in HBM whose
``__array__`` copies to the host, so NumPy-protocol consumers keep working unchanged.
"""
import numpy as np
import torch

from . import _cabi


class DeviceArray(torch.Tensor):
    """A CUDA tensor that also satisfies the NumPy array protocol (device -> host copy on demand)."""

    @staticmethod
    def wrap(t):
        return t.as_subclass(DeviceArray) if isinstance(t, torch.Tensor) else t

    def __array__(self, dtype=None, copy=None):
        a = self.detach().as_subclass(torch.Tensor).cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def torch(self):
        return self.as_subclass(torch.Tensor)


def require_gpu():
    """The product path never falls back to the CPU: fail loudly without CUDA or without the library."""
    _cabi.load()
    if not torch.cuda.is_available():
        raise _cabi.MahakalaB200Error(
            "mahakala_b200 needs a CUDA device (built for sm_100a / B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def as_device(x, dtype=torch.float64):
    """array-like -> contiguous CUDA tensor (host data goes through pinned memory)."""
    dev = require_gpu()
    if isinstance(x, torch.Tensor):
        t = x.as_subclass(torch.Tensor)
        if t.device != dev or t.dtype != dtype:
            t = t.to(device=dev, dtype=dtype)
        return t.contiguous()
    a = np.ascontiguousarray(np.asarray(x), dtype={torch.float64: np.float64, torch.int32: np.int32,
                                                    torch.int64: np.int64, torch.float32: np.float32}[dtype])
    h = torch.from_numpy(a)
    if a.nbytes >= (1 << 16):
        h = h.pin_memory()
    return h.to(dev, non_blocking=True)


def empty(shape, dtype=torch.float64):
    return torch.empty(shape, dtype=dtype, device=require_gpu())


def zeros(shape, dtype=torch.float64):
    return torch.zeros(shape, dtype=dtype, device=require_gpu())


def bind_host_to_gpu_numa_node(device_index=None):
    """Pin this process (and therefore its first-touch page placement, pinned staging buffers included) to the
    CPUs of the NUMA node the GPU hangs off.  Matters for the zero-copy host-to-host path at one process per GPU:
    the kernel's PCIe reads of pinned host memory should not cross the inter-socket link.  Returns the node or
    None when the topology is not exposed (virtualised hosts report -1)."""
    import os
    try:
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        bus = torch.cuda.get_device_properties(idx).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(idx), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(idx), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return None
