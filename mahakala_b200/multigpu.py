"""Multi-GPU rendering: one process per GPU (torch.distributed), image tiles sharded over the ranks.

The reference is single-device (its only scaling device is the sequential pixel-chunk loop of
/root/reference/mahakala/images.py:67-78).  Rays are independent and the snapshot is read-only, so the
image plane shards naturally into 32-ray patches:

* ``mode='queue'``  — ONE dynamic tile queue for all GPUs.  Rank 0 owns a small buffer (queue counter +
  the image) that every other rank maps with CUDA IPC; the persistent render kernels of all ranks pull
  patches with ``atomicAdd`` on that counter over NVLink and store finished pixels directly into rank 0's
  image (the gather is fused into the kernel; no collective on the data path).
* ``mode='static'`` — patch p goes to rank p % world; every rank renders into a private zero image and the
  disjoint tiles are combined with one NCCL reduce (sum of disjoint supports is exact).  Used when peer
  mapping is unavailable, and with the gloo backend in the CPU tests of the host logic.

The snapshot is replicated: rank ``src`` repacks it, the others receive the cell array with one NCCL
broadcast (``replicate_snapshot``).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi


# ---------------------------------------------------------------------------------------------------
# host logic (backend-agnostic; exercised with gloo on CPU)
# ---------------------------------------------------------------------------------------------------
PATCH_X, PATCH_Y = 4, 8


def patch_count(res):
    return (-(-res // PATCH_X)) * (-(-res // PATCH_Y))


def patch_pixels(patch, res):
    """Flat pixel indices (ix*res + iy) of a grid-camera patch, as the render kernel assigns them."""
    py_n = -(-res // PATCH_Y)
    px, py = divmod(patch, py_n)
    ix = px * PATCH_X + np.arange(32) // 8
    iy = py * PATCH_Y + np.arange(32) % 8
    ok = (ix < res) & (iy < res)
    return (ix * res + iy)[ok]


def static_assignment(npatches, rank, world):
    """(begin, end, stride) of the interleaved static sharding: patches rank, rank+world, ..."""
    return rank, npatches, world


def combine_static(local_image, dst=0, group=None):
    """Sum the disjoint per-rank images onto ``dst`` (in place).  Works on CUDA (NCCL) and CPU (gloo)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(local_image, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return local_image


def world():
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ---------------------------------------------------------------------------------------------------
# peer memory
# ---------------------------------------------------------------------------------------------------
class _RawCuda:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def tensor_from_pointer(ptr, nbytes, device):
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=device)


class SharedImage:
    """Rank-``owner`` buffer [256 B queue counter | image (nfreq, npx) f64] mapped by every rank."""

    def __init__(self, nfreq, npx, owner=0):
        rank, _ = world()
        self.owner, self.rank = owner, rank
        self.nfreq, self.npx = nfreq, npx
        self.nbytes = 256 + 8 * nfreq * npx
        handle = (ctypes.c_ubyte * 64)()
        ptr = ctypes.c_void_p()
        box = [None]
        if rank == owner:
            _cabi.call("mk_ipc_alloc", self.nbytes, ctypes.byref(ptr), handle)
            box = [bytes(handle)]
        if dist.is_initialized():
            dist.broadcast_object_list(box, src=owner)
        if rank != owner:
            handle = (ctypes.c_ubyte * 64).from_buffer_copy(box[0])
            _cabi.call("mk_ipc_open", handle, ctypes.byref(ptr))
        self.ptr = ptr.value
        self.queue_ptr = self.ptr
        self.image_ptr = self.ptr + 256

    def local_view(self):
        """(counter tensor, image tensor) on the owner rank."""
        assert self.rank == self.owner
        dev = torch.device("cuda", torch.cuda.current_device())
        raw = tensor_from_pointer(self.ptr, self.nbytes, dev)
        return raw[:4].view(torch.int32), raw[256:].view(torch.float64).view(self.nfreq, self.npx)

    def reset(self):
        if self.rank == self.owner:
            q, _ = self.local_view()
            q.zero_()

    def close(self):
        if self.ptr:
            _cabi.call("mk_ipc_free" if self.rank == self.owner else "mk_ipc_close", self.ptr)
            self.ptr = 0


# ---------------------------------------------------------------------------------------------------
# snapshot replication
# ---------------------------------------------------------------------------------------------------
def replicate_snapshot(model=None, src=0):
    """Give every rank the device snapshot held by rank ``src``.

    Rank ``src`` passes its ``AthenakFluidModel``; the other ranks may pass ``None`` (they receive the mesh
    geometry with ``broadcast_object_list`` and build a geometry-only replica) or a model of the same shape.
    Rank ``src`` repacks its host arrays once; the cell array then travels to all other ranks with ONE NCCL
    broadcast over NVLink.  Returns the (replica) model on every rank.
    """
    from .grmhd.athenak import AthenakFluidModel
    rank, nranks = world()
    if nranks == 1:
        model.snapshot()
        return model
    meta = [None]
    if rank == src:
        model.snapshot()
        meta = [model.replica_meta()]
    dist.broadcast_object_list(meta, src=src)
    if rank != src:
        if model is None:
            model = AthenakFluidModel.replica(**meta[0])
        else:
            model._storage = meta[0]["storage"]
        model.snapshot(fill=False)
    cells = ctypes.c_void_p()
    nbytes = ctypes.c_long()
    _cabi.call("mk_snapshot_cells", model.snapshot(), ctypes.byref(cells), ctypes.byref(nbytes))
    dev = torch.device("cuda", torch.cuda.current_device())
    t = tensor_from_pointer(cells.value, nbytes.value, dev)
    dist.broadcast(t, src=src)
    torch.cuda.synchronize()
    return model


# ---------------------------------------------------------------------------------------------------
# distributed render
# ---------------------------------------------------------------------------------------------------
def render_distributed(model, mode="queue", shared=None, dst=0, **render_kwargs):
    """Render one image with all ranks.  Returns the (nfreq, npx) image tensor on rank ``dst`` (None elsewhere).

    ``shared`` (a ``SharedImage``) can be passed to reuse the peer mapping across frames.
    """
    from . import images
    rank, nranks = world()
    res = int(render_kwargs.get("resolution", 160))
    nus = np.atleast_1d(render_kwargs.get("observing_frequencies", (230.e9,)))
    npx = res * res
    if nranks == 1:
        return images.render(model, **render_kwargs)
    if mode == "queue":
        own = shared is None
        if own:
            shared = SharedImage(len(nus), npx, owner=dst)
        shared.reset()
        dist.barrier()
        images.render(model, image_out=shared.image_ptr, queue=shared.queue_ptr, **render_kwargs)
        torch.cuda.synchronize()
        dist.barrier()
        out = None
        if rank == dst:
            out = shared.local_view()[1].clone()
        if own:
            dist.barrier()
            shared.close()
        return out
    begin, end, stride = static_assignment(patch_count(res), rank, nranks)
    img = images.render(model, patch_range=(begin, -1, stride), **render_kwargs)
    combine_static(img, dst=dst)
    return img if rank == dst else None
