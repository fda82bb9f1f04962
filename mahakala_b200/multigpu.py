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
        if local_image.is_cuda and dist.get_backend(group) == "gloo":
            dist.all_reduce(local_image, op=dist.ReduceOp.SUM, group=group)     # gloo has no CUDA reduce
        else:
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


HEADER_BYTES = 512          # queue counters, 128 B apart (0: patch / ray queue, 128: photon-ring patch queue)
PAGE_RANK_SHIFT = 26        # page locator of a multi-GPU dump: (rank << 26) + page number in that rank's own pool


class SharedBuffer:
    """``nbytes`` of rank-``owner`` device memory (zero-filled) mapped into every rank with CUDA IPC.  The first
    ``HEADER_BYTES`` hold the queue counters that the persistent kernels of all GPUs advance with system-scope
    atomics over NVLink; the payload behind them receives the results (the gather is fused into the kernels)."""

    def __init__(self, payload_bytes, owner=0):
        rank, _ = world()
        self.owner, self.rank = owner, rank
        self.nbytes = HEADER_BYTES + int(payload_bytes)
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
        self.ring_queue_ptr = self.ptr + 128
        self.payload_ptr = self.ptr + HEADER_BYTES

    def _raw(self):
        assert self.rank == self.owner
        return tensor_from_pointer(self.ptr, self.nbytes, torch.device("cuda", torch.cuda.current_device()))

    def reset(self):
        """Zero the queue counters (owner rank; call between jobs, followed by a barrier)."""
        if self.rank == self.owner:
            self._raw()[:HEADER_BYTES].zero_()
            torch.cuda.synchronize()        # a host-side (gloo) barrier after this must imply "counters are zero"

    def close(self):
        if self.ptr:
            _cabi.call("mk_ipc_free" if self.rank == self.owner else "mk_ipc_close", self.ptr)
            self.ptr = 0


class SharedImage(SharedBuffer):
    """[queue counters | image (nfreq, npx) f64] in rank ``owner``'s memory, mapped by every rank."""

    def __init__(self, nfreq, npx, owner=0):
        self.nfreq, self.npx = nfreq, npx
        super().__init__(8 * nfreq * npx, owner)
        self.image_ptr = self.payload_ptr

    def local_view(self):
        """(counter tensor, image tensor) on the owner rank."""
        raw = self._raw()
        return raw[:4].view(torch.int32), raw[HEADER_BYTES:].view(torch.float64).view(self.nfreq, self.npx)


class SharedRays(SharedBuffer):
    """[queue counters | final (n, 8) f64 | r_last (n) f64 | nsteps (n) i32 | page_first (n, 2) i32]: the per-ray
    results of a bundle integrated by all ranks together, gathered in rank ``owner``'s memory by the kernels."""

    def __init__(self, n, owner=0):
        self.n = int(n)
        super().__init__(self.n * (64 + 8 + 4 + 8), owner)
        self.final_ptr = self.payload_ptr
        self.r_last_ptr = self.final_ptr + 64 * self.n
        self.nsteps_ptr = self.r_last_ptr + 8 * self.n
        self.page_first_ptr = self.nsteps_ptr + 4 * self.n

    def results(self):
        return self.final_ptr, self.nsteps_ptr, self.r_last_ptr, self.page_first_ptr

    def local_views(self):
        """dict of tensors (final, r_last, nsteps, page_first) on the owner rank."""
        raw, n, o = self._raw(), self.n, HEADER_BYTES
        return {"final": raw[o:o + 64 * n].view(torch.float64).view(n, 8),
                "r_last": raw[o + 64 * n:o + 72 * n].view(torch.float64),
                "nsteps": raw[o + 72 * n:o + 76 * n].view(torch.int32),
                "page_first": raw[o + 76 * n:o + 84 * n].view(torch.int32).view(n, 2)}


def longest_first_ray_order(res, frames=1):
    """Queue order for ``frames`` bundles of res x res grid-camera rays stored back to back: pixels by distance from
    the image centre (the long photon-ring rays sit a few M from it for any spin and inclination), frames interleaved,
    so that every long ray of every frame starts early and the short outer rays fill the tail of the queue.  Returns an
    int32 NumPy array: queue position -> frame * res^2 + ix * res + iy."""
    c = (np.arange(res) + 0.5) - res / 2.0
    rho = np.hypot(c[:, None], c[None, :]).reshape(-1)
    pix = np.argsort(rho, kind="stable").astype(np.int64)
    order = (np.arange(frames, dtype=np.int64)[None, :] * (res * res) + pix[:, None]).reshape(-1)
    return order.astype(np.int32)


def interleaved_pixel_ray_order(res, frames=1):
    """Queue order for ``frames`` bundles stored back to back: pixel order, frames interleaved (queue position q ->
    frame q % frames, pixel q // frames), so that all frames progress together and every GPU sees the same mix of long
    and short rays at any time.  Measured on B200 (cfg2, paged dump, one GPU, L2 flushed between launches,
    ``scripts/dev/order_matrix_probe.py``): pixel order 16.3 ms against 17.1 ms centre-out -- with the dump, warps that
    all hold rays of the same length stay in lock-step and their slot stores arrive in bursts."""
    pix = np.arange(res * res, dtype=np.int64)
    order = (np.arange(frames, dtype=np.int64)[None, :] * (res * res) + pix[:, None]).reshape(-1)
    return order.astype(np.int32)


def integrate_distributed(N, s0, div, tol, bhspin, store, shared, ray_order=None):
    """Integrate ONE bundle ``s0`` (n, 8), resident on every rank, with all ranks together (trajectory-dump mode).

    All GPUs pull rays from the one queue in ``shared`` (a ``SharedRays``) and store each ray's final state, step
    count, classifier radius and page locator straight into the owner's memory over NVLink; trajectories are logged
    in each rank's own ``store`` (``page_first[:, 0] >> PAGE_RANK_SHIFT`` names the rank that holds a ray, the low bits
    the page in that rank's pool -- a fixed stride, because the pools of different ranks need not be equally large).  No
    collective on the data path; the two barriers only fence the queue reset.  Returns the owner's result views
    (``SharedRays.local_views()``) on the owner rank, None elsewhere."""
    from . import geodesics as geo
    rank, nranks = world()
    if store.max_pages >= (1 << PAGE_RANK_SHIFT) or nranks > (1 << (31 - PAGE_RANK_SHIFT)):
        raise ValueError("page pool or rank count too large for the rank-tagged page locator")
    shared.reset()
    if nranks > 1:
        dist.barrier()
    geo.integrate_paged(N, s0, div, tol, bhspin, store=store, queue=shared.queue_ptr, ray_order=ray_order,
                        results=shared.results(), page_id_offset=rank << PAGE_RANK_SHIFT, participants=nranks)
    torch.cuda.synchronize()
    if nranks > 1:
        dist.barrier()
    return shared.local_views() if rank == shared.owner else None


# ---------------------------------------------------------------------------------------------------
# snapshot replication
# ---------------------------------------------------------------------------------------------------
def warm_communicator():
    """Create the communicator and its channels (first-collective cost of NCCL: a few hundred ms) with one tiny
    all-reduce and one tiny broadcast, so that later timings measure the collectives and not NCCL start-up.
    Returns the milliseconds it took."""
    import time
    rank, nranks = world()
    if nranks == 1:
        return 0.0
    dev = torch.device("cuda", torch.cuda.current_device())
    t0 = time.perf_counter()
    x = torch.ones(1024, dtype=torch.float32, device=dev)
    dist.all_reduce(x)
    dist.broadcast(x, src=0)
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0)


def _cells_tensor(model):
    cells = ctypes.c_void_p()
    nbytes = ctypes.c_long()
    _cabi.call("mk_snapshot_cells", model.snapshot(), ctypes.byref(cells), ctypes.byref(nbytes))
    dev = torch.device("cuda", torch.cuda.current_device())
    return tensor_from_pointer(cells.value, nbytes.value, dev)


def broadcast_large(w, src=0):
    """Broadcast of a large 1-D device tensor.  Default: ``dist.broadcast`` -- NCCL's broadcast moves the 0.64 GB of
    float32 cells of the cfg4 snapshot to 8 B200s in 1.23 ms = 524 GB/s (``scripts/dev/bcast_probe.py``,
    ``profiles/r02_bcast_probe.txt``).  ``MK_BCAST=sag`` selects the scatter + all-gather formulation (van de Geijn: the
    root sends each rank one n-th, every rank collects the rest from its peers, in place), built to beat the ring and
    measured SLOWER on the NVSwitch node: 1.79 ms = 359 GB/s.  What the replication's ``broadcast`` phase costs beyond
    the wire (5-6 ms in all) is the float64 -> float32 -> float64 conversion and the 0.64 GB temporary on either side."""
    import os
    rank, n = world()
    numel = w.numel()
    if (n <= 2 or dist.get_backend() != "nccl" or numel * w.element_size() < (32 << 20) or numel % n != 0
            or os.environ.get("MK_BCAST", "nccl") != "sag"):
        dist.broadcast(w, src=src)
        return
    pieces = w.view(n, numel // n)
    mine = pieces[rank]
    dist.scatter(mine, scatter_list=list(pieces.unbind(0)) if rank == src else None, src=src)
    dist.all_gather_into_tensor(w, mine)


def replicate_snapshot(model=None, src=0, wire="auto"):
    """Give every rank the device snapshot held by rank ``src``.

    Rank ``src`` passes its ``AthenakFluidModel``; the other ranks may pass ``None`` (they receive the mesh
    geometry with ``broadcast_object_list`` and build a geometry-only replica) or a model of the same shape.
    Rank ``src`` uploads its interior arrays and runs the ghost-fill / repack kernel once; the cell array then travels
    to all other ranks with ONE NCCL broadcast over NVLink.  ``wire``: 'f64' sends float64 cells as they are, 'f32'
    / 'auto' send them as float32 when that is lossless (AthenaK writes float32, and so are its ghost-zone averages
    checked on the device) -- half the bytes on the wire, expanded again by the receivers; float32 snapshots always
    travel as stored.  Returns the (replica) model on every rank; ``model.replication_timing`` holds the phases in
    ms: host_prep / upload / ghost_fill (rank ``src``; zero elsewhere), meta (geometry exchange), broadcast (device
    time of the collective INCLUDING the wire conversion on both sides), collective (the NCCL broadcast alone,
    ``broadcast_large``), wire_bytes, effective GB/s of both.
    """
    import time
    from .grmhd.athenak import AthenakFluidModel
    rank, nranks = world()
    if nranks == 1:
        model.snapshot()
        model.replication_timing = dict(getattr(model, "setup_timing", {}), meta=0.0, broadcast=0.0, wire_bytes=0)
        return model
    dev = torch.device("cuda", torch.cuda.current_device())
    meta = [None]
    t64 = None
    if rank == src:
        model.snapshot()
        m = model.replica_meta()
        m["wire"] = model.storage
        if model.storage == "f64" and wire in ("auto", "f32"):
            t64 = _cells_tensor(model).view(torch.float64)
            lossless = bool(torch.equal(t64.to(torch.float32).to(torch.float64), t64))
            if lossless:
                m["wire"] = "f32"
            elif wire == "f32":
                raise ValueError("float32 wire format requested but the cells are not float32-representable")
        meta = [m]
    t0 = time.perf_counter()
    dist.broadcast_object_list(meta, src=src)
    if rank != src:
        if model is None:
            kw = dict(meta[0])
            kw.pop("wire")
            model = AthenakFluidModel.replica(**kw)
        else:
            model._storage = meta[0]["storage"]
        model.snapshot(fill=False)
    torch.cuda.synchronize()
    t_meta = 1e3 * (time.perf_counter() - t0)
    t = _cells_tensor(model)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if meta[0]["wire"] == "f32" and meta[0]["storage"] == "f64":
        t64 = t.view(torch.float64)
        w = t64.to(torch.float32) if rank == src else torch.empty(t64.shape, dtype=torch.float32, device=dev)
        w0.record()
        broadcast_large(w.view(-1), src=src)
        w1.record()
        if rank != src:
            t64.copy_(w)                      # float32 -> float64 is exact
        wire_bytes = w.numel() * 4
        del w
    else:
        w0.record()
        broadcast_large(t.view(-1), src=src)
        w1.record()
        wire_bytes = t.numel()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    wire_ms = w0.elapsed_time(w1)
    # the root enqueues its broadcast first and then waits for the receivers (still allocating): the rank that arrives
    # LAST sees the transfer alone, so the minimum over ranks is the wire time
    wmin = torch.tensor([wire_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
    wire_min = float(wmin)
    base = getattr(model, "setup_timing", {}) if rank == src else {}
    model.replication_timing = dict(host_prep=base.get("host_prep", 0.0), upload=base.get("upload", 0.0),
                                    ghost_fill=base.get("ghost_fill", 0.0), meta=t_meta, broadcast=ms,
                                    wire_bytes=int(wire_bytes), wire_format=meta[0]["wire"],
                                    broadcast_GBps=wire_bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
                                    collective=wire_ms, collective_min=wire_min,
                                    collective_GBps=wire_bytes / (wire_min * 1e-3) / 1e9 if wire_min > 0 else 0.0)
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
        images.render(model, image_out=shared.image_ptr, queue=shared.queue_ptr, long_queue=shared.ring_queue_ptr,
                      participants=nranks, **render_kwargs)
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
