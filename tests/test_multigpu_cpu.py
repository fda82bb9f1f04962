"""Host logic of the multi-GPU tile sharding, run with the gloo backend on CPU (world_size 2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, res, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mahakala_b200 import multigpu
    assert multigpu.world() == (rank, world)
    npatch = multigpu.patch_count(res)
    begin, end, stride = multigpu.static_assignment(npatch, rank, world)
    img = torch.zeros((2, res * res), dtype=torch.float64)
    mine = 0
    for p in range(begin, end, stride):
        pix = multigpu.patch_pixels(p, res)
        img[0, pix] = torch.from_numpy(pix + 1.0)
        img[1, pix] = float(rank + 1)
        mine += len(pix)
    multigpu.combine_static(img, dst=0)
    counts = torch.tensor([mine], dtype=torch.int64)
    dist.all_reduce(counts)
    if rank == 0:
        ok = bool(torch.equal(img[0], torch.arange(1, res * res + 1, dtype=torch.float64)))
        owners = set(img[1].tolist())
        out.put((ok, int(counts.item()), sorted(owners)))
    dist.barrier()
    dist.destroy_process_group()


def test_static_sharding_covers_image_once_gloo(built):
    res = 20                       # not a multiple of the 4x8 patch: edge patches are partial
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, res, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, total, owners = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and total == res * res and owners == [1.0, 2.0]


def test_patch_geometry_matches_library(built):
    from mahakala_b200 import _cabi, multigpu
    lib = _cabi.load()
    for res in (8, 20, 33, 1024):
        n = multigpu.patch_count(res)
        assert n == lib.mk_render_patch_count(res, None, 0)
        if res <= 33:
            seen = np.concatenate([multigpu.patch_pixels(p, res) for p in range(n)])
            assert np.array_equal(np.sort(seen), np.arange(res * res))
    assert multigpu.static_assignment(10, 1, 4) == (1, 10, 4)


def test_longest_first_ray_order_is_an_interleaved_permutation(built):
    """Queue order of a multi-frame integration job: a permutation of all rays, frames interleaved position by
    position, pixels by distance from the image centre (host logic of multigpu.integrate_distributed)."""
    from mahakala_b200 import multigpu
    res, frames = 12, 3
    order = multigpu.longest_first_ray_order(res, frames)
    n = res * res
    assert order.dtype == np.int32 and order.shape == (frames * n,)
    assert np.array_equal(np.sort(order), np.arange(frames * n))
    assert np.array_equal(order.reshape(n, frames) // n, np.tile(np.arange(frames), (n, 1)))      # frames interleaved
    pix = order.reshape(n, frames)[:, 0]
    ix, iy = pix // res, pix % res
    rho = np.hypot(ix + 0.5 - res / 2, iy + 0.5 - res / 2)
    assert np.all(np.diff(rho) >= 0)                                                               # centre first
    assert np.array_equal(multigpu.longest_first_ray_order(5, 1), multigpu.longest_first_ray_order(5))
    # SharedRays layout: offsets of the four result arrays (bytes) behind the 512 B of queue counters
    assert multigpu.HEADER_BYTES == 512


def test_pixel_interleaved_order_and_long_patch_split(built):
    """Host logic of the round-2 schedulers: the pixel-order / frames-interleaved queue of the multi-frame integration job
    (multigpu.interleaved_pixel_ray_order) and the split of a learned patch order between the warp-specialised
    long-patch kernel and the fused kernel (images.long_patch_count, images._long_grid_cap)."""
    from mahakala_b200 import images, multigpu
    res, frames = 10, 4
    n = res * res
    order = multigpu.interleaved_pixel_ray_order(res, frames)
    assert order.dtype == np.int32 and np.array_equal(np.sort(order), np.arange(frames * n))
    assert np.array_equal(order.reshape(n, frames) // n, np.tile(np.arange(frames), (n, 1)))      # frames interleaved
    assert np.array_equal(order.reshape(n, frames)[:, 2] % n, np.arange(n))                       # pixel order in a frame
    assert np.array_equal(multigpu.interleaved_pixel_ray_order(7, 1), np.arange(49))
    # learned lengths (descending): patches >= threshold x the longest, at most one per SM of every participant
    L = np.array([3765] * 10 + [2000] * 50 + [1000] * 400 + [900] * 3000 + [400] * 20000)
    assert images.long_patch_count(L, 1, threshold=0.5, sms=148) == 60
    assert images.long_patch_count(L, 1, threshold=0.25, sms=148) == 148           # capped: one per SM
    assert images.long_patch_count(L, 8, threshold=0.25, sms=148) == 460
    assert images.long_patch_count(L, 8, threshold=0.2, sms=148) == 8 * 148
    assert images.long_patch_count(None, 8, sms=148) == 0 and images.long_patch_count(np.zeros(4), 1, sms=148) == 0
    assert images._long_grid_cap(460, 1) == 0 and images._long_grid_cap(460, 8) >= 460 // (8 * max(1, images._LONG_EXCLUSIVE))
