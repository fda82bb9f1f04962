"""Drop-in for ``mahakala.images`` (reference: /root/reference/mahakala/images.py:30-144).

``make_image`` keeps the reference signature.  With an ``AthenakFluidModel`` it runs the FUSED kernel
(camera -> RK4 geodesic -> snapshot sample -> j, alpha -> intensity in registers, no trajectories in
memory, no pixel chunking).  ``make_image_unfused`` executes the reference's stage-by-stage chain on the
device (integrate/dump -> sample -> Theta_e -> j, alpha -> sigma cut -> back-to-front transfer) and is
what any other fluid-model duck type goes through.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from . import geodesics as geo
from .constants import Msun
from .electrons import rlow_rhigh_model
from .transfer import emission_params, solve_specific_intensity, synchrotron_coefficients
from ._device import as_device, empty, require_gpu, stream_ptr


def _params_for(fluid_model, M_bh, mass_scale, r_high):
    units = fluid_model.get_units(M_bh, mass_scale)
    return emission_params(fluid_gamma=fluid_model.fluid_gamma, r_high=r_high, Ne_unit=units['Ne_unit'],
                           B_unit=units['B_unit'], L_unit=units['L_unit']), units


_order_cache = {}
_staging = {}


def _pinned_staging(n):
    """Reusable pinned host buffer for device -> host image copies (cudaHostAlloc costs milliseconds)."""
    if n not in _staging:
        if len(_staging) > 8:
            _staging.clear()
        _staging[n] = torch.empty((n,), dtype=torch.float64, pin_memory=True)
    return _staging[n]


def centre_out_patch_order(res, device):
    """Scheduling order of the 4x8-pixel patches of a res x res grid camera: by distance of the patch centre
    from the image centre.  The long rays (photon ring, a few M from the centre for any spin/inclination) are
    then handed out early and the short outer rays fill the tail of the dynamic queue."""
    key = (int(res), str(device))
    if key not in _order_cache:
        px_n, py_n = -(-res // 4), -(-res // 8)
        cx = (np.arange(px_n) * 4 + 2.0) - res / 2.0
        cy = (np.arange(py_n) * 8 + 4.0) - res / 2.0
        rho = np.hypot(cx[:, None], cy[None, :]).reshape(-1)          # patch index = px * py_n + py
        order = np.argsort(rho, kind="stable").astype(np.int32)
        _order_cache[key] = torch.from_numpy(order).to(device)
    return _order_cache[key]


_learned_order = {}
_learned_lengths = {}        # same keys: steps of the longest ray of every patch, in the learned (descending) order
_priority_streams = {}


def _priority_stream(dev):
    if dev not in _priority_streams:
        _priority_streams[dev] = torch.cuda.Stream(device=dev, priority=-1)
    return _priority_streams[dev]


import os as _os
QUICK_ORDER = _os.environ.get("MK_QUICK_ORDER", "1") == "1"    # coarse pre-pass in front of a cold multi-GPU frame
_DEV_SKIP_BULK = False      # scripts/dev/strong_probe.py: time the long-patch launch alone
_LONG_EXCLUSIVE = int(_os.environ.get("MK_LONG_EXCLUSIVE", "2"))     # 0 = shared SMs; 1..4 = groups per exclusive CTA


def _long_grid_cap(n_long, participants):
    """CTAs of the long-patch launch per GPU: its share of the long patches (+ 25 %), so that with several GPUs on
    one queue they spread over all of them instead of being taken by the GPU that starts first"""
    per_cta = max(1, int(_LONG_EXCLUSIVE))
    if participants <= 1:
        return 0
    return max(1, int(np.ceil(1.25 * n_long / per_cta / participants)))


def long_patch_count(lengths, participants=1, threshold=None, device=None, sms=None):
    """How many patches at the head of a learned (longest-first) order go to the warp-specialised long-patch kernel
    (``mk_render_long``): those whose longest ray takes at least ``threshold`` x the longest ray of the frame (default
    0.25, ``MK_LONG_THRESHOLD``), at most one per SM of every participating GPU, so that each of them has a CTA from
    time zero.  Measured on 8 B200s, one 1024^2 cfg4 frame (``scripts/dev/strong_probe.py``): 9.1-10.1 ms without the
    long-patch launch, 7.2 / 5.6 / 4.6-5.0 / 4.4-4.5 / 4.8-5.0 ms for thresholds 0.5 / 0.4 / 0.3 / 0.25 / 0.2 -- below
    0.25 the long-patch kernel (68 % of the fused kernel's throughput per SM) takes too much of the machine, above it the
    longest patch left to the fused kernel (3.1 us per step at full occupancy) sets the time."""
    import os
    if lengths is None or len(lengths) == 0 or lengths[0] <= 0:
        return 0
    if threshold is None:
        threshold = float(os.environ.get("MK_LONG_THRESHOLD", "0.25"))
    n = int((lengths >= threshold * float(lengths[0])).sum())
    if sms is None:
        sms = torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).multi_processor_count
    return max(0, min(n, int(participants) * sms))


def _camera_key(bhspin, camera_inclination, camera_distance, fov, resolution, max_nsteps, div, tol, device):
    return (float(bhspin), float(camera_inclination), float(camera_distance), float(fov), int(resolution),
            int(max_nsteps), float(div), float(tol), str(device))


def learn_patch_order(bhspin, camera_inclination=60, camera_distance=1000, fov=20, resolution=160, max_nsteps=10000,
                      div=40, tol=1e-4):
    """Profile-guided scheduling for REPEATED frames of one camera (a movie, a frequency or model sweep): the step
    count of a ray depends on the camera and the spacetime only, not on the fluid, so one geodesics-only pass
    (``integrate_final``, ~15 ms for 1024^2 rays) gives the exact length of every 4x8-pixel patch.  Later
    ``render`` calls with the same camera hand the patches out longest first, the optimal list schedule for the
    dynamic queue: the patch that contains the longest photon-ring ray (a chain of dependent RK4 steps nothing can
    shorten) starts at time zero instead of whenever the centre-out heuristic reaches it.  Results are unaffected
    (scheduling only).  Returns the permutation (device int32 tensor)."""
    dev = require_gpu()
    res = int(resolution)
    s0 = geo.initialize_geodesics_at_camera(bhspin, camera_inclination, camera_distance, -fov / 2., fov / 2., res)
    _, nsteps, _ = geo.integrate_final(max_nsteps, s0, div, tol, bhspin)
    px_n, py_n = -(-res // 4), -(-res // 8)
    img = torch.zeros((px_n * 4, py_n * 8), dtype=torch.int32, device=dev)
    img[:res, :res] = nsteps.view(res, res)
    longest = img.view(px_n, 4, py_n, 8).amax(dim=(1, 3)).reshape(-1)          # patch index = px * py_n + py
    order = torch.argsort(longest, descending=True, stable=True).to(torch.int32)
    key = _camera_key(bhspin, camera_inclination, camera_distance, fov, res, max_nsteps, div, tol, dev)
    _learned_order[key] = order
    _learned_lengths[key] = longest[order.long()].cpu().numpy()
    _quick_long.pop(key, None)
    return order


_quick_long = {}            # camera key -> number of long patches found by quick_patch_order (absolute rule)


def forget_patch_orders():
    """Drop every learned / quick patch order (scheduling state only)."""
    _learned_order.clear(); _learned_lengths.clear(); _quick_long.clear()


def quick_patch_order(bhspin, camera_inclination=60, camera_distance=1000, fov=20, resolution=160, max_nsteps=10000,
                      div=40, tol=1e-4, coarse=8, cap=1024, long_factor=2.5):
    """A COARSE, CAPPED version of ``learn_patch_order`` that is cheap enough to run in front of a single cold frame:
    geodesics only, every ``coarse``-th pixel per axis, at most ``cap`` steps (~0.8 ms on B200: 1024 dependent steps of
    0.64 us; the full pass of ``learn_patch_order`` costs the longest ray's 3765 steps plus 15 ms of throughput at 1024^2).
    A patch counts as LONG when the coarse rays around it (3x3 dilation: the photon ring is thinner than a coarse
    pixel, its neighbourhood is not) need at least ``long_factor`` x the shortest ray of the frame -- the absolute form of
    "a quarter of the longest ray", which a capped pass cannot know.  Patches are ordered by that estimate, longest
    first.  Deterministic, so every rank of a multi-GPU job computes the same order by itself.  ``render`` calls it on
    its own when several GPUs share one frame and nothing better has been learned (``patch_order='auto'``); results
    never depend on it.  Returns ``(order, n_long)``."""
    dev = require_gpu()
    res = int(resolution)
    px_n, py_n = -(-res // 4), -(-res // 8)
    cres = max(8, -(-res // int(coarse)))
    s0 = geo.initialize_geodesics_at_camera(bhspin, camera_inclination, camera_distance, -fov / 2., fov / 2., cres)
    _, nsteps, _ = geo.integrate_final(min(int(cap), int(max_nsteps)), s0, div, tol, bhspin)
    est = nsteps.view(1, 1, cres, cres).to(torch.float32)
    est = torch.nn.functional.max_pool2d(est, kernel_size=3, stride=1, padding=1)            # 3x3 dilation
    # patch (px, py) covers pixels [4 px, 4 px + 4) x [8 py, 8 py + 8): its centre in coarse-pixel units
    cx = ((torch.arange(px_n, device=dev, dtype=torch.float32) * 4 + 2.0) * (cres / float(res))).long().clamp_(0, cres - 1)
    cy = ((torch.arange(py_n, device=dev, dtype=torch.float32) * 8 + 4.0) * (cres / float(res))).long().clamp_(0, cres - 1)
    longest = est[0, 0][cx][:, cy].reshape(-1)                                                # patch index = px * py_n + py
    order = torch.argsort(longest, descending=True, stable=True).to(torch.int32)
    shortest = float(nsteps[nsteps > 0].min()) if bool((nsteps > 0).any()) else 0.0
    n_long = int((longest >= long_factor * shortest).sum()) if shortest > 0 else 0
    key = _camera_key(bhspin, camera_inclination, camera_distance, fov, res, max_nsteps, div, tol, dev)
    _learned_order[key] = order
    _learned_lengths[key] = longest[order.long()].cpu().numpy()
    _quick_long[key] = n_long
    return order, n_long


def render(fluid_model, camera_inclination=60, camera_distance=1000, mass_scale=1.e26, M_bh=6.2e9 * Msun,
           r_high=40, observing_frequencies=(230.e9,), fov=20, resolution=160, max_nsteps=10000, s0=None,
           div=40, tol=1e-4, image_out=None, queue=None, patch_range=(0, -1, 1), want_counters=False,
           patch_order="auto", long_patches="auto", long_queue=None, participants=1):
    """Fused multi-frequency render.  Returns ``image (nfreq, npx)`` on the device (plus counters).

    ``s0`` (npx, 8) replaces the grid camera by explicit rays.  ``image_out`` / ``queue`` may be tensors
    or raw device pointers (possibly in a peer GPU's memory) — see ``mahakala_b200.multigpu``.
    ``patch_order``: 'auto' = the order learned for this camera by ``learn_patch_order`` if there is one, else -- when
    several GPUs share the frame (``participants`` > 1 with both queues) -- the order of ``quick_patch_order``, computed
    here once per camera, else 'centre_out'; 'centre_out'; None = row-major; or an int32 device tensor.
    ``long_patches``: how many patches at the head of ``patch_order`` run through the warp-specialised long-patch kernel
    (``mk_render_long``: the sample leaves the critical path of the ray's dependent RK4 steps) on a high-priority
    stream next to the bulk launch; 'auto' = ``long_patch_count`` of the learned order (0 without one); built-in
    spacetime, grid camera, whole frames only.  ``long_queue``: its queue counter when ``queue`` is shared by
    several GPUs (``participants`` of them); pixels are bit-identical either way.
    """
    dev = require_gpu()
    snap = fluid_model.snapshot()
    P, units = _params_for(fluid_model, M_bh, mass_scale, r_high)
    nus = np.atleast_1d(np.asarray(observing_frequencies, dtype=np.float64))
    nfreq = nus.size
    if nfreq < 1:
        raise ValueError("at least one observing frequency")
    if nfreq > 8:
        # one launch carries up to 8 frequencies (they share the geodesic and the samples); more are batched
        if image_out is not None or queue is not None or want_counters:
            raise ValueError("at most 8 observing frequencies per launch with image_out / queue / want_counters")
        parts = [render(fluid_model, camera_inclination, camera_distance, mass_scale, M_bh, r_high, nus[k:k + 8], fov,
                        resolution, max_nsteps, s0, div, tol, None, None, patch_range, False, patch_order)
                 for k in range(0, nfreq, 8)]
        return torch.cat(parts, dim=0)
    c_nu = (ctypes.c_double * 8)(*(list(nus) + [nus[-1]] * (8 - nfreq)))
    if s0 is not None:
        s0d = as_device(s0)
        npx = s0d.shape[0]
        res = 0
    else:
        s0d = None
        res = int(resolution)
        npx = res * res
    if image_out is not None:
        img = image_out
    elif tuple(patch_range[:2]) == (0, -1) and (len(patch_range) < 3 or patch_range[2] == 1):
        img = empty((nfreq, npx))
    else:
        img = torch.zeros((nfreq, npx), dtype=torch.float64, device=dev)    # partial render: others stay 0
    nsteps = None
    counters = torch.zeros(2, dtype=torch.int64, device=dev) if want_counters else None
    order = None
    lengths = None
    quick_n = None
    if s0d is None:
        if isinstance(patch_order, torch.Tensor):
            order = patch_order
        elif patch_order == "auto":
            key = _camera_key(fluid_model.bhspin, camera_inclination, camera_distance, fov, res, max_nsteps, div, tol, dev)
            order = _learned_order.get(key)
            if (order is None and participants > 1 and QUICK_ORDER and long_patches == "auto"
                    and geo._active_metric == geo.KERR_SCHILD and queue is not None and long_queue is not None):
                # several GPUs share ONE frame and nothing is known about this camera: a coarse, capped geodesics-only
                # pre-pass (~1 ms, once per camera, identical on every rank) finds the photon-ring patches
                order, _ = quick_patch_order(fluid_model.bhspin, camera_inclination, camera_distance, fov, res,
                                             max_nsteps, div, tol)
            lengths = _learned_lengths.get(key) if order is not None else None
            quick_n = _quick_long.get(key) if order is not None else None
            if order is None:
                order = centre_out_patch_order(res, dev)
        elif patch_order == "centre_out":
            order = centre_out_patch_order(res, dev)
    i = camera_inclination * np.pi / 180
    # spacetime of the launch: the built-in closed-form Kerr-Schild kernel, or the NVRTC-built kernel of the active
    # run-time registered metric (geodesics AND fluid frame in that metric)
    metric_id = geo._active_metric if geo._active_metric >= geo.PLUGIN_BASE else geo.KERR_SCHILD
    if geo._active_metric == geo.KERR_SCHILD_STRICT:
        raise ValueError("the strict (literal IEEE) integrator has no fused render: use make_image_unfused")
    whole = tuple(patch_range[:2]) == (0, -1) and (len(patch_range) < 3 or patch_range[2] == 1)
    n_long = 0
    if order is not None and whole and metric_id == geo.KERR_SCHILD and (queue is None) == (long_queue is None):
        if long_patches == "auto" and quick_n is not None:
            n_long = min(int(quick_n), int(participants) * torch.cuda.get_device_properties(dev).multi_processor_count)
        elif long_patches == "auto":
            n_long = long_patch_count(lengths, participants, device=dev)
        elif long_patches:
            n_long = min(int(long_patches), int(order.numel()))
    c_steps = counters[0:1] if want_counters else None
    c_samples = counters[1:2] if want_counters else None
    if n_long > 0:
        # positions [0, n_long) of the order: long-patch kernel, launched first on a high-priority stream (its CTAs are
        # placed before the bulk kernel's, one per SM); the rest: the fused kernel on the caller's stream
        cur, side = torch.cuda.current_stream(), _priority_stream(dev)
        side.wait_stream(cur)
        _cabi.call("mk_render_long", float(fluid_model.bhspin), float(np.cos(i)), float(np.sin(i)), float(camera_distance),
                   -fov / 2., fov / 2., res, s0d, npx, int(max_nsteps), float(div), float(tol), snap, P, nfreq, c_nu,
                   img, nsteps, c_steps, c_samples, long_queue, 0, n_long, 1, order, int(_LONG_EXCLUSIVE),
                   _long_grid_cap(n_long, participants), side.cuda_stream)
        patch_range = (n_long, -1, 1)
    if not (n_long > 0 and _DEV_SKIP_BULK):
        _cabi.call("mk_render_metric", int(metric_id), float(fluid_model.bhspin), float(np.cos(i)), float(np.sin(i)), float(camera_distance),
               -fov / 2., fov / 2., res, s0d, npx, int(max_nsteps), float(div), float(tol), snap, P, nfreq, c_nu,
               img, nsteps, c_steps, c_samples,
               queue, int(patch_range[0]), int(patch_range[1]), int(patch_range[2]) if len(patch_range) > 2 else 1,
               order, stream_ptr())
    if n_long > 0:
        cur.wait_stream(side)
    if want_counters:
        return img, counters
    return img


def make_image(fluid_model, camera_inclination=60, camera_distance=1000,
               mass_scale=1.e26, M_bh=6.2e9 * Msun, r_high=40,
               observing_frequency=230.e9,
               fov=20, resolution=160,
               max_nsteps=10000,
               max_chunk_bytes=None):
    """images.py:30-144.  Returns a (resolution, resolution) NumPy array of specific intensities in cgs.

    ``max_chunk_bytes`` is accepted for signature compatibility; the fused kernel stores no trajectories,
    so there is nothing to chunk (the unfused fallback for foreign fluid models does honour it).
    """
    # The fused kernel runs in the active spacetime: the built-in closed-form Kerr-Schild metric, or a user-registered
    # one (geodesics.register_metric / set_metric), for which NVRTC built the same kernel around the user's metric --
    # geodesics from its dual-number derivatives, fluid frame (athenak.py:760-786) from its g and g^-1 at every
    # sample.  (The stage-by-stage chain make_image_unfused follows the active metric in the integrator only, its
    # sampling kernel keeps the Kerr-Schild frame -- the reference's behaviour when only geodesics.metric is swapped,
    # since athenak.py:34 binds the Kerr-Schild metric at import.)
    if hasattr(fluid_model, "snapshot") and geo._active_metric != geo.KERR_SCHILD_STRICT:
        # the kernel stores the pixels straight into pinned host memory (mapped under unified addressing):
        # no device image, no separate device -> host copy
        host = _pinned_staging(resolution * resolution)
        render(fluid_model, camera_inclination, camera_distance, mass_scale, M_bh, r_high,
               (observing_frequency,), fov, resolution, max_nsteps, image_out=host)
        torch.cuda.current_stream().synchronize()
        return host.numpy().copy().reshape((resolution, resolution))
    return make_image_unfused(fluid_model, camera_inclination, camera_distance, mass_scale, M_bh, r_high,
                              observing_frequency, fov, resolution, max_nsteps, max_chunk_bytes)


def intensity_from_trajectories(fluid_model, S, final_dt, mass_scale, M_bh, r_high, observing_frequency):
    """images.py:84-120 for one chunk, stage by stage on the device (reference order of operations)."""
    fluid_gamma = fluid_model.fluid_gamma
    fs = fluid_model.get_fluid_scalars_from_geodesics(S)
    dens, u, b = (as_device(fs[k]) for k in ('dens', 'u', 'b'))
    bsq = b * b
    beta = u * (fluid_gamma - 1.) / bsq / 0.5
    sigma = bsq / dens
    Theta_e = rlow_rhigh_model(dens, u, beta, r_high=r_high)
    units = fluid_model.get_units(M_bh, mass_scale)
    Ne_in_cgs = units['Ne_unit'] * dens
    B_in_gauss = units['B_unit'] * b
    local_nu = - as_device(fs['kdotu']) * observing_frequency
    em, ab = synchrotron_coefficients(Ne_in_cgs, Theta_e, B_in_gauss, fs['pitch_angle'], local_nu,
                                      invariant=True, rescale_nu=1. / observing_frequency)
    cut = sigma > 100.
    em = torch.where(cut, torch.zeros_like(em), em)
    ab = torch.where(cut, torch.zeros_like(ab), ab)
    return solve_specific_intensity(em, ab, final_dt, units['L_unit'])


def make_image_unfused(fluid_model, camera_inclination=60, camera_distance=1000, mass_scale=1.e26,
                       M_bh=6.2e9 * Msun, r_high=40, observing_frequency=230.e9, fov=20, resolution=160,
                       max_nsteps=10000, max_chunk_bytes=None):
    """The reference's chunked stage-by-stage pipeline (images.py:56-144) with every stage on the GPU."""
    bhspin = fluid_model.bhspin
    s0 = geo.initialize_geodesics_at_camera(bhspin, camera_inclination, camera_distance, -fov / 2., fov / 2.,
                                            resolution)
    npx = s0.shape[0]
    num_pixels_per_chunk = npx + 10
    if max_chunk_bytes is not None:
        num_pixels_per_chunk = int(max_chunk_bytes // 4 // 20 // max_nsteps)     # images.py:69
    else:
        free, _ = torch.cuda.mem_get_info()
        # a trajectory row costs 72 B/ray plus ~10 stage arrays of 8 B; keep a chunk under ~40% of free HBM
        num_pixels_per_chunk = max(1024, min(npx + 10, int(0.4 * free / (160 * 2500))))
    out = np.zeros((0))
    lower = 0
    while lower < npx:
        S, final_dt = geo.geodesic_integrator(max_nsteps, s0[lower:lower + num_pixels_per_chunk], 40, 1e-4, bhspin)
        I_nu = intensity_from_trajectories(fluid_model, S, final_dt, mass_scale, M_bh, r_high, observing_frequency)
        out = np.append(out, np.asarray(I_nu))
        del S, final_dt, I_nu
        lower += num_pixels_per_chunk
    return np.array(out).reshape((resolution, resolution))
