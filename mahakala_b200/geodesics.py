"""Drop-in for ``mahakala.geodesics`` (reference: /root/reference/mahakala/geodesics.py).

Same names, positional order, keyword names and defaults as the reference; the arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI (``include/mahakala_b200.h``).  Arrays returned are
``DeviceArray`` objects (HBM-resident, NumPy protocol via ``__array__``), the analogue of the
``jax.Array`` results of the reference.  There is no CPU fallback.
"""
import numpy as np
import torch

from . import _cabi
from ._device import DeviceArray, as_device, empty, require_gpu, stream_ptr

KERR_SCHILD = 0
KERR_SCHILD_DUAL = 1
KERR_SCHILD_STRICT = 2       # literal IEEE evaluation (integrate_final only), bit-identical to the CPU restatement
PLUGIN_BASE = 16             # MK_METRIC_PLUGIN_BASE: ids of run-time registered spacetimes start here

_METRICS = {"kerr_schild": KERR_SCHILD, "kerr_schild_dual": KERR_SCHILD_DUAL, "kerr_schild_strict": KERR_SCHILD_STRICT}
_active_metric = KERR_SCHILD


def set_metric(name):
    """Select the spacetime plugin used by the module-level functions (the reference's equivalent is
    replacing the module globals ``metric``/``imetric``, geodesics.py:304-305)."""
    global _active_metric
    _active_metric = _METRICS[name] if isinstance(name, str) else int(name)
    return _active_metric


def register_metric(name, cuda_source, params=None):
    """Register a user-defined spacetime at run time and return its metric id.

    ``cuda_source`` is CUDA C++ defining ``struct UserMetric`` (contract: ``csrc/plugin_tu.cuh``): the
    covariant metric evaluated on a generic scalar type, the step-rule radius and the horizon radius.  NVRTC
    compiles it for sm_100a together with the integrate kernel; derivatives come from forward-mode dual
    numbers (the role ``jax.jacfwd`` plays at geodesics.py:305).  ``params`` (up to 7 floats) fill
    ``UserMetric::params[1..7]``; ``params[0]`` is the ``bhspin`` argument of each call.  Select it with
    ``set_metric(name)``.  Compilation itself needs no GPU.
    """
    import ctypes
    import os
    mid = ctypes.c_int(-1)
    log = ctypes.create_string_buffer(8192)
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
    _cabi.call("mk_register_metric", name.encode(), cuda_source.encode(), inc.encode(), ctypes.byref(mid), log, 8192)
    _METRICS[name] = mid.value
    if params is not None:
        set_metric_params(name, params)
    return mid.value


def set_metric_params(name, params):
    """Set ``UserMetric::params[1..7]`` of a registered spacetime."""
    import ctypes
    mid = _METRICS[name] if isinstance(name, str) else int(name)
    vals = [0.0] + [float(q) for q in params]
    vals += [0.0] * (8 - len(vals))
    _cabi.call("mk_metric_set_params", mid, (ctypes.c_double * 8)(*vals[:8]))


def _cos_sin_deg(inclination):
    i = inclination * np.pi / 180          # geodesics.py:205
    return float(np.cos(i)), float(np.sin(i))


# -------------------------------------------------------------------------------------------------
# camera
# -------------------------------------------------------------------------------------------------
def initialize_geodesics_at_camera(bhspin, inclination, distance, fov_lower, fov_upper, pixels_per_side,
                                   camera_type='grid'):
    """geodesics.py:29-55.  Returns s0 (npx, 8) = [t, x, y, z, k^t, k^x, k^y, k^z]."""
    if camera_type.lower() == 'grid' and _active_metric < PLUGIN_BASE:
        # one kernel: image-plane geometry + nullification with the built-in Kerr-Schild metric
        n = int(pixels_per_side)
        ci, si = _cos_sin_deg(inclination)
        s0 = empty((n * n, 8))
        _cabi.call("mk_camera_grid", float(bhspin), ci, si, float(distance), float(fov_lower),
                   float(fov_upper), n, 1, s0, stream_ptr())
        return DeviceArray.wrap(s0)
    # a user-registered spacetime is active: the wavevectors must be null in THAT metric (the reference's
    # initial_condition takes the module-level metric, geodesics.py:225-230)
    grid = get_initial_grid(inclination, distance, fov_lower, fov_upper, pixels_per_side, camera_type)
    if grid is None:
        return None
    return initial_condition(grid[0], grid[1], bhspin)


def get_initial_grid(inclination, distance, fov_lower, fov_upper, spacing, camera_type):
    """geodesics.py:137-201.  Returns (s0_x, s0_v), each (4, npx); wavevectors not yet null."""
    if camera_type.lower() == 'grid':
        n = int(spacing)
        ci, si = _cos_sin_deg(inclination)
        raw = empty((n * n, 8))
        _cabi.call("mk_camera_grid", 0.0, ci, si, float(distance), float(fov_lower), float(fov_upper), n,
                   0, raw, stream_ptr())
        return DeviceArray.wrap(raw[:, :4].T.contiguous()), DeviceArray.wrap(raw[:, 4:].T.contiguous())
    elif camera_type.lower() == 'equator':
        grid_list = np.linspace(fov_lower, fov_upper, 2 * spacing + 1)[1::2]
        s0_x = np.zeros((4, len(grid_list)))
        s0_x[1] = distance
        s0_x[2] = grid_list
        s0_v = np.ones((4, len(grid_list)))
        s0_v[2] = 0
        s0_v[3] = 0
        return s0_x, s0_v
    else:
        print(f'Unexpected camera type "{camera_type}".')
        print('Please choose either "grid" or "equator"')


def _image_points(radius, angle):
    size = np.size(radius)
    x = np.ones(size) * np.cos(angle) * radius     # geodesics.py:117-118
    y = np.ones(size) * np.sin(angle) * radius
    return x, y


def get_camera_pixel(inclination, distance, radius, angle):
    """geodesics.py:107-134.  Returns (s0_x, s0_v), each (4, n)."""
    x, y = _image_points(np.asarray(radius, dtype=np.float64), np.asarray(angle, dtype=np.float64))
    ci, si = _cos_sin_deg(inclination)
    raw = empty((x.size, 8))
    _cabi.call("mk_camera_points", 0.0, ci, si, float(distance), as_device(x), as_device(y), x.size, 0,
               raw, stream_ptr())
    return DeviceArray.wrap(raw[:, :4].T.contiguous()), DeviceArray.wrap(raw[:, 4:].T.contiguous())


def initial_condition(s0_x, s0_v, bhspin):
    """geodesics.py:219-230: make the wavevectors null.  (4, n), (4, n) -> (n, 8)."""
    sx = as_device(s0_x)
    sv = as_device(s0_v)
    n = sx.shape[1]
    s0 = empty((n, 8))
    _cabi.call("mk_initial_condition_metric", _active_metric, float(bhspin), sx, sv, n, s0, stream_ptr())
    return DeviceArray.wrap(s0)


def _camera_pixels_state(inclination, distance, radius, angle, bhspin):
    """get_camera_pixel + initial_condition in one kernel launch (built-in spacetimes; a registered spacetime goes
    through its own initial_condition so that the rays are null in the user's metric)."""
    if _active_metric >= PLUGIN_BASE:
        sx, sv = get_camera_pixel(inclination, distance, radius, angle)
        return as_device(initial_condition(sx, sv, bhspin))
    x, y = _image_points(np.asarray(radius, dtype=np.float64), np.asarray(angle, dtype=np.float64))
    ci, si = _cos_sin_deg(inclination)
    s0 = empty((x.size, 8))
    _cabi.call("mk_camera_points", float(bhspin), ci, si, float(distance), as_device(x), as_device(y),
               x.size, 1, s0, stream_ptr())
    return s0


# -------------------------------------------------------------------------------------------------
# metric utilities
# -------------------------------------------------------------------------------------------------
def _metric_pair(x, bhspin, want):
    xs = as_device(x)
    single = xs.dim() == 1
    pts = xs.reshape(-1, xs.shape[-1])[:, :4].contiguous()
    n = pts.shape[0]
    g = empty((n, 4, 4)) if want == 'g' else None
    gi = empty((n, 4, 4)) if want == 'gi' else None
    _cabi.call("mk_metric", _active_metric, float(bhspin), pts, n, g, gi, stream_ptr())
    out = g if want == 'g' else gi
    out = out[0] if single else out.reshape(tuple(xs.shape[:-1]) + (4, 4))
    return DeviceArray.wrap(out)


def metric(x, bhspin):
    """geodesics.py:88-104: covariant Cartesian Kerr-Schild metric at x (..., 4) -> (..., 4, 4)."""
    return _metric_pair(x, bhspin, 'g')


def imetric(x, bhspin):
    """geodesics.py:339-347: contravariant metric (closed form eta - f l l instead of linalg.inv)."""
    return _metric_pair(x, bhspin, 'gi')


def radius_cal(x, bhspin):
    """geodesics.py:284-291: Kerr-Schild radius of points x (..., >=4)."""
    xs = as_device(x)
    stride = xs.shape[-1]
    flat = xs.reshape(-1, stride)
    n = flat.shape[0]
    r = empty((n,))
    _cabi.call("mk_radius_cal", float(bhspin), flat, n, stride, r, stream_ptr())
    return DeviceArray.wrap(r.reshape(xs.shape[:-1]))


def radius_EH(a_spin):
    """geodesics.py:350-351."""
    return 1 + np.sqrt(1 - a_spin**2)


def rhs(state1, bhspin):
    """geodesics.py:294-309: right-hand side of the geodesic equation; (8,) or (n, 8)."""
    s = as_device(state1)
    single = s.dim() == 1
    s2 = s.reshape(-1, 8)
    out = empty(s2.shape)
    _cabi.call("mk_rhs", _active_metric, float(bhspin), s2, s2.shape[0], out, stream_ptr())
    return DeviceArray.wrap(out[0] if single else out)


_vectorized_rhs = rhs


def RK4_gen(state1, dt, bhspin):
    """geodesics.py:317-336: one RK4 step of a bundle (n, 8) with per-ray dt (n,)."""
    s = as_device(state1)
    d = as_device(dt).reshape(-1)
    out = empty(s.shape)
    _cabi.call("mk_rk4_step", _active_metric, float(bhspin), s, d, s.shape[0], out, stream_ptr())
    return DeviceArray.wrap(out)


# -------------------------------------------------------------------------------------------------
# integrator
# -------------------------------------------------------------------------------------------------
def integrate_final(N, s0, div, tol, bhspin, want_total=False):
    """Final-state mode of the integrate kernel (no trajectory storage).

    Returns ``(final_state (npx, 8), nsteps (npx,) int32, r_last (npx,))`` as device tensors, where
    ``r_last`` is the reference's last-point radius ``radius_cal(S)[argmax(dt) - 1]`` (geodesics.py:370-378).
    """
    s = as_device(s0)
    npx = s.shape[0]
    final = empty((npx, 8))
    nsteps = empty((npx,), dtype=torch.int32)
    r_last = empty((npx,))
    total = torch.zeros(1, dtype=torch.int64, device=s.device) if want_total else None
    _cabi.call("mk_integrate", _active_metric, float(bhspin), int(N), npx, s, float(div), float(tol),
               final, nsteps, r_last, None, None, 0, total, stream_ptr())
    if want_total:
        return final, nsteps, r_last, total
    return final, nsteps, r_last


def integrate_adaptive(N, s0, tol, bhspin, rtol=1e-9, atol=1e-12, cap=0.5):
    """OPTIONAL integrator, not in the reference: embedded Dormand-Prince 5(4) with step-size control instead of
    classical RK4 under the fixed rule ``dt = -(r - r_H)/div`` (geodesics.py:246-269, :317-336).  The step follows the
    local truncation error (``atol + rtol |y|`` per step on position and wavevector), so a ray needs tens of steps
    instead of hundreds to thousands, and a user-registered spacetime is no longer stepped by a Kerr-specific rule.
    Termination and freezing are the reference's: a ray lives while ``tol <= radius - r_H <= 1500``, a step that would
    end outside that range is rejected; ``|h| <= cap (radius - r_H)``.  Runs in the active metric (Kerr-Schild closed
    form, its dual-number twin, or a registered one).

    Returns ``(final_state (npx, 8), nsteps (npx,) int32 accepted, nrejected (npx,) int32, r_last (npx,))`` as device
    tensors; ``r_last`` is the radius of the final state (about ``r_H + tol`` for a captured ray, hundreds of M for an
    escaped one), which classifies rays like ``select_photons_integrator``'s last-point radius does."""
    if _active_metric == KERR_SCHILD_STRICT:
        raise ValueError("the strict (literal IEEE) integrator follows the reference's fixed rule only")
    s = as_device(s0)
    npx = s.shape[0]
    final = empty((npx, 8))
    nsteps = empty((npx,), dtype=torch.int32)
    nrej = empty((npx,), dtype=torch.int32)
    r_last = empty((npx,))
    _cabi.call("mk_integrate_adaptive", _active_metric, float(bhspin), int(N), npx, s, float(rtol), float(atol),
               float(tol), float(cap), final, nsteps, nrej, r_last, stream_ptr())
    return final, nsteps, nrej, r_last


def dump_rows(N, max_steps):
    """Row count of the reference's truncated scan output (geodesics.py:275-281): first all-zero row
    + 2, or N (+2, clipped to N) when there is none or it is row 0."""
    first_zero = max_steps if (1 <= max_steps <= N - 1) else N
    return min(first_zero + 2, N)


def geodesic_integrator(N, s0, div, tol, bhspin):
    """geodesics.py:233-281.  Returns ``(S (nrows, npx, 8), final_dt (nrows, npx))`` in HBM.

    One pass of the persistent integrate kernel dumps the trajectories into a paged store (no row count has
    to be known in advance), then a gather kernel lays them out as the reference does, truncated at the first
    all-zero row + 2 (geodesics.py:275-281) with the frozen state repeated below each ray's last row.  If the
    page pool does not fit next to the padded result the function falls back to two integration passes
    (count, then write).  Use the fused ``images.make_image`` path when the trajectories themselves are not
    needed, and ``integrate_paged`` for bundles whose padded rectangle would not fit in HBM.
    """
    s = as_device(s0)
    npx = s.shape[0]
    N = int(N)
    if npx == 0 or N == 0:       # lax.scan over zero rays / zero iterations: empty outputs of the reference's shapes
        return DeviceArray.wrap(empty((N, npx, 8))), DeviceArray.wrap(empty((N, npx)))
    free, _total = torch.cuda.mem_get_info()
    # typical rays take a few hundred to a few thousand steps; size the pool for min(N, 4096) rows per ray
    want_pages = 2 * (-(-npx // 32)) * (-(-(min(N, 4096) + 1) // TrajectoryStore.PAGE_SLOTS)) + 64
    if want_pages * TrajectoryStore.PAGE_DOUBLES * 8 < 0.35 * free:
        store = TrajectoryStore(npx, N, want_pages, s.device)
        integrate_paged(N, s, div, tol, bhspin, store=store)
        if not store.overflowed:
            nrows = dump_rows(N, int(store.nsteps.max().item()))
            if nrows * npx * 72 > torch.cuda.mem_get_info()[0]:
                raise MemoryError(f"the padded trajectory array ({nrows} rows x {npx} rays) does not fit in HBM: "
                                  "integrate the bundle in chunks (s0[lo:hi]) or keep it paged (integrate_paged)")
            return store.padded()
        del store
    final, nsteps, _ = integrate_final(N, s, div, tol, bhspin)
    nrows = dump_rows(N, int(nsteps.max().item()))
    need = nrows * npx * 72
    free, _total = torch.cuda.mem_get_info()
    if need > free:
        raise MemoryError(f"trajectory dump needs {need / 2**30:.1f} GiB for {nrows} rows x {npx} rays but only "
                          f"{free / 2**30:.1f} GiB are free: integrate the bundle in chunks (s0[lo:hi])")
    S = empty((nrows, npx, 8))
    dt = empty((nrows, npx))
    _cabi.call("mk_integrate", _active_metric, float(bhspin), N, npx, s, float(div), float(tol),
               None, None, None, S, dt, nrows, None, stream_ptr())
    _cabi.call("mk_fill_frozen_rows", S, dt, final, nsteps, npx, nrows, stream_ptr())
    return DeviceArray.wrap(S), DeviceArray.wrap(dt)


class TrajectoryStore:
    """Paged, ragged trajectory dump held in HBM (single integration pass, no padding).

    The reference materialises ``S (nrows, npx, 8)`` padded to the longest ray (and its scan to all ``N``
    iterations); at 1024^2 rays that is hundreds of GB.  Here every warp of the persistent kernel appends
    one 32-lane slot per iteration to its own log of 16-slot pages taken from a pool (fully coalesced
    stores), so memory and write traffic are proportional to the steps actually taken (72 B per ray-step).
    ``padded()`` materialises the reference layout for any subset of rays on demand.
    """
    PAGE_SLOTS = 16
    PAGE_DOUBLES = 16 * 32 * 9

    def __init__(self, npx, N, max_pages, device):
        self.npx, self.N, self.max_pages = int(npx), int(N), int(max_pages)
        self.pages = torch.empty((self.max_pages, self.PAGE_DOUBLES), dtype=torch.float64, device=device)
        self.page_next = torch.empty((self.max_pages,), dtype=torch.int32, device=device)
        self.page_first = torch.empty((self.npx, 2), dtype=torch.int32, device=device)
        self.ctrl = torch.zeros(2, dtype=torch.int32, device=device)       # [page counter, overflow flag]
        self.final = torch.empty((self.npx, 8), dtype=torch.float64, device=device)
        self.nsteps = torch.empty((self.npx,), dtype=torch.int32, device=device)
        self.r_last = torch.empty((self.npx,), dtype=torch.float64, device=device)
        self.total_steps = torch.zeros(1, dtype=torch.int64, device=device)
        self._host_results = None      # set by integrate_paged_host: the per-ray results live in host memory

    @classmethod
    def allocate(cls, npx, N, max_pages=None, mem_fraction=0.6):
        """Reserve the page pool.  Without ``max_pages`` the pool is sized for the worst case but capped at
        ``mem_fraction`` of the currently free HBM (a cfg2-sized dump needs ~42 GB); an undersized pool is
        reported through ``overflowed``, never silently truncated."""
        dev = require_gpu()
        worst = 2 * (-(-int(npx) // 32)) * (-(-(int(N) + 1) // cls.PAGE_SLOTS)) + 148 * 64
        if max_pages is None:
            free, _ = torch.cuda.mem_get_info()
            max_pages = min(worst, int(mem_fraction * free) // (cls.PAGE_DOUBLES * 8))
        return cls(npx, N, max(int(max_pages), 1), dev)

    def reset(self):
        self.ctrl.zero_()
        self.total_steps.zero_()
        self._host_results = None

    def _sync_results(self):
        """After a zero-copy launch the step counts (needed to gather trajectories) are in host memory: bring
        them (and the other per-ray results) back into the store's device arrays, once, on demand."""
        if self._host_results is not None:
            h, self._host_results = self._host_results, None
            self.final.copy_(h["final"]); self.nsteps.copy_(h["nsteps"]); self.r_last.copy_(h["r_last"])

    @property
    def overflowed(self):
        return bool(self.ctrl[1].item())

    @property
    def pages_used(self):
        return int(self.ctrl[0].item())

    def padded(self, rays=None):
        """Reference layout ``(S (nrows, nsel, 8), final_dt (nrows, nsel))`` for ``rays`` (default: all), with
        ``nrows`` following geodesics.py:275-281 for that selection."""
        if self.overflowed:
            raise MemoryError("the page pool overflowed during integration; allocate more pages")
        self._sync_results()
        if rays is None:
            idx = None
            nsel = self.npx
            nmax = int(self.nsteps.max().item()) if nsel else 0
        else:
            idx = as_device(np.asarray(rays, dtype=np.int64), dtype=torch.int64)
            nsel = idx.numel()
            nmax = int(self.nsteps[idx].max().item()) if nsel else 0
        nrows = dump_rows(self.N, nmax)
        S = empty((nrows, nsel, 8))
        dt = empty((nrows, nsel))
        _cabi.call("mk_paged_gather", self.pages, self.page_next, self.page_first, self.nsteps, idx, nsel, nrows,
                   self.N, S, dt, stream_ptr())
        return DeviceArray.wrap(S), DeviceArray.wrap(dt)


def integrate_paged(N, s0, div, tol, bhspin, store=None, queue=None, ray_order=None, results=None, page_id_offset=0,
                    participants=1):
    """Single-pass trajectory dump of a whole bundle into a ``TrajectoryStore`` (see there).

    ``queue`` (a device pointer, possibly in a peer GPU's memory) makes this launch one participant of a multi-GPU
    job that shares ONE dynamic ray queue (``mahakala_b200.multigpu.integrate_distributed``): ``ray_order`` (int32
    tensor) is the order in which rays are handed out, ``results`` = (final, nsteps, r_last, page_first) pointers in
    the gathering GPU's memory, ``page_id_offset`` tags the page numbers with their owner, ``participants`` is the
    number of GPUs pulling from the queue (sizes the chunks warps take from it)."""
    s = as_device(s0)
    npx = s.shape[0]
    if store is None:
        store = TrajectoryStore.allocate(npx, N)
    if store.npx != npx or store.N != int(N):
        raise ValueError("TrajectoryStore was allocated for a different bundle")
    store.reset()                 # a reused store starts from an empty page pool (page counter, overflow flag, total)
    if queue is not None:
        final, nsteps, r_last, page_first = results if results is not None else (store.final, store.nsteps,
                                                                                 store.r_last, store.page_first)
        _cabi.call("mk_integrate_shared", _active_metric, float(bhspin), int(N), npx, s, float(div), float(tol),
                   final, nsteps, r_last, store.pages, store.page_next, page_first, store.ctrl[0:1], store.max_pages,
                   store.ctrl[1:2], store.total_steps, queue, ray_order, int(page_id_offset), int(participants),
                   stream_ptr())
        return store
    _cabi.call("mk_integrate_paged", _active_metric, float(bhspin), int(N), npx, s, float(div), float(tol),
               store.final, store.nsteps, store.r_last, store.pages, store.page_next, store.page_first,
               store.ctrl[0:1], store.max_pages, store.ctrl[1:2], store.total_steps, stream_ptr())
    return store


_stream_pool = {}


def _side_streams(dev):
    if dev not in _stream_pool:
        _stream_pool[dev] = [torch.cuda.Stream(device=dev) for _ in range(4)]      # copy-in, 2 x compute, copy-out
    return _stream_pool[dev]


def integrate_paged_host(N, s0_host, div, tol, bhspin, store, host_out, ray_order=None):
    """Host-to-host ``integrate_paged`` without staging copies (zero-copy).  ``ray_order`` (device int32 tensor):
    optional order in which the rays are handed out (longest first shortens the tail of the launch).

    ``s0_host`` (npx, 8) and ``host_out`` = {``final`` (npx, 8), ``nsteps`` (npx,) int32, ``r_last`` (npx,)} are
    PINNED CPU tensors.  Under unified addressing pinned host memory is mapped into the device address space, so
    the persistent kernel itself pulls each ray's 64 B initial state over PCIe when a lane picks the ray up and
    stores the 76 B of per-ray results straight into host memory when the ray freezes.  The transfers ride along
    inside ONE launch (147 MB per 18 ms at cfg2, a fraction of the PCIe bandwidth; the refill latency of ~2 us is
    hidden by the other warps), so there is a single drain phase instead of one per chunk: measured on B200,
    cfg2, host to host: 18.8 ms against 22.3 ms for the 4-chunk copy/compute pipeline (``integrate_paged_streamed``)
    and 18.4 ms for the device-resident launch.  The trajectories stay in ``store`` in HBM.
    Returns after the kernel has completed (results are visible to the host).
    """
    require_gpu()
    npx = s0_host.shape[0]
    if store.npx != npx or store.N != int(N):
        raise ValueError("TrajectoryStore was allocated for a different bundle")
    tensors = (s0_host, host_out["final"], host_out["nsteps"], host_out["r_last"])
    if not all(t.device.type == "cpu" and t.is_pinned() and t.is_contiguous() for t in tensors):
        raise ValueError("integrate_paged_host needs pinned, contiguous CPU tensors")
    if (s0_host.dtype, host_out["final"].dtype, host_out["nsteps"].dtype, host_out["r_last"].dtype) != \
            (torch.float64, torch.float64, torch.int32, torch.float64):
        raise ValueError("s0 / final / r_last must be float64 and nsteps int32")
    store.reset()
    if ray_order is not None:
        if not hasattr(store, "_queue"):
            store._queue = torch.zeros(64, dtype=torch.int32, device=store.pages.device)
        store._queue.zero_()
        _cabi.call("mk_integrate_shared", _active_metric, float(bhspin), int(N), npx, s0_host, float(div), float(tol),
                   host_out["final"], host_out["nsteps"], host_out["r_last"], store.pages, store.page_next,
                   store.page_first, store.ctrl[0:1], store.max_pages, store.ctrl[1:2], store.total_steps,
                   store._queue, ray_order, 0, 1, stream_ptr())
    else:
        _cabi.call("mk_integrate_paged", _active_metric, float(bhspin), int(N), npx, s0_host, float(div), float(tol),
                   host_out["final"], host_out["nsteps"], host_out["r_last"], store.pages, store.page_next,
                   store.page_first, store.ctrl[0:1], store.max_pages, store.ctrl[1:2], store.total_steps, stream_ptr())
    torch.cuda.current_stream().synchronize()
    store._host_results = host_out
    return store


def integrate_paged_streamed(N, s0_host, div, tol, bhspin, store, host_out, chunks=4):
    """Host-to-host variant of ``integrate_paged`` that overlaps explicit PCIe copies with the kernel (kept for
    comparison with the zero-copy ``integrate_paged_host``, which is faster: every chunk pays its own drain phase).

    ``s0_host`` is a pinned CPU tensor (npx, 8); ``host_out`` a dict of pinned CPU tensors ``final`` (npx, 8),
    ``nsteps`` (npx,) int32 and ``r_last`` (npx,) that receive the per-ray results (the trajectories stay in
    ``store`` in HBM).  The bundle is cut into ``chunks`` ray ranges: the upload of range i+1, the persistent
    kernel on range i (two alternating compute streams, so that the tail of one launch overlaps the head of
    the next) and the download of range i-1 run concurrently.  All ranges share the store's page pool.
    Returns after the last download has completed.
    """
    dev = require_gpu()
    npx = s0_host.shape[0]
    if store.npx != npx or store.N != int(N):
        raise ValueError("TrajectoryStore was allocated for a different bundle")
    if not hasattr(store, "s0_dev") or store.s0_dev.shape[0] != npx:
        store.s0_dev = torch.empty((npx, 8), dtype=torch.float64, device=dev)
    store.reset()
    cin, c0, c1, cout = _side_streams(dev)
    cur = torch.cuda.current_stream()
    for st in (cin, c0, c1, cout):
        st.wait_stream(cur)
    bounds = [npx * k // chunks for k in range(chunks + 1)]
    for k in range(chunks):
        lo, hi = bounds[k], bounds[k + 1]
        if hi <= lo:
            continue
        with torch.cuda.stream(cin):
            store.s0_dev[lo:hi].copy_(s0_host[lo:hi], non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(cin)
        comp = c0 if k % 2 == 0 else c1
        with torch.cuda.stream(comp):
            comp.wait_event(ev_in)
            _cabi.call("mk_integrate_paged", _active_metric, float(bhspin), int(N), hi - lo, store.s0_dev[lo:hi],
                       float(div), float(tol), store.final[lo:hi], store.nsteps[lo:hi], store.r_last[lo:hi],
                       store.pages, store.page_next, store.page_first[lo:hi], store.ctrl[0:1], store.max_pages,
                       store.ctrl[1:2], store.total_steps, comp.cuda_stream)
            ev_done = torch.cuda.Event()
            ev_done.record(comp)
        with torch.cuda.stream(cout):
            cout.wait_event(ev_done)
            host_out["final"][lo:hi].copy_(store.final[lo:hi], non_blocking=True)
            host_out["nsteps"][lo:hi].copy_(store.nsteps[lo:hi], non_blocking=True)
            host_out["r_last"][lo:hi].copy_(store.r_last[lo:hi], non_blocking=True)
    cout.synchronize()
    cur.wait_stream(c0)
    cur.wait_stream(c1)
    return store


# -------------------------------------------------------------------------------------------------
# shadow finder
# -------------------------------------------------------------------------------------------------
def select_photons_integrator(inc, angle, radius, bhspin, distance=1000, max_steps=2000):
    """geodesics.py:354-378: last-point radius of each photon (used to classify captured / escaped)."""
    s0 = _camera_pixels_state(inc, distance, radius, angle, bhspin)
    _, _, r_last = integrate_final(max_steps, s0, 40, 1e-2, bhspin)
    return DeviceArray.wrap(r_last)


def select_photons_adaptive(inc, angle, radius, bhspin, distance=1000, max_steps=2000):
    """``select_photons_integrator`` with the optional adaptive integrator (extension, see ``integrate_adaptive``): the
    radius at which each photon ended, in the active spacetime -- no Kerr-specific step rule involved."""
    s0 = _camera_pixels_state(inc, distance, radius, angle, bhspin)
    return DeviceArray.wrap(integrate_adaptive(max_steps, s0, 1e-2, bhspin)[3])


def find_shadow_bisection(bhspin, inc, num_angles, max_steps=2000, error_allowed=0.001, max_it=40):
    """geodesics.py:381-402."""
    angles = np.arange(num_angles) / num_angles * 2. * np.pi
    radii = find_shadow_bisection_angles(bhspin, inc, angles, max_it=max_it, error_allowed=error_allowed,
                                         max_steps=max_steps)
    radii = np.append(radii, radii[0])
    angles = np.append(angles, angles[0])
    return angles, radii


def _bisection_iterations(lo, hi, error_allowed, max_it):
    """How many times the reference's loop (geodesics.py:420-434) runs: every angle starts from the same bracket and
    halves it each time, so ``max(error)`` follows one sequence.  Returns None when a width comes within rounding of
    ``error_allowed`` (the two halves of a bracket can differ in the last bit; then the host loop decides)."""
    n = 0
    width = float(hi) - float(lo)
    while width > error_allowed and n < max_it:
        if abs(width - error_allowed) <= 1e-9 * error_allowed:
            return None
        width = width / 2
        n += 1
    return None if abs(width - error_allowed) <= 1e-9 * error_allowed else n


def find_shadow_bisection_angles(bhspin, inc, angles, max_steps=2000, error_allowed=0.001, max_it=40):
    """geodesics.py:405-435: bisection on the image-plane radius of the shadow edge, per angle.

    With the built-in spacetime the whole bisection runs in ONE kernel launch (``mk_shadow_bisection``: a lane owns an
    angle and iterates ray -> captured? -> halve); a user-registered spacetime, or a bracket that stalls (NaN
    classifier radius), goes through the reference's own iteration-by-iteration loop below."""
    require_gpu()
    angles = np.asarray(angles, dtype=np.float64)
    if angles.size == 0:
        return np.zeros(angles.shape)
    n_iter = _bisection_iterations(0.5, 10, error_allowed, max_it)
    if _active_metric == KERR_SCHILD and n_iter is not None:
        flat = angles.reshape(-1)
        ci, si = _cos_sin_deg(inc)
        inner, outer = empty((flat.size,)), empty((flat.size,))
        _cabi.call("mk_shadow_bisection", float(bhspin), ci, si, 1000.0, as_device(np.cos(flat)), as_device(np.sin(flat)),
                   flat.size, int(max_steps), 40.0, 1e-2, n_iter, 0.5, 10.0, 100.0, inner, outer, stream_ptr())
        inner_h, outer_h = inner.cpu().numpy(), outer.cpu().numpy()
        if not (np.max(outer_h - inner_h) > error_allowed and n_iter < max_it):      # the loop would have stopped here too
            return inner_h.reshape(angles.shape)
    return _find_shadow_bisection_angles_host(bhspin, inc, angles, max_steps, error_allowed, max_it)


def _find_shadow_bisection_angles_host(bhspin, inc, angles, max_steps=2000, error_allowed=0.001, max_it=40):
    """The reference's loop, one bundle launch per bisection iteration (geodesics.py:417-435): brackets start at
    [0.5, 10]; a mid-radius ray whose last-point radius is below 100 fell in (edge further out), at or above 100 it
    got away; a NaN radius moves neither end."""
    require_gpu()
    angles = np.asarray(angles, dtype=np.float64)
    lo = np.full(angles.shape, 0.5)
    hi = np.full(angles.shape, 10.0)
    for _ in range(max_it):
        if not np.max(hi - lo) > error_allowed:
            break
        mid = (hi - lo) / 2 + lo
        r_end = np.asarray(select_photons_integrator(inc, angles, mid, bhspin, max_steps=max_steps))
        lo = np.where(r_end < 100, mid, lo)
        hi = np.where(r_end >= 100, mid, hi)
    return lo
