"""GPU parity of sampling, thermodynamics, synchrotron, transfer and the fused render against the oracle."""
import numpy as np
import pytest

from helpers import M_BH, MASS_SCALE, device_model, oracle_model, snapshot_arrays

pytestmark = pytest.mark.gpu
A = 0.94


@pytest.fixture(scope="module")
def setup(built):
    from oracle import c_oracle, mahakala_oracle as onp
    arr = snapshot_arrays(ncells=32, block=16, extent=16.0)
    om = oracle_model(arr, A)
    dm = device_model(arr, A)
    s0 = onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 10)
    S, dt = c_oracle.geodesic_integrator(10000, s0, 40, 1e-4, A)
    return dict(arr=arr, om=om, dm=dm, s0=s0, S=S, dt=dt)


def test_ghost_fill_matches_oracle(setup):
    assert np.array_equal(setup["dm"].all_meshblocks, setup["om"].all_meshblocks)
    # the device snapshot was built by the fused ghost-fill + repack kernel from the interior arrays only
    assert setup["dm"]._ghost_fill == "device"
    assert np.array_equal(np.asarray(setup["dm"].device_meshblocks()), setup["om"].all_meshblocks)


def test_device_ghost_fill_variants(setup):
    """Device ghost fill (mk_snapshot_create_from_interiors) against the oracle's host fill: float32 and float64
    inputs, permuted variable order, non-float32-representable data (auto storage must fall back to float64),
    and identity with the host-filled upload route."""
    arr, om = setup["arr"], setup["om"]
    ref = om.all_meshblocks
    m32 = device_model(dict(arr, uov=arr["uov"].astype(np.float32), B=arr["B"].astype(np.float32)), A)
    assert np.array_equal(np.asarray(m32.device_meshblocks()), ref) and m32.storage == "f32"
    assert m32._uov.dtype == np.float32
    mh = device_model(arr, A, ghost_fill="host")
    assert np.array_equal(np.asarray(mh.device_meshblocks()), ref) and mh.storage == "f32"
    m64 = device_model(arr, A, storage="f64")
    assert np.array_equal(np.asarray(m64.device_meshblocks()), ref) and m64.storage == "f64"
    # permuted file order: eint first, dens last among the hydro variables
    perm = [4, 1, 2, 3, 0]
    names = [arr["VariableNames"][q] for q in perm] + list(arr["VariableNames"][5:])
    mp = device_model(dict(arr, uov=arr["uov"][perm], VariableNames=names), A)
    got = np.asarray(mp.device_meshblocks())
    assert np.array_equal(got[:, :5], ref[:, perm]) and np.array_equal(got[:, 5:], ref[:, 5:])
    S = setup["S"]
    for k, v in setup["dm"].get_prims_from_geodesics(S).items():
        assert np.array_equal(np.asarray(v), np.asarray(mp.get_prims_from_geodesics(S)[k])), k
    # values that are not float32-representable: 'auto' must store float64 and reproduce them exactly
    lossy = dict(arr, uov=arr["uov"] * (1.0 + 1e-11))
    ml = device_model(lossy, A)
    assert np.array_equal(np.asarray(ml.device_meshblocks()), oracle_model(lossy, A).all_meshblocks)
    assert ml.storage == "f64"
    for m in (m32, mh, m64, mp, ml):
        m.release()


def test_sample_prims_and_scalars(setup):
    from oracle import c_oracle
    om, dm, S = setup["om"], setup["dm"], setup["S"]
    ref_p = c_oracle.sample(om, S, mode="prims")
    ref_s = c_oracle.sample(om, S, mode="scalars")
    got_p = dm.get_prims_from_geodesics(S)
    got_s = dm.get_fluid_scalars_from_geodesics(S)
    indom = ref_p["dens"] != 0
    assert indom.sum() > 1000 and (~indom).sum() > 1000
    for k in ref_p:
        g = np.asarray(got_p[k])
        assert g.shape == S.shape[:2]
        assert np.array_equal(g == 0, ref_p[k] == 0), k
        scale = np.abs(ref_p[k]).max()
        assert np.abs(g - ref_p[k]).max() <= 1e-14 * scale, k
    for k in ("dens", "u", "kdotu", "b"):
        g = np.asarray(got_s[k])
        assert np.abs(g - ref_s[k]).max() <= 1e-11 * np.abs(ref_s[k]).max(), k
    # pitch angle: acos amplifies rounding near 0 and pi; compare cosines
    assert np.abs(np.cos(np.asarray(got_s["pitch_angle"])) - np.cos(ref_s["pitch_angle"])).max() < 1e-9
    # the NumPy oracle agrees with the C oracle on a thin slice (pins the C sampling code)
    sl = S[40:60]
    np_s = om.get_fluid_scalars_from_geodesics(sl)
    for k in ("dens", "u", "kdotu", "b"):
        assert np.allclose(np_s[k], ref_s[k][40:60], rtol=1e-10, atol=1e-14 * np.abs(ref_s[k]).max()), k


def test_lookup_and_storage_variants_identical(setup):
    arr, S = setup["arr"], setup["S"][::7]
    base = {k: np.asarray(v) for k, v in setup["dm"].get_prims_from_geodesics(S).items()}
    assert setup["dm"].storage == "f32" and setup["dm"].lookup == "grid"      # synthetic data are f32-exact
    for kw in (dict(storage="f64"), dict(lookup="scan"), dict(storage="f64", lookup="scan")):
        m = device_model(arr, A, **kw)
        got = m.get_prims_from_geodesics(S)
        for k in base:
            assert np.array_equal(np.asarray(got[k]), base[k]), (kw, k)
        m.release()


def test_points_on_faces_outside_and_nan(setup):
    from oracle import c_oracle
    om, dm = setup["om"], setup["dm"]
    f = om.x1f[0]
    pts = []
    for x in (f[0], f[1], f[-1], -16.0, 16.0, 0.0, 15.999999999999998, 16.000000000000004, -15.75, 1e300, np.nan):
        for y in (0.0, -16.0, 16.0, 2.0):
            pts.append([0.0, x, y, 0.5, 1.0, 0.3, -0.2, 0.1])
    S = np.array(pts)[None]
    ref = c_oracle.sample(om, S, mode="prims")
    got = dm.get_prims_from_geodesics(S)
    for k in ref:
        assert np.allclose(np.asarray(got[k]), ref[k], rtol=1e-13, atol=1e-14 * np.abs(ref[k]).max()), k


def test_thermo_synchrotron_transfer_elementwise(built):
    import mahakala_b200 as ma
    from mahakala_b200.electrons import rlow_rhigh_model
    from mahakala_b200 import transfer
    from oracle import mahakala_oracle as onp
    rng = np.random.default_rng(5)
    n = 4000
    dens = np.exp(rng.normal(0, 2, n)); u = dens * np.exp(rng.normal(-1, 1, n)); beta = np.exp(rng.normal(0, 2, n))
    dens[:5] = 0; beta[5:9] = np.inf; u[9] = np.nan
    th = np.asarray(rlow_rhigh_model(dens, u, beta, r_high=40))
    th_ref = onp.rlow_rhigh_model(dens, u, beta, r_high=40)
    assert np.allclose(th, th_ref, rtol=1e-14, equal_nan=True)
    Ne = np.exp(rng.normal(10, 2, n)); Th = np.exp(rng.normal(1, 1.5, n)); B = np.exp(rng.normal(1, 1, n))
    pitch = rng.uniform(0, np.pi, n); nu = 230e9 * np.exp(rng.normal(0, 0.5, n))
    Th[:20] = 0.1; B[20:25] = 0; Ne[25:28] = np.nan; nu[28:40] *= 1e-7; pitch[40:44] = 0.0
    for inv, resc in ((True, 1 / 230e9), (True, 1.0), (False, 1.0)):
        em, ab = transfer.synchrotron_coefficients(Ne, Th, B, pitch, nu, invariant=inv, rescale_nu=resc)
        em_r, ab_r = onp.synchrotron_coefficients(Ne, Th, B, pitch, nu, invariant=inv, rescale_nu=resc)
        assert np.array_equal(np.asarray(em) == 0, em_r == 0)
        assert np.allclose(np.asarray(em), em_r, rtol=1e-12, atol=0)
        assert np.allclose(np.asarray(ab), ab_r, rtol=1e-12, atol=0)
    # scalar broadcasting as in the reference's elementwise arithmetic
    em, ab = ma.synchrotron_coefficients(Ne, 10.0, B, np.pi / 3, 230e9)
    em_r, ab_r = onp.synchrotron_coefficients(Ne, 10.0, B, np.pi / 3, 230e9)
    assert np.allclose(np.asarray(em), em_r, rtol=1e-12) and np.allclose(np.asarray(ab), ab_r, rtol=1e-12)
    # transfer scans
    nrows, npx = 57, 33
    emm = np.exp(rng.normal(-3, 1, (nrows, npx))); abb = np.exp(rng.normal(-2, 1, (nrows, npx)))
    dt = -np.abs(rng.normal(0.3, 0.1, (nrows, npx))); dt[40:, ::3] = 0
    I = np.asarray(ma.solve_specific_intensity(emm, abb, dt, 2.5))
    I_r = onp.solve_specific_intensity(emm, abb, dt, 2.5)
    assert np.array_equal(I, I_r)                       # same operation order -> bit-exact
    I2, dI = ma.solve_specific_intensity(emm, abb, dt, 2.5, dIs=True)
    I2_r, dI_r = onp.solve_specific_intensity(emm, abb, dt, 2.5, dIs=True)
    assert np.array_equal(np.asarray(dI), dI_r) and np.array_equal(np.asarray(I2), I2_r)
    att = np.asarray(ma.solve_attenuated_emissivity(emm, abb, dt, 2.5))
    att_r = onp.solve_attenuated_emissivity(emm, abb, dt, 2.5)
    assert att.shape == att_r.shape == (nrows - 1, npx)
    assert np.allclose(att, att_r, rtol=1e-13, atol=0)
    # degenerate shapes
    assert np.asarray(ma.solve_specific_intensity(emm[:1], abb[:1], dt[:1], 1.0)).tolist() == [0.0] * npx


def _image_errors(img, ref):
    scale = np.abs(ref).max()
    per_px = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-6 * scale)
    flux = abs(img.sum() - ref.sum()) / abs(ref.sum())
    return per_px.max(), flux


def test_fused_and_unfused_images_match_oracle(setup):
    """north_star tolerances: per-pixel intensity 1e-6 relative, total flux 1e-8."""
    from mahakala_b200 import images
    from oracle import c_oracle, mahakala_oracle as onp
    om, dm = setup["om"], setup["dm"]
    res = 24
    s0 = onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, res)
    units = om.get_units(M_BH, MASS_SCALE)
    ref, nsteps, nin = c_oracle.render(om, s0, units, [230e9], r_high=40)
    ref = ref[0].reshape(res, res)
    assert ref.max() > 0 and nin > 0
    img = images.make_image(dm, resolution=res)
    assert img.shape == (res, res) and isinstance(img, np.ndarray)
    e_px, e_flux = _image_errors(img, ref)
    assert e_px < 1e-6 and e_flux < 1e-8, (e_px, e_flux)
    img_u = images.make_image_unfused(dm, resolution=res)
    e_px, e_flux = _image_errors(img_u, ref)
    assert e_px < 1e-6 and e_flux < 1e-8, (e_px, e_flux)
    # NumPy oracle end to end on a coarser image (pins the C render chain)
    s0c = onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 8)
    ref_np = onp.make_image(om, resolution=8, integrator=c_oracle.geodesic_integrator)
    ref_c, _, _ = c_oracle.render(om, s0c, units, [230e9])
    assert np.allclose(ref_c[0].reshape(8, 8), ref_np, rtol=1e-9, atol=1e-12 * ref_np.max())


def test_fused_multifrequency_and_explicit_rays(setup):
    from mahakala_b200 import images
    from oracle import c_oracle, mahakala_oracle as onp
    om, dm = setup["om"], setup["dm"]
    res = 12
    nus = [43e9, 86e9, 230e9, 345e9, 690e9]
    s0 = onp.initialize_geodesics_at_camera(A, 30, 1000, -9, 9, res)
    units = om.get_units(M_BH, MASS_SCALE)
    ref, nsteps, nin = c_oracle.render(om, s0, units, nus)
    img, counters = images.render(dm, camera_inclination=30, fov=18, resolution=res, observing_frequencies=nus,
                                  want_counters=True)
    img = np.asarray(img.cpu())
    assert img.shape == (5, res * res)
    for f in range(5):
        e_px, e_flux = _image_errors(img[f], ref[f])
        assert e_px < 1e-6 and e_flux < 1e-8, (f, e_px, e_flux)
    assert int(counters[1]) == nin
    assert abs(int(counters[0]) - int(nsteps.sum())) <= 4 * res * res
    # explicit rays (ragged count, not a multiple of 32)
    sub = s0[5:5 + 77]
    img2 = np.asarray(images.render(dm, s0=sub, observing_frequencies=[230e9]).cpu())
    assert np.allclose(img2[0], ref[2][5:5 + 77], rtol=1e-6, atol=1e-9 * ref[2].max())


def test_distributed_render_two_ranks(built):
    """The N > 1 code paths against the single-GPU results, bit for bit (scripts/multigpu_check.py): shared tile queue
    + in-kernel gather and static sharding for the render, shared ray queue + gathered per-ray results + rank-tagged
    page locator for the integration.  With two GPUs: one rank per GPU over NCCL / NVLink.  On a single-GPU box: two
    ranks on GPU 0 (gloo for the plumbing; queue counter and result buffers are CUDA-IPC mappings, the same kernel
    code path), so the split-vs-single parity is checked wherever the GPU tests run."""
    import os
    import subprocess
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    if torch.cuda.device_count() < 2:
        env["MK_SAME_GPU"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "scripts", "multigpu_check.py"), "128", "32"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mode=queue: identical to single-GPU image: True" in r.stdout
    assert "mode=static: identical to single-GPU image: True" in r.stdout
    assert "long-patch pipeline" in r.stdout and "quick patch order" in r.stdout
    assert r.stdout.count("identical to the fused single-GPU image: True") == 2
    assert "shared ray queue: identical to single-GPU integration: True" in r.stdout
    assert "identical to the single-GPU dump: True" in r.stdout


def test_two_level_mesh_sampling(built):
    """Level-aware O(1) block lookup + sampling on a refined mesh, against the oracle's linear scan."""
    from helpers import two_level_mesh
    from oracle import c_oracle, mahakala_oracle as onp
    arr, expected = two_level_mesh(n=8)
    dm = device_model(arr, A)
    om = oracle_model(arr, A)
    om.all_meshblocks = expected                      # brute-force ghost cells (tests/helpers.py)
    assert np.abs(dm.all_meshblocks - expected).max() < 1e-15
    # refinement boundaries on the device: injection from the coarse neighbour and the mean of the 8 fine cells
    # are evaluated in the host fill's operation order -> bit-identical (the means are not float32 numbers, so
    # 'auto' storage has to choose float64 cells)
    assert np.array_equal(np.asarray(dm.device_meshblocks()), dm.all_meshblocks) and dm.storage == "f64"
    rng = np.random.default_rng(2)
    pts = rng.uniform(-8.5, 8.5, (4000, 3))
    faces = np.array([-8.0, -4.0, 0.0, 2.0, 4.0, 6.0, 8.0])        # points exactly on coarse and fine faces
    pts[:600] = rng.choice(faces, (600, 3))
    pts[600:900, 0] = rng.choice(faces, 300)
    S = np.concatenate([np.zeros((4000, 1)), pts, np.ones((4000, 1)), rng.normal(0, 0.5, (4000, 3))], 1)[None]
    ref = c_oracle.sample(om, S, mode="prims")
    for lookup in ("grid", "scan"):
        m = device_model(arr, A, lookup=lookup, storage="f64")
        got = m.get_prims_from_geodesics(S)
        for k in ref:
            g = np.asarray(got[k])
            assert np.array_equal(g == 0, ref[k] == 0), (lookup, k)
            assert np.allclose(g, ref[k], rtol=1e-13, atol=1e-14 * np.abs(ref[k]).max()), (lookup, k)
        m.release()
    mb_ref = om._meshblock_indices(S)[0]
    assert set(np.unique(mb_ref)) == set(range(-1, 15))            # every block (and "outside") is exercised


@pytest.mark.parametrize("exclusive", [0, 2, 4])
def test_long_patch_pipeline_is_bit_identical(setup, exclusive, monkeypatch):
    """mk_render_long (producer warp = geodesic, consumer warps = sample + emission through a shared-memory ring) against
    the fused kernel: the same device functions on the same operands in the same order, so every pixel, step count and
    counter must be IDENTICAL -- whole frame through the pipeline kernel alone, a learned order with the head of it on
    the pipeline kernel (two concurrent launches), an odd resolution (NaN centre ray, partial patches), the iteration
    cap, explicit rays, f32 cells and the analytic torus; 128-thread CTAs (one patch) and 512-thread CTAs (four patches,
    an SM to themselves)."""
    import torch
    import mahakala_b200 as ma
    from mahakala_b200 import images
    from mahakala_b200.grmhd import AnalyticTorusFluidModel
    monkeypatch.setattr(images, "_LONG_EXCLUSIVE", int(exclusive))
    dm = setup["dm"]
    for kw in (dict(resolution=40), dict(resolution=21), dict(resolution=24, max_nsteps=300),
               dict(resolution=32, camera_inclination=17, fov=14.0)):
        ref, cref = images.render(dm, want_counters=True, patch_order="centre_out", **kw)
        npatch = (-(-kw["resolution"] // 4)) * (-(-kw["resolution"] // 8))
        order = images.centre_out_patch_order(kw["resolution"], ref.device)
        for n_long in (npatch, npatch // 3):
            img, c = images.render(dm, want_counters=True, patch_order=order, long_patches=n_long, **kw)
            assert torch.equal(img, ref), (kw, n_long, float((img - ref).abs().max()))
            assert torch.equal(c, cref), (kw, n_long, c, cref)
    # several frequencies: the consumers evaluate all of them per sample and fold each into its own (I_f, T_f)
    for nus in ((86e9, 230e9), (43e9, 86e9, 230e9, 345e9, 690e9), (43e9, 86e9, 130e9, 230e9, 345e9, 460e9, 690e9, 870e9)):
        kw = dict(resolution=24, observing_frequencies=nus, mass_scale=2e24)
        ref, cref = images.render(dm, want_counters=True, patch_order="centre_out", **kw)
        order = images.centre_out_patch_order(24, ref.device)
        for n_long in (18, 7):
            img, c = images.render(dm, want_counters=True, patch_order=order, long_patches=n_long, **kw)
            assert torch.equal(img, ref) and torch.equal(c, cref), (len(nus), n_long, float((img - ref).abs().max()))
    # learned order: 'auto' sends the photon-ring patches to the pipeline kernel
    ref = images.render(dm, resolution=48)
    images.learn_patch_order(A, resolution=48)
    try:
        key = next(iter(images._learned_lengths))
        assert images.long_patch_count(images._learned_lengths[key]) > 0
        assert torch.equal(images.render(dm, resolution=48), ref)
        assert np.array_equal(images.make_image(dm, resolution=48).reshape(-1), np.asarray(ref.cpu())[0])
    finally:
        images.forget_patch_orders()
    # the coarse, capped pre-pass used in front of a cold multi-GPU frame: finds long patches, same pixels
    order, n_long = images.quick_patch_order(A, resolution=48)
    try:
        assert n_long > 0 and sorted(order.tolist()) == list(range(12 * 6))
        assert torch.equal(images.render(dm, resolution=48, long_patches=n_long), ref)
    finally:
        images.forget_patch_orders()
    # explicit rays (ragged), f32 cells, analytic torus
    s0 = ma.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 14)[3:3 + 150]
    ref = images.render(dm, s0=s0)
    assert torch.equal(_render_long_only(dm, exclusive, s0=s0), ref)
    d32 = device_model(setup["arr"], A, storage="f32")
    assert torch.equal(_render_long_only(d32, exclusive, resolution=24), images.render(d32, resolution=24))
    tm = AnalyticTorusFluidModel(A)
    assert torch.equal(_render_long_only(tm, exclusive, resolution=24), images.render(tm, resolution=24))


def _render_long_only(model, exclusive, **kw):
    """every patch through mk_render_long (explicit rays have no patch order: call the ABI entry directly)"""
    import ctypes
    import torch
    from mahakala_b200 import _cabi, images
    from mahakala_b200._device import as_device, stream_ptr
    snap = model.snapshot()
    P, _ = images._params_for(model, M_BH, MASS_SCALE, 40)
    nu = (ctypes.c_double * 8)(*([230e9] * 8))
    s0 = kw.get("s0")
    res = 0 if s0 is not None else int(kw["resolution"])
    s0d = as_device(s0) if s0 is not None else None
    npx = s0d.shape[0] if s0 is not None else res * res
    img = torch.full((1, npx), -1.0, dtype=torch.float64, device="cuda")
    i = np.pi / 3
    _cabi.call("mk_render_long", float(model.bhspin), float(np.cos(i)), float(np.sin(i)), 1000.0, -10.0, 10.0, res, s0d, npx,
               10000, 40.0, 1e-4, snap, P, 1, nu, img, None, None, None, None, 0, -1, 1, None, int(exclusive), 0, stream_ptr())
    torch.cuda.synchronize()
    return img


def test_odd_resolution_dead_centre_pixel_and_cap(setup):
    """Reference quirk: with an odd resolution the centre pixel has a zero direction -> NaN
    wavevector; such a ray never moves and contributes zero intensity.  Also the iteration cap N."""
    import mahakala_b200 as ma
    from mahakala_b200 import images
    from oracle import c_oracle, mahakala_oracle as onp
    om, dm = setup["om"], setup["dm"]
    res = 9
    s0 = np.asarray(ma.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, res))
    ref_s0 = onp.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, res)
    centre = (res // 2) * res + res // 2
    assert np.isnan(s0[centre, 5:]).all() and np.isnan(ref_s0[centre, 5:]).all()
    assert np.array_equal(np.isnan(s0), np.isnan(ref_s0))
    units = om.get_units(M_BH, MASS_SCALE)
    ref, nsteps, _ = c_oracle.render(om, ref_s0, units, [230e9])
    img = images.make_image(dm, resolution=res)
    assert img.shape == (res, res) and img[res // 2, res // 2] == 0.0 and ref[0][centre] == 0.0
    assert np.allclose(img.reshape(-1), ref[0], rtol=1e-6, atol=1e-12 * ref[0].max())
    # iteration cap: max_nsteps smaller than the natural ray length truncates the transfer integral identically
    ref_cap, _, _ = c_oracle.render(om, ref_s0, units, [230e9], N=260)
    img_cap = images.make_image(dm, resolution=res, max_nsteps=260)
    assert np.allclose(img_cap.reshape(-1), ref_cap[0], rtol=1e-6, atol=1e-12 * ref[0].max())
    assert not np.allclose(img_cap, img)


def test_against_frozen_oracle_fixtures(built):
    """The CUDA path against committed golden vectors (tests/golden/oracle_fixtures.npz)."""
    import os
    import mahakala_b200 as ma
    from mahakala_b200 import geodesics as geo, images
    from mahakala_b200.grmhd import AnalyticTorusFluidModel
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_fixtures.npz"))
    s0 = np.asarray(ma.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 12))
    assert np.allclose(s0, g["geo_s0"], rtol=1e-14, atol=0)
    f, n, rl = geo.integrate_final(2000, s0, 40, 1e-2, A)
    cap = g["geo_r_last"] < 100
    assert np.array_equal(np.asarray(rl.cpu()) < 100, cap)
    assert np.array_equal(np.asarray(n.cpu())[~cap], g["geo_nsteps"][~cap])
    assert np.allclose(np.asarray(f.cpu())[~cap], g["geo_final"][~cap], rtol=1e-9, atol=1e-9)
    arr = snapshot_arrays(ncells=32, block=16, extent=16.0)
    dm = device_model(arr, A)
    img, cnt = images.render(dm, resolution=12, observing_frequencies=(230e9, 345e9), want_counters=True)
    ref = g["img_230_345"]
    assert np.allclose(np.asarray(img.cpu()), ref, rtol=1e-6, atol=1e-12 * ref.max())
    assert int(cnt[1]) == int(g["img_in_domain_samples"])
    timg = images.render(AnalyticTorusFluidModel(A), resolution=12)
    assert np.allclose(np.asarray(timg.cpu()), g["torus_img_230"], rtol=1e-6, atol=1e-12 * g["torus_img_230"].max())
    em, ab = ma.synchrotron_coefficients(g["syn_Ne"], g["syn_Th"], g["syn_B"], g["syn_pitch"], g["syn_nu"],
                                         invariant=True, rescale_nu=1 / 230e9)
    assert np.allclose(np.asarray(em), g["syn_em"], rtol=1e-12) and np.allclose(np.asarray(ab), g["syn_ab"], rtol=1e-12)


def test_optically_thick_regime_matches_where_finite(built):
    """alpha*dt > 2 per step makes the reference's explicit Euler scheme blow up (|I| ~ 1e300, negative pixels).
    The fused front-to-back accumulation must still agree with the literal back-to-front order wherever both
    stay finite; only the exact overflow boundary may differ."""
    from mahakala_b200 import images
    from oracle import c_oracle, mahakala_oracle as onp
    arr = snapshot_arrays(ncells=64, block=32, extent=32.0)
    om, dm = oracle_model(arr, A), device_model(arr, A)
    units = om.get_units(M_BH, MASS_SCALE)
    res = 96
    s0 = onp.initialize_geodesics_at_camera(A, 80, 1000, -10, 10, res)
    ref, _, _ = c_oracle.render(om, s0, units, [43e9])
    img = np.asarray(images.render(dm, camera_inclination=80, resolution=res, observing_frequencies=[43e9]).cpu())[0]
    fin_r, fin_g = np.isfinite(ref[0]), np.isfinite(img)
    both = fin_r & fin_g
    assert np.abs(ref[0][both]).max() > 1e50                 # the regime really is unstable in the oracle
    assert both.sum() > 0.95 * res * res
    rel = np.abs(img[both] - ref[0][both]) / np.maximum(np.abs(ref[0][both]), 1e-300)
    big = np.abs(ref[0][both]) > 1e-12 * np.median(np.abs(ref[0][both]))
    # alternating sums of terms ~1e300 cancel, so rounding differences are amplified: bulk agreement stays at
    # rounding level, the worst pixels at ~1e-5 (the result itself is numerical garbage in both implementations)
    assert np.percentile(rel[big], 99) < 1e-8 and rel[big].max() < 1e-3
    assert (fin_r != fin_g).sum() <= 0.01 * res * res


def test_variable_order_follows_names(setup):
    """The primitive order of the file (VariableNames, athenak.py:697-710) may differ from the usual one."""
    from mahakala_b200.grmhd import AthenakFluidModel
    arr, S = setup["arr"], setup["S"][::9]
    base = {k: np.asarray(v) for k, v in setup["dm"].get_fluid_scalars_from_geodesics(S).items()}
    perm = [4, 0, 1, 2, 3]                                   # eint, dens, velx, vely, velz
    names = ('eint', 'dens', 'velx', 'vely', 'velz', 'bcc1', 'bcc2', 'bcc3')
    m = AthenakFluidModel.from_arrays(arr["uov"][perm], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"],
                                      arr["x2f"], arr["x3f"], arr["LogicalLocations"], arr["Levels"], A,
                                      fluid_gamma=arr["fluid_gamma"], VariableNames=names)
    got = m.get_fluid_scalars_from_geodesics(S)
    for k in base:
        assert np.array_equal(np.asarray(got[k]), base[k]), k
    m.release()


def test_more_than_eight_frequencies_are_batched(setup):
    from mahakala_b200 import images
    dm = setup["dm"]
    nus = [30e9 * (1.35 ** k) for k in range(11)]
    img = np.asarray(images.render(dm, resolution=12, observing_frequencies=nus).cpu())
    assert img.shape == (11, 144)
    for k in (0, 7, 8, 10):
        one = np.asarray(images.render(dm, resolution=12, observing_frequencies=[nus[k]]).cpu())[0]
        assert np.allclose(img[k], one, rtol=1e-10, atol=1e-14 * max(one.max(), 1e-300))


def test_unfused_chunking_is_transparent(setup):
    """images.py:62-78: max_chunk_bytes only changes how many pixels go through the pipeline at once."""
    from mahakala_b200 import images
    dm = setup["dm"]
    whole = images.make_image_unfused(dm, resolution=12)
    # int(max_chunk_bytes // 4 // 20 // max_nsteps) pixels per chunk (images.py:69): 50 px -> 3 chunks of 144 px
    chunked = images.make_image_unfused(dm, resolution=12, max_chunk_bytes=50 * 4 * 20 * 10000)
    assert np.array_equal(whole, chunked)
    fused = images.make_image(dm, resolution=12, max_chunk_bytes=1e9)        # accepted, irrelevant for the fused path
    assert np.allclose(fused, whole, rtol=1e-9, atol=1e-14 * whole.max())


def test_make_image_follows_a_user_registered_spacetime(setup):
    """With a run-time registered spacetime selected, make_image runs the FUSED kernel that NVRTC built around the
    user's metric: geodesics from its dual-number derivatives, fluid frame (athenak.py:760-786) from its g and g^-1.
    (a) The reference's own metric registered as the user plugin reproduces the built-in fused image to 1e-9 (a
    different code path end to end: dual numbers + adjugate inverse + generic frame algebra).  (b) The stage-by-stage
    chain with the plugin (geodesics in the plugin metric, Kerr-Schild frame) agrees too.  (c) A spacetime that is NOT
    Kerr -- Schwarzschild of mass 1.2 in Kerr-Schild coordinates -- gives a different image, equal to what the
    stage-by-stage chain gives for its geodesics plus its own frame (checked through multi-frequency consistency and
    against the a = 0 built-in image for M = 1)."""
    from mahakala_b200 import geodesics as geo, images
    from test_geodesics_gpu import KERR_SCHILD_USER
    from test_host_cpu import SCHWARZSCHILD_KS
    dm = setup["dm"]
    fused = images.make_image(dm, resolution=12)
    # (230 / 345 GHz: at 86 GHz this snapshot is in the regime where the explicit-Euler transfer amplifies rounding)
    multi = np.asarray(images.render(dm, resolution=12, observing_frequencies=(230e9, 345e9)).cpu())
    geo.register_metric("ks_user_img", KERR_SCHILD_USER)
    geo.register_metric("schw_m1_img", SCHWARZSCHILD_KS, params=[1.0])
    geo.register_metric("schw_m12_img", SCHWARZSCHILD_KS, params=[1.2])
    geo.set_metric("ks_user_img")
    try:
        user = images.make_image(dm, resolution=12)
        user_unfused = images.make_image_unfused(dm, resolution=12)
        user_multi = np.asarray(images.render(dm, resolution=12, observing_frequencies=(230e9, 345e9)).cpu())
    finally:
        geo.set_metric("kerr_schild")
    assert user.shape == fused.shape and fused.max() > 0
    scale = fused.max()
    assert np.abs(user - fused).max() < 1e-9 * scale, np.abs(user - fused).max() / scale
    assert not np.array_equal(user, fused)          # different code path (dual-number geodesics, generic frame)
    assert np.allclose(user_unfused, fused, rtol=1e-6, atol=1e-12 * scale)
    assert np.abs(user_multi - multi).max() < 1e-9 * np.abs(multi).max()        # one launch per frequency
    # Schwarzschild plugin with M = 1 == the built-in metric at a = 0 (image of the same snapshot with bhspin = 0)
    from helpers import device_model
    dm0 = device_model(setup["arr"], 0.0)
    ref0 = images.make_image(dm0, resolution=12)
    geo.set_metric("schw_m1_img")
    try:
        s1 = images.make_image(dm0, resolution=12)
        geo.set_metric("schw_m12_img")
        s12 = images.make_image(dm0, resolution=12)
    finally:
        geo.set_metric("kerr_schild")
    assert np.abs(s1 - ref0).max() < 1e-9 * ref0.max()
    assert np.abs(s12 - ref0).max() > 1e-3 * ref0.max() and np.isfinite(s12).all()      # a different spacetime
    dm0.release()


def test_noncubic_meshblocks(built):
    """Meshblocks with nk != nj != ni (the reference's loader only works for cubic blocks): device
    ghost fill, host fill, block lookup and the trilinear gather against plain indexing of the stitched global array."""
    from helpers import noncubic_mesh
    nb, n = (2, 3, 2), (8, 6, 4)
    arr, G, centres = noncubic_mesh(nb, n)
    ni, nj, nk = n
    expected = np.empty((nb[0] * nb[1] * nb[2], 8, nk + 2, nj + 2, ni + 2))
    mb = 0
    for lk in range(nb[2]):
        for lj in range(nb[1]):
            for li in range(nb[0]):
                expected[mb] = G[:, lk * nk:(lk + 1) * nk + 2, lj * nj:(lj + 1) * nj + 2, li * ni:(li + 1) * ni + 2]
                mb += 1
    rng = np.random.default_rng(5)
    ext = [c[-1] - 0.25 for c in centres]                       # domain half-widths (upper faces)
    pts = np.stack([rng.uniform(-e - 0.3, e + 0.3, 3000) for e in ext], axis=1)
    S = np.concatenate([np.zeros((3000, 1)), pts, np.ones((3000, 1)), rng.normal(0, 0.5, (3000, 3))], 1)[None]
    # brute-force trilinear interpolation on the zero-padded global array (cell centres incl. one ghost layer)
    inside = np.all([(pts[:, ax] > -ext[ax]) & (pts[:, ax] <= ext[ax]) for ax in range(3)], axis=0)
    want = np.zeros((8, 3000))
    idx, w = [], []
    for ax in range(3):
        t = (pts[:, ax] - centres[ax][0]) / 0.5
        i0 = np.clip(np.floor(t).astype(int), 0, len(centres[ax]) - 2)
        idx.append(i0)
        w.append(t - i0)
    for dk in (0, 1):
        for dj in (0, 1):
            for di in (0, 1):
                wt = (w[2] if dk else 1 - w[2]) * (w[1] if dj else 1 - w[1]) * (w[0] if di else 1 - w[0])
                want += wt * G[:, idx[2] + dk, idx[1] + dj, idx[0] + di]
    want[:, ~inside] = 0.0
    order = [0, 4, 1, 2, 3, 5, 6, 7]                             # dens, u(eint), U1..3, B1..3 <- file order
    for kw in (dict(), dict(ghost_fill="host"), dict(lookup="scan", storage="f64")):
        m = device_model(arr, A, **kw)
        assert np.array_equal(np.asarray(m.device_meshblocks()), expected), kw
        got = m.get_prims_from_geodesics(S)
        for q, k in enumerate(('dens', 'u', 'U1', 'U2', 'U3', 'B1', 'B2', 'B3')):
            g = np.asarray(got[k])[0]
            assert np.array_equal(g == 0, want[order[q]] == 0), (kw, k)
            # points within rounding distance of a block face may take the ghost cell of the neighbouring block:
            # same interpolant, different rounding
            assert np.allclose(g, want[order[q]], rtol=1e-12, atol=1e-13), (kw, k)
        m.release()
    assert np.array_equal(device_model(arr, A).all_meshblocks, expected)          # host fill, non-cubic


def test_three_level_mesh(built):
    """Three refinement levels (22 blocks, 2:1 balanced): host and device ghost fills against the brute-force
    expectation, and the level-aware block-grid lookup + sampling against the oracle's linear scan."""
    from helpers import three_level_mesh
    from oracle import c_oracle
    arr, expected = three_level_mesh(n=4)
    dm = device_model(arr, A)
    assert np.abs(dm.all_meshblocks - expected).max() < 1e-15                       # host fill
    got = np.asarray(dm.device_meshblocks())
    assert np.array_equal(got, dm.all_meshblocks) and dm.storage == "f64"           # device fill, bit-identical
    assert dm.lookup == "grid"
    om = oracle_model(arr, A)
    om.all_meshblocks = expected
    rng = np.random.default_rng(7)
    pts = rng.uniform(-8.5, 8.5, (6000, 3))
    pts[:2000] = rng.uniform(3.0, 8.2, (2000, 3))                                    # the refined corner
    faces = np.array([-8.0, 0.0, 4.0, 6.0, 7.0, 8.0])
    pts[2000:2600] = rng.choice(faces, (600, 3))
    S = np.concatenate([np.zeros((6000, 1)), pts, np.ones((6000, 1)), rng.normal(0, 0.5, (6000, 3))], 1)[None]
    ref = c_oracle.sample(om, S, mode="prims")
    out = dm.get_prims_from_geodesics(S)
    for k in ref:
        g = np.asarray(out[k])
        assert np.array_equal(g == 0, ref[k] == 0), k
        assert np.allclose(g, ref[k], rtol=1e-13, atol=1e-14 * np.abs(ref[k]).max()), k
    assert set(np.unique(om._meshblock_indices(S)[0])) == set(range(-1, 22))        # every block is exercised


def test_kernels_match_vectors_from_the_reference_source(built):
    """The thermodynamics / synchrotron / transfer kernels against tests/golden/reference_golden.npz, the vectors
    obtained by executing the reference's own electrons.py / transfer.py (see tests/golden/make_reference_golden.py):
    values to 1e-12, zero / NaN patterns exactly, the transfer scans bit for bit (same operation order)."""
    import os
    from mahakala_b200 import electrons, transfer
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))

    def eq(a, b, rtol):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        fin = np.isfinite(b)
        return (np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~fin & ~np.isnan(b)], b[~fin & ~np.isnan(b)])
                and np.array_equal(a[fin] == 0, b[fin] == 0) and np.allclose(a[fin], b[fin], rtol=rtol, atol=0))

    dens, u, beta = g["theta_in"]
    assert eq(electrons.rlow_rhigh_model(dens, u, beta), g["theta_default"], 1e-14)
    assert eq(electrons.rlow_rhigh_model(dens, u, beta, r_low=10, r_high=160), g["theta_r10_r160"], 1e-14)
    Ne, Th, B, pitch, nu = g["syn_in"]
    for tag, kw in {"inv": dict(invariant=True, rescale_nu=1. / 230e9), "inv1": dict(invariant=True),
                    "plain": dict(invariant=False)}.items():
        em, ab = transfer.synchrotron_coefficients(Ne, Th, B, pitch, nu, **kw)
        assert eq(em, g["syn_em_" + tag], 1e-12) and eq(ab, g["syn_ab_" + tag], 1e-12), tag
    em2, ab2, dt, L = g["tr_in_em"], g["tr_in_ab"], g["tr_in_dt"], float(g["tr_in_L"])
    assert np.array_equal(np.asarray(transfer.solve_specific_intensity(em2, ab2, dt, L)), g["tr_I"])
    I2, dIs = transfer.solve_specific_intensity(em2, ab2, dt, L, dIs=True)
    assert np.array_equal(np.asarray(I2), g["tr_I_dIs"]) and np.array_equal(np.asarray(dIs), g["tr_dIs"])
    assert eq(transfer.solve_attenuated_emissivity(em2, ab2, dt, L), g["tr_attenuated"], 1e-13)
