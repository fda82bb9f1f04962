"""CUDA path against the vectors produced by the whole reference package (tests/golden/reference_fluid_golden.npz).

Kept in its own module, collected last (written at the end of round 1; green on B200 since round 2).
"""
import numpy as np
import pytest

from helpers import device_model

pytestmark = pytest.mark.gpu
A = 0.94


def test_cuda_path_matches_the_reference_package_on_snapshots_and_images(built):
    """The CUDA sampling / image path against tests/golden/reference_fluid_golden.npz, produced by the whole reference
    package (its own loader, get_prims_from_geodesics, get_fluid_scalars_from_geodesics and make_image) running under a
    NumPy-backed `jax` (tests/golden/make_reference_fluid_golden.py): device ghost fill bit for bit, sampled
    primitives to 1e-13 with the exact zero pattern, fluid scalars to 1e-10, images to the north-star 1e-6 per pixel
    and 1e-8 in flux."""
    import os
    from mahakala_b200 import images
    from mahakala_b200.synthetic import make_synthetic_snapshot
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_fluid_golden.npz"))
    single = make_synthetic_snapshot(ncells=16, block=8, extent=16.0, seed=0)
    dm = device_model(single, A)
    assert np.array_equal(np.asarray(dm.device_meshblocks()), g["single_all_meshblocks"])
    S = g["sample_S"]
    got = dm.get_prims_from_geodesics(S)
    for q, k in enumerate(('dens', 'u', 'U1', 'U2', 'U3', 'B1', 'B2', 'B3')):
        ref = g["sample_prims"][q]
        v = np.asarray(got[k])
        assert np.array_equal(v == 0, ref == 0), k
        assert np.abs(v - ref).max() <= 1e-13 * np.abs(ref).max(), k
    sc = dm.get_fluid_scalars_from_geodesics(S)
    for q, k in enumerate(('dens', 'u', 'pitch_angle', 'kdotu', 'b')):
        ref = g["sample_scalars"][q]
        assert np.abs(np.asarray(sc[k]) - ref).max() <= 1e-10 * np.abs(ref).max(), k

    def check(img, ref):
        err = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-6 * ref.max())
        assert ref.max() > 0 and err.max() < 1e-6 and abs(img.sum() - ref.sum()) / ref.sum() < 1e-8, (err.max(),)

    check(images.make_image(dm, resolution=6), g["image_res6"])
    check(images.make_image_unfused(dm, resolution=6, max_chunk_bytes=12 * 4 * 20 * 10000), g["image_res6_chunked"])
    check(images.make_image(dm, camera_inclination=30, observing_frequency=345e9, r_high=10, resolution=6, max_nsteps=3000),
          g["image_res6_345GHz_i30"])
    dm.release()


def test_cuda_cfg1_grid_against_the_reference_source(built):
    """BASELINE config 1 against tests/golden/reference_cfg1_golden.npz (the reference's own geodesics.py on the full
    64x64 grid under the NumPy stand-in): captured / escaped classification bit-exact (792 captured), step counts
    identical on escaped rays, end states of escaped rays within the north-star 1e-9.  Mirrors
    test_geodesics_gpu.py::test_cfg1_grid_classification_and_states with the oracle replaced by reference output."""
    import os
    import mahakala_b200 as ma
    from mahakala_b200 import geodesics as geo
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cfg1_golden.npz"))
    s0 = np.asarray(ma.initialize_geodesics_at_camera(A, 60, 1000, -10, 10, 64))
    assert np.array_equal(s0[:, :4], g["s0"][:, :4]) and np.allclose(s0[:, 4:], g["s0"][:, 4:], rtol=1e-13, atol=1e-300)
    final, nsteps, r_last = (np.asarray(q.cpu()) for q in geo.integrate_final(2000, g["s0"], 40, 1e-2, A))
    cap_ref = g["r_last"] < 100
    assert cap_ref.sum() == 792 and np.array_equal(r_last < 100, cap_ref)
    esc = ~cap_ref
    assert np.array_equal(nsteps[esc], g["nsteps"][esc])
    assert abs(int(nsteps.sum()) - 2079364) <= 64          # captured rays may differ by a step in the chaotic tail
    err = np.abs(final[esc] - g["final"][esc]).max(axis=1) / np.abs(g["final"][esc]).max(axis=1)
    assert np.median(err) < 1e-12 and err.max() < 1e-9, err.max()


def test_cuda_cfg2_sublattice_against_the_reference_source(built):
    """BASELINE config 2 settings on the every-16th-pixel sub-lattice against tests/golden/reference_cfg2_golden.npz
    (the reference's own geodesics.py under the NumPy stand-in; 799 captured rays, 2 177 333 ray-steps): classification
    bit-exact, step counts identical on escaped rays, end states of escaped rays within the north-star 1e-9."""
    import os
    from mahakala_b200 import geodesics as geo
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cfg2_golden.npz"))
    final, nsteps, r_last = (np.asarray(q.cpu()) for q in geo.integrate_final(10000, g["s0"], 40, 1e-4, A))
    cap_ref = g["r_last"] < 100
    assert cap_ref.sum() == 799 and np.array_equal(r_last < 100, cap_ref)
    esc = ~cap_ref
    assert np.array_equal(nsteps[esc], g["nsteps"][esc])
    assert abs(int(nsteps.sum()) - 2177333) <= 4 * 799          # captured rays: chaotic tail, a few steps either way
    err = np.abs(final[esc] - g["final"][esc]).max(axis=1) / np.abs(g["final"][esc]).max(axis=1)
    assert np.median(err) < 1e-12 and err.max() < 1e-9, err.max()


def test_cuda_image12_against_the_reference_package(built):
    """The fused and the stage-by-stage CUDA image paths against the 12x12 image produced by the reference's own
    make_image (tests/golden/reference_image12_golden.npz): north-star tolerances, 1e-6 per pixel and 1e-8 in flux."""
    import os
    from helpers import snapshot_arrays
    from mahakala_b200 import images
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_image12_golden.npz"))["image_res12"]
    dm = device_model(snapshot_arrays(ncells=32, block=16, extent=16.0), A)
    for img in (images.make_image(dm, resolution=12), images.make_image_unfused(dm, resolution=12)):
        err = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-6 * ref.max())
        assert img.shape == (12, 12) and err.max() < 1e-6 and abs(img.sum() - ref.sum()) / ref.sum() < 1e-8, err.max()
    dm.release()
