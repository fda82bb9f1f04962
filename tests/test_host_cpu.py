"""CPU suite: C-ABI surface, host-side logic of the package, loud failure without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names += re.findall(r"\b(mk_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(built):
    from mahakala_b200 import _cabi
    lib = _cabi.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    bound = set(_cabi.SIGNATURES) | set(_cabi.OTHER)
    assert bound == set(declared), (bound ^ set(declared))
    assert _cabi.call("mk_abi_version") == 1
    assert lib.mk_page_rows() == 16
    # argument-validation paths return an error string without touching the GPU
    with pytest.raises(_cabi.MahakalaB200Error, match="metric"):
        _cabi.call("mk_rhs", 99, 0.5, 1, 1, 1, None)       # unknown metric id is rejected before any launch
    assert lib.mk_render_patch_count(1024, None, 0) == 256 * 128
    assert lib.mk_render_patch_count(10, None, 0) == 3 * 2
    assert lib.mk_render_patch_count(0, 8, 77) == 3


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu(built):
    import mahakala_b200 as ma
    from mahakala_b200 import _cabi
    with pytest.raises(_cabi.MahakalaB200Error, match="no CPU fallback"):
        ma.initialize_geodesics_at_camera(0.9, 60, 1000, -10, 10, 8)
    with pytest.raises(_cabi.MahakalaB200Error):
        ma.geodesic_integrator(10, np.zeros((4, 8)), 40, 1e-2, 0.9)
    with pytest.raises(_cabi.MahakalaB200Error):
        ma.find_shadow_bisection_angles(0.9, 60, np.array([0.0, 1.0]))
    with pytest.raises(_cabi.MahakalaB200Error):
        ma.synchrotron_coefficients(np.ones(3), np.ones(3), np.ones(3), np.ones(3), np.ones(3))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mahakala_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.lower().replace("# oracle", ""), os.path.join(dirpath, fn)


def test_api_surface_matches_reference_exports():
    import inspect
    import mahakala_b200 as ma
    from mahakala_b200 import geodesics, images, electrons
    assert ma.__all__ == ["find_shadow_bisection", "find_shadow_bisection_angles", "geodesic_integrator",
                          "initialize_geodesics_at_camera", "synchrotron_coefficients", "solve_specific_intensity",
                          "solve_attenuated_emissivity"]
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(ma.initialize_geodesics_at_camera) == ["bhspin", "inclination", "distance", "fov_lower", "fov_upper",
                                                      "pixels_per_side", "camera_type"]
    assert sig(ma.geodesic_integrator) == ["N", "s0", "div", "tol", "bhspin"]
    assert sig(ma.find_shadow_bisection_angles) == ["bhspin", "inc", "angles", "max_steps", "error_allowed", "max_it"]
    assert sig(ma.find_shadow_bisection) == ["bhspin", "inc", "num_angles", "max_steps", "error_allowed", "max_it"]
    assert sig(geodesics.select_photons_integrator) == ["inc", "angle", "radius", "bhspin", "distance", "max_steps"]
    assert sig(ma.synchrotron_coefficients) == ["Ne", "Theta_e", "B", "pitch_angle", "nu", "invariant", "rescale_nu"]
    assert sig(ma.solve_specific_intensity) == ["emissivity", "absorptivity", "dt", "L_unit", "dIs"]
    assert sig(ma.solve_attenuated_emissivity) == ["emissivity", "absorptivity", "dt", "L_unit"]
    assert sig(electrons.rlow_rhigh_model) == ["dens", "u", "beta", "r_low", "r_high", "electron_gamma", "ion_gamma"]
    assert sig(images.make_image) == ["fluid_model", "camera_inclination", "camera_distance", "mass_scale", "M_bh",
                                      "r_high", "observing_frequency", "fov", "resolution", "max_nsteps", "max_chunk_bytes"]
    d = inspect.signature(images.make_image).parameters
    assert (d["camera_inclination"].default, d["camera_distance"].default, d["mass_scale"].default, d["r_high"].default,
            d["observing_frequency"].default, d["fov"].default, d["resolution"].default, d["max_nsteps"].default) == \
        (60, 1000, 1.e26, 40, 230.e9, 20, 160, 10000)
    for name in ("metric", "imetric", "rhs", "RK4_gen", "radius_cal", "radius_EH", "get_camera_pixel",
                 "get_initial_grid", "initial_condition"):
        assert callable(getattr(geodesics, name))
    from mahakala_b200 import constants as c
    assert (c.EE, c.KB, c.CL, c.ME, c.HPL, c.GNEWT, c.Msun, c.MP) == \
        (4.8032e-10, 1.3807e-16, 2.99792458e10, 9.1094e-28, 6.6261e-27, 6.6743e-8, 1.989e33, 1.6726e-24)
    assert geodesics.radius_EH(0.6) == 1.8
    me = ma.install_as_mahakala()
    import mahakala
    from mahakala.images import make_image
    assert mahakala is me and make_image is images.make_image


def test_dump_rows_rule():
    from mahakala_b200.geodesics import dump_rows
    assert dump_rows(2000, 1291) == 1293          # first all-zero row 1291 -> +2   (SURVEY cfg1)
    assert dump_rows(2000, 0) == 2000             # all rays frozen at row 0 -> N rows (geodesics.py:277-279)
    assert dump_rows(2000, 2000) == 2000          # nobody froze -> N rows
    assert dump_rows(2000, 1999) == 2000          # clipped to the N rows the scan produced
    assert dump_rows(2000, 1998) == 2000
    assert dump_rows(2000, 1997) == 1999


def test_units_and_ghost_fill_and_block_grid(built):
    from helpers import oracle_model, snapshot_arrays
    from mahakala_b200.grmhd import GRMHDFluidModel
    from mahakala_b200.grmhd.athenak import build_block_grid, fill_ghost_zones
    from oracle import mahakala_oracle as onp
    u = GRMHDFluidModel().get_units(6.2e9 * 1.989e33, 1e26)
    assert u == onp.GRMHDFluidModel().get_units(6.2e9 * 1.989e33, 1e26)
    arr = snapshot_arrays(ncells=24, block=8, extent=12.0)
    amb, index = fill_ghost_zones(arr["uov"], arr["B"], arr["LogicalLocations"], arr["Levels"])
    om = oracle_model(arr, 0.5)
    assert np.array_equal(amb, om.all_meshblocks) and index == om.mb_index_map
    assert np.array_equal(amb.astype(np.float32).astype(np.float64), amb)       # f32 storage is lossless
    g, gn, g0, ginv = build_block_grid(arr["x1f"], arr["x2f"], arr["x3f"])
    assert g.shape == (3, 3, 3) and list(gn) == [3, 3, 3] and (g >= 0).all() and len(set(g.ravel())) == 27
    rng = np.random.default_rng(0)
    pts = rng.uniform(-12, 12, (500, 3))
    S = np.concatenate([np.zeros((500, 1)), pts, np.zeros((500, 4))], 1)
    mb_ref = om._meshblock_indices(S)
    c = np.clip(np.floor((pts - g0) * ginv).astype(int), 0, gn - 1)
    assert np.array_equal(g[c[:, 2], c[:, 1], c[:, 0]], mb_ref)
    # irregular mesh (blocks that are not integer multiples of the smallest) -> scan fallback
    x1f = arr["x1f"].copy(); x1f[0, -1] += 0.37
    assert build_block_grid(x1f, arr["x2f"], arr["x3f"]) is None


def test_device_array_numpy_protocol():
    from mahakala_b200._device import DeviceArray
    t = DeviceArray.wrap(torch.arange(6, dtype=torch.float64).reshape(2, 3))
    assert np.allclose(t, [[0, 1, 2], [3, 4, 5]])
    assert isinstance(t[0], DeviceArray) and np.asarray(t * t)[1, 2] == 25.0
    assert np.where(np.asarray(t) > 2)[0].tolist() == [1, 1, 1]


SCHWARZSCHILD_KS = r"""
// Schwarzschild in Kerr-Schild coordinates with mass M = params[1]: g = eta + (2M/r) l l, l = (1, x/r, y/r, z/r)
struct UserMetric {
    static constexpr bool stationary = true;      // optional: lets the plugin skip the t tangent
    double params[8];
    template <class T> __device__ void operator()(const T x[4], T g[4][4]) const {
        const double M = params[1];
        T r = mk_sqrt(x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
        T f = (2.0 * M) / r;
        T l[4];
        l[0] = T(1.0); l[1] = x[1] / r; l[2] = x[2] / r; l[3] = x[3] / r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                T e = f * (l[i] * l[j]);
                g[i][j] = (i == j) ? e + (i == 0 ? -1.0 : 1.0) : e;
            }
    }
    __device__ double radius(const double x[4]) const { return sqrt(x[1] * x[1] + x[2] * x[2] + x[3] * x[3]); }
    __device__ double horizon() const { return 2.0 * params[1]; }
};
"""


def test_user_metric_compiles_without_gpu(built):
    """NVRTC compiles a user-registered spacetime for sm_100a on a CPU-only host; errors carry the log."""
    from mahakala_b200 import _cabi, geodesics as geo
    mid = geo.register_metric("schwarzschild_ks_cpu_test", SCHWARZSCHILD_KS, params=[1.0])
    assert mid >= 16 and geo._METRICS["schwarzschild_ks_cpu_test"] == mid
    with pytest.raises(_cabi.MahakalaB200Error, match="compilation of metric"):
        geo.register_metric("broken", "struct UserMetric { double params[8]; this is not C++ };")
    with pytest.raises(_cabi.MahakalaB200Error, match="unknown metric id"):
        _cabi.call("mk_metric_set_params", 999, (__import__("ctypes").c_double * 8)())


def test_amr_ghost_fill_two_levels_bruteforce(built):
    """Ghost cells across a refinement boundary (athenak.py:231-514): coarser neighbour -> injection, finer
    neighbour -> average of the 8 children.  Checked against plain indexing of global fine / coarse arrays."""
    from helpers import two_level_mesh
    from mahakala_b200.grmhd.athenak import build_block_grid, fill_ghost_zones
    arr, expected = two_level_mesh(n=8)
    amb, index = fill_ghost_zones(arr["uov"], arr["B"], arr["LogicalLocations"], arr["Levels"])
    assert amb.shape == expected.shape == (15, 8, 10, 10, 10)
    assert np.abs(amb - expected).max() < 1e-15
    assert (expected[:, 0] > 0).sum() > 0.7 * expected[:, 0].size     # only domain-boundary ghosts stay zero
    grid, gn, g0, ginv = build_block_grid(arr["x1f"], arr["x2f"], arr["x3f"])
    assert grid.shape == (4, 4, 4) and (grid >= 0).all()
    assert len(set(grid[:2].ravel()) | set(grid[:, :2].ravel()) | set(grid[:, :, :2].ravel())) == 7   # 7 coarse blocks
    assert len(set(grid[2:, 2:, 2:].ravel())) == 8                                                     # 8 fine blocks


def test_primitive_name_mapping_and_errors():
    """athenak.py:55-65 / :697-710: primitives are found by name; non-contiguous velocity or field components
    raise ValueError (checked on the host before anything is uploaded)."""
    from helpers import snapshot_arrays
    from mahakala_b200.grmhd import AthenakFluidModel
    arr = snapshot_arrays(ncells=16, block=8, extent=8.0)

    def model(names):
        return AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"],
                                             arr["x2f"], arr["x3f"], arr["LogicalLocations"], arr["Levels"], 0.5,
                                             fluid_gamma=4. / 3, VariableNames=names)

    m = model(('dens', 'velx', 'vely', 'velz', 'eint', 'bcc1', 'bcc2', 'bcc3'))
    assert m.get_index_for_primitive_by_name(' EINT ') == 4 and m.get_index_for_primitive_by_name('nope') == -1
    assert m._prim_index() == [0, 4, 1, 2, 3, 5, 6, 7]
    # a different but valid file order is honoured
    m2 = model(('eint', 'dens', 'bcc1', 'bcc2', 'bcc3', 'velx', 'vely', 'velz'))
    assert m2._prim_index() == [1, 0, 5, 6, 7, 2, 3, 4]
    with pytest.raises(ValueError, match="Velocity"):
        model(('dens', 'velx', 'eint', 'vely', 'velz', 'bcc1', 'bcc2', 'bcc3'))._prim_index()
    with pytest.raises(ValueError, match="Magnetic"):
        model(('dens', 'velx', 'vely', 'velz', 'bcc1', 'eint', 'bcc2', 'bcc3'))._prim_index()
    assert m.all_meshblocks.shape == (8, 8, 10, 10, 10) and m.nprim_all == 8
    assert set(m.mb_index_map) == {(0, i, j, k) for i in range(2) for j in range(2) for k in range(2)}


def test_file_constructor_npz_roundtrip(tmp_path):
    """AthenakFluidModel(filename, bhspin, fluid_gamma) keeps the reference signature (athenak.py:50); .npz files
    with the .athdf datasets are read as well; a missing file is an OSError."""
    from helpers import snapshot_arrays
    from mahakala_b200.grmhd import AthenakFluidModel
    arr = snapshot_arrays(ncells=16, block=8, extent=8.0)
    fn = tmp_path / "snap.npz"
    np.savez(fn, **{k: arr[k] for k in AthenakFluidModel.DATASETS}, VariableNames=np.array(arr["VariableNames"]))
    m = AthenakFluidModel(str(fn), 0.7, fluid_gamma=13. / 9)
    ref = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                        arr["x3f"], arr["LogicalLocations"], arr["Levels"], 0.7, fluid_gamma=13. / 9)
    assert np.array_equal(m.all_meshblocks, ref.all_meshblocks) and m.bhspin == 0.7 and m.fluid_gamma == 13. / 9
    assert list(m.variable_names) == list(ref.variable_names)
    with pytest.raises(OSError):
        AthenakFluidModel("missing.athdf", 0.7)


MATLAB_HDF5 = os.path.join(os.path.dirname(__import__("scipy").__file__), "io", "matlab", "tests", "data",
                           "testhdf5_7.4_GLNX86.mat")


@pytest.mark.skipif(not os.path.exists(MATLAB_HDF5), reason="scipy's MATLAB v7.3 test file is not installed")
def test_minimal_hdf5_reader_on_a_third_party_file():
    """The built-in HDF5 reader (used for .athdf dumps when h5py is missing) parses a file written by someone
    else's HDF5 library: MATLAB 7.4's v7.3 MAT-file from SciPy's test data (512 B user block, superblock 0,
    symbol-table group, version-1 object headers, version-2 layout message).  Its one variable is the same
    0 : pi/4 : 2 pi ramp that SciPy reads from the v5 twin of the file."""
    import scipy.io
    from mahakala_b200.grmhd._hdf5_min import Hdf5File
    f = Hdf5File(MATLAB_HDF5)
    assert f.base == 512 and f.keys() == ["testdouble"]
    got = f["testdouble"]
    twin = scipy.io.loadmat(MATLAB_HDF5.replace("testhdf5", "testdouble"))["testdouble"]
    assert got.dtype == np.float64 and np.array_equal(got.ravel(), twin.ravel())
    assert np.array_equal(got.ravel(), np.arange(9) * np.pi / 4)
    with pytest.raises(KeyError):
        f["nope"]


def test_minimal_hdf5_reader_latest_format_flavour(tmp_path):
    """Superblock 2, version-2 object headers, compact link messages, version-2 dataspaces, version-3 attributes
    (what h5py writes with libver='latest'); dense link storage is refused loudly."""
    from hdf5_writer import write_hdf5_v2
    from mahakala_b200.grmhd._hdf5_min import Hdf5File
    rng = np.random.default_rng(0)
    data = {"uov": rng.normal(size=(5, 3, 4, 4, 4)).astype(np.float32), "Levels": np.arange(3, dtype=np.int32),
            "x1f": rng.normal(size=(3, 5)), "LogicalLocations": rng.integers(0, 9, (3, 3)).astype(np.int64)}
    fn = str(tmp_path / "latest.h5")
    write_hdf5_v2(fn, data, attrs={"VariableNames": np.array(["dens", "velx"], dtype="S"), "Time": np.float64(3.25)})
    f = Hdf5File(fn)
    assert sorted(f.keys()) == sorted(data) and f.attrs["Time"] == 3.25
    assert [n.decode() for n in f.attrs["VariableNames"]] == ["dens", "velx"]
    for k, v in data.items():
        got = f[k]
        assert got.dtype == v.dtype and np.array_equal(got, v), k


@pytest.mark.parametrize("variant", ["contiguous_f32", "chunked_userblock_f64"])
def test_file_constructor_reads_athdf_without_h5py(tmp_path, variant):
    """AthenakFluidModel("x.athdf", bhspin, fluid_gamma) (athenak.py:50, :79-103) through the built-in reader: an
    .athdf-shaped HDF5 file (float32 data as AthenaK writes it, int LogicalLocations / Levels, VariableNames as a
    fixed-length byte-string attribute) gives the same model as from_arrays."""
    from hdf5_writer import write_hdf5
    from helpers import snapshot_arrays
    from mahakala_b200.grmhd import AthenakFluidModel
    from mahakala_b200.grmhd._hdf5_min import Hdf5File, Hdf5FormatError
    arr = snapshot_arrays(ncells=16, block=8, extent=8.0)
    f32 = variant.endswith("f32")
    fdt = np.float32 if f32 else np.float64
    data = {k: np.asarray(arr[k], dtype=fdt) for k in ('x1v', 'x2v', 'x3v', 'x1f', 'x2f', 'x3f', 'uov', 'B')}
    data["LogicalLocations"] = np.asarray(arr["LogicalLocations"], dtype=np.int64)
    data["Levels"] = np.asarray(arr["Levels"], dtype=np.int32)
    fn = str(tmp_path / "snap.athdf")
    write_hdf5(fn, data, attrs={"VariableNames": np.array(arr["VariableNames"], dtype="S"), "Time": np.float64(12.5),
                                "NumCycles": np.int32(7)},
               chunked=("uov", "B", "x1v") if variant.startswith("chunked") else (),
               userblock=512 if "userblock" in variant else 0)
    f = Hdf5File(fn)
    assert sorted(f.keys()) == sorted(data) and f.attrs["Time"] == 12.5 and f.attrs["NumCycles"] == 7
    for k, v in data.items():
        got = f[k]
        assert got.dtype == v.dtype and np.array_equal(got, v), k
    m = AthenakFluidModel(fn, 0.7, fluid_gamma=13. / 9)
    ref = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"], arr["x2f"],
                                        arr["x3f"], arr["LogicalLocations"], arr["Levels"], 0.7, fluid_gamma=13. / 9)
    assert list(m.variable_names) == list(ref.variable_names) and m._uov.dtype == fdt
    assert np.array_equal(m.all_meshblocks, ref.all_meshblocks)         # the synthetic values are float32-exact
    assert np.array_equal(m.x1f, ref.x1f) and np.array_equal(m.Levels, ref.Levels)
    bad = tmp_path / "bad.athdf"
    bad.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(Hdf5FormatError):
        AthenakFluidModel(str(bad), 0.7)
