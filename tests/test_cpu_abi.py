"""CPU tier: the C-ABI library loads and exports every symbol include/ssb.h declares, fails loudly without a GPU,
the code generator emits what the reference's would, and the host-side mirror keeps the reference's surface."""
import ctypes
import inspect
import os
import pickle
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, has_reference
from util import load_model


def _lib():
    from spatialpy_b200 import codegen
    return codegen.build_core()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ssb.h")).read()
    declared = sorted(set(re.findall(r"\b(ssb_[a-z_]+)\s*\(", hdr)) - {"ssb_progress_cb"})
    assert len(declared) >= 14
    lib = ctypes.CDLL(_lib())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ssb.h but not exported"
    from spatialpy_b200 import engine
    assert sorted(engine.EXPORTS) == declared
    assert lib.ssb_abi_version() == engine.SSB_ABI_VERSION


def test_only_sm100a_code_in_the_library():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out)


def test_peaks_library_exports_its_header_and_fails_loudly_without_a_gpu():
    """include/ssb_peaks.h (the fp64 peak microbenchmark bench.py runs beside the roofline): symbol exported, sm_100a DFMA code,
    and no silent number when there is no device."""
    from spatialpy_b200 import codegen, peaks
    path = codegen.build_peaks()
    hdr = open(os.path.join(ROOT, "include", "ssb_peaks.h")).read()
    declared = sorted(set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == ["ssb_fp64_peak"]
    lib = ctypes.CDLL(path)
    assert all(hasattr(lib, name) for name in declared)
    elf = subprocess.run(["cuobjdump", "--list-elf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf)
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    assert sass.count("DFMA") >= 64
    if not os.path.exists("/dev/nvidia0"):
        with pytest.raises(RuntimeError, match="return code = 3"):
            peaks.fp64_peak(0)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_cpu_fallback_fails_loudly():
    from spatialpy_b200.engine import Engine, EngineError
    with pytest.raises(EngineError) as ei:
        Engine(load_model("birth_death"))
    assert ei.value.code == 3 and "no CPU fallback" in str(ei.value)


def test_flatmodel_roundtrip(tmp_path):
    from spatialpy_b200 import FlatModel
    fm = load_model("cylinder")
    p = str(tmp_path / "m.npz")
    fm.save(p)
    fm2 = FlatModel.load(p)
    for k in FlatModel._ARRAYS:
        np.testing.assert_array_equal(getattr(fm, k), getattr(fm2, k))
    assert [r.propensity for r in fm.reactions] == [r.propensity for r in fm2.reactions]
    assert fm2.h == fm.h and fm2.type_constants == fm.type_constants


def test_codegen_emits_reference_propensity_text():
    from spatialpy_b200 import codegen
    fm = load_model("cylinder")
    src = codegen.generate_unit_source(fm)
    assert "return (((P0*x[0])*x[1])/vol);" in src                      # mass-action text from the reference's converter
    assert "if(sd == type_Edge1){" in src and "return 0.0;}" in src       # restrict_to wrapper (solver.py:356-368)
    assert "#define SSB_SD 2" in src and "#define SSB_RD 3" in src
    fm = load_model("cavity2d")
    assert "me->v[0]=1.0;" in codegen.generate_unit_source(fm)           # BC text passes through unmodified


def test_model_unit_builds_for_sm100a_and_exports_the_table():
    from spatialpy_b200 import codegen
    so = codegen.build_model_unit(load_model("birth_death"))
    lib = ctypes.CDLL(so)
    assert hasattr(lib, "ssbm_get_unit")
    sass = subprocess.run(["cuobjdump", "--list-elf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_solver_surface_and_pickle():
    from spatialpy_b200 import Solver, SimulationError
    with pytest.raises(SimulationError):
        Solver(object())
    sol = Solver(load_model("birth_death"), debug_level=0)
    for attr in ("model", "is_compiled", "debug_level", "model_name", "build_dir", "propfilename", "prop_file_name",
                 "executable_name", "h"):
        assert hasattr(sol, attr)
    sol.compile()
    assert sol.is_compiled and abs(sol.h - 0.24444444444444458) < 1e-15
    sol2 = pickle.loads(pickle.dumps(sol))
    assert sol2.is_compiled and sol2.unit_path == sol.unit_path
    run_params = list(inspect.signature(Solver.run).parameters)
    assert run_params[:8] == ["self", "number_of_trajectories", "seed", "timeout", "number_of_threads", "debug", "profile", "verbose"]


def test_ensemble_sharding_partitions_trajectories():
    from spatialpy_b200.ensemble import shard_trajectories
    for n in (0, 1, 7, 1024):
        for world in (1, 2, 8):
            shards = [shard_trajectories(n, r, world) for r in range(world)]
            assert sorted(sum(shards, [])) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def _need_spatialpy():
    """The reference front-end, importable in the dev container only (plotly stubbed, oracle/build_ref.py)."""
    if not has_reference():
        pytest.skip("needs /root/reference")
    import build_ref
    build_ref.add_reference_to_path()
    pytest.importorskip("spatialpy")


@pytest.mark.reference
@pytest.mark.skipif(not has_reference(), reason="needs /root/reference")
def test_flattening_matches_reference_codegen_inputs():
    """FlatModel.from_spatialpy reproduces the committed fixture, and the reference's Solver has the surface we mirror."""
    import build_ref
    build_ref.add_reference_to_path()
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import models
    from spatialpy_b200 import FlatModel, Solver
    m = models.cavity2d()
    np.random.seed(12345)
    fm = FlatModel.from_spatialpy(m)
    ref = load_model("cavity2d")
    for k in ("x", "nu", "mass", "rho", "solid", "output_steps"):
        np.testing.assert_array_equal(getattr(fm, k), getattr(ref, k))
    # type INDICES come from iterating a Python set and change with PYTHONHASHSEED (domain.py:141-146): compare by name
    inv_a = {v: k for k, v in fm.type_constants.items()}
    inv_b = {v: k for k, v in ref.type_constants.items()}
    assert [inv_a[t] for t in fm.type] == [inv_b[t] for t in ref.type]
    assert fm.bc_source == ref.bc_source and fm.h == ref.h and fm.nt == ref.nt
    from spatialpy.solvers.solver import Solver as RefSolver
    ref_params = list(inspect.signature(RefSolver.run).parameters)
    assert list(inspect.signature(Solver.run).parameters)[:len(ref_params)] == ref_params
    assert list(inspect.signature(Solver.__init__).parameters) == list(inspect.signature(RefSolver.__init__).parameters)
    assert list(inspect.signature(Solver.compile).parameters) == list(inspect.signature(RefSolver.compile).parameters)


def test_binary_side_store_round_trip(tmp_path):
    """outputN.ssb layout (written by ssb_core.cu write_bin on the GPU box): the Python twin round-trips, read_output prefers it."""
    import numpy as np
    from spatialpy_b200.solver import FlatResult
    from spatialpy_b200.vtk import read_output, read_ssb, write_ssb
    rng = np.random.default_rng(3)
    n = 37
    x, v, scal = rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), rng.random((4, n))
    C, D, typ = rng.random((2, n)), rng.integers(0, 1000, (2, n)), rng.integers(1, 4, n)
    path = tmp_path / "output3.ssb"
    write_ssb(path, x, v, scal, C, typ, D, ["A", "B_long_name"], step=30)
    assert path.stat().st_size % 4 == 0 and (path.stat().st_size - n * (80 + 16 + 4 + 8)) % 64 == 0
    pts, arr = read_ssb(path)
    assert pts.dtype == np.float32 and np.array_equal(pts, x.astype(np.float32))
    assert list(arr) == ["id", "type", "v", "rho", "mass", "bvf_phi", "nu", "C[A]", "C[B_long_name]", "D[A]", "D[B_long_name]",
                         "__nfields_header__"]
    assert np.array_equal(arr["v"], v) and np.array_equal(arr["nu"], scal[3]) and np.array_equal(arr["C[B_long_name]"], C[1])
    assert arr["D[A]"].dtype == np.int64 and np.array_equal(arr["D[A]"], D[0]) and np.array_equal(arr["type"], typ)
    assert arr["__nfields_header__"] == 11
    p2, a2 = read_output(str(tmp_path), 3)
    assert np.array_equal(p2, pts)
    res = FlatResult(None, str(tmp_path))
    assert np.array_equal(res.get_species("A", timepoints=3), D[0])
    with open(path, "r+b") as f:
        f.truncate(path.stat().st_size - 8)
    with pytest.raises(ValueError):
        read_ssb(path)


def test_reference_result_subclass_reads_the_binary_store(tmp_path):
    """With the spatialpy front-end importable the Solver returns the reference's own Result type; only read_step is overridden."""
    _need_spatialpy()
    import numpy as np
    from spatialpy.core.result import Result
    from spatialpy_b200.solver import _result_class
    from spatialpy_b200.vtk import write_ssb
    n = 5
    write_ssb(tmp_path / "output0.ssb", np.zeros((n, 3)), np.ones((n, 3)), np.full((4, n), 2.0), None, np.ones(n, int),
              np.arange(n)[None, :], ["A"], rdme_initialized=0)
    cls = _result_class()
    assert issubclass(cls, Result)
    res = cls(None, str(tmp_path))
    pts, arr = res.read_step(0)
    assert pts.shape == (n, 3) and "__nfields_header__" not in arr and np.array_equal(arr["D[A]"], np.arange(n))
    res.result_dir = None          # keep Result.__del__ away from pytest's tmp_path


def _write_series(tmp_path, n=23, T=6, seed=4):
    import numpy as np
    from spatialpy_b200.vtk import write_ssb
    rng = np.random.default_rng(seed)
    snaps = []
    for t in range(T):
        snap = dict(x=rng.normal(size=(n, 3)), v=rng.normal(size=(n, 3)), scal=rng.random((4, n)) + 0.5, C=rng.random((2, n)),
                    typ=rng.integers(1, 4, n), D=rng.integers(0, 1000, (2, n)))
        write_ssb(tmp_path / f"output{t}.ssb", snap["x"], snap["v"], snap["scal"], snap["C"], snap["typ"], snap["D"],
                  ["A", "B"], step=10 * t)
        snaps.append(snap)
    return snaps


def test_binary_store_field_reads_match_full_reads(tmp_path):
    """read_ssb_field (one array by offset) == the same key of read_ssb, for every key and dtype."""
    import numpy as np
    from spatialpy_b200.vtk import read_ssb, read_ssb_field
    _write_series(tmp_path, T=2)
    path = tmp_path / "output1.ssb"
    pts, arr = read_ssb(path)
    for key, val in arr.items():
        if key.startswith("__"):
            continue
        got = read_ssb_field(path, key)
        assert got.dtype == val.dtype and got.shape == val.shape and np.array_equal(got, val), key
    assert np.array_equal(read_ssb_field(path, "points"), pts)
    for bad in ("C[Z]", "D[]", "nope", "E[A]"):
        with pytest.raises(KeyError):
            read_ssb_field(path, bad)


def test_all_timepoint_getters_from_the_binary_store(tmp_path):
    """FlatResult.get_species / get_property: the reference's argument handling and shapes (result.py:334-402,601-655)."""
    import numpy as np
    from spatialpy_b200.solver import FlatResult
    snaps = _write_series(tmp_path)
    res = FlatResult(None, str(tmp_path))
    allA = res.get_species("A")
    assert allA.dtype == np.float64 and allA.shape == (6, 23)
    assert np.array_equal(allA, np.array([s["D"][0] for s in snaps], dtype=float))
    assert np.array_equal(res.get_species("B", timepoints=4), snaps[4]["D"][1].astype(float))
    assert np.array_equal(res.get_species("B", timepoints=[1, 5]), np.array([snaps[1]["D"][1], snaps[5]["D"][1]], dtype=float))
    assert np.array_equal(res.get_species("A", timepoints=slice(2, 4)), allA[2:4])
    assert np.array_equal(res.get_species("A", deterministic=True), np.array([s["C"][0] for s in snaps]))
    conc = res.get_species("A", timepoints=2, concentration=True)
    assert np.array_equal(conc, snaps[2]["D"][0] / (snaps[2]["scal"][1] / snaps[2]["scal"][0]))
    with pytest.raises(Exception):
        res.get_species("A", timepoints=1.0)
    v = res.get_property("v")
    assert v.shape == (6, 23, 3) and np.array_equal(v[3], snaps[3]["v"])
    assert np.array_equal(res.get_property("rho"), np.array([s["scal"][0] for s in snaps]))
    # result.py:648 reads step `ndx` (the position in the selection), so a single timepoint always returns step 0
    assert np.array_equal(res.get_property("nu", timepoints=4), snaps[0]["scal"][3])


def test_reference_result_getters_agree_with_the_inherited_ones(tmp_path):
    """B200Result.get_species/get_property from the binary store == the reference's own methods run over read_step."""
    _need_spatialpy()
    import numpy as np
    from spatialpy.core.result import Result
    from spatialpy_b200.solver import _result_class

    class _Dom:
        def get_num_voxels(self):
            return 23

    class _Model:
        tspan = np.arange(6.0)
        domain = _Dom()
        listOfSpecies = {"A": None, "B": None}

    _write_series(tmp_path)
    res = _result_class()(_Model(), str(tmp_path))
    for kw in ({}, {"timepoints": 3}, {"timepoints": [0, 5]}, {"concentration": True}, {"deterministic": True, "timepoints": 1}):
        fast = res.get_species("B", **kw)
        slow = Result.get_species(res, "B", **kw)          # the inherited loop over read_step (binary-backed, full snapshots)
        assert fast.shape == slow.shape and fast.dtype == slow.dtype and np.array_equal(fast, slow), kw
    for name in ("v", "rho", "mass", "type", "bvf_phi", "nu"):
        for kw in ({}, {"timepoints": 2}):
            fast, slow = res.get_property(name, **kw), Result.get_property(res, name, **kw)
            assert fast.shape == slow.shape and np.array_equal(fast, slow), (name, kw)
    with pytest.raises(Exception, match="not found"):
        res.get_species("Z")
    res.result_dir = None          # keep Result.__del__ away from pytest's tmp_path


def test_domain_from_arrays_equals_the_add_point_loop():
    """builders.domain_from_arrays == Domain.add_point per particle (domain.py:203-255), attribute for attribute, and the
    flattened model the engine sees is the same."""
    _need_spatialpy()
    import numpy as np
    import spatialpy
    from spatialpy.core.spatialpyerror import DomainError
    from spatialpy_b200 import FlatModel
    from spatialpy_b200.builders import domain_from_arrays
    rng = np.random.default_rng(7)
    n = 200
    pts = rng.random((n, 3))
    vol = rng.random(n) + 0.5
    mass = rng.random(n) + 0.5
    tid = np.where(pts[:, 0] < 0.3, "Left", np.where(pts[:, 0] > 0.7, "Right_side", "Mid")).astype(object)
    nu = rng.random(n)
    fixed = pts[:, 2] < 0.2
    lim = dict(xlim=(0, 1), ylim=(0, 1), zlim=(0, 1))
    slow = spatialpy.Domain(0, **lim, rho0=2.0, c0=5, gravity=[0, 0, -1])
    for i in range(n):
        slow.add_point(pts[i], vol=vol[i], mass=mass[i], type_id=tid[i], nu=nu[i], fixed=fixed[i])
    fast = domain_from_arrays(pts, type_id=tid, vol=vol, mass=mass, nu=nu, fixed=fixed, rho0=2.0, c0=5, gravity=[0, 0, -1], **lim)
    for attr in ("vertices", "vol", "mass", "nu", "c", "rho", "fixed", "type_id"):
        a, b = getattr(slow, attr), getattr(fast, attr)
        assert a.shape == b.shape and a.dtype == b.dtype and (a == b).all(), attr
    assert (slow.P0, slow.rho0, slow.c0, slow.gravity) == (fast.P0, fast.rho0, fast.c0, fast.gravity)
    assert fast.get_num_voxels() == n and abs(fast.find_h() - slow.find_h()) == 0.0

    def model(dom):
        m = spatialpy.Model("m")
        m.add_domain(dom)
        a = spatialpy.Species("A", diffusion_coefficient=0.1, restrict_to=["Left", "Mid"])
        m.add_species(a)
        k = spatialpy.Parameter("k", expression=2.0)
        m.add_parameter(k)
        m.add_reaction(spatialpy.Reaction(name="r", reactants={}, products={"A": 1}, rate="k", restrict_to="Left"))
        m.add_initial_condition(spatialpy.PlaceInitialCondition(a, 50, [0.5, 0.5, 0.5]))
        m.timespan(spatialpy.TimeSpan.linspace(t=1, num_points=3, timestep_size=0.5))
        return m
    fa, fb = FlatModel.from_spatialpy(model(slow)), FlatModel.from_spatialpy(model(fast))
    for name in FlatModel._ARRAYS:
        assert np.array_equal(getattr(fa, name), getattr(fb, name)), name
    assert fa.type_constants == fb.type_constants and fa.h == fb.h
    # scalars broadcast, 2-D points get z = 0, the reference's validation errors are raised
    d2 = domain_from_arrays(rng.random((10, 2)), type_id=3, mass=2.0, vol=4.0)
    assert d2.vertices.shape == (10, 3) and not d2.vertices[:, 2].any() and (d2.rho == 0.5).all() and d2.type_id[0] == "type_3"
    for bad in (dict(vol=-1.0), dict(type_id=0), dict(type_id="a b"), dict(type_id="a-b")):
        with pytest.raises(DomainError):
            domain_from_arrays(pts, **bad)


def test_install_adds_the_solver_keyword_to_model_run():
    """spatialpy_b200.install(): `Model.run(solver=...)` (the keyword the reference's README promises, model.py:1021-1056) hands
    the reference's own arguments to the given solver class or instance; without it the reference's path runs untouched."""
    _need_spatialpy()
    import spatialpy
    from spatialpy.core.model import Model
    import spatialpy_b200
    orig = Model.run
    try:
        spatialpy_b200.install()
        patched = Model.run
        assert patched is not orig and getattr(patched, "_ssb_patched", False)
        spatialpy_b200.install()                       # idempotent
        assert Model.run is patched
        assert spatialpy.B200Solver is spatialpy_b200.Solver
        calls = []

        class Dummy:
            def __init__(self, model, debug_level=0):
                calls.append(("init", model, debug_level))

            def run(self, **kw):
                calls.append(("run", kw))
                return "result"

        m = spatialpy.Model("m")
        assert m.run(solver=Dummy, number_of_trajectories=3, seed=7, timeout=5, debug_level=2, devices=[0, 1]) == "result"
        assert calls[0] == ("init", m, 2)
        assert calls[1] == ("run", dict(number_of_trajectories=3, seed=7, timeout=5, number_of_threads=None, debug=False,
                                        profile=False, devices=[0, 1]))
        inst = Dummy(m)
        calls.clear()
        m.run(solver=inst, seed=1)
        assert [c[0] for c in calls] == ["run"] and calls[0][1]["seed"] == 1
        # no solver keyword: the reference's own Model.run runs (an empty model fails inside the reference's compile_prep)
        with pytest.raises(Exception) as info:
            m.run()
        assert "reference/spatialpy" in str(info.traceback[-1].path) and not calls[1:]      # raised inside the reference
    finally:
        Model.run = orig


def _parse_reference_vtk(text):
    """Full-precision parse of a reference output file: float64 points (read_vtk keeps the reader's float32) + arrays."""
    lines = text.split("\n")
    n = int(lines[4].split()[1])
    vals, k = [], 5
    while len(vals) < 3 * n:
        vals.extend(lines[k].split())
        k += 1
    return np.array(vals, dtype=np.float64).reshape(n, 3)


@pytest.mark.parametrize("name", ["output0.vtk", "output1.vtk", "output10.vtk"])
def test_python_vtk_writer_reproduces_the_reference_files_byte_for_byte(tmp_path, name):
    """vtk.write_vtk (host-assembled snapshots of slab runs) against files the REFERENCE engine wrote (tests/golden/
    vtk_diffusion3d, made by make_golden.py): parse -> rewrite -> identical bytes, FIELD undercount of output0 included."""
    import gzip
    from spatialpy_b200.vtk import read_vtk, write_bounding_box, write_vtk
    src = os.path.join(ROOT, "tests", "golden", "vtk_diffusion3d", name + ".gz")
    text = gzip.open(src, "rt", encoding="ascii").read()
    raw = tmp_path / "ref.vtk"
    raw.write_text(text)
    x = _parse_reference_vtk(text)
    _, arr = read_vtk(str(raw))
    species = ["A", "B"]
    scal = np.stack([arr[k] for k in ("rho", "mass", "bvf_phi", "nu")])
    C = np.stack([arr[f"C[{s}]"] for s in species])
    D = np.stack([arr[f"D[{s}]"] for s in species])
    out = tmp_path / "mine.vtk"
    write_vtk(str(out), x, arr["v"], scal, C, arr["type"], D, species, rdme_initialized=int(arr["__nfields_header__"] == 11))
    assert out.read_bytes() == text.encode("ascii")
    if name == "output0.vtk":
        assert arr["__nfields_header__"] == 9                       # output.cpp:151-154
        bb = gzip.open(os.path.join(ROOT, "tests", "golden", "vtk_diffusion3d", "output0_boundingBox.vtk.gz"), "rt").read()
        lims = [tuple(float(t) for t in bb.split("\n")[k].split()) for k in (6, 8, 10)]
        write_bounding_box(str(tmp_path), *lims)
        assert (tmp_path / "output0_boundingBox.vtk").read_text() == bb


def test_python_vtk_writer_line_layout_for_every_remainder(tmp_path):
    """Particle counts that are not multiples of 3 or 9, negative values, more particles than one write chunk."""
    from spatialpy_b200.vtk import read_vtk, write_vtk
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 8, 9, 10, 28, 90001):
        x, v = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
        scal, C, D = rng.random((4, n)), -rng.random((1, n)), rng.integers(0, 5000, (1, n))
        typ = rng.integers(1, 4, n)
        p = tmp_path / f"o{n}.vtk"
        write_vtk(str(p), x, v, scal, C, typ, D, ["S"])
        text = p.read_text()
        body = text.split("\n")
        assert body[4] == f"POINTS {n} float" and len(body[5].split()) == min(9, 3 * n)
        pts, arr = read_vtk(str(p))
        assert np.array_equal(pts, x.astype(np.float32)) or np.allclose(pts, x, rtol=1e-6)
        assert np.array_equal(arr["type"], typ) and np.array_equal(arr["D[S]"], D[0]) and np.array_equal(arr["id"], np.arange(n))
        assert np.allclose(arr["v"], v, atol=5e-7) and np.allclose(arr["C[S]"], C[0], atol=5e-7) and arr["__nfields_header__"] == 9
        i = body.index(f"id 1 {n} int")
        assert len(body[i + 1].split()) == min(9, n) and all(len(l.split()) <= 9 for l in body[i + 1:i + 1 + (n + 8) // 9])


@pytest.mark.parametrize("mesh,sub", [("cylinder.xml", None), ("GRN_Spatial.mesh.xml", "GRN_Spatial.subdomains.txt")])
def test_fast_xml_mesh_reader_equals_the_reference_reader(mesh, sub):
    """builders.read_xml_mesh == Domain.read_xml_mesh (domain.py:1148-1180) on the meshes the reference ships: every Domain
    attribute identical, volumes to one ulp (the reference's 3-vector dot goes through BLAS)."""
    _need_spatialpy()
    import spatialpy
    from spatialpy_b200.builders import read_xml_mesh
    base = os.path.join(os.environ.get("SSB_REFERENCE_ROOT", "/root/reference"), "examples", "Domain_Files")
    kw = {} if sub is None else {"subdomain_file": os.path.join(base, sub)}
    ref = spatialpy.Domain.read_xml_mesh(os.path.join(base, mesh), **kw)
    fast = read_xml_mesh(os.path.join(base, mesh), **kw)
    for key, val in vars(ref).items():
        got = getattr(fast, key)
        if key in ("vol", "mass", "tetrahedron_vol"):
            np.testing.assert_allclose(got, val, rtol=4e-16 * 8, atol=0)
        elif isinstance(val, np.ndarray):
            assert got.dtype == val.dtype and got.shape == val.shape and (got == val).all(), key
        elif key != "actions":
            assert got == val, key
    assert abs(fast.find_h() - ref.find_h()) == 0.0


def _same_domain(fast, ref, ulp_keys=("vol", "mass", "tetrahedron_vol", "rho")):
    for key, val in vars(ref).items():
        got = getattr(fast, key)
        if key in ulp_keys and val is not None:
            np.testing.assert_allclose(got, val, rtol=4e-16 * 8, atol=0, err_msg=key)
        elif isinstance(val, np.ndarray):
            assert got.dtype == val.dtype and got.shape == val.shape and (got == val).all(), key
        elif key != "actions":
            assert got == val, key


def test_fast_msh_reader_equals_the_reference_paths():
    """builders.read_msh_file (Gmsh ASCII 2.2 / 4.1, no meshio needed) against the reference two ways: (i) the 2.2 mesh the
    reference ships in both formats — coli.msh must give the Domain that the reference's OWN xml reader gives for its dolfin twin
    (vertices, tetrahedra, volumes); (ii) the 4.1 cube and the 2.2 hes1 cell: the parsed points / cell blocks handed to the
    reference's `Domain.import_meshio_object` (domain.py:865-898, `MeshIOLattice.apply` lattice.py:732-806: first triangle block,
    first tetra block, one add_point per node) give the same Domain attribute for attribute."""
    _need_spatialpy()
    import types
    import spatialpy
    from spatialpy_b200.builders import parse_msh, read_msh_file
    root = os.environ.get("SSB_REFERENCE_ROOT", "/root/reference")
    base = os.path.join(root, "examples", "Under_Construction", "Domain_Files")
    twin = spatialpy.Domain.read_xml_mesh(os.path.join(base, "mesh", "coli.xml"))
    fast = read_msh_file(os.path.join(base, "coli.msh"))
    assert (fast.vertices == twin.vertices).all() and (fast.tetrahedrons == twin.tetrahedrons).all()
    np.testing.assert_allclose(fast.vol, twin.vol, rtol=4e-16 * 8, atol=0)
    assert fast.triangles.shape == (1492, 3) and fast.triangles.max() < len(fast.vertices)
    for path in (os.path.join(root, "examples", "Tests", "Mesh_Files", "cube.msh"),):
        pts, cells = parse_msh(path)
        mesh = types.SimpleNamespace(points=pts, cells=[types.SimpleNamespace(type=name, data=conn) for name, conn in cells])
        ref = spatialpy.Domain.import_meshio_object(mesh)
        _same_domain(read_msh_file(path), ref)
        from spatialpy_b200.builders import import_meshio_object
        sub = os.path.join(root, "examples", "Domain_Files", "GRN_Spatial.subdomains.txt")      # index,type lines; sticky types
        _same_domain(import_meshio_object(mesh, subdomain_file=sub), spatialpy.Domain.import_meshio_object(mesh, subdomain_file=sub))
        assert abs(ref.vol.sum() - 1.0) < 1e-12                  # the unit cube
    with pytest.raises(spatialpy.core.spatialpyerror.LatticeError):
        parse_msh(os.path.join(base, "mesh", "coli.xml"))


def test_fast_stochss_domain_reader_equals_the_reference_reader(tmp_path):
    """builders.read_stochss_domain == Domain.read_stochss_domain (domain.py:1098-1118, `StochSSLattice.apply` lattice.py:845-903)
    on a .domn file and on the same domain wrapped in a .smdl model file: optional `rho` / `c`, '-' stripped from type names,
    domain-wide constants, limits; a file without the expected keys raises the reference's LatticeError."""
    _need_spatialpy()
    import json
    import spatialpy
    from spatialpy_b200.builders import read_stochss_domain
    rng = np.random.default_rng(8)
    parts = []
    for k in range(57):
        p = {"point": rng.normal(size=3).tolist(), "type": int(k % 3), "volume": float(rng.uniform(0.5, 2)), "mass": float(rng.uniform(0.5, 2)),
             "nu": float(rng.random()), "fixed": bool(k % 5 == 0)}
        if k % 2:
            p["rho"] = float(rng.uniform(0.9, 1.1))
        if k % 4 == 0:
            p["c"] = float(rng.uniform(5, 15))
        if k % 7 == 0:
            p["rho"] = None
        parts.append(p)
    dom = {"rho_0": 1.5, "c_0": 12.0, "p_0": 3.0, "gravity": [0, -9.8, 0], "particles": parts,
           "types": [{"typeID": 0, "name": "Un-Assigned"}, {"typeID": 1, "name": "Cyto-plasm"}, {"typeID": 2, "name": "Membrane"}]}
    for name, body in (("d.domn", dom), ("m.smdl", {"name": "model", "domain": dom})):
        path = tmp_path / name
        path.write_text(json.dumps(body))
        _same_domain(read_stochss_domain(str(path)), spatialpy.Domain.read_stochss_domain(str(path)), ulp_keys=())
    bad = tmp_path / "bad.domn"
    bad.write_text(json.dumps({"particles": []}))
    with pytest.raises(spatialpy.core.spatialpyerror.LatticeError):
        read_stochss_domain(str(bad))


def test_header_is_valid_c_and_matches_the_ctypes_binding(tmp_path):
    """include/ssb.h + include/ssb_peaks.h compile as pedantic C11 and link from a plain-C program (tests/c/abi_smoke.c); the C
    struct the header declares has the size of the ctypes mirror in spatialpy_b200/engine.py; argument errors are return codes."""
    from spatialpy_b200 import codegen, engine
    _lib()
    codegen.build_peaks()
    exe = tmp_path / "abi_smoke"
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", str(exe), "-L", codegen.LIB_DIR, "-lssb_core", "-lssb_peaks",
           f"-Wl,-rpath,{codegen.LIB_DIR}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    m = re.search(r"abi (\d+), sizeof\(ssb_model\) (\d+), device_count rc (\d+)", run.stdout)
    assert m and int(m.group(1)) == engine.SSB_ABI_VERSION
    assert int(m.group(2)) == ctypes.sizeof(engine.SsbModel)
    offs = re.search(r"offsets (.*)", run.stdout).group(1).split()
    for name, off in zip(offs[0::2], offs[1::2]):
        assert getattr(engine.SsbModel, name).offset == int(off), name
    if not os.path.exists("/dev/nvidia0"):
        assert int(m.group(3)) == 3                      # SSB_ERR_CUDA: no device, reported as a code, not a crash
    # the engine's file writers driven from C (ssb_write_snapshot): the Python readers see what the C program wrote
    from spatialpy_b200.vtk import read_ssb, read_vtk
    out = tmp_path / "snap"
    out.mkdir()
    assert subprocess.run([str(exe), "snapshot", str(out)]).returncode == 0
    pts, arr = read_vtk(str(out / "output1.vtk"))
    assert pts.shape == (3, 3) and arr["D[A]"].tolist() == [7, 0, 42] and arr["type"].tolist() == [1, 2, 1] and arr["nu"].tolist() == [2.0] * 3
    _, arb = read_ssb(str(out / "output1.ssb"))
    assert arb["C[A]"].tolist() == [0.5, 1.5, 2.5] and arb["mass"].tolist() == [0.1] * 3


def test_converted_expressions_evaluate_like_python():
    """The reference's own expression test (test/integration_tests/test_solver.py:76-121,202-256: BuildExpression's C text must
    evaluate like Python to 3 places) replayed on THIS code generator: every converted expression becomes a reaction propensity,
    the generated `ssb_gen` functions are compiled for the host (the same text nvcc compiles for sm_100a — also checked) and
    evaluated at the reference's value sets."""
    _need_spatialpy()
    import ctypes as C
    import nsm_oracle
    from spatialpy.core.model import Model  # noqa: F401  (front-end importable)
    from spatialpy.solvers.build_expression import BuildExpression, ExpressionConverter
    from spatialpy_b200 import FlatModel, ReactionSource, codegen
    numeric = [({"x": 0}, "x*2", [[0.0], [1.0], [-1.0], [9.999], [-9.999]]),
               ({"x": 0}, "x*2 + x/2 - (x*3)^2 + x/3^2", [[0.0], [1.0], [-1.0], [3.333], [-3.333], [9.8765], [-9.8765]]),
               ({"x": 0, "y": 1}, "(x-1)*y^2+x", [[1.0, 2.4], [5.1, 0.0], [5.1, 1.0], [5.1, -1.0], [9.8765, -1.0], [-1.0, 9.8765]]),
               ({"x": 0, "y": 1, "z": 2}, "(x^2/y^2/z^2)/x^2/y^2/z^2**1/x**1/y**1/z",
                [[5.1, 0.1, 2.0], [0.1, 5.1, 2.0], [2.0, 0.1, 5.1], [2.0, 5.1, 0.1]])]
    boolean = [({"x": 0}, "x > 0", [[100], [0], [0.001], [-1]]),
               ({"x": 0, "y": 1}, "x > y", [[100, 99], [99, 100], [-10, 10], [10, -10], [0.001, 0.0], [0.0, 0.001], [-99.999, -99.998]]),
               ({"x": 0, "y": 1}, "x > 0 and y < x", [[100, 99], [99, 100], [0, -100], [-0.001, -99.0], [0, 0.001], [-0.001, 0]]),
               ({"x": 0, "y": 1}, "x > 0 and y < 10 and x > y",
                [[100, 9], [0.01, 0.00], [100, 200], [0.01, 0.02], [0, 0], [-0.01, -0.02], [-0.01, 0]]),
               ({"x": 0, "y": 1}, "x > 0 and y < 10 or y > 100", [[10, 9], [0.01, 9.99], [0, 10], [-1.0, -1.0]]),
               ({"x": 0, "y": 1, "z": 2}, "x^2>x and y<y^2 or z^2!=z^3 and y!=z",
                [[1.0, 1.0, 1.0], [99.9, 99.9, 100.0], [0.0, -1.0, 99.9], [-1.0, -1.0, 0.00]])]
    cases = [(a, e, v, False) for a, e, v in numeric] + [(a, e, v, True) for a, e, v in boolean]
    reactions, pyfuncs = [], []
    for k, (args, expr, _, _) in enumerate(cases):
        converted = ExpressionConverter.convert_str(expr)
        text = BuildExpression(namespace={name: f"data_fn[{slot}]" for name, slot in args.items()}, sanitize=True).getexpr_cpp(converted)
        reactions.append(ReactionSource(name=f"e{k}", propensity=text, ode_propensity=text, restrict_to=None))
        pyfuncs.append(eval(f"lambda {','.join(args)}: {converted}"))
    n = 2
    fm = FlatModel(name="expr", x=np.array([[0.0, 0, 0], [0.1, 0, 0]]), type=np.ones(n, np.int32), nu=np.ones(n), mass=np.ones(n),
                   c=np.zeros(n), rho=np.ones(n), solid=np.ones(n, np.int32), species_names=["A"], reactions=reactions,
                   u0=np.ones((n, 1), np.uint32), N_dense=np.zeros((1, len(reactions)), np.int32),
                   diffusion_matrix=np.zeros((1, 1)), data_fn=np.zeros((3, n)), static_domain=True, dt=1.0, nt=1,
                   output_steps=np.array([0, 1], np.uint32), h=0.25, dimension=1).finalize()
    assert os.path.exists(codegen.build_model_unit(fm))          # the same text compiles as sm_100a device code
    lib = nsm_oracle.build(fm)
    out = np.zeros(len(reactions))
    x = np.ones(1, np.int32)
    for k, (args, expr, values, is_bool) in enumerate(cases):
        for vals in values:
            df = np.zeros(3)
            df[:len(vals)] = vals
            lib.nsm_oracle_eval_propensities(x.ctypes.data_as(C.c_void_p), C.c_double(0.0), C.c_double(1.0),
                                             df.ctypes.data_as(C.c_void_p), C.c_int(1), out.ctypes.data_as(C.c_void_p))
            want = pyfuncs[k](*vals)
            if is_bool:
                assert bool(out[k]) == bool(want), (expr, vals, out[k])
            else:
                assert abs(out[k] - want) < 5e-4 * max(1.0, abs(want)) or round(out[k] - want, 3) == 0, (expr, vals, out[k], want)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the arm the driver runs beside ours): one JSON line with the contract's keys, timed on the
    unmodified reference executable built from /root/reference (oracle/_ref); skipped where that binary does not exist."""
    import json
    import sys
    exe = os.path.join(ROOT, "oracle", "_ref", "bench_cylinder", "fast", "ssa_sdpd.exe")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref is not built here")
    run = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [l for l in run.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "particle-steps/s" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_bench_helper_measurements_degrade_to_an_error_field():
    """The fp64 peak helper runs in a child process and reports failure as data (it must never cost the bench line)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    res = bench.fp64_peak_sample(0)
    if os.path.exists("/dev/nvidia0"):
        assert res.get("fp64_tflops", 0) > 1.0
    else:
        assert set(res) == {"error"} and "return code = 3" in res["error"]
    from spatialpy_b200 import configs
    fm = configs.tank_sdpd(n=12, nt=10, output_every=10)
    assert bench.algorithmic_bytes(fm, True) == 698 + 64 * fm.num_chem_species + (68 + 12 * fm.num_stoch_species + 8 * fm.num_stoch_rxns)


@pytest.mark.parametrize("name", ["output0.vtk", "output1.vtk", "output10.vtk"])
def test_engine_writers_on_host_snapshots_reproduce_the_reference_files(tmp_path, name):
    """ssb_write_snapshot = the engine's own C++ writers fed from caller memory (no device): byte for byte the files the
    REFERENCE wrote (tests/golden/vtk_diffusion3d), bounding box included; the binary side-store equals the Python twin's."""
    import gzip
    from spatialpy_b200.vtk import read_vtk, write_snapshot, write_ssb
    gold = os.path.join(ROOT, "tests", "golden", "vtk_diffusion3d")
    text = gzip.open(os.path.join(gold, name + ".gz"), "rt", encoding="ascii").read()
    raw = tmp_path / "ref.vtk"
    raw.write_text(text)
    x = _parse_reference_vtk(text)
    _, arr = read_vtk(str(raw))
    species = ["A", "B"]
    scal = np.stack([arr[k] for k in ("rho", "mass", "bvf_phi", "nu")])
    C = np.stack([arr[f"C[{s}]"] for s in species])
    D = np.stack([arr[f"D[{s}]"] for s in species])
    k = int(name[6:-4])
    init = int(arr["__nfields_header__"] == 11)
    bb = gzip.open(os.path.join(gold, "output0_boundingBox.vtk.gz"), "rt").read()
    lims = [tuple(float(t) for t in bb.split("\n")[q].split()) for q in (6, 8, 10)]
    out = tmp_path / "run"
    out.mkdir()
    write_snapshot(str(out), k, x, arr["v"], scal, C, arr["type"], D, species, lims, step=k, rdme_initialized=init, vtk=True, binary=True)
    assert (out / name).read_bytes() == text.encode("ascii")
    if k == 0:
        assert (out / "output0_boundingBox.vtk").read_text() == bb
    else:
        assert not (out / "output0_boundingBox.vtk").exists()
    twin = tmp_path / "twin.ssb"
    write_ssb(str(twin), x, arr["v"], scal, C, arr["type"], D, species, step=k, rdme_initialized=init)
    assert (out / f"output{k}.ssb").read_bytes() == twin.read_bytes()
    with pytest.raises(OSError):
        write_snapshot(str(tmp_path / "missing_dir"), 1, x, arr["v"], scal, C, arr["type"], D, species, lims)


def test_engine_writers_large_snapshot_threads_and_python_twin_agree(tmp_path, monkeypatch):
    """>= 200 k particles are formatted by several host threads over line-aligned ranges: the bytes must not depend on the thread
    count and must equal the independent Python twin (vtk.write_vtk)."""
    from spatialpy_b200.vtk import write_snapshot, write_vtk
    rng = np.random.default_rng(5)
    n = 200_003
    x, v = rng.normal(size=(n, 3)), rng.normal(size=(n, 3)) * 1e-3
    scal, C = rng.random((4, n)), -rng.random((1, n))
    D, typ = rng.integers(0, 100000, (1, n)), rng.integers(1, 4, n)
    lims = [(-1.0, 1.0), (0.0, 2.0), (0.0, 0.0)]
    outs = []
    for threads in ("1", "7", None):
        d = tmp_path / f"t{threads}"
        d.mkdir()
        if threads is None:
            monkeypatch.delenv("SSB_VTK_THREADS", raising=False)
        else:
            monkeypatch.setenv("SSB_VTK_THREADS", threads)
        write_snapshot(str(d), 3, x, v, scal, C, typ, D, ["S"], lims, step=30)
        outs.append((d / "output3.vtk").read_bytes())
    assert outs[0] == outs[1] == outs[2]
    write_vtk(str(tmp_path / "twin.vtk"), x, v, scal, C, typ, D, ["S"])
    assert (tmp_path / "twin.vtk").read_bytes() == outs[0]


# ----------------------------------------------------------------------------------------------------------------------
# the moving-domain force sweep run ON THE HOST: its CUDA source executed by a block emulator (tests/cuda_emu) — the default record
# layout (SSB_REC2: the neighbour's 1/rho, P/rho^2, C from one derived 32-byte sector written by the predictor) against the layout
# that recomputes / gathers them per pair, and against a plain numpy evaluation of the reference's pair formulas
# ----------------------------------------------------------------------------------------------------------------------
def _build_force_emulator(fm, tmp_path, defines=()):
    from spatialpy_b200 import codegen
    src = open(os.path.join(codegen.CSRC, "ssb_model_unit.cuh")).read()
    a = src.index("// per-particle state of the sweep + the pair body")
    b = src.index("// ---------------------------------------------------------------------------------------------\n// Static-domain fast path")
    kern = tmp_path / "kernels.inc"
    kern.write_text(src[a:b])
    hdr = tmp_path / "model.h"
    hdr.write_text(codegen.generate_model_header(fm))
    so = tmp_path / ("force_emu" + "".join(d.replace("=", "_") for d in defines) + ".so")
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(codegen.nvcc_path())), "include")
    cmd = ["g++", "-std=c++20", "-O1", "-w", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-I", cuda_inc, "-I", codegen.CSRC,
           "-I", os.path.join(ROOT, "tests", "cuda_emu"), f"-DSSB_MODEL_HEADER=\"{hdr}\"", f"-DEMU_KERNELS=\"{kern}\""]
    cmd += [f"-D{d}" for d in defines] + [os.path.join(ROOT, "tests", "cuda_emu", "force_emu.cpp"), "-o", str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return ctypes.CDLL(str(so))


class _EmuArgs(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("dim", ctypes.c_int), ("num_types", ctypes.c_int), ("filter", ctypes.c_int),
                ("flags", ctypes.c_uint), ("dt", ctypes.c_double), ("h", ctypes.c_double), ("rho0", ctypes.c_double),
                ("P0", ctypes.c_double), ("rec", ctypes.c_void_p), ("nbr", ctypes.c_void_p), ("nbr_count", ctypes.c_void_p),
                ("nbr_cap", ctypes.c_int), ("owned", ctypes.c_void_p), ("F", ctypes.c_void_p * 3), ("Fbp", ctypes.c_void_p * 3),
                ("Frho", ctypes.c_void_p), ("C", ctypes.c_void_p), ("Q", ctypes.c_void_p), ("Ddiag", ctypes.c_void_p),
                ("data_fn", ctypes.c_void_p), ("dmat", ctypes.c_void_p), ("max_bits", ctypes.c_void_p),
                ("rec2", ctypes.c_void_p)]


def _force_sweep_inputs(fm, seed=2):
    """A mid-trajectory state in the engine's layout: cell-sorted storage order, ascending candidate lists of radius 1.1 h
    (ELL, transposed), 128-byte gather records, a few ghost particles."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    n = fm.num_particles
    rad = fm.h * 1.1
    x0 = fm.x + rng.uniform(-0.004, 0.004, fm.x.shape)
    cell = np.floor((x0 - x0.min(axis=0)) / rad).astype(np.int64)
    nc = cell.max(axis=0) + 1
    key = (cell[:, 2] * nc[1] + cell[:, 1]) * nc[0] + cell[:, 0]
    order = np.argsort(key, kind="stable")
    x0 = x0[order]
    cell_start = np.searchsorted(key[order], np.arange(int(nc.prod()) + 1)).astype(np.int64)      # [ncells + 1], last = n
    typ, solid, mass, nu = fm.type[order], fm.solid[order], fm.mass[order], fm.nu[order]
    xl = x0 + rng.uniform(-1e-4, 1e-4, x0.shape) * (solid == 0)[:, None]       # live positions after the predictor
    v = rng.normal(size=(n, 3)) * 0.1 * (solid == 0)[:, None]
    vt = v + rng.normal(size=(n, 3)) * 0.01
    rho = 1.0 + 0.02 * rng.random(n)
    bits = (order.astype(np.int64) & 0xffffffff) | (typ.astype(np.int64) << 32) | (solid.astype(np.int64) << 48)
    rec = np.zeros((n, 16))
    rec[:, 0:3], rec[:, 3:6], rec[:, 6:9], rec[:, 9:12] = x0, xl, v, vt
    rec[:, 12], rec[:, 13], rec[:, 14] = rho, mass, nu
    rec[:, 15] = bits.view(np.float64)
    lists = [sorted(l) for l in cKDTree(x0).query_ball_point(x0, rad)]
    cnt = np.array([len(l) for l in lists], np.int32)
    cap = int(cnt.max())
    nbr = np.zeros((cap, n), np.int32)
    for i, l in enumerate(lists):
        nbr[:len(l), i] = l
    owned = np.ones(n, np.int32)
    owned[rng.choice(n, 9, replace=False)] = 0
    Sc, Sd = fm.num_chem_species, fm.num_stoch_species
    Sc0 = fm.num_chem_species
    Cfull = rng.random((max(Sc0, 1), n))
    rec2 = np.zeros((n, 4))
    rec2[:, 0] = 1.0 / rho
    rec2[:, 1] = (fm.P0 * (rho / fm.rho0 - 1.0)) * rec2[:, 0] * rec2[:, 0]
    for s_ in range(min(Sc0, 2)):
        rec2[:, 2 + s_] = Cfull[s_]
    st = dict(rec2=rec2, rec=rec, nbr=nbr, cnt=cnt, cap=cap, owned=owned, F=rng.normal(size=(3, n)), Fbp=rng.normal(size=(3, n)),
              Frho=rng.normal(size=n), C=Cfull, Q=rng.normal(size=(max(Sc, 1), n)),
              dmat=np.ascontiguousarray(fm.diffusion_matrix, dtype=np.float64), Sd=Sd, n=n)
    return st


def _run_force_emulator(lib, fm, st, flags=8):
    n = st["n"]
    out = {k: np.ascontiguousarray(st[k]).copy() for k in ("F", "Fbp", "Frho", "Q")}
    out["Ddiag"] = np.full((max(st["Sd"], 1), n), -1.0)
    mb = np.zeros(1, np.uint64)
    keep = [np.ascontiguousarray(st[k]) for k in ("rec", "nbr", "cnt", "owned", "C", "dmat")]
    a = _EmuArgs()
    a.N, a.dim, a.num_types, a.filter, a.flags = n, fm.dimension, fm.num_types, 1, flags
    a.dt, a.h, a.rho0, a.P0 = fm.dt, fm.h, fm.rho0, fm.P0
    a.rec, a.nbr, a.nbr_count, a.nbr_cap, a.owned = keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, st["cap"], keep[3].ctypes.data
    for d in range(3):
        a.F[d] = out["F"][d].ctypes.data
        a.Fbp[d] = out["Fbp"][d].ctypes.data
    a.Frho, a.C, a.Q, a.Ddiag = out["Frho"].ctypes.data, keep[4].ctypes.data, out["Q"].ctypes.data, out["Ddiag"].ctypes.data
    a.data_fn, a.dmat, a.max_bits = None, keep[5].ctypes.data, mb.ctypes.data
    r2 = np.ascontiguousarray(st["rec2"])
    keep.append(r2)
    a.rec2 = r2.ctypes.data
    assert lib.emu_force(ctypes.byref(a), 3) == 0
    out["max_bits"] = mb
    return out


@pytest.mark.parametrize("which", ["tank", "cavity2d_rdme"])
def test_force_sweep_source_record_layouts_agree_on_the_host(tmp_path, which):
    """k_force_mv as CUDA SOURCE run by the block emulator, compiled with the derived sector (SSB_REC2 = 1, default for S_c <= 2)
    and without it (SSB_REC2 = 0: 1/rho_j and P_j/rho_j^2 recomputed per pair, C_j gathered from V.C): same pair formulas, the
    outputs agree to rounding (1e-13 of each field's scale) on a 3-D and on a 2-D moving model."""
    from spatialpy_b200 import configs
    fm = configs.tank_sdpd(n=14, nt=10, output_every=10) if which == "tank" else load_model("cavity2d_rdme")
    st = _force_sweep_inputs(fm, seed=2 if which == "tank" else 7)
    a = _run_force_emulator(_build_force_emulator(fm, tmp_path), fm, st)
    b = _run_force_emulator(_build_force_emulator(fm, tmp_path, ("SSB_REC2=0",)), fm, st)
    assert np.abs(a["F"] - st["F"]).max() > 0 and (a["Ddiag"][:, st["owned"] == 1] >= 0).all()   # the sweep did something
    for k in ("F", "Fbp", "Frho", "Q", "Ddiag"):
        scale = max(float(np.abs(b[k]).max()), 1e-300)
        assert float(np.abs(a[k] - b[k]).max()) / scale <= 1e-13, k


def test_force_sweep_source_honours_the_corrected_pde_index_flag(tmp_path):
    """SSB_FLAG_CORRECTED_PDE_INDEX (512) in k_force_mv's SOURCE, on the host: with two species and two types the reference's index
    [S_c*(type-1)+s] (E/src/model.cpp:163) reads entry [type-1][s] of the 2x2 table and the corrected index reads [s][type-1], so
    the corrected sweep on a table M must equal the parity sweep on M transposed bit for bit — and differ from the parity sweep
    on M itself."""
    from spatialpy_b200 import configs
    fm = configs.box_sdpd_rdme(nx=7, ny=7, nz=7, nt=4, output_every=4, dt=1e-5)
    fm.type = np.where(fm.x[:, 0] > np.median(fm.x[:, 0]), 2, 1).astype(fm.type.dtype)
    fm.diffusion_matrix = np.array([[0.01, 0.03], [0.02, 0.005]])
    lib = _build_force_emulator(fm, tmp_path)
    st = _force_sweep_inputs(fm, seed=4)
    corrected = _run_force_emulator(lib, fm, st, flags=8 | 512)
    parity = _run_force_emulator(lib, fm, st, flags=8)
    st_t = dict(st, dmat=np.ascontiguousarray(fm.diffusion_matrix.T))
    parity_t = _run_force_emulator(lib, fm, st_t, flags=8)
    np.testing.assert_array_equal(corrected["Q"], parity_t["Q"])
    assert np.abs(corrected["Q"] - parity["Q"]).max() > 1e-3 * np.abs(parity["Q"]).max()
    for k in ("F", "Fbp", "Frho"):                      # the flag touches nothing but the chemistry flux
        np.testing.assert_array_equal(corrected[k], parity[k])


def test_codegen_guards_bc_density_assignment_and_reaction_count():
    """(i) the static fast path is switched off by a boundary condition that ASSIGNS me->rho (boundarycondition.py:147-166, target
    'rho') — not by any mention of the word; (ii) a model with more reactions than the 64-bit dependency masks hold is refused
    at build time instead of silently never refreshing reactions 64 and above."""
    import copy
    from spatialpy_b200 import codegen
    fm = load_model("cavity2d_bc")
    assert "#define SSB_BC_TOUCHES_RHO 1" in codegen.generate_model_header(fm)
    plain = load_model("cavity2d")
    assert "#define SSB_BC_TOUCHES_RHO 0" in codegen.generate_model_header(plain)
    mention = copy.copy(plain)
    mention.bc_source = "if((me->x[0] >= system->rho0)){me->v[0]=0.0;}"          # reads a name containing "rho", assigns nothing to it
    assert "#define SSB_BC_TOUCHES_RHO 0" in codegen.generate_model_header(mention)
    big = copy.copy(load_model("birth_death"))
    big.reactions = [big.reactions[k % 2] for k in range(65)]
    with pytest.raises(codegen.BuildError):
        codegen.generate_model_header(big)


def test_batched_ensembles_refuse_models_whose_boundary_conditions_test_coordinates():
    """replicate_model translates the copies of the model; a coordinate predicate (`me->x[0] >= xmin`, boundarycondition.py:129-140)
    would select the wrong region in every copy but the first — refused, and Solver's automatic batching skips such models."""
    import copy
    from spatialpy_b200.ensemble import replicate_model
    from spatialpy_b200.solver import Solver
    fm = copy.copy(load_model("birth_death"))
    assert replicate_model(fm, 3).num_particles == 3 * fm.num_particles
    fm.bc_source = "if((me->x[0] >= 0.5)){me->nu=2.0;}"
    with pytest.raises(ValueError):
        replicate_model(fm, 3)
    sol = Solver(fm)
    sol.flat = fm.finalize()
    assert sol._auto_batch(1024) is None
    sol.flat = load_model("birth_death")
    assert sol._auto_batch(1024) is True and sol._auto_batch(8) is None
