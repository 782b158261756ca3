"""GPU tier: the drop-in Solver surface and the result-directory file contract."""
import filecmp
import gzip
import os

import numpy as np
import pytest

from util import GOLDEN, load_model

pytestmark = pytest.mark.gpu


def _ref_file(name):
    with gzip.open(os.path.join(GOLDEN, "vtk_diffusion3d", name + ".gz"), "rb") as f:
        return f.read().decode()


def _sections(text):
    """split a VTK file into {array header line: body text}; everything before POINT_DATA under key 'geometry'."""
    head, _, rest = text.partition("POINT_DATA")
    out = {"geometry": head}
    lines = ("POINT_DATA" + rest).split("\n")
    out["point_data_header"] = "\n".join(lines[:2])
    cur, buf = None, []
    for ln in lines[2:]:
        parts = ln.split()
        if len(parts) == 4 and parts[3] in ("int", "double") and parts[1].isdigit():
            if cur:
                out[cur] = "\n".join(buf)
            cur, buf = ln, []
        else:
            buf.append(ln)
    if cur:
        out[cur] = "\n".join(buf)
    return out


@pytest.fixture(scope="module")
def diffusion_run():
    from spatialpy_b200 import Solver
    sol = Solver(load_model("diffusion3d"))
    res = sol.run(number_of_trajectories=1, seed=1000)
    return sol, res


def test_result_directory_has_the_reference_file_set(diffusion_run):
    _, res = diffusion_run
    listing = open(os.path.join(GOLDEN, "vtk_diffusion3d", "listing.txt")).read().split()
    assert sorted(os.listdir(res.result_dir)) == sorted(listing)       # incl. the file->step map of simulate_threads.cpp:231-247
    assert res.success and not res.timeout and len(res) == 1


def test_vtk_bytes_match_reference_writer(diffusion_run):
    """Every deterministic byte equals the reference writer's output (E/src/output.cpp:104-229): geometry, header lines,
    id/type/v/rho/mass/bvf_phi/nu and the deterministic C[] arrays; D[] arrays are stochastic (different RNG) except at t=0."""
    _, res = diffusion_run
    assert open(os.path.join(res.result_dir, "output0_boundingBox.vtk")).read() == _ref_file("output0_boundingBox.vtk")
    for k in (0, 1, 10):
        mine = _sections(open(os.path.join(res.result_dir, f"output{k}.vtk")).read())
        ref = _sections(_ref_file(f"output{k}.vtk"))
        assert list(mine) == list(ref), f"output{k}: array order / headers differ"
        for key in ref:
            if key.startswith("D[") and k > 0:
                continue
            assert mine[key] == ref[key], f"output{k}.vtk section {key!r} differs"
    assert open(os.path.join(res.result_dir, "output0.vtk")).read() == _ref_file("output0.vtk")


def test_vtk_bytes_identical_with_threaded_formatting(diffusion_run, monkeypatch):
    """Large snapshots are formatted by several host threads; the bytes must not depend on the thread count."""
    from spatialpy_b200 import Solver
    _, res1 = diffusion_run
    monkeypatch.setenv("SSB_VTK_THREADS", "5")
    res5 = Solver(load_model("diffusion3d")).run(number_of_trajectories=1, seed=1000)
    for k in (0, 1, 10):
        a = open(os.path.join(res1.result_dir, f"output{k}.vtk"), "rb").read()
        b = open(os.path.join(res5.result_dir, f"output{k}.vtk"), "rb").read()
        assert a == b, f"output{k}.vtk differs between 1 and 5 formatting threads"


def test_reader_roundtrip_and_step0_equals_u0(diffusion_run):
    sol, res = diffusion_run
    pts, data = res.read_step(0)
    fm = sol.flat
    assert pts.dtype == np.float32 and data["id"].dtype == np.int64
    np.testing.assert_array_equal(data["D[A]"], fm.u0[:, 0])                  # test_solver.py:127-131
    np.testing.assert_array_equal(data["D[B]"], fm.u0[:, 1])
    np.testing.assert_allclose(pts, fm.x.astype(np.float32))
    assert data["__nfields_header__"] == 7 + 2                               # S_d not counted in output0 (output.cpp:151-154)
    assert res.read_step(1)[1]["__nfields_header__"] == 7 + 2 + 2
    assert res.get_species("A").shape == (11, fm.num_particles)
    np.testing.assert_array_equal(res.get_species("A").sum(axis=1), np.full(11, fm.u0[:, 0].sum()))


def test_same_seed_identical_dirs_different_seed_differs():
    """test/integration_tests/test_solver.py:140-165."""
    from spatialpy_b200 import Solver
    sol = Solver(load_model("birth_death"))
    a = sol.run(seed=5)
    b = sol.run(seed=5)
    c = sol.run(seed=6)
    cmp_ab = filecmp.dircmp(a.result_dir, b.result_dir)
    assert not (cmp_ab.left_only or cmp_ab.right_only)
    match, mismatch, errors = filecmp.cmpfiles(a.result_dir, b.result_dir, cmp_ab.common_files, shallow=False)
    assert not mismatch and not errors
    _, mismatch_c, _ = filecmp.cmpfiles(a.result_dir, c.result_dir, cmp_ab.common_files, shallow=False)
    assert mismatch_c


def test_ensemble_length_and_seed_mapping():
    """number_of_trajectories -> list-like Result (test_solver.py:197-200); trajectory k == a single run with seed+k."""
    from spatialpy_b200 import Solver
    sol = Solver(load_model("birth_death"))
    ens = sol.run(number_of_trajectories=3, seed=40)
    assert len(ens) == 3 and all(r.success for r in ens.listOfResultObjects)
    single = sol.run(seed=42)
    np.testing.assert_array_equal(ens[2].read_step(10)[1]["D[Rabbits]"], single.read_step(10)[1]["D[Rabbits]"])


def test_timeout_sets_flag():
    from spatialpy_b200 import Solver, configs
    fm = configs.cylinder_rdme(delta=0.08, nt=200000, output_every=100000)
    res = Solver(fm).run(timeout=1)
    assert res.timeout and not res.success


def test_nan_raises_simulation_error():
    """check_particle_nan -> exit(1) -> SimulationError (particle.cpp:88-126, solver.py:595-597)."""
    from spatialpy_b200 import SimulationError, Solver
    fm = load_model("cavity2d")
    fm.x = fm.x.copy()
    fm.x[7, 0] = np.nan
    with pytest.raises(SimulationError, match="return code = 1"):
        Solver(fm).run(seed=1)


def test_binary_side_store_matches_the_vtk_files(diffusion_run):
    """`binary_store=True` keeps outputN.ssb next to outputN.vtk; read_step then serves the same keys / shapes / dtypes from the
    raw arrays.  Integers are identical, fp64 fields agree to the six decimals the text keeps (output.cpp:170-230 prints %lf)."""
    from spatialpy_b200 import Solver
    from spatialpy_b200.vtk import read_ssb, read_vtk
    _, plain = diffusion_run
    res = Solver(load_model("diffusion3d")).run(number_of_trajectories=1, seed=1000, binary_store=True)
    names = sorted(os.listdir(res.result_dir))
    assert sorted(n for n in names if n.endswith(".vtk")) == sorted(os.listdir(plain.result_dir))
    assert len([n for n in names if n.endswith(".ssb")]) == len(names) // 2      # one per outputN.vtk (bounding box has none)
    for k in (0, 1, 10):
        pv, av = read_vtk(os.path.join(res.result_dir, f"output{k}.vtk"))
        pb, ab = read_ssb(os.path.join(res.result_dir, f"output{k}.ssb"))
        assert list(av) == list(ab)
        assert pb.dtype == pv.dtype == np.float32 and np.allclose(pb, pv, rtol=2e-7, atol=0)
        for key in av:
            if key.startswith("__"):
                assert av[key] == ab[key]
            elif av[key].dtype == np.int64:
                assert ab[key].dtype == np.int64 and np.array_equal(av[key], ab[key]), key
            else:
                assert ab[key].shape == av[key].shape and np.max(np.abs(av[key] - ab[key])) <= 5.0000001e-7, key
        assert filecmp.cmp(os.path.join(res.result_dir, f"output{k}.vtk"), os.path.join(plain.result_dir, f"output{k}.vtk"),
                           shallow=False) or k > 0        # same seed: output0 is byte-identical; later D[] too (checked next)
    assert np.array_equal(res.read_step(10)[1]["D[A]"], plain.read_step(10)[1]["D[A]"])     # same seed, same trajectory
    assert res.read_step(10)[1]["C[A]"].dtype == np.float64


def test_binary_only_run(tmp_path):
    from spatialpy_b200 import Solver
    from spatialpy_b200.solver import SimulationError
    sol = Solver(load_model("diffusion3d"))
    res = sol.run(number_of_trajectories=2, seed=7, binary_store=True, vtk=False)
    for r in res.listOfResultObjects:
        assert all(n.endswith(".ssb") for n in os.listdir(r.result_dir)) and len(os.listdir(r.result_dir)) == 11
        pts, arr = r.read_step(5)
        assert pts.shape[1] == 3 and arr["D[A]"].sum() > 0
        assert r.get_species("A").shape[0] == 11
    with pytest.raises(SimulationError):
        sol.run(vtk=False)


def test_corrected_output_steps_flag_writes_one_file_per_time_point(tmp_path):
    """SSB_FLAG_CORRECTED_OUTPUT_STEPS: file k holds step output_steps[k].  The reference's gate (simulate_threads.cpp:231-247,
    283-288) writes steps 0, 1, f, 2f, ... and the final state — one file more, and file k >= 2 holds step (k-1) f.  Same seed,
    both gates: the corrected file k >= 1 must be byte-identical to the reference gate's file k+1, and file 0 to file 0."""
    from spatialpy_b200.engine import Engine, FLAG_CORRECTED_OUTPUT_STEPS, FLAG_SKIP_STATIC_FORCES
    from spatialpy_b200.slab import output_schedule
    fm = load_model("diffusion3d")
    fm.output_steps = np.array([0, 5, 10], dtype=fm.output_steps.dtype)      # nt = 10: reference gate -> steps 0, 1, 5, 10
    assert [s for _, s in output_schedule(fm.nt, fm.output_steps)] == [0, 1, 5, 10]
    dirs = {}
    for name, extra in (("reference", 0), ("corrected", FLAG_CORRECTED_OUTPUT_STEPS)):
        d = tmp_path / name
        d.mkdir()
        with Engine(fm, flags=FLAG_SKIP_STATIC_FORCES | extra) as eng:
            eng.run(1000, [str(d)])
        dirs[name] = d
    vtk = lambda d: sorted((f for f in os.listdir(d) if f.startswith("output") and "bounding" not in f), key=lambda f: int(f[6:-4]))
    ref_files, cor_files = vtk(dirs["reference"]), vtk(dirs["corrected"])
    sched = output_schedule(fm.nt, fm.output_steps, corrected=True)
    assert [s for _, s in sched] == [int(v) for v in fm.output_steps if v <= fm.nt]
    assert len(cor_files) == len(sched) == len(ref_files) - 1
    read = lambda d, f: open(os.path.join(d, f), "rb").read()
    assert read(dirs["corrected"], cor_files[0]) == read(dirs["reference"], ref_files[0])
    for k in range(1, len(cor_files)):
        assert read(dirs["corrected"], cor_files[k]) == read(dirs["reference"], ref_files[k + 1]), k
