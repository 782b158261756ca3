"""GPU tier: the REFERENCE's own front door on the CUDA engine.  `import spatialpy` (the reference's Python package, staged by
oracle/build_ref.py under the untracked oracle/_ref/py/ so it travels to the GPU box), `spatialpy_b200.install()`, then the
reference's own models (test/models/birth_death.py, the diffusion model of test/integration_tests/test_solver.py:28-47) through
`Model.run(solver=spatialpy_b200.Solver)` into the reference's own `Result` — replaying the pins of
test/integration_tests/test_solver.py:127-200 (step-0 species == u0, same seed => equal results, different seed => different,
pickling of Solver / Result / Model, ensemble length) and system_tests/test_compiler.py:34-36 (birth-death compiles and runs)."""
import os
import pickle
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


@pytest.fixture(scope="module")
def spatialpy():
    import build_ref
    path = build_ref.staged_python_path()
    if path is None:
        pytest.skip("oracle/_ref/py not staged (run __graft_entry__.build() where /root/reference exists)")
    if path not in sys.path:
        sys.path.insert(0, path)
    import spatialpy as sp
    import spatialpy_b200
    spatialpy_b200.install()
    return sp


def create_diffusion_debug(sp):
    """test/integration_tests/test_solver.py:28-47, verbatim arguments."""
    model = sp.Model("diffusion_debug_test")
    A = sp.Species(name="A", diffusion_coefficient=0.01)
    model.add_species([A])
    domain = sp.Domain.create_2D_domain(xlim=[-1, 1], ylim=[-1, 1], numx=50, numy=50, type_id=1, mass=1.0, nu=1.0, fixed=True,
                                        rho0=1.0, c0=1.0, P0=1.0)
    model.add_domain(domain)
    model.add_initial_condition(sp.PlaceInitialCondition(A, 100000, [0, 0, 0]))
    model.timespan(sp.TimeSpan.linspace(t=10, num_points=11, timestep_size=0.1))
    return model


@pytest.fixture(scope="module")
def model(spatialpy):
    return create_diffusion_debug(spatialpy)


def B(sp):
    import spatialpy_b200
    return spatialpy_b200.Solver


def test_solver_io(spatialpy, model):
    """test_solver.py:127-131: the initial value in the solver output equals the input initial value — through the reference's
    own Result.get_species (its VTK reader parsing the files this engine wrote)."""
    result = model.run(solver=B(spatialpy), seed=3)
    assert type(result).__mro__[1].__module__.startswith("spatialpy") or type(result).__module__.startswith("spatialpy")
    A = result.get_species("A", 0)
    assert not (A - model.u0).any()
    assert int(result.get_species("A", -1).sum()) == 100000           # pure diffusion conserves the molecules
    assert result.get_species("A", -1)[np.argmax(model.u0)] < 100000  # ... and they have moved


def test_same_seed_explicit_solver(spatialpy, model):
    """test_solver.py:140-145 (Result.__eq__ = filecmp.dircmp of the result directories, result.py:173-183)."""
    solver = B(spatialpy)(model)
    assert solver.run(seed=1) == solver.run(seed=1)


def test_same_seed_model_run(spatialpy, model):
    """test_solver.py:147-151."""
    assert model.run(solver=B(spatialpy), seed=1) == model.run(solver=B(spatialpy), seed=1)


def test_different_seeds_and_default_seed(spatialpy, model):
    """test_solver.py:153-165."""
    solver = B(spatialpy)(model)
    assert not (solver.run(seed=1) == solver.run(seed=100))
    assert not (solver.run() == solver.run())


def test_model_solver_result_pickle(spatialpy, model):
    """test_solver.py:167-195."""
    model2 = pickle.loads(pickle.dumps(model))
    assert model.run(solver=B(spatialpy), seed=1) == model2.run(solver=B(spatialpy), seed=1)
    sol = B(spatialpy)(model)
    sol2 = pickle.loads(pickle.dumps(sol))
    r1 = sol.run(seed=1)
    assert r1 == sol2.run(seed=1)
    assert r1 == pickle.loads(pickle.dumps(r1))


def test_run_ensemble_and_timeout(spatialpy, model):
    """test_solver.py:197-200 (ensemble length), solver.py:579-586 (timeout => result.timeout, no exception)."""
    results = model.run(3, solver=B(spatialpy), seed=5)
    assert len(results) == 3
    assert not (results[0] == results[1])                              # trajectory k runs with seed + k (solver.py:558-559)
    slow = create_diffusion_debug(spatialpy)
    slow.timespan(spatialpy.TimeSpan.linspace(t=2000, num_points=11, timestep_size=0.1))
    r = slow.run(solver=B(spatialpy), seed=1, timeout=1)
    assert r.timeout is True


def test_reference_birth_death_model_compiles_and_runs(spatialpy):
    """system_tests/test_compiler.py:34-36 + README: the reference's own test/models/birth_death.py through the front door."""
    import build_ref
    sys.path.insert(0, os.path.join(build_ref.staged_python_path(), "ref_test_models"))
    import birth_death
    model = birth_death.create_birth_death()
    result = model.run(solver=B(spatialpy), seed=7)
    rabbits0 = result.get_species("Rabbits", 0)
    assert int(rabbits0.sum()) == 100                                   # ScatterInitialCondition(100)
    assert result.get_species("Rabbits", -1).sum() > rabbits0.sum()     # the reference's NSM only ever fires the birth
    ts = result.get_timespan()
    assert len(ts) == 11


def test_debug_level_prints_the_nsm_totals(spatialpy, model, capsys):
    """E/src/simulate_rdme.cpp:71-72: the debug build reports the event totals when the NSM is torn down."""
    model.run(solver=B(spatialpy), seed=2, debug_level=1)
    out = capsys.readouterr().out
    assert "NSM: total # diffusion events" in out and "NSM: total # reacton events 0" in out
