"""GPU tier: the CUDA engine against the CPU oracle restatement (oracle/sdpd_oracle.py, oracle/nsm_oracle.cpp — both
pinned to the unmodified reference by tests/test_cpu_oracle.py) on seeded inputs that have no golden fixture, and
size-independent properties at the BASELINE sizes."""
import numpy as np
import pytest
from scipy import stats

from util import RTOL_STEP, RTOL_TRAJ, csr_sorted, rel_err

import sdpd_oracle

pytestmark = pytest.mark.gpu


def _jitter(fm, seed, amp):
    rng = np.random.default_rng(seed)
    fm.x = fm.x + rng.uniform(-amp, amp, size=fm.x.shape) * (np.arange(3) < fm.dimension)
    return fm


@pytest.mark.parametrize("seed", [11, 12])
def test_moving_tank_matches_oracle(seed):
    """Ragged (jittered) 3-D tank: neighbour sets bit-exact, every field within tolerance over 22 steps (Shepard at 0 and 20)."""
    from spatialpy_b200 import configs
    from spatialpy_b200.engine import Engine
    fm = _jitter(configs.tank_sdpd(n=11, nt=30, output_every=30, dt=2e-5), seed, 0.012)
    o = sdpd_oracle.SdpdOracle(fm)
    with Engine(fm) as eng:
        eng.reset(1)
        for s in (1, 2, 21, 22):
            while o.step_no < s:
                o.step()
            eng.step(s - (s - 1 if s in (2, 22) else 0 if s == 1 else 2))
            ptr, idx, dist, dWdr, Dij = eng.neighbors()
            np.testing.assert_array_equal(ptr, o.nbr["ptr"])
            gi, gd = csr_sorted(ptr, idx, dist)
            ri, rd = csr_sorted(o.nbr["ptr"], o.nbr["j"].astype(np.int32), o.nbr["dist"])
            np.testing.assert_array_equal(gi, ri)
            tol = RTOL_STEP if s == 1 else RTOL_TRAJ
            for f, a in (("x", o.x), ("v", o.v), ("rho", o.rho), ("F", o.F), ("Fbp", o.Fbp), ("Frho", o.Frho),
                         ("bvf_phi", o.bvf), ("C", o.C)):
                err = rel_err(eng.get(f), a)
                assert err <= tol, f"seed {seed} step {s} {f}: {err:.3e}"


def test_edge_cases_isolated_particle_and_1d():
    """A particle with no neighbour inside h (empty list) and a 1-D domain (alpha = `5/4*h` == h, particle.cpp:174)."""
    from spatialpy_b200 import FlatModel
    from spatialpy_b200.engine import Engine
    N = 12
    x = np.zeros((N, 3))
    x[:, 0] = np.arange(N) * 0.1
    x[-1, 0] = 5.0                                     # isolated: no neighbour
    fm = FlatModel(name="line1d", x=x, type=np.ones(N, np.int32), nu=np.ones(N), mass=np.ones(N), c=np.zeros(N),
                   rho=np.ones(N), solid=np.ones(N, np.int32), static_domain=True, dt=0.01, nt=2,
                   output_steps=np.array([0, 1, 2], np.uint32), h=0.25, dimension=1, xlim=(0, 5), ylim=(0, 0), zlim=(0, 0)).finalize()
    o = sdpd_oracle.SdpdOracle(fm)
    nb = o.find_neighbors(o.x, o.x)
    with Engine(fm, flags=0) as eng:
        eng.reset(1)
        eng.step(1)
        ptr, idx, dist, dWdr, Dij = eng.neighbors()
        np.testing.assert_array_equal(ptr, nb["ptr"])
        assert ptr[-1] - ptr[-2] == 0
        gi, gw = csr_sorted(ptr, idx, dWdr)
        ri, rw = csr_sorted(nb["ptr"], nb["j"].astype(np.int32), nb["dWdr"])
        np.testing.assert_array_equal(gi, ri)
        assert rel_err(gw, rw) <= RTOL_STEP


def test_small_cylinder_rdme_matches_nsm_oracle():
    """Synthetic config 2b geometry at oracle size: sSSA ensemble vs the serial NSM restatement, KS p > 0.01."""
    import nsm_oracle
    from spatialpy_b200 import configs
    from spatialpy_b200.engine import Engine
    fm = configs.cylinder_rdme(delta=0.25, nt=40, output_every=40, dt=0.05, enable_pde=False)
    o = sdpd_oracle.SdpdOracle(fm)
    nb = o.find_neighbors(o.x, o.x)
    lib = nsm_oracle.build(fm)
    ntraj = 400
    ref = np.array([nsm_oracle.run(lib, fm, nb, 9000 + k, fm.nt * fm.dt)[0] for k in range(ntraj)]).astype(np.int64)
    got = []
    with Engine(fm) as eng:
        for k in range(ntraj):
            eng.reset(100 + k)
            eng.step(fm.nt)
            got.append(eng.get("xx").astype(np.int64))
    got = np.array(got)
    for j in range(2):
        p = stats.ks_2samp(got[:, :, j].sum(axis=1), ref[:, :, j].sum(axis=1)).pvalue
        assert p > 0.01, f"species {j}: KS p={p:.4f}"
    mg, mr = got.mean(axis=0), ref.mean(axis=0)
    se = np.sqrt(got.var(axis=0, ddof=1) / ntraj + ref.var(axis=0, ddof=1) / ntraj)
    ok = se > 0
    z = np.abs(mg - mr)[ok] / se[ok]
    assert (z > 3).mean() <= 0.01 and z.max() < 4.5, (z.max(), (z > 3).mean())


def test_full_size_static_cylinder_properties():
    """BASELINE configs[1] at full size (~1.0 M voxels): molecule balance = u0 + creations - 2*annihilations is checked through
    the counters, neighbour relation is symmetric in count, C stays finite, same seed => identical state."""
    from spatialpy_b200 import configs
    from spatialpy_b200.engine import Engine
    fm = configs.cylinder_rdme(nt=20, output_every=20)
    with Engine(fm) as eng:
        outs = []
        for rep in range(2):
            eng.reset(77)
            eng.step(10)
            outs.append((eng.get("xx").copy(), eng.get("C").copy()))
        np.testing.assert_array_equal(outs[0][0], outs[1][0])
        np.testing.assert_array_equal(outs[0][1], outs[1][1])
        xx = outs[0][0].astype(np.int64)
        c = eng.counters()
        # A and B are created one at a time and annihilated in pairs: #A - #B changes only by creations, and
        # reactions >= |#A - #B| ; diffusion conserves both
        assert c["reactions"] >= abs(int(xx[:, 0].sum()) - int(xx[:, 1].sum()))
        assert np.isfinite(outs[0][1]).all()
        cnt = eng.get("nbr_count")
        cap, total = eng.nbr_stats()
        assert total == int(cnt.sum()) and total % 2 == 0          # j in N(i) <=> i in N(j) on a static domain
        # species restriction: A never enters Edge2 voxels, B never enters Edge1 voxels (diffusion matrix zeros)
        assert xx[fm.type == 1, 0].sum() == 0 and xx[fm.type == 2, 1].sum() == 0


def test_full_size_moving_tank_properties():
    """BASELINE configs[2] stand-in at full size (~1.0 M moving particles): same seed => identical state, walls never move,
    the fluid stays finite and inside the tank, molecule bookkeeping balances, and the Verlet-skin candidate lists give the
    same neighbour counts as exact per-step lists (SSB_SKIN=0)."""
    import os
    from spatialpy_b200 import configs
    from spatialpy_b200.engine import Engine
    fm = configs.tank_sdpd(n=120, nt=20, output_every=20)
    fm.parameters = {"P0": 0.0, "P1": 0.0}            # reactions off: diffusion must conserve the molecule count exactly
    outs = {}
    for skin in ("auto", "0"):
        if skin == "0":
            os.environ["SSB_SKIN"] = "0"
        try:
            with Engine(fm) as eng:
                eng.reset(5)
                eng.step(6)
                outs[skin] = dict(x=eng.get("x"), v=eng.get("v"), rho=eng.get("rho"), xx=eng.get("xx"), stats=eng.skin_stats(),
                                  nbr=eng.neighbors()[0], counters=eng.counters())
                if skin == "auto":
                    eng.reset(5)
                    eng.step(6)
                    np.testing.assert_array_equal(eng.get("x"), outs[skin]["x"])       # bit-reproducible
                    np.testing.assert_array_equal(eng.get("xx"), outs[skin]["xx"])
        finally:
            os.environ.pop("SSB_SKIN", None)
    a, b = outs["auto"], outs["0"]
    walls = fm.solid == 1
    np.testing.assert_array_equal(a["x"][walls], fm.x[walls])
    assert np.isfinite(a["x"]).all() and np.isfinite(a["v"]).all() and np.isfinite(a["rho"]).all()
    assert a["x"].min() >= -1e-9 and a["x"].max() <= 1.0 + 1e-9
    assert int(a["xx"].sum()) == int(fm.u0.sum()) and a["counters"]["reactions"] == 0 and a["counters"]["diffusions"] > 0
    assert a["stats"]["skin"] > 0 and b["stats"]["skin"] == 0 and a["stats"]["rebuilds"] == 1     # one list build served all 6 steps
    np.testing.assert_array_equal(a["nbr"], b["nbr"])                               # identical neighbour counts per particle
    assert rel_err(a["x"], b["x"]) <= RTOL_TRAJ and rel_err(a["rho"], b["rho"]) <= RTOL_TRAJ


def test_corrected_stoichiometry_flag_matches_the_restatement_with_the_right_index():
    """SSB_FLAG_CORRECTED_STOICH: the deterministic reaction term reads N[s][rxn] (species x reactions, row-major) instead of the
    reference's transposed / out-of-bounds index (E/src/model.cpp:186-187).  On the Cdc42 model (9 species, 13 reactions: the
    two indexings differ) the engine with the flag must follow the numpy restatement evaluated with the correct index — and
    must differ from the parity mode, or the flag does nothing."""
    import sdpd_oracle
    from spatialpy_b200.engine import Engine, FLAG_CORRECTED_STOICH, FLAG_SKIP_STATIC_FORCES
    from util import load_model, rel_err
    fm = load_model("cdc42")
    o = sdpd_oracle.SdpdOracle(fm)
    o.corrected_stoich = True
    out = {}
    for name, flags in (("corrected", FLAG_SKIP_STATIC_FORCES | FLAG_CORRECTED_STOICH), ("parity", FLAG_SKIP_STATIC_FORCES)):
        with Engine(fm, flags=flags) as eng:
            eng.reset(1000)
            eng.step(2)
            out[name] = (eng.get("C"), eng.get("Q"))
    for _ in range(2):
        o.step()
    assert rel_err(out["corrected"][1], o.Q) <= 1e-12 and rel_err(out["corrected"][0], o.C) <= 1e-12
    assert rel_err(out["parity"][1], o.Q) > 1e-6


def test_corrected_pde_index_flag_reads_the_species_major_entry():
    """SSB_FLAG_CORRECTED_PDE_INDEX: the PDE flux takes D[s, type] from the species-major table (the entry simulate_rdme.cpp:146
    reads) instead of the reference's [S_c*(type-1)+s] (E/src/model.cpp:163).  Two species x two types with type-dependent
    coefficients make the two indexings differ; all three sites are covered — the optimised force sweep, the literal sweep and
    (on Cdc42, static) the streaming sweep — each against the numpy restatement evaluated with the same index."""
    from spatialpy_b200 import configs
    from spatialpy_b200.engine import (Engine, FLAG_CORRECTED_PDE_INDEX, FLAG_CORRECTED_STOICH, FLAG_LITERAL_KERNELS,
                                       FLAG_SKIP_STATIC_FORCES)
    from util import load_model
    fm = _jitter(configs.box_sdpd_rdme(nx=7, ny=7, nz=7, nt=4, output_every=4, dt=1e-5), 5, 0.01)
    fm.type = np.where(fm.x[:, 0] > np.median(fm.x[:, 0]), 2, 1).astype(fm.type.dtype)
    fm.diffusion_matrix = np.array([[0.01, 0.03], [0.02, 0.005]])
    fm.u0 = np.random.default_rng(6).integers(0, 40, size=fm.u0.shape).astype(fm.u0.dtype)     # concentration gradients
    ref = {}
    for corrected in (True, False):
        o = sdpd_oracle.SdpdOracle(fm)
        o.corrected_pde_index = corrected
        o.step()
        ref[corrected] = (o.Q.copy(), o.C.copy())
    assert rel_err(ref[False][0], ref[True][0]) > 1e-3, "the fixture does not separate the two indexings"
    for flags, want in ((FLAG_CORRECTED_PDE_INDEX, True), (FLAG_CORRECTED_PDE_INDEX | FLAG_LITERAL_KERNELS, True), (0, False)):
        with Engine(fm, flags=flags) as eng:
            eng.reset(1)
            eng.step(1)
            for f, a in (("Q", ref[want][0]), ("C", ref[want][1])):
                err = rel_err(eng.get(f), a)
                assert err <= RTOL_STEP, f"flags {flags} {f}: {err:.3e}"
    # static streaming sweep (k_static_step): Cdc42, 9 species x 3 types
    fm = load_model("cdc42")
    o = sdpd_oracle.SdpdOracle(fm)
    o.corrected_stoich = True
    o.corrected_pde_index = True
    for _ in range(2):
        o.step()
    out = {}
    for name, extra in (("both", FLAG_CORRECTED_PDE_INDEX), ("stoich", 0)):
        with Engine(fm, flags=FLAG_SKIP_STATIC_FORCES | FLAG_CORRECTED_STOICH | extra) as eng:
            eng.reset(1000)
            eng.step(2)
            out[name] = (eng.get("C"), eng.get("Q"))
    assert rel_err(out["both"][1], o.Q) <= 1e-12 and rel_err(out["both"][0], o.C) <= 1e-12
    assert rel_err(out["stoich"][1], o.Q) > 1e-6
