"""CPU tier: pins the oracle restatement (oracle/sdpd_oracle.py) to the full-precision taps of the UNMODIFIED
reference engine (tests/golden/*.ref.npz from oracle/_ref), the way the GPU tier pins the CUDA engine."""
import numpy as np
import pytest

from util import RTOL_STEP, RTOL_TRAJ, csr_sorted, load_model, load_ref, rel_err

import sdpd_oracle


@pytest.mark.parametrize("name", ["birth_death", "diffusion3d", "cavity2d", "tank3d", "cylinder", "cdc42", "cdc42_full", "line1d"])
def test_oracle_neighbours_match_reference(name):
    fm, ref = load_model(name), load_ref(name)
    o = sdpd_oracle.SdpdOracle(fm)
    nb = o.find_neighbors(o.x, o.x)
    np.testing.assert_array_equal(nb["ptr"], ref["s1_nbr_ptr"] if fm.static_domain else ref["s0_nbr_ptr"] if ref["s0_nbr_ptr"][-1] else nb["ptr"])
    # step-1 tap = lists built during step 0 (positions unchanged at step 0: v = 0, F = 0)
    rptr, ridx = ref["s1_nbr_ptr"], ref["s1_nbr_idx"]
    np.testing.assert_array_equal(nb["ptr"], rptr)
    gi, gd, gw, gD = csr_sorted(nb["ptr"], nb["j"].astype(np.int32), nb["dist"], nb["dWdr"], nb["Dij"])
    ri, rd, rw, rD = csr_sorted(rptr, ridx, ref["s1_nbr_dist"], ref["s1_nbr_dWdr"], ref["s1_nbr_Dij"])
    np.testing.assert_array_equal(gi, ri)
    assert rel_err(gd, rd) <= RTOL_STEP and rel_err(gw, rw) <= RTOL_STEP and rel_err(gD, rD) <= RTOL_STEP


@pytest.mark.parametrize("name", ["cavity2d", "tank3d", "diffusion3d", "cavity2d_rdme", "line1d", "cavity2d_bc"])
def test_oracle_trajectory_matches_reference(name):
    fm, ref = load_model(name), load_ref(name)
    o = sdpd_oracle.SdpdOracle(fm)
    for s in [int(v) for v in ref["steps"] if int(v) >= 1]:
        while o.step_no < s:
            o.step()
        tol = RTOL_STEP if s == 1 else RTOL_TRAJ
        for f, a in (("x", o.x), ("v", o.v), ("vt", o.vt), ("F", o.F), ("Fbp", o.Fbp), ("rho", o.rho), ("Frho", o.Frho),
                     ("bvf_phi", o.bvf), ("old_rho", o.old_rho), ("C", o.C), ("Q", o.Q)):
            err = rel_err(a, ref[f"s{s}_{f}"])
            assert err <= tol, f"{name} step {s} {f}: {err:.3e}"
        if not fm.static_domain:
            np.testing.assert_array_equal(o.nbr["ptr"], ref[f"s{s}_nbr_ptr"])


@pytest.mark.parametrize("name", ["birth_death", "diffusion3d", "cylinder", "cdc42", "cdc42_full"])
def test_oracle_ddiag_matches_reference(name):
    fm, ref = load_model(name), load_ref(name)
    o = sdpd_oracle.SdpdOracle(fm)
    o.find_neighbors(o.x, o.x)
    assert rel_err(o.ddiag(), ref["s1_Ddiag"]) <= RTOL_STEP


def test_oracle_reproduces_the_reference_abort_on_an_isolated_particle():
    """Error behaviour: a particle with no neighbour on a MOVING domain gets a Shepard-filtered density of 0/0 at step 0
    (model.cpp:194-233), and the reference exits at the NaN check of step 1 (particle.cpp:88-126) — recorded from the
    reference itself in line1d_isolated.ref.npz.  The restatement must produce the same NaN on the same particle at the same
    step (the CUDA engine reports it as SSB_ERR_NAN, tests/test_gpu_oracle.py)."""
    fm, ref = load_model("line1d_isolated"), load_ref("line1d_isolated")
    assert int(ref["exit_code"]) == 1 and str(ref["message"]) == "ERROR: nan/inf detected!!!"
    o = sdpd_oracle.SdpdOracle(fm)
    assert np.isfinite(o.rho).all()
    o.step()                                              # step 0 runs through: the check sits at the top of a step
    bad = np.nonzero(~np.isfinite(o.rho))[0]
    assert bad.tolist() == [int(ref["particle"])] and o.step_no == int(ref["step"])
    assert np.isfinite(o.x).all() and np.isfinite(o.v).all()


# ---------------------------------------------------------------------------------------------------------------
# NSM restatement vs ensembles of the unmodified reference (statistical pin)
# ---------------------------------------------------------------------------------------------------------------
def _nsm_ensemble(name, ntraj, t_end, **kw):
    import nsm_oracle
    fm = load_model(name)
    o = sdpd_oracle.SdpdOracle(fm)
    nb = o.find_neighbors(o.x, o.x)
    lib = nsm_oracle.build(fm)
    return fm, np.array([nsm_oracle.run(lib, fm, nb, 7000 + k, t_end, **kw)[0] for k in range(ntraj)]).astype(np.int64)


@pytest.mark.parametrize("name", ["birth_death", "cylinder", "diffusion3d", "cdc42"])
def test_nsm_oracle_matches_reference_ensemble(name):
    from scipy import stats
    from util import load_ens
    ens = load_ens(name)
    fm = load_model(name)
    last = int(ens["steps"][1])
    # cdc42 (9 species, 13 reactions, a data function, type-dependent voxel volumes) costs 8e5 events per trajectory
    fm2, xx = _nsm_ensemble(name, 200 if name == "cdc42" else 600, last * fm.dt)
    for j in range(xx.shape[2]):
        a, r = xx[:, :, j].sum(axis=1), ens["t1_totals"][:, j]
        if a.std() == 0 and r.std() == 0:
            assert a[0] == r[0]
            continue
        p = stats.ks_2samp(a, r).pvalue
        assert p > 0.01, f"{name} species {j}: KS p={p:.4f}"
    mean_g, var_g = xx.mean(axis=0), xx.var(axis=0, ddof=1)
    se = np.sqrt(var_g / xx.shape[0] + ens["t1_vox_var"] / int(ens["ntraj"]))
    ok = se > 0
    z = np.abs(mean_g - ens["t1_vox_mean"])[ok] / se[ok]
    assert (z > 3).mean() <= 0.01 and z.max() < 4.5, (z.max(), (z > 3).mean())


def test_data_function_reaches_the_propensities():
    """test/integration_tests/test_model.py:81-104 (propensity = data function 10000*x): no births where x = 0, Poisson(10^4)
    births per unit time where x = 1 — the reference's own run is in datafn.ref.npz; the NSM restatement agrees."""
    import nsm_oracle
    fm, ref = load_model("datafn"), load_ref("datafn")
    rx = ref["s1_xx"].ravel().astype(np.int64)
    assert rx[0] == 0 and rx[1] == 0 and rx[2] > 0 and rx[3] > 0          # the reference test's own assertions
    o = sdpd_oracle.SdpdOracle(fm)
    nb = o.find_neighbors(o.x, o.x)
    xx, n_rx, n_df = nsm_oracle.run(nsm_oracle.build(fm), fm, nb, 5, fm.nt * fm.dt)
    xx = xx.ravel().astype(np.int64)
    assert xx[0] == 0 and xx[1] == 0 and n_df == 0 and n_rx == xx.sum()
    assert abs(xx[2] - 10000) < 500 and abs(xx[3] - 10000) < 500 and abs(rx[2] - 10000) < 500     # 5 sigma of Poisson(10^4)


def test_single_letter_species_names_compile_for_the_device():
    """test/integration_tests/test_model.py:60-78: all 51 non-reserved single-letter species names.  The propensity ABI's own
    identifiers (x, t, vol, sd ...) must not collide with species named like them: the model unit builds for sm_100a and the
    fixture's step-0 state is what the reference produced."""
    from spatialpy_b200 import codegen
    fm, ref = load_model("letters"), load_ref("letters")
    assert fm.num_species == 51 and {"x", "e", "s", "d", "N", "S"} <= set(fm.species_names) and "t" not in fm.species_names
    assert ref["s1_xx"].shape == (fm.num_particles, 51) and not ref["s1_xx"].any()
    import os
    assert os.path.exists(codegen.build_model_unit(fm))
