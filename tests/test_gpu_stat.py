"""Statistical parity of the sSSA against ensembles of the UNMODIFIED reference NSM (tests/golden/*.ens.npz,
>= 1000 reference trajectories each).  BASELINE.json: per-voxel mean and variance within 3 sigma, KS p > 0.01."""
import numpy as np
import pytest
from scipy import stats

from util import load_ens, load_model

pytestmark = pytest.mark.gpu

XBINS = 8


def gpu_ensemble(name, ntraj, steps, seed0=50_000, workers=16, **kw):
    """xx of `ntraj` trajectories (seed0 + k) at the requested steps.  Small models are launch-latency bound, so `workers` engine
    handles run side by side on their own streams (ctypes releases the GIL during the engine calls); trajectory k's result does not
    depend on which handle ran it (Philox is keyed by seed + particle id)."""
    from concurrent.futures import ThreadPoolExecutor
    from spatialpy_b200.engine import Engine
    from spatialpy_b200 import codegen
    fm = load_model(name)
    codegen.build_model_unit(fm)                      # once, before the handles race for it
    workers = max(1, min(workers, ntraj))
    res = [None] * ntraj

    def work(w):
        with Engine(fm, **kw) as eng:
            for k in range(w, ntraj, workers):
                eng.reset(seed0 + k)
                done, snaps = 0, {}
                for s in steps:
                    eng.step(s - done)
                    done = s
                    snaps[s] = eng.get("xx").astype(np.int64)
                res[k] = snaps
    with ThreadPoolExecutor(max_workers=workers) as ex:
        list(ex.map(work, range(workers)))
    return {s: np.array([r[s] for r in res]) for s in steps}       # [ntraj, N, S]


def check_against_reference(name, ntraj, **kw):
    ens = load_ens(name)
    steps = [int(s) for s in ens["steps"]]
    bins = ens["bins"]
    g = gpu_ensemble(name, ntraj, steps, **kw)
    nref = int(ens["ntraj"])
    report = []
    for ti, s in enumerate(steps):
        xx = g[s]                                           # [ntraj, N, S]
        S = xx.shape[2]
        # (1) KS on the total population of every species
        tot = xx.sum(axis=1)
        for j in range(S):
            ref_tot = ens[f"t{ti}_totals"][:, j]
            if ref_tot.std() == 0 and tot[:, j].std() == 0:
                assert ref_tot[0] == tot[0, j]
                continue
            p = stats.ks_2samp(tot[:, j], ref_tot).pvalue
            report.append((f"step {s} total[{j}]", p))
            assert p > 0.01, f"{name} step {s} species {j}: KS p={p:.4f} on totals"
        # (2) KS on coarse spatial bins (Bonferroni over the family)
        binned = np.stack([np.stack([np.bincount(bins, weights=xx[k, :, j], minlength=XBINS) for j in range(S)])
                           for k in range(xx.shape[0])])    # [ntraj, S, XBINS]
        ps = []
        for j in range(S):
            for b in range(XBINS):
                a, r = binned[:, j, b], ens[f"t{ti}_binned"][:, j, b]
                if a.std() == 0 and r.std() == 0:
                    continue
                ps.append(stats.ks_2samp(a, r).pvalue)
        if ps:
            assert min(ps) > 0.01 / len(ps), f"{name} step {s}: min binned KS p={min(ps):.2e} over {len(ps)} tests"
        # (3) per-voxel mean and variance within 3 sigma (a 1% outlier allowance for the multiplicity, none beyond 4.5)
        mean_g, var_g = xx.mean(axis=0), xx.var(axis=0, ddof=1)
        mean_r, var_r = ens[f"t{ti}_vox_mean"], ens[f"t{ti}_vox_var"]
        se_mean = np.sqrt(var_g / ntraj + var_r / nref)
        ok = se_mean > 0
        z = np.abs(mean_g - mean_r)[ok] / se_mean[ok]
        zz = np.where(ok, np.abs(mean_g - mean_r) / np.where(ok, se_mean, 1.0), 0.0)
        worst = [(int(v), int(sp), float(mean_g[v, sp]), float(mean_r[v, sp]), float(zz[v, sp]))
                 for v, sp in zip(*np.unravel_index(np.argsort(-zz, axis=None)[:6], zz.shape))]
        assert (z > 3).mean() <= 0.01 and z.max() < 4.5, \
            f"{name} step {s}: mean z max {z.max():.2f}, frac>3 {(z > 3).mean():.4f}; worst (voxel, species, gpu, ref, z): {worst}"
        # variance: standard error from the fourth central moment; only where the count statistics are rich enough
        # for that estimate to mean anything (>= 100 expected molecules seen over the reference ensemble)
        m4 = ((xx - mean_g) ** 4).mean(axis=0)
        se_var = np.sqrt(np.maximum(m4 - var_g ** 2, 0) * (1.0 / ntraj + 1.0 / nref))
        okv = (se_var > 0) & (mean_r * nref >= 100)
        zv = np.abs(var_g - var_r)[okv] / se_var[okv]
        # thresholds calibrated on the null: the same statistic computed between two independent ensembles of the REFERENCE itself
        # (cdc42, 400 vs 600 trajectories, 200 random splits) has frac>3 mean 0.8 %, 99th percentile 2.8 %, and max z 99th
        # percentile 6.3 — the fourth-moment standard error is heavy-tailed for count data
        assert (zv > 3).mean() <= 0.035 and zv.max() < 7.5, f"{name} step {s}: var z max {zv.max():.2f}, frac>3 {(zv > 3).mean():.4f}"
    return report


def test_birth_death_ensemble_matches_reference():
    """BASELINE config 1 — note the reference's channel pick makes `death` unreachable (simulate_rdme.cpp:260):
    the population GROWS (100 -> ~480 in 10 s); parity mode reproduces that law."""
    check_against_reference("birth_death", 1000)


def test_cylinder_ensemble_matches_reference():
    """BASELINE config 2a — shipped cylinder mesh, A+B annihilation with sources at both ends, unequal voxel volumes."""
    check_against_reference("cylinder", 1000)


def test_cdc42_ensemble_matches_reference():
    """BASELINE config 4 (yeast polarisation, coarse lattice, short horizon): 9 type-restricted species, 13 reactions, a custom
    propensity reading a data function, voxel volumes that differ across the membrane/cytoplasm interface."""
    check_against_reference("cdc42", 1000)         # >= 1000 trajectories on both sides (north_star), reference ensemble: 1000


def test_moving_domain_rdme_matches_reference():
    """Moving SDPD domain with an advected, decaying species: the reference rebuilds its NSM every step and executes one event
    past each step's end (about 8 % of all events in this model); totals, spatial bins and per-voxel moments must agree."""
    check_against_reference("cavity2d_rdme", 1000)


def test_pure_diffusion_ensemble_matches_reference():
    check_against_reference("diffusion3d", 1000)


def test_corrected_mode_birth_death_is_poisson():
    """SSB_FLAG_CORRECTED_NSM_SELECT: textbook NSM; stationary law is Poisson(k_birth*vol/k_death = 1) per voxel,
    so total over 121 voxels ~ Poisson(121) once relaxed (t = 10 s is one death time constant: mean
    100*e^-1 + 121*(1-e^-1))."""
    from spatialpy_b200.engine import FLAG_CORRECTED_NSM_SELECT, FLAG_SKIP_STATIC_FORCES
    g = gpu_ensemble("birth_death", 300, [10], flags=FLAG_CORRECTED_NSM_SELECT | FLAG_SKIP_STATIC_FORCES)
    tot = g[10].sum(axis=(1, 2))
    expect = 100 * np.exp(-1.0) + 121 * (1 - np.exp(-1.0))
    # variance of the total: survivors Binomial(100, e^-1) + Poisson(121(1-e^-1))
    var = 100 * np.exp(-1.0) * (1 - np.exp(-1.0)) + 121 * (1 - np.exp(-1.0))
    z = abs(tot.mean() - expect) / np.sqrt(var / len(tot))
    assert z < 4, (tot.mean(), expect, z)


def test_leap_form_diffusion_parity_and_conservation():
    """SSB_FLAG_LEAP_DIFFUSION (opt-in, for very crowded voxels): binomial jump counts + multinomial destinations per window.
    Same law as the event form in the windowed scheme: checked against the reference ensemble on the pure-diffusion model
    (the four-model check — birth-death, cylinder, Cdc42, diffusion — was run once by hand: all KS p > 0.01)."""
    from spatialpy_b200.engine import FLAG_LEAP_DIFFUSION, FLAG_SKIP_STATIC_FORCES
    flags = FLAG_LEAP_DIFFUSION | FLAG_SKIP_STATIC_FORCES
    check_against_reference("diffusion3d", 800, flags=flags)
    g = gpu_ensemble("diffusion3d", 3, [10], flags=flags)
    fm = load_model("diffusion3d")
    np.testing.assert_array_equal(g[10].sum(axis=1), np.tile(fm.u0.sum(axis=0), (3, 1)))


# written after the round-1 GPU budget was spent; the construction is pinned on the CPU (tests/test_cpu_gloo.py: the copies of a
# replicated model pass the same KS test through the serial NSM restatement); the GPU run of it is pending
@pytest.mark.parametrize("name", ["birth_death", "cdc42"])
def test_batched_ensemble_has_the_reference_law(name):
    """run_ensemble_batched: 256 trajectories per engine handle as disjoint copies of the model.  Totals of every species at
    the last output step vs the reference ensemble (KS p > 0.01), and the copies of one batch are uncorrelated."""
    from spatialpy_b200.ensemble import run_ensemble_batched
    fm = load_model(name)
    ens = load_ens(name)
    ntraj = 1024
    res = run_ensemble_batched(fm, ntraj, 70_000, batch=256)
    xx = np.array([res[k]["xx_final"] for k in range(ntraj)]).astype(np.int64)        # [ntraj, N, S]
    # the fixture's last ensemble tap is the end of the run only if nt matches; both fixtures are generated that way
    assert int(ens["steps"][1]) == fm.nt
    tot = xx.sum(axis=1)
    for j in range(tot.shape[1]):
        ref = ens["t1_totals"][:, j]
        if ref.std() == 0 and tot[:, j].std() == 0:
            assert ref[0] == tot[0, j]
            continue
        p = stats.ks_2samp(tot[:, j], ref).pvalue
        assert p > 0.01, f"{name} species {j}: KS p={p:.4f}"
    j = int(np.argmax(tot.std(axis=0)))
    a = tot[:, j].reshape(4, 256)
    r_adj = np.corrcoef(a[:, :-1].ravel(), a[:, 1:].ravel())[0, 1]
    assert abs(r_adj) < 4.0 / np.sqrt(a[:, 1:].size), r_adj
