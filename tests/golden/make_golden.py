"""Generate the committed parity fixtures from the UNMODIFIED reference engine (dev container only).

    python tests/golden/make_golden.py [names...]

For every model in tests/golden/models.py this
  1. flattens the spatialpy.Model with FlatModel.from_spatialpy  -> tests/golden/<name>.model.npz
     (the exact inputs both engines consume; same Python process => same type indices),
  2. builds oracle/_ref/<name>/parity_dump/ssa_sdpd.exe (reference sources + oracle/ref_dump_output.cpp),
  3. runs it with `-t 1 -s <seed>` and stores selected full-precision taps -> tests/golden/<name>.ref.npz,
  4. for stochastic models runs an ensemble and stores per-trajectory observables -> tests/golden/<name>.ens.npz.
The GPU box never runs this script (no /root/reference there); it only reads the .npz files.
"""
import os
import shutil
import sys
import tempfile

if os.environ.get("PYTHONHASHSEED") != "0":     # type indices come from iterating a Python set (domain.py:141-146):
    os.environ["PYTHONHASHSEED"] = "0"           # pin the hash seed so every fixture of a model agrees on them
    os.execv(sys.executable, [sys.executable] + sys.argv)
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import build_ref  # noqa: E402
import ref_dump  # noqa: E402

build_ref.add_reference_to_path()
import models  # noqa: E402
from spatialpy_b200 import FlatModel  # noqa: E402

SEED = 1000
TAP_FIELDS = ("x", "v", "vt", "F", "Fbp", "rho", "old_rho", "Frho", "bvf_phi", "C", "Q", "xx")
NBR_FIELDS = ("nbr_ptr", "nbr_idx", "nbr_dist", "nbr_dWdr", "nbr_Dij")
RDME_FIELDS = ("srrate", "sdrate", "Ddiag", "rrate")

# which dump steps to keep per model (dump k = state at the top of step k = after k completed steps)
TAPS = {
    "birth_death": [0, 1],
    "diffusion3d": [0, 1, 2, 10],
    "cavity2d": [0, 1, 2, 3, 20, 21, 22, 45],
    "tank3d": [0, 1, 2, 21, 25],
    "cylinder": [0, 1],
    "cdc42": [0, 1, 2],
    "cavity2d_rdme": [0, 1, 22, 45],
    "cdc42_full": [0, 1, 2],
    "line1d": [0, 1, 2, 3, 4],
    "letters": [0, 1],
    "datafn": [0, 1],
    "cavity2d_bc": [0, 1, 2, 21, 22],
    "mid_cylinder": [0, 1, 25],
    "mid_tank": [0, 1, 2, 21, 25],
}
# mid-size fixtures (~15 k particles): the models are regenerated from spatialpy_b200.configs by the tests (deterministic builders,
# a checksum of x is stored) and the neighbour lists are stored as per-particle DIGESTS (count, sum and sum of squares of the
# neighbour ids, sums of dist / dWdr / D_i_j in id order) instead of 20 MB of CSR per tap
MID = ("mid_cylinder", "mid_tank")
EXPECTED_ABORT = {"line1d_isolated": "ERROR: nan/inf detected!!!"}
ENSEMBLES = {"birth_death": 1500, "cylinder": 1000, "diffusion3d": 1000, "cdc42": 1000, "cavity2d_rdme": 1000}
XBINS = 8


def coarse_bins(fm):
    """voxel -> coarse bin along x (observables for the KS tests)."""
    x = fm.x[:, 0]
    lo, hi = x.min(), x.max()
    b = np.minimum(((x - lo) / max(hi - lo, 1e-300) * XBINS).astype(int), XBINS - 1)
    return b


def nbr_digest(d):
    """Per-particle digest of a CSR neighbour tap: equal digests <=> equal neighbour SETS (count, sum of ids, sum of squared ids
    in uint64 arithmetic) and, summed in neighbour-id order, the frozen pair quantities."""
    ptr, idx = d["nbr_ptr"], d["nbr_idx"].astype(np.int64)
    n = len(ptr) - 1
    row = np.repeat(np.arange(n), np.diff(ptr))
    order = np.lexsort((idx, row))
    u = idx.astype(np.uint64)
    dig = {"nbr_count": np.diff(ptr).astype(np.int32),
           "nbr_idsum": np.bincount(row, weights=None, minlength=n) * 0}
    dig["nbr_idsum"] = np.zeros(n, np.uint64)
    dig["nbr_idsq"] = np.zeros(n, np.uint64)
    np.add.at(dig["nbr_idsum"], row, u)
    np.add.at(dig["nbr_idsq"], row, u * u)
    for f in ("dist", "dWdr", "Dij"):
        acc = np.zeros(n)
        np.add.at(acc, row[order], d[f"nbr_{f}"][order])
        dig[f"nbr_{f}_sum"] = acc
    return dig


def run_one(exe, seed):
    d = tempfile.mkdtemp(prefix="ssb_golden_")
    try:
        build_ref.run_exe(exe, d, seed, threads=1)
        steps = sorted(int(f[5:11]) for f in os.listdir(d) if f.startswith("dump_"))
        dumps = {s: ref_dump.read_dump(os.path.join(d, f"dump_{s:06d}.bin")) for s in steps}
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return dumps


def make(name):
    model = models.BUILDERS[name]()
    # compile_prep() re-applies the (random) scatter initial conditions every time it is called
    # (model.py:178-185), and both flattenings call it: re-seed so the two engines get the same u0.
    if isinstance(model, FlatModel):          # synthetic edge-case models: the reference's own template filled from the arrays
        fm = model
        fm.save(os.path.join(HERE, f"{name}.model.npz"))
        exe = build_ref.build_flat(fm, name, variant="parity", dump=True)
    else:
        np.random.seed(12345)
        fm = FlatModel.from_spatialpy(model)
        fm.save(os.path.join(HERE, f"{name}.model.npz"))
        np.random.seed(12345)
        exe = build_ref.build_model(model, name, variant="parity", dump=True, h=fm.h)
    if name in EXPECTED_ABORT:                # error-behaviour fixture: the reference must exit(1) with this message
        try:
            run_one(exe, SEED)
        except RuntimeError as err:
            text = str(err)
            assert EXPECTED_ABORT[name] in text, text[:400]
            ident = int(text.split("\nid=")[1].split("\n")[0])
            step = int(text.split("sys->current_step=")[1].split("\n")[0])
            np.savez_compressed(os.path.join(HERE, f"{name}.ref.npz"), exit_code=np.array(1), message=np.array(EXPECTED_ABORT[name]),
                                particle=np.array(ident), step=np.array(step))
            print(f"{name}: reference aborted as expected: '{EXPECTED_ABORT[name]}' particle {ident} at step {step}")
            return
        raise SystemExit(f"{name}: the reference was expected to abort but ran through")
    dumps = run_one(exe, SEED)
    out = {"steps": np.array(TAPS[name]), "seed": np.array(SEED)}
    if name in MID:
        os.remove(os.path.join(HERE, f"{name}.model.npz"))
        out["x_checksum"] = np.array(float(np.sum(fm.x * np.arange(1, fm.x.size + 1).reshape(fm.x.shape))))
    for s in TAPS[name]:
        d = dumps[s]
        for f in TAP_FIELDS:
            if name in MID and (f == "xx" or (s > 1 and f in ("vt", "F", "Fbp", "Frho", "old_rho", "Q"))):
                continue
            out[f"s{s}_{f}"] = d[f]
        if name in MID:
            if s == 1 or (s > 1 and not fm.static_domain):
                out.update({f"s{s}_{k}": v for k, v in nbr_digest(d).items()})
        elif s <= 1 or not fm.static_domain:
            for f in NBR_FIELDS:
                out[f"s{s}_{f}"] = d[f]
        if d["initialized"]:
            for f in RDME_FIELDS:
                out[f"s{s}_{f}"] = d[f]
        out[f"s{s}_events"] = np.array([d["total_reactions"], d["total_diffusion"]])
    np.savez_compressed(os.path.join(HERE, f"{name}.ref.npz"), **out)
    print(f"{name}: N={fm.num_particles} taps={TAPS[name]} final events={dumps[max(dumps)]['total_reactions']}/{dumps[max(dumps)]['total_diffusion']}")
    if name in ENSEMBLES:
        ntraj = ENSEMBLES[name]
        bins = coarse_bins(fm)
        last = max(dumps)
        mid = sorted(dumps)[len(dumps) // 2]

        def obs(k):
            dd = run_one(exe, SEED + k)
            res = []
            for s in (mid, last):
                xx = dd[s]["xx"].astype(np.int64)                       # [N, S]
                binned = np.stack([np.bincount(bins, weights=xx[:, j], minlength=XBINS) for j in range(xx.shape[1])])
                res.append((xx.sum(axis=0), binned, xx))
            return res, (dd[last]["total_reactions"], dd[last]["total_diffusion"])
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            results = list(ex.map(obs, range(ntraj)))
        ens = {"ntraj": np.array(ntraj), "seed0": np.array(SEED), "steps": np.array([mid, last]), "bins": bins}
        for ti, s in enumerate((mid, last)):
            ens[f"t{ti}_totals"] = np.array([r[0][ti][0] for r in results])                 # [ntraj, S]
            ens[f"t{ti}_binned"] = np.array([r[0][ti][1] for r in results])                 # [ntraj, S, XBINS]
            allxx = np.array([r[0][ti][2] for r in results], dtype=np.float64)              # [ntraj, N, S]
            ens[f"t{ti}_vox_mean"] = allxx.mean(axis=0)
            ens[f"t{ti}_vox_var"] = allxx.var(axis=0, ddof=1)
        ens["events"] = np.array([r[1] for r in results])
        np.savez_compressed(os.path.join(HERE, f"{name}.ens.npz"), **ens)
        print(f"{name}: ensemble of {ntraj}: mean totals at last = {ens['t1_totals'].mean(axis=0)}, events/traj = {ens['events'].mean(axis=0)}")


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a != "vtk"] or (list(models.BUILDERS) if len(sys.argv) == 1 else [])
    for n in names:
        make(n)


def make_vtk(name="diffusion3d", keep=(0, 1, 10)):
    """Reference VTK byte-format fixtures (E/src/output.cpp:104-229) from the unmodified engine incl. its own output.cpp."""
    import gzip
    model = models.BUILDERS[name]()
    np.random.seed(12345)
    fm = FlatModel.from_spatialpy(model)
    np.random.seed(12345)
    exe = build_ref.build_model(model, name, variant="parity", dump=False, h=fm.h)
    d = tempfile.mkdtemp(prefix="ssb_golden_vtk_")
    build_ref.run_exe(exe, d, SEED, threads=1)
    out = os.path.join(HERE, f"vtk_{name}")
    os.makedirs(out, exist_ok=True)
    files = sorted(os.listdir(d))
    with open(os.path.join(out, "listing.txt"), "w") as f:
        f.write("\n".join(files) + "\n")
    for fn in ["output0_boundingBox.vtk"] + [f"output{k}.vtk" for k in keep]:
        with open(os.path.join(d, fn), "rb") as src, gzip.open(os.path.join(out, fn + ".gz"), "wb") as dst:
            dst.write(src.read())
    shutil.rmtree(d, ignore_errors=True)
    print(f"vtk fixtures for {name}: {files}")


if __name__ == "__main__" and (len(sys.argv) == 1 or "vtk" in sys.argv[1:]):
    make_vtk()
