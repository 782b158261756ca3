"""TEST SCAFFOLDING — builds the checkers that travel to the GPU box under oracle/_ref/ (called by
__graft_entry__.build(); building the checker is not using it).

  * bench reference executables: the UNMODIFIED reference engine (oracle/build_ref.py) baked with a bounded-size
    instance of each bench workload, variants `fast` (-O3) and `shipped` (no -O flag, E/build/SConstruct:23).
    The reference cannot be built at the bench's 1 M particles (one source line per particle, solver.py:312-331;
    13.8 s of g++ at 8 k), so the CPU arm is timed at the size stated in BENCH_REF below and reported per
    particle-step with that size named.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import build_ref  # noqa: E402

# name -> (config builder name, kwargs)
BENCH_REF = {
    "bench_cylinder": ("cylinder_rdme", dict(delta=0.125, nt=100, output_every=100, dt=1e-3)),
    "bench_tank": ("tank_sdpd", dict(n=26, nt=20, output_every=20, dt=1e-5)),
    # the bench's default workload (BASELINE configs[4] per-GPU unit, SDPD + sSSA) at a size the reference compiles: 24^3 = 13 824
    "bench_box": ("box_sdpd_rdme", dict(nx=24, ny=24, nz=24, nt=20, output_every=20, dt=1e-5)),
}
# the stepping loop = wall of the executable minus the wall of a ZERO-STEP run of the same model (process
# start, particle construction, the step-0 output); built for the default workload only
ZERO_STEP = ("bench_box",)


def build_bench_refs(variants=("fast", "shipped")):
    from spatialpy_b200 import configs
    out = {}
    for name, (builder, kw) in BENCH_REF.items():
        fm = getattr(configs, builder)(**kw)
        for variant in variants:
            exe = build_ref.build_flat(fm, name, variant=variant, dump=False, model_opt="-O1" if variant == "fast" else None)
            out[f"{name}/{variant}"] = exe
        if name in ZERO_STEP:
            kw0 = dict(kw, nt=0)
            fm0 = getattr(configs, builder)(**kw0)
            out[f"{name}/fast0"] = build_ref.build_flat(fm0, name + "_zero", variant="fast", dump=False, model_opt="-O1")
        meta = dict(N=fm.num_particles, nt=int(fm.nt), dt=float(fm.dt), builder=builder, kwargs=kw,
                    Sc=fm.num_chem_species, Sd=fm.num_stoch_species, R=fm.num_reactions)
        with open(os.path.join(build_ref.OUT, name, "meta.json"), "w") as f:
            json.dump(meta, f)
    return out


def build_ensemble_refs(names=("cdc42_full", "birth_death")):
    """BASELINE configs[3] (Cdc42 at its named size, the `ens_cdc42_full` bench workload) and configs[0] (birth-death) as timing
    executables: the unmodified reference through its own code generator (spatialpy front end), g++ -O3."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    build_ref.add_reference_to_path()
    import numpy as np
    import models
    from spatialpy_b200 import FlatModel
    out = {}
    for name in names:
        model = models.BUILDERS[name]()
        np.random.seed(12345)
        fm = FlatModel.from_spatialpy(model)
        np.random.seed(12345)
        out[name] = build_ref.build_model(model, f"bench_{name}", variant="fast", dump=False, h=fm.h)
        with open(os.path.join(build_ref.OUT, f"bench_{name}", "meta.json"), "w") as f:
            json.dump(dict(N=fm.num_particles, nt=int(fm.nt), dt=float(fm.dt), builder=f"tests/golden/models.py:{name}", kwargs={}), f)
    return out


def build_all():
    if not build_ref.reference_available():
        print("oracle: /root/reference not present — using the prebuilt oracle/_ref/")
        return
    for k, v in build_bench_refs().items():
        print("built", k, v)
    for k, v in build_ensemble_refs().items():
        print(f"built bench_{k}/fast", v)
    print("staged", build_ref.stage_reference_python())


if __name__ == "__main__":
    build_all()
