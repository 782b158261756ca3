"""TEST SCAFFOLDING — builds the UNMODIFIED reference engine as the parity/timing oracle.

Nothing here is imported by the product path (spatialpy_b200/).  Only tests/, bench.py's
cpu_baseline / --impl reference legs, __graft_entry__.build() and the golden-fixture generators use it.

The reference's own build goes through SCons (spatialpy/solvers/solver.py:474-492), which is not
installed here, so this recipe compiles the reference sources *where they lie* under
/root/reference with g++ directly and writes outputs only under oracle/_ref/
(git-ignored, but shipped to the GPU box by gpurun like any other built artefact).

Variants
  shipped : -std=c++14, no optimisation flag  (what E/build/SConstruct:23 does)       -> timing baseline
  fast    : -std=c++14 -O3                                                            -> timing baseline
  parity  : -std=c++14 -O2 -ffp-contract=off -ftrivial-auto-var-init=zero             -> parity oracle
            (zero-fills the stack temporary Particle in init_create_particle, template:74, so the
            fields the reference never initialises — F, Fbp, Frho, vt, bvf_phi, normal, old_* —
            are a defined 0 instead of stack noise; particle.cpp:70-83)
  dump=True replaces E/src/output.cpp by oracle/ref_dump_output.cpp (full-precision binary taps).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SSB_REFERENCE_ROOT", "/root/reference")
ENGINE = os.path.join(REF_ROOT, "spatialpy/solvers/c_base/ssa_sdpd-c-simulation-engine")
OUT = os.path.join(HERE, "_ref")

CORE_SOURCES = ["count_cores", "model", "NRMConstant_v5", "output", "particle", "pthread_barrier",
                "simulate", "simulate_rdme", "simulate_threads"]  # E/src/SConscript:2-12

FLAGS = {
    "shipped": ["-std=c++14", "-w"],
    "fast": ["-std=c++14", "-w", "-O3"],
    "parity": ["-std=c++14", "-w", "-O2", "-ffp-contract=off", "-ftrivial-auto-var-init=zero"],
}


def reference_available():
    return os.path.isdir(ENGINE)


def add_reference_to_path():
    """Make `import spatialpy` work: reference checkout + the no-op plotly stub."""
    for p in (os.path.join(HERE, "stubs"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)


def stage_reference_python():
    """Copy the reference's PYTHON front end (spatialpy/ without the C++ engine tree), its own test models and the no-op plotly
    stub into the untracked oracle/_ref/py/ so that the GPU box — which has no /root/reference — can `import spatialpy` and run the
    reference's front door (Model.run -> Solver -> Result) against the CUDA engine (tests/test_gpu_frontdoor.py).  Build output
    only: oracle/_ref/ is git-ignored and never imported by the product path."""
    import shutil
    dst = os.path.join(OUT, "py")
    if not reference_available():
        return dst if os.path.isdir(os.path.join(dst, "spatialpy")) else None
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    shutil.copytree(os.path.join(REF_ROOT, "spatialpy"), os.path.join(dst, "spatialpy"),
                    ignore=shutil.ignore_patterns("c_base", "__pycache__", "*.pyc"))
    shutil.copytree(os.path.join(HERE, "stubs", "plotly"), os.path.join(dst, "plotly"), ignore=shutil.ignore_patterns("__pycache__"))
    shutil.copytree(os.path.join(REF_ROOT, "test", "models"), os.path.join(dst, "ref_test_models"), ignore=shutil.ignore_patterns("__pycache__"))
    return dst


def staged_python_path():
    """sys.path entry under which `import spatialpy` finds the staged reference front end (None if it was never staged)."""
    dst = os.path.join(OUT, "py")
    return dst if os.path.isdir(os.path.join(dst, "spatialpy")) else None


def _run(cmd, **kw):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if res.returncode != 0:
        raise RuntimeError(f"command failed: {' '.join(cmd)}\n{res.stdout}")
    return res.stdout


def build_core(variant="parity", dump=False):
    """Static library of the reference engine + vendored ANN; cached (sources never change)."""
    tag = f"core_{variant}{'_dump' if dump else ''}"
    cdir = os.path.join(OUT, tag)
    lib = os.path.join(cdir, "libcore.a")
    if os.path.exists(lib):
        return lib
    if not reference_available():
        raise RuntimeError("reference checkout not present; oracle/_ref must be prebuilt")
    os.makedirs(cdir, exist_ok=True)
    inc = ["-I", os.path.join(ENGINE, "include"), "-I", os.path.join(ENGINE, "external/ANN/include")]
    jobs = []
    for s in CORE_SOURCES:
        src = os.path.join(ENGINE, "src", s + ".cpp")
        if s == "output" and dump:
            src = os.path.join(HERE, "ref_dump_output.cpp")
        jobs.append((src, os.path.join(cdir, s + ".o")))
    anndir = os.path.join(ENGINE, "external/ANN/src")
    for f in sorted(os.listdir(anndir)):
        if f.endswith(".cpp"):
            jobs.append((os.path.join(anndir, f), os.path.join(cdir, "ann_" + f[:-4] + ".o")))

    def cc(job):
        _run(["g++"] + FLAGS[variant] + inc + ["-c", job[0], "-o", job[1]])
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(cc, jobs))
    _run(["ar", "rcs", lib] + [j[1] for j in jobs])
    return lib


def generate_model_source(model, path, debug_level=0, h=None):
    """Reference codegen without SCons: Solver.__create_propensity_file (solver.py:100-158)."""
    add_reference_to_path()
    from spatialpy.solvers.solver import Solver
    sm, dg = model.compile_prep()
    sol = Solver(model, debug_level=debug_level)
    if h is not None:
        sol.h = h
    sol._Solver__create_propensity_file(sm, dg, file_name=path)
    return sol


def build_model(model, name, variant="parity", dump=False, debug_level=0, h=None, model_opt=None):
    """Generate + compile + link one reference executable for `model`; returns the exe path."""
    lib = build_core(variant, dump)
    mdir = os.path.join(OUT, name, f"{variant}{'_dump' if dump else ''}")
    os.makedirs(mdir, exist_ok=True)
    src = os.path.join(mdir, "model.cpp")
    exe = os.path.join(mdir, "ssa_sdpd.exe")
    generate_model_source(model, src + ".new", debug_level=debug_level, h=h)
    with open(src + ".new", "rb") as f:
        new_hash = hashlib.sha256(f.read()).hexdigest()
    stamp = os.path.join(mdir, "model.sha256")
    if os.path.exists(exe) and os.path.exists(stamp) and open(stamp).read() == new_hash:
        os.remove(src + ".new")
        return exe
    os.replace(src + ".new", src)
    flags = list(FLAGS[variant])
    if model_opt is not None:  # big literal-laden TUs: allow -O0 for the model file only
        flags = [f for f in flags if not f.startswith("-O")] + [model_opt]
    inc = ["-I", os.path.join(ENGINE, "include"), "-I", os.path.join(ENGINE, "external/ANN/include")]
    _run(["g++"] + flags + inc + [src, lib, "-lpthread", "-o", exe])
    with open(stamp, "w") as f:
        f.write(new_hash)
    return exe


def build_flat(fm, name, variant="parity", dump=False, debug_level=0, model_opt=None):
    """Same as build_model, for a FlatModel (synthetic configs): the reference's own template is filled by
    oracle/emit_ref_model.py with the substitutions of solver.py:135-153."""
    import emit_ref_model
    lib = build_core(variant, dump)
    mdir = os.path.join(OUT, name, f"{variant}{'_dump' if dump else ''}")
    os.makedirs(mdir, exist_ok=True)
    src = os.path.join(mdir, "model.cpp")
    exe = os.path.join(mdir, "ssa_sdpd.exe")
    emit_ref_model.emit(fm, src + ".new", os.path.join(ENGINE, "propensity_file_template.cpp"), debug_level)
    with open(src + ".new", "rb") as f:
        new_hash = hashlib.sha256(f.read()).hexdigest()
    stamp = os.path.join(mdir, "model.sha256")
    if os.path.exists(exe) and os.path.exists(stamp) and open(stamp).read() == new_hash:
        os.remove(src + ".new")
        return exe
    os.replace(src + ".new", src)
    flags = list(FLAGS[variant])
    if model_opt is not None:
        flags = [f for f in flags if not f.startswith("-O")] + [model_opt]
    inc = ["-I", os.path.join(ENGINE, "include"), "-I", os.path.join(ENGINE, "external/ANN/include")]
    _run(["g++"] + flags + inc + [src, lib, "-lpthread", "-o", exe])
    os.remove(src)      # the literal-laden TU can be hundreds of MB; the stamp identifies it
    with open(stamp, "w") as f:
        f.write(new_hash)
    return exe


def run_exe(exe, out_dir, seed, threads=1, timeout=None):
    """`exe -t T -s SEED` with cwd = out_dir (the reference's process contract, solver.py:553-569)."""
    os.makedirs(out_dir, exist_ok=True)
    return _run([exe, "-t", str(threads), "-s", str(seed)], cwd=out_dir, timeout=timeout)
