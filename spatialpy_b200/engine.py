"""ctypes binding of include/ssb.h — the thin layer between Python and the CUDA engine.

There is no CPU fallback: constructing an `Engine` without libssb_core.so (or without a usable GPU) raises.
"""
import ctypes as C
import os

# Ensemble lanes run one engine stream per concurrent trajectory.  With the default of 8 hardware work queues, more than 8
# streams alias onto the same queue and serialise behind each other's long sSSA kernels (measured: 16 lanes ran 4x slower per
# trajectory); 32 queues restore full concurrency.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# Slab ranks that run as THREADS of one process (Solver.run(decomposition="slab"), the 1-GPU tests) share one CUDA context, and a
# kernel's first launch under lazy module loading synchronises that context: a rank reaching a not-yet-loaded kernel would wait for
# its neighbour's spinning wait kernel, which waits for this rank's message.  Load every kernel when its module is loaded.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import numpy as np

from . import codegen
from .flatmodel import FlatModel

SSB_ABI_VERSION = 4
FLAG_CORRECTED_NSM_SELECT = 1
FLAG_CORRECTED_STOICH = 2
FLAG_NO_VTK = 4
FLAG_SKIP_STATIC_FORCES = 8
FLAG_LITERAL_KERNELS = 16
FLAG_LEAP_DIFFUSION = 32
FLAG_BINARY_STORE = 64
FLAG_NO_STEP_OVERSHOOT = 128
FLAG_CORRECTED_OUTPUT_STEPS = 256
FLAG_CORRECTED_PDE_INDEX = 512

ERR_NAMES = {1: "SSB_ERR_NAN", 2: "SSB_ERR_RDME", 3: "SSB_ERR_CUDA", 4: "SSB_ERR_ARG", 5: "SSB_ERR_IO",
             6: "SSB_ERR_CANCELLED", 7: "SSB_ERR_MODEL_UNIT", 8: "SSB_ERR_HALO"}

# every symbol include/ssb.h declares (tests check the library exports exactly these)
EXPORTS = ["ssb_abi_version", "ssb_device_count", "ssb_create", "ssb_load_kernels", "ssb_destroy", "ssb_run",
           "ssb_reset", "ssb_step", "ssb_counters", "ssb_get_field", "ssb_get_neighbors", "ssb_cancel",
           "ssb_last_error", "ssb_launch_count", "ssb_step_timed", "ssb_profile", "ssb_profile_read", "ssb_io_bytes",
           "ssb_nbr_stats", "ssb_step_phase", "ssb_halo_pack", "ssb_halo_unpack", "ssb_halo_inbox_pack", "ssb_halo_inbox_add",
           "ssb_halo_width", "ssb_mark", "ssb_mark_elapsed_ms", "ssb_skin_stats", "ssb_set_field", "ssb_get_step", "ssb_set_step", "ssb_write_snapshot",
           "ssb_slab_setup", "ssb_slab_blob_bytes", "ssb_slab_export", "ssb_slab_connect", "ssb_slab_disconnect", "ssb_slab_step"]

PH_PRE, PH_CORRECTOR, PH_FINISH, PH_RDME_PREP, PH_RDME_INIT, PH_RDME_WINDOW, PH_RDME_CLOSE, PH_END, PH_RDME_MIN, PH_RDME_EXTRA = range(10)

PROFILE_CATEGORIES = ["cells", "predictor", "search", "force", "corrector", "finish", "diff_init", "rdme_init",
                      "rdme_window", "output"]


class EngineError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"Solver execution failed, return code = {code} ({ERR_NAMES.get(code, '?')}): {message}")
        self.code = code
        self.message = message


class SsbModel(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("flags", C.c_uint32), ("n_particles", C.c_int64),
        ("dimension", C.c_int32), ("static_domain", C.c_int32), ("num_types", C.c_int32),
        ("num_chem_species", C.c_int32), ("num_chem_rxns", C.c_int32), ("num_stoch_species", C.c_int32),
        ("num_stoch_rxns", C.c_int32), ("num_data_fn", C.c_int32),
        ("dt", C.c_double), ("nt", C.c_uint32), ("n_output_steps", C.c_uint32),
        ("output_steps", C.POINTER(C.c_uint32)),
        ("h", C.c_double), ("rho0", C.c_double), ("c0", C.c_double), ("P0", C.c_double),
        ("xlo", C.c_double), ("xhi", C.c_double), ("ylo", C.c_double), ("yhi", C.c_double),
        ("zlo", C.c_double), ("zhi", C.c_double), ("gravity", C.c_double * 3),
        ("x", C.POINTER(C.c_double)), ("type", C.POINTER(C.c_int32)),
        ("nu", C.POINTER(C.c_double)), ("mass", C.POINTER(C.c_double)), ("c", C.POINTER(C.c_double)),
        ("rho", C.POINTER(C.c_double)), ("solid", C.POINTER(C.c_int32)), ("u0", C.POINTER(C.c_uint32)),
        ("data_fn", C.POINTER(C.c_double)), ("N_dense", C.POINTER(C.c_int32)),
        ("irN", C.POINTER(C.c_int64)), ("jcN", C.POINTER(C.c_int64)), ("prN", C.POINTER(C.c_int32)),
        ("irG", C.POINTER(C.c_int64)), ("jcG", C.POINTER(C.c_int64)),
        ("diffusion_matrix", C.POINTER(C.c_double)), ("species_names", C.POINTER(C.c_char_p)),
        ("rdme_epsilon", C.c_double), ("device", C.c_int32), ("reserved", C.c_int32),
        ("owned", C.POINTER(C.c_int32)), ("rng_id", C.POINTER(C.c_int32)),
    ]


_LIB = None


def load_library(path=None):
    """dlopen libssb_core.so (built in-tree by __graft_entry__.build / codegen.build_core)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or codegen.CORE_LIB
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the B200 engine has no CPU fallback)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    H = C.c_void_p
    lib.ssb_abi_version.restype = C.c_int
    lib.ssb_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.ssb_create.argtypes = [C.POINTER(SsbModel), C.POINTER(H)]
    lib.ssb_load_kernels.argtypes = [H, C.c_char_p]
    lib.ssb_destroy.argtypes = [H]
    lib.ssb_run.argtypes = [H, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p]
    lib.ssb_reset.argtypes = [H, C.c_uint64]
    lib.ssb_step.argtypes = [H, C.c_uint32]
    lib.ssb_counters.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.ssb_get_field.argtypes = [H, C.c_char_p, C.c_void_p, C.c_int64]
    lib.ssb_get_neighbors.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    lib.ssb_cancel.argtypes = [H]
    lib.ssb_last_error.argtypes = [H]
    lib.ssb_last_error.restype = C.c_char_p
    lib.ssb_launch_count.argtypes = [H, C.POINTER(C.c_int64)]
    lib.ssb_step_timed.argtypes = [H, C.c_uint32, C.POINTER(C.c_double)]
    lib.ssb_profile.argtypes = [H, C.c_int]
    lib.ssb_profile_read.argtypes = [H, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.ssb_io_bytes.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.ssb_nbr_stats.argtypes = [H, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    lib.ssb_step_phase.argtypes = [H, C.c_int, C.c_double, C.POINTER(C.c_double)]
    lib.ssb_halo_pack.argtypes = [H, C.c_int, C.c_void_p, C.c_int32, C.c_void_p]
    lib.ssb_halo_unpack.argtypes = [H, C.c_int, C.c_void_p, C.c_int32, C.c_void_p]
    lib.ssb_halo_inbox_pack.argtypes = [H, C.c_void_p, C.c_int32, C.c_void_p]
    lib.ssb_halo_inbox_add.argtypes = [H, C.c_void_p, C.c_int32, C.c_void_p]
    lib.ssb_halo_width.argtypes = [H, C.c_int, C.POINTER(C.c_int32)]
    lib.ssb_skin_stats.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.ssb_set_field.argtypes = [H, C.c_char_p, C.c_void_p, C.c_int64]
    lib.ssb_get_step.argtypes = [H, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    lib.ssb_set_step.argtypes = [H, C.c_uint32, C.c_uint64]
    lib.ssb_write_snapshot.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                       C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_uint32]
    lib.ssb_slab_setup.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    lib.ssb_slab_blob_bytes.argtypes = []
    lib.ssb_slab_export.argtypes = [H, C.c_void_p, C.c_int64]
    lib.ssb_slab_connect.argtypes = [H, C.c_void_p, C.c_int64]
    lib.ssb_slab_disconnect.argtypes = [H]
    lib.ssb_slab_step.argtypes = [H, C.c_uint32, C.c_double, C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
    lib.ssb_mark.argtypes = [H, C.c_int]
    lib.ssb_mark_elapsed_ms.argtypes = [H, C.POINTER(C.c_double)]
    for name in EXPORTS:
        if getattr(lib, name).restype is None:
            getattr(lib, name).restype = C.c_int
    if path == codegen.CORE_LIB:
        _LIB = lib
    return lib


def device_count():
    n = C.c_int(0)
    load_library().ssb_device_count(C.byref(n))
    return n.value


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class Engine:
    """One engine handle = one model on one GPU (include/ssb.h ssb_handle)."""

    _FIELDS = {  # name -> (dtype, columns or key)
        "x": (np.float64, 3), "v": (np.float64, 3), "vt": (np.float64, 3), "F": (np.float64, 3), "Fbp": (np.float64, 3),
        "x0": (np.float64, 3),
        "rho": (np.float64, 1), "old_rho": (np.float64, 1), "Frho": (np.float64, 1), "bvf_phi": (np.float64, 1),
        "mass": (np.float64, 1), "nu": (np.float64, 1), "srrate": (np.float64, 1), "sdrate": (np.float64, 1),
        "tnext": (np.float64, 1), "rho_search": (np.float64, 1),
        "type": (np.int32, 1), "solid": (np.int32, 1), "nbr_count": (np.int32, 1), "id": (np.int32, 1),
        "C": (np.float64, "Sc"), "Q": (np.float64, "Sc"), "Ddiag": (np.float64, "Sd"), "rrate": (np.float64, "Rd"),
        "xx": (np.uint32, "Sd"),
    }

    def __init__(self, fm: FlatModel, device=0, flags=FLAG_SKIP_STATIC_FORCES, rdme_epsilon=0.0, unit_path=None,
                 unit_flags=(), owned=None, rng_id=None):
        self.lib = load_library()
        self.fm = fm.finalize()
        self._h = C.c_void_p(None)
        m = SsbModel()
        self._keep = []
        m.abi_version = SSB_ABI_VERSION
        m.flags = int(flags)
        m.n_particles = fm.num_particles
        m.dimension = int(fm.dimension)
        m.static_domain = int(bool(fm.static_domain))
        m.num_types = int(fm.num_types)
        m.num_chem_species, m.num_chem_rxns = fm.num_chem_species, fm.num_chem_rxns
        m.num_stoch_species, m.num_stoch_rxns = fm.num_stoch_species, fm.num_stoch_rxns
        m.num_data_fn = fm.num_data_fn
        m.dt, m.nt = float(fm.dt), int(fm.nt)
        m.n_output_steps = int(fm.output_steps.shape[0])
        m.output_steps = _ptr(fm.output_steps, C.c_uint32)
        m.h, m.rho0, m.c0, m.P0 = float(fm.h), float(fm.rho0), float(fm.c0), float(fm.P0)
        m.xlo, m.xhi = fm.xlim
        m.ylo, m.yhi = fm.ylim
        m.zlo, m.zhi = fm.zlim
        m.gravity = (C.c_double * 3)(*fm.gravity)
        m.x = _ptr(fm.x, C.c_double)
        m.type = _ptr(fm.type, C.c_int32)
        m.nu, m.mass, m.c, m.rho = (_ptr(a, C.c_double) for a in (fm.nu, fm.mass, fm.c, fm.rho))
        m.solid = _ptr(fm.solid, C.c_int32)
        m.u0 = _ptr(fm.u0, C.c_uint32)
        m.data_fn = _ptr(fm.data_fn, C.c_double)
        m.N_dense = _ptr(fm.N_dense, C.c_int32)
        m.irN, m.jcN, m.prN = _ptr(fm.irN, C.c_int64), _ptr(fm.jcN, C.c_int64), _ptr(fm.prN, C.c_int32)
        m.irG, m.jcG = _ptr(fm.irG, C.c_int64), _ptr(fm.jcG, C.c_int64)
        m.diffusion_matrix = _ptr(fm.diffusion_matrix, C.c_double)
        names = (C.c_char_p * max(1, fm.num_species))(*[n.encode() for n in fm.species_names])
        m.species_names = C.cast(names, C.POINTER(C.c_char_p))
        m.rdme_epsilon = float(rdme_epsilon)
        m.device = int(device)
        self._owned = None if owned is None else np.ascontiguousarray(owned, dtype=np.int32)
        self._rng_id = None if rng_id is None else np.ascontiguousarray(rng_id, dtype=np.int32)
        m.owned = None if self._owned is None else _ptr(self._owned, C.c_int32)
        m.rng_id = None if self._rng_id is None else _ptr(self._rng_id, C.c_int32)
        self._check(self.lib.ssb_create(C.byref(m), C.byref(self._h)))
        self.unit_path = unit_path or codegen.build_model_unit(fm, extra_flags=unit_flags)
        self._check(self.lib.ssb_load_kernels(self._h, self.unit_path.encode()))
        self.N = fm.num_particles
        self._sizes = {"Sc": fm.num_chem_species, "Sd": fm.num_stoch_species, "Rd": fm.num_stoch_rxns}

    # ------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self.lib.ssb_last_error(self._h) if self._h else b""
            raise EngineError(rc, (msg or b"").decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.ssb_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------
    def reset(self, seed):
        self._check(self.lib.ssb_reset(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF))

    def step(self, n=1):
        self._check(self.lib.ssb_step(self._h, int(n)))

    def run(self, seed, out_dirs, first_traj=0):
        """Run len(out_dirs) trajectories (seed+first_traj+k each), writing VTK files into out_dirs[k]."""
        n = len(out_dirs)
        arr = (C.c_char_p * max(1, n))(*[os.fsencode(d) for d in out_dirs])
        self._check(self.lib.ssb_run(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, n, int(first_traj),
                                     C.cast(arr, C.POINTER(C.c_char_p)), None, None))

    def run_no_files(self, seed, ntraj=1, first_traj=0):
        self._check(self.lib.ssb_run(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(ntraj), int(first_traj), None, None, None))

    def cancel(self):
        self.lib.ssb_cancel(self._h)

    def counters(self):
        r, d, w = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        s = C.c_double(0)
        self._check(self.lib.ssb_counters(self._h, C.byref(r), C.byref(d), C.byref(s), C.byref(w)))
        return {"reactions": r.value, "diffusions": d.value, "seconds": s.value, "windows": w.value}

    def launch_count(self):
        n = C.c_int64(0)
        self._check(self.lib.ssb_launch_count(self._h, C.byref(n)))
        return n.value

    def step_timed(self, n=1):
        """n engine steps; returns the device time in ms (CUDA events on the engine stream)."""
        ms = C.c_double(0)
        self._check(self.lib.ssb_step_timed(self._h, int(n), C.byref(ms)))
        return ms.value

    def profile(self, enable=True):
        self._check(self.lib.ssb_profile(self._h, int(bool(enable))))

    def profile_read(self):
        out = {}
        for k, name in enumerate(PROFILE_CATEGORIES):
            ms, n = C.c_double(0), C.c_int64(0)
            self._check(self.lib.ssb_profile_read(self._h, k, C.byref(ms), C.byref(n)))
            out[name] = {"ms": ms.value, "launches": n.value}
        return out

    def io_bytes(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.ssb_io_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def nbr_stats(self):
        cap, tot = C.c_int32(0), C.c_int64(0)
        self._check(self.lib.ssb_nbr_stats(self._h, C.byref(cap), C.byref(tot)))
        return cap.value, tot.value

    # -- slab decomposition primitives (spatialpy_b200/slab.py) ---------------------------------------------------
    def phase(self, phase, arg=0.0):
        out = C.c_double(0.0)
        self._check(self.lib.ssb_step_phase(self._h, int(phase), float(arg), C.byref(out)))
        return out.value

    def halo_width(self, group):
        w = C.c_int32(0)
        self._check(self.lib.ssb_halo_width(self._h, int(group), C.byref(w)))
        return w.value

    def halo_pack(self, group, ids_ptr, n, out_ptr):
        self._check(self.lib.ssb_halo_pack(self._h, int(group), ids_ptr, int(n), out_ptr))

    def halo_unpack(self, group, ids_ptr, n, in_ptr):
        self._check(self.lib.ssb_halo_unpack(self._h, int(group), ids_ptr, int(n), in_ptr))

    def inbox_pack(self, ids_ptr, n, out_ptr):
        self._check(self.lib.ssb_halo_inbox_pack(self._h, ids_ptr, int(n), out_ptr))

    def inbox_add(self, ids_ptr, n, in_ptr):
        self._check(self.lib.ssb_halo_inbox_add(self._h, ids_ptr, int(n), in_ptr))

    # -- native slab transport: windows in peer-mapped memory, stream-ordered exchanges (include/ssb.h ssb_slab_*) ----------
    def slab_setup(self, rank, world, send_ids, recv_ids):
        """send_ids / recv_ids: {neighbour rank: int32 local particle ids} as SlabPartition holds them."""
        arrs = []
        for nb in (rank - 1, rank + 1):
            for d in (send_ids, recv_ids):
                arrs.append(np.ascontiguousarray(d.get(nb, np.zeros(0, np.int32)), dtype=np.int32))
        args = []
        for a in arrs:
            args += [a.ctypes.data_as(C.c_void_p), int(a.size)]
        self._check(self.lib.ssb_slab_setup(self._h, int(rank), int(world), *args))

    def slab_export(self):
        n = self.lib.ssb_slab_blob_bytes()
        buf = C.create_string_buffer(n)
        self._check(self.lib.ssb_slab_export(self._h, buf, n))
        return buf.raw

    def slab_connect(self, blobs):
        """blobs: the slab_export() of every rank, in rank order."""
        raw = b"".join(blobs)
        self._check(self.lib.ssb_slab_connect(self._h, raw, len(raw)))

    def slab_disconnect(self):
        if getattr(self, "_h", None):
            self.lib.ssb_slab_disconnect(self._h)

    def slab_step(self, n, travel_limit=0.0):
        """n engine steps with device-side halo exchanges; returns (steps done, travel bound)."""
        done, travel = C.c_uint32(0), C.c_double(0.0)
        self._check(self.lib.ssb_slab_step(self._h, int(n), float(travel_limit), C.byref(done), C.byref(travel)))
        return done.value, travel.value

    def skin_stats(self):
        a, b, n = C.c_double(0), C.c_double(0), C.c_int64(0)
        self._check(self.lib.ssb_skin_stats(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return {"skin": a.value, "step_disp_max": b.value, "rebuilds": n.value}

    def mark(self, which):
        self._check(self.lib.ssb_mark(self._h, int(which)))

    def mark_elapsed_ms(self):
        ms = C.c_double(0.0)
        self._check(self.lib.ssb_mark_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def get(self, name):
        dtype, cols = self._FIELDS[name]
        k = self._sizes[cols] if isinstance(cols, str) else cols
        out = np.empty((self.N, k) if (k != 1 or isinstance(cols, str)) else (self.N,), dtype=dtype)
        self._check(self.lib.ssb_get_field(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    _SETTABLE = ("x", "v", "vt", "F", "Fbp", "rho", "old_rho", "Frho", "bvf_phi", "nu", "C", "Q", "xx")

    def set(self, name, values):
        """Replace a state field (particle-id order, same shapes as `get`) — the hand-over of a slab re-partition."""
        if name not in self._SETTABLE:
            raise KeyError(f"field '{name}' cannot be set")
        dtype, cols = self._FIELDS[name]
        k = self._sizes[cols] if isinstance(cols, str) else cols
        a = np.ascontiguousarray(values, dtype=dtype)
        if a.size != self.N * k:
            raise ValueError(f"field '{name}' needs {self.N * k} values, got {a.size}")
        self._check(self.lib.ssb_set_field(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.nbytes))

    def get_step(self):
        """(engine step counter, Philox window epoch) of the running trajectory."""
        step, epoch = C.c_uint32(0), C.c_uint64(0)
        self._check(self.lib.ssb_get_step(self._h, C.byref(step), C.byref(epoch)))
        return step.value, epoch.value

    def set_step(self, step, epoch):
        self._check(self.lib.ssb_set_step(self._h, int(step), int(epoch)))

    def neighbors(self):
        """Neighbour lists in id space: (ptr[N+1], idx, dist, dWdr, Dij) — same shape as oracle dumps."""
        nnz = C.c_int64(0)
        self._check(self.lib.ssb_get_neighbors(self._h, None, None, None, None, None, C.byref(nnz)))
        n = nnz.value
        ptr = np.empty(self.N + 1, np.int64)
        idx = np.empty(n, np.int32)
        dist, dWdr, Dij = (np.empty(n, np.float64) for _ in range(3))
        self._check(self.lib.ssb_get_neighbors(self._h, ptr.ctypes.data, idx.ctypes.data, dist.ctypes.data,
                                               dWdr.ctypes.data, Dij.ctypes.data, C.byref(nnz)))
        return ptr, idx, dist, dWdr, Dij
