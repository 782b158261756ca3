"""TEST INFRASTRUCTURE — builds and calls oracle/nsm_oracle.cpp for one FlatModel (g++, ctypes)."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def build(fm):
    from spatialpy_b200 import codegen
    hdr = codegen.generate_model_header(fm)
    src = open(os.path.join(HERE, "nsm_oracle.cpp")).read()
    key = hashlib.sha256((hdr + src).encode()).hexdigest()[:20]
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, f"nsm_{key}.so")
    if not os.path.exists(so):
        hpath = os.path.join(out_dir, f"model_{key}.h")
        with open(hpath, "w") as f:
            f.write(hdr)
        subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", f"-DSSB_MODEL_HEADER=\"{hpath}\"",
                        os.path.join(HERE, "nsm_oracle.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.nsm_oracle_run.restype = C.c_int
    return lib


def run(lib, fm, nbr, seed, t_end, corrected=False):
    """nbr: dict(ptr, j, Dij) from sdpd_oracle.SdpdOracle.find_neighbors; returns (xx[N,S], n_rx, n_df)."""
    N, S = fm.num_particles, fm.num_stoch_species
    xx = np.ascontiguousarray(fm.u0[:, :S], dtype=np.uint32).copy()
    ptr = np.ascontiguousarray(nbr["ptr"], np.int64)
    idx = np.ascontiguousarray(nbr["j"], np.int32)
    Dij = np.ascontiguousarray(nbr["Dij"], np.float64)
    vol = np.ascontiguousarray(fm.mass / fm.rho)
    nrx, ndf = C.c_int64(0), C.c_int64(0)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.nsm_oracle_run(C.c_int(N), P(ptr), P(idx), P(Dij), P(fm.type), P(vol), P(fm.data_fn), P(fm.diffusion_matrix),
                            C.c_int(fm.num_types), P(xx), C.c_double(0.0), C.c_double(t_end), C.c_uint64(seed),
                            C.c_int(int(corrected)), C.byref(nrx), C.byref(ndf))
    if rc:
        raise RuntimeError(f"nsm oracle failed rc={rc}")
    return xx, nrx.value, ndf.value
