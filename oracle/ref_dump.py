"""TEST SCAFFOLDING — reader for the binary taps written by oracle/ref_dump_output.cpp."""
import numpy as np


def read_dump(path):
    """Parse one dump_<step>.bin into a dict of numpy arrays (layout: ref_dump_output.cpp header)."""
    with open(path, "rb") as f:
        buf = f.read()
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a.copy()

    N, Sc, Sd, step, init, nnz = (int(v) for v in take(np.int64, 6))
    d = {"N": N, "Sc": Sc, "Sd": Sd, "step": step, "initialized": init}
    for name in ("x", "v", "vt", "F", "Fbp", "normal"):
        d[name] = take(np.float64, N * 3).reshape(N, 3)
    for name in ("rho", "old_rho", "Frho", "bvf_phi", "mass", "nu"):
        d[name] = take(np.float64, N)
    for name in ("type", "solid", "id"):
        d[name] = take(np.int32, N)
    d["C"] = take(np.float64, N * Sc).reshape(N, Sc)
    d["Q"] = take(np.float64, N * Sc).reshape(N, Sc)
    d["xx"] = take(np.uint32, N * Sd).reshape(N, Sd)
    d["nbr_ptr"] = take(np.int64, N + 1)
    d["nbr_idx"] = take(np.int32, nnz)
    d["nbr_dist"] = take(np.float64, nnz)
    d["nbr_dWdr"] = take(np.float64, nnz)
    d["nbr_Dij"] = take(np.float64, nnz)
    if init:
        R = int(take(np.int64, 1)[0])
        d["srrate"] = take(np.float64, N)
        d["sdrate"] = take(np.float64, N)
        d["Ddiag"] = take(np.float64, N * Sd).reshape(N, Sd)
        d["rrate"] = take(np.float64, N * R).reshape(N, R)
    cnt = take(np.int64, 2)
    d["total_reactions"], d["total_diffusion"] = int(cnt[0]), int(cnt[1])
    assert off == len(buf), (off, len(buf))
    return d
