"""Reader for the engine's output files — same parsing rules as the reference's VTKReader
(spatialpy/core/vtkreader.py:80-194): 4 header lines, `POINTS n float` -> float32 points, everything up to the `FIELD`
line skipped, then arrays `name ncomp ntuples dtype` with `int` -> numpy int (int64) and `double` -> float64."""
import json
import os

import numpy as np

SSB_MAGIC = b"SSBOUT1\0"


def read_vtk(path):
    with open(path, "r", encoding="utf-8") as f:
        lines = f.read().split("\n")
    k = 4
    _, n, _ = lines[k].split()
    n = int(n)
    k += 1
    vals = []
    while len(vals) < 3 * n:
        vals.extend(lines[k].split())
        k += 1
    points = np.array(vals, dtype=np.float32).reshape(n, 3)
    while not lines[k].startswith("FIELD"):
        k += 1
    nfields_header = int(lines[k].split()[2])
    k += 1
    arrays = {}
    while k < len(lines):
        parts = lines[k].split()
        k += 1
        if len(parts) != 4:
            continue
        name, ncomp, ntup, dtype = parts[0], int(parts[1]), int(parts[2]), parts[3]
        vals = []
        while len(vals) < ncomp * ntup:
            vals.extend(lines[k].split())
            k += 1
        a = np.array(vals, dtype=np.int64 if dtype == "int" else np.float64)
        arrays[name] = a.reshape(ntup, ncomp) if ncomp > 1 else a
    arrays["__nfields_header__"] = nfields_header
    return points, arrays


def _ssb_layout(np_, Sc, Sd):
    """(name, dtype, count) of the raw sections of an outputN.ssb file, in file order (ssb_core.cu write_bin)."""
    return [("x", "<f8", 3 * np_), ("v", "<f8", 3 * np_), ("scal", "<f8", 4 * np_), ("C", "<f8", Sc * np_),
            ("type", "<i4", np_), ("D", "<u4", Sd * np_)]


def read_ssb(path):
    """Read the binary side-store written under SSB_FLAG_BINARY_STORE.  Returns the same `(points, arrays)` pair, keys, shapes
    and dtypes as `read_vtk` / the reference's VTKReader (vtkreader.py:29-56,160-194) — float32 points, `int` arrays as int64 —
    but the fp64 fields carry full precision instead of the six decimals `%lf` leaves in the text file (output.cpp:170-230),
    and a 1 M-particle snapshot loads in milliseconds instead of the ~10 s the ASCII parser needs."""
    with open(path, "rb") as f:
        hdr, _ = _ssb_header(f, path)
        n, Sc, Sd = hdr["np"], hdr["Sc"], hdr["Sd"]
        raw = {}
        for name, dt, cnt in _ssb_layout(n, Sc, Sd):
            raw[name] = np.fromfile(f, dtype=dt, count=cnt)
            if raw[name].size != cnt:
                raise ValueError(f"{path} is truncated in section {name}")
    arrays = {"id": np.arange(n, dtype=np.int64), "type": raw["type"].astype(np.int64), "v": raw["v"].reshape(n, 3)}
    for k, name in enumerate(("rho", "mass", "bvf_phi", "nu")):
        arrays[name] = raw["scal"][k * n:(k + 1) * n]
    for s in range(Sc):
        arrays[f"C[{hdr['species'][s]}]"] = raw["C"][s * n:(s + 1) * n]
    for s in range(Sd):
        arrays[f"D[{hdr['species'][s]}]"] = raw["D"][s * n:(s + 1) * n].astype(np.int64)
    # output.cpp:151-154 undercounts FIELD in output0 (the RDME is not initialised yet); keep the reader-visible number
    arrays["__nfields_header__"] = 7 + Sc + (Sd if hdr["rdme_initialized"] else 0)
    return raw["x"].reshape(n, 3).astype(np.float32), arrays


def _ssb_header(f, path):
    if f.read(8) != SSB_MAGIC:
        raise ValueError(f"{path} is not an SSB output file")
    hl = int(np.frombuffer(f.read(8), dtype="<u8")[0])
    return json.loads(f.read(hl).decode("ascii")), 16 + hl


def read_ssb_field(path, key):
    """One array of an outputN.ssb file, located by offset instead of loading the whole snapshot — what the all-timepoints
    getters of Result (`get_species`, `get_property`; result.py:334-402,601-655) need: one field of every step.  Same keys,
    shapes and dtypes as `read_ssb`; raises KeyError for a name the snapshot does not hold."""
    with open(path, "rb") as f:
        hdr, base = _ssb_header(f, path)
        n, Sc, Sd = hdr["np"], hdr["Sc"], hdr["Sd"]
        off = {}
        for name, dt, cnt in _ssb_layout(n, Sc, Sd):
            off[name] = base
            base += cnt * np.dtype(dt).itemsize

        def sect(name, dt, skip, cnt):
            f.seek(off[name] + skip * np.dtype(dt).itemsize)
            a = np.fromfile(f, dtype=dt, count=cnt)
            if a.size != cnt:
                raise ValueError(f"{path} is truncated in section {name}")
            return a
        if key == "id":
            return np.arange(n, dtype=np.int64)
        if key == "type":
            return sect("type", "<i4", 0, n).astype(np.int64)
        if key == "v":
            return sect("v", "<f8", 0, 3 * n).reshape(n, 3)
        if key == "points":
            return sect("x", "<f8", 0, 3 * n).reshape(n, 3).astype(np.float32)
        if key in ("rho", "mass", "bvf_phi", "nu"):
            return sect("scal", "<f8", ("rho", "mass", "bvf_phi", "nu").index(key) * n, n)
        if len(key) > 3 and key[1] == "[" and key[-1] == "]" and key[2:-1] in hdr["species"]:
            s = hdr["species"].index(key[2:-1])
            if key[0] == "C" and s < Sc:
                return sect("C", "<f8", s * n, n)
            if key[0] == "D" and s < Sd:
                return sect("D", "<u4", s * n, n).astype(np.int64)
        raise KeyError(key)


def write_ssb(path, x, v, scal, C, type_, D, species, step=0, rdme_initialized=1):
    """Python twin of ssb_core.cu's write_bin (tests pin the two against each other)."""
    n = len(type_)
    C = np.zeros((0, n)) if C is None else np.asarray(C, dtype="<f8").reshape(-1, n)
    D = np.zeros((0, n), dtype="<u4") if D is None else np.asarray(D, dtype="<u4").reshape(-1, n)
    hdr = json.dumps({"np": n, "Sc": int(C.shape[0]), "Sd": int(D.shape[0]), "step": int(step),
                      "rdme_initialized": int(rdme_initialized), "species": list(species)})
    while (16 + len(hdr)) % 64:
        hdr += " "
    with open(path, "wb") as f:
        f.write(SSB_MAGIC)
        f.write(np.array([len(hdr)], dtype="<u8").tobytes())
        f.write(hdr.encode("ascii"))
        for a, dt in ((x, "<f8"), (v, "<f8"), (scal, "<f8"), (C, "<f8"), (type_, "<i4"), (D, "<u4")):
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())


def _section(f, fmt, values, per_line, chunk=90000):
    """`values` formatted with `fmt` + one blank each, a newline after every `per_line`-th item — the layout of every
    array of the reference's writer (output.cpp:133-229); written in chunks that end on a line boundary."""
    values = np.asarray(values)
    for lo in range(0, len(values), chunk):
        part = values[lo:lo + chunk]
        ends = np.full(len(part), " ", dtype="<U2")
        ends[per_line - 1::per_line] = " \n"
        f.write("".join(np.char.add(np.char.mod(fmt, part), ends).tolist()))


def write_vtk(path, x, v, scal, C, type_, D, species, rdme_initialized=1):
    """Python twin of the engine's VTK writer (ssb_core.cu write_vtk), i.e. of the reference's `output_vtk__async_step`
    (E/src/output.cpp:130-229): ASCII VTK 4.1 POLYDATA, `POINTS n float` as `%.10e` three points per line, `VERTICES`,
    `FIELD FieldData k` with k = 7 + S_c + (S_d once the RDME is initialised — the reference undercounts in output0,
    output.cpp:151-154), arrays `id type` (`%u`), `v` (`%lf`, one particle's three components, three particles per line),
    `rho mass bvf_phi nu C[name]` (`%lf`, nine per line), `D[name]` (`%u`).  Used where a snapshot is assembled on the host
    (slab-decomposed runs gather the owned particles of every rank); pinned byte for byte against files written by the
    reference itself (tests/golden/vtk_diffusion3d, tests/test_cpu_abi.py).  Arguments as `write_ssb`."""
    n = len(type_)
    x = np.asarray(x, dtype=np.float64).reshape(n, 3)
    v = np.asarray(v, dtype=np.float64).reshape(n, 3)
    scal = np.asarray(scal, dtype=np.float64).reshape(4, n)
    C = np.zeros((0, n)) if C is None else np.asarray(C, dtype=np.float64).reshape(-1, n)
    D = np.zeros((0, n), dtype=np.uint32) if D is None else np.asarray(D, dtype=np.uint32).reshape(-1, n)
    Sc, Sd = C.shape[0], D.shape[0]
    with open(path, "w", encoding="ascii", newline="") as f:
        f.write("# vtk DataFile Version 4.1\nGenerated by SpatialPy\nASCII\nDATASET POLYDATA\n")
        f.write(f"POINTS {n} float\n")
        _section(f, "%.10e", x.reshape(-1), 9)
        f.write(f"\nVERTICES {n} {2 * n}\n")
        for lo in range(0, n, 90000):
            f.write("".join(f"1 {i}\n" for i in range(lo, min(n, lo + 90000))))
        f.write(f"\nPOINT_DATA {n}\n")
        f.write(f"FIELD FieldData {7 + Sc + (Sd if rdme_initialized else 0)}\n")
        f.write(f"id 1 {n} int\n")
        _section(f, "%u", np.arange(n, dtype=np.int64), 9)
        f.write(f"\ntype 1 {n} int\n")
        _section(f, "%u", np.asarray(type_, dtype=np.int64), 9)
        f.write(f"\nv 3 {n} double\n")
        _section(f, "%f", v.reshape(-1), 9)
        f.write("\n")
        for k, name in enumerate(("rho", "mass", "bvf_phi", "nu")):
            f.write(f"{name} 1 {n} double\n")
            _section(f, "%f", scal[k], 9)
            f.write("\n")
        for s_ in range(Sc):
            f.write(f"C[{species[s_]}] 1 {n} double\n")
            _section(f, "%f", C[s_], 9)
            f.write("\n")
        for s_ in range(Sd):
            f.write(f"D[{species[s_]}] 1 {n} int\n")
            _section(f, "%u", D[s_].astype(np.int64), 9)
            f.write("\n")


def write_bounding_box(result_dir, xlim, ylim, zlim):
    """output0_boundingBox.vtk (E/src/output.cpp:110-128)."""
    with open(os.path.join(result_dir, "output0_boundingBox.vtk"), "w", encoding="ascii", newline="") as f:
        f.write("# vtk DataFile Version 4.1\nGenerated by ssa_sdpd\nASCII\nDATASET RECTILINEAR_GRID\nDIMENSIONS 2 2 2\n")
        for axis, (lo, hi) in zip("XYZ", (xlim, ylim, zlim)):
            f.write(f"{axis}_COORDINATES 2 double\n%f %f\n" % (lo, hi))


def write_snapshot(result_dir, file_index, x, v, scal, C, type_, D, species, lims, step=0, rdme_initialized=1, vtk=True,
                   binary=False):
    """One output of a run from HOST arrays through the engine's own C++ writers (`ssb_write_snapshot`, include/ssb.h): the
    byte format of E/src/output.cpp:104-229 (and/or the outputN.ssb side-store), multi-threaded for large snapshots — what
    slab-decomposed runs and batched ensembles call once they have assembled a snapshot.  Arguments as `write_vtk`;
    lims = (xlim, ylim, zlim); file 0 also gets output0_boundingBox.vtk."""
    import ctypes as C_
    from .engine import load_library
    n = len(type_)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(n, 3)
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(n, 3)
    scal = np.ascontiguousarray(scal, dtype=np.float64).reshape(4, n)
    Cc = None if C is None else np.ascontiguousarray(C, dtype=np.float64).reshape(-1, n)
    Dd = None if D is None else np.ascontiguousarray(D, dtype=np.uint32).reshape(-1, n)
    Sc = 0 if Cc is None else Cc.shape[0]
    Sd = 0 if Dd is None else Dd.shape[0]
    typ = np.ascontiguousarray(type_, dtype=np.int32)
    names = (C_.c_char_p * max(1, len(species)))(*[str(s_).encode() for s_ in species])
    lim = np.array([lims[0][0], lims[0][1], lims[1][0], lims[1][1], lims[2][0], lims[2][1]], dtype=np.float64)
    ptr = lambda a: None if a is None or a.size == 0 else a.ctypes.data_as(C_.c_void_p)      # noqa: E731
    rc = load_library().ssb_write_snapshot(os.fsencode(result_dir), int(file_index), int(step), int(rdme_initialized), n, Sc, Sd,
                                           C_.cast(names, C_.POINTER(C_.c_char_p)), ptr(lim), ptr(x), ptr(v), ptr(scal),
                                           ptr(Cc) if Sc else None, ptr(typ), ptr(Dd) if Sd else None,
                                           (1 if vtk else 0) | (2 if binary else 0))
    if rc != 0:
        raise OSError(f"ssb_write_snapshot failed, return code = {rc} (directory {result_dir})")


def write_snapshot_py(result_dir, file_index, x, v, scal, C, type_, D, species, lims, step=0, rdme_initialized=1, vtk=True,
                      binary=False):
    """`write_snapshot` with the pure-Python twins of the writers (same files, no libssb_core.so needed; slow for large N)."""
    if file_index == 0 and vtk:
        write_bounding_box(result_dir, *lims)
    if vtk:
        write_vtk(os.path.join(result_dir, f"output{file_index}.vtk"), x, v, scal, C, type_, D, species, rdme_initialized=rdme_initialized)
    if binary:
        write_ssb(os.path.join(result_dir, f"output{file_index}.ssb"), x, v, scal, C, type_, D, species, step=step,
                  rdme_initialized=rdme_initialized)


def read_output(result_dir, step_num):
    """outputN.ssb when the run kept a binary side-store, else outputN.vtk."""
    b = os.path.join(result_dir, f"output{step_num}.ssb")
    if os.path.exists(b):
        return read_ssb(b)
    return read_vtk(os.path.join(result_dir, f"output{step_num}.vtk"))
