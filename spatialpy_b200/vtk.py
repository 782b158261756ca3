"""Reader for the engine's output files — same parsing rules as the reference's VTKReader
(spatialpy/core/vtkreader.py:80-194): 4 header lines, `POINTS n float` -> float32 points, everything up to the `FIELD`
line skipped, then arrays `name ncomp ntuples dtype` with `int` -> numpy int (int64) and `double` -> float64."""
import numpy as np


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
