"""ctypes binding of include/ssb_peaks.h: the measured fp64 FMA peak of a device (bench.py's fp64 companion of the HBM
roofline).  `python -m spatialpy_b200.peaks [device]` prints one JSON object — bench.py runs it as a child
process so that a failure of the microbenchmark can never take the bench line with it."""
import ctypes as C
import json
import os
import sys

from . import codegen


def fp64_peak(device=0):
    """(TFLOP/s, ms of the best launch) of an issue-bound DFMA loop on every SM."""
    if not os.path.exists(codegen.PEAKS_LIB):
        raise ImportError(f"{codegen.PEAKS_LIB} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(codegen.PEAKS_LIB)
    lib.ssb_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ssb_fp64_peak.restype = C.c_int
    tf, ms = C.c_double(0.0), C.c_double(0.0)
    rc = lib.ssb_fp64_peak(int(device), C.byref(tf), C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"ssb_fp64_peak failed, return code = {rc}")
    return tf.value, ms.value


if __name__ == "__main__":
    tflops, ms = fp64_peak(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    print(json.dumps({"fp64_tflops": tflops, "launch_ms": ms}))
