"""Regenerate profiles/r1_resource_usage.md: static resource usage (cuobjdump --dump-resource-usage) and load-instruction
variants (cuobjdump -sass) of the in-tree binaries.  Needs no GPU:  python profiles/make_resource_usage.py"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spatialpy_b200 import codegen, configs  # noqa: E402


def usage(path, title, out):
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", path], capture_output=True, text=True).stdout
    out.append(f"## {title}  ({os.path.relpath(path, ROOT)})")
    out.append("| kernel | registers | shared B | stack / local B | const B |")
    out.append("|---|---|---|---|---|")
    fn = None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+) CONSTANT\[0\]:(\d+)", line)
        if m and fn:
            out.append(f"| `{fn}` | {m.group(1)} | {m.group(3)} | {max(int(m.group(2)), int(m.group(4)))} | {m.group(5)} |")
            fn = None
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cnt = {k: len(re.findall(k, sass)) for k in ("DFMA", "DMUL", "DADD", "MUFU.RCP64H", "SHFL", "ATOM", "RED", "BAR.SYNC", "LDS", "STS")}
    out.append("")
    out.append("SASS mnemonic counts: " + ", ".join(f"{k} {v}" for k, v in cnt.items()))
    loads = Counter(re.findall(r"LDG[.A-Za-z0-9]*", sass))
    out.append("Load variants: " + ", ".join(f"{k} {v}" for k, v in loads.most_common(8)))
    out.append("")


def main():
    out = ["# Static resource usage of the in-tree binaries (cuobjdump --dump-resource-usage / -sass, sm_100a, nvcc 12.9)", "",
           "The neighbour sweeps and the static step do not spill; the sSSA window kernels keep 16 B of stack (one 64-byte case in "
           "the leap form). `k_force_mv` is capped at 128 registers (`__launch_bounds__(128, 4)`) and the cooperative window kernel "
           "sits at 116-126, i.e. 24-25 % occupancy by design (profiles/README.md). The 256-bit record gathers appear as "
           "`LDG.E.ENL2.256.CONSTANT`. `k_force_mv_tile` is the opt-in shared-memory form of the force sweep (DESIGN.md section 9).", ""]
    usage(codegen.build_core(), "libssb_core.so (model independent kernels)", out)
    usage(codegen.build_model_unit(configs.tank_sdpd(n=12, nt=10, output_every=10)), "model unit: tank (moving SDPD + 1 species sSSA)", out)
    usage(codegen.build_model_unit(configs.cylinder_rdme()), "model unit: cylinder (static RDME+PDE, 2 species, 3 reactions)", out)
    usage(codegen.build_peaks(), "libssb_peaks.so (fp64 peak microbenchmark)", out)
    with open(os.path.join(ROOT, "profiles", "r1_resource_usage.md"), "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
