"""Instruction histogram of one kernel of a model unit, from `cuobjdump -sass` (run here, no GPU needed).
usage: python profiles/sass_histogram.py <unit.so> <kernel substring>   ->  mnemonic counts, whole kernel and hottest loop body"""
import collections
import re
import subprocess
import sys


def kernel_sass(so, name):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    blocks = out.split("Function : ")
    for b in blocks[1:]:
        if name in b.split("\n", 1)[0]:
            return b
    raise SystemExit(f"no kernel matching {name}")


def main():
    so, name = sys.argv[1], sys.argv[2]
    body = kernel_sass(so, name)
    ins = []
    for ln in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    print(f"{name}: {len(ins)} SASS instructions")
    # back edges: BRA to a lower address; the hottest loop = the innermost back edge spanning LDG.E.ENL2.256 / LDG..256 loads
    loops = []
    for addr, op, rest in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) < addr:
                loops.append((int(m.group(1), 16), addr))
    def hist(lo, hi):
        h = collections.Counter()
        for addr, op, _ in ins:
            if lo <= addr <= hi:
                h[op] += 1
        return h
    whole = hist(0, 1 << 60)
    print("whole kernel, top mnemonics:", ", ".join(f"{k} {v}" for k, v in whole.most_common(14)))
    best = None
    for lo, hi in loops:
        h = hist(lo, hi)
        wide = sum(v for k, v in h.items() if k.startswith("LDG") and "256" in k)
        if wide and (best is None or (hi - lo) < (best[1] - best[0])):
            best = (lo, hi, h, wide)
    if best:
        lo, hi, h, wide = best
        n = sum(h.values())
        fp64 = sum(v for k, v in h.items() if k.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RCP64H", "MUFU.RSQ64H")))
        print(f"innermost loop with 256-bit loads: {lo:#x}..{hi:#x}, {n} instructions, {wide} x 256-bit LDG, "
              f"{sum(v for k, v in h.items() if k.startswith('LDG'))} LDG in all, {fp64} fp64-pipe instructions")
        print("  ", ", ".join(f"{k} {v}" for k, v in h.most_common(20)))


if __name__ == "__main__":
    main()
