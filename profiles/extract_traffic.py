#!/usr/bin/env python
"""profiles/extract_traffic.py <workload> <particles> <report.ncu-rep> [...] — read `ncu --set full` captures and write
profiles/ncu_traffic.json: per workload and kernel category, dram__bytes_read.sum + dram__bytes_write.sum per launch and per
particle (what bench.py reports as roofline.traffic), plus the raw-page rows the roofline discussion quotes, appended to
profiles/<tag>_metrics.csv.  Runs where ncu is installed (the dev container reads the .ncu-rep files gpurun brought back)."""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CATEGORY = {"k_force_mv": "force", "k_static_step": "force", "k_finish": "finish", "k_predictor": "predictor", "k_search": "search",
            "k_rdme_windows_coop": "rdme_window", "k_rdme_window": "rdme_window", "k_corrector": "corrector"}
KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "l1tex__throughput",
        "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "lts__throughput", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "smsp__issue_active", "smsp__average_warp", "smsp__warp_issue_stalled",
        "l1tex__data_pipe_lsu_wavefronts", "l1tex__m_xbar2l1tex_read_sectors", "sm__throughput.avg.pct", "smsp__inst_executed.sum")


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    workload, particles = sys.argv[1], int(sys.argv[2])
    out_path = os.path.join(HERE, "ncu_traffic.json")
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for rep in sys.argv[3:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        tag = os.path.splitext(os.path.basename(rep))[0]
        with open(os.path.join(HERE, f"{tag}_metrics.csv"), "w", newline="") as f:
            w = csv.writer(f)
            keep = [i for i, h in enumerate(hdr) if any(k in h for k in KEEP)]
            w.writerow([hdr[i] for i in keep])
            w.writerow([units[i] for i in keep])
            for r in rows[2:]:
                w.writerow([r[i] for i in keep])
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            cat = next((c for k, c in CATEGORY.items() if name.startswith(k) or ("::" + k) in name or k in name), None)
            if cat is None:
                continue
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            table.setdefault(workload, {})[cat] = {"kernel": name.split("(")[0], "bytes_per_launch": rd + wr, "particles": particles,
                                                    "bytes_per_particle": (rd + wr) / particles, "source": os.path.basename(rep)}
    with open(out_path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(json.dumps(table, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
