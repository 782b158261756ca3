#!/usr/bin/env bash
# One gpurun call = the evidence a round needs for one workload (B200_PROFILING.md commands):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/capture.sh r2 tank [extra-flags]'
# Writes into gpurun_out/: <tag>_bench_<wl>.json (a clean bench line, measured FIRST and outside any profiler),
# <tag>_launches_<wl>.csv (every launch with its device time; cold-cache, serialised: compare SHARES), <tag>_prof_<wl>_<kernel>.ncu-rep
# (ncu --set full of the dominant kernels) and <tag>_prof_<wl>_metrics.csv (the raw-page rows DESIGN.md / bench.py quote).
# Copy what you want judged into profiles/ afterwards (gpurun_out/ is scratch).
set -u
TAG=${1:-r2}; WL=${2:-tank}; XF=${3:-0}
OUT=gpurun_out; mkdir -p "$OUT"
case "$WL" in
  tank)     KERNELS="k_force_mv k_finish k_rdme_windows_coop k_predictor" ;;
  cylinder) KERNELS="k_static_step k_rdme_windows_coop" ;;
  *)        KERNELS="k_force_mv" ;;
esac
python bench.py --workload "$WL" --extra-flags "$XF" --no-slab > "$OUT/${TAG}_bench_${WL}.json" 2> "$OUT/${TAG}_bench_${WL}.err"
tail -c 600 "$OUT/${TAG}_bench_${WL}.json"
# launch list: two bench steps after one warm-up step are enough to see every kernel of the step
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file "$OUT/${TAG}_launches_${WL}.csv" \
    python bench.py --workload "$WL" --extra-flags "$XF" --steps 1 --warmup 1 --sps 20 --no-cpu --no-slab > /dev/null 2>&1
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k "regex:$K" -s 6 -c 2 -f -o "$OUT/${TAG}_prof_${WL}_${K}" \
      python bench.py --workload "$WL" --extra-flags "$XF" --steps 1 --warmup 1 --sps 12 --no-cpu --no-slab > /dev/null 2>&1
  ncu -i "$OUT/${TAG}_prof_${WL}_${K}.ncu-rep" --page raw --csv 2>/dev/null | python - "$K" >> "$OUT/${TAG}_prof_${WL}_metrics.csv" <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) > 2:
    hdr = rows[0]
    keep = [i for i, h in enumerate(hdr) if any(k in h for k in (
        "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput",
        "l1tex__throughput", "lts__t_sector_hit_rate", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "sm__inst_executed_pipe_fp64", "smsp__pcsamp_warps_issue_stalled", "sm__throughput.avg.pct"))]
    w = csv.writer(sys.stdout)
    w.writerow([hdr[i] for i in keep])
    for r in rows[2:]:
        w.writerow([r[i] for i in keep])
PY
done
ls -la "$OUT" | tail -12
