#!/bin/bash
# a few counters of both kernels of the headline run, for each SDIMB_RUN_LPS given
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct
for lps in "$@"; do
  SDIMB_RUN_LPS=$lps timeout 600 ncu --metrics $M --clock-control none -k regex:'run_tail|interp_planes' -s 2 -c 2 --csv --log-file gpurun_out/r2_quick_lps$lps.csv \
    python tools/run_case.py 256 3 16384 auto headline 2 > gpurun_out/r2_quick_lps$lps.log 2>&1
done
python - <<'P'
import csv, glob
for f in sorted(glob.glob("gpurun_out/r2_quick_lps*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    for r in rows[1:]:
        print(f.split("/")[-1], r[ki][:60], r[mi], r[vi])
P
