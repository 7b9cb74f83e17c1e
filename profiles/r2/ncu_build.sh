#!/bin/bash
# ncu --set full over the kernels of one device build of the cfg2 tree (7.73M points, sliding midpoint):
# DRAM bytes per kernel against the algorithmic traffic of a level (~100 MB: 31 MB of indices read + written,
# coordinates gathered through them).
set -x
ncu --set full --clock-control none -k regex:"huge_count|huge_scatter|huge_swap|huge_slide|split_level_block|split_level_warp|root_box_kernel" \
    -c 56 -o gpurun_out/r2_build -f python profiles/build_only.py > gpurun_out/ncu_build.log 2>&1
ncu --set full --clock-control none -k regex:"pack_points4|emit_nodes|fat_nodes" \
    -c 4 -o gpurun_out/r2_build_tail -f python profiles/build_only.py >> gpurun_out/ncu_build.log 2>&1
ncu -i gpurun_out/r2_build.ncu-rep --page raw --csv > gpurun_out/r2_build_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_build_tail.ncu-rep --page raw --csv | tail -n +3 >> gpurun_out/r2_build_raw.csv 2>/dev/null
rm -f gpurun_out/r2_build_tail.ncu-rep
python profiles/ncu_pick.py gpurun_out/r2_build_raw.csv > gpurun_out/r2_build_summary.txt
rm -f gpurun_out/r2_build.ncu-rep
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_build_raw.csv")))
hdr, vals = rows[0], rows[2:]
def col(name): return hdr.index(name)
out = open("gpurun_out/r2_build_table.txt", "w")
out.write("%-34s %9s %10s %10s %8s %8s %8s\n" % ("kernel", "us", "dram_rd_MB", "dram_wr_MB", "dram%", "issue%", "lanes"))
for v in vals:
    def g(n, scale=1.0):
        try: return float(v[col(n)]) * scale
        except (ValueError, IndexError): return float("nan")
    unit_rd = rows[1][col("dram__bytes_read.sum")]
    sc = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    out.write("%-34s %9.1f %10.2f %10.2f %8.1f %8.1f %8.1f\n" % (
        v[col("Kernel Name")].split("<")[0].split("::")[-1][:34],
        g("gpu__time_duration.sum", {"us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(rows[1][col("gpu__time_duration.sum")], 1.0)),
        g("dram__bytes_read.sum", sc.get(unit_rd, 1.0)), g("dram__bytes_write.sum", sc.get(rows[1][col("dram__bytes_write.sum")], 1.0)),
        g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        g("smsp__thread_inst_executed_per_inst_executed.ratio")))
out.close()
print(open("gpurun_out/r2_build_table.txt").read())
PY
