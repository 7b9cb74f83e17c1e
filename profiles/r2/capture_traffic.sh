#!/bin/bash
# One ncu --set full capture of the headline traversal kernel and the isolated leaf-scan kernels at the bench.py
# default workload; writes gpurun_out/traffic.json (copy to profiles/traffic.json) with the hash of the kernel
# sources it is valid for, plus the metric summary and the hottest SASS lines.
set -x
ncu --set full --import-source on --clock-control none -k regex:"knn_thread_kernel|leaf_scan_kernel|first_leaf_kernel" -c 6 \
    -o gpurun_out/r2_final_kernels -f python profiles/ncu_target.py knn1 leaf > gpurun_out/ncu_final.log 2>&1
ncu -i gpurun_out/r2_final_kernels.ncu-rep --page raw --csv > gpurun_out/r2_final_kernels_raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r2_final_kernels_raw.csv smsp__inst_executed.sum > gpurun_out/r2_final_kernels_summary.txt
ncu -i gpurun_out/r2_final_kernels.ncu-rep --page source --csv -k regex:knn_thread_kernel > gpurun_out/r2_final_source.csv 2>/dev/null
python profiles/ncu_src.py gpurun_out/r2_final_source.csv > gpurun_out/r2_final_hot_sass.txt 2>/dev/null
rm -f gpurun_out/r2_final_kernels.ncu-rep gpurun_out/r2_final_source.csv
python - <<'PY'
import csv, json, sys
sys.path.insert(0, ".")
from bench import kernel_sources_sha16
rows = list(csv.reader(open("gpurun_out/r2_final_kernels_raw.csv")))
hdr, units, vals = rows[0], rows[1], rows[2:]
sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def dram(v):
    r, w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    return float(v[r]) * sc[units[r]] + float(v[w]) * sc[units[w]]
knn = [dram(v) for v in vals if "knn_thread_kernel" in v[hdr.index("Kernel Name")]]
leaf = [dram(v) for v in vals if "leaf_scan_kernel" in v[hdr.index("Kernel Name")]]
out = {"knn1_dram_bytes_per_launch": int(sum(knn) / len(knn)), "leaf_scan_dram_bytes_per_launch": int(sum(leaf) / len(leaf)),
       "sources_sha16": kernel_sources_sha16(),
       "source": "profiles/r2/final_kernels_ncu_raw.csv: one ncu --set full capture at the bench.py default workload "
                 "(7,200,863 queries per launch), dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of %d / %d "
                 "launches" % (len(knn), len(leaf))}
json.dump(out, open("gpurun_out/traffic.json", "w"), indent=1)
print(out)
PY
cat gpurun_out/r2_final_kernels_summary.txt | head -40
