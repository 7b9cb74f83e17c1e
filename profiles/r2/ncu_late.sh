#!/bin/bash
# Late round-2 captures: knn=16 after the segmented insertion network, radius count / fill, box count / fill, and the
# median rule's cooperative kernel. Summaries land in gpurun_out/ (copy to profiles/r2/).
set -x
timeout 400 ncu --set full --clock-control none -k regex:"knn_thread_kernel|radius_thread_kernel" -c 4 \
    -o gpurun_out/r2_late_search -f python profiles/ncu_target.py knn16 radius > gpurun_out/ncu_late_search.log 2>&1
ncu -i gpurun_out/r2_late_search.ncu-rep --page raw --csv > gpurun_out/r2_late_search_raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r2_late_search_raw.csv smsp__inst_executed.sum > gpurun_out/late_search_ncu_summary.txt
timeout 300 ncu --set full --clock-control none -k regex:"median_huge_level" -c 3 \
    -o gpurun_out/r2_late_median -f python profiles/build_median_once.py > gpurun_out/ncu_late_median.log 2>&1
ncu -i gpurun_out/r2_late_median.ncu-rep --page raw --csv > gpurun_out/r2_late_median_raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r2_late_median_raw.csv smsp__inst_executed.sum > gpurun_out/late_median_ncu_summary.txt
rm -f gpurun_out/r2_late_search.ncu-rep gpurun_out/r2_late_median.ncu-rep
tail -5 gpurun_out/ncu_late_median.log
cat gpurun_out/late_search_ncu_summary.txt gpurun_out/late_median_ncu_summary.txt | cut -c1-230
