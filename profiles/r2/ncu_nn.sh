set -x
for m in 0 1 17; do
PICO_B200_NN=$m ncu --set full --import-source on --clock-control none -k regex:"knn_thread_kernel|nn_kernel" -c 2 -o gpurun_out/r2_nn_mode$m -f python profiles/ncu_target.py knn1 > gpurun_out/ncu_nn_mode$m.log 2>&1
ncu -i gpurun_out/r2_nn_mode$m.ncu-rep --page raw --csv > gpurun_out/r2_nn_mode${m}_raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r2_nn_mode${m}_raw.csv l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum smsp__inst_executed.sum l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum lts__t_sectors_op_read.sum > gpurun_out/r2_nn_mode${m}_summary.txt
cat gpurun_out/r2_nn_mode${m}_summary.txt
done
python profiles/nn_sweep.py NN=17 NN=21 NN=9
