set -x
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
cat gpurun_out/bench_e.json | cut -c1-400
python tools/step_kernel_times.py 20 gpurun_out/step_kernels_e.md 2>&1 | grep -v gemm | head -32
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bn_act|stats_pool|head_finish_dw|head_prep_weights|opt_step|pack_input|l2_loss|head_combine|bn_rows' -s 60 -c 30 -f -o gpurun_out/layers_full python tools/profile_step.py 3 > gpurun_out/ncu_layers.log 2>&1
tail -2 gpurun_out/ncu_layers.log
