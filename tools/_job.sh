#!/bin/bash
# GPU job: sharded-head GPU test, bench, ncu --set full (GEMM launches of one step; the other kernels of one step), launch list
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_head_shard_gpu.py -x -q 2>&1 | tail -25) > gpurun_out/s5c_pytest.log
timeout 300 python bench.py --verbose > gpurun_out/s5c_bench.json 2> gpurun_out/s5c_bench.err
timeout 500 ncu --set full --clock-control none -k regex:gemm_kernel -s 24 -c 24 -o /tmp/s5_gemm -f python tools/profile_step.py 2 > gpurun_out/s5c_ncu.log 2>&1
ncu -i /tmp/s5_gemm.ncu-rep --page raw --csv > gpurun_out/s5_gemm_raw.csv 2>> gpurun_out/s5c_ncu.log
timeout 500 ncu --set full --clock-control none -k regex:'^(?!.*gemm_kernel)' -s 30 -c 34 -o /tmp/s5_rest -f python tools/profile_step.py 2 >> gpurun_out/s5c_ncu.log 2>&1
ncu -i /tmp/s5_rest.ncu-rep --page raw --csv > gpurun_out/s5_rest_raw.csv 2>> gpurun_out/s5c_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 70 --csv --log-file gpurun_out/s5_launches.csv python tools/profile_step.py 3 >> gpurun_out/s5c_ncu.log 2>&1
ls -la gpurun_out/ /tmp/*.ncu-rep >> gpurun_out/s5c_ncu.log
du -sh gpurun_out
tail -5 gpurun_out/s5c_pytest.log; cat gpurun_out/s5c_bench.json | cut -c1-400; tail -3 gpurun_out/s5c_ncu.log
