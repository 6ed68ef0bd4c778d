#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_extract_gpu.py tests/test_train_step_gpu.py -x -q 2>&1 | tail -4) > gpurun_out/s10_pytest.log
timeout 100 python tools/layers_bench.py --json gpurun_out/s10_layers_mb4.json > gpurun_out/s10_layers_mb4.log 2>&1
XV_LIB_PATH=$PWD/tf_kaldi_speaker_b200/libxvector_b200.mb1.so timeout 100 python tools/layers_bench.py --json gpurun_out/s10_layers_mb1.json > gpurun_out/s10_layers_mb1.log 2>&1
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/s10_bench_mb4.json 2> gpurun_out/s10_bench_mb4.err
XV_LIB_PATH=$PWD/tf_kaldi_speaker_b200/libxvector_b200.mb1.so timeout 200 python bench.py --no-cpu-baseline > gpurun_out/s10_bench_mb1.json 2> gpurun_out/s10_bench_mb1.err
timeout 200 python tools/config_bench.py > gpurun_out/s10_config_bench.log 2>&1
tail -n 2 gpurun_out/s10_pytest.log
for v in mb4 mb1; do echo "== $v"; grep "bwd_apply C" gpurun_out/s10_layers_$v.log; python -c "import json;d=json.load(open('gpurun_out/s10_bench_$v.json'));print('bench',d['value'],d['ms_per_step'])"; done
grep "C5" gpurun_out/s10_config_bench.log | cut -c1-300
