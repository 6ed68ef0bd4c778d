#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_attention_gpu.py tests/test_head_shard_gpu.py tests/test_kernels_gpu.py -x -q 2>&1 | tail -6) > gpurun_out/s8_pytest.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
tail -n 3 gpurun_out/s8_pytest.log; python -c "import json;d=json.load(open('gpurun_out/s8_bench.json'));print('bench',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])"; tail -3 gpurun_out/s8_bench.err
