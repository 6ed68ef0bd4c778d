#!/bin/bash
# GPU job: correctness of the flat BN kernels + cm decode, then timing sweeps
mkdir -p gpurun_out
for m in 1 2; do
  (XV_FLAT=$m timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/s6_pytest_flat$m.log
done
(timeout 600 python -m pytest tests/test_cm_decode_gpu.py tests/test_train_step_gpu.py -x -q 2>&1 | tail -8) > gpurun_out/s6_pytest_misc.log
for m in 0 1 2; do
  XV_FLAT=$m timeout 200 python tools/layers_bench.py --json gpurun_out/s6_layers_flat$m.json > gpurun_out/s6_layers_flat$m.log 2>&1
done
for v in pr2 pr4 cpt8; do
  XV_LIB_PATH=$PWD/tf_kaldi_speaker_b200/libxvector_b200.$v.so timeout 200 python tools/layers_bench.py --json gpurun_out/s6_layers_$v.json > gpurun_out/s6_layers_$v.log 2>&1
done
for m in 0 1 2; do
  XV_FLAT=$m timeout 200 python bench.py --no-cpu-baseline > gpurun_out/s6_bench_flat$m.json 2> gpurun_out/s6_bench_flat$m.err
done
tail -2 gpurun_out/s6_pytest_*.log
for m in 0 1 2; do echo "== flat $m"; grep -E "bn_act|stats_pool" gpurun_out/s6_layers_flat$m.log; python -c "import json;d=json.load(open('gpurun_out/s6_bench_flat$m.json'));print('bench',d['value'],d['ms_per_step'])"; done
for v in pr2 pr4 cpt8; do echo "== $v"; grep -E "stats_pool" gpurun_out/s6_layers_$v.log; done
