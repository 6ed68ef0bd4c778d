#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29621 tools/dist_check_syncbn.py > gpurun_out/n2_check_syncbn.json 2> gpurun_out/n2_check_syncbn.err
echo "check syncbn rc=$?"
timeout 400 $TR --master-port 29622 tools/dist_check_sharded.py --variant bf16 > gpurun_out/n2_check_bf16.json 2> gpurun_out/n2_check_bf16.err
echo "check bf16 rc=$?"
timeout 400 $TR --master-port 29623 bench.py --gpus 2 --no-cpu-baseline --grad-dtype bf16 > gpurun_out/n2c_bench_bf16.json 2> gpurun_out/n2c_bench_bf16.err
for f in n2_check_syncbn n2_check_bf16; do echo "== $f"; grep '"check"' gpurun_out/$f.json | cut -c1-1500; tail -4 gpurun_out/$f.err | cut -c1-300; done
for f in n2c_bench_bf16; do grep '"metric"' gpurun_out/$f.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$f',d['value'],d['ms_per_step'])" || tail -12 gpurun_out/$f.err | cut -c1-250; done
