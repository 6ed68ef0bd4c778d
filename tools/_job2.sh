#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29622 tools/dist_check_sharded.py --variant multimem > gpurun_out/n2_check_multimem.json 2> gpurun_out/n2_check_multimem.err
echo "check multimem rc=$?"
timeout 240 $TR --master-port 29623 bench.py --gpus 2 --no-cpu-baseline --allreduce multimem > gpurun_out/n2e_bench_multimem.json 2> gpurun_out/n2e_bench_multimem.err
echo "bench multimem rc=$?"
nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv
for f in n2_check_multimem; do echo "== $f"; grep '"check"' gpurun_out/$f.json | cut -c1-1500; grep -E "Error|error" gpurun_out/$f.err | head -5 | cut -c1-300; done
for f in n2e_bench_multimem; do grep '"metric"' gpurun_out/$f.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$f',d['value'],d['ms_per_step'],d['config']['parallelism'])" || (grep -E "Error|error" gpurun_out/$f.err | head -8 | cut -c1-300); done
