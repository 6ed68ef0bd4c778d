#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29751 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/n${N}g_bench.json 2> gpurun_out/n${N}g_bench.err
echo "bench rc=$?"
grep '"metric"' gpurun_out/n${N}g_bench.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('bench',d['value'],d['ms_per_step'],d['e2e']['value'],d['config']['parallelism'],d['clocks'])" || (grep -E "Error|error" gpurun_out/n${N}g_bench.err | head -8 | cut -c1-300)
timeout 200 $TR --master-port 29752 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/n${N}g_ref.json 2> gpurun_out/n${N}g_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/n${N}g_ref.json
