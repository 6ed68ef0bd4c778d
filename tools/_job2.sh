#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29640 tools/allreduce_bench.py > gpurun_out/n${N}_arbench.json 2> gpurun_out/n${N}_arbench.err
echo "arbench rc=$?"; grep '"world"' gpurun_out/n${N}_arbench.json || grep -E "Error|error" gpurun_out/n${N}_arbench.err | head -5 | cut -c1-300
timeout 240 $TR --master-port 29641 bench.py --gpus $N --no-cpu-baseline --allreduce multimem > gpurun_out/n${N}f_bench_mm.json 2> gpurun_out/n${N}f_bench_mm.err
echo "bench rc=$?"
grep '"metric"' gpurun_out/n${N}f_bench_mm.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('bench',d['value'],d['ms_per_step'],d['config']['parallelism'])" || (grep -E "Error|error" gpurun_out/n${N}f_bench_mm.err | head -8 | cut -c1-300)
