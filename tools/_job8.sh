#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29713 bench.py --gpus $N --no-cpu-baseline > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "bench rc=$?"
timeout 400 $TR --master-port 29714 bench.py --gpus $N --no-cpu-baseline --head-shard > gpurun_out/n${N}_bench_shard.json 2> gpurun_out/n${N}_bench_shard.err
echo "bench shard rc=$?"
for f in n${N}_bench n${N}_bench_shard; do grep '"metric"' gpurun_out/$f.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$f',d['value'],d['ms_per_step'])" || tail -12 gpurun_out/$f.err | cut -c1-250; done
