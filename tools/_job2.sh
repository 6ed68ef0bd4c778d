#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29650 tools/dist_check_sharded.py --variant bucket > gpurun_out/n${N}_check_bucket.json 2> gpurun_out/n${N}_check_bucket.err
echo "check rc=$?"; grep '"check"' gpurun_out/n${N}_check_bucket.json | cut -c1-900 || true; grep -E "Error|error" gpurun_out/n${N}_check_bucket.err | head -5 | cut -c1-300
timeout 200 $TR --master-port 29651 bench.py --gpus $N --no-cpu-baseline --bucket-overlap > gpurun_out/n${N}h_bench_bucket.json 2> gpurun_out/n${N}h_bench_bucket.err
echo "bench rc=$?"
grep '"metric"' gpurun_out/n${N}h_bench_bucket.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('bench bucket',d['value'],d['ms_per_step'])" || (grep -E "Error|error" gpurun_out/n${N}h_bench_bucket.err | head -8 | cut -c1-300)
timeout 200 $TR --master-port 29652 bench.py --gpus $N --no-cpu-baseline > gpurun_out/n${N}h_bench_flat.json 2> gpurun_out/n${N}h_bench_flat.err
grep '"metric"' gpurun_out/n${N}h_bench_flat.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('bench flat',d['value'],d['ms_per_step'])" || (grep -E "Error|error" gpurun_out/n${N}h_bench_flat.err | head -8 | cut -c1-300)
