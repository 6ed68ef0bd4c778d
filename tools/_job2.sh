#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29621 tools/dist_check_syncbn.py > gpurun_out/n2_check_syncbn.json 2> gpurun_out/n2_check_syncbn.err
echo "check syncbn rc=$?"
timeout 300 $TR --master-port 29622 tools/dist_check_sharded.py --variant symm > gpurun_out/n2_check_symm.json 2> gpurun_out/n2_check_symm.err
echo "check symm rc=$?"
timeout 300 $TR --master-port 29623 bench.py --gpus 2 --no-cpu-baseline --allreduce symm > gpurun_out/n2d_bench_symm.json 2> gpurun_out/n2d_bench_symm.err
echo "bench symm rc=$?"
for f in n2_check_syncbn n2_check_symm; do echo "== $f"; grep '"check"' gpurun_out/$f.json | cut -c1-1500; grep -E "Error|error" gpurun_out/$f.err | head -5 | cut -c1-300; done
for f in n2d_bench_symm; do grep '"metric"' gpurun_out/$f.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$f',d['value'],d['ms_per_step'],d['config']['parallelism'])" || (grep -E "Error|error" gpurun_out/$f.err | head -8 | cut -c1-300); done
