#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for c in 0 1 2 3; do
  XV_AR_CFG=$c timeout 120 $TR --master-port $((29630 + c)) tools/allreduce_bench.py > gpurun_out/n${N}_arbench_cfg$c.json 2> gpurun_out/n${N}_arbench_cfg$c.err
  echo "cfg $c rc=$?"; grep '"world"' gpurun_out/n${N}_arbench_cfg$c.json || grep -E "Error|error" gpurun_out/n${N}_arbench_cfg$c.err | head -5 | cut -c1-300
done
