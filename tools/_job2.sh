#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 tools/dist_check_sharded.py > gpurun_out/n2_check_aam.json 2> gpurun_out/n2_check_aam.err
echo "check aam rc=$?"
timeout 400 $TR --master-port 29612 tools/dist_check_sharded.py --loss softmax > gpurun_out/n2_check_softmax.json 2> gpurun_out/n2_check_softmax.err
echo "check softmax rc=$?"
timeout 400 $TR --master-port 29615 tools/dist_check_sharded.py --loss asoftmax --speakers 4300 > gpurun_out/n2_check_asoftmax.json 2> gpurun_out/n2_check_asoftmax.err
echo "check asoftmax rc=$?"
for f in n2_check_aam n2_check_softmax n2_check_asoftmax; do echo "== $f"; grep '"check"' gpurun_out/$f.json | cut -c1-1200; tail -3 gpurun_out/$f.err; done
