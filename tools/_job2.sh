#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 tools/dist_check_sharded.py > gpurun_out/n2_check_aam.json 2> gpurun_out/n2_check_aam.err
echo "check aam rc=$?"
timeout 400 $TR --master-port 29612 tools/dist_check_sharded.py --loss softmax > gpurun_out/n2_check_softmax.json 2> gpurun_out/n2_check_softmax.err
echo "check softmax rc=$?"
timeout 400 $TR --master-port 29613 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/n2b_bench.json 2> gpurun_out/n2b_bench.err
timeout 400 $TR --master-port 29614 bench.py --gpus 2 --no-cpu-baseline --dp-overlap --overlap-sms 16 > gpurun_out/n2b_bench_ov16.json 2> gpurun_out/n2b_bench_ov16.err
timeout 400 $TR --master-port 29615 bench.py --gpus 2 --no-cpu-baseline --dp-overlap --overlap-sms 8 > gpurun_out/n2b_bench_ov8.json 2> gpurun_out/n2b_bench_ov8.err
timeout 400 $TR --master-port 29616 bench.py --gpus 2 --no-cpu-baseline --dp-overlap --overlap-sms 32 > gpurun_out/n2b_bench_ov32.json 2> gpurun_out/n2b_bench_ov32.err
for f in n2_check_aam n2_check_softmax; do echo "== $f"; grep '"check"' gpurun_out/$f.json | cut -c1-1500; tail -2 gpurun_out/$f.err; done
for f in n2b_bench n2b_bench_ov16 n2b_bench_ov8 n2b_bench_ov32; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'])" || tail -5 gpurun_out/$f.err; done
