mkdir -p gpurun_out/r02h
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r02h/bench_2gpu.json 2> gpurun_out/r02h/bench_2gpu.err; echo bench2 rc=$?
python -c "import json; j=json.load(open('gpurun_out/r02h/bench_2gpu.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['config']['parallelism'])"
timeout 300 $TR --master-port 29602 tools/dist_check_sharded.py --variant multimem > gpurun_out/r02h/n2_allreduce_check.json 2> gpurun_out/r02h/n2_allreduce_check.err; echo check rc=$?; tail -c 600 gpurun_out/r02h/n2_allreduce_check.json
timeout 300 $TR --master-port 29603 tools/dist_check_syncbn.py > gpurun_out/r02h/n2_syncbn_check.json 2> gpurun_out/r02h/n2_syncbn.err; echo syncbn rc=$?; tail -c 500 gpurun_out/r02h/n2_syncbn_check.json
timeout 300 $TR --master-port 29604 tools/dist_check_syncbn.py --attention > gpurun_out/r02h/n2_syncbn_attention_check.json 2> gpurun_out/r02h/n2_syncbn_att.err; echo syncbn-att rc=$?; tail -c 700 gpurun_out/r02h/n2_syncbn_attention_check.json; tail -3 gpurun_out/r02h/n2_syncbn_att.err
timeout 300 $TR --master-port 29605 tools/dist_check_sharded.py --variant shard > gpurun_out/r02h/n2_shard_check.json 2> gpurun_out/r02h/n2_shard.err; echo shard rc=$?; tail -c 400 gpurun_out/r02h/n2_shard_check.json
timeout 200 $TR --master-port 29606 bench.py --workload extract --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02h/bench_extract_2gpu.json 2>/dev/null; echo ex2 rc=$?; python -c "import json; j=json.load(open('gpurun_out/r02h/bench_extract_2gpu.json')); print(j['value'], j['e2e']['value'])"
